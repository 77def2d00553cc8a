// pair_example.cu — the pair_plugin pattern: a user evaluator header plus ONE .cu that instantiates
// the kernel template for it (example_plugins/pair_plugin/PotentialPairExampleGPUKernel.cu:13-15).
#include "pair_kernel.cuh"
namespace hb2
    {
hb2_status pair_forces_example(hb2_shift_mode mode, const PairArgs& a, const hb2_box& box,
                               cudaStream_t stream)
    {
    return pair_dispatch<EvaluatorPairExample>(mode, a, box, stream);
    }
    } // namespace hb2
