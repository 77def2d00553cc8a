// pair_lj.cu — explicit instantiation of the pair-force pass for EvaluatorPairLJ (the counterpart of
// the reference's generated PotentialPairGPUKernel.cu per evaluator, hoomd/md/CMakeLists.txt:505-554).
#include "pair_kernel.cuh"
namespace hb2
    {
hb2_status pair_forces_lj(hb2_shift_mode mode, const PairArgs& a, const hb2_box& box,
                          cudaStream_t stream)
    {
    return pair_dispatch<EvaluatorPairLJ>(mode, a, box, stream);
    }
    } // namespace hb2
