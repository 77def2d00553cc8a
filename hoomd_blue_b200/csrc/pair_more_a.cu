// pair_more_a.cu — explicit instantiations: Gauss, Yukawa, Mie, Morse (one line per evaluator, the
// pair_plugin pattern: example_plugins/pair_plugin/PotentialPairExampleGPUKernel.cu:13-15).
#include "pair_kernel.cuh"
#include "evaluators_more.cuh"
namespace hb2
    {
hb2_status pair_forces_gauss(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairGauss>(m, a, b, s);
    }
hb2_status pair_forces_yukawa(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairYukawa>(m, a, b, s);
    }
hb2_status pair_forces_mie(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairMie>(m, a, b, s);
    }
hb2_status pair_forces_morse(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairMorse>(m, a, b, s);
    }
    } // namespace hb2
