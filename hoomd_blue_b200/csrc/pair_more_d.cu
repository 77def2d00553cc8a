// pair_more_d.cu — explicit instantiations: ExpandedMie, OPP, Fourier (five and six packed parameters).
#include "pair_kernel.cuh"
#include "evaluators_more.cuh"
namespace hb2
    {
hb2_status pair_forces_expanded_mie(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairExpandedMie>(m, a, b, s);
    }
hb2_status pair_forces_opp(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairOPP>(m, a, b, s);
    }
hb2_status pair_forces_fourier(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairFourier>(m, a, b, s);
    }
    } // namespace hb2
