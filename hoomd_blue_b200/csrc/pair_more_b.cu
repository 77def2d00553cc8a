// pair_more_b.cu — explicit instantiations: LJ1208, LJ0804, ForceShiftedLJ.
#include "pair_kernel.cuh"
#include "evaluators_more.cuh"
namespace hb2
    {
hb2_status pair_forces_lj1208(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairLJ1208>(m, a, b, s);
    }
hb2_status pair_forces_lj0804(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairLJ0804>(m, a, b, s);
    }
hb2_status pair_forces_force_shifted_lj(hb2_shift_mode m, const PairArgs& a, const hb2_box& b,
                                        cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairForceShiftedLJ>(m, a, b, s);
    }
    } // namespace hb2
