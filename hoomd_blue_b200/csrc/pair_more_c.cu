// pair_more_c.cu — explicit instantiations: Buckingham, ExpandedLJ, ExpandedGaussian, LJGauss, TWF.
#include "pair_kernel.cuh"
#include "evaluators_more.cuh"
namespace hb2
    {
hb2_status pair_forces_buckingham(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairBuckingham>(m, a, b, s);
    }
hb2_status pair_forces_expanded_lj(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairExpandedLJ>(m, a, b, s);
    }
hb2_status pair_forces_expanded_gaussian(hb2_shift_mode m, const PairArgs& a, const hb2_box& b,
                                         cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairExpandedGaussian>(m, a, b, s);
    }
hb2_status pair_forces_lj_gauss(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairLJGauss>(m, a, b, s);
    }
hb2_status pair_forces_twf(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairTWF>(m, a, b, s);
    }
hb2_status pair_forces_table(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairTable>(m, a, b, s);
    }
    } // namespace hb2
