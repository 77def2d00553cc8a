// pair_more_c.cu — explicit instantiations: Buckingham, ExpandedLJ, ExpandedGaussian. (One translation unit per two or three evaluators: every
// evaluator instantiates the row and tile kernels in all their variants, minutes of nvcc time each.)
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
hb2_status pair_forces_expanded_gaussian(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairExpandedGaussian>(m, a, b, s);
    }
    } // namespace hb2
