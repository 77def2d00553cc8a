// pair_more_g.cu — explicit instantiations: LJGauss, TWF, Table. (One translation unit per two or three evaluators: every
// evaluator instantiates the row and tile kernels in all their variants, minutes of nvcc time each.)
#include "pair_kernel.cuh"
#include "evaluators_more.cuh"
namespace hb2
    {
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
