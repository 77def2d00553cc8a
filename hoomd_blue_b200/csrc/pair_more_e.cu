// pair_more_e.cu — explicit instantiations: OPP (six packed parameters). (One translation unit per two or three evaluators: every
// evaluator instantiates the row and tile kernels in all their variants, minutes of nvcc time each.)
#include "pair_kernel.cuh"
#include "evaluators_more.cuh"
namespace hb2
    {
hb2_status pair_forces_opp(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairOPP>(m, a, b, s);
    }
    } // namespace hb2
