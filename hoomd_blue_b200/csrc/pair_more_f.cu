// pair_more_f.cu — explicit instantiations: Fourier (six packed parameters). (One translation unit per two or three evaluators: every
// evaluator instantiates the row and tile kernels in all their variants, minutes of nvcc time each.)
#include "pair_kernel.cuh"
#include "evaluators_more.cuh"
namespace hb2
    {
hb2_status pair_forces_fourier(hb2_shift_mode m, const PairArgs& a, const hb2_box& b, cudaStream_t s)
    {
    return pair_dispatch<EvaluatorPairFourier>(m, a, b, s);
    }
    } // namespace hb2
