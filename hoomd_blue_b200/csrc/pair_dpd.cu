// pair_dpd.cu — explicit instantiation for the DPD thermostat evaluator
// (hoomd/md/CMakeLists.txt:482-503 generates the reference's DPD kernels the same way) and for its
// conservative-only use through the plain PotentialPair (hoomd.md.pair.DPDConservative).
#include "pair_kernel.cuh"
namespace hb2
    {
class EvaluatorPairDPDConservative : public EvaluatorPairDPDThermoDPD
    {
    public:
    static constexpr bool needs_thermo = false;
    HB2_DEVICE EvaluatorPairDPDConservative(Scalar _rsq, Scalar _rcutsq, const param_type& _params)
        : EvaluatorPairDPDThermoDPD(_rsq, _rcutsq, _params)
        {
        }
    };

hb2_status pair_forces_dpd(hb2_shift_mode mode, const PairArgs& a, const hb2_box& box,
                           cudaStream_t stream)
    {
    return pair_dispatch<EvaluatorPairDPDThermoDPD>(mode, a, box, stream);
    }
hb2_status pair_forces_dpdlj(hb2_shift_mode mode, const PairArgs& a, const hb2_box& box, cudaStream_t stream)
    {
    return pair_dispatch<EvaluatorPairDPDThermoLJ>(mode, a, box, stream);
    }
hb2_status pair_forces_dpd_conservative(hb2_shift_mode mode, const PairArgs& a, const hb2_box& box,
                                        cudaStream_t stream)
    {
    return pair_dispatch<EvaluatorPairDPDConservative>(mode, a, box, stream);
    }
    } // namespace hb2
