// evaluators.cuh — pair evaluators with the interface of the reference's pair_plugin template
// (example_plugins/pair_plugin/EvaluatorPairExample.h:33-160; canonical doc
// hoomd/md/EvaluatorPairLJ.h:34-80): nested param_type, ctor (rsq, rcutsq, params),
// evalForceAndEnergy(force_divr, pair_eng, energy_shift), static getName(). A user evaluator written
// for the reference compiles against pair_kernel.cuh by adding one .cu that instantiates
// hb2::pair_dispatch<Evaluator> (see pair_example.cu and INTEGRATION.md).
//
// Arithmetic is fp32 (Scalar = float, the HOOMD_LONGREAL_SIZE=32 layout). Host-side `pack` is the
// param_type(pybind11::dict) constructor restated over a plain array of user parameters.
//
// The built-in evaluators are written branch-free (compute, then select on the cut-off test): in a
// liquid nearly every warp has a lane inside the cut-off, so a branch only adds divergence
// bookkeeping; the kernel discards the values of a pair that returns false.
#pragma once
#include "common.cuh"
#include "rng.cuh"

namespace hb2
    {
typedef float Scalar;

// 1/x to ~0.5 ulp: MUFU.RCP (1 ulp) plus one Newton step in FMA form.
HB2_DEVICE float rcp_refined(float x)
    {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    const float e = fmaf(-x, y, 1.0f);
    return fmaf(y, e, y);
    }
// 1/sqrt(x) to ~0.5 ulp: MUFU.RSQ plus one Newton step.
HB2_DEVICE float rsqrt_refined(float x)
    {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    const float e = fmaf(-x * y, y, 1.0f);
    return fmaf(0.5f * y, e, y);
    }

// ---------------------------------------------------------------------------------------------
// Lennard-Jones (hoomd/md/EvaluatorPairLJ.h:95-270)
// ---------------------------------------------------------------------------------------------
class EvaluatorPairLJ
    {
    public:
    struct param_type
        {
        Scalar lj1; // 4 eps sigma^12
        Scalar lj2; // 4 eps sigma^6
        };
    static constexpr bool needs_thermo = false;
    static constexpr bool tune_tpp = true;
    static constexpr unsigned n_user_params = 2; // {epsilon, sigma}

    // hoomd/md/EvaluatorPairLJ.h:119-140: sigma_6 = sigma^6, epsilon_x_4 = 4 eps (fp64 on the host)
    static void pack(const double* user, param_type& p)
        {
        const double eps = user[0], sigma = user[1];
        const double s6 = sigma * sigma * sigma * sigma * sigma * sigma;
        p.lj1 = (Scalar)(4.0 * eps * s6 * s6);
        p.lj2 = (Scalar)(4.0 * eps * s6);
        }

    HB2_DEVICE EvaluatorPairLJ(Scalar _rsq, Scalar _rcutsq, const param_type& _params)
        : rsq(_rsq), rcutsq(_rcutsq), lj1(_params.lj1), lj2(_params.lj2)
        {
        }

    // hoomd/md/EvaluatorPairLJ.h:162-211
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        const bool within = rsq < rcutsq && lj1 != 0;
        const Scalar r2inv = rcp_refined(rsq);
        const Scalar r6inv = r2inv * r2inv * r2inv;
        force_divr = r2inv * r6inv * (Scalar(12.0) * lj1 * r6inv - Scalar(6.0) * lj2);
        pair_eng = r6inv * (lj1 * r6inv - lj2);
        if (energy_shift)
            {
            const Scalar rcut2inv = rcp_refined(rcutsq);
            const Scalar rcut6inv = rcut2inv * rcut2inv * rcut2inv;
            pair_eng -= rcut6inv * (lj1 * rcut6inv - lj2);
            }
        return within;
        }
    static const char* getName()
        {
        return "lj";
        }

    // Two pairs per instruction on Blackwell's packed fp32 pipe (FFMA2 / FMUL2 / FADD2, sm_100):
    // the same arithmetic as evalForceAndEnergy, component-wise on (pair 0, pair 1). Used by the pair
    // kernel for the one-type, non-xplor instantiations (uniform parameters).
    static constexpr bool has_packed = true;
    HB2_DEVICE static void evalPacked(const float2 rsq, const Scalar rcutsq, const param_type& p,
                                      const bool energy_shift, const bool want_energy,
                                      float2& force_divr, float2& pair_eng, bool& in0, bool& in1)
        {
        const bool on = p.lj1 != 0;
        in0 = rsq.x < rcutsq && on;
        in1 = rsq.y < rcutsq && on;
        float2 y;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.x) : "f"(rsq.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.y) : "f"(rsq.y));
        const float2 one = make_float2(1.0f, 1.0f);
        const float2 e = __ffma2_rn(make_float2(-rsq.x, -rsq.y), y, one);
        const float2 r2inv = __ffma2_rn(y, e, y);
        const float2 r6inv = __fmul2_rn(__fmul2_rn(r2inv, r2inv), r2inv);
        const Scalar a12 = Scalar(12.0) * p.lj1, b6 = Scalar(-6.0) * p.lj2;
        const float2 t = __ffma2_rn(make_float2(a12, a12), r6inv, make_float2(b6, b6));
        force_divr = __fmul2_rn(__fmul2_rn(r2inv, r6inv), t);
        if (want_energy)
            {
            const float2 u = __ffma2_rn(make_float2(p.lj1, p.lj1), r6inv, make_float2(-p.lj2, -p.lj2));
            pair_eng = __fmul2_rn(r6inv, u);
            if (energy_shift)
                {
                const Scalar rcut2inv = rcp_refined(rcutsq);
                const Scalar rcut6inv = rcut2inv * rcut2inv * rcut2inv;
                const Scalar sft = rcut6inv * (p.lj1 * rcut6inv - p.lj2);
                pair_eng = __fadd2_rn(pair_eng, make_float2(-sft, -sft));
                }
            }
        }

    protected:
    Scalar rsq, rcutsq, lj1, lj2;
    };

// ---------------------------------------------------------------------------------------------
// The pair_plugin example potential V = k/2 (sigma - r)^2
// (example_plugins/pair_plugin/EvaluatorPairExample.h:94-127) — kept in the reference's branchy form
// to show that an unmodified user evaluator works.
// ---------------------------------------------------------------------------------------------
class EvaluatorPairExample
    {
    public:
    static constexpr bool has_packed = false;
    static constexpr bool tune_tpp = true;
    struct param_type
        {
        Scalar k;
        Scalar sigma;
        };
    static constexpr bool needs_thermo = false;
    static constexpr unsigned n_user_params = 2; // {k, sigma}
    static void pack(const double* user, param_type& p)
        {
        p.k = (Scalar)user[0];
        p.sigma = (Scalar)user[1];
        }
    HB2_DEVICE EvaluatorPairExample(Scalar _rsq, Scalar _rcutsq, const param_type& _params)
        : rsq(_rsq), rcutsq(_rcutsq), k(_params.k), sigma(_params.sigma)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq)
            {
            const Scalar r = sqrtf(rsq);
            const Scalar rinv = Scalar(1.0) / r;
            const Scalar overlap = sigma - r;
            force_divr = k * overlap * rinv;
            pair_eng = Scalar(0.5) * k * overlap * overlap;
            if (energy_shift)
                {
                const Scalar rcut = sqrtf(rcutsq);
                const Scalar cut_overlap = sigma - rcut;
                pair_eng -= Scalar(0.5) * k * cut_overlap * cut_overlap;
                }
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "example_pair";
        }

    protected:
    Scalar rsq, rcutsq, k, sigma;
    };

// ---------------------------------------------------------------------------------------------
// DPD (hoomd/md/EvaluatorPairDPDThermoDPD.h:144-277). evalForceAndEnergy is the conservative part
// (:186-203); evalForceEnergyThermo adds drag and the per-pair Philox random force (:222-277).
// ---------------------------------------------------------------------------------------------
class EvaluatorPairDPDThermoDPD
    {
    public:
    static constexpr bool has_packed = false;
    static constexpr bool tune_tpp = true;
    struct param_type
        {
        Scalar A;
        Scalar gamma;
        };
    static constexpr bool needs_thermo = true;
    static constexpr unsigned n_user_params = 2; // {A, gamma}
    static void pack(const double* user, param_type& p)
        {
        p.A = (Scalar)user[0];
        p.gamma = (Scalar)user[1];
        }
    HB2_DEVICE EvaluatorPairDPDThermoDPD(Scalar _rsq, Scalar _rcutsq, const param_type& _params)
        : rsq(_rsq), rcutsq(_rcutsq), a(_params.A), gamma(_params.gamma)
        {
        }
    HB2_DEVICE void set_seed_ij_timestep(uint16_t seed, uint32_t i, uint32_t j, uint64_t timestep)
        {
        m_seed = seed;
        m_i = i;
        m_j = j;
        m_timestep = timestep;
        }
    HB2_DEVICE void setDeltaT(Scalar dt)
        {
        m_deltaT = dt;
        }
    HB2_DEVICE void setRDotV(Scalar dot)
        {
        m_dot = dot;
        }
    HB2_DEVICE void setT(Scalar Temp)
        {
        m_T = Temp;
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        const bool within = rsq < rcutsq;
        const Scalar rinv = rsqrt_refined(rsq);
        const Scalar r = rsq * rinv;
        const Scalar rcutinv = rsqrt_refined(rcutsq);
        const Scalar rcut = rcutsq * rcutinv;
        force_divr = a * (rinv - rcutinv);
        pair_eng = a * (rcut - r) - Scalar(0.5) * a * rcutinv * (rcutsq - rsq);
        return within;
        }
    HB2_DEVICE bool evalForceEnergyThermo(Scalar& force_divr, Scalar& force_divr_cons,
                                          Scalar& pair_eng, bool energy_shift)
        {
        // Unlike the LJ case most listed DPD neighbours are OUTSIDE the cut-off (r_cut = 1,
        // r_list = 1.4: 64 % of the list) and the Philox draw is 40+ integer instructions, so this
        // one keeps the branch.
        if (rsq < rcutsq)
            {
            const Scalar rinv = rsqrt_refined(rsq);
            const Scalar r = rsq * rinv;
            const Scalar rcutinv = rsqrt_refined(rcutsq);
            const Scalar rcut = rcutsq * rcutinv;
            const Scalar w = rinv - rcutinv;
            // uniform in (-1,1) from the pair's Philox draw; the top 32 bits carry every bit an
            // fp32 alpha can hold (the fp64 reference uses 64)
            uint32_t w0, w1;
            dpd_pair_bits(m_seed, m_i, m_j, m_timestep, w0, w1);
            const Scalar alpha = fmaf((Scalar)w0, 4.656612873077393e-10f, -1.0f);
            force_divr = a * w;
            force_divr_cons = force_divr;
            force_divr -= gamma * m_dot * w * w;
            force_divr += rsqrt_refined(m_deltaT / (m_T * gamma * Scalar(6.0))) * w * alpha;
            pair_eng = a * (rcut - r) - Scalar(0.5) * a * rcutinv * (rcutsq - rsq);
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "dpd";
        }

    protected:
    Scalar rsq, rcutsq, a, gamma;
    uint16_t m_seed;
    uint32_t m_i, m_j;
    uint64_t m_timestep;
    Scalar m_deltaT, m_dot, m_T;
    };

// DPDLJ: LJ conservative force + the DPD drag and random forces
// (hoomd/md/EvaluatorPairDPDThermoLJ.h:66-76 parameters, :204-296 evaluation)
class EvaluatorPairDPDThermoLJ
    {
    public:
    static constexpr bool has_packed = false;
    static constexpr bool tune_tpp = false;
    struct param_type
        {
        Scalar lj1, lj2, gamma;
        };
    static constexpr bool needs_thermo = true;
    static constexpr unsigned n_user_params = 3; // {epsilon, sigma, gamma}
    static void pack(const double* user, param_type& p)
        {
        const Scalar sigma = (Scalar)user[1];
        const Scalar sigma_6 = sigma * sigma * sigma * sigma * sigma * sigma;
        const Scalar epsilon_x_4 = Scalar(4.0) * (Scalar)user[0];
        p.lj1 = epsilon_x_4 * sigma_6 * sigma_6;
        p.lj2 = epsilon_x_4 * sigma_6;
        p.gamma = (Scalar)user[2];
        }
    HB2_DEVICE EvaluatorPairDPDThermoLJ(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), lj1(p.lj1), lj2(p.lj2), gamma(p.gamma)
        {
        }
    HB2_DEVICE void set_seed_ij_timestep(uint16_t seed, uint32_t i, uint32_t j, uint64_t timestep)
        {
        m_seed = seed;
        m_i = i;
        m_j = j;
        m_timestep = timestep;
        }
    HB2_DEVICE void setDeltaT(Scalar dt)
        {
        m_deltaT = dt;
        }
    HB2_DEVICE void setRDotV(Scalar dot)
        {
        m_dot = dot;
        }
    HB2_DEVICE void setT(Scalar Temp)
        {
        m_T = Temp;
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq && lj1 != 0)
            {
            const Scalar r2inv = Scalar(1.0) / rsq;
            const Scalar r6inv = r2inv * r2inv * r2inv;
            force_divr = r2inv * r6inv * (Scalar(12.0) * lj1 * r6inv - Scalar(6.0) * lj2);
            pair_eng = r6inv * (lj1 * r6inv - lj2);
            if (energy_shift)
                {
                const Scalar rcut2inv = Scalar(1.0) / rcutsq;
                const Scalar rcut6inv = rcut2inv * rcut2inv * rcut2inv;
                pair_eng -= rcut6inv * (lj1 * rcut6inv - lj2);
                }
            return true;
            }
        return false;
        }
    HB2_DEVICE bool evalForceEnergyThermo(Scalar& force_divr, Scalar& force_divr_cons,
                                          Scalar& pair_eng, bool energy_shift)
        {
        if (!evalForceAndEnergy(force_divr, pair_eng, energy_shift))
            return false;
        force_divr_cons = force_divr;
        const Scalar w = rsqrt_refined(rsq) - rsqrt_refined(rcutsq);
        uint32_t w0, w1;
        dpd_pair_bits(m_seed, m_i, m_j, m_timestep, w0, w1);
        const Scalar alpha = fmaf((Scalar)w0, 4.656612873077393e-10f, -1.0f);
        force_divr -= gamma * m_dot * w * w;
        force_divr += rsqrt_refined(m_deltaT / (m_T * gamma * Scalar(6.0))) * w * alpha;
        return true;
        }
    static const char* getName()
        {
        return "dpdlj";
        }

    protected:
    Scalar rsq, rcutsq, lj1, lj2, gamma;
    uint16_t m_seed;
    uint32_t m_i, m_j;
    uint64_t m_timestep;
    Scalar m_deltaT, m_dot, m_T;
    };
    } // namespace hb2
