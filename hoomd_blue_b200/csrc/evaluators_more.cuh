// evaluators_more.cuh — further pair potentials through the same evaluator template
// (the reference generates one kernel per evaluator the same way,
// hoomd/md/CMakeLists.txt:505-554). Each class restates the arithmetic of the cited reference header in
// fp32; `pack` is its param_type(pybind11::dict) constructor over a plain array of user parameters.
// They are instantiated for the default threads-per-particle only (tune_tpp = false).
#pragma once
#include "evaluators.cuh"

namespace hb2
    {
#define HB2_EVAL_COMMON(NUSER)                          \
    static constexpr bool needs_thermo = false;         \
    static constexpr bool has_packed = false;           \
    static constexpr bool tune_tpp = false;             \
    static constexpr unsigned n_user_params = NUSER;

// V = eps exp(-r^2 / (2 sigma^2))   (hoomd/md/EvaluatorPairGauss.h:126-147)
class EvaluatorPairGauss
    {
    public:
    struct param_type
        {
        Scalar epsilon, sigma;
        };
    HB2_EVAL_COMMON(2) // {epsilon, sigma}
    static void pack(const double* user, param_type& p)
        {
        p.epsilon = (Scalar)user[0];
        p.sigma = (Scalar)user[1];
        }
    HB2_DEVICE EvaluatorPairGauss(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), epsilon(p.epsilon), sigma(p.sigma)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq)
            {
            const Scalar sigma_sq = sigma * sigma;
            const Scalar exp_val = __expf(Scalar(-0.5) * rsq / sigma_sq);
            force_divr = epsilon / sigma_sq * exp_val;
            pair_eng = epsilon * exp_val;
            if (energy_shift)
                pair_eng -= epsilon * __expf(Scalar(-0.5) * rcutsq / sigma_sq);
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "gauss";
        }

    protected:
    Scalar rsq, rcutsq, epsilon, sigma;
    };

// V = eps exp(-kappa r) / r   (hoomd/md/EvaluatorPairYukawa.h:128-154)
class EvaluatorPairYukawa
    {
    public:
    struct param_type
        {
        Scalar epsilon, kappa;
        };
    HB2_EVAL_COMMON(2) // {epsilon, kappa}
    static void pack(const double* user, param_type& p)
        {
        p.epsilon = (Scalar)user[0];
        p.kappa = (Scalar)user[1];
        }
    HB2_DEVICE EvaluatorPairYukawa(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), epsilon(p.epsilon), kappa(p.kappa)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq && epsilon != 0)
            {
            const Scalar rinv = rsqrt_refined(rsq);
            const Scalar r = rsq * rinv;
            const Scalar r2inv = rinv * rinv;
            const Scalar exp_val = expf(-kappa * r);
            force_divr = epsilon * exp_val * r2inv * (rinv + kappa);
            pair_eng = epsilon * exp_val * rinv;
            if (energy_shift)
                {
                const Scalar rcutinv = rsqrt_refined(rcutsq);
                pair_eng -= epsilon * expf(-kappa * rcutsq * rcutinv) * rcutinv;
                }
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "yukawa";
        }

    protected:
    Scalar rsq, rcutsq, epsilon, kappa;
    };

// V = 4 eps [(sigma/r)^12 - (sigma/r)^8]   (hoomd/md/EvaluatorPairLJ1208.h:74-88,139-165)
class EvaluatorPairLJ1208
    {
    public:
    struct param_type
        {
        Scalar lj1, lj2; // 4 eps sigma^12, 4 eps sigma^8
        };
    HB2_EVAL_COMMON(2) // {epsilon, sigma}
    static void pack(const double* user, param_type& p)
        {
        const double s4 = user[1] * user[1] * user[1] * user[1];
        p.lj1 = (Scalar)(4.0 * user[0] * s4 * s4 * s4);
        p.lj2 = (Scalar)(4.0 * user[0] * s4 * s4);
        }
    HB2_DEVICE EvaluatorPairLJ1208(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), lj1(p.lj1), lj2(p.lj2)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq && lj1 != 0)
            {
            const Scalar r2inv = rcp_refined(rsq);
            const Scalar r4inv = r2inv * r2inv;
            const Scalar r8inv = r4inv * r4inv;
            force_divr = r2inv * r8inv * (Scalar(12.0) * lj1 * r4inv - Scalar(8.0) * lj2);
            pair_eng = r8inv * (lj1 * r4inv - lj2);
            if (energy_shift)
                {
                const Scalar c2 = rcp_refined(rcutsq);
                const Scalar c4 = c2 * c2, c8 = c4 * c4;
                pair_eng -= c8 * (lj1 * c4 - lj2);
                }
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "lj1208";
        }

    protected:
    Scalar rsq, rcutsq, lj1, lj2;
    };

// V = 4 eps [(sigma/r)^8 - (sigma/r)^4]   (hoomd/md/EvaluatorPairLJ0804.h)
class EvaluatorPairLJ0804
    {
    public:
    struct param_type
        {
        Scalar lj1, lj2; // 4 eps sigma^8, 4 eps sigma^4
        };
    HB2_EVAL_COMMON(2) // {epsilon, sigma}
    static void pack(const double* user, param_type& p)
        {
        const double s4 = user[1] * user[1] * user[1] * user[1];
        p.lj1 = (Scalar)(4.0 * user[0] * s4 * s4);
        p.lj2 = (Scalar)(4.0 * user[0] * s4);
        }
    HB2_DEVICE EvaluatorPairLJ0804(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), lj1(p.lj1), lj2(p.lj2)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq && lj1 != 0)
            {
            const Scalar r2inv = rcp_refined(rsq);
            const Scalar r4inv = r2inv * r2inv;
            force_divr = r2inv * r4inv * (Scalar(8.0) * lj1 * r4inv - Scalar(4.0) * lj2);
            pair_eng = r4inv * (lj1 * r4inv - lj2);
            if (energy_shift)
                {
                const Scalar c2 = rcp_refined(rcutsq);
                const Scalar c4 = c2 * c2;
                pair_eng -= c4 * (lj1 * c4 - lj2);
                }
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "lj0804";
        }

    protected:
    Scalar rsq, rcutsq, lj1, lj2;
    };

// LJ with the force shifted to zero at r_cut   (hoomd/md/EvaluatorPairForceShiftedLJ.h:144-176)
class EvaluatorPairForceShiftedLJ
    {
    public:
    struct param_type
        {
        Scalar lj1, lj2;
        };
    HB2_EVAL_COMMON(2) // {epsilon, sigma}
    static void pack(const double* user, param_type& p)
        {
        const double s6 = user[1] * user[1] * user[1] * user[1] * user[1] * user[1];
        p.lj1 = (Scalar)(4.0 * user[0] * s6 * s6);
        p.lj2 = (Scalar)(4.0 * user[0] * s6);
        }
    HB2_DEVICE EvaluatorPairForceShiftedLJ(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), lj1(p.lj1), lj2(p.lj2)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq && lj1 != 0)
            {
            const Scalar r2inv = rcp_refined(rsq);
            const Scalar r6inv = r2inv * r2inv * r2inv;
            force_divr = r2inv * r6inv * (Scalar(12.0) * lj1 * r6inv - Scalar(6.0) * lj2);
            pair_eng = r6inv * (lj1 * r6inv - lj2);
            const Scalar rcut2inv = rcp_refined(rcutsq);
            const Scalar rcut6inv = rcut2inv * rcut2inv * rcut2inv;
            if (energy_shift)
                pair_eng -= rcut6inv * (lj1 * rcut6inv - lj2);
            const Scalar rcut_r_inv = rsqrt_refined(rsq * rcutsq);
            const Scalar force_rcut_at_rcut = rcut6inv * (Scalar(12.0) * lj1 * rcut6inv - Scalar(6.0) * lj2);
            force_divr -= rcut_r_inv * force_rcut_at_rcut;
            pair_eng += (rsq * rcut_r_inv - Scalar(1.0)) * force_rcut_at_rcut;
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "force_shifted_lj";
        }

    protected:
    Scalar rsq, rcutsq, lj1, lj2;
    };

// Mie n-m potential   (hoomd/md/EvaluatorPairMie.h:71-82,129-153)
class EvaluatorPairMie
    {
    public:
    struct param_type
        {
        Scalar m1, m2, m3, m4; // prefactor eps sigma^n, prefactor eps sigma^m, n, m
        };
    HB2_EVAL_COMMON(4) // {epsilon, sigma, n, m}
    static void pack(const double* user, param_type& p)
        {
        const double eps = user[0], sigma = user[1], n = user[2], m = user[3];
        const double out_front = (n / (n - m)) * pow(n / m, m / (n - m));
        p.m1 = (Scalar)(out_front * eps * pow(sigma, n));
        p.m2 = (Scalar)(out_front * eps * pow(sigma, m));
        p.m3 = (Scalar)n;
        p.m4 = (Scalar)m;
        }
    HB2_DEVICE EvaluatorPairMie(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), mie1(p.m1), mie2(p.m2), mie3(p.m3), mie4(p.m4)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq && mie1 != 0)
            {
            const Scalar r2inv = rcp_refined(rsq);
            const Scalar rninv = powf(r2inv, mie3 * Scalar(0.5));
            const Scalar rminv = powf(r2inv, mie4 * Scalar(0.5));
            force_divr = r2inv * (mie3 * mie1 * rninv - mie4 * mie2 * rminv);
            pair_eng = mie1 * rninv - mie2 * rminv;
            if (energy_shift)
                pair_eng -= mie1 * powf(rcutsq, -mie3 * Scalar(0.5)) - mie2 * powf(rcutsq, -mie4 * Scalar(0.5));
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "mie";
        }

    protected:
    Scalar rsq, rcutsq, mie1, mie2, mie3, mie4;
    };

// V = D0 [exp(-2 alpha (r - r0)) - 2 exp(-alpha (r - r0))]   (hoomd/md/EvaluatorPairMorse.h:125-148)
class EvaluatorPairMorse
    {
    public:
    struct param_type
        {
        Scalar D0, alpha, r0, _pad;
        };
    HB2_EVAL_COMMON(3) // {D0, alpha, r0}
    static void pack(const double* user, param_type& p)
        {
        p.D0 = (Scalar)user[0];
        p.alpha = (Scalar)user[1];
        p.r0 = (Scalar)user[2];
        p._pad = 0;
        }
    HB2_DEVICE EvaluatorPairMorse(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), D0(p.D0), alpha(p.alpha), r0(p.r0)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq)
            {
            const Scalar rinv = rsqrt_refined(rsq);
            const Scalar r = rsq * rinv;
            const Scalar ef = expf(-alpha * (r - r0));
            pair_eng = D0 * ef * (ef - Scalar(2.0));
            force_divr = Scalar(2.0) * D0 * alpha * ef * (ef - Scalar(1.0)) * rinv;
            if (energy_shift)
                {
                const Scalar rcut = sqrtf(rcutsq);
                const Scalar efc = expf(-alpha * (rcut - r0));
                pair_eng -= D0 * efc * (efc - Scalar(2.0));
                }
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "morse";
        }

    protected:
    Scalar rsq, rcutsq, D0, alpha, r0;
    };

// V = A exp(-r / rho) - C / r^6   (hoomd/md/EvaluatorPairBuckingham.h:127-158)
class EvaluatorPairBuckingham
    {
    public:
    struct param_type
        {
        Scalar A, rho, C;
        };
    HB2_EVAL_COMMON(3) // {A, rho, C}
    static void pack(const double* user, param_type& p)
        {
        p.A = (Scalar)user[0];
        p.rho = (Scalar)user[1];
        p.C = (Scalar)user[2];
        }
    HB2_DEVICE EvaluatorPairBuckingham(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), A(p.A), rho(p.rho), C(p.C)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq && rho > Scalar(0.0))
            {
            const Scalar r = sqrtf(rsq);
            const Scalar exp_factor = A * __expf(-r / rho);
            const Scalar r2inv = Scalar(1.0) / rsq;
            const Scalar r6inv = r2inv * r2inv * r2inv;
            force_divr = (exp_factor / (rho * r)) - (r2inv * r6inv * Scalar(6.0) * C);
            pair_eng = exp_factor - r6inv * C;
            if (energy_shift)
                {
                const Scalar rcut = sqrtf(rcutsq);
                const Scalar rcut2inv = Scalar(1.0) / rcutsq;
                const Scalar rcut6inv = rcut2inv * rcut2inv * rcut2inv;
                pair_eng -= A * __expf(-rcut / rho) - rcut6inv * C;
                }
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "buckingham";
        }

    protected:
    Scalar rsq, rcutsq, A, rho, C;
    };

// LJ in r - delta   (hoomd/md/EvaluatorPairExpandedLJ.h:66-84 parameters, :150-186 evaluation)
class EvaluatorPairExpandedLJ
    {
    public:
    struct param_type
        {
        Scalar lj1, lj2, delta;
        };
    HB2_EVAL_COMMON(3) // {epsilon, sigma, delta}
    static void pack(const double* user, param_type& p)
        {
        // the reference keeps sigma^6 and 4 epsilon as Scalars and forms lj1, lj2 from them
        const Scalar sigma = (Scalar)user[1];
        const Scalar sigma_3 = sigma * sigma * sigma;
        const Scalar sigma_6 = sigma_3 * sigma_3;
        const Scalar epsilon_x_4 = Scalar(4.0) * (Scalar)user[0];
        p.lj1 = epsilon_x_4 * sigma_6 * sigma_6;
        p.lj2 = epsilon_x_4 * sigma_6;
        p.delta = (Scalar)user[2];
        }
    HB2_DEVICE EvaluatorPairExpandedLJ(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), lj1(p.lj1), lj2(p.lj2), delta(p.delta)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        const Scalar rinv = rsqrt_refined(rsq);
        const Scalar r = Scalar(1.0) / rinv;
        if (rsq < rcutsq && lj1 != 0)
            {
            const Scalar rmd = r - delta;
            const Scalar rmdinv = Scalar(1.0) / rmd;
            const Scalar rmd2inv = rmdinv * rmdinv;
            const Scalar rmd6inv = rmd2inv * rmd2inv * rmd2inv;
            force_divr = rinv * rmdinv * rmd6inv * (Scalar(12.0) * lj1 * rmd6inv - Scalar(6.0) * lj2);
            pair_eng = rmd6inv * (lj1 * rmd6inv - lj2);
            if (energy_shift)
                {
                const Scalar r_cut_shifted_inv = Scalar(1.0) / (sqrtf(rcutsq) - delta);
                const Scalar r_cut2_inv = r_cut_shifted_inv * r_cut_shifted_inv;
                const Scalar r_cut6_inv = r_cut2_inv * r_cut2_inv * r_cut2_inv;
                pair_eng -= r_cut6_inv * (lj1 * r_cut6_inv - lj2);
                }
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "expanded_lj";
        }

    protected:
    Scalar rsq, rcutsq, lj1, lj2, delta;
    };

// V = eps exp(-(r - delta)^2 / (2 sigma^2))   (hoomd/md/EvaluatorPairExpandedGaussian.h:127-162)
class EvaluatorPairExpandedGaussian
    {
    public:
    struct param_type
        {
        Scalar epsilon, sigma, delta;
        };
    HB2_EVAL_COMMON(3) // {epsilon, sigma, delta}
    static void pack(const double* user, param_type& p)
        {
        p.epsilon = (Scalar)user[0];
        p.sigma = (Scalar)user[1];
        p.delta = (Scalar)user[2];
        }
    HB2_DEVICE EvaluatorPairExpandedGaussian(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), epsilon(p.epsilon), sigma(p.sigma), delta(p.delta)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq)
            {
            const Scalar r = sqrtf(rsq);
            const Scalar rinv = Scalar(1.0) / r;
            const Scalar rmd = r - delta;
            const Scalar sigma_sq = sigma * sigma;
            const Scalar exp_val = __expf(Scalar(-0.5) * (rmd * rmd) / sigma_sq);
            force_divr = epsilon / sigma_sq * exp_val * rmd * rinv;
            pair_eng = epsilon * exp_val;
            if (energy_shift)
                {
                const Scalar rcutmd = sqrtf(rcutsq) - delta;
                pair_eng -= epsilon * __expf(Scalar(-0.5) * (rcutmd * rcutmd) / sigma_sq);
                }
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "expanded_gaussian";
        }

    protected:
    Scalar rsq, rcutsq, epsilon, sigma, delta;
    };

// V = 1/r^12 - 2/r^6 - eps exp(-(r - r0)^2 / (2 sigma^2))   (hoomd/md/EvaluatorPairLJGauss.h:131-166)
class EvaluatorPairLJGauss
    {
    public:
    struct param_type
        {
        Scalar epsilon, sigma, r0;
        };
    HB2_EVAL_COMMON(3) // {epsilon, sigma, r0}
    static void pack(const double* user, param_type& p)
        {
        p.epsilon = (Scalar)user[0];
        p.sigma = (Scalar)user[1];
        p.r0 = (Scalar)user[2];
        }
    HB2_DEVICE EvaluatorPairLJGauss(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), epsilon(p.epsilon), sigma(p.sigma), r0(p.r0)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq >= rcutsq)
            return false;
        const Scalar r = sqrtf(rsq);
        const Scalar sigma2 = sigma * sigma;
        const Scalar rdiff = r - r0;
        const Scalar rdiff_sigma2 = rdiff / sigma2;
        const Scalar exp_val = __expf(Scalar(-0.5) * rdiff_sigma2 * rdiff);
        const Scalar r2inv = Scalar(1.0) / rsq;
        const Scalar r6inv = r2inv * r2inv * r2inv;
        force_divr = (r2inv * r6inv * Scalar(12.0) * (r6inv - Scalar(1.0)))
                     - (exp_val * epsilon * rdiff_sigma2 / r);
        pair_eng = r6inv * (r6inv - Scalar(2.0)) - exp_val * epsilon;
        if (energy_shift)
            {
            const Scalar rcut2inv = Scalar(1.0) / rcutsq;
            const Scalar rcut6inv = rcut2inv * rcut2inv * rcut2inv;
            const Scalar d = sqrtf(rcutsq) - r0;
            pair_eng -= rcut6inv * (rcut6inv - Scalar(2.0)) - (epsilon * __expf(Scalar(-0.5) * d * d / sigma2));
            }
        return true;
        }
    static const char* getName()
        {
        return "lj_gauss";
        }

    protected:
    Scalar rsq, rcutsq, epsilon, sigma, r0;
    };

// ten Wolde-Frenkel: V = 4 eps / alpha^2 [1/(r^2/sigma^2 - 1)^6 - alpha /(r^2/sigma^2 - 1)^3]
// (hoomd/md/EvaluatorPairTWF.h:58-64 parameters, :139-180 evaluation; overlapping pairs are infinite)
class EvaluatorPairTWF
    {
    public:
    struct param_type
        {
        Scalar sigma, alpha, prefactor;
        };
    HB2_EVAL_COMMON(3) // {epsilon, sigma, alpha}
    static void pack(const double* user, param_type& p)
        {
        p.sigma = (Scalar)user[1];
        p.alpha = (Scalar)user[2];
        p.prefactor = (Scalar)(4.0 * (double)(Scalar)user[0] / ((double)p.alpha * (double)p.alpha));
        }
    HB2_DEVICE EvaluatorPairTWF(Scalar _rsq, Scalar _rcutsq, const param_type& p)
        : rsq(_rsq), rcutsq(_rcutsq), sigma(p.sigma), alpha(p.alpha), prefactor(p.prefactor)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (rsq < rcutsq)
            {
            const Scalar sigma2 = sigma * sigma;
            if (rsq <= sigma2)
                {
                pair_eng = INFINITY;
                force_divr = INFINITY;
                return true;
                }
            const Scalar common_term = Scalar(1.0) / (rsq / sigma2 - Scalar(1.0));
            const Scalar common_term3 = common_term * common_term * common_term;
            const Scalar common_term6 = common_term3 * common_term3;
            pair_eng = prefactor * (common_term6 - alpha * common_term3);
            const Scalar force_term = Scalar(6.0) * common_term / sigma2;
            force_divr = prefactor * force_term * (Scalar(2.0) * common_term6 - alpha * common_term3);
            if (energy_shift)
                {
                const Scalar ct = Scalar(1.0) / (rcutsq / sigma2 - Scalar(1.0));
                const Scalar ct3 = ct * ct * ct;
                pair_eng -= prefactor * (ct3 * ct3 - alpha * ct3);
                }
            return true;
            }
        return false;
        }
    static const char* getName()
        {
        return "twf";
        }

    protected:
    Scalar rsq, rcutsq, sigma, alpha, prefactor;
    };
// ---------------------------------------------------------------------------------------------
// Tabulated potential (hoomd/md/EvaluatorPairTable.h:42-205): U(r) and F(r) = -dU/dr sampled at
// r_min + i (r_cut - r_min) / width, linear interpolation, nothing outside [r_min, r_cut). The
// variable-size tables live in one device buffer of the pair force; param_type carries a pointer into
// it (V[width] followed by F[width]) — the role ManagedArray plays in the reference.
// ---------------------------------------------------------------------------------------------
class EvaluatorPairTable
    {
    public:
    static constexpr bool has_packed = false;
    static constexpr bool tune_tpp = false;
    static constexpr bool needs_thermo = false;
    struct param_type
        {
        Scalar rmin;
        uint32_t width;
        const Scalar* table;
        };
    static constexpr unsigned n_user_params = 0; // set through hb2_pair_set_table
    HB2_DEVICE EvaluatorPairTable(Scalar _rsq, Scalar _rcutsq, const param_type& _params)
        : rsq(_rsq), rcutsq(_rcutsq), rmin(_params.rmin), width(_params.width), table(_params.table)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        const Scalar r = sqrtf(rsq);
        if (rsq >= rcutsq || r < rmin || width == 0)
            return false;
        const Scalar rcut = sqrtf(rcutsq);
        const Scalar delta_r = (rcut - rmin) / (Scalar)width;
        const Scalar value_f = (r - rmin) / delta_r;
        uint32_t value_i = (uint32_t)floorf(value_f);
        if (value_i >= width) // r within rounding of r_cut
            value_i = width - 1;
        const Scalar V0 = __ldg(table + value_i), F0 = __ldg(table + width + value_i);
        Scalar V1 = 0, F1 = 0;
        if (value_i + 1 < width)
            {
            V1 = __ldg(table + value_i + 1);
            F1 = __ldg(table + width + value_i + 1);
            }
        const Scalar f = value_f - (Scalar)value_i;
        const Scalar V = V0 + f * (V1 - V0);
        const Scalar F = F0 + f * (F1 - F0);
        if (rsq > Scalar(0.0))
            force_divr = F / r;
        pair_eng = V;
        return true;
        }
    static const char* getName()
        {
        return "table";
        }

    protected:
    Scalar rsq, rcutsq, rmin;
    uint32_t width;
    const Scalar* table;
    };

// ---------------------------------------------------------------------------------------------
// Evaluators with five and six packed parameters (the parameter stride is HB2_PARAM_FLOATS = 8).
// ---------------------------------------------------------------------------------------------

// Mie n-m in r - delta: V = A (sigma/(r-delta))^n - B (sigma/(r-delta))^m with the Mie prefactor
// (hoomd/md/EvaluatorPairExpandedMie.h:60-75 parameters, :127-160 evaluation)
class EvaluatorPairExpandedMie
    {
    public:
    struct param_type
        {
        Scalar rep, att, n, m, delta;
        };
    HB2_EVAL_COMMON(5) // {epsilon, sigma, n, m, delta}
    static void pack(const double* user, param_type& p)
        {
        const double eps = user[0], sigma = user[1], n = user[2], m = user[3];
        const double pref = (n / (n - m)) * pow(n / m, m / (n - m));
        p.rep = (Scalar)(pref * eps * pow(sigma, n));
        p.att = (Scalar)(pref * eps * pow(sigma, m));
        p.n = (Scalar)n;
        p.m = (Scalar)m;
        p.delta = (Scalar)user[4];
        }
    HB2_DEVICE EvaluatorPairExpandedMie(Scalar _rsq, Scalar _rcutsq, const param_type& q)
        : rsq(_rsq), rcutsq(_rcutsq), p(q)
        {
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (!(rsq < rcutsq) || p.rep == 0)
            return false;
        const Scalar rinv = rsqrt_refined(rsq);
        const Scalar s = Scalar(1.0) / rinv - p.delta; // r - delta
        const Scalar sn = p.rep * powf(s, -p.n), sm = p.att * powf(s, -p.m);
        pair_eng = sn - sm;
        force_divr = rinv * (p.n * sn - p.m * sm) / s;
        if (energy_shift)
            {
            const Scalar sc = sqrtf(rcutsq) - p.delta;
            pair_eng -= p.rep * powf(sc, -p.n) - p.att * powf(sc, -p.m);
            }
        return true;
        }
    static const char* getName()
        {
        return "expanded_mie";
        }

    protected:
    Scalar rsq, rcutsq;
    param_type p;
    };

// Oscillating pair potential V = C1 r^-eta1 + C2 r^-eta2 cos(k r - phi)
// (hoomd/md/EvaluatorPairOPP.h:46-82 parameters, :128-163 evaluation)
class EvaluatorPairOPP
    {
    public:
    struct param_type
        {
        Scalar C1, C2, eta1, eta2, k, phi;
        };
    HB2_EVAL_COMMON(6) // {C1, C2, eta1, eta2, k, phi}
    static void pack(const double* user, param_type& p)
        {
        p.C1 = (Scalar)user[0];
        p.C2 = (Scalar)user[1];
        p.eta1 = (Scalar)user[2];
        p.eta2 = (Scalar)user[3];
        p.k = (Scalar)user[4];
        p.phi = (Scalar)user[5];
        }
    HB2_DEVICE EvaluatorPairOPP(Scalar _rsq, Scalar _rcutsq, const param_type& q) : rsq(_rsq), rcutsq(_rcutsq), p(q) { }
    HB2_DEVICE static void terms(const param_type& p, const Scalar r, Scalar& t1, Scalar& env, Scalar& sn, Scalar& cs)
        {
        t1 = p.C1 * powf(r, -p.eta1);  // C1 r^-eta1
        env = p.C2 * powf(r, -p.eta2); // C2 r^-eta2, the envelope of the oscillation
        sincosf(p.k * r - p.phi, &sn, &cs);
        }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (!(rsq < rcutsq))
            return false;
        const Scalar r = sqrtf(rsq), rinv = Scalar(1.0) / r;
        Scalar t1, env, sn, cs;
        terms(p, r, t1, env, sn, cs);
        pair_eng = t1 + env * cs;
        // -dV/dr = eta1 t1 / r + env (eta2 cos / r + k sin)
        force_divr = rinv * (rinv * (p.eta1 * t1 + p.eta2 * env * cs) + p.k * env * sn);
        if (energy_shift)
            {
            Scalar c1, ce, s2, c2;
            terms(p, sqrtf(rcutsq), c1, ce, s2, c2);
            pair_eng -= c1 + ce * c2;
            }
        return true;
        }
    static const char* getName()
        {
        return "opp";
        }

    protected:
    Scalar rsq, rcutsq;
    param_type p;
    };

// Fourier pair potential: V = r^-12 + r^-2 sum_{n=1..4} [a_n cos(n x) + b_n sin(n x)], x = pi r / r_cut, with
// a_1 and b_1 fixed by a_2..4, b_2..4 so that V and its slope vanish at the cut-off
// (hoomd/md/EvaluatorPairFourier.h:55-99 parameters, :135-180 evaluation). No energy shift.
class EvaluatorPairFourier
    {
    public:
    struct param_type
        {
        Scalar a[3], b[3];
        };
    HB2_EVAL_COMMON(6) // {a2, a3, a4, b2, b3, b4}
    static void pack(const double* user, param_type& p)
        {
        for (int i = 0; i < 3; i++)
            {
            p.a[i] = (Scalar)user[i];
            p.b[i] = (Scalar)user[3 + i];
            }
        }
    HB2_DEVICE EvaluatorPairFourier(Scalar _rsq, Scalar _rcutsq, const param_type& q) : rsq(_rsq), rcutsq(_rcutsq), p(q) { }
    HB2_DEVICE bool evalForceAndEnergy(Scalar& force_divr, Scalar& pair_eng, bool energy_shift)
        {
        if (!(rsq < rcutsq))
            return false;
        const Scalar r = sqrtf(rsq), rinv = Scalar(1.0) / r, r2inv = rinv * rinv;
        const Scalar w = Scalar(3.14159265358979323846) / sqrtf(rcutsq); // dx/dr
        const Scalar x = r * w;
        // first harmonic: a_1 = a_2 - a_3 + a_4, b_1 = 2 b_2 - 3 b_3 + 4 b_4
        const Scalar an[4] = {p.a[0] - p.a[1] + p.a[2], p.a[0], p.a[1], p.a[2]};
        const Scalar bn[4] = {Scalar(2.0) * p.b[0] - Scalar(3.0) * p.b[1] + Scalar(4.0) * p.b[2], p.b[0], p.b[1], p.b[2]};
        Scalar series = 0, dseries = 0; // S(x) and -dS/dx
#pragma unroll
        for (int n = 1; n <= 4; n++)
            {
            Scalar sn, cs;
            sincosf((Scalar)n * x, &sn, &cs);
            series += an[n - 1] * cs + bn[n - 1] * sn;
            dseries += (Scalar)n * (an[n - 1] * sn - bn[n - 1] * cs);
            }
        const Scalar r6inv = r2inv * r2inv * r2inv, r12inv = r6inv * r6inv;
        pair_eng = r12inv + r2inv * series;
        // -dV/dr / r = 12 r^-14 + 2 S r^-4 + w (-dS/dx) r^-3
        force_divr = r2inv * (Scalar(12.0) * r12inv + Scalar(2.0) * r2inv * series + w * rinv * dseries);
        return true;
        }
    static const char* getName()
        {
        return "fourier";
        }

    protected:
    Scalar rsq, rcutsq;
    param_type p;
    };

#undef HB2_EVAL_COMMON
    } // namespace hb2
