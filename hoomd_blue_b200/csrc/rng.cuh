// rng.cuh — counter-based RNG of the pair-force path: Philox4x32-10 (Salmon, Moraes, Dror, Shaw,
// SC'11; the reference vendors random123, hoomd/extern/random123 philox.h) with HOOMD's Seed/Counter
// packing (hoomd/RandomNumbers.h:136-199) and the distributions the path draws from
// (uniform :305-335, normal :361-374, gamma :461-505). Host+device so the host can check itself.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace hb2
    {
#define HB2_HD __host__ __device__ __forceinline__

HB2_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
    {
#ifdef __CUDA_ARCH__
    hi = __umulhi(a, b);
    lo = a * b;
#else
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
    }

// Ten rounds; multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85.
HB2_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
    {
#pragma unroll
    for (int r = 0; r < 10; r++)
        {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c[0], hi0, lo0);
        mulhilo32(0xCD9E8D57u, c[2], hi1, lo1);
        const uint32_t n0 = hi1 ^ c[1] ^ k0;
        const uint32_t n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0;
        c[1] = lo1;
        c[2] = n2;
        c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
        }
    }

// RNG ids (hoomd/RNGIdentifiers.h)
constexpr uint8_t RNG_ID_DPD = 22;   // EvaluatorPairDPDThermo
constexpr uint8_t RNG_ID_BUSSI = 45; // BussiThermostat
constexpr uint8_t RNG_ID_LANGEVIN = 25; // TwoStepLangevin

// hoomd/RandomNumbers.h:153-158: key = {id<<24 | seed16<<8 | timestep bits 39..32, timestep low 32}
HB2_HD void seed_key(uint8_t id, uint64_t timestep, uint16_t seed, uint32_t& k0, uint32_t& k1)
    {
    k0 = (uint32_t)id << 24 | (uint32_t)seed << 8 | (uint32_t)((timestep >> 32) & 0xffu);
    k1 = (uint32_t)(timestep & 0xffffffffull);
    }

struct Rng
    {
    uint32_t k0, k1, ctr[4];
    // Counter(a, b, c, d) -> ctr = {d<<16, c, b, a} (hoomd/RandomNumbers.h:185-187)
    HB2_HD Rng(uint8_t id, uint64_t timestep, uint16_t seed, uint32_t a, uint32_t b = 0,
               uint32_t c = 0, uint16_t d = 0)
        {
        seed_key(id, timestep, seed, k0, k1);
        ctr[0] = (uint32_t)d << 16;
        ctr[1] = c;
        ctr[2] = b;
        ctr[3] = a;
        }
    // hoomd/RandomNumbers.h:264-270: one Philox call per draw, low counter word increments
    HB2_HD void next(uint32_t u[4])
        {
        u[0] = ctr[0];
        u[1] = ctr[1];
        u[2] = ctr[2];
        u[3] = ctr[3];
        philox4x32_10(u, k0, k1);
        ctr[0] += 1;
        }
    };

// hoomd/RandomNumbers.h:63-74: u64 -> (0,1): in * 2^-64 + 2^-65
HB2_HD double u01_u64(uint64_t in)
    {
    return (double)in * 5.421010862427522e-20 + 2.710505431213761e-20;
    }
// hoomd/RandomNumbers.h:76-87: u64 -> (-1,1): int64(in) * 2^-63 + 2^-64
HB2_HD double uneg11_u64(uint64_t in)
    {
    return (double)(int64_t)in * 1.0842021724855044e-19 + 5.421010862427522e-20;
    }
HB2_HD double rng_canonical(Rng& g)
    {
    uint32_t u[4];
    g.next(u);
    return u01_u64((uint64_t)u[0] << 32 | u[1]);
    }
// Box-Muller, one value per draw (hoomd/RandomNumbers.h:361-374)
HB2_HD double rng_normal(Rng& g)
    {
    uint32_t u[4];
    g.next(u);
    const uint64_t u0 = (uint64_t)u[0] << 32 | u[1], u1 = (uint64_t)u[2] << 32 | u[3];
    const double x = sin(3.14159265358979323846 * uneg11_u64(u0));
    const double r = sqrt(-2.0 * log(u01_u64(u1)));
    return x * r;
    }
// Marsaglia-Tsang (hoomd/RandomNumbers.h:461-505)
HB2_HD double rng_gamma(Rng& g, double alpha_in, double bparam)
    {
    double alpha = alpha_in;
    if (alpha_in < 1.0)
        alpha += 1.0;
    const double d = alpha - (1. / 3.);
    const double c = (1.0 / sqrt(d)) / 3.;
    double v;
    for (;;)
        {
        double x;
        do
            {
            x = rng_normal(g);
            v = 1.0 + c * x;
            } while (v <= 0.);
        v = v * v * v;
        const double x2 = x * x;
        const double u = rng_canonical(g);
        if (u < 1.0 - 0.0331 * x2 * x2)
            break;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v)))
            break;
        }
    v *= d * bparam;
    if (alpha_in < 1.0)
        v *= pow(rng_canonical(g), 1.0 / alpha_in);
    return v;
    }

// BussiThermostat::getRescalingFactorsOne (hoomd/md/Thermostat.h:357-447)
HB2_HD double bussi_factor(double K, double dof, double dt, double kT, double tau, uint64_t timestep,
                           uint16_t seed, uint32_t first_tag)
    {
    if (dt == 0.0 || dof == 0.0)
        return 1.0;
    Rng g(RNG_ID_BUSSI, timestep, seed, first_tag);
    double decay = 0.0;
    if (tau != 0.0)
        decay = exp(-dt / tau);
    const double rn = rng_normal(g);
    double r_gamma = 0.0;
    if (dof > 1.0)
        r_gamma = 2.0 * rng_gamma(g, (dof - 1.0) / 2.0, 1.0);
    const double v = kT / 2.0 / K;
    const double term1 = v * (1.0 - decay) * (r_gamma + rn * rn);
    const double term2 = 2.0 * rn * sqrt(v * (1.0 - decay) * decay);
    return sqrt(decay + term1 + term2);
    }

// DPD pair draw: uniform in (-1,1) keyed (22, seed, timestep), Counter(min_tag, max_tag)
// (hoomd/md/EvaluatorPairDPDThermoDPD.h:237-254)
HB2_HD void dpd_pair_bits(uint16_t seed, uint32_t tag_i, uint32_t tag_j, uint64_t timestep,
                          uint32_t& w0, uint32_t& w1)
    {
    const uint32_t lo = tag_i < tag_j ? tag_i : tag_j, hi = tag_i < tag_j ? tag_j : tag_i;
    uint32_t k0, k1;
    seed_key(RNG_ID_DPD, timestep, seed, k0, k1);
    uint32_t c[4] = {0u, 0u, hi, lo};
    philox4x32_10(c, k0, k1);
    w0 = c[0];
    w1 = c[1];
    }
HB2_HD double dpd_alpha(uint16_t seed, uint32_t tag_i, uint32_t tag_j, uint64_t timestep)
    {
    uint32_t w0, w1;
    dpd_pair_bits(seed, tag_i, tag_j, timestep, w0, w1);
    return -1.0 + 2.0 * u01_u64((uint64_t)w0 << 32 | w1);
    }
    } // namespace hb2
