// bond.cu — bond forces (harmonic, FENE + WCA) and the bond-exclusion filter of the neighbour list.
//
// Replaces PotentialBond<EvaluatorBondHarmonic>::computeForces (hoomd/md/PotentialBond.h:176-312,
// GPU: gpu_compute_bond_forces_kernel hoomd/md/PotentialBondGPU.cuh) and NeighborList::filterNlist with
// updateExListIdx (hoomd/md/NeighborList.cc:1448-1536, GPU: gpu_nlist_filter_kernel
// hoomd/md/NeighborListGPU.cu:158-267).
//
// Bonds live in a table keyed by particle TAG (CSR: head[tag] .. head[tag + 1] -> partner tag, bond
// type), so the re-ordering of particles at every rebuild does not touch it; a partner's current index
// comes from rtag. Owner-computes: particle i sums its own bonds (F_i = F/r (r_i - r_j), half the bond
// energy and half the virial), no atomics, fixed order -> bitwise reproducible.
#include "common.cuh"
#include "kernels.cuh"

namespace hb2
    {
constexpr int BOND_BLOCK = 128;

// EvaluatorBondHarmonic (hoomd/md/EvaluatorBondHarmonic.h:104-124): p = (k, r0)
struct BondHarmonic
    {
    HB2_DEVICE static bool eval(float rsq, const float* p, float& fdr, float& eng)
        {
        const float r = sqrtf(rsq);
        fdr = p[0] * (p[1] / r - 1.0f);
        if (!isfinite(fdr))
            fdr = 0.f;
        const float d = p[1] - r;
        eng = 0.5f * p[0] * d * d;
        return true;
        }
    };

// EvaluatorBondFENE (hoomd.md.bond.FENEWCA, hoomd/md/EvaluatorBondFENE.h:117-158):
// p = (k, r0, lj1, lj2, delta); false when the bond is stretched beyond r0
struct BondFENEWCA
    {
    HB2_DEVICE static bool eval(float rsq_in, const float* p, float& fdr, float& eng)
        {
        const float K = p[0], r_0 = p[1], lj1 = p[2], lj2 = p[3], delta = p[4];
        const float r = sqrtf(rsq_in);
        const float rtemp = r - delta;
        const float rmdoverr = rtemp / r;
        const float rsq = rtemp * rtemp;
        const float r2inv = 1.0f / rsq;
        const float r6inv = r2inv * r2inv * r2inv;
        float wca = 0.f, pair_eng = 0.f;
        const float sigma6inv = lj2 / lj1;
        const float epsilon = lj2 * lj2 / 4.0f / lj1;
        if (lj1 != 0.f && r6inv > sigma6inv / 2.0f)
            {
            wca = r2inv * r6inv * (12.0f * lj1 * r6inv - 6.0f * lj2);
            pair_eng = r6inv * (lj1 * r6inv - lj2) + epsilon;
            }
        fdr = wca * rmdoverr;
        eng = pair_eng;
        if (K != 0.f)
            {
            if (rsq >= r_0 * r_0)
                return false;
            fdr += -K / (1.0f - rsq / (r_0 * r_0)) * rmdoverr;
            eng += -0.5f * K * (r_0 * r_0) * logf(1.0f - rsq / (r_0 * r_0));
            }
        return true;
        }
    };

// ACCUM: add to `force` / `virial` (the integrator's net force = pair force + bonds); otherwise
// overwrite (the bond force's own arrays, for its getters).
template<class Ev, bool ACCUM>
__global__ void __launch_bounds__(BOND_BLOCK)
    bond_force_kernel(const float4* __restrict__ pos, const uint32_t* __restrict__ tag,
                         const uint32_t* __restrict__ rtag, uint32_t N, uint32_t Ntot,
                         const uint32_t* __restrict__ head, const uint32_t* __restrict__ partner,
                         const uint32_t* __restrict__ btype, const float* __restrict__ params, BoxD box,
                         float4* __restrict__ force, float* __restrict__ virial, size_t virial_pitch,
                         const uint32_t* __restrict__ skip_flag, uint32_t* __restrict__ err)
    {
    if (skip_flag && *skip_flag)
        return;
    const uint32_t i = blockIdx.x * BOND_BLOCK + threadIdx.x;
    if (i >= N)
        return;
    const uint32_t t = tag[i];
    const uint32_t k0 = head[t], k1 = head[t + 1];
    if (ACCUM && k0 == k1)
        return;
    const float4 pi = pos[i];
    float fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (uint32_t k = k0; k < k1; k++)
        {
        const uint32_t j = rtag[partner[k]];
        if (j >= Ntot)
            {
            atomicMax(err, partner[k] + 1u); // "bond ... is incomplete" (PotentialBond.h:229-234)
            continue;
            }
        const float4 pj = pos[j];
        // separation in fp64 with the exact minimum image (a bond may cross the box; an fp32 +-L
        // would cost half an ulp of L, i.e. k * 4e-6 of absolute force error)
        double dx = (double)pi.x - (double)pj.x, dy = (double)pi.y - (double)pj.y,
               dz = (double)pi.z - (double)pj.z;
        min_image_exact(box, dx, dy, dz);
        const float x = (float)dx, y = (float)dy, z = (float)dz;
        const float rsq = x * x + y * y + z * z;
        float fdr, eng;
        if (!Ev::eval(rsq, params + HB2_BOND_PARAM_FLOATS * btype[k], fdr, eng))
            {
            atomicMax(err + 1, t + 1u); // "bond out of bounds" (PotentialBond.h:297-303)
            continue;
            }
        fx += fdr * x;
        fy += fdr * y;
        fz += fdr * z;
        pe += 0.5f * eng;
        if (virial)
            {
            const float f2 = 0.5f * fdr;
            v[0] += x * x * f2;
            v[1] += x * y * f2;
            v[2] += x * z * f2;
            v[3] += y * y * f2;
            v[4] += y * z * f2;
            v[5] += z * z * f2;
            }
        }
    if (ACCUM)
        {
        float4 f = force[i];
        f.x += fx;
        f.y += fy;
        f.z += fz;
        f.w += pe;
        force[i] = f;
        if (virial)
#pragma unroll
            for (int q = 0; q < 6; q++)
                virial[q * virial_pitch + i] += v[q];
        }
    else
        {
        force[i] = make_float4(fx, fy, fz, pe);
        if (virial)
#pragma unroll
            for (int q = 0; q < 6; q++)
                virial[q * virial_pitch + i] = v[q];
        }
    }

template<class Ev>
static hb2_status launch_bond_ev(const BondArgs& a, bool accumulate, const hb2_box& box, cudaStream_t stream)
    {
    const uint32_t blocks = (a.N + BOND_BLOCK - 1) / BOND_BLOCK;
    if (accumulate)
        bond_force_kernel<Ev, true><<<blocks, BOND_BLOCK, 0, stream>>>(
            a.pos, a.tag, a.rtag, a.N, a.Ntot, a.head, a.partner, a.btype, a.params, boxd(box), a.force,
            a.virial, a.virial_pitch, a.skip_flag, a.err);
    else
        bond_force_kernel<Ev, false><<<blocks, BOND_BLOCK, 0, stream>>>(
            a.pos, a.tag, a.rtag, a.N, a.Ntot, a.head, a.partner, a.btype, a.params, boxd(box), a.force,
            a.virial, a.virial_pitch, a.skip_flag, a.err);
    HB2_CHECK_LAUNCH("bond_force_kernel");
    return HB2_OK;
    }

hb2_status launch_bond_forces(hb2_bond_kind kind, const BondArgs& a, bool accumulate, const hb2_box& box,
                              cudaStream_t stream)
    {
    if (a.N == 0)
        return HB2_OK;
    if (kind == HB2_BOND_FENEWCA)
        return launch_bond_ev<BondFENEWCA>(a, accumulate, box, stream);
    return launch_bond_ev<BondHarmonic>(a, accumulate, box, stream);
    }

// Angle forces, owner-computes: particle i walks the angles it is a member of (table keyed by tag:
// the three member tags and `type | role << 30`), recomputes the triplet's forces exactly as the
// reference's loop over angles does and keeps its own share (role 0: a, 1: the vertex b, 2: c).
// KIND 0: hoomd/md/HarmonicAngleForceCompute.cc:104-258; 1: hoomd/md/CosineSqAngleForceCompute.cc.
template<int KIND, bool ACCUM>
__global__ void __launch_bounds__(BOND_BLOCK)
    angle_force_kernel(const float4* __restrict__ pos, const uint32_t* __restrict__ tag,
                       const uint32_t* __restrict__ rtag, uint32_t N, uint32_t Ntot,
                       const uint32_t* __restrict__ head, const uint4* __restrict__ members,
                       const float* __restrict__ params, BoxD box, float4* __restrict__ force,
                       float* __restrict__ virial, size_t virial_pitch,
                       const uint32_t* __restrict__ skip_flag, uint32_t* __restrict__ err)
    {
    if (skip_flag && *skip_flag)
        return;
    const uint32_t i = blockIdx.x * BOND_BLOCK + threadIdx.x;
    if (i >= N)
        return;
    const uint32_t t = tag[i];
    const uint32_t k0 = head[t], k1 = head[t + 1];
    if (ACCUM && k0 == k1)
        return;
    float fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (uint32_t k = k0; k < k1; k++)
        {
        const uint4 m = members[k];
        const uint32_t ia = rtag[m.x], ib = rtag[m.y], ic = rtag[m.z];
        if (ia >= Ntot || ib >= Ntot || ic >= Ntot)
            {
            atomicMax(err, t + 1u);
            continue;
            }
        const uint32_t role = m.w >> 30, type = m.w & 0x3fffffffu;
        const float4 pa = pos[ia], pb = pos[ib], pc = pos[ic];
        double ax = (double)pa.x - (double)pb.x, ay = (double)pa.y - (double)pb.y, az = (double)pa.z - (double)pb.z;
        double cx = (double)pc.x - (double)pb.x, cy = (double)pc.y - (double)pb.y, cz = (double)pc.z - (double)pb.z;
        min_image_exact(box, ax, ay, az);
        min_image_exact(box, cx, cy, cz);
        // The triplet's geometry is evaluated in fp64: near 0 and 180 degrees 1 - cos^2 and acos lose
        // half of fp32's digits (chains with t0 = pi sit exactly there), and there are only a few
        // angles per particle.
        const double dab[3] = {ax, ay, az}, dcb[3] = {cx, cy, cz};
        const double rsqab = dab[0] * dab[0] + dab[1] * dab[1] + dab[2] * dab[2];
        const double rab = sqrt(rsqab);
        const double rsqcb = dcb[0] * dcb[0] + dcb[1] * dcb[1] + dcb[2] * dcb[2];
        const double rcb = sqrt(rsqcb);
        double c_abbc = (dab[0] * dcb[0] + dab[1] * dcb[1] + dab[2] * dcb[2]) / (rab * rcb);
        c_abbc = fmin(1.0, fmax(-1.0, c_abbc));
        const double K = (double)params[HB2_BOND_PARAM_FLOATS * type],
                     t_0 = (double)params[HB2_BOND_PARAM_FLOATS * type + 1];
        double a, eng;
        if (KIND == 0)
            {
            double s_abbc = sqrt(1.0 - c_abbc * c_abbc);
            if (s_abbc < 0.001)
                s_abbc = 0.001;
            s_abbc = 1.0 / s_abbc;
            const double dth = acos(c_abbc) - t_0;
            const double tk = K * dth;
            a = -1.0 * tk * s_abbc;
            eng = (tk * dth) * (1.0 / 6.0);
            }
        else
            {
            const double dcosth = c_abbc - cos(t_0);
            const double tk = K * dcosth;
            a = tk;
            eng = (tk * dcosth) * (1.0 / 6.0);
            }
        const double a11 = a * c_abbc / rsqab, a12 = -a / (rab * rcb), a22 = a * c_abbc / rsqcb;
        double fab[3], fcb[3];
#pragma unroll
        for (int d = 0; d < 3; d++)
            {
            fab[d] = a11 * dab[d] + a12 * dcb[d];
            fcb[d] = a22 * dcb[d] + a12 * dab[d];
            }
        const double sa = role == 0 ? 1.0 : (role == 1 ? -1.0 : 0.0);
        const double sc = role == 2 ? 1.0 : (role == 1 ? -1.0 : 0.0);
        fx += (float)(sa * fab[0] + sc * fcb[0]);
        fy += (float)(sa * fab[1] + sc * fcb[1]);
        fz += (float)(sa * fab[2] + sc * fcb[2]);
        pe += (float)eng;
        if (virial)
            {
            v[0] += (float)((1. / 3.) * (dab[0] * fab[0] + dcb[0] * fcb[0]));
            v[1] += (float)((1. / 3.) * (dab[1] * fab[0] + dcb[1] * fcb[0]));
            v[2] += (float)((1. / 3.) * (dab[2] * fab[0] + dcb[2] * fcb[0]));
            v[3] += (float)((1. / 3.) * (dab[1] * fab[1] + dcb[1] * fcb[1]));
            v[4] += (float)((1. / 3.) * (dab[2] * fab[1] + dcb[2] * fcb[1]));
            v[5] += (float)((1. / 3.) * (dab[2] * fab[2] + dcb[2] * fcb[2]));
            }
        }
    if (ACCUM)
        {
        float4 f = force[i];
        f.x += fx;
        f.y += fy;
        f.z += fz;
        f.w += pe;
        force[i] = f;
        if (virial)
#pragma unroll
            for (int q = 0; q < 6; q++)
                virial[q * virial_pitch + i] += v[q];
        }
    else
        {
        force[i] = make_float4(fx, fy, fz, pe);
        if (virial)
#pragma unroll
            for (int q = 0; q < 6; q++)
                virial[q * virial_pitch + i] = v[q];
        }
    }

hb2_status launch_angle_forces(hb2_bond_kind kind, const BondArgs& a, const uint4* members, bool accumulate,
                               const hb2_box& box, cudaStream_t stream)
    {
    if (a.N == 0)
        return HB2_OK;
    const uint32_t blocks = (a.N + BOND_BLOCK - 1) / BOND_BLOCK;
#define HB2_ANGLE_LAUNCH(K, ACC)                                                                          \
    angle_force_kernel<K, ACC><<<blocks, BOND_BLOCK, 0, stream>>>(                                        \
        a.pos, a.tag, a.rtag, a.N, a.Ntot, a.head, members, a.params, boxd(box), a.force, a.virial,       \
        a.virial_pitch, a.skip_flag, a.err)
    if (kind == HB2_ANGLE_COSINESQ)
        {
        if (accumulate)
            HB2_ANGLE_LAUNCH(1, true);
        else
            HB2_ANGLE_LAUNCH(1, false);
        }
    else
        {
        if (accumulate)
            HB2_ANGLE_LAUNCH(0, true);
        else
            HB2_ANGLE_LAUNCH(0, false);
        }
#undef HB2_ANGLE_LAUNCH
    HB2_CHECK_LAUNCH("angle_force_kernel");
    return HB2_OK;
    }

// Dihedrals: KIND 0 Periodic (hoomd/md/HarmonicDihedralForceCompute.cc:104-356), KIND 1 OPLS
// (hoomd/md/OPLSDihedralForceCompute.cc:100-352, parameters k_i / 2),
// owner-computes like the angles: table entries hold the four member tags and `type | role << 30`;
// the quadruplet is evaluated in fp64 with the reference's arithmetic and the particle keeps its share.
template<int KIND, bool ACCUM>
__global__ void __launch_bounds__(BOND_BLOCK)
    dihedral_force_kernel(const float4* __restrict__ pos, const uint32_t* __restrict__ tag,
                          const uint32_t* __restrict__ rtag, uint32_t N, uint32_t Ntot,
                          const uint32_t* __restrict__ head, const uint4* __restrict__ members,
                          const uint32_t* __restrict__ type_role, const float* __restrict__ params,
                          BoxD box, float4* __restrict__ force, float* __restrict__ virial,
                          size_t virial_pitch, const uint32_t* __restrict__ skip_flag,
                          uint32_t* __restrict__ err)
    {
    if (skip_flag && *skip_flag)
        return;
    const uint32_t i = blockIdx.x * BOND_BLOCK + threadIdx.x;
    if (i >= N)
        return;
    const uint32_t t = tag[i];
    const uint32_t k0 = head[t], k1 = head[t + 1];
    if (ACCUM && k0 == k1)
        return;
    float fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (uint32_t k = k0; k < k1; k++)
        {
        const uint4 m = members[k];
        const uint32_t tr = type_role[k];
        const uint32_t ia = rtag[m.x], ib = rtag[m.y], ic = rtag[m.z], id = rtag[m.w];
        if (ia >= Ntot || ib >= Ntot || ic >= Ntot || id >= Ntot)
            {
            atomicMax(err, t + 1u);
            continue;
            }
        const uint32_t role = tr >> 30, type = tr & 0x3fffffffu;
        const float4 pa = pos[ia], pb = pos[ib], pc = pos[ic], pd = pos[id];
        double dab[3] = {(double)pa.x - (double)pb.x, (double)pa.y - (double)pb.y, (double)pa.z - (double)pb.z};
        double dcb[3] = {(double)pc.x - (double)pb.x, (double)pc.y - (double)pb.y, (double)pc.z - (double)pb.z};
        double ddc[3] = {(double)pd.x - (double)pc.x, (double)pd.y - (double)pc.y, (double)pd.z - (double)pc.z};
        min_image_exact(box, dab[0], dab[1], dab[2]);
        min_image_exact(box, dcb[0], dcb[1], dcb[2]);
        min_image_exact(box, ddc[0], ddc[1], ddc[2]);
        const float* prm = params + HB2_BOND_PARAM_FLOATS * type;
        const double prm_d[4] = {(double)prm[0], (double)prm[1], (double)prm[2], (double)prm[3]};

        double dcbm[3] = {-dcb[0], -dcb[1], -dcb[2]};
        min_image_exact(box, dcbm[0], dcbm[1], dcbm[2]);
        const double aax = dab[1] * dcbm[2] - dab[2] * dcbm[1];
        const double aay = dab[2] * dcbm[0] - dab[0] * dcbm[2];
        const double aaz = dab[0] * dcbm[1] - dab[1] * dcbm[0];
        const double bbx = ddc[1] * dcbm[2] - ddc[2] * dcbm[1];
        const double bby = ddc[2] * dcbm[0] - ddc[0] * dcbm[2];
        const double bbz = ddc[0] * dcbm[1] - ddc[1] * dcbm[0];
        const double raasq = aax * aax + aay * aay + aaz * aaz;
        const double rbbsq = bbx * bbx + bby * bby + bbz * bbz;
        const double rgsq = dcbm[0] * dcbm[0] + dcbm[1] * dcbm[1] + dcbm[2] * dcbm[2];
        const double rg = sqrt(rgsq);
        double rginv = 0.0, raa2inv = 0.0, rbb2inv = 0.0;
        if (rg > 0.0)
            rginv = 1.0 / rg;
        if (raasq > 0.0)
            raa2inv = 1.0 / raasq;
        if (rbbsq > 0.0)
            rbb2inv = 1.0 / rbbsq;
        const double rabinv = sqrt(raa2inv * rbb2inv);
        double c_abcd = (aax * bbx + aay * bby + aaz * bbz) * rabinv;
        const double s_abcd = rg * rabinv * (aax * ddc[0] + aay * ddc[1] + aaz * ddc[2]);
        if (c_abcd > 1.0)
            c_abcd = 1.0;
        if (c_abcd < -1.0)
            c_abcd = -1.0;
        double df, eng;
        if (KIND == 1)
            {
            // hoomd/md/OPLSDihedralForceCompute.cc:236-279: p = sum_i k_i/2 (1 + (-1)^(i+1) cos(i phi)),
            // df = dp/dphi through the Chebyshev-like recursion; the halved k_i are the parameters
            const double k1 = prm_d[0], k2 = prm_d[1], k3 = prm_d[2], k4 = prm_d[3];
            double ddf1 = c_abcd, df1 = s_abcd, cos_term = ddf1;
            double p = k1 * (1.0 + cos_term);
            df = k1 * df1;
            ddf1 = cos_term * c_abcd - df1 * s_abcd;
            df1 = cos_term * s_abcd + df1 * c_abcd;
            cos_term = ddf1;
            p += k2 * (1.0 - cos_term);
            df += -2.0 * k2 * df1;
            ddf1 = cos_term * c_abcd - df1 * s_abcd;
            df1 = cos_term * s_abcd + df1 * c_abcd;
            cos_term = ddf1;
            p += k3 * (1.0 + cos_term);
            df += 3.0 * k3 * df1;
            ddf1 = cos_term * c_abcd - df1 * s_abcd;
            df1 = cos_term * s_abcd + df1 * c_abcd;
            cos_term = ddf1;
            p += k4 * (1.0 - cos_term);
            df += -4.0 * k4 * df1;
            eng = 0.25 * p;
            }
        else
            {
            const double K = prm_d[0], sign = prm_d[1], phi_0 = prm_d[3];
            const int multi = (int)prm_d[2];
            double p = 1.0, dfab = 0.0, ddfab = 0.0;
            for (int j = 0; j < multi; j++)
                {
                ddfab = p * c_abcd - dfab * s_abcd;
                dfab = p * s_abcd + dfab * c_abcd;
                p = ddfab;
                }
            const double sin_phi_0 = sin(phi_0), cos_phi_0 = cos(phi_0);
            p = p * cos_phi_0 + dfab * sin_phi_0;
            p = p * sign;
            dfab = dfab * cos_phi_0 - ddfab * sin_phi_0;
            dfab = dfab * sign;
            dfab *= (double)-multi;
            p += 1.0;
            if (multi == 0)
                {
                p = 1.0 + sign;
                dfab = 0.0;
                }
            df = -K * dfab * 0.500;
            eng = p * K * 0.125;
            }
        const double fg = dab[0] * dcbm[0] + dab[1] * dcbm[1] + dab[2] * dcbm[2];
        const double hg = ddc[0] * dcbm[0] + ddc[1] * dcbm[1] + ddc[2] * dcbm[2];
        const double fga = fg * raa2inv * rginv, hgb = hg * rbb2inv * rginv;
        const double gaa = -raa2inv * rg, gbb = rbb2inv * rg;
        const double dtf[3] = {gaa * aax, gaa * aay, gaa * aaz};
        const double dtg[3] = {fga * aax - hgb * bbx, fga * aay - hgb * bby, fga * aaz - hgb * bbz};
        const double dth[3] = {gbb * bbx, gbb * bby, gbb * bbz};
        double ffa[3], ffb[3], ffc[3], ffd[3];
        for (int q = 0; q < 3; q++)
            {
            const double s2 = df * dtg[q];
            ffa[q] = df * dtf[q];
            ffb[q] = s2 - ffa[q];
            ffd[q] = df * dth[q];
            ffc[q] = -s2 - ffd[q];
            }
        const double vir[6] = {(1. / 4.) * (dab[0] * ffa[0] + dcb[0] * ffc[0] + (ddc[0] + dcb[0]) * ffd[0]),
                               (1. / 4.) * (dab[1] * ffa[0] + dcb[1] * ffc[0] + (ddc[1] + dcb[1]) * ffd[0]),
                               (1. / 4.) * (dab[2] * ffa[0] + dcb[2] * ffc[0] + (ddc[2] + dcb[2]) * ffd[0]),
                               (1. / 4.) * (dab[1] * ffa[1] + dcb[1] * ffc[1] + (ddc[1] + dcb[1]) * ffd[1]),
                               (1. / 4.) * (dab[2] * ffa[1] + dcb[2] * ffc[1] + (ddc[2] + dcb[2]) * ffd[1]),
                               (1. / 4.) * (dab[2] * ffa[2] + dcb[2] * ffc[2] + (ddc[2] + dcb[2]) * ffd[2])};

        const double* mine = role == 0 ? ffa : (role == 1 ? ffb : (role == 2 ? ffc : ffd));
        fx += (float)mine[0];
        fy += (float)mine[1];
        fz += (float)mine[2];
        pe += (float)eng;
        if (virial)
#pragma unroll
            for (int q = 0; q < 6; q++)
                v[q] += (float)vir[q];
        }
    if (ACCUM)
        {
        float4 f = force[i];
        f.x += fx;
        f.y += fy;
        f.z += fz;
        f.w += pe;
        force[i] = f;
        if (virial)
#pragma unroll
            for (int q = 0; q < 6; q++)
                virial[q * virial_pitch + i] += v[q];
        }
    else
        {
        force[i] = make_float4(fx, fy, fz, pe);
        if (virial)
#pragma unroll
            for (int q = 0; q < 6; q++)
                virial[q * virial_pitch + i] = v[q];
        }
    }

hb2_status launch_dihedral_forces(hb2_bond_kind kind, const BondArgs& a, const uint4* members,
                                  const uint32_t* type_role, bool accumulate, const hb2_box& box,
                                  cudaStream_t stream)
    {
    if (a.N == 0)
        return HB2_OK;
    const uint32_t blocks = (a.N + BOND_BLOCK - 1) / BOND_BLOCK;
#define HB2_DIHEDRAL_LAUNCH(K, ACC)                                                                       \
    dihedral_force_kernel<K, ACC><<<blocks, BOND_BLOCK, 0, stream>>>(                                     \
        a.pos, a.tag, a.rtag, a.N, a.Ntot, a.head, members, type_role, a.params, boxd(box), a.force,      \
        a.virial, a.virial_pitch, a.skip_flag, a.err)
    if (kind == HB2_DIHEDRAL_OPLS)
        {
        if (accumulate)
            HB2_DIHEDRAL_LAUNCH(1, true);
        else
            HB2_DIHEDRAL_LAUNCH(1, false);
        }
    else
        {
        if (accumulate)
            HB2_DIHEDRAL_LAUNCH(0, true);
        else
            HB2_DIHEDRAL_LAUNCH(0, false);
        }
#undef HB2_DIHEDRAL_LAUNCH
    HB2_CHECK_LAUNCH("dihedral_force_kernel");
    return HB2_OK;
    }

// One warp per row: entries whose particle is bonded to the row's owner are dropped, the row is
// compacted in place (a 32-entry chunk is read completely before the chunk's survivors are written at
// or below its start) and padded again with the dummy index up to a multiple of 32.
__global__ void __launch_bounds__(128)
    nlist_filter_bonds_kernel(uint32_t N, const uint32_t* __restrict__ tag,
                              const uint32_t* __restrict__ rtag, const uint32_t* __restrict__ head,
                              const uint32_t* __restrict__ partner, uint32_t* __restrict__ nlist,
                              uint32_t pitch, uint32_t* __restrict__ n_neigh, uint32_t dummy_index)
    {
    const uint32_t i = (blockIdx.x * 128 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= N)
        return;
    const uint32_t t = tag[i];
    const uint32_t k0 = head[t], n_ex = head[t + 1] - k0;
    if (n_ex == 0)
        return;
    uint32_t* row = nlist + (size_t)i * pitch;
    const uint32_t n = n_neigh[i];
    uint32_t o = 0;
    for (uint32_t c = 0; c < n; c += 32)
        {
        const uint32_t e = c + lane < n ? row[c + lane] : 0xffffffffu;
        bool excluded = false;
        for (uint32_t x0 = 0; x0 < n_ex; x0 += 32)
            {
            const uint32_t ex = x0 + lane < n_ex ? rtag[partner[k0 + x0 + lane]] : 0xfffffffeu;
            const uint32_t m = min(32u, n_ex - x0);
            for (uint32_t l = 0; l < m; l++)
                excluded |= (e == __shfl_sync(0xffffffffu, ex, l));
            }
        const bool keep = (c + lane < n) && !excluded;
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep)
            row[o + __popc(bal & ((1u << lane) - 1u))] = e;
        o += __popc(bal);
        __syncwarp();
        }
    const uint32_t end = min((o + 31u) & ~31u, pitch);
    const uint32_t old_end = min((n + 31u) & ~31u, pitch);
    for (uint32_t k = o + lane; k < max(end, old_end); k += 32)
        if (k < end)
            row[k] = dummy_index;
    if (lane == 0)
        n_neigh[i] = o;
    }

hb2_status launch_nlist_filter_bonds(uint32_t N, const uint32_t* tag, const uint32_t* rtag,
                                     const uint32_t* head, const uint32_t* partner, uint32_t* nlist,
                                     uint32_t pitch, uint32_t* n_neigh, uint32_t dummy_index,
                                     cudaStream_t stream)
    {
    if (N == 0)
        return HB2_OK;
    const uint32_t blocks = (uint32_t)(((uint64_t)N * 32 + 127) / 128);
    nlist_filter_bonds_kernel<<<blocks, 128, 0, stream>>>(N, tag, rtag, head, partner, nlist, pitch,
                                                         n_neigh, dummy_index);
    HB2_CHECK_LAUNCH("nlist_filter_bonds_kernel");
    return HB2_OK;
    }
    } // namespace hb2
