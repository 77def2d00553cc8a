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

// ---- fp64 3-vectors for the bonded geometry (a few triplets / quadruplets per particle)
struct V3
    {
    double x, y, z;
    };
HB2_DEVICE V3 v3(double x, double y, double z)
    {
    V3 r = {x, y, z};
    return r;
    }
HB2_DEVICE V3 operator+(V3 a, V3 b)
    {
    return v3(a.x + b.x, a.y + b.y, a.z + b.z);
    }
HB2_DEVICE V3 operator-(V3 a, V3 b)
    {
    return v3(a.x - b.x, a.y - b.y, a.z - b.z);
    }
HB2_DEVICE V3 operator*(double s, V3 a)
    {
    return v3(s * a.x, s * a.y, s * a.z);
    }
HB2_DEVICE double dot(V3 a, V3 b)
    {
    return a.x * b.x + a.y * b.y + a.z * b.z;
    }
HB2_DEVICE V3 cross(V3 a, V3 b)
    {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
    }
// minimum-image separation p - q of two stored positions
HB2_DEVICE V3 sep(const BoxD& box, const float4& p, const float4& q)
    {
    double x = (double)p.x - (double)q.x, y = (double)p.y - (double)q.y, z = (double)p.z - (double)q.z;
    min_image_exact(box, x, y, z);
    return v3(x, y, z);
    }
// upper triangle (xx, xy, xz, yy, yz, zz) of r (x) f added to w with weight s
HB2_DEVICE void add_outer(double (&w)[6], double s, V3 r, V3 f)
    {
    w[0] += s * r.x * f.x;
    w[1] += s * r.x * f.y;
    w[2] += s * r.x * f.z;
    w[3] += s * r.y * f.y;
    w[4] += s * r.y * f.z;
    w[5] += s * r.z * f.z;
    }

// Angle forces, owner-computes: particle i walks the angles it is a member of (table keyed by tag:
// the three member tags and `type | role << 30`), evaluates the triplet and keeps its own share (role
// 0: a, 1: the vertex b, 2: c). Same potentials and conventions as the reference (KIND 0:
// hoomd/md/HarmonicAngleForceCompute.cc:104-258, V = K/2 (theta - t0)^2 with sin(theta) clamped at
// 0.001; KIND 1: hoomd/md/CosineSqAngleForceCompute.cc, V = K/2 (cos theta - cos t0)^2; a third of
// the energy and of the virial to each member), written as a torque: with u = r_a - r_b and
// w = r_c - r_b, the part of w^ perpendicular to u^ is p_a = w^ - cos(theta) u^ (length sin theta,
// in the plane, pointing from u towards w), and the force on the end particle a is
//     F_a = V'(theta) / (|u| sin theta) p_a        (harmonic)
//     F_a = -dV/dcos(theta) / |u| p_a               (cosine squared: no division by sin theta)
// likewise F_c with u and w exchanged; F_b = -(F_a + F_c); W = (u (x) F_a + w (x) F_c) / 3.
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
        // fp64: near 0 and 180 degrees sin(theta) and acos lose half of fp32's digits (chains with
        // t0 = pi sit exactly there), and there are only a few angles per particle
        const V3 u = sep(box, pa, pb), w = sep(box, pc, pb);
        const double lu = sqrt(dot(u, u)), lw = sqrt(dot(w, w));
        const V3 uh = (1.0 / lu) * u, wh = (1.0 / lw) * w;
        const double c = fmin(1.0, fmax(-1.0, dot(uh, wh)));
        const V3 p_a = wh - c * uh, p_c = uh - c * wh;
        const double K = (double)params[HB2_BOND_PARAM_FLOATS * type],
                     t_0 = (double)params[HB2_BOND_PARAM_FLOATS * type + 1];
        double ga, gc, energy; // F_a = ga p_a, F_c = gc p_c
        if (KIND == 0)
            {
            const double s = fmax(sqrt(1.0 - c * c), 0.001);
            const double dth = acos(c) - t_0;
            ga = K * dth / (lu * s);
            gc = K * dth / (lw * s);
            energy = 0.5 * K * dth * dth;
            }
        else
            {
            const double dc = c - cos(t_0);
            ga = -K * dc / lu;
            gc = -K * dc / lw;
            energy = 0.5 * K * dc * dc;
            }
        const V3 Fa = ga * p_a, Fc = gc * p_c;
        const V3 mine = role == 0 ? Fa : (role == 2 ? Fc : v3(0, 0, 0) - (Fa + Fc));
        fx += (float)mine.x;
        fy += (float)mine.y;
        fz += (float)mine.z;
        pe += (float)(energy * (1.0 / 3.0));
        if (virial)
            {
            double wv[6] = {0, 0, 0, 0, 0, 0};
            add_outer(wv, 1.0 / 3.0, u, Fa);
            add_outer(wv, 1.0 / 3.0, w, Fc);
#pragma unroll
            for (int q = 0; q < 6; q++)
                v[q] += (float)wv[q];
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

// Dihedrals, owner-computes like the angles (table entries hold the four member tags and
// `type | role << 30`). Same potentials and conventions as the reference — KIND 0 Periodic
// V = k/2 (1 + d cos(n phi - phi0)) (hoomd/md/HarmonicDihedralForceCompute.cc:104-356), KIND 1 OPLS
// V = 1/2 [k1 (1 + cos phi) + k2 (1 - cos 2 phi) + k3 (1 + cos 3 phi) + k4 (1 - cos 4 phi)]
// (hoomd/md/OPLSDihedralForceCompute.cc:100-352); a quarter of the energy and of the virial to each
// member — evaluated from the angle itself: with F = r_a - r_b, G = r_b - r_c, H = r_d - r_c, A = F x G
// and B = H x G,
//     phi = atan2((B x A) . G / |G|, A . B),
// V and dV/dphi come straight from cos / sin of the multiples of phi, and the forces are -dV/dphi times
// the gradients of phi (Blondel and Karplus, J. Comput. Chem. 17, 1132 (1996)):
//     dphi/dr_a = -|G| A / A^2,  dphi/dr_d = |G| B / B^2,
//     dphi/dr_b = -dphi/dr_a + S,  dphi/dr_c = -dphi/dr_d - S,  S = (F.G) A / (A^2 |G|) - (H.G) B / (B^2 |G|).
// The virial is sum_i (r_i - r_b) (x) F_i (torque-free, hence symmetric).
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
        const V3 F = sep(box, pa, pb), G = sep(box, pb, pc), H = sep(box, pd, pc);
        const V3 A = cross(F, G), B = cross(H, G);
        const double A2 = dot(A, A), B2 = dot(B, B), lG = sqrt(dot(G, G));
        const double phi = atan2(dot(cross(B, A), G) / lG, dot(A, B));
        const float* prm = params + HB2_BOND_PARAM_FLOATS * type;
        double energy, dE; // V(phi), dV/dphi
        if (KIND == 0)
            {
            const double k = (double)prm[0], d = (double)prm[1], n = (double)prm[2], phi0 = (double)prm[3];
            double sn, cn;
            sincos(n * phi - phi0, &sn, &cn);
            energy = 0.5 * k * (1.0 + d * cn);
            dE = -0.5 * k * d * n * sn;
            }
        else
            {
            // (the host packs the OPLS coefficients as k_i / 2, like hoomd/md/OPLSDihedralForceCompute.cc:73-78)
            const double k1 = 2.0 * (double)prm[0], k2 = 2.0 * (double)prm[1], k3 = 2.0 * (double)prm[2],
                         k4 = 2.0 * (double)prm[3];
            double s1, c1, s2, c2, s3, c3, s4, c4;
            sincos(phi, &s1, &c1);
            sincos(2.0 * phi, &s2, &c2);
            sincos(3.0 * phi, &s3, &c3);
            sincos(4.0 * phi, &s4, &c4);
            energy = 0.5 * (k1 * (1.0 + c1) + k2 * (1.0 - c2) + k3 * (1.0 + c3) + k4 * (1.0 - c4));
            dE = 0.5 * (-k1 * s1 + 2.0 * k2 * s2 - 3.0 * k3 * s3 + 4.0 * k4 * s4);
            }
        const V3 ga = (-lG / A2) * A, gd = (lG / B2) * B;
        const V3 S = (dot(F, G) / (A2 * lG)) * A - (dot(H, G) / (B2 * lG)) * B;
        const V3 Fa = (-dE) * ga, Fd = (-dE) * gd;
        const V3 Fb = (-dE) * (S - ga), Fc = dE * (gd + S);
        const V3 mine = role == 0 ? Fa : (role == 1 ? Fb : (role == 2 ? Fc : Fd));
        fx += (float)mine.x;
        fy += (float)mine.y;
        fz += (float)mine.z;
        pe += (float)(0.25 * energy);
        if (virial)
            {
            // relative to b: r_a - r_b = F, r_c - r_b = -G, r_d - r_b = H - G
            double wv[6] = {0, 0, 0, 0, 0, 0};
            add_outer(wv, 0.25, F, Fa);
            add_outer(wv, -0.25, G, Fc);
            add_outer(wv, 0.25, H - G, Fd);
#pragma unroll
            for (int q = 0; q < 6; q++)
                v[q] += (float)wv[q];
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
