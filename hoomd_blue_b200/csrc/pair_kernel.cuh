// pair_kernel.cuh — PotentialPair<Evaluator> force / energy / virial pass over a full neighbour list.
//
// Replaces PotentialPair<ev>::computeForces (hoomd/md/PotentialPair.h:587-794) and
// gpu_compute_pair_forces_shared_kernel (hoomd/md/PotentialPairGPU.cuh:146-404), and the DPD
// variants (hoomd/md/PotentialPairDPDThermo.h:102-268, hoomd/md/PotentialPairDPDThermoGPU.cuh:148-330).
//
// Layout contract with the build kernel (nlist.cu): row i starts at nlist + i * pitch, pitch is a
// multiple of 32, entries are ascending in j and the row is padded to a multiple of 32 with the index
// of a far-away dummy particle (pos[N_total]), so the loop has no tail handling and never tests
// k < n_neigh.
//
// TPP lanes share one particle. The rows of a warp's particles are staged in shared memory by one
// bulk copy (cp.async.bulk, TMA 1D) per row completing on a per-warp mbarrier: the streaming read of
// the list bypasses the LSU wavefront queue that the position gathers saturate (ncu r01: 79 % of the
// L1 data pipe, a quarter of it index loads) and its DRAM latency is paid once per warp.
// One loop iteration consumes a 32-entry chunk of the row: for u = 0 .. 32/TPP-1 lane `sub` takes
// entry u*TPP + sub (a conflict-free 32-bit shared load), and — because rows are ascending and the
// particles are in cell order — the TPP position gathers issued by one instruction fall on TPP nearly
// consecutive particles, i.e. on one or two 128-byte lines. All 32/TPP gathers are issued before the
// first pair is evaluated (memory-level parallelism).
// Owner-computes on a full list: no atomics, fixed __shfl_xor reduction tree, bitwise reproducible.
// Each visit adds E/2 and (F/r) dx_a dx_b / 2 to the visiting particle only
// (hoomd/md/PotentialPairGPU.cuh:365-370, hoomd/md/PotentialPair.h:739-752).
#pragma once
#include "common.cuh"
#include "evaluators.cuh"
#include "kernels.cuh"
#include "tile.cuh"
#include <algorithm>
#include <vector>

namespace hb2
    {
constexpr int PAIR_BLOCK = 128;
constexpr int PAIR_MAX_TYPES_SMEM = 16; // ntypes^2 <= 256 entries staged in shared memory
constexpr int PAIR_STAGE_MAX = 8;       // 32-entry chunks of a row staged in shared memory per pass
// 10 resident blocks per SM (48 registers): with the packed arithmetic the kernel is latency-bound on
// its gathers and 10 blocks measure 7 % faster than 8 (0.193 vs 0.207 ms at C2)
#ifndef PAIR_MIN_BLOCKS
#define PAIR_MIN_BLOCKS 10
#endif
// (the virial / thermostatted / multi-type instantiations carry more state: 8 blocks, 64 registers)
#define PAIR_BOUNDS \
    __launch_bounds__(PAIR_BLOCK, (VIRIAL || Ev::needs_thermo || !ONETYPE || TPP < 8) ? 8 : PAIR_MIN_BLOCKS)
constexpr int PAIR_ROW_PAD = 8;         // words of padding per staged row (bank spread across rows)

// ---- bulk-copy (TMA 1D) row staging: one cp.async.bulk per neighbour-list row, completion on a
// per-warp mbarrier. The row never passes through registers or the LSU wavefront queue, and the DRAM
// latency of the list is paid once per warp instead of once per 32-entry chunk.
HB2_DEVICE uint32_t smem_addr(const void* p)
    {
    return (uint32_t)__cvta_generic_to_shared(p);
    }
HB2_DEVICE void mbar_init(uint64_t* bar, uint32_t count)
    {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
HB2_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
    {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
    }
HB2_DEVICE void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
    {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
    }
HB2_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity)
    {
    uint32_t done;
    do
        {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(parity)
                     : "memory");
        } while (!done);
    }

template<class Ev> struct PairKArgs
    {
    const float4* pos;
    const float4* vel;
    const uint32_t* tag;
    uint32_t N;
    const uint32_t* n_neigh;
    const uint32_t* nlist;
    uint32_t pitch;
    const float* params; // HB2_PARAM_FLOATS floats per type pair
    const float* rcutsq;
    const float* ronsq;
    typename Ev::param_type param0;
    float rcutsq0, ronsq0;
    const double* rcutsq_d; // fp64 r_cut^2 per type pair (may be null) and for type pair 0
    double rcutsq0_d;
    uint32_t ntypes;
    float4* force;
    float* virial;
    size_t virial_pitch;
    uint32_t seed;
    const uint64_t* d_timestep;
    uint64_t timestep;
    float dt, kT;
    float face_margin; // >= r_cut_max: farther than this from every face => no pair inside the cut-off crosses a face
    uint32_t stage;    // chunks of a row staged per pass (<= PAIR_STAGE_MAX)
    bool energy;       // accumulate the per-particle potential energy
    const uint32_t* skip_flag; // non-zero word => return at once (speculative launch, engine.cu)
    BoxF box;
    BoxD boxd;
    // tile-shared list (tile.cuh / pair_tile.cuh); rows16 == nullptr: per-particle rows of global indices
    const uint16_t* rows16;
    TileGeom tg;
    DomainGeom geom;
    const uint32_t* cell_start;
    const uint32_t* cell_end;
    uint32_t stage_cap;
    uint32_t tile_flags;            // bit 0: L2 prefetch of the rows, bit 1: of the positions (producer warp)
    const unsigned char* tile_desc; // per-tile descriptors written by the list build (tile.cuh)
    uint32_t* tile_ticket; // [0] next tile, [1] blocks that ran out (pair_tile.cuh); zero between launches
    };

// MODE: 0 none, 1 shift, 2 xplor (hoomd/md/PotentialPair.h:104-109)
// ENERGY = false skips the per-particle potential energy (force.w is written as 0): the engine asks
// for it only on steps whose energies can be observed (the last step of a run, or lazily when a getter
// is called), like the reference does for the virial through PDataFlags (hoomd/ParticleData.h:70-84).
template<class Ev, int MODE, bool VIRIAL, int TPP, bool ONETYPE, bool ENERGY>
__global__ void PAIR_BOUNDS pair_force_kernel(const PairKArgs<Ev> a)
    {
    typedef typename Ev::param_type param_type;
    constexpr int U = 32 / TPP; // entries per lane per 32-entry chunk
    if (a.skip_flag && *a.skip_flag)
        return;
    __shared__ param_type s_params[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ float s_rcutsq[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ float s_ronsq[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ __align__(8) uint64_t s_bar[PAIR_BLOCK / 32];
    extern __shared__ __align__(128) uint32_t s_rows[]; // PAIR_BLOCK / TPP rows of stage*32 + pad
    const int lane = threadIdx.x & 31;
    uint64_t* bar = &s_bar[threadIdx.x >> 5];
    if (lane == 0)
        mbar_init(bar, 1);
    __syncwarp();
    if (!ONETYPE)
        {
        const uint32_t nt2 = a.ntypes * a.ntypes;
        for (uint32_t t = threadIdx.x; t < nt2; t += blockDim.x)
            {
            s_params[t] = *reinterpret_cast<const param_type*>(a.params + HB2_PARAM_FLOATS * t);
            s_rcutsq[t] = a.rcutsq[t];
            s_ronsq[t] = a.ronsq[t];
            }
        __syncthreads();
        }

    const uint32_t gid = blockIdx.x * PAIR_BLOCK + threadIdx.x;
    const uint32_t i = gid / TPP;
    const uint32_t sub = gid % TPP;
    const bool active = i < a.N;

    float fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
    float vxx = 0.f, vxy = 0.f, vxz = 0.f, vyy = 0.f, vyz = 0.f, vzz = 0.f;
    // Packed path (Blackwell FFMA2/FMUL2/FADD2: two pairs per instruction): one-type, non-xplor,
    // non-thermostatted evaluators that provide evalPacked. Accumulators hold (even, odd) pair sums.
    constexpr bool PACKED = Ev::has_packed && ONETYPE && MODE != 2 && !Ev::needs_thermo && (U % 2 == 0);
    float2 fx2 = make_float2(0.f, 0.f), fy2 = fx2, fz2 = fx2, pe2 = fx2;
    float2 wxx = fx2, wxy = fx2, wxz = fx2, wyy = fx2, wyz = fx2, wzz = fx2;

    // Image handling without per-pair fp32 cancellation error: a particle farther than
    // r_list + buffer from every face has all its listed neighbours in the same image, so the plain
    // difference pi - pj is the minimum image (and is exact or correctly rounded). A warp that holds
    // only such particles runs the shift-free loop. Otherwise the particle carries pi - L and pi + L
    // as unevaluated (hi, lo) float pairs (from fp64), and a pair across a face uses
    // (hi - pj) + lo: the first difference is exact (Sterbenz), so dx is correctly rounded instead
    // of carrying half an ulp of L.
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
    bool near_face = false;
    if (active)
        {
        pi = __ldg(a.pos + i);
        // a non-positive margin disables the shortcut (every particle takes the shifting loop)
        const float m = a.face_margin > 0.f ? a.face_margin : 3.0e38f;
        near_face = (pi.x >= a.box.hi[0] - m) || (pi.x < a.box.lo[0] + m)
                    || (pi.y >= a.box.hi[1] - m) || (pi.y < a.box.lo[1] + m)
                    || (pi.z >= a.box.hi[2] - m) || (pi.z < a.box.lo[2] + m);
        }
    const bool warp_shift = __any_sync(0xffffffffu, near_face);

    // ---- per-particle set-up (inactive lanes of the last warp carry n = 0 and only take part in the
    // warp-wide staging protocol)
    const uint32_t ti = ONETYPE ? 0u : __float_as_uint(pi.w);
    const uint32_t n = active ? __ldg(a.n_neigh + i) : 0u;
    const uint32_t nchunks = (n + 31u) >> 5;
    float4 vi = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t tagi = 0;
    uint64_t timestep = 0;
    if (Ev::needs_thermo && active)
        {
        vi = __ldg(a.vel + i);
        tagi = __ldg(a.tag + i);
        timestep = a.d_timestep ? *a.d_timestep : a.timestep;
        }

    // The cut-off decision of the reference is made in fp64 (rsq < rcutsq with Scalar = double), and
    // an unshifted LJ pair flipped at the cut-off costs 8e-4 of the rms force. The hot loop stays
    // branch-free: it decides in fp32 and only notes (two integer instructions per pair) whether any
    // rsq came within 32 ulp of r_cut^2 — the separations are exact or correctly rounded, so rsq
    // carries < 2 ulp. Such a pass (about one in 1e5) is then re-walked by `fix_pass`, which decides
    // the near pairs with the CPU build's exact arithmetic (dist_sq_exact) and adds or removes the
    // contribution of the pairs where fp32 and the reference disagree.
    // near_min = min over the pass of (bits(rsq) - bits(r_cut^2) + 32) as unsigned: <= 64 iff some rsq
    // came within 32 ulp of the cut-off (one VIADDMNMX per pair).
    uint32_t near_min = 0xffffffffu;
    auto note_near = [&](const float rsq, const float rcutsq)
    {
#ifndef PAIR_NO_EXACT
        near_min = min(near_min, (uint32_t)(__float_as_int(rsq) + (32 - __float_as_int(rcutsq))));
#endif
    };

    // one pair; sgn = +1 adds it, -1 removes it (fix_pass). rsq is passed in so that the caller
    // controls which side of the cut-off the evaluator sees.
    auto pair_body = [&](const uint32_t j, const float4 pj, const float dx, const float dy,
                         const float dz, const float rsq, const float sgn)
    {
        float rcutsq = a.rcutsq0, ronsq = a.ronsq0;
        const param_type* prm = &a.param0;
        if (!ONETYPE)
            {
            // Index2D(ntypes)(ti, tj) = tj * ntypes + ti (hoomd/Index1D.h:53)
            const uint32_t tp = __float_as_uint(pj.w) * a.ntypes + ti;
            rcutsq = s_rcutsq[tp];
            ronsq = s_ronsq[tp];
            prm = &s_params[tp];
            }
        bool energy_shift = (MODE == 1);
        if (MODE == 2)
            energy_shift = ronsq > rcutsq; // hoomd/md/PotentialPair.h:697-700

        float force_divr = 0.f, force_divr_cons = 0.f, pair_eng = 0.f;
        Ev eval(rsq, rcutsq, *prm);
        bool evaluated;
        if constexpr (Ev::needs_thermo)
            {
            const float4 vj = __ldg(a.vel + j);
            const uint32_t tagj = __ldg(a.tag + j);
            const float rdotv = dx * (vi.x - vj.x) + dy * (vi.y - vj.y) + dz * (vi.z - vj.z);
            eval.set_seed_ij_timestep((uint16_t)a.seed, tagi, tagj, timestep);
            eval.setDeltaT(a.dt);
            eval.setRDotV(rdotv);
            eval.setT(a.kT);
            evaluated = eval.evalForceEnergyThermo(force_divr, force_divr_cons, pair_eng,
                                                   energy_shift);
            }
        else
            {
            evaluated = eval.evalForceAndEnergy(force_divr, pair_eng, energy_shift);
            force_divr_cons = force_divr;
            }
        if (MODE == 2)
            {
            // hoomd/md/PotentialPair.h:711-737
            if (evaluated && rsq >= ronsq && rsq < rcutsq)
                {
                const float old_eng = pair_eng, old_f = force_divr;
                const float d = rcutsq - ronsq;
                const float denom_inv = 1.0f / (d * d * d);
                const float m = rsq - rcutsq;
                const float s = m * m * (rcutsq + 2.0f * rsq - 3.0f * ronsq) * denom_inv;
                const float ds = 12.0f * (rsq - ronsq) * m * denom_inv;
                pair_eng = old_eng * s;
                force_divr = s * old_f - ds * old_eng;
                force_divr_cons = force_divr;
                }
            }
        // branch-free accumulation: a pair beyond the cut-off (or the row padding) adds zeros
        force_divr = evaluated ? sgn * force_divr : 0.f;
        pair_eng = evaluated ? sgn * pair_eng : 0.f;
        fx += dx * force_divr;
        fy += dy * force_divr;
        fz += dz * force_divr;
        if (ENERGY)
            pe += pair_eng * 0.5f;
        if (VIRIAL)
            {
            const float f2 = evaluated ? 0.5f * sgn * force_divr_cons : 0.f;
            vxx += f2 * dx * dx;
            vxy += f2 * dx * dy;
            vxz += f2 * dx * dz;
            vyy += f2 * dy * dy;
            vyz += f2 * dy * dz;
            vzz += f2 * dz * dz;
            }
    };

    // two neighbours at once; dx2 = (dx of pair 0, dx of pair 1) etc.
    auto pair_body2 = [&](const float2 dx2, const float2 dy2, const float2 dz2)
    {
        const float2 rsq = __ffma2_rn(dz2, dz2, __ffma2_rn(dy2, dy2, __fmul2_rn(dx2, dx2)));
        note_near(rsq.x, a.rcutsq0);
        note_near(rsq.y, a.rcutsq0);
        float2 fdr = make_float2(0.f, 0.f), eng = fdr;
        bool in0 = false, in1 = false;
        if constexpr (PACKED)
            Ev::evalPacked(rsq, a.rcutsq0, a.param0, MODE == 1, ENERGY, fdr, eng, in0, in1);
        fdr.x = in0 ? fdr.x : 0.f;
        fdr.y = in1 ? fdr.y : 0.f;
        fx2 = __ffma2_rn(dx2, fdr, fx2);
        fy2 = __ffma2_rn(dy2, fdr, fy2);
        fz2 = __ffma2_rn(dz2, fdr, fz2);
        if (ENERGY)
            {
            eng.x = in0 ? eng.x : 0.f;
            eng.y = in1 ? eng.y : 0.f;
            pe2 = __ffma2_rn(eng, make_float2(0.5f, 0.5f), pe2);
            }
        if (VIRIAL)
            {
            const float2 h = __fmul2_rn(fdr, make_float2(0.5f, 0.5f));
            const float2 hx = __fmul2_rn(h, dx2), hy = __fmul2_rn(h, dy2), hz = __fmul2_rn(h, dz2);
            wxx = __ffma2_rn(hx, dx2, wxx);
            wxy = __ffma2_rn(hx, dy2, wxy);
            wxz = __ffma2_rn(hx, dz2, wxz);
            wyy = __ffma2_rn(hy, dy2, wyy);
            wyz = __ffma2_rn(hy, dz2, wyz);
            wzz = __ffma2_rn(hz, dz2, wzz);
            }
    };

    // (hi, lo) of the one image shift a neighbour of this particle can need per component: a particle
    // in the upper half of the box can only see neighbours across the upper face (pi - L), one in the
    // lower half only across the lower face (pi + L). Only the face-shifting loop uses them.
    float sh[3], sl[3];
    const float pc[3] = {pi.x, pi.y, pi.z};
    const float hc[3] = {a.box.hi[0], a.box.hi[1], a.box.hi[2]};
    if (warp_shift)
        {
#pragma unroll
        for (int c = 0; c < 3; c++)
            {
            const double m = pc[c] >= 0.f ? (double)pc[c] - a.boxd.L[c] : (double)pc[c] + a.boxd.L[c];
            sh[c] = (float)m;
            sl[c] = (float)(m - (double)sh[c]);
            }
        }

    // minimum-image separation pi - pj: the plain difference when no particle of the warp is near a
    // face, the (hi, lo) form otherwise
    auto separation_shift = [&](const float4 q, float& dx, float& dy, float& dz)
    {
        float d[3];
        const float qc[3] = {q.x, q.y, q.z};
#pragma unroll
        for (int c = 0; c < 3; c++)
            {
            const float d0 = pc[c] - qc[c];
            const float ds = (sh[c] - qc[c]) + sl[c];
            d[c] = fabsf(d0) >= hc[c] ? ds : d0;
            }
        dx = d[0];
        dy = d[1];
        dz = d[2];
    };
    auto separation = [&](const float4 q, float& dx, float& dy, float& dz)
    {
        if (!warp_shift)
            {
            dx = pi.x - q.x;
            dy = pi.y - q.y;
            dz = pi.z - q.z;
            }
        else
            separation_shift(q, dx, dy, dz);
    };
    // the one rounding sequence both the hot loop (scalar and packed) and fix_pass use
    auto rsq_of = [](const float dx, const float dy, const float dz)
    { return fmaf(dz, dz, fmaf(dy, dy, dx * dx)); };

    // ---- staged sweep over the row. Pass p brings chunks [p*stage, (p+1)*stage) of every row of the
    // warp into shared memory with one bulk copy per row; lanes then read their indices with
    // conflict-free 32-bit shared loads (row stride = stage*32 + 8 words).
    const uint32_t stage = a.stage;
    const uint32_t row_words = stage * 32u + PAIR_ROW_PAD;
    uint32_t* __restrict__ myrow = s_rows + (threadIdx.x / TPP) * row_words;
    const uint32_t* __restrict__ grow = a.nlist + (size_t)i * a.pitch;
    const uint32_t nch_warp = __reduce_max_sync(0xffffffffu, nchunks);
    // rare: re-walk the staged pass, decide the pairs near the cut-off exactly, repair disagreements
    auto fix_pass = [&](const uint32_t cnt)
    {
        for (uint32_t c = 0; c < cnt; c++)
            for (int u = 0; u < U; u++)
                {
                const uint32_t j = myrow[c * 32u + u * TPP + sub];
                const float4 q = __ldg(a.pos + j);
                float dx, dy, dz;
                separation(q, dx, dy, dz);
                const float rsq = rsq_of(dx, dy, dz);
                const uint32_t tp = ONETYPE ? 0u : __float_as_uint(q.w) * a.ntypes + ti;
                const float rcutsq = ONETYPE ? a.rcutsq0 : s_rcutsq[tp];
                if ((uint32_t)(__float_as_int(rsq) - __float_as_int(rcutsq) + 32) > 64u)
                    continue;
                const double rc2 = ONETYPE ? a.rcutsq0_d : (a.rcutsq_d ? a.rcutsq_d[tp] : (double)rcutsq);
                const bool in_ref = dist_sq_exact(a.boxd, pi, q) < rc2;
                const bool in_f32 = rsq < rcutsq;
                if (in_ref == in_f32)
                    continue;
                // add the pair just inside the cut-off, or take back what the hot loop added
                const float below = __uint_as_float(__float_as_uint(rcutsq) - 1u);
                pair_body(j, q, dx, dy, dz, in_ref ? below : rsq, in_ref ? 1.0f : -1.0f);
                }
    };

    uint32_t parity = 0;
    for (uint32_t c0 = 0; c0 < nch_warp; c0 += stage)
        {
        const uint32_t cnt = nchunks > c0 ? min(nchunks - c0, stage) : 0u;
        const uint32_t bytes = cnt * 128u;
        const uint32_t tot = __reduce_add_sync(0xffffffffu, sub == 0 ? bytes : 0u);
        if (c0 > 0)
            {
            // every lane is done reading the previous pass before the async proxy overwrites it
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
        if (lane == 0)
            mbar_expect_tx(bar, tot);
        __syncwarp();
        if (sub == 0 && bytes)
            bulk_copy_g2s(myrow, grow + c0 * 32u, bytes, bar);
        mbar_wait(bar, parity);
        parity ^= 1u;

        for (uint32_t c = 0; c < cnt; c++)
            {
            uint32_t j[U];
            float4 pj[U];
#pragma unroll
            for (int u = 0; u < U; u++)
                j[u] = myrow[c * 32u + u * TPP + sub];
#pragma unroll
            for (int u = 0; u < U; u++)
                pj[u] = __ldg(a.pos + j[u]);
            // separations of the U neighbours (scalar: each result lands in its own register, so
            // consecutive pairs form the float2 operands of the packed path without moves)
            float dxs[U], dys[U], dzs[U];
            if (!warp_shift) // hoisted: one warp-uniform branch per chunk, not per neighbour
                {
#pragma unroll
                for (int u = 0; u < U; u++)
                    {
                    dxs[u] = pi.x - pj[u].x;
                    dys[u] = pi.y - pj[u].y;
                    dzs[u] = pi.z - pj[u].z;
                    }
                }
            else
                {
#pragma unroll
                for (int u = 0; u < U; u++)
                    separation_shift(pj[u], dxs[u], dys[u], dzs[u]);
                }
            if constexpr (PACKED)
                {
#pragma unroll
                for (int u = 0; u < U; u += 2)
                    pair_body2(make_float2(dxs[u], dxs[u + 1]), make_float2(dys[u], dys[u + 1]),
                               make_float2(dzs[u], dzs[u + 1]));
                }
            else
                {
#pragma unroll
                for (int u = 0; u < U; u++)
                    {
                    const float rsq = rsq_of(dxs[u], dys[u], dzs[u]);
                    note_near(rsq, ONETYPE ? a.rcutsq0
                                           : s_rcutsq[__float_as_uint(pj[u].w) * a.ntypes + ti]);
                    pair_body(j[u], pj[u], dxs[u], dys[u], dzs[u], rsq, 1.0f);
                    }
                }
            }
        if (near_min <= 64u)
            fix_pass(cnt);
        near_min = 0xffffffffu;
        }

    if constexpr (PACKED)
        {
        // fixed order: even-pair sums + odd-pair sums
        // (the scalar accumulators hold fix_pass's rare corrections)
        fx += fx2.x + fx2.y;
        fy += fy2.x + fy2.y;
        fz += fz2.x + fz2.y;
        pe += pe2.x + pe2.y;
        if (VIRIAL)
            {
            vxx += wxx.x + wxx.y;
            vxy += wxy.x + wxy.y;
            vxz += wxz.x + wxz.y;
            vyy += wyy.x + wyy.y;
            vyz += wyz.x + wyz.y;
            vzz += wzz.x + wzz.y;
            }
        }

    // combine the TPP lanes of a particle (fixed butterfly -> reproducible)
#pragma unroll
    for (int off = TPP >> 1; off > 0; off >>= 1)
        {
        fx += __shfl_xor_sync(0xffffffffu, fx, off);
        fy += __shfl_xor_sync(0xffffffffu, fy, off);
        fz += __shfl_xor_sync(0xffffffffu, fz, off);
        pe += __shfl_xor_sync(0xffffffffu, pe, off);
        if (VIRIAL)
            {
            vxx += __shfl_xor_sync(0xffffffffu, vxx, off);
            vxy += __shfl_xor_sync(0xffffffffu, vxy, off);
            vxz += __shfl_xor_sync(0xffffffffu, vxz, off);
            vyy += __shfl_xor_sync(0xffffffffu, vyy, off);
            vyz += __shfl_xor_sync(0xffffffffu, vyz, off);
            vzz += __shfl_xor_sync(0xffffffffu, vzz, off);
            }
        }
    if (active && sub == 0)
        {
        a.force[i] = make_float4(fx, fy, fz, pe);
        if (VIRIAL)
            {
            a.virial[0 * a.virial_pitch + i] = vxx;
            a.virial[1 * a.virial_pitch + i] = vxy;
            a.virial[2 * a.virial_pitch + i] = vxz;
            a.virial[3 * a.virial_pitch + i] = vyy;
            a.virial[4 * a.virial_pitch + i] = vyz;
            a.virial[5 * a.virial_pitch + i] = vzz;
            }
        }
    }

    } // namespace hb2
#include "pair_tile.cuh"
namespace hb2
    {
template<class Ev, int MODE, bool VIRIAL, int TPP>
hb2_status pair_launch_tpp(const PairKArgs<Ev>& k, cudaStream_t stream)
    {
    const uint64_t threads = (uint64_t)k.N * TPP;
    const uint32_t blocks = (uint32_t)((threads + PAIR_BLOCK - 1) / PAIR_BLOCK);
    // staged rows: PAIR_BLOCK / TPP rows of (stage * 32 + pad) words; at most 33 KB (TPP = 4)
    const size_t smem = (size_t)(PAIR_BLOCK / TPP) * (k.stage * 32u + PAIR_ROW_PAD) * sizeof(uint32_t);
    // the energy-free variant exists for the default TPP only (it is the step loop's kernel)
    if (TPP == 8 && !k.energy)
        {
        if (k.ntypes == 1)
            pair_force_kernel<Ev, MODE, VIRIAL, TPP, true, (TPP != 8)><<<blocks, PAIR_BLOCK, smem, stream>>>(k);
        else
            pair_force_kernel<Ev, MODE, VIRIAL, TPP, false, (TPP != 8)><<<blocks, PAIR_BLOCK, smem, stream>>>(k);
        }
    else if (k.ntypes == 1)
        pair_force_kernel<Ev, MODE, VIRIAL, TPP, true, true><<<blocks, PAIR_BLOCK, smem, stream>>>(k);
    else
        pair_force_kernel<Ev, MODE, VIRIAL, TPP, false, true><<<blocks, PAIR_BLOCK, smem, stream>>>(k);
    HB2_CHECK_LAUNCH("pair_force_kernel");
    return HB2_OK;
    }

template<class Ev, int MODE, bool VIRIAL>
hb2_status pair_launch(const PairKArgs<Ev>& k, uint32_t tpp, cudaStream_t stream)
    {
    // evaluators outside the benchmarked set are built for the default threads-per-particle only
    if constexpr (!Ev::tune_tpp)
        return pair_launch_tpp<Ev, MODE, VIRIAL, 8>(k, stream);
    else
    switch (tpp)
        {
    case 4:
        return pair_launch_tpp<Ev, MODE, VIRIAL, 4>(k, stream);
    case 8:
        return pair_launch_tpp<Ev, MODE, VIRIAL, 8>(k, stream);
    case 16:
        return pair_launch_tpp<Ev, MODE, VIRIAL, 16>(k, stream);
    case 32:
        return pair_launch_tpp<Ev, MODE, VIRIAL, 32>(k, stream);
    default:
        set_error("threads_per_particle must be 4, 8, 16 or 32");
        return HB2_ERR_INVALID;
        }
    }

template<class Ev, int MODE>
hb2_status pair_launch_mode(const PairKArgs<Ev>& k, bool virial, uint32_t tpp, cudaStream_t stream)
    {
    return virial ? pair_launch<Ev, MODE, true>(k, tpp, stream)
                  : pair_launch<Ev, MODE, false>(k, tpp, stream);
    }

// The one entry point an evaluator's .cu instantiates (the counterpart of the reference's
// `template hipError_t gpu_compute_pair_forces<Evaluator>(...)`,
// example_plugins/pair_plugin/PotentialPairExampleGPUKernel.cu:13-15).
template<class Ev>
hb2_status pair_dispatch(hb2_shift_mode mode, const PairArgs& a, const hb2_box& box,
                         cudaStream_t stream)
    {
    static_assert(sizeof(typename Ev::param_type) <= HB2_PARAM_FLOATS * sizeof(float)
                      && alignof(typename Ev::param_type) <= 16,
                  "packed param_type is at most HB2_PARAM_FLOATS floats on this ABI");
    if (a.N == 0)
        return HB2_OK;
    if (a.ntypes > PAIR_MAX_TYPES_SMEM)
        {
        set_error("pair forces: at most 16 particle types are supported");
        return HB2_ERR_INVALID;
        }
    if (a.pitch % 32u != 0)
        {
        set_error("pair forces: the neighbour-list pitch must be a multiple of 32");
        return HB2_ERR_INVALID;
        }
    PairKArgs<Ev> k;
    k.pos = a.pos;
    k.vel = a.vel;
    k.tag = a.tag;
    k.N = a.N;
    k.n_neigh = a.n_neigh;
    k.nlist = a.nlist;
    k.pitch = a.pitch;
    k.params = a.params;
    k.rcutsq = a.rcutsq;
    k.ronsq = a.ronsq;
    memcpy(&k.param0, a.params0, sizeof(k.param0));
    k.rcutsq0 = a.rcutsq0;
    k.ronsq0 = a.ronsq0;
    k.rcutsq_d = a.rcutsq_d;
    k.rcutsq0_d = a.rcutsq_d ? a.rcutsq0_d : (double)a.rcutsq0;
    k.ntypes = a.ntypes;
    k.force = a.force;
    k.virial = a.virial;
    k.virial_pitch = a.virial_pitch;
    k.seed = a.seed;
    k.d_timestep = a.d_timestep;
    k.timestep = a.timestep;
    k.dt = a.dt;
    k.kT = a.kT;
    k.box = boxf(box);
    k.boxd = boxd(box);
    k.face_margin = a.face_margin;
    k.energy = a.energy;
    k.skip_flag = a.skip_flag;
    k.stage = std::min<uint32_t>(std::max<uint32_t>(a.pitch / 32u, 1u), PAIR_STAGE_MAX);
    k.rows16 = a.rows16;
    k.tg = a.tg;
    k.geom = a.geom;
    k.cell_start = a.cell_start;
    k.cell_end = a.cell_end;
    k.stage_cap = a.stage_cap;
    k.tile_ticket = a.tile_ticket;
    k.tile_desc = a.tile_desc;
        {
        static const uint32_t flags = getenv("HB2_TILE_PREFETCH") ? (uint32_t)atoi(getenv("HB2_TILE_PREFETCH")) : 2u;
        k.tile_flags = flags;
        }
    if (Ev::needs_thermo && (!a.vel || !a.tag))
        {
        set_error("pair forces: this evaluator needs velocities and tags");
        return HB2_ERR_INVALID;
        }
    const bool virial = a.virial != nullptr;
    if (a.rows16)
        {
        if constexpr (Ev::needs_thermo)
            {
            set_error("pair forces: thermostatted evaluators read the per-particle rows, not the tile list");
            return HB2_ERR_INVALID;
            }
        else
            switch (mode)
                {
            case HB2_SHIFT_NONE:
                return pair_tile_launch_mode<Ev, 0>(k, virial, stream);
            case HB2_SHIFT_SHIFT:
                return pair_tile_launch_mode<Ev, 1>(k, virial, stream);
            case HB2_SHIFT_XPLOR:
                return pair_tile_launch_mode<Ev, 2>(k, virial, stream);
                }
        }
    switch (mode)
        {
    case HB2_SHIFT_NONE:
        return pair_launch_mode<Ev, 0>(k, virial, a.tpp, stream);
    case HB2_SHIFT_SHIFT:
        return pair_launch_mode<Ev, 1>(k, virial, a.tpp, stream);
    case HB2_SHIFT_XPLOR:
        return pair_launch_mode<Ev, 2>(k, virial, a.tpp, stream);
        }
    set_error("pair forces: unknown shift mode");
    return HB2_ERR_INVALID;
    }

// per-evaluator translation units
hb2_status pair_forces_lj(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_dpd(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_example(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_dpd_conservative(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_gauss(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_yukawa(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_lj1208(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_lj0804(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_force_shifted_lj(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_mie(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_morse(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_buckingham(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_expanded_lj(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_expanded_gaussian(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_lj_gauss(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_twf(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_dpdlj(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_table(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_expanded_mie(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_opp(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
hb2_status pair_forces_fourier(hb2_shift_mode, const PairArgs&, const hb2_box&, cudaStream_t);
    } // namespace hb2
