// pair_tile.cuh — PotentialPair<Evaluator> force / energy / virial pass over the tile-shared list
// (tile.cuh). Same semantics as pair_kernel.cuh (hoomd/md/PotentialPair.h:587-794,
// hoomd/md/PotentialPairGPU.cuh:146-404); included by it.
//
// A PERSISTENT kernel: a few blocks per SM, each a producer warp and eight consumer warps joined by a
// two-stage pipeline of tiles in shared memory (mbarriers "full" / "empty" per stage, no
// __syncthreads after start-up).
//  * The producer draws the next tile from a global ticket, reads the tile's descriptor — header, home
//    records (global index, staged offset, row length) and copy runs, written once per list build
//    (tile.cuh, nlist_tile.cu) — and issues one bulk copy (cp.async.bulk, TMA 1-D, UBLKCP in SASS) for
//    the header + home records and one per run of stencil cells that are contiguous in memory: the stored
//    global coordinates arrive bit for bit, without a register pass or LSU wavefronts, while the
//    consumers are still busy with the previous tile. It also asks for the tile's rows to be brought into
//    L2 (cp.async.bulk.prefetch.L2) so that the consumers' row loads are short.
//  * A consumer warp takes four home particles per pass, eight lanes per particle. A lane keeps its four
//    16-bit neighbour slots of every 32-entry chunk of the row in registers (rows are stored
//    lane-interleaved, TILE_SLOT: one coalesced 8-byte load per chunk) — the rows of the NEXT pass, also
//    across a tile boundary, are requested before the current pass is evaluated — and gathers neighbour
//    positions with LDS.128 from the staged stencil instead of through L1.
//  * Passes are dealt to the warps round-robin and the deal continues across tiles, so no warp collects
//    the odd passes; a warp that finishes a tile walks on into the next one without waiting for the others.
//  * Separations: a tile none of whose stencil cells lies across a periodic boundary uses the plain
//    difference (exact or correctly rounded); otherwise the (hi, lo) image-shift form of
//    pair_kernel.cuh. Arithmetic, exact cut-off fix-up and the reduction tree are those of
//    pair_kernel.cuh: owner-computes, no atomics, bitwise reproducible (the result does not depend on
//    which block or warp evaluates a particle).
#pragma once

namespace hb2
    {
constexpr int PT_CONSUMERS = 8;                     // consumer warps per block
constexpr int PT_THREADS = 32 * (PT_CONSUMERS + 1); // + the producer warp
constexpr int PT_NPRE = 4;                          // chunks of a row held in registers (longer rows: loads in the loop)

// one pipeline stage in shared memory: [TileDescHeader][home records: hcap x uint2][staged positions]
inline size_t pair_tile_slot_bytes(uint32_t stage_cap)
    {
    return sizeof(TileDescHeader) + sizeof(uint2) * (size_t)tile_hcap(stage_cap) + sizeof(float4) * stage_cap;
    }
inline size_t pair_tile_smem(uint32_t stage_cap)
    {
    return 2 * pair_tile_slot_bytes(stage_cap);
    }

HB2_DEVICE void mbar_arrive(uint64_t* bar)
    {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
    }
HB2_DEVICE void bulk_prefetch_l2(const void* src, uint32_t bytes)
    {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
    }

template<class Ev, int MODE, bool VIRIAL, bool ONETYPE, bool ENERGY>
__global__ void __launch_bounds__(PT_THREADS, (VIRIAL || !ONETYPE) ? 2 : 3)
    pair_force_tile_kernel(const PairKArgs<Ev> a)
    {
    typedef typename Ev::param_type param_type;
    if (a.skip_flag && *a.skip_flag)
        return;
    __shared__ param_type s_params[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ float s_rcutsq[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ float s_ronsq[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
    __shared__ float2 s_shift[3][PT_CONSUMERS * 4]; // (hi, lo) of pi -/+ L per component and particle
    extern __shared__ __align__(128) unsigned char s_raw[];
    const uint32_t hcap = tile_hcap(a.stage_cap);
    const uint32_t home_off = (uint32_t)sizeof(TileDescHeader);
    const uint32_t stage_off = home_off + (uint32_t)sizeof(uint2) * hcap;
    const uint32_t slot_bytes = stage_off + (uint32_t)sizeof(float4) * a.stage_cap;
    const size_t desc_stride = tile_desc_stride(hcap);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t ntiles = a.tg.nt[0] * a.tg.nt[1] * a.tg.nt[2];
    const uint32_t pitch_b = a.pitch * 2u; // bytes per row

    if (threadIdx.x == 0)
        {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        mbar_init(&s_empty[0], PT_CONSUMERS);
        mbar_init(&s_empty[1], PT_CONSUMERS);
        // slot 0 of either stage: the far-away dummy (the copies start at slot 1)
        *reinterpret_cast<float4*>(s_raw + stage_off) = make_float4(NL_FAR, NL_FAR, NL_FAR, 0.f);
        *reinterpret_cast<float4*>(s_raw + slot_bytes + stage_off) = make_float4(NL_FAR, NL_FAR, NL_FAR, 0.f);
        }
    if (!ONETYPE)
        {
        const uint32_t nt2 = a.ntypes * a.ntypes;
        for (uint32_t t = threadIdx.x; t < nt2; t += blockDim.x)
            {
            s_params[t] = *reinterpret_cast<const param_type*>(a.params + HB2_PARAM_FLOATS * t);
            s_rcutsq[t] = a.rcutsq[t];
            s_ronsq[t] = a.ronsq[t];
            }
        }
    __syncthreads();

    if (w == PT_CONSUMERS)
        {
        // ================= producer =================
        // The ticket and the descriptor of the NEXT tile are drawn right after a tile's copies are issued,
        // and the next tile's positions and rows are asked into L2 before its stage is free: the path from
        // "stage released" to "tile landed" is one bulk copy out of L2.
        uint32_t tile = 0;
        uint4 hdr = make_uint4(0u, 0u, 0u, 0u), r0 = hdr, r1 = hdr;
        const unsigned char* desc = nullptr;
        auto draw = [&]()
        {
            if (lane == 0)
                tile = atomicAdd(a.tile_ticket, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile < ntiles)
                {
                desc = a.tile_desc + (size_t)tile * desc_stride;
                // (the runs follow the hcap home records: the same offset as the staged array in a stage)
                const uint4* const runs = reinterpret_cast<const uint4*>(desc + stage_off);
                hdr = __ldg(reinterpret_cast<const uint4*>(desc)); // valid, nhome, shifting, total
                r0 = __ldg(runs + lane);
                r1 = __ldg(runs + 32 + lane);
                }
        };
        draw();
        for (uint32_t it = 0;; it++)
            {
            const uint32_t b = it & 1u;
            unsigned char* const slot = s_raw + b * slot_bytes;
            if (tile >= ntiles)
                {
                if (it >= 2u)
                    mbar_wait(&s_empty[b], ((it >> 1) - 1u) & 1u);
                if (lane == 0)
                    {
                    reinterpret_cast<TileDescHeader*>(slot)->valid = 0u;
                    mbar_arrive(&s_full[b]);
                    // the last block to run out re-arms the ticket for the next launch
                    if (atomicAdd(a.tile_ticket + 1, 1u) == gridDim.x - 1u)
                        {
                        a.tile_ticket[0] = 0u;
                        a.tile_ticket[1] = 0u;
                        }
                    }
                break;
                }
            const uint32_t nhome = hdr.y;
            // into L2 ahead of use: the stencil's positions and the rows of the home particles
            if (r0.z && (a.tile_flags & 2u))
                bulk_prefetch_l2(a.pos + r0.x, r0.z * (uint32_t)sizeof(float4));
            if (r1.z && (a.tile_flags & 2u))
                bulk_prefetch_l2(a.pos + r1.x, r1.z * (uint32_t)sizeof(float4));
            const uint2* const dh = reinterpret_cast<const uint2*>(desc + home_off);
            for (uint32_t hs = lane; hs < ((a.tile_flags & 1u) ? nhome : 0u); hs += 32u)
                {
                const uint2 h = __ldg(dh + hs);
                const uint32_t nch = ((h.y >> 16) + 31u) >> 5;
                if (nch)
                    bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(a.rows16) + (size_t)h.x * pitch_b, nch * 64u);
                }
            if (it >= 2u) // the consumers have left the tile this stage held before
                mbar_wait(&s_empty[b], ((it >> 1) - 1u) & 1u);
            const uint32_t head_bytes = (home_off + (uint32_t)sizeof(uint2) * nhome + 15u) & ~15u;
            if (lane == 0)
                {
                mbar_expect_tx(&s_full[b], head_bytes + (hdr.w - 1u) * (uint32_t)sizeof(float4));
                bulk_copy_g2s(slot, desc, head_bytes, &s_full[b]);
                }
            if (r0.z)
                bulk_copy_g2s(slot + stage_off + r0.y * 16u, a.pos + r0.x, r0.z * (uint32_t)sizeof(float4), &s_full[b]);
            if (r1.z)
                bulk_copy_g2s(slot + stage_off + r1.y * 16u, a.pos + r1.x, r1.z * (uint32_t)sizeof(float4), &s_full[b]);
            __syncwarp();
            draw();
            }
        return;
        }

    // ================= consumers =================
    const uint32_t sub = lane & 7u, grp = lane >> 3;
    const unsigned char* const rows_b = reinterpret_cast<const unsigned char*>(a.rows16) + sub * 8u;
    auto header_of = [&](const uint32_t b) -> const TileDescHeader&
    { return *reinterpret_cast<const TileDescHeader*>(s_raw + b * slot_bytes); };
    auto home = [&](const uint32_t b, const uint32_t hs) -> uint2
    {
        if (hs >= header_of(b).nhome)
            return make_uint2(0u, 0u); // no particle: the far-away dummy with an empty row
        return *reinterpret_cast<const uint2*>(s_raw + b * slot_bytes + home_off + hs * 8u);
    };
    // this lane's four slots of each of the first PT_NPRE chunks of a row
    auto load_rows = [&](const uint2 h, uint2 (&ix)[PT_NPRE])
    {
        const uint32_t nch = ((h.y >> 16) + 31u) >> 5;
        const uint2* src = reinterpret_cast<const uint2*>(rows_b + (size_t)h.x * pitch_b);
#pragma unroll
        for (int c = 0; c < PT_NPRE; c++)
            ix[c] = (uint32_t)c < nch ? __ldg(src + 8 * c) : make_uint2(0u, 0u);
    };

    const float hc[3] = {a.box.hi[0], a.box.hi[1], a.box.hi[2]};
    const uint32_t shift_s = smem_u32(&s_shift[0][w * 4 + grp]);

    // ---- one pass: four home particles of the tile in stage b
    auto compute_pass = [&](const uint32_t b, const uint2 hcur, const uint2 (&ic)[PT_NPRE])
    {
        const uint32_t stage_s = smem_u32(s_raw + b * slot_bytes + stage_off);
        const bool valid = hcur.y != 0u;
        const uint32_t gi = hcur.x, n = hcur.y >> 16;
        const float4 pi = lds128(stage_s + (hcur.y & 0xffffu));
        // In a tile that touches a periodic boundary only the particles within reach of a (global) face
        // can have a neighbour across it (the face_margin test of pair_kernel.cuh): half the passes of a
        // face tile take the plain loop.
        bool SHIFT = header_of(b).shifting != 0u;
        if (SHIFT)
            {
            const float m = a.face_margin > 0.f ? a.face_margin : 3.0e38f;
            const bool near_face = valid
                                   && ((pi.x >= a.box.hi[0] - m) || (pi.x < a.box.lo[0] + m) || (pi.y >= a.box.hi[1] - m)
                                       || (pi.y < a.box.lo[1] + m) || (pi.z >= a.box.hi[2] - m) || (pi.z < a.box.lo[2] + m));
            SHIFT = __any_sync(0xffffffffu, near_face);
            }
        const uint32_t ti = ONETYPE ? 0u : __float_as_uint(pi.w);
        const uint32_t nchunks = (n + 31u) >> 5;
        const uint32_t nch_warp = __reduce_max_sync(0xffffffffu, nchunks);

        float fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
        float vxx = 0.f, vxy = 0.f, vxz = 0.f, vyy = 0.f, vyz = 0.f, vzz = 0.f;
        constexpr bool PACKED = Ev::has_packed && ONETYPE && MODE != 2 && !Ev::needs_thermo;
        float2 fx2 = make_float2(0.f, 0.f), fy2 = fx2, fz2 = fx2, pe2 = fx2;
        float2 wxx = fx2, wxy = fx2, wxz = fx2, wyy = fx2, wyz = fx2, wzz = fx2;
        uint32_t near_min = 0xffffffffu; // see pair_kernel.cuh: exact cut-off fix-up
        auto note_near = [&](const float rsq, const float rcutsq)
        {
#ifndef PAIR_NO_EXACT
            near_min = min(near_min, (uint32_t)(__float_as_int(rsq) + (32 - __float_as_int(rcutsq))));
#endif
        };
        auto pair_body = [&](const float4 pj, const float dx, const float dy, const float dz, const float rsq,
                             const float sgn)
        {
            float rcutsq = a.rcutsq0, ronsq = a.ronsq0;
            const param_type* prm = &a.param0;
            if (!ONETYPE)
                {
                const uint32_t tp = __float_as_uint(pj.w) * a.ntypes + ti; // hoomd/Index1D.h:53
                rcutsq = s_rcutsq[tp];
                ronsq = s_ronsq[tp];
                prm = &s_params[tp];
                }
            bool energy_shift = (MODE == 1);
            if (MODE == 2)
                energy_shift = ronsq > rcutsq; // hoomd/md/PotentialPair.h:697-700
            float force_divr = 0.f, pair_eng = 0.f;
            Ev eval(rsq, rcutsq, *prm);
            const bool evaluated = eval.evalForceAndEnergy(force_divr, pair_eng, energy_shift);
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
                    }
                }
            force_divr = evaluated ? sgn * force_divr : 0.f;
            pair_eng = evaluated ? sgn * pair_eng : 0.f;
            fx += dx * force_divr;
            fy += dy * force_divr;
            fz += dz * force_divr;
            if (ENERGY)
                pe += pair_eng * 0.5f;
            if (VIRIAL)
                {
                const float f2 = 0.5f * force_divr;
                vxx += f2 * dx * dx;
                vxy += f2 * dx * dy;
                vxz += f2 * dx * dz;
                vyy += f2 * dy * dy;
                vyz += f2 * dy * dz;
                vzz += f2 * dz * dz;
                }
        };
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

        // (hi, lo) of pi -/+ L, see pair_kernel.cuh; kept in shared memory so that the plain tiles' loop
        // — all tiles but a surface layer — does not carry six more registers
        const float pc[3] = {pi.x, pi.y, pi.z};
        if (SHIFT)
            {
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 3; c++)
                {
                const double m = pc[c] >= 0.f ? (double)pc[c] - a.boxd.L[c] : (double)pc[c] + a.boxd.L[c];
                const float hi = (float)m;
                if (sub == 0u)
                    s_shift[c][w * 4 + grp] = make_float2(hi, (float)(m - (double)hi));
                }
            __syncwarp();
            }
        auto separation_shift = [&](const float4 (&pj)[4], float (&dx)[4], float (&dy)[4], float (&dz)[4])
        {
            float* const d[3] = {dx, dy, dz};
#pragma unroll
            for (int c = 0; c < 3; c++)
                {
                const uint2 hl = lds64(shift_s + c * (uint32_t)(PT_CONSUMERS * 4 * sizeof(float2)));
                const float sh = __uint_as_float(hl.x), sl = __uint_as_float(hl.y);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    {
                    const float qc = c == 0 ? pj[u].x : (c == 1 ? pj[u].y : pj[u].z);
                    const float d0 = pc[c] - qc;
                    const float ds = (sh - qc) + sl;
                    d[c][u] = fabsf(d0) >= hc[c] ? ds : d0;
                    }
                }
        };
        auto rsq_of = [](const float dx, const float dy, const float dz)
        { return fmaf(dz, dz, fmaf(dy, dy, dx * dx)); };

        // one 32-entry chunk: wc = this lane's four staged byte offsets
        auto chunk = [&](const uint2 wc, const bool shifted)
        {
            float4 pj[4];
            pj[0] = lds128(stage_s + (wc.x & 0xffffu));
            pj[1] = lds128(stage_s + (wc.x >> 16));
            pj[2] = lds128(stage_s + (wc.y & 0xffffu));
            pj[3] = lds128(stage_s + (wc.y >> 16));
            float dxs[4], dys[4], dzs[4];
            if (!shifted)
                {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    {
                    dxs[u] = pi.x - pj[u].x;
                    dys[u] = pi.y - pj[u].y;
                    dzs[u] = pi.z - pj[u].z;
                    }
                }
            else
                separation_shift(pj, dxs, dys, dzs);
            if constexpr (PACKED)
                {
#pragma unroll
                for (int u = 0; u < 4; u += 2)
                    pair_body2(make_float2(dxs[u], dxs[u + 1]), make_float2(dys[u], dys[u + 1]),
                               make_float2(dzs[u], dzs[u + 1]));
                }
            else
                {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    {
                    const float rsq = rsq_of(dxs[u], dys[u], dzs[u]);
                    note_near(rsq, ONETYPE ? a.rcutsq0 : s_rcutsq[__float_as_uint(pj[u].w) * a.ntypes + ti]);
                    pair_body(pj[u], dxs[u], dys[u], dzs[u], rsq, 1.0f);
                    }
                }
        };
        auto walk = [&](const bool shifted)
        {
#pragma unroll
            for (int c = 0; c < PT_NPRE; c++)
                if ((uint32_t)c < nch_warp)
                    chunk(ic[c], shifted);
            if (nch_warp > (uint32_t)PT_NPRE)
                {
                // rows longer than the register window
                const uint2* src = reinterpret_cast<const uint2*>(rows_b + (size_t)gi * pitch_b);
                for (uint32_t c = PT_NPRE; c < nch_warp; c++)
                    chunk(c < nchunks ? __ldg(src + 8u * c) : make_uint2(0u, 0u), shifted);
                }
        };
        if (SHIFT)
            walk(true);
        else
            walk(false);

        if (near_min <= 64u)
            {
            // rare (about one pass in 1e5): some rsq came within 32 ulp of r_cut^2. Re-walk the row, decide
            // those pairs with the CPU build's exact arithmetic — the staged positions ARE the stored
            // global coordinates — and repair the pairs where fp32 and the reference disagree.
            const uint16_t* __restrict__ row = a.rows16 + (size_t)gi * a.pitch + 4u * sub;
            const float4* stage = reinterpret_cast<const float4*>(s_raw + b * slot_bytes + stage_off);
            for (uint32_t e = 0; e < nchunks * 4u; e++)
                {
                // this lane's storage slots: 4 sub .. 4 sub + 3 of every chunk
                const uint32_t jj = __ldg(row + (e >> 2) * 32u + (e & 3u));
                if (jj == 0u)
                    continue;
                const float4 qq = stage[jj >> 4];
                float dx = pi.x - qq.x, dy = pi.y - qq.y, dz = pi.z - qq.z;
                if (SHIFT)
                    {
                    const float2 sx = s_shift[0][w * 4 + grp], sy = s_shift[1][w * 4 + grp], sz = s_shift[2][w * 4 + grp];
                    dx = fabsf(dx) >= hc[0] ? (sx.x - qq.x) + sx.y : dx;
                    dy = fabsf(dy) >= hc[1] ? (sy.x - qq.y) + sy.y : dy;
                    dz = fabsf(dz) >= hc[2] ? (sz.x - qq.z) + sz.y : dz;
                    }
                const float rsq = rsq_of(dx, dy, dz);
                const uint32_t tp = ONETYPE ? 0u : __float_as_uint(qq.w) * a.ntypes + ti;
                const float rcutsq = ONETYPE ? a.rcutsq0 : s_rcutsq[tp];
                if ((uint32_t)(__float_as_int(rsq) - __float_as_int(rcutsq) + 32) > 64u)
                    continue;
                const double rc2 = ONETYPE ? a.rcutsq0_d : (a.rcutsq_d ? a.rcutsq_d[tp] : (double)rcutsq);
                const bool in_ref = dist_sq_exact(a.boxd, pi, qq) < rc2;
                const bool in_f32 = rsq < rcutsq;
                if (in_ref == in_f32)
                    continue;
                const float below = __uint_as_float(__float_as_uint(rcutsq) - 1u);
                pair_body(qq, dx, dy, dz, in_ref ? below : rsq, in_ref ? 1.0f : -1.0f);
                }
            }

        if constexpr (PACKED)
            {
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
#pragma unroll
        for (int off = 4; off > 0; off >>= 1)
            {
            fx += __shfl_xor_sync(0xffffffffu, fx, off);
            fy += __shfl_xor_sync(0xffffffffu, fy, off);
            fz += __shfl_xor_sync(0xffffffffu, fz, off);
            if (ENERGY)
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
        if (valid && sub == 0)
            {
            a.force[gi] = make_float4(fx, fy, fz, pe);
            if (VIRIAL)
                {
                a.virial[0 * a.virial_pitch + gi] = vxx;
                a.virial[1 * a.virial_pitch + gi] = vxy;
                a.virial[2 * a.virial_pitch + gi] = vxz;
                a.virial[3 * a.virial_pitch + gi] = vyy;
                a.virial[4 * a.virial_pitch + gi] = vyz;
                a.virial[5 * a.virial_pitch + gi] = vzz;
                }
            }
    };

    // ---- the walk over tiles. Pass p of a tile goes to warp (p + rot) mod 8 and rot advances by the
    // tile's pass count: the deal continues across tiles. (it, b, rot, npass) is the tile the warp is
    // positioned on — acquired: its stage has landed; q is this warp's next pass in it.
    // The rows of an item are loaded at ONE place in the loop and reach the evaluation through register
    // copies only, so that the evaluation of one pass never waits on the loads of the next.
    // The walk's state is five registers — it, rot, q, npass and a word of flags (separate bools ended up
    // in local memory, and a local load queues behind the row loads in the L1 pipe).
    uint32_t it = 0, rot = 0, q = 0, npass = 0;
    constexpr uint32_t F_ITEM = 1u, F_CUR = 2u, F_LAST = 4u, F_ENDED = 8u;
    // F_ITEM: (it, q) names the next pass to fetch; F_CUR: hcur / ic hold a pass to evaluate;
    // F_LAST: that pass is the warp's last in its tile (stage (it - 1) & 1, released after it);
    // F_ENDED: the ticket ran out
    uint32_t flags = 0;
    auto arrive = [&](const uint32_t stage)
    {
        __syncwarp();
        if (lane == 0)
            mbar_arrive(&s_empty[stage]);
    };
    // from the acquired tile on: the first tile that holds a pass for this warp (tiles without one are
    // released at once); false: the ticket ran out
    auto seek = [&]() -> bool
    {
        for (;;)
            {
            q = ((uint32_t)w - rot) & (PT_CONSUMERS - 1);
            if (q < npass)
                return true;
            arrive(it & 1u);
            rot = (rot + npass) & (PT_CONSUMERS - 1);
            it++;
            mbar_wait(&s_full[it & 1u], (it >> 1) & 1u);
            if (header_of(it & 1u).valid == 0u)
                return false;
            npass = (header_of(it & 1u).nhome + 3u) >> 2;
            }
    };
    mbar_wait(&s_full[0], 0u);
    if (header_of(0).valid == 0u)
        return;
    npass = (header_of(0).nhome + 3u) >> 2;
    flags = seek() ? F_ITEM : F_ENDED;
    uint2 hcur = make_uint2(0u, 0u), ic[PT_NPRE];
#pragma unroll
    for (int c = 0; c < PT_NPRE; c++)
        ic[c] = make_uint2(0u, 0u);
    for (;;)
        {
        uint2 hnext = make_uint2(0u, 0u), in_[PT_NPRE];
#pragma unroll
        for (int c = 0; c < PT_NPRE; c++)
            in_[c] = make_uint2(0u, 0u);
        if (flags & F_ITEM)
            {
            hnext = home(it & 1u, 4u * q + grp);
            load_rows(hnext, in_);
            }
        if (flags & F_CUR)
            {
            const uint32_t bcur = (it - ((flags & F_LAST) ? 1u : 0u)) & 1u;
            compute_pass(bcur, hcur, ic);
            if (flags & F_LAST)
                arrive(bcur);
            }
        if (!(flags & F_ITEM))
            {
            if (flags & F_ENDED)
                break;
            // the tile the warp is positioned on holds no pass for it (a small tile): walk on
            flags = seek() ? F_ITEM : F_ENDED;
            if (flags & F_ENDED)
                break;
            continue;
            }
        hcur = hnext;
#pragma unroll
        for (int c = 0; c < PT_NPRE; c++)
            ic[c] = in_[c];
        flags = F_ITEM | F_CUR;
        // the item after this one
        q += PT_CONSUMERS;
        if (q >= npass)
            {
            // look into the next tile: its stage was released by this warp before it entered the current
            // one, which is released once the item just fetched has been evaluated
            flags = F_CUR | F_LAST;
            rot = (rot + npass) & (PT_CONSUMERS - 1);
            it++;
            mbar_wait(&s_full[it & 1u], (it >> 1) & 1u);
            if (header_of(it & 1u).valid == 0u)
                flags |= F_ENDED;
            else
                {
                npass = (header_of(it & 1u).nhome + 3u) >> 2;
                q = ((uint32_t)w - rot) & (PT_CONSUMERS - 1);
                if (q < npass)
                    flags |= F_ITEM;
                }
            }
        }
    }

// blocks per SM of a kernel instantiation at this shared-memory size (cached per kernel: instantiations
// share the function-pointer type, so the key is the pointer)
template<class K> int pair_tile_occupancy(K kern, size_t smem)
    {
    struct Entry
        {
        const void* fn;
        size_t smem;
        int blocks;
        };
    static std::vector<Entry> cache;
    for (const Entry& e : cache)
        if (e.fn == (const void*)kern && e.smem == smem)
            return e.blocks;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 0;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, PT_THREADS, smem) != cudaSuccess)
        return 0;
    for (Entry& e : cache)
        if (e.fn == (const void*)kern)
            {
            e.smem = smem;
            e.blocks = nb;
            return nb;
            }
    cache.push_back({(const void*)kern, smem, nb});
    return nb;
    }

template<class Ev, int MODE, bool VIRIAL>
hb2_status pair_tile_launch(const PairKArgs<Ev>& k, cudaStream_t stream)
    {
    const uint32_t ntiles = k.tg.nt[0] * k.tg.nt[1] * k.tg.nt[2];
    const size_t smem = pair_tile_smem(k.stage_cap);
    static int sms = 0;
    if (sms == 0)
        {
        int dev = 0;
        HB2_CUDA(cudaGetDevice(&dev));
        HB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
    auto go = [&](auto kern) -> hb2_status
    {
        const int per_sm = pair_tile_occupancy(kern, smem);
        if (per_sm <= 0)
            {
            cudaGetLastError();
            set_error("pair forces (tiles): the pipeline does not fit in shared memory");
            return HB2_ERR_INVALID;
            }
        const uint32_t blocks = std::min<uint32_t>(ntiles, (uint32_t)(per_sm * sms));
        kern<<<blocks, PT_THREADS, smem, stream>>>(k);
        HB2_CHECK_LAUNCH("pair_force_tile_kernel");
        return HB2_OK;
    };
    if (k.ntypes == 1)
        return k.energy ? go(pair_force_tile_kernel<Ev, MODE, VIRIAL, true, true>)
                        : go(pair_force_tile_kernel<Ev, MODE, VIRIAL, true, false>);
    return k.energy ? go(pair_force_tile_kernel<Ev, MODE, VIRIAL, false, true>)
                    : go(pair_force_tile_kernel<Ev, MODE, VIRIAL, false, false>);
    }

template<class Ev, int MODE>
hb2_status pair_tile_launch_mode(const PairKArgs<Ev>& k, bool virial, cudaStream_t stream)
    {
    return virial ? pair_tile_launch<Ev, MODE, true>(k, stream) : pair_tile_launch<Ev, MODE, false>(k, stream);
    }
    } // namespace hb2
