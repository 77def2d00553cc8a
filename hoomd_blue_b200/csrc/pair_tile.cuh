// pair_tile.cuh — PotentialPair<Evaluator> force / energy / virial pass over the tile-shared list
// (tile.cuh). Same semantics as pair_kernel.cuh (hoomd/md/PotentialPair.h:587-794,
// hoomd/md/PotentialPairGPU.cuh:146-404); included by it.
//
// One block per tile of 2 x 2 x 2 home cells.
//  1. tile_layout, then one bulk copy (cp.async.bulk, TMA 1-D, UBLKCP in SASS) per stencil cell brings
//     the 4 x 4 x 4 stencil's float4 positions — the stored global coordinates, bit for bit — into shared
//     memory; completion on one mbarrier. No register pass, no LSU wavefronts.
//  2. Each warp takes four home particles per pass, eight lanes per particle. Lane `sub` reads its four
//     16-bit neighbour slots of a 32-entry chunk with one coalesced 8-byte load (rows are stored
//     lane-interleaved, TILE_SLOT) — the next chunk, and the first chunk of the warp's next pass, are
//     requested before the current one is evaluated — and gathers the four positions with LDS.128 from
//     the staged stencil instead of through L1.
//  3. Separations: a tile none of whose stencil cells lies across a periodic boundary uses the plain
//     difference (exact or correctly rounded); otherwise the (hi, lo) image-shift form of
//     pair_kernel.cuh. Arithmetic, exact cut-off fix-up and the reduction tree are those of
//     pair_kernel.cuh: owner-computes, no atomics, bitwise reproducible.
#pragma once

namespace hb2
    {
#ifndef PAIR_TILE_MIN_BLOCKS
#define PAIR_TILE_MIN_BLOCKS 4
#endif

constexpr int PT_HCAP = 384; // home particles of a tile with a record in shared memory

inline size_t pair_tile_smem(uint32_t stage_cap, uint32_t pitch)
    {
    return ((sizeof(TileLayout) + 15) & ~(size_t)15) + sizeof(float4) * stage_cap
           + sizeof(uint16_t) * TILE_WARPS * 2 * 4 * pitch;
    }

// does any stencil cell of this tile lie across a periodic boundary? (geometry only: the same answer as
// TileLayout::any_shift, before anything is loaded)
HB2_DEVICE bool tile_is_shifting(const uint32_t tile, const TileGeom& tg, const DomainGeom& geom)
    {
    const uint32_t tt[3] = {tile % tg.nt[0], (tile / tg.nt[0]) % tg.nt[1], tile / (tg.nt[0] * tg.nt[1])};
    bool sh = false;
#pragma unroll
    for (int d = 0; d < 3; d++)
        {
        const int c0 = (int)(tg.first[d] + 2u * tt[d]);
        const int nc = min(2, (int)(tg.first[d] + geom.dimi[d]) - c0);
        const int below = c0 - 1, above = c0 + nc;
        if (geom.n[d] > 1)
            sh = sh || (below == 0 && geom.coord[d] == 0) || (above == (int)geom.dimt[d] - 1 && geom.coord[d] == geom.n[d] - 1);
        else
            sh = sh || below < 0 || above >= (int)geom.dimt[d];
        }
    return sh;
    }

// Tiles that touch a periodic boundary use the (hi, lo) image-shift separations of pair_kernel.cuh; their
// per-particle shift operands live in shared memory (one slot per thread, read inside the pair loop), so
// that the plain tiles' loop — all tiles but a surface layer — does not carry nine more registers.
template<class Ev, int MODE, bool VIRIAL, bool ONETYPE, bool ENERGY>
__global__ void __launch_bounds__(TILE_THREADS, (VIRIAL || !ONETYPE) ? 3 : PAIR_TILE_MIN_BLOCKS)
    pair_force_tile_kernel(const PairKArgs<Ev> a)
    {
    typedef typename Ev::param_type param_type;
    if (a.skip_flag && *a.skip_flag)
        return;
    const bool SHIFT = tile_is_shifting(blockIdx.x, a.tg, a.geom);
    __shared__ param_type s_params[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ float s_rcutsq[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ float s_ronsq[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(8) uint2 s_home[PT_HCAP];   // global index, staged byte offset | n_neigh << 16
    __shared__ float2 s_shift[3][TILE_THREADS / 8];  // (hi, lo) of pi -/+ L per component and particle
    extern __shared__ __align__(128) unsigned char s_raw[];
    TileLayout& L = *reinterpret_cast<TileLayout*>(s_raw);
    float4* __restrict__ stage = reinterpret_cast<float4*>(s_raw + ((sizeof(TileLayout) + 15) & ~(size_t)15));
    uint16_t* rowbuf_all = reinterpret_cast<uint16_t*>(stage + a.stage_cap);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    if (threadIdx.x == 0)
        mbar_init(&s_bar, 1);
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
    const uint32_t total = tile_layout(L, blockIdx.x, a.tg, a.geom, a.cell_start, a.cell_end);
    const uint32_t ns = L.nk[0] * L.nk[1] * L.nk[2];
    if (total > a.stage_cap)
        __trap(); // the build ran with the same capacity over the same cells
    // ---- 1. stage the stencil: warp 0 issues the bulk copies
    if (w == 0)
        {
        if (lane == 0)
            {
            stage[0] = make_float4(NL_FAR, NL_FAR, NL_FAR, 0.f);
            mbar_expect_tx(&s_bar, (total - 1u) * (uint32_t)sizeof(float4));
            }
        __syncwarp();
        // stencil cells that follow each other in memory (the x-neighbours of a row-major cell order)
        // travel as one copy: the lane of the first cell of such a run issues it
#pragma unroll
        for (int h = 0; h < 2; h++)
            {
            const uint32_t s = lane + 32 * h;
            if (s < ns)
                {
                const uint32_t o = L.off[s], g0 = L.gstart[s];
                const bool head = s == 0u || L.gstart[s - 1] + (o - L.off[s - 1]) != g0;
                if (head)
                    {
                    uint32_t t = s + 1u, tot = L.off[t] - o;
                    while (t < ns && L.gstart[t] == g0 + tot)
                        {
                        t++;
                        tot = L.off[t] - o;
                        }
                    if (tot)
                        bulk_copy_g2s(stage + o, a.pos + g0, tot * (uint32_t)sizeof(float4), &s_bar);
                    }
                }
            }
        }
    const uint32_t nhome = L.hoff[8];
    const uint32_t npass = (nhome + 3u) >> 2;
    const uint32_t sub = lane & 7u, grp = lane >> 3;

    // home slot hs of the tile -> (global index, staged byte offset | n_neigh << 16)
    auto resolve = [&](const uint32_t hs) -> uint2
    {
        uint32_t c = (hs >= L.hoff[4]) ? 4u : 0u;
        c += (hs >= L.hoff[c + 2]) ? 2u : 0u;
        c += (hs >= L.hoff[c + 1]) ? 1u : 0u;
        const uint32_t sc = L.hcell[c], k = hs - L.hoff[c];
        const uint32_t gi = L.gstart[sc] + k;
        return make_uint2(gi, ((L.off[sc] + k) * 16u) | (__ldg(a.n_neigh + gi) << 16));
    };
    for (uint32_t hs = threadIdx.x; hs < min(nhome, (uint32_t)PT_HCAP); hs += TILE_THREADS)
        s_home[hs] = resolve(hs);
    // the positions have landed (one warp polls the mbarrier; the others sleep in the block barrier)
    if (w == 0)
        mbar_wait(&s_bar, 0);
    __syncthreads();
    const uint32_t home_s = smem_u32(s_home);
    auto home = [&](const uint32_t hs) -> uint2
    {
        if (hs >= nhome)
            return make_uint2(0u, 0u); // no particle: the far-away dummy with an empty row
        if (hs < (uint32_t)PT_HCAP)
            return lds64(home_s + hs * 8u);
        return resolve(hs);
    };

    // ---- 2. rows: 16-byte asynchronous copies (cp.async, LDGSTS) into a per-warp double buffer, requested
    // one pass ahead. Lane `sub` of a particle copies quarter sub & 3 of chunks sub >> 2, sub >> 2 + 2, ...
    const uint32_t stage_s = smem_u32(stage);
    const uint32_t pitch_b = a.pitch * 2u; // bytes per row
    const uint32_t rowbuf_s = smem_u32(rowbuf_all) + (uint32_t)w * (8u * pitch_b) + grp * pitch_b;
    const uint32_t lane_off = (sub >> 2) * 64u + (sub & 3u) * 16u;
    auto request_rows = [&](const uint2 h, const uint32_t b)
    {
        const uint32_t nch = ((h.y >> 16) + 31u) >> 5;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.rows16) + (size_t)h.x * pitch_b + lane_off;
        uint32_t dst = rowbuf_s + b * (4u * pitch_b) + lane_off;
        for (uint32_t c = sub >> 2; c < nch; c += 2u)
            {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            src += 128;
            dst += 128u;
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint32_t q = w;
    uint2 hcur = home(4u * q + grp);
    request_rows(hcur, 0u);
    uint32_t b = 0;

    while (q < npass)
        {
        const uint32_t qn = q + TILE_WARPS;
        // this pass's rows have landed; the other buffer was last read in pass q - 1 (the reduction's
        // shuffles at its end joined the warp), so the next pass's rows can be requested into it
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const uint2 hnext = home(4u * qn + grp);
        request_rows(hnext, b ^ 1u);

        const bool valid = 4u * q + grp < nhome;
        const uint32_t gi = hcur.x, n = hcur.y >> 16;
        const float4 pi = lds128(stage_s + (hcur.y & 0xffffu));
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

        // (hi, lo) of pi -/+ L, see pair_kernel.cuh
        const float pc[3] = {pi.x, pi.y, pi.z};
        const float hc[3] = {a.box.hi[0], a.box.hi[1], a.box.hi[2]};
        const uint32_t shift_s = smem_u32(&s_shift[0][threadIdx.x >> 3]);
        if (SHIFT)
            {
#pragma unroll
            for (int c = 0; c < 3; c++)
                {
                const double m = pc[c] >= 0.f ? (double)pc[c] - a.boxd.L[c] : (double)pc[c] + a.boxd.L[c];
                const float hi = (float)m;
                if (sub == 0u)
                    s_shift[c][threadIdx.x >> 3] = make_float2(hi, (float)(m - (double)hi));
                }
            __syncwarp();
            }
        auto separation_shift = [&](const float4 (&pj)[4], float (&dx)[4], float (&dy)[4], float (&dz)[4])
        {
            float* const d[3] = {dx, dy, dz};
#pragma unroll
            for (int c = 0; c < 3; c++)
                {
                const uint2 hl = lds64(shift_s + c * (uint32_t)(TILE_THREADS / 8 * sizeof(float2)));
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

        const uint32_t idx_s = rowbuf_s + b * (4u * pitch_b) + sub * 8u;
        for (uint32_t c = 0; c < nch_warp; c++)
            {
            uint2 wc = lds64(idx_s + c * 64u);
            if (c >= nchunks) // a shorter row than the warp's longest: the dummy slot
                wc = make_uint2(0u, 0u);
            // row entries are staged slots times 16: byte offsets into the staged array
            float4 pj[4];
            pj[0] = lds128(stage_s + (wc.x & 0xffffu));
            pj[1] = lds128(stage_s + (wc.x >> 16));
            pj[2] = lds128(stage_s + (wc.y & 0xffffu));
            pj[3] = lds128(stage_s + (wc.y >> 16));
            float dxs[4], dys[4], dzs[4];
            if (!SHIFT)
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
            }
        if (near_min <= 64u)
            {
            // rare (about one pass in 1e5): some rsq came within 32 ulp of r_cut^2. Re-walk the row, decide
            // those pairs with the CPU build's exact arithmetic — the staged positions ARE the stored
            // global coordinates — and repair the pairs where fp32 and the reference disagree.
            const uint16_t* __restrict__ row
                = rowbuf_all + (size_t)((w * 2u + b) * 4u + grp) * a.pitch + 4u * sub;
            for (uint32_t e = 0; e < nchunks * 4u; e++)
                {
                // this lane's storage slots: 4 sub .. 4 sub + 3 of every chunk
                const uint32_t jj = row[(e >> 2) * 32u + (e & 3u)];
                if (jj == 0u)
                    continue;
                const float4 qq = stage[jj >> 4];
                float dx = pi.x - qq.x, dy = pi.y - qq.y, dz = pi.z - qq.z;
                if (SHIFT)
                    {
                    const float2 sx = s_shift[0][threadIdx.x >> 3], sy = s_shift[1][threadIdx.x >> 3], sz = s_shift[2][threadIdx.x >> 3];
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
        q = qn;
        hcur = hnext;
        b ^= 1u;
        }
    }

template<class Ev, int MODE, bool VIRIAL>
hb2_status pair_tile_launch(const PairKArgs<Ev>& k, cudaStream_t stream)
    {
    const uint32_t blocks = k.tg.nt[0] * k.tg.nt[1] * k.tg.nt[2];
    const size_t smem = pair_tile_smem(k.stage_cap, k.pitch);
    auto go = [&](auto kern) -> hb2_status
    {
        HB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, TILE_THREADS, smem, stream>>>(k);
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
