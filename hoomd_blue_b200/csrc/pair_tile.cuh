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

constexpr int PT_HCAP = 448; // home particles of a tile resolved through the shared-memory table

inline size_t pair_tile_smem(uint32_t stage_cap, uint32_t pitch)
    {
    return ((sizeof(TileLayout) + 15) & ~(size_t)15) + sizeof(float4) * stage_cap
           + sizeof(uint16_t) * TILE_WARPS * 2 * 4 * pitch;
    }

template<class Ev, int MODE, bool VIRIAL, bool ONETYPE, bool ENERGY>
__global__ void __launch_bounds__(TILE_THREADS, (VIRIAL || !ONETYPE) ? 3 : PAIR_TILE_MIN_BLOCKS)
    pair_force_tile_kernel(const PairKArgs<Ev> a)
    {
    typedef typename Ev::param_type param_type;
    if (a.skip_flag && *a.skip_flag)
        return;
    __shared__ param_type s_params[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ float s_rcutsq[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ float s_ronsq[ONETYPE ? 1 : PAIR_MAX_TYPES_SMEM * PAIR_MAX_TYPES_SMEM];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_hgi[PT_HCAP];
    __shared__ uint16_t s_hli[PT_HCAP];
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
    // ---- 1. stage the stencil: warp 0 issues one bulk copy per stencil cell
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
    const bool shifting = L.any_shift != 0u;
    const bool tabled = nhome <= (uint32_t)PT_HCAP;

    // home slot hs of the tile -> (staged slot, global index)
    auto resolve = [&](const uint32_t hs, uint32_t& li, uint32_t& gi)
    {
        uint32_t c = (hs >= L.hoff[4]) ? 4u : 0u;
        c += (hs >= L.hoff[c + 2]) ? 2u : 0u;
        c += (hs >= L.hoff[c + 1]) ? 1u : 0u;
        const uint32_t sc = L.hcell[c], k = hs - L.hoff[c];
        li = L.off[sc] + k;
        gi = L.gstart[sc] + k;
    };
    if (tabled)
        for (uint32_t hs = threadIdx.x; hs < nhome; hs += TILE_THREADS)
            {
            uint32_t li, gi;
            resolve(hs, li, gi);
            s_hgi[hs] = gi;
            s_hli[hs] = (uint16_t)li;
            }
    // the positions have landed (one warp polls the mbarrier; the others sleep in the block barrier)
    if (w == 0)
        mbar_wait(&s_bar, 0);
    __syncthreads();
    auto home = [&](const uint32_t hs, uint32_t& li, uint32_t& gi)
    {
        if (tabled)
            {
            li = s_hli[hs];
            gi = s_hgi[hs];
            }
        else
            resolve(hs, li, gi);
    };

    // ---- 2. rows: 16-byte asynchronous copies (cp.async, LDGSTS) into a per-warp double buffer, requested
    // one pass ahead (their length n_neigh two passes ahead). Lane `sub` of a particle copies quarter
    // sub & 3 of chunks sub >> 2, sub >> 2 + 2, ...: one instruction moves two 64-byte chunks per particle.
    const uint32_t stage_s = smem_addr(stage);
    uint16_t* const rowbuf = rowbuf_all + (size_t)w * (2u * 4u * a.pitch);
    const uint32_t rowbuf_s = smem_addr(rowbuf);
    const uint32_t pitch_b = a.pitch * 2u; // bytes per row
    auto load_n = [&](const uint32_t q) -> uint32_t
    {
        const uint32_t hs = 4u * q + grp;
        if (q >= npass || hs >= nhome)
            return 0u;
        uint32_t li, gi;
        home(hs, li, gi);
        return __ldg(a.n_neigh + gi);
    };
    auto request_rows = [&](const uint32_t q, const uint32_t nq)
    {
        const uint32_t hs = 4u * q + grp;
        if (q < npass && hs < nhome)
            {
            uint32_t li, gi;
            home(hs, li, gi);
            const uint32_t b = (q / TILE_WARPS) & 1u;
            const uint32_t nch = (nq + 31u) >> 5;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(a.rows16) + (size_t)gi * pitch_b
                                       + (sub >> 2) * 64u + (sub & 3u) * 16u;
            const uint32_t dst = rowbuf_s + (b * 4u + grp) * pitch_b + (sub >> 2) * 64u + (sub & 3u) * 16u;
            for (uint32_t c = sub >> 2; c < nch; c += 2u)
                {
                const uint32_t o = (c - (sub >> 2)) * 64u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
                }
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint32_t q = w;
    uint32_t n = load_n(q), n_next = load_n(q + TILE_WARPS);
    request_rows(q, n);

    while (q < npass)
        {
        const uint32_t b = (q / TILE_WARPS) & 1u;
        const uint32_t qn = q + TILE_WARPS;
        // this pass's rows have landed; the other buffer was last read in pass q - 1 (the reduction's
        // shuffles at its end joined the warp), so the next pass's rows can be requested into it
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        request_rows(qn, n_next);
        const uint32_t n_next2 = load_n(qn + TILE_WARPS);

        const uint32_t hs = 4u * q + grp;
        const bool valid = hs < nhome;
        uint32_t li = 0, gi = 0;
        if (valid)
            home(hs, li, gi);
        const float4 pi = lds128(stage_s + li * 16u);
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
        float sh[3], sl[3];
        const float pc[3] = {pi.x, pi.y, pi.z};
        const float hc[3] = {a.box.hi[0], a.box.hi[1], a.box.hi[2]};
        if (shifting)
            {
#pragma unroll
            for (int c = 0; c < 3; c++)
                {
                const double m = pc[c] >= 0.f ? (double)pc[c] - a.boxd.L[c] : (double)pc[c] + a.boxd.L[c];
                sh[c] = (float)m;
                sl[c] = (float)(m - (double)sh[c]);
                }
            }
        auto separation_shift = [&](const float4 qq, float& dx, float& dy, float& dz)
        {
            float d[3];
            const float qc[3] = {qq.x, qq.y, qq.z};
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
        auto rsq_of = [](const float dx, const float dy, const float dz)
        { return fmaf(dz, dz, fmaf(dy, dy, dx * dx)); };

        const uint32_t idx_s = rowbuf_s + (b * 4u + grp) * pitch_b + sub * 8u;
        uint2 wc = nchunks ? lds64(idx_s) : make_uint2(0u, 0u);
        for (uint32_t c = 0; c < nch_warp; c++)
            {
            uint2 wn = make_uint2(0u, 0u);
            if (c + 1u < nchunks)
                wn = lds64(idx_s + (c + 1u) * 64u);
            // row entries are staged slots times 16: byte offsets into the staged array
            float4 pj[4];
            pj[0] = lds128(stage_s + (wc.x & 0xffffu));
            pj[1] = lds128(stage_s + (wc.x >> 16));
            pj[2] = lds128(stage_s + (wc.y & 0xffffu));
            pj[3] = lds128(stage_s + (wc.y >> 16));
            float dxs[4], dys[4], dzs[4];
            if (!shifting)
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
                {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    separation_shift(pj[u], dxs[u], dys[u], dzs[u]);
                }
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
            wc = wn;
            }
        if (near_min <= 64u)
            {
            // rare (about one pass in 1e5): some rsq came within 32 ulp of r_cut^2. Re-walk the row, decide
            // those pairs with the CPU build's exact arithmetic — the staged positions ARE the stored
            // global coordinates — and repair the pairs where fp32 and the reference disagree.
            const uint16_t* __restrict__ row = rowbuf + (size_t)(b * 4u + grp) * a.pitch + 4u * sub;
            for (uint32_t e = 0; e < nchunks * 4u; e++)
                {
                // this lane's storage slots: 4 sub .. 4 sub + 3 of every chunk
                const uint32_t jj = row[(e >> 2) * 32u + (e & 3u)];
                if (jj == 0u)
                    continue;
                const float4 qq = stage[jj >> 4];
                float dx, dy, dz;
                if (!shifting)
                    {
                    dx = pi.x - qq.x;
                    dy = pi.y - qq.y;
                    dz = pi.z - qq.z;
                    }
                else
                    separation_shift(qq, dx, dy, dz);
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
        n = n_next;
        n_next = n_next2;
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
