// nlist.cu — warp-cooperative stencil neighbour-list build over cell-ordered particles.
//
// Replaces NeighborListBinned::buildNlist (hoomd/md/NeighborListBinned.cc:36-179) and
// gpu_compute_nlist_binned_kernel (hoomd/md/NeighborListGPUBinned.cu:56-298).
//
// One warp owns one home cell.
//  1. Stencil: lane l < 27 resolves its stencil cell (ascending cell id, the order
//     hoomd/CellList.cc:393-400 sorts cell_adj into), reads that cell's [start, end) from cell_start
//     and its periodic image shift; a warp scan lays the cells out back to back.
//  2. Staging: the cells' float4 positions are copied into shared memory with coalesced loads,
//     already shifted into the home cell's image (no minimum-image arithmetic in the sweep) and with
//     the particle index in .w; the tail is padded with far-away dummies up to 32*C candidates.
//  3. Sweep: lane l owns candidates [l*C, l*C + C) (C odd => conflict-free LDS.128). For TWO home
//     particles at a time it tests its C candidates in fp32 against a guard band and records hits and
//     borderline cases as bits of per-lane registers — no ballot, no store inside the sweep.
//  4. Borderline candidates (a few per million) are re-decided in fp64 with the CPU build's exact
//     operation order, which makes the pair set bit-exact with the reference.
//  5. Compaction: one warp scan of popc(hits) per home particle, then every lane appends its own
//     hits: rows come out in ascending particle index, deterministically, without atomics.
#include "common.cuh"
#include "kernels.cuh"
#include "nlist_common.cuh"
#include <stdlib.h>
#include <algorithm>

namespace hb2
    {
constexpr int NL_WARPS = 4;     // warps (home cells) per block
#ifndef NL_CMAX_DEF
#define NL_CMAX_DEF 21
#endif
#ifndef NL_MIN_BLOCKS
#define NL_MIN_BLOCKS 5
#endif
constexpr int NL_CMAX = NL_CMAX_DEF; // candidates per lane per round (odd); 21 -> 5 blocks per SM
constexpr int NL_CAP = 32 * NL_CMAX;

template<bool MULTITYPE> struct NlWarpSmem
    {
    float4 cand[NL_CAP]; // x, y, z (image-shifted), particle index (bits)
    uint8_t type[MULTITYPE ? NL_CAP : 16];
    };

// MULTITYPE: per-type-pair r_list; SMALLDIM: some grid dimension < 3, so a stencil cell can hold
// both images of a neighbour and the sweep applies the minimum image itself.
template<bool MULTITYPE, bool SMALLDIM>
__global__ void __launch_bounds__(NL_WARPS * 32, NL_MIN_BLOCKS)
    nlist_build_kernel(const float4* __restrict__ pos, uint32_t N, DomainGeom geom,
                       const uint32_t* __restrict__ cell_start,
                       const uint32_t* __restrict__ cell_end,
                       const double* __restrict__ r_listsq, uint32_t ntypes, double r_listsq_single,
                       NlConst cst, uint32_t* __restrict__ nlist, uint32_t pitch,
                       uint32_t* __restrict__ n_neigh, float4* __restrict__ last_pos,
                       uint32_t* __restrict__ conditions, uint32_t dummy_index)
    {
    extern __shared__ __align__(16) unsigned char s_raw[];
    NlWarpSmem<MULTITYPE>& sm = reinterpret_cast<NlWarpSmem<MULTITYPE>*>(s_raw)[threadIdx.x >> 5];
    // per-type-pair (lo, hi) thresholds of the fp32 test, after the per-warp regions
    float2* s_thr = reinterpret_cast<float2*>(s_raw + sizeof(NlWarpSmem<MULTITYPE>) * NL_WARPS);

    const int lane = threadIdx.x & 31;
    const float band = cst.band;
    if (MULTITYPE)
        {
        for (uint32_t t = threadIdx.x; t < ntypes * ntypes; t += blockDim.x)
            {
            const float rl2 = (float)r_listsq[t];
            s_thr[t] = make_float2(rl2 * (1.0f - band), rl2 * (1.0f + band));
            }
        __syncthreads();
        }

    const uint3 dim = make_uint3(geom.dimt[0], geom.dimt[1], geom.dimt[2]);
    const uint32_t Nc = dim.x * dim.y * dim.z;
    const uint32_t cell = blockIdx.x * NL_WARPS + (threadIdx.x >> 5);
    if (cell >= Nc)
        return;
    // owned particles occupy [0, N) in cell order; ghost-layer cells start at or beyond N
    const uint32_t h0 = cell_start[cell];
    const uint32_t hend = min(cell_end[cell], N);
    if (h0 >= hend)
        return;

    // ---- 1. stencil cells, one per lane, ascending cell id
    const int cx = cell % dim.x, cy = (cell / dim.x) % dim.y, cz = cell / (dim.x * dim.y);
    int dummy_c, dummy_s;
    const int cnx = stencil3(cx, dim.x, geom.n[0], geom.coord[0], 0, dummy_c, dummy_s);
    const int cny = stencil3(cy, dim.y, geom.n[1], geom.coord[1], 0, dummy_c, dummy_s);
    const int cnz = stencil3(cz, dim.z, geom.n[2], geom.coord[2], 0, dummy_c, dummy_s);
    const int ncells = cnx * cny * cnz;
    uint32_t c_start = 0, c_cnt = 0;
    float shx = 0.f, shy = 0.f, shz = 0.f;
    bool is_home = false;
    if (lane < ncells)
        {
        const int kx = lane % cnx, ky = (lane / cnx) % cny, kz = lane / (cnx * cny);
        int nx, ny, nz, sx, sy, sz;
        stencil3(cx, dim.x, geom.n[0], geom.coord[0], kx, nx, sx);
        stencil3(cy, dim.y, geom.n[1], geom.coord[1], ky, ny, sy);
        stencil3(cz, dim.z, geom.n[2], geom.coord[2], kz, nz, sz);
        const uint32_t cid = ((uint32_t)nz * dim.y + (uint32_t)ny) * dim.x + (uint32_t)nx;
        c_start = __ldg(cell_start + cid);
        c_cnt = __ldg(cell_end + cid) - c_start;
        shx = (float)sx * cst.bf.L[0];
        shy = (float)sy * cst.bf.L[1];
        shz = (float)sz * cst.bf.L[2];
        is_home = (cid == cell);
        }
    uint32_t T;
    const uint32_t c_off = warp_excl_scan(c_cnt, lane, T);
    const uint32_t home_off = __shfl_sync(0xffffffffu, c_off, __ffs(__ballot_sync(0xffffffffu, is_home)) - 1);

    const float rl2_1 = (float)r_listsq_single;
    const float rl2_lo1 = rl2_1 * (1.0f - band), rl2_hi1 = rl2_1 * (1.0f + band);
    const float Lx = cst.bf.L[0], Ly = cst.bf.L[1], Lz = cst.bf.L[2];
    const float hx = cst.bf.hi[0], hy = cst.bf.hi[1], hz = cst.bf.hi[2];

    uint32_t max_n = 0;
    // ---- rounds of at most NL_CAP staged candidates (one round unless the stencil is very full)
    for (uint32_t r0 = 0; r0 < T; r0 += NL_CAP)
        {
        const uint32_t r1 = min(T, r0 + NL_CAP);
        const uint32_t ncand = r1 - r0;
        // ---- 2. staging (cells intersecting [r0, r1)): batches of 9 cells so that nine independent
        // coalesced loads are in flight per lane before the first shared-memory store
        auto stage = [&](uint32_t q, uint32_t j, float4 p, float sx, float sy, float sz)
        {
            if (MULTITYPE)
                sm.type[q - r0] = (uint8_t)__float_as_uint(p.w);
            if (!SMALLDIM)
                {
                p.x += sx;
                p.y += sy;
                p.z += sz;
                }
            p.w = __uint_as_float(j);
            sm.cand[q - r0] = p;
        };
        for (int c0 = 0; c0 < ncells; c0 += 9)
            {
            float4 v[9];
            bool ok[9];
#pragma unroll
            for (int u = 0; u < 9; u++)
                {
                const int c = min(c0 + u, 31);
                const uint32_t s = __shfl_sync(0xffffffffu, c_start, c);
                const uint32_t cnt = __shfl_sync(0xffffffffu, c_cnt, c);
                const uint32_t q = __shfl_sync(0xffffffffu, c_off, c) + lane;
                ok[u] = (c0 + u < ncells) && ((uint32_t)lane < cnt) && q >= r0 && q < r1;
                v[u] = ok[u] ? __ldg(pos + s + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
            for (int u = 0; u < 9; u++)
                {
                const int c = min(c0 + u, 31);
                const uint32_t s = __shfl_sync(0xffffffffu, c_start, c);
                const uint32_t cnt = __shfl_sync(0xffffffffu, c_cnt, c);
                const uint32_t off = __shfl_sync(0xffffffffu, c_off, c);
                const float sx = __shfl_sync(0xffffffffu, shx, c);
                const float sy = __shfl_sync(0xffffffffu, shy, c);
                const float sz = __shfl_sync(0xffffffffu, shz, c);
                if (ok[u])
                    stage(off + lane, s + lane, v[u], sx, sy, sz);
                if (c0 + u < ncells && cnt > 32) // crowded cell: the rest, one load at a time
                    for (uint32_t k = 32 + lane; k < cnt; k += 32)
                        if (off + k >= r0 && off + k < r1)
                            stage(off + k, s + k, __ldg(pos + s + k), sx, sy, sz);
                }
            }
        // lane-blocked ownership: C candidates per lane, C odd (bank-conflict-free float4 reads)
        const uint32_t C = ((ncand + 31) >> 5) | 1u;
        for (uint32_t q = ncand + lane; q < 32 * C; q += 32)
            {
            sm.cand[q] = make_float4(NL_FAR, NL_FAR, NL_FAR, __uint_as_float(0xffffffffu));
            if (MULTITYPE)
                sm.type[q] = 0;
            }
        __syncwarp();
        const float4* __restrict__ mine = sm.cand + lane * C;
        const uint8_t* __restrict__ mine_t = sm.type + lane * C;

        // ---- 3./4./5. H home particles per sweep: every candidate read from shared memory is tested
        // against all of them, and their rows are emitted in lock step (H independent dependency
        // chains: this stage is latency-bound). H = 4 was measured too: same 0.93 ms per build at C2,
        // the extra registers cancel the extra parallelism.
        constexpr int H = 2;
        for (uint32_t i0 = h0; i0 < hend; i0 += H)
            {
            uint32_t ih[H], kh[H], th[H];
            bool has[H], sh_[H];
            float4 ph[H];
#pragma unroll
            for (int h = 0; h < H; h++)
                {
                has[h] = i0 + h < hend;
                ih[h] = has[h] ? i0 + h : i0;
                // the home cell is part of its own stencil with zero shift: read the home particles
                // from shared memory when this round staged them (no global latency in the loop)
                kh[h] = home_off + (ih[h] - h0);
                sh_[h] = kh[h] >= r0 && kh[h] < r1;
                ph[h] = sh_[h] ? sm.cand[kh[h] - r0] : __ldg(pos + ih[h]);
                th[h] = 0u;
                if (MULTITYPE)
                    th[h] = sh_[h] ? (uint32_t)sm.type[kh[h] - r0] : __float_as_uint(ph[h].w);
                }

            // Distances are non-negative floats, so their bit patterns order like integers: the sign
            // bit of (bits(r) - bits(threshold)) is the comparison result, and a funnel shift pushes it
            // into the per-lane mask — no predicates, no branches.
            // (The masks come out bit-reversed: candidate t lands in bit C-1-t; fixed up after the loop.)
            uint32_t hit[H], may[H];
            const int lo1 = __float_as_int(rl2_lo1), hi1 = __float_as_int(rl2_hi1) + 1;
            if constexpr (!MULTITYPE && !SMALLDIM)
                {
                // One-type fast path. The home particles ride pairwise in the two halves of Blackwell's
                // packed fp32 instructions (FADD2 / FMUL2 / FFMA2 with the candidate as a broadcast
                // scalar operand): 6 instead of 12 fp instructions per candidate and pair. The guard
                // band is tracked as a running unsigned minimum of bits(r) - bits(lo) instead of a
                // second mask (r < lo wraps to a huge value and drops out).
                float2 pxx[H / 2], pyy[H / 2], pzz[H / 2];
#pragma unroll
                for (int g = 0; g < H / 2; g++)
                    {
                    pxx[g] = make_float2(ph[2 * g].x, ph[2 * g + 1].x);
                    pyy[g] = make_float2(ph[2 * g].y, ph[2 * g + 1].y);
                    pzz[g] = make_float2(ph[2 * g].z, ph[2 * g + 1].z);
                    }
                uint32_t in[H], mn[H];
#pragma unroll
                for (int h = 0; h < H; h++)
                    {
                    in[h] = 0u;
                    mn[h] = 0xffffffffu;
                    }
#pragma unroll 2
                for (uint32_t t = 0; t < C; t++)
                    {
                    const float4 c = mine[t];
#pragma unroll
                    for (int g = 0; g < H / 2; g++)
                        {
                        const float2 dx = __fadd2_rn(pxx[g], make_float2(-c.x, -c.x));
                        const float2 dy = __fadd2_rn(pyy[g], make_float2(-c.y, -c.y));
                        const float2 dz = __fadd2_rn(pzz[g], make_float2(-c.z, -c.z));
                        const float2 r = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                        const uint32_t k0 = (uint32_t)(__float_as_int(r.x) - lo1);
                        const uint32_t k1 = (uint32_t)(__float_as_int(r.y) - lo1);
                        in[2 * g] = __funnelshift_l(k0, in[2 * g], 1); // r < lo : certainly inside
                        in[2 * g + 1] = __funnelshift_l(k1, in[2 * g + 1], 1);
                        mn[2 * g] = min(mn[2 * g], k0);
                        mn[2 * g + 1] = min(mn[2 * g + 1], k1);
                        }
                    }
                const uint32_t W = (uint32_t)(hi1 - lo1); // lo <= r <= hi  <=>  k < W
                bool any_band = false;
#pragma unroll
                for (int h = 0; h < H; h++)
                    {
                    hit[h] = __brev(in[h]) >> (32u - C);
                    may[h] = 0;
                    any_band |= mn[h] < W;
                    }
                if (any_band)
                    {
                    // rare: find which candidates sit in the band
                    for (uint32_t t = 0; t < C; t++)
                        {
                        const float4 c = mine[t];
#pragma unroll
                        for (int h = 0; h < H; h++)
                            {
                            const float ax = ph[h].x - c.x, ay = ph[h].y - c.y, az = ph[h].z - c.z;
                            const uint32_t k
                                = (uint32_t)(__float_as_int(fmaf(az, az, fmaf(ay, ay, ax * ax))) - lo1);
                            if (k < W)
                                may[h] |= 1u << t;
                            }
                        }
                    }
                }
            else
                {
                uint32_t in[H], cd[H];
                int lo[H], hi[H];
#pragma unroll
                for (int h = 0; h < H; h++)
                    {
                    in[h] = cd[h] = 0u;
                    lo[h] = lo1;
                    hi[h] = hi1;
                    }
#pragma unroll 4
                for (uint32_t t = 0; t < C; t++)
                    {
                    const float4 c = mine[t];
                    uint32_t tc = 0;
                    if (MULTITYPE)
                        tc = mine_t[t];
#pragma unroll
                    for (int h = 0; h < H; h++)
                        {
                        float ax = ph[h].x - c.x, ay = ph[h].y - c.y, az = ph[h].z - c.z;
                        if (SMALLDIM)
                            {
                            ax = min_image_f(ax, Lx, hx);
                            ay = min_image_f(ay, Ly, hy);
                            az = min_image_f(az, Lz, hz);
                            }
                        const int r = __float_as_int(ax * ax + ay * ay + az * az);
                        if (MULTITYPE)
                            {
                            const float2 thr = s_thr[tc * ntypes + th[h]];
                            lo[h] = __float_as_int(thr.x);
                            hi[h] = __float_as_int(thr.y) + 1;
                            }
                        in[h] = __funnelshift_l((uint32_t)(r - lo[h]), in[h], 1); // r <  lo : inside
                        cd[h] = __funnelshift_l((uint32_t)(r - hi[h]), cd[h], 1); // r <= hi : candidate
                        }
                    }
#pragma unroll
                for (int h = 0; h < H; h++)
                    {
                    hit[h] = __brev(in[h]) >> (32u - C);
                    may[h] = (__brev(cd[h]) >> (32u - C)) & ~hit[h];
                    }
                }
            // a particle is not its own neighbour (hoomd/md/NeighborListBinned.cc:144)
            uint32_t may_any = 0;
#pragma unroll
            for (int h = 0; h < H; h++)
                {
                if (sh_[h])
                    {
                    const uint32_t ks = kh[h] - r0;
                    if (ks / C == (uint32_t)lane)
                        {
                        hit[h] &= ~(1u << (ks % C));
                        may[h] &= ~(1u << (ks % C));
                        }
                    }
                may_any |= may[h];
                }
            // ---- 4. borderline: decide exactly as the fp64 CPU build does
            if (__any_sync(0xffffffffu, may_any != 0u))
                {
#pragma unroll
                for (int h = 0; h < H; h++)
                    while (may[h])
                        {
                        const uint32_t t = __ffs(may[h]) - 1;
                        may[h] &= may[h] - 1;
                        const uint32_t j = __float_as_uint(mine[t].w);
                        if (j == 0xffffffffu)
                            continue;
                        const float4 pj = __ldg(pos + j);
                        const double rl2d = MULTITYPE ? r_listsq[__float_as_uint(pj.w) * ntypes + th[h]]
                                                      : r_listsq_single;
                        if (rl2d > 0.0 && dist_sq_exact(cst.bd, ph[h], pj) <= rl2d)
                            hit[h] |= 1u << t;
                        }
                }

            // ---- 5. compaction, output-parallel: lane l of pass p writes row slot 32 p + l. It finds
            // the lane that owns that slot by binary search over the exclusive scan of the per-lane
            // hit counts (shuffles), fetches that lane's hit mask and picks the right set bit. Every
            // pass is one fully coalesced 128-byte store, ascending candidate order == ascending
            // particle index, and slots past the last hit are filled with the dummy particle's index
            // so that rows are padded to a multiple of 32 for the force kernel.
            const bool last_round = (r1 == T);
            uint32_t off[H], tot[H], base[H], end[H], s0[H];
            uint32_t npass = 0;
#pragma unroll
            for (int h = 0; h < H; h++)
                {
                if (!has[h])
                    hit[h] = 0;
                off[h] = warp_excl_scan(__popc(hit[h]), lane, tot[h]);
                base[h] = r0 > 0 ? n_neigh[ih[h]] : 0u;
                // slots [base, base + tot) receive this round's hits; on the last round the tail of
                // the final 32-slot group receives the dummy index
                end[h] = has[h] ? (last_round ? ((base[h] + tot[h] + 31u) & ~31u) : base[h] + tot[h]) : 0u;
                s0[h] = base[h] & ~31u;
                npass = max(npass, end[h] > s0[h] ? (end[h] - s0[h] + 31u) >> 5 : 0u);
                }
            for (uint32_t p = 0; p < npass; p++)
                {
                uint32_t slot[H], rk[H], src[H];
#pragma unroll
                for (int h = 0; h < H; h++)
                    {
                    slot[h] = s0[h] + 32u * p + lane;
                    rk[h] = slot[h] - base[h]; // wanted hit rank
                    src[h] = 0;
                    }
                // source lane: the last lane whose exclusive offset is <= rank
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    {
                    uint32_t probe[H];
#pragma unroll
                    for (int h = 0; h < H; h++)
                        probe[h] = __shfl_sync(0xffffffffu, off[h], (src[h] + step) & 31);
#pragma unroll
                    for (int h = 0; h < H; h++)
                        if (probe[h] <= rk[h])
                            src[h] += step;
                    }
                uint32_t m[H], o[H];
#pragma unroll
                for (int h = 0; h < H; h++)
                    {
                    m[h] = __shfl_sync(0xffffffffu, hit[h], src[h]);
                    o[h] = __shfl_sync(0xffffffffu, off[h], src[h]);
                    }
#pragma unroll
                for (int h = 0; h < H; h++)
                    {
                    const bool have = slot[h] >= base[h] && rk[h] < tot[h];
                    // (a slot without a hit evaluates a harmless pick on a clamped rank)
                    const uint32_t t = nth_set_bit(m[h], have ? rk[h] - o[h] : 0u);
                    const uint32_t v = have ? __float_as_uint(sm.cand[src[h] * C + min(t, C - 1u)].w)
                                            : dummy_index;
                    if (slot[h] >= base[h] && slot[h] < end[h] && slot[h] < pitch)
                        nlist[(size_t)ih[h] * pitch + slot[h]] = v;
                    }
                }
#pragma unroll
            for (int h = 0; h < H; h++)
                {
                if (lane == 0 && has[h])
                    n_neigh[ih[h]] = base[h] + tot[h];
                if (has[h])
                    max_n = max(max_n, base[h] + tot[h]);
                }
            }
        __syncwarp();
        }

    if (last_pos)
        for (uint32_t i = h0 + lane; i < hend; i += 32)
            last_pos[i] = pos[i];
    if (lane == 0)
        atomicMax(&conditions[0], max_n);
    }

template<bool MT, bool SD>
static hb2_status launch_variant(const NlistBuildArgs& a, const NlConst& cst, cudaStream_t stream)
    {
    const uint32_t Nc = a.geom.dimt[0] * a.geom.dimt[1] * a.geom.dimt[2];
    const uint32_t blocks = (Nc + NL_WARPS - 1) / NL_WARPS;
    const size_t smem = sizeof(NlWarpSmem<MT>) * NL_WARPS + sizeof(float2) * (MT ? a.ntypes * a.ntypes : 1);
    static bool attr = false;
    static size_t attr_smem = 0;
    if (!attr || smem > attr_smem)
        {
        HB2_CUDA(cudaFuncSetAttribute(nlist_build_kernel<MT, SD>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
        attr_smem = smem;
        }
    nlist_build_kernel<MT, SD><<<blocks, NL_WARPS * 32, smem, stream>>>(
        a.pos, a.N, a.geom, a.cell_start, a.cell_end, a.r_listsq, a.ntypes, a.r_listsq_single, cst, a.nlist,
        a.pitch, a.n_neigh, a.last_pos, a.conditions, a.dummy_index);
    HB2_CHECK_LAUNCH("nlist_build_kernel");
    return HB2_OK;
    }

hb2_status launch_nlist_build(const NlistBuildArgs& a, const hb2_box& box, cudaStream_t stream)
    {
    if (a.N == 0)
        return HB2_OK;
    if (a.ntypes > 64)
        {
        set_error("nlist build: at most 64 particle types are supported");
        return HB2_ERR_INVALID;
        }
    const NlConst cst = nl_make_const(box, a.r_list_min, a.Lmax);

    // only a periodic ring of fewer than 3 cells can hold two images of one neighbour
    bool small = false;
    for (int d = 0; d < 3; d++)
        small = small || (a.geom.n[d] == 1 && a.geom.dimt[d] < 3);
    if (a.ntypes > 1)
        return small ? launch_variant<true, true>(a, cst, stream)
                     : launch_variant<true, false>(a, cst, stream);
    return small ? launch_variant<false, true>(a, cst, stream)
                 : launch_variant<false, false>(a, cst, stream);
    }
    } // namespace hb2
