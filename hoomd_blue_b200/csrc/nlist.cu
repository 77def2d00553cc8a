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
constexpr float NL_FAR = 1.0e18f;

template<bool MULTITYPE> struct NlWarpSmem
    {
    float4 cand[NL_CAP]; // x, y, z (image-shifted), particle index (bits)
    uint8_t type[MULTITYPE ? NL_CAP : 16];
    };

struct NlConst
    {
    BoxF bf;
    BoxD bd;
    float band; // relative guard band of the fp32 distance test
    };

// k-th (ascending) member of the sorted, de-duplicated periodic neighbourhood {c-1, c, c+1} of
// coordinate c on a ring of size dim, and the image shift (in units of L) that brings that cell next
// to c. Returns the neighbourhood size.
HB2_DEVICE int ring3(int c, int dim, int k, int& coord, int& shift)
    {
    shift = 0;
    if (dim == 1)
        {
        coord = 0;
        return 1;
        }
    if (dim == 2)
        {
        coord = k & 1;
        return 2;
        }
    if (c == 0)
        {
        coord = (k == 2) ? dim - 1 : k;
        shift = (k == 2) ? -1 : 0;
        }
    else if (c == dim - 1)
        {
        coord = (k == 0) ? 0 : dim - 3 + k;
        shift = (k == 0) ? 1 : 0;
        }
    else
        coord = c - 1 + k;
    return 3;
    }

// position of the q-th (0-based) set bit of m (m < 2^24 and popc(m) > q): binary search on popc
HB2_DEVICE uint32_t nth_set_bit(uint32_t m, uint32_t q)
    {
    uint32_t t = 0, c;
    c = __popc(m & 0xFFFu);
    if (q >= c)
        {
        q -= c;
        t = 12;
        m >>= 12;
        }
    c = __popc(m & 0x3Fu);
    if (q >= c)
        {
        q -= c;
        t += 6;
        m >>= 6;
        }
    c = __popc(m & 0x7u);
    if (q >= c)
        {
        q -= c;
        t += 3;
        m >>= 3;
        }
    c = m & 1u;
    if (q >= c)
        {
        q -= c;
        t += 1;
        m >>= 1;
        }
    c = m & 1u;
    if (q >= c)
        t += 1;
    return t;
    }

// The same along a dimension that may be decomposed over n ranks (this rank at grid coordinate rc):
// the local grid is then open — interior home cells 1..dim-2 always have both neighbours — and only
// the ghost-cell layers that lie across the GLOBAL periodic boundary carry an image shift.
HB2_DEVICE int stencil3(int c, int dim, int n, int rc, int k, int& coord, int& shift)
    {
    if (n > 1)
        {
        coord = c - 1 + k;
        shift = 0;
        if (coord == 0 && rc == 0)
            shift = -1;
        else if (coord == dim - 1 && rc == n - 1)
            shift = 1;
        return 3;
        }
    return ring3(c, dim, k, coord, shift);
    }

HB2_DEVICE uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total)
    {
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
        {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off)
            incl += t;
        }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
    }

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

// ------------------------------------------------------------------------------------------------
// Row-per-lane build (one particle type, every periodic ring >= 3 cells, pitch <= NR_MAX_PITCH).
//
// Same stencil and staging as above, but the sweep is transposed: LANE l OWNS HOME PARTICLE l of the
// cell and every lane walks the whole staged stencil. Candidates are staged structure-of-arrays, so
// one broadcast LDS.128 per coordinate feeds four candidates to all 32 home particles, and the
// distance arithmetic runs two candidates per packed fp32 instruction. A hit appends the candidate's
// 16-bit staged index to the lane's private row in shared memory (predicated STS + pointer bump): the
// row comes out ascending with no masks, no scans and no compaction pass — that pass was 40 % of
// the mask-based kernel's instructions. Rows are then written to HBM one after the other with full
// 128-byte coalesced stores (the staged index is translated to the particle index on the way out).
// The kernel holds ~10 warps per SM (shared memory), so everything outside the sweep is written to
// keep many independent loads in flight: all 27 stencil cells are fetched before the first is
// staged, and the write-out reads four 32-entry groups of a row before it stores the first.
// Bit-exactness: hits are collected for r <= hi (the guard band's upper edge); a lane that saw any
// r >= lo re-walks its own row and re-decides those entries in fp64 like the CPU build.
constexpr int NR_WARPS = 2;
constexpr int NR_CAP = 704; // staged candidates per round (multiple of 8)
constexpr uint32_t NR_MAX_PITCH = 256;
constexpr uint32_t NR_SLACK = 16;    // row entries beyond the pitch: one sweep trip appends <= 16

HB2_HOSTDEVICE uint32_t nr_row_stride_words(uint32_t pitch)
    {
    return ((pitch + NR_SLACK) >> 1) | 1u; // odd: the lanes' row heads fall into distinct banks
    }
HB2_HOSTDEVICE size_t nr_warp_smem(uint32_t pitch)
    {
    return 16u * (size_t)NR_CAP + 128u * (size_t)nr_row_stride_words(pitch);
    }

// rows are addressed by their 32-bit shared-window byte address (one register, one add per hit)
HB2_DEVICE void sts_u16(uint32_t addr, int v)
    {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v));
    }

// eight staged candidates (two float4 groups per coordinate), starting at staged index idx0, against
// this lane's home particle. SELF: the group may hold the home particle itself, which must not list
// itself (hoomd/md/NeighborListBinned.cc:144).
template<bool SELF>
HB2_DEVICE void nr_test8(const float4 (&X)[2], const float4 (&Y)[2], const float4 (&Z)[2], int idx0,
                         float2 npx, float2 npy, float2 npz, int hi1, int self, uint32_t& rp, int& mx)
    {
#pragma unroll
    for (int u = 0; u < 2; u++)
        {
#pragma unroll
        for (int v = 0; v < 2; v++)
            {
            const float2 cx = v ? make_float2(X[u].z, X[u].w) : make_float2(X[u].x, X[u].y);
            const float2 cy = v ? make_float2(Y[u].z, Y[u].w) : make_float2(Y[u].x, Y[u].y);
            const float2 cz = v ? make_float2(Z[u].z, Z[u].w) : make_float2(Z[u].x, Z[u].y);
            const float2 dx = __fadd2_rn(cx, npx);
            const float2 dy = __fadd2_rn(cy, npy);
            const float2 dz = __fadd2_rn(cz, npz);
            const float2 r = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            const int idx = idx0 + 4 * u + 2 * v;
            const int ra = __float_as_int(r.x), rb = __float_as_int(r.y);
            bool pa = ra < hi1, pb = rb < hi1;
            if (SELF)
                {
                pa = pa && idx != self;
                pb = pb && idx + 1 != self;
                }
            if (pa)
                {
                sts_u16(rp, idx);
                rp += 2;
                mx = max(mx, ra);
                }
            if (pb)
                {
                sts_u16(rp, idx + 1);
                rp += 2;
                mx = max(mx, rb);
                }
            }
        }
    }

// one sweep segment: steps [it0, it1) of 8 candidates each, two steps per loop trip and the
// shared-memory loads of a step issued one step ahead of its arithmetic (software pipeline: at ~2.5
// warps per scheduler the LDS latency is otherwise exposed once per step).
template<bool SELF>
HB2_DEVICE void nr_sweep(const float4* __restrict__ sx4, const float4* __restrict__ sy4,
                         const float4* __restrict__ sz4, uint32_t it0, uint32_t it1, float2 npx,
                         float2 npy, float2 npz, int hi1, int self, uint32_t& rp,
                         const uint32_t rp_lim, uint32_t& over, int& mx)
    {
    if (it0 >= it1)
        return;
    float4 XA[2], YA[2], ZA[2], XB[2], YB[2], ZB[2];
    auto load = [&](float4 (&X)[2], float4 (&Y)[2], float4 (&Z)[2], uint32_t it)
    {
#pragma unroll
        for (int u = 0; u < 2; u++)
            {
            X[u] = sx4[2 * it + u];
            Y[u] = sy4[2 * it + u];
            Z[u] = sz4[2 * it + u];
            }
    };
    auto check = [&]()
    {
        if (rp > rp_lim)
            {
            // row full: keep counting, stop advancing (the host regrows the list and rebuilds)
            over += (rp - rp_lim) >> 1;
            rp = rp_lim;
            }
    };
    uint32_t it = it0;
    load(XA, YA, ZA, it);
#pragma unroll 1
    for (; it + 1 < it1; it += 2)
        {
        load(XB, YB, ZB, it + 1);
        nr_test8<SELF>(XA, YA, ZA, (int)(8 * it), npx, npy, npz, hi1, self, rp, mx);
        load(XA, YA, ZA, min(it + 2, it1 - 1)); // (the last trip re-reads a valid step)
        nr_test8<SELF>(XB, YB, ZB, (int)(8 * it + 8), npx, npy, npz, hi1, self, rp, mx);
        check();
        }
    if (it < it1)
        {
        nr_test8<SELF>(XA, YA, ZA, (int)(8 * it), npx, npy, npz, hi1, self, rp, mx);
        check();
        }
    }

// per-lane description of one stencil cell of the warp's home cell (lane < 27)
struct NrStencil
    {
    uint32_t c_start, c_cnt;
    float shx, shy, shz;
    bool is_home;
    uint32_t h0, hend; // (uniform) the home cell's owned particles
    };

HB2_DEVICE NrStencil nr_load_stencil(uint32_t cell, uint3 dim, const DomainGeom& geom, const BoxF& bf,
                                     const uint32_t* __restrict__ cell_start,
                                     const uint32_t* __restrict__ cell_end, uint32_t N, int lane)
    {
    NrStencil st;
    st.h0 = __ldg(cell_start + cell);
    st.hend = min(__ldg(cell_end + cell), N); // ghost-layer cells start at or beyond N
    st.c_start = st.c_cnt = 0;
    st.shx = st.shy = st.shz = 0.f;
    st.is_home = false;
    if (lane < 27)
        {
        const int cx = cell % dim.x, cy = (cell / dim.x) % dim.y, cz = cell / (dim.x * dim.y);
        const int kx = lane % 3, ky = (lane / 3) % 3, kz = lane / 9;
        int nx, ny, nz, sx_, sy_, sz_;
        stencil3(cx, dim.x, geom.n[0], geom.coord[0], kx, nx, sx_);
        stencil3(cy, dim.y, geom.n[1], geom.coord[1], ky, ny, sy_);
        stencil3(cz, dim.z, geom.n[2], geom.coord[2], kz, nz, sz_);
        const uint32_t cid = ((uint32_t)nz * dim.y + (uint32_t)ny) * dim.x + (uint32_t)nx;
        st.c_start = __ldg(cell_start + cid);
        st.c_cnt = __ldg(cell_end + cid) - st.c_start;
        st.shx = (float)sx_ * bf.L[0];
        st.shy = (float)sy_ * bf.L[1];
        st.shz = (float)sz_ * bf.L[2];
        st.is_home = (cid == cell);
        }
    return st;
    }

// Persistent warps: warp w builds the rows of cells w, w + W, w + 2W, ... While it writes one cell's
// rows out it already has the next cell's 27 position loads in flight (and that cell's table
// entries were requested before the sweep), so the global-memory latency of the staging is off the
// critical path — with ~10 warps per SM it would otherwise be a third of the kernel's time.
__global__ void __launch_bounds__(NR_WARPS * 32, 5)
    nlist_build_rows_kernel(const float4* __restrict__ pos, uint32_t N, DomainGeom geom,
                            const uint32_t* __restrict__ cell_start,
                            const uint32_t* __restrict__ cell_end, double r_listsq_single,
                            NlConst cst, uint32_t* __restrict__ nlist, uint32_t pitch,
                            uint32_t* __restrict__ n_neigh, float4* __restrict__ last_pos,
                            uint32_t* __restrict__ conditions, uint32_t dummy_index)
    {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31;
    const uint32_t rstride = nr_row_stride_words(pitch);
    unsigned char* s_warp = s_raw + nr_warp_smem(pitch) * (threadIdx.x >> 5);
    float* sx = reinterpret_cast<float*>(s_warp);
    float* sy = sx + NR_CAP;
    float* sz = sy + NR_CAP;
    uint32_t* sj = reinterpret_cast<uint32_t*>(sz + NR_CAP);
    uint16_t* rows = reinterpret_cast<uint16_t*>(sj + NR_CAP);
    uint16_t* const myrow = rows + (size_t)lane * 2u * rstride;
    const float4* sx4 = reinterpret_cast<const float4*>(sx);
    const float4* sy4 = reinterpret_cast<const float4*>(sy);
    const float4* sz4 = reinterpret_cast<const float4*>(sz);

    const uint3 dim = make_uint3(geom.dimt[0], geom.dimt[1], geom.dimt[2]);
    const uint32_t Nc = dim.x * dim.y * dim.z;
    const uint32_t n_warps = gridDim.x * NR_WARPS;
    const float rl2 = (float)r_listsq_single;
    const int lo1 = __float_as_int(rl2 * (1.0f - cst.band));
    const int hi1 = __float_as_int(rl2 * (1.0f + cst.band)) + 1;

    uint32_t max_n = 0;
    uint32_t cell = blockIdx.x * NR_WARPS + (threadIdx.x >> 5);
    if (cell >= Nc)
        return;

    // the first 32 particles of each stencil cell of the NEXT cell to process, and its layout
    float4 v[27];
    NrStencil nx = nr_load_stencil(cell, dim, geom, cst.bf, cell_start, cell_end, N, lane);
    uint32_t nx_T = 0, nx_off = 0;
    auto prefetch = [&]()
    {
        nx_off = warp_excl_scan(nx.c_cnt, lane, nx_T);
        if (nx.h0 >= nx.hend)
            return; // no owned particles in that cell: nothing to stage
#pragma unroll
        for (int u = 0; u < 27; u++)
            {
            const uint32_t s = __shfl_sync(0xffffffffu, nx.c_start, u);
            const uint32_t cnt = __shfl_sync(0xffffffffu, nx.c_cnt, u);
            const uint32_t q = __shfl_sync(0xffffffffu, nx_off, u) + lane;
            v[u] = ((uint32_t)lane < cnt && q < (uint32_t)NR_CAP) ? __ldg(pos + s + lane)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
    };
    prefetch();

    for (; cell < Nc; cell += n_warps)
        {
        const NrStencil st = nx;
        const uint32_t T = nx_T, c_off = nx_off;
        const uint32_t h0 = st.h0, hend = st.hend;
        const bool cell_live = h0 < hend;
        const uint32_t cell_next = cell + n_warps;
        uint32_t home_off = 0;
        if (cell_live)
            {
            home_off = __shfl_sync(0xffffffffu, c_off, __ffs(__ballot_sync(0xffffffffu, st.is_home)) - 1);
            // ---- staging of round 0 from the prefetched registers
#pragma unroll
            for (int u = 0; u < 27; u++)
                {
                const uint32_t s = __shfl_sync(0xffffffffu, st.c_start, u);
                const uint32_t cnt = __shfl_sync(0xffffffffu, st.c_cnt, u);
                const uint32_t q = __shfl_sync(0xffffffffu, c_off, u) + lane;
                const float ox = __shfl_sync(0xffffffffu, st.shx, u);
                const float oy = __shfl_sync(0xffffffffu, st.shy, u);
                const float oz = __shfl_sync(0xffffffffu, st.shz, u);
                if ((uint32_t)lane < cnt && q < (uint32_t)NR_CAP)
                    {
                    sx[q] = v[u].x + ox;
                    sy[q] = v[u].y + oy;
                    sz[q] = v[u].z + oz;
                    sj[q] = s + lane;
                    }
                }
            }
        // the next cell's table entries: requested now, consumed after the sweep
        if (cell_next < Nc)
            nx = nr_load_stencil(cell_next, dim, geom, cst.bf, cell_start, cell_end, N, lane);
        bool prefetched = false;

        for (uint32_t r0 = 0; cell_live && r0 < T; r0 += NR_CAP)
            {
            const uint32_t r1 = min(T, r0 + NR_CAP);
            const uint32_t ncand = r1 - r0;
            const bool last_round = (r1 == T);
            // ---- what the prefetch did not cover: later rounds, and cells with more than 32 particles
            auto stage = [&](uint32_t q, uint32_t j, float4 p, float ox, float oy, float oz)
            {
                sx[q - r0] = p.x + ox;
                sy[q - r0] = p.y + oy;
                sz[q - r0] = p.z + oz;
                sj[q - r0] = j;
            };
            if (r0 > 0 || __any_sync(0xffffffffu, st.c_cnt > 32u))
                for (int u = 0; u < 27; u++)
                    {
                    const uint32_t s = __shfl_sync(0xffffffffu, st.c_start, u);
                    const uint32_t cnt = __shfl_sync(0xffffffffu, st.c_cnt, u);
                    const uint32_t off = __shfl_sync(0xffffffffu, c_off, u);
                    const float ox = __shfl_sync(0xffffffffu, st.shx, u);
                    const float oy = __shfl_sync(0xffffffffu, st.shy, u);
                    const float oz = __shfl_sync(0xffffffffu, st.shz, u);
                    for (uint32_t k = (r0 > 0 ? 0u : 32u) + lane; k < cnt; k += 32)
                        if (off + k >= r0 && off + k < r1)
                            stage(off + k, s + k, __ldg(pos + s + k), ox, oy, oz);
                    }
            const uint32_t n_it = (ncand + 7u) >> 3; // sweep steps of 8 candidates
            if (ncand + lane < 8u * n_it)
                {
                sx[ncand + lane] = sy[ncand + lane] = sz[ncand + lane] = NL_FAR;
                sj[ncand + lane] = dummy_index;
                }
            __syncwarp();

            // ---- batches of 32 home particles (one batch unless the cell is crowded)
            for (uint32_t hb = h0; hb < hend; hb += 32)
                {
                const uint32_t i = hb + lane;
                const bool has = i < hend;
                // staged entries [hs_lo, hs_lo + 32) are this batch's own particles (zero image shift)
                const int hs_lo = (int)(home_off + (hb - h0)) - (int)r0;
                const int self = hs_lo + lane;
                // an idle lane sits far away from everything, padding included
                float4 ph = make_float4(-NL_FAR, -NL_FAR, -NL_FAR, 0.f);
                if (has)
                    ph = (self >= 0 && self < (int)ncand) ? make_float4(sx[self], sy[self], sz[self], 0.f)
                                                          : __ldg(pos + i);
                const uint32_t it_a = (uint32_t)min(max(hs_lo >> 3, 0), (int)n_it);
                const uint32_t it_b = (uint32_t)min(max((hs_lo + 32 + 7) >> 3, 0), (int)n_it);

                const uint32_t rp0 = (uint32_t)__cvta_generic_to_shared(myrow);
                uint32_t rp = rp0;
                const uint32_t rp_lim = rp0 + 2u * pitch;
                uint32_t over = 0;
                int mx = 0;
                const float2 npx = make_float2(-ph.x, -ph.x), npy = make_float2(-ph.y, -ph.y),
                             npz = make_float2(-ph.z, -ph.z);
                nr_sweep<false>(sx4, sy4, sz4, 0, it_a, npx, npy, npz, hi1, self, rp, rp_lim, over, mx);
                nr_sweep<true>(sx4, sy4, sz4, it_a, it_b, npx, npy, npz, hi1, self, rp, rp_lim, over, mx);
                nr_sweep<false>(sx4, sy4, sz4, it_b, n_it, npx, npy, npz, hi1, self, rp, rp_lim, over, mx);
                uint32_t n = (rp - rp0) >> 1;
                __syncwarp(); // (also orders the sweep's asm stores before the reads below)

                // ---- guard band: entries with lo <= r <= hi are decided exactly as the fp64 CPU build
                if (mx >= lo1 && over == 0 && has)
                    {
                    uint32_t n2 = 0;
                    for (uint32_t s = 0; s < n; s++)
                        {
                        const uint32_t idx = myrow[s];
                        // the sweep's arithmetic, operation for operation
                        const float ax = __fadd_rn(sx[idx], -ph.x), ay = __fadd_rn(sy[idx], -ph.y),
                                    az = __fadd_rn(sz[idx], -ph.z);
                        const int rb = __float_as_int(
                            __fmaf_rn(az, az, __fmaf_rn(ay, ay, __fmul_rn(ax, ax))));
                        bool keep = true;
                        if (rb >= lo1)
                            keep = dist_sq_exact(cst.bd, ph, __ldg(pos + sj[idx])) <= r_listsq_single;
                        if (keep)
                            myrow[n2++] = (uint16_t)idx;
                        }
                    n = n2;
                    }
                n += over;
                const uint32_t n_out = over ? 0u : n; // an overflowed row holds nothing worth writing
                if (r0 == 0 && last_pos && has)
                    last_pos[i] = ph;
                __syncwarp();

                // the staged candidates are not needed any more after this write-out: get the next
                // cell's positions on their way
                if (last_round && hb + 32 >= hend && cell_next < Nc)
                    {
                    prefetch();
                    prefetched = true;
                    }

                // ---- rows to HBM, one after the other, 128 bytes per store instruction; the last
                // 32-slot group of a row is filled up with the dummy particle's index. Four groups are
                // read before the first is stored.
                const uint32_t base = (r0 > 0 && has) ? n_neigh[i] : 0u;
                const uint32_t nh = min(32u, hend - hb);
                const uint16_t* row = rows;
                uint32_t* out = nlist + (size_t)hb * pitch;
                for (uint32_t r = 0; r < nh; r++, row += 2u * rstride, out += pitch)
                    {
                    const uint32_t n_r = __shfl_sync(0xffffffffu, n_out, r);
                    const uint32_t b_r = r0 > 0 ? min(__shfl_sync(0xffffffffu, base, r), pitch) : 0u;
                    const uint32_t tot = b_r + n_r;
                    const uint32_t kend = min(last_round ? ((tot + 31u) & ~31u) : tot, pitch) - b_r;
                    for (uint32_t k0 = lane; k0 < kend; k0 += 128)
                        {
                        uint32_t idx[4], val[4];
#pragma unroll
                        for (int p = 0; p < 4; p++)
                            idx[p] = k0 + 32u * p < n_r ? row[k0 + 32u * p] : 0xffffu;
#pragma unroll
                        for (int p = 0; p < 4; p++)
                            val[p] = idx[p] != 0xffffu ? sj[idx[p]] : dummy_index;
#pragma unroll
                        for (int p = 0; p < 4; p++)
                            if (k0 + 32u * p < kend)
                                out[b_r + k0 + 32u * p] = val[p];
                        }
                    }
                if (has)
                    {
                    n_neigh[i] = base + n;
                    max_n = max(max_n, base + n);
                    }
                __syncwarp();
                }
            }
        if (!prefetched && cell_next < Nc)
            prefetch(); // (this cell had no owned particles)
        }

    max_n = max(max_n, __shfl_xor_sync(0xffffffffu, max_n, 16));
    max_n = max(max_n, __shfl_xor_sync(0xffffffffu, max_n, 8));
    max_n = max(max_n, __shfl_xor_sync(0xffffffffu, max_n, 4));
    max_n = max(max_n, __shfl_xor_sync(0xffffffffu, max_n, 2));
    max_n = max(max_n, __shfl_xor_sync(0xffffffffu, max_n, 1));
    if (lane == 0 && max_n)
        atomicMax(&conditions[0], max_n);
    }

static hb2_status launch_rows(const NlistBuildArgs& a, const NlConst& cst, cudaStream_t stream)
    {
    const uint32_t Nc = a.geom.dimt[0] * a.geom.dimt[1] * a.geom.dimt[2];
    const size_t smem = nr_warp_smem(a.pitch) * NR_WARPS;
    // persistent warps: as many blocks as fit on the 148 SMs at once
    const uint32_t per_sm = (uint32_t)std::max<size_t>(1, (size_t)(227 * 1024) / (smem + 1024));
    const uint32_t blocks = std::min<uint32_t>((Nc + NR_WARPS - 1) / NR_WARPS, NUM_SMS * per_sm);
    static size_t attr_smem = 0;
    if (smem > attr_smem)
        {
        HB2_CUDA(cudaFuncSetAttribute(nlist_build_rows_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
        }
    nlist_build_rows_kernel<<<blocks, NR_WARPS * 32, smem, stream>>>(
        a.pos, a.N, a.geom, a.cell_start, a.cell_end, a.r_listsq_single, cst, a.nlist, a.pitch, a.n_neigh,
        a.last_pos, a.conditions, a.dummy_index);
    HB2_CHECK_LAUNCH("nlist_build_rows_kernel");
    return HB2_OK;
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
    NlConst cst;
    cst.bf = boxf(box);
    cst.bd = boxd(box);
    // fp32 error model of rsq (DESIGN.md "guard band"): each component of the separation carries at
    // most 2^-23 Lmax absolute error (one rounding of a difference of magnitude <= Lmax, one of the
    // image shift), so rsq carries <= 2 sqrt(3) r 2^-23 Lmax + 2^-22 rsq. The band is twice that,
    // relative to the smallest r_list in use.
    const double rl_min = a.r_list_min > 0 ? a.r_list_min : 1.0;
    double band = 2.0 * (2.0 * 1.7320508 * ldexp(1.0, -23) * (double)a.Lmax / rl_min + ldexp(1.0, -22));
    if (band > 0.25)
        band = 0.25;
    cst.band = (float)band;

    // only a periodic ring of fewer than 3 cells can hold two images of one neighbour
    bool small = false;
    for (int d = 0; d < 3; d++)
        small = small || (a.geom.n[d] == 1 && a.geom.dimt[d] < 3);
    // Kernel choice: the row-per-lane build (HB2_NLIST_KERNEL=rows) needs one type, full rings and
    // rows that fit in shared memory.
    const bool rows_ok = a.ntypes == 1 && !small && a.pitch <= NR_MAX_PITCH && a.r_listsq_single > 0.0;
    // The row-per-lane kernel is opt-in: at C2 it runs level with the mask kernel (0.84 vs 0.88 ms under
    // ncu, 0.98 vs 0.94 ms per rebuild in the step loop).
    const char* pick = getenv("HB2_NLIST_KERNEL");
    if (pick && pick[0] == 'r' && rows_ok)
        return launch_rows(a, cst, stream);
    if (a.ntypes > 1)
        return small ? launch_variant<true, true>(a, cst, stream)
                     : launch_variant<true, false>(a, cst, stream);
    return small ? launch_variant<false, true>(a, cst, stream)
                 : launch_variant<false, false>(a, cst, stream);
    }
    } // namespace hb2
