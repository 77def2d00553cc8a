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

namespace hb2
    {
constexpr int NL_WARPS = 4;     // warps (home cells) per block
#ifndef NL_CMAX_DEF
#define NL_CMAX_DEF 21
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
__global__ void __launch_bounds__(NL_WARPS * 32)
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

        // ---- 3./4./5. two home particles per sweep
        for (uint32_t ia = h0; ia < hend; ia += 2)
            {
            const bool has_b = ia + 1 < hend;
            const uint32_t ib = has_b ? ia + 1 : ia;
            // the home cell is part of its own stencil with zero shift: read the home particles
            // from shared memory when this round staged them (no global latency in the loop)
            const uint32_t ka = home_off + (ia - h0), kb = home_off + (ib - h0);
            const bool sa = ka >= r0 && ka < r1, sb = kb >= r0 && kb < r1;
            const float4 pa = sa ? sm.cand[ka - r0] : __ldg(pos + ia);
            const float4 pb = sb ? sm.cand[kb - r0] : __ldg(pos + ib);
            uint32_t ta = 0u, tb = 0u;
            if (MULTITYPE)
                {
                ta = sa ? (uint32_t)sm.type[ka - r0] : __float_as_uint(pa.w);
                tb = sb ? (uint32_t)sm.type[kb - r0] : __float_as_uint(pb.w);
                }

            // Distances are non-negative floats, so their bit patterns order like integers: the sign
            // bit of (bits(r) - bits(threshold)) is the comparison result, and a funnel shift pushes it
            // into the per-lane mask — two integer instructions per test, no predicates, no branches.
            // (The masks come out bit-reversed: candidate t lands in bit C-1-t; fixed up after the loop.)
            uint32_t in_a = 0, in_b = 0, cd_a = 0, cd_b = 0;
            int lo_a = __float_as_int(rl2_lo1), hi_a = __float_as_int(rl2_hi1) + 1;
            int lo_b = lo_a, hi_b = hi_a;
            uint32_t hit_a, hit_b, may_a, may_b;
            if constexpr (!MULTITYPE && !SMALLDIM)
                {
                // One-type fast path. Both home particles ride in the two halves of Blackwell's packed
                // fp32 instructions (FADD2 / FMUL2 / FFMA2 with the candidate as a broadcast scalar
                // operand): 6 instead of 12 fp instructions per candidate. The guard band is tracked as
                // a running unsigned minimum of bits(r) - bits(lo) instead of a second mask (r < lo
                // wraps to a huge value and drops out), so a test costs three integer instructions.
                const float2 pxx = make_float2(pa.x, pb.x), pyy = make_float2(pa.y, pb.y),
                             pzz = make_float2(pa.z, pb.z);
                uint32_t m_a = 0xffffffffu, m_b = 0xffffffffu;
#pragma unroll 4
                for (uint32_t t = 0; t < C; t++)
                    {
                    const float4 c = mine[t];
                    const float2 dx = __fadd2_rn(pxx, make_float2(-c.x, -c.x));
                    const float2 dy = __fadd2_rn(pyy, make_float2(-c.y, -c.y));
                    const float2 dz = __fadd2_rn(pzz, make_float2(-c.z, -c.z));
                    const float2 r = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                    const uint32_t ka2 = (uint32_t)(__float_as_int(r.x) - lo_a);
                    const uint32_t kb2 = (uint32_t)(__float_as_int(r.y) - lo_a);
                    in_a = __funnelshift_l(ka2, in_a, 1); // r < lo : certainly inside
                    in_b = __funnelshift_l(kb2, in_b, 1);
                    m_a = min(m_a, ka2);
                    m_b = min(m_b, kb2);
                    }
                hit_a = __brev(in_a) >> (32u - C);
                hit_b = __brev(in_b) >> (32u - C);
                may_a = 0;
                may_b = 0;
                const uint32_t W = (uint32_t)(hi_a - lo_a); // lo <= r <= hi  <=>  k < W
                if (m_a < W || m_b < W)
                    {
                    // rare: find which candidates sit in the band
                    for (uint32_t t = 0; t < C; t++)
                        {
                        const float4 c = mine[t];
                        const float ax = pa.x - c.x, ay = pa.y - c.y, az = pa.z - c.z;
                        const float bx = pb.x - c.x, by = pb.y - c.y, bz = pb.z - c.z;
                        const uint32_t ka2
                            = (uint32_t)(__float_as_int(fmaf(az, az, fmaf(ay, ay, ax * ax))) - lo_a);
                        const uint32_t kb2
                            = (uint32_t)(__float_as_int(fmaf(bz, bz, fmaf(by, by, bx * bx))) - lo_a);
                        if (ka2 < W)
                            may_a |= 1u << t;
                        if (kb2 < W)
                            may_b |= 1u << t;
                        }
                    }
                }
            else
                {
#pragma unroll 4
                for (uint32_t t = 0; t < C; t++)
                    {
                    const float4 c = mine[t];
                    float ax = pa.x - c.x, ay = pa.y - c.y, az = pa.z - c.z;
                    float bx = pb.x - c.x, by = pb.y - c.y, bz = pb.z - c.z;
                    if (SMALLDIM)
                        {
                        ax = min_image_f(ax, Lx, hx);
                        ay = min_image_f(ay, Ly, hy);
                        az = min_image_f(az, Lz, hz);
                        bx = min_image_f(bx, Lx, hx);
                        by = min_image_f(by, Ly, hy);
                        bz = min_image_f(bz, Lz, hz);
                        }
                    const int ra = __float_as_int(ax * ax + ay * ay + az * az);
                    const int rb = __float_as_int(bx * bx + by * by + bz * bz);
                    if (MULTITYPE)
                        {
                        const uint32_t tc = mine_t[t];
                        const float2 tha = s_thr[tc * ntypes + ta], thb = s_thr[tc * ntypes + tb];
                        lo_a = __float_as_int(tha.x);
                        hi_a = __float_as_int(tha.y) + 1;
                        lo_b = __float_as_int(thb.x);
                        hi_b = __float_as_int(thb.y) + 1;
                        }
                    in_a = __funnelshift_l((uint32_t)(ra - lo_a), in_a, 1); // r <  lo : certainly inside
                    cd_a = __funnelshift_l((uint32_t)(ra - hi_a), cd_a, 1); // r <= hi : candidate
                    in_b = __funnelshift_l((uint32_t)(rb - lo_b), in_b, 1);
                    cd_b = __funnelshift_l((uint32_t)(rb - hi_b), cd_b, 1);
                    }
                hit_a = __brev(in_a) >> (32u - C);
                hit_b = __brev(in_b) >> (32u - C);
                may_a = (__brev(cd_a) >> (32u - C)) & ~hit_a;
                may_b = (__brev(cd_b) >> (32u - C)) & ~hit_b;
                }
            // a particle is not its own neighbour (hoomd/md/NeighborListBinned.cc:144)
            if (sa)
                {
                const uint32_t ks = ka - r0;
                if (ks / C == (uint32_t)lane)
                    {
                    hit_a &= ~(1u << (ks % C));
                    may_a &= ~(1u << (ks % C));
                    }
                }
            if (sb)
                {
                const uint32_t ks = kb - r0;
                if (ks / C == (uint32_t)lane)
                    {
                    hit_b &= ~(1u << (ks % C));
                    may_b &= ~(1u << (ks % C));
                    }
                }
            // ---- 4. borderline: decide exactly as the fp64 CPU build does
            if (__any_sync(0xffffffffu, (may_a | may_b) != 0u))
                {
                while (may_a)
                    {
                    const uint32_t t = __ffs(may_a) - 1;
                    may_a &= may_a - 1;
                    const uint32_t j = __float_as_uint(mine[t].w);
                    if (j == 0xffffffffu)
                        continue;
                    const float4 pj = __ldg(pos + j);
                    const double rl2d = MULTITYPE ? r_listsq[__float_as_uint(pj.w) * ntypes + ta]
                                                  : r_listsq_single;
                    if (rl2d > 0.0 && dist_sq_exact(cst.bd, pa, pj) <= rl2d)
                        hit_a |= 1u << t;
                    }
                while (may_b)
                    {
                    const uint32_t t = __ffs(may_b) - 1;
                    may_b &= may_b - 1;
                    const uint32_t j = __float_as_uint(mine[t].w);
                    if (j == 0xffffffffu)
                        continue;
                    const float4 pj = __ldg(pos + j);
                    const double rl2d = MULTITYPE ? r_listsq[__float_as_uint(pj.w) * ntypes + tb]
                                                  : r_listsq_single;
                    if (rl2d > 0.0 && dist_sq_exact(cst.bd, pb, pj) <= rl2d)
                        hit_b |= 1u << t;
                    }
                }
            if (!has_b)
                hit_b = 0;

            // ---- 5. compaction, output-parallel: lane l of pass p writes row slot 32 p + l. It finds
            // the lane that owns that slot by binary search over the exclusive scan of the per-lane
            // hit counts (shuffles), fetches that lane's hit mask and picks the right set bit. Every
            // pass is one fully coalesced 128-byte store, ascending candidate order == ascending
            // particle index, and slots past the last hit are filled with the dummy particle's index
            // so that rows are padded to a multiple of 32 for the force kernel.
            uint32_t tot_a, tot_b;
            const uint32_t off_a = warp_excl_scan(__popc(hit_a), lane, tot_a);
            const uint32_t off_b = warp_excl_scan(__popc(hit_b), lane, tot_b);
            uint32_t base_a = 0, base_b = 0;
            if (r0 > 0)
                {
                base_a = n_neigh[ia];
                base_b = n_neigh[ib];
                }
            const bool last_round = (r1 == T);
            // Both rows are emitted in lock step: the source-lane search is a chain of five dependent
            // shuffles and the bit pick a chain of dependent integer ops, so two independent chains
            // per warp double the instruction-level parallelism of this latency-bound stage.
            uint32_t* __restrict__ row_a = nlist + (size_t)ia * pitch;
            uint32_t* __restrict__ row_b = nlist + (size_t)ib * pitch;
            // slots [base, base + tot) receive this round's hits; on the last round the tail of the
            // final 32-slot group receives the dummy index
            const uint32_t end_a = last_round ? ((base_a + tot_a + 31u) & ~31u) : base_a + tot_a;
            const uint32_t end_b = has_b ? (last_round ? ((base_b + tot_b + 31u) & ~31u) : base_b + tot_b) : 0u;
            const uint32_t s0a = base_a & ~31u, s0b = base_b & ~31u;
            const uint32_t npass = max(end_a > s0a ? (end_a - s0a + 31u) >> 5 : 0u,
                                       end_b > s0b ? (end_b - s0b + 31u) >> 5 : 0u);
            // (four chains — two passes of both rows — measured slower: 0.98 vs 0.93 ms per build, the
            // odd pass count wastes a quarter of the work)
            for (uint32_t p = 0; p < npass; p++)
                {
                const uint32_t slot_a = s0a + 32u * p + lane, slot_b = s0b + 32u * p + lane;
                const uint32_t ra = slot_a - base_a, rb = slot_b - base_b; // wanted hit ranks
                // source lane: the last lane whose exclusive offset is <= rank
                uint32_t src_a = 0, src_b = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    {
                    const uint32_t pa_ = __shfl_sync(0xffffffffu, off_a, (src_a + step) & 31);
                    const uint32_t pb_ = __shfl_sync(0xffffffffu, off_b, (src_b + step) & 31);
                    if (pa_ <= ra)
                        src_a += step;
                    if (pb_ <= rb)
                        src_b += step;
                    }
                const uint32_t ma = __shfl_sync(0xffffffffu, hit_a, src_a);
                const uint32_t mb = __shfl_sync(0xffffffffu, hit_b, src_b);
                const uint32_t oa = __shfl_sync(0xffffffffu, off_a, src_a);
                const uint32_t ob = __shfl_sync(0xffffffffu, off_b, src_b);
                const bool have_a = slot_a >= base_a && ra < tot_a;
                const bool have_b = slot_b >= base_b && rb < tot_b;
                // (a slot without a hit evaluates a harmless pick on a clamped rank)
                const uint32_t ta_ = nth_set_bit(ma, have_a ? ra - oa : 0u);
                const uint32_t tb_ = nth_set_bit(mb, have_b ? rb - ob : 0u);
                const uint32_t va = have_a ? __float_as_uint(sm.cand[src_a * C + min(ta_, C - 1u)].w) : dummy_index;
                const uint32_t vb = have_b ? __float_as_uint(sm.cand[src_b * C + min(tb_, C - 1u)].w) : dummy_index;
                if (slot_a >= base_a && slot_a < end_a && slot_a < pitch)
                    row_a[slot_a] = va;
                if (slot_b >= base_b && slot_b < end_b && slot_b < pitch)
                    row_b[slot_b] = vb;
                }
            if (lane == 0)
                {
                n_neigh[ia] = base_a + tot_a;
                if (has_b)
                    n_neigh[ib] = base_b + tot_b;
                }
            max_n = max(max_n, max(base_a + tot_a, base_b + tot_b));
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
    if (a.ntypes > 1)
        return small ? launch_variant<true, true>(a, cst, stream)
                     : launch_variant<true, false>(a, cst, stream);
    return small ? launch_variant<false, true>(a, cst, stream)
                 : launch_variant<false, false>(a, cst, stream);
    }
    } // namespace hb2
