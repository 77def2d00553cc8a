// nlist_tile.cu — neighbour-list build in the tile-shared format (tile.cuh).
//
// Replaces NeighborListBinned::buildNlist (hoomd/md/NeighborListBinned.cc:36-179) and
// gpu_compute_nlist_binned_kernel (hoomd/md/NeighborListGPUBinned.cu:56-298); same pair set as nlist.cu
// (bit-exact with the reference's CPU build), different structure:
//
// One block per tile of 2 x 2 x 2 home cells, one warp per home cell.
//  1. tile_layout: the 4 x 4 x 4 stencil's cell ranges, laid out back to back.
//  2. Staging: every stencil cell's float4 positions go to shared memory once per EIGHT home cells
//     (the per-cell kernel staged 27 cells per home cell), already shifted into the tile's periodic
//     image.
//  3. Each warp writes the candidate table of its home cell: the 27 cells around it are 9 contiguous
//     runs of the staged array (three x-neighbours are adjacent in stencil order); the table maps a
//     running candidate number to a staged slot, so that every sweep iteration has 32 live lanes.
//  4. Sweep, two home particles at a time: lane l tests candidate 32 k + l against both (packed fp32:
//     FADD2 / FFMA2 with the candidate as the broadcast operand), the hits of each home particle are
//     compacted with ballot + popc and appended to its row in shared memory — ascending by
//     construction, no masks to expand afterwards (the mask -> row expansion was 40 % of the per-cell
//     kernel's instructions).
//  5. Candidates inside the fp32 guard band are decided in fp64 with the CPU build's operation order
//     from the stored global coordinates (dist_sq_exact) — membership is bit-exact.
//  6. Rows leave as 16-bit staged slots, lane-interleaved per 32-entry chunk (TILE_SLOT), padded with the
//     dummy slot 0, with full 128-byte stores.
#include "tile.cuh"
#include "kernels.cuh"
#include <algorithm>

namespace hb2
    {
struct TileBuildKArgs
    {
    TileBuildArgs a;
    NlConst cst;
    uint32_t vcap; // candidate-table entries per warp (multiple of 32)
    uint32_t prune; // prune the candidate tables against the home cells' bounding boxes
    };

HB2_DEVICE size_t align16(size_t x)
    {
    return (x + 15) & ~(size_t)15;
    }
static inline size_t align16h(size_t x)
    {
    return (x + 15) & ~(size_t)15;
    }
static inline uint32_t tile_vcap(uint32_t stage_cap)
    {
    // 27 of 64 stencil cells, with head-room for uneven cells
    return ((stage_cap * 9u) / 16u + 31u) & ~31u;
    }
size_t tile_build_smem(uint32_t stage_cap, uint32_t pitch, uint32_t ntypes)
    {
    const size_t b = align16h(sizeof(TileLayout)) + sizeof(float4) * stage_cap
                     + sizeof(uint16_t) * TILE_WARPS * tile_vcap(stage_cap)
                     + sizeof(uint16_t) * TILE_WARPS * 2 * pitch + sizeof(float2) * (ntypes > 1 ? ntypes * ntypes : 0);
    // (row entries are 16-bit byte offsets of float4 slots: at most 4096 slots)
    return (b <= 200 * 1024 && stage_cap <= 4096u) ? b : 0;
    }

template<bool MULTITYPE>
__global__ void __launch_bounds__(TILE_THREADS, 2) nlist_build_tile_kernel(const TileBuildKArgs ka)
    {
    const TileBuildArgs& a = ka.a;
    extern __shared__ __align__(16) unsigned char s_raw[];
    TileLayout& L = *reinterpret_cast<TileLayout*>(s_raw);
    float4* __restrict__ stage = reinterpret_cast<float4*>(s_raw + align16(sizeof(TileLayout)));
    uint16_t* tab_all = reinterpret_cast<uint16_t*>(stage + a.stage_cap);
    uint16_t* row_all = tab_all + TILE_WARPS * ka.vcap;
    float2* s_thr = reinterpret_cast<float2*>(row_all + TILE_WARPS * 2 * a.pitch);

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t total = tile_layout(L, blockIdx.x, a.tg, a.geom, a.cell_start, a.cell_end);
    const uint32_t ns = L.nk[0] * L.nk[1] * L.nk[2];
    const uint32_t hcap = tile_hcap(a.stage_cap);
    if (total > a.stage_cap || L.hoff[8] > hcap)
        {
        if (threadIdx.x == 0)
            atomicMax(&a.conditions[4], max(total, 4u * L.hoff[8]));
        return;
        }
    // ---- 2. staging, shifted into the tile's image (fp32 add: the guard band covers its rounding)
    const float band = ka.cst.band;
    for (uint32_t s = w; s < ns; s += TILE_WARPS)
        {
        const uint32_t g0 = L.gstart[s], o = L.off[s], cnt = L.off[s + 1] - o;
        const int sc = L.shift[s];
        const float sx = (float)((sc & 3) - 1) * ka.cst.bf.L[0];
        const float sy = (float)(((sc >> 2) & 3) - 1) * ka.cst.bf.L[1];
        const float sz = (float)(((sc >> 4) & 3) - 1) * ka.cst.bf.L[2];
        for (uint32_t k = lane; k < cnt; k += 32)
            {
            float4 p = __ldg(a.pos + g0 + k);
            p.x += sx;
            p.y += sy;
            p.z += sz;
            stage[o + k] = p;
            }
        }
    if (threadIdx.x == 0)
        stage[0] = make_float4(NL_FAR, NL_FAR, NL_FAR, 0.f);
    if (MULTITYPE)
        for (uint32_t t = threadIdx.x; t < a.ntypes * a.ntypes; t += TILE_THREADS)
            {
            const float rl2 = (float)a.r_listsq[t];
            s_thr[t] = make_float2(rl2 * (1.0f - band), rl2 * (1.0f + band));
            }
    __syncthreads();

    // ---- 3. this warp's home cell and its candidate table
    const uint32_t nh = (L.nk[0] - 2u) * (L.nk[1] - 2u) * (L.nk[2] - 2u);
    if ((uint32_t)w >= nh)
        return;
    const uint32_t hc = L.hcell[w];
    const uint32_t hstart = L.off[hc], hcnt = L.off[hc + 1] - hstart, hg0 = L.gstart[hc];
    if (hcnt == 0)
        return;
    uint16_t* __restrict__ tab = tab_all + w * ka.vcap;
    uint16_t* __restrict__ row0 = row_all + (size_t)(2 * w) * a.pitch;
    uint16_t* __restrict__ row1 = row0 + a.pitch;
    uint32_t seg_start = 0, seg_len = 0;
    if (lane < 9)
        {
        const uint32_t kx = hc % L.nk[0], ky = (hc / L.nk[0]) % L.nk[1], kz = hc / (L.nk[0] * L.nk[1]);
        const uint32_t s0 = ((kz + lane / 3 - 1u) * L.nk[1] + (ky + lane % 3 - 1u)) * L.nk[0] + kx - 1u;
        seg_start = L.off[s0];
        seg_len = L.off[s0 + 3] - seg_start;
        }
    uint32_t V;
    const uint32_t voff = warp_excl_scan(seg_len, lane, V);
    if (V > ka.vcap)
        {
        if (lane == 0)
            atomicMax(&a.conditions[5], V);
        return;
        }
#pragma unroll
    for (int r = 0; r < 9; r++)
        {
        const uint32_t st = __shfl_sync(0xffffffffu, seg_start, r);
        const uint32_t ln = __shfl_sync(0xffffffffu, seg_len, r);
        const uint32_t vo = __shfl_sync(0xffffffffu, voff, r);
        for (uint32_t k = lane; k < ln; k += 32)
            tab[vo + k] = (uint16_t)(st + k);
        }
    __syncwarp();
    // ---- 3b. prune the table against the bounding box of the home cell's particles: a candidate farther
    // than the largest r_list from the box is farther than that from every home particle. At a cell width of
    // r_list this drops the far parts of the 12 edge and 8 corner cells (about a quarter of the candidates).
    // fp32 on the staged coordinates — the ones the sweep uses — with the sweep's guard band as margin.
    if (ka.prune)
        {
        float blo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bhi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        for (uint32_t k = lane; k < hcnt; k += 32)
            {
            const float4 p = stage[hstart + k];
            blo[0] = fminf(blo[0], p.x);
            bhi[0] = fmaxf(bhi[0], p.x);
            blo[1] = fminf(blo[1], p.y);
            bhi[1] = fmaxf(bhi[1], p.y);
            blo[2] = fminf(blo[2], p.z);
            bhi[2] = fmaxf(bhi[2], p.z);
            }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
            for (int c = 0; c < 3; c++)
                {
                blo[c] = fminf(blo[c], __shfl_xor_sync(0xffffffffu, blo[c], off));
                bhi[c] = fmaxf(bhi[c], __shfl_xor_sync(0xffffffffu, bhi[c], off));
                }
        const float thr = (float)a.r_listsq_max * (1.0f + 2.0f * band) ;
        const uint32_t Vin = (V + 31u) & ~31u;
        uint32_t Vout = 0;
        for (uint32_t v = 0; v < Vin; v += 32)
            {
            const uint32_t e = v + lane;
            const uint32_t slot = e < V ? tab[e] : 0u;
            const float4 c = stage[slot];
            const float dx = fmaxf(fmaxf(blo[0] - c.x, c.x - bhi[0]), 0.f);
            const float dy = fmaxf(fmaxf(blo[1] - c.y, c.y - bhi[1]), 0.f);
            const float dz = fmaxf(fmaxf(blo[2] - c.z, c.z - bhi[2]), 0.f);
            const bool keep = e < V && fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= thr;
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            __syncwarp(); // every lane has read its entry before the compacted ones overwrite the front
            if (keep)
                tab[Vout + __popc(m & ((1u << lane) - 1u))] = (uint16_t)slot;
            Vout += __popc(m);
            }
        V = Vout;
        __syncwarp();
        }
    const uint32_t Vpad = (V + 31u) & ~31u;
    if (V + lane < Vpad)
        tab[V + lane] = 0;
    __syncwarp();

    const float rl2_1 = (float)a.r_listsq_single;
    const float lo1 = rl2_1 * (1.0f - band), hi1 = rl2_1 * (1.0f + band);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t stage_s = smem_u32(stage), tab_s = smem_u32(tab) + 2u * lane;
    const uint32_t row0_s = smem_u32(row0), row1_s = smem_u32(row1);
    const uint32_t cap = a.pitch;
    uint32_t max_n = 0;

    // ---- 4./5. sweep, two home particles per pass
    for (uint32_t k0 = 0; k0 < hcnt; k0 += 2)
        {
        const bool has1 = k0 + 1 < hcnt;
        const uint32_t li0 = hstart + k0, li1 = li0 + 1u;
        // (the second home particle of an odd tail is the far-away dummy: it finds no neighbours)
        float2 px, py, pz;
        uint32_t t0 = 0, t1 = 0;
            {
            const float4 p0 = lds128(stage_s + li0 * 16u), p1 = lds128(has1 ? stage_s + li1 * 16u : stage_s);
            // packed once per pass and kept as 64-bit register pairs (the empty asm hides where the halves
            // came from, or the compiler re-packs them from the two float4 loads inside the sweep)
            px = make_float2(p0.x, p1.x);
            py = make_float2(p0.y, p1.y);
            pz = make_float2(p0.z, p1.z);
            asm volatile("" : "+f"(px.x), "+f"(px.y), "+f"(py.x), "+f"(py.y), "+f"(pz.x), "+f"(pz.y));
            if (MULTITYPE)
                {
                t0 = __float_as_uint(p0.w);
                t1 = __float_as_uint(p1.w);
                }
            }
        uint32_t cnt0 = 0, cnt1 = 0;
#pragma unroll 2
        for (uint32_t v = 0; v < Vpad; v += 32)
            {
            const uint32_t idx = lds_u16(tab_s + 2u * v);
            const float4 c = lds128(stage_s + idx * 16u);
            const float2 dx = __fadd2_rn(px, make_float2(-c.x, -c.x));
            const float2 dy = __fadd2_rn(py, make_float2(-c.y, -c.y));
            const float2 dz = __fadd2_rn(pz, make_float2(-c.z, -c.z));
            const float2 r = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            float lo_a = lo1, lo_b = lo1, hi_a = hi1, hi_b = hi1;
            if (MULTITYPE)
                {
                const uint32_t tc = __float_as_uint(c.w);
                const float2 ta = s_thr[tc * a.ntypes + t0], tb = s_thr[tc * a.ntypes + t1];
                lo_a = ta.x;
                hi_a = ta.y;
                lo_b = tb.x;
                hi_b = tb.y;
                }
            // r < lo: certainly inside; lo <= r <= hi: inside the fp32 guard band, decided exactly below
            bool in0 = r.x < lo_a, in1 = r.y < lo_b;
            const bool cd0 = r.x <= hi_a, cd1 = r.y <= hi_b;
            if (__any_sync(0xffffffffu, (cd0 != in0) || (cd1 != in1)))
                {
                const bool b0 = cd0 && !in0, b1 = cd1 && !in1;
                if ((b0 || b1) && idx != 0u)
                    {
                    const float4 pj = __ldg(a.pos + tile_global_of(L, ns, idx));
                    const uint32_t tj = __float_as_uint(pj.w);
                    if (b0)
                        {
                        const double rl2d = MULTITYPE ? a.r_listsq[tj * a.ntypes + t0] : a.r_listsq_single;
                        in0 = rl2d > 0.0 && dist_sq_exact(ka.cst.bd, stage[li0], pj) <= rl2d;
                        }
                    if (b1 && has1)
                        {
                        const double rl2d = MULTITYPE ? a.r_listsq[tj * a.ntypes + t1] : a.r_listsq_single;
                        in1 = rl2d > 0.0 && dist_sq_exact(ka.cst.bd, stage[li1], pj) <= rl2d;
                        }
                    }
                }
            // a particle is not its own neighbour (hoomd/md/NeighborListBinned.cc:144)
            in0 = in0 && idx != li0;
            in1 = in1 && idx != li1;
            const uint32_t m0 = __ballot_sync(0xffffffffu, in0);
            const uint32_t m1 = __ballot_sync(0xffffffffu, in1);
            const uint32_t q0 = cnt0 + __popc(m0 & lt_mask), q1 = cnt1 + __popc(m1 & lt_mask);
            if (in0 && q0 < cap)
                sts_u16(row0_s + 2u * q0, idx);
            if (in1 && q1 < cap)
                sts_u16(row1_s + 2u * q1, idx);
            cnt0 += __popc(m0);
            cnt1 += __popc(m1);
            }
        __syncwarp();
        // ---- 6. rows to HBM: lane-interleaved chunks, two 16-bit slots per lane and store
#pragma unroll
        for (int h = 0; h < 2; h++)
            {
            if (h == 1 && !has1)
                break;
            const uint32_t n = h ? cnt1 : cnt0;
            const uint16_t* __restrict__ row = h ? row1 : row0;
            const uint32_t gi = hg0 + k0 + h;
            const uint32_t lim = min((n + 31u) & ~31u, a.pitch);
            uint32_t* __restrict__ out = reinterpret_cast<uint32_t*>(a.rows16 + (size_t)gi * a.pitch);
            for (uint32_t s = 2u * lane; s < lim; s += 64)
                {
                const uint32_t base = s & ~31u, sl = s & 31u;
                const uint32_t e0 = base + TILE_ENTRY(sl), e1 = base + TILE_ENTRY(sl + 1u);
                const uint32_t v0 = e0 < n ? row[e0] : 0u, v1 = e1 < n ? row[e1] : 0u;
                out[s >> 1] = (v0 << 4) | (v1 << 20); // staged slot times 16: a byte offset
                }
            if (lane == 0)
                a.n_neigh[gi] = n;
            max_n = max(max_n, n);
            }
        __syncwarp();
        }
    if (a.last_pos)
        for (uint32_t k = lane; k < hcnt; k += 32)
            a.last_pos[hg0 + k] = stage[hstart + k]; // home cells carry no shift: the stored position
    if (lane == 0)
        atomicMax(&a.conditions[0], max_n);
    }

// One warp per tile: the descriptor the force pass's producer warp reads instead of laying the stencil
// out again on every step (tile.cuh) — header, copy runs (stencil cells that follow each other in memory
// travel as one bulk copy) and the home records (global index, staged byte offset | row length << 16).
// Tiles the build kernel rejected (conditions[4]) are skipped: the build is repeated with more room.
__global__ void __launch_bounds__(TILE_THREADS) tile_desc_kernel(const TileBuildArgs a, const uint32_t ntiles)
    {
    __shared__ TileLayout s_L[TILE_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * TILE_WARPS + w;
    if (tile >= ntiles)
        return;
    TileLayout& L = s_L[w];
    const uint32_t total = tile_layout_warp(L, tile, a.tg, a.geom, a.cell_start, a.cell_end, lane);
    const uint32_t ns = L.nk[0] * L.nk[1] * L.nk[2];
    const uint32_t hcap = tile_hcap(a.stage_cap);
    if (total > a.stage_cap || L.hoff[8] > hcap)
        return;
    unsigned char* const desc = a.desc + (size_t)tile * tile_desc_stride(hcap);
    uint2* const desc_home = reinterpret_cast<uint2*>(desc + sizeof(TileDescHeader));
    uint4* const runs = reinterpret_cast<uint4*>(desc + sizeof(TileDescHeader) + sizeof(uint2) * (size_t)hcap);
    uint32_t nruns = 0;
#pragma unroll
    for (int h = 0; h < 2; h++)
        {
        const uint32_t s = lane + 32 * h;
        bool head = false;
        uint32_t o = 0, g0 = 0, tot = 0;
        if (s < ns)
            {
            o = L.off[s];
            g0 = L.gstart[s];
            head = s == 0u || L.gstart[s - 1] + (o - L.off[s - 1]) != g0;
            if (head)
                {
                uint32_t t = s + 1u;
                tot = L.off[t] - o;
                while (t < ns && L.gstart[t] == g0 + tot)
                    {
                    t++;
                    tot = L.off[t] - o;
                    }
                head = tot != 0u;
                }
            }
        const uint32_t hm = __ballot_sync(0xffffffffu, head);
        if (head)
            runs[nruns + __popc(hm & ((1u << lane) - 1u))] = make_uint4(g0, o, tot, 0u);
        nruns += __popc(hm);
        }
    for (uint32_t r = nruns + lane; r < 64u; r += 32u)
        runs[r] = make_uint4(0u, 0u, 0u, 0u);
    if (lane == 0)
        *reinterpret_cast<uint4*>(desc) = make_uint4(1u, L.hoff[8], L.any_shift, total);
    const uint32_t nh = (L.nk[0] - 2u) * (L.nk[1] - 2u) * (L.nk[2] - 2u);
    for (uint32_t c = 0; c < nh; c++)
        {
        const uint32_t sc = L.hcell[c], g0 = L.gstart[sc], o = L.off[sc], cnt = L.off[sc + 1] - o;
        for (uint32_t k = lane; k < cnt; k += 32)
            {
            const uint32_t n = a.n_neigh[g0 + k];
            desc_home[L.hoff[c] + k] = make_uint2(g0 + k, ((o + k) * 16u) | (min(n, 0xffffu) << 16));
            }
        }
    }

// rows16 -> global indices in nlist.cu's row layout (entry e of a row at nlist + i * pitch + e)
__global__ void __launch_bounds__(TILE_THREADS) tile_rows_to_global_kernel(const TileBuildArgs a,
                                                                          uint32_t* __restrict__ nlist,
                                                                          uint32_t dummy_index)
    {
    __shared__ TileLayout L;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    tile_layout(L, blockIdx.x, a.tg, a.geom, a.cell_start, a.cell_end);
    const uint32_t ns = L.nk[0] * L.nk[1] * L.nk[2];
    const uint32_t nh = (L.nk[0] - 2u) * (L.nk[1] - 2u) * (L.nk[2] - 2u);
    if ((uint32_t)w >= nh)
        return;
    const uint32_t hc = L.hcell[w];
    const uint32_t hcnt = L.off[hc + 1] - L.off[hc], hg0 = L.gstart[hc];
    for (uint32_t k = 0; k < hcnt; k++)
        {
        const uint32_t gi = hg0 + k;
        const uint32_t n = a.n_neigh[gi];
        const uint32_t lim = min((n + 31u) & ~31u, a.pitch);
        for (uint32_t s = lane; s < lim; s += 32)
            {
            const uint32_t v = a.rows16[(size_t)gi * a.pitch + s];
            const uint32_t e = (s & ~31u) + TILE_ENTRY(s & 31u);
            nlist[(size_t)gi * a.pitch + e] = v ? tile_global_of(L, ns, v >> 4) : dummy_index;
            }
        }
    }

hb2_status launch_nlist_build_tiles(const TileBuildArgs& a, const hb2_box& box, cudaStream_t stream)
    {
    if (a.N == 0)
        return HB2_OK;
    if (a.ntypes > 64)
        {
        set_error("nlist build: at most 64 particle types are supported");
        return HB2_ERR_INVALID;
        }
    TileBuildKArgs ka;
    ka.a = a;
    ka.cst = nl_make_const(box, a.r_list_min, a.Lmax);
    ka.vcap = tile_vcap(a.stage_cap);
    static const uint32_t prune = getenv("HB2_TILE_PRUNE") ? (uint32_t)atoi(getenv("HB2_TILE_PRUNE")) : 1u;
    ka.prune = prune;
    const size_t smem = tile_build_smem(a.stage_cap, a.pitch, a.ntypes);
    if (smem == 0)
        {
        set_error("nlist build (tiles): the stencil does not fit in shared memory");
        return HB2_ERR_INVALID;
        }
    const uint32_t blocks = a.tg.nt[0] * a.tg.nt[1] * a.tg.nt[2];
    auto go = [&](auto kern) -> hb2_status
    {
        HB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, TILE_THREADS, smem, stream>>>(ka);
        return HB2_OK;
    };
    const hb2_status st = a.ntypes > 1 ? go(nlist_build_tile_kernel<true>) : go(nlist_build_tile_kernel<false>);
    if (st != HB2_OK)
        return st;
    HB2_CHECK_LAUNCH("nlist_build_tile_kernel");
    // the tiles' descriptors for the force pass (tile.cuh): layout, copy runs, home records
    tile_desc_kernel<<<(blocks + TILE_WARPS - 1) / TILE_WARPS, TILE_THREADS, 0, stream>>>(a, blocks);
    HB2_CHECK_LAUNCH("tile_desc_kernel");
    return HB2_OK;
    }

hb2_status launch_tile_rows_to_global(const TileBuildArgs& a, uint32_t* nlist, uint32_t dummy_index,
                                      cudaStream_t stream)
    {
    if (a.N == 0)
        return HB2_OK;
    const uint32_t blocks = a.tg.nt[0] * a.tg.nt[1] * a.tg.nt[2];
    tile_rows_to_global_kernel<<<blocks, TILE_THREADS, 0, stream>>>(a, nlist, dummy_index);
    HB2_CHECK_LAUNCH("tile_rows_to_global_kernel");
    return HB2_OK;
    }
    } // namespace hb2
