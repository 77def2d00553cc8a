// tile.cuh — the tile-shared neighbour-list format.
//
// A TILE is a block of up to 2 x 2 x 2 home cells. Its STENCIL is the surrounding block of up to
// 4 x 4 x 4 cells (one cell beyond the tile on every side, through the periodic boundary or into the
// ghost-cell layer of a domain decomposition). Both the list build (nlist_tile.cu) and the force pass
// (pair_tile.cuh) stage the particles of the stencil in shared memory in the SAME order — stencil cells
// in (kz, ky, kx) order, the particles of a cell in their memory order, slot 0 reserved for a far-away
// dummy — so a neighbour is named by its 16-bit slot in that staged array instead of a 32-bit global
// index: half the list traffic, and the force pass gathers neighbour positions from shared memory
// (LDS.128) instead of through L1 (the L1 wavefront pipe bound the per-particle-row kernel at 78 %,
// profiles/r01_ncu_full_v4_summary.txt).
//
// Row format: row i (global particle index, owned particles only) starts at rows16 + i * pitch, holds
// n_neigh[i] entries — the staged slot times 16, i.e. the byte offset of the neighbour's float4 in the staged
// array, which caps a tile's stencil at 4096 particles — padded with the dummy slot 0 to a multiple of 32, and within every 32-entry chunk the
// entries are stored lane-interleaved: the force kernel's lane `sub` (0..7) of a particle reads entries
// sub, sub + 8, sub + 16, sub + 24 of the chunk with ONE 8-byte load, so storage slot 4 sub + u holds
// ascending entry 8 u + sub (TILE_PERM).
//
// Semantics replaced: the same as nlist.cu / pair_kernel.cuh (hoomd/md/NeighborListBinned.cc:36-179,
// hoomd/md/PotentialPair.h:587-794); this is a second memory layout of the same full neighbour list.
#pragma once
#include "common.cuh"
#include "comm.cuh"
#include "nlist_common.cuh"

namespace hb2
    {
constexpr int TILE_MAXC = 64;    // stencil cells per tile
constexpr int TILE_THREADS = 256;
constexpr int TILE_WARPS = TILE_THREADS / 32;

// ascending entry e of a 32-entry chunk lives in storage slot TILE_SLOT(e); storage slot s holds entry
// TILE_ENTRY(s)
#define TILE_SLOT(e) ((((e) & 7u) << 2) | ((e) >> 3))
#define TILE_ENTRY(s) ((((s) & 3u) << 3) | ((s) >> 2))

// ---- per-tile descriptor: written by the list build (the layout only changes when the particles are
// re-binned), read by the force pass's producer warp, so that no step lays a stencil out again.
//   [TileDescHeader][home records: hcap x uint2][runs: TILE_MAXC x uint4]
// home record: (global index, staged byte offset | n_neigh << 16), in the tile's home order;
// run: (global index of the first particle, staged slot of the first particle, particles, 0) — the stencil
// cells that follow each other in memory travel as one bulk copy; unused runs have zero particles.
struct TileDescHeader
    {
    uint32_t valid;    // 1 (the force pass writes 0 into its shared-memory copy when the tiles ran out)
    uint32_t nhome;    // home particles
    uint32_t shifting; // some stencil cell lies across a periodic boundary
    uint32_t total;    // staged slots including the dummy slot 0
    };
HB2_HOSTDEVICE uint32_t tile_hcap(uint32_t stage_cap) // home records per tile: a quarter of the stencil's slots
    {
    return (stage_cap / 4u + 1u) & ~1u;
    }
HB2_HOSTDEVICE size_t tile_desc_stride(uint32_t hcap)
    {
    return sizeof(TileDescHeader) + sizeof(uint2) * (size_t)hcap + sizeof(uint4) * 64;
    }

struct TileGeom // passed by value to kernels
    {
    uint32_t nt[3];    // tiles along x, y, z (over the interior cells)
    uint32_t first[3]; // index of the first interior cell along each dimension (1 where decomposed)
    };

inline TileGeom tile_geom(const DomainGeom& g)
    {
    TileGeom t;
    for (int d = 0; d < 3; d++)
        {
        t.nt[d] = (g.dimi[d] + 1u) / 2u;
        t.first[d] = g.n[d] > 1 ? 1u : 0u;
        }
    return t;
    }
// the tile path needs every periodic ring to hold at least 3 cells (two images of one neighbour never
// meet in a stencil) — smaller grids take the per-particle-row kernels
inline bool tile_capable(const DomainGeom& g)
    {
    for (int d = 0; d < 3; d++)
        if (g.n[d] == 1 && g.dimt[d] < 3)
            return false;
    return true;
    }

struct TileLayout // shared memory, one per block
    {
    uint32_t off[TILE_MAXC + 1]; // staged slot of the first particle of stencil cell s (off[0] = 1)
    uint32_t gstart[TILE_MAXC];  // global index of that particle
    int shift[TILE_MAXC];        // image shift, 2 bits per component biased by 1: sx+1 | (sy+1)<<2 | (sz+1)<<4
    uint32_t nk[3];              // stencil extent per dimension (home cells + 2)
    uint32_t any_shift;          // some stencil cell lies across a periodic boundary
    uint32_t hoff[9];            // home particles: prefix over the (up to 8) home cells
    uint32_t hcell[8];           // stencil cell index of each home cell
    };

// Stencil coordinate k of a tile along one dimension -> cell coordinate and image shift (in units of
// L, added to the positions of that cell's particles to bring them next to the tile). k = 0 is the cell
// below the tile's first cell, k = 1 .. ncell the tile's own cells, k = ncell + 1 the cell above.
// Decomposed dimensions (n > 1) are open and only the ghost layers that lie across the GLOBAL periodic
// boundary carry a shift, exactly as stencil3 (nlist_common.cuh).
HB2_DEVICE void tile_dim(int c0, int ncell, int dimt, int n, int rc, int k, int& coord, int& shift)
    {
    shift = 0;
    if (k >= 1 && k <= ncell)
        {
        coord = c0 + k - 1;
        return;
        }
    coord = (k == 0) ? c0 - 1 : c0 + ncell;
    if (n > 1)
        {
        if (coord == 0 && rc == 0)
            shift = -1;
        else if (coord == dimt - 1 && rc == n - 1)
            shift = 1;
        }
    else if (coord < 0)
        {
        coord = dimt - 1;
        shift = -1;
        }
    else if (coord >= dimt)
        {
        coord = 0;
        shift = 1;
        }
    }

// One warp: resolve the stencil of tile `tile`, lay its cells out back to back (slot 0 = dummy) and
// list the home cells (interior cells: they hold owned particles only). t = lane. Ends with a
// __syncwarp(); returns the number of staged slots.
HB2_DEVICE uint32_t tile_layout_warp(TileLayout& L, uint32_t tile, const TileGeom& tg, const DomainGeom& geom,
                                     const uint32_t* __restrict__ cell_start,
                                     const uint32_t* __restrict__ cell_end, const int t)
    {
    const uint32_t tx = tile % tg.nt[0], ty = (tile / tg.nt[0]) % tg.nt[1], tz = tile / (tg.nt[0] * tg.nt[1]);
    const uint32_t tt[3] = {tx, ty, tz};
    int c0[3], nc[3], nk[3];
#pragma unroll
    for (int d = 0; d < 3; d++)
        {
        c0[d] = (int)(tg.first[d] + 2u * tt[d]);
        nc[d] = min(2, (int)(tg.first[d] + geom.dimi[d]) - c0[d]);
        nk[d] = nc[d] + 2;
        }
    const int ns = nk[0] * nk[1] * nk[2];
        {
        // lane l resolves stencil cells l and l + 32
        uint32_t cnt[2] = {0u, 0u};
#pragma unroll
        for (int h = 0; h < 2; h++)
            {
            const int s = t + 32 * h;
            if (s < ns)
                {
                const int kx = s % nk[0], ky = (s / nk[0]) % nk[1], kz = s / (nk[0] * nk[1]);
                int cx, cy, cz, sx, sy, sz;
                tile_dim(c0[0], nc[0], (int)geom.dimt[0], (int)geom.n[0], (int)geom.coord[0], kx, cx, sx);
                tile_dim(c0[1], nc[1], (int)geom.dimt[1], (int)geom.n[1], (int)geom.coord[1], ky, cy, sy);
                tile_dim(c0[2], nc[2], (int)geom.dimt[2], (int)geom.n[2], (int)geom.coord[2], kz, cz, sz);
                const uint32_t cid = ((uint32_t)cz * geom.dimt[1] + (uint32_t)cy) * geom.dimt[0] + (uint32_t)cx;
                const uint32_t g0 = __ldg(cell_start + cid);
                cnt[h] = __ldg(cell_end + cid) - g0;
                L.gstart[s] = g0;
                L.shift[s] = (sx + 1) | ((sy + 1) << 2) | ((sz + 1) << 4);
                }
            }
        // exclusive scan in stencil order: cells 0..31 first, then 32..63
        uint32_t tot0, tot1;
        const uint32_t e0 = warp_excl_scan(cnt[0], t, tot0);
        const uint32_t e1 = warp_excl_scan(cnt[1], t, tot1);
        if (t < ns)
            L.off[t] = 1u + e0;
        if (t + 32 < ns)
            L.off[t + 32] = 1u + tot0 + e1;
        const uint32_t shifted = __ballot_sync(0xffffffffu, (t < ns && L.shift[t] != 0x15)
                                                                || (t + 32 < ns && L.shift[t + 32] != 0x15));
        if (t == 0)
            {
            L.off[ns] = 1u + tot0 + tot1;
            L.nk[0] = nk[0];
            L.nk[1] = nk[1];
            L.nk[2] = nk[2];
            L.any_shift = shifted != 0u;
            }
        // home cells (kz, ky, kx in 1..nc): lane h < 8
        const int nh = nc[0] * nc[1] * nc[2];
        uint32_t hc = 0, hn = 0;
        if (t < nh)
            {
            const int hx = t % nc[0], hy = (t / nc[0]) % nc[1], hz = t / (nc[0] * nc[1]);
            hc = (uint32_t)(((hz + 1) * nk[1] + (hy + 1)) * nk[0] + (hx + 1));
            }
        __syncwarp();
        if (t < nh)
            hn = L.off[hc + 1] - L.off[hc];
        uint32_t htot;
        const uint32_t he = warp_excl_scan(hn, t, htot);
        if (t < 8)
            {
            L.hcell[t] = hc;
            L.hoff[t] = t < nh ? he : htot;
            }
        if (t == 8)
            L.hoff[8] = htot;
        }
    __syncwarp();
    return L.off[ns];
    }

// Block-wide form: warp 0 lays the tile out; ends with a __syncthreads().
HB2_DEVICE uint32_t tile_layout(TileLayout& L, uint32_t tile, const TileGeom& tg, const DomainGeom& geom,
                                const uint32_t* __restrict__ cell_start,
                                const uint32_t* __restrict__ cell_end)
    {
    if (threadIdx.x < 32)
        tile_layout_warp(L, tile, tg, geom, cell_start, cell_end, (int)threadIdx.x);
    __syncthreads();
    return L.off[L.nk[0] * L.nk[1] * L.nk[2]];
    }

// stencil cell that holds staged slot q (1 <= q < L.off[ns]): the last s with off[s] <= q
HB2_DEVICE uint32_t tile_cell_of(const TileLayout& L, uint32_t ns, uint32_t q)
    {
    uint32_t lo = 0, hi = ns; // invariant: off[lo] <= q < off[hi]
    while (hi - lo > 1u)
        {
        const uint32_t mid = (lo + hi) >> 1;
        if (L.off[mid] <= q)
            lo = mid;
        else
            hi = mid;
        }
    return lo;
    }
HB2_DEVICE uint32_t tile_global_of(const TileLayout& L, uint32_t ns, uint32_t q)
    {
    const uint32_t s = tile_cell_of(L, ns, q);
    return L.gstart[s] + (q - L.off[s]);
    }

// ---- shared-memory accesses by 32-bit shared-window address (no generic-to-shared conversion in the
// inner loops). volatile: ordered with the barriers and with each other, free to move past arithmetic.
HB2_DEVICE uint32_t smem_u32(const void* p)
    {
    return (uint32_t)__cvta_generic_to_shared(p);
    }
HB2_DEVICE float4 lds128(const uint32_t addr)
    {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
    }
HB2_DEVICE uint2 lds64(const uint32_t addr)
    {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
    }
HB2_DEVICE uint32_t lds_u16(const uint32_t addr)
    {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
    }
HB2_DEVICE void sts_u16(const uint32_t addr, const uint32_t v)
    {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v));
    }

// ---- launch interface (nlist_tile.cu)
struct TileBuildArgs
    {
    const float4* pos;
    uint32_t N; // owned particles (rows)
    DomainGeom geom;
    TileGeom tg;
    const uint32_t* cell_start;
    const uint32_t* cell_end;
    const double* r_listsq; // ntypes^2 (device), tj * ntypes + ti
    uint32_t ntypes;
    double r_listsq_single;
    double r_listsq_max; // largest r_list^2 over the type pairs (candidate pruning against a home cell's box)
    double r_list_min;
    float Lmax;
    uint16_t* rows16;
    uint32_t pitch;       // entries per row (multiple of 32)
    uint32_t* n_neigh;
    float4* last_pos;     // may be null
    uint32_t* conditions; // [0] <- max n_neigh, [4] <- staged slots needed when they exceed stage_cap,
                          // [5] <- candidates of one home cell when they exceed the table capacity
    uint32_t stage_cap;   // staged slots (float4) a block may hold
    unsigned char* desc;  // per-tile descriptors (tile_desc_stride(tile_hcap(stage_cap)) bytes each)
    };
hb2_status launch_nlist_build_tiles(const TileBuildArgs& a, const hb2_box& box, cudaStream_t stream);
// rows16 -> rows of global indices (the layout of nlist.cu, without its ascending order across cells)
hb2_status launch_tile_rows_to_global(const TileBuildArgs& a, uint32_t* nlist, uint32_t dummy_index,
                                      cudaStream_t stream);
// shared-memory bytes of the build kernel for a given capacity (0: does not fit)
size_t tile_build_smem(uint32_t stage_cap, uint32_t pitch, uint32_t ntypes);
    } // namespace hb2
