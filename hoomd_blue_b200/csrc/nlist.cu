// nlist.cu — warp-cooperative stencil neighbour-list build over cell-ordered particles.
//
// Replaces NeighborListBinned::buildNlist (hoomd/md/NeighborListBinned.cc:36-179) and
// gpu_compute_nlist_binned_kernel (hoomd/md/NeighborListGPUBinned.cu:56-298).
//
// One warp owns one home cell. It stages the contents of the <= 27 stencil cells (ascending cell id,
// the order hoomd/CellList.cc:393-400 sorts cell_adj into) as one concatenated candidate array in
// shared memory with coalesced float4 loads — x-adjacent cells are contiguous in the cell-ordered
// particle array, so the stencil is at most 9 long runs. Then, for each home particle, the 32 lanes
// sweep the candidates 32 at a time: fp32 distance with a guard band, the few borderline candidates
// are re-decided in fp64 with the CPU build's exact operation order (bit-exact membership), and
// ballot/popc compaction appends the hits to the particle's row in ascending j. No atomics on the
// data path; the row order is deterministic.
#include "common.cuh"
#include "kernels.cuh"

namespace hb2
    {
constexpr int NL_WARPS = 4;  // warps (home cells) per block
constexpr int NL_CAP = 768;  // staged candidates per warp per round
constexpr int NL_MAXRUNS = 27;

struct NlWarpSmem
    {
    float4 cand[NL_CAP];            // x, y, z, particle index (bits)
    uint32_t runs[NL_MAXRUNS][2];   // [start, end) particle ranges, ascending
    uint8_t type[NL_CAP];           // candidate type (MULTITYPE only)
    uint8_t pad[8];
    };

struct NlConst
    {
    BoxF bf;
    BoxD bd;
    float band; // relative guard band of the fp32 distance test
    };

// sorted, de-duplicated periodic neighbours {c-1, c, c+1} of coordinate c on a ring of size dim
HB2_DEVICE int ring3(int c, int dim, int out[3])
    {
    if (dim == 1)
        {
        out[0] = 0;
        return 1;
        }
    if (dim == 2)
        {
        out[0] = 0;
        out[1] = 1;
        return 2;
        }
    if (c == 0)
        {
        out[0] = 0;
        out[1] = 1;
        out[2] = dim - 1;
        }
    else if (c == dim - 1)
        {
        out[0] = 0;
        out[1] = dim - 2;
        out[2] = dim - 1;
        }
    else
        {
        out[0] = c - 1;
        out[1] = c;
        out[2] = c + 1;
        }
    return 3;
    }

template<bool MULTITYPE>
__global__ void __launch_bounds__(NL_WARPS * 32)
    nlist_build_kernel(const float4* __restrict__ pos, uint32_t N, uint3 dim,
                       const uint32_t* __restrict__ cell_start,
                       const double* __restrict__ r_listsq, uint32_t ntypes, double r_listsq_single,
                       NlConst cst, uint32_t* __restrict__ nlist, uint32_t pitch,
                       uint32_t* __restrict__ n_neigh, float4* __restrict__ last_pos,
                       uint32_t* __restrict__ conditions)
    {
    extern __shared__ __align__(16) unsigned char s_raw[];
    NlWarpSmem& sm = reinterpret_cast<NlWarpSmem*>(s_raw)[threadIdx.x >> 5];

    const int lane = threadIdx.x & 31;
    const uint32_t Nc = dim.x * dim.y * dim.z;
    const uint32_t cell = blockIdx.x * NL_WARPS + (threadIdx.x >> 5);
    if (cell >= Nc)
        return;
    const uint32_t h0 = cell_start[cell];
    const uint32_t hend = min(cell_start[cell + 1], N);
    if (h0 >= hend)
        return;

    // --- the stencil as ascending [start,end) runs; x-neighbours merge when contiguous
    int nruns = 0;
    if (lane == 0)
        {
        const int cx = cell % dim.x, cy = (cell / dim.x) % dim.y, cz = cell / (dim.x * dim.y);
        int nx[3], ny[3], nz[3];
        const int cnx = ring3(cx, dim.x, nx), cny = ring3(cy, dim.y, ny),
                  cnz = ring3(cz, dim.z, nz);
        for (int a = 0; a < cnz; a++)
            for (int b = 0; b < cny; b++)
                {
                const uint32_t row = ((uint32_t)nz[a] * dim.y + (uint32_t)ny[b]) * dim.x;
                for (int c = 0; c < cnx; c++)
                    {
                    const uint32_t s = cell_start[row + nx[c]], e = cell_start[row + nx[c] + 1];
                    if (e <= s)
                        continue;
                    if (nruns > 0 && sm.runs[nruns - 1][1] == s)
                        sm.runs[nruns - 1][1] = e; // contiguous with the previous run
                    else
                        {
                        sm.runs[nruns][0] = s;
                        sm.runs[nruns][1] = e;
                        nruns++;
                        }
                    }
                }
        }
    nruns = __shfl_sync(0xffffffffu, nruns, 0);
    __syncwarp();

    const float band = cst.band;
    const float rl2_1 = (float)r_listsq_single;
    const float rl2_lo1 = rl2_1 * (1.0f - band), rl2_hi1 = rl2_1 * (1.0f + band);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const float Lx = cst.bf.L[0], Ly = cst.bf.L[1], Lz = cst.bf.L[2];
    const float hx = cst.bf.hi[0], hy = cst.bf.hi[1], hz = cst.bf.hi[2];

    int run = 0;
    uint32_t run_off = 0; // progress inside runs[run]
    bool first_round = true;
    uint32_t max_n = 0;
    while (run < nruns)
        {
        // ---- stage up to NL_CAP candidates (coalesced float4 loads of contiguous runs)
        uint32_t ncand = 0;
        while (run < nruns && ncand < NL_CAP)
            {
            const uint32_t s = sm.runs[run][0] + run_off, e = sm.runs[run][1];
            const uint32_t take = min(e - s, (uint32_t)NL_CAP - ncand);
            for (uint32_t k = lane; k < take; k += 32)
                {
                float4 p = __ldg(pos + s + k);
                if (MULTITYPE)
                    sm.type[ncand + k] = (uint8_t)__float_as_uint(p.w);
                p.w = __uint_as_float(s + k);
                sm.cand[ncand + k] = p;
                }
            ncand += take;
            if (s + take == e)
                {
                run++;
                run_off = 0;
                }
            else
                run_off += take;
            }
        __syncwarp();

        // ---- every home particle sweeps the staged candidates
        for (uint32_t i = h0; i < hend; i++)
            {
            const float4 pi = __ldg(pos + i);
            const uint32_t ti = MULTITYPE ? __float_as_uint(pi.w) : 0u;
            uint32_t n = first_round ? 0u : n_neigh[i];
            uint32_t* __restrict__ row = nlist + (size_t)i * pitch;
            for (uint32_t k0 = 0; k0 < ncand; k0 += 32)
                {
                const uint32_t k = k0 + lane;
                bool hit = false;
                uint32_t j = 0;
                if (k < ncand)
                    {
                    const float4 c = sm.cand[k];
                    j = __float_as_uint(c.w);
                    const float dx = min_image_f(pi.x - c.x, Lx, hx);
                    const float dy = min_image_f(pi.y - c.y, Ly, hy);
                    const float dz = min_image_f(pi.z - c.z, Lz, hz);
                    const float rsq = dx * dx + dy * dy + dz * dz;
                    float lo = rl2_lo1, hi = rl2_hi1;
                    uint32_t tp = 0;
                    if (MULTITYPE)
                        {
                        tp = (uint32_t)sm.type[k] * ntypes + ti;
                        const float rl2 = (float)r_listsq[tp];
                        lo = rl2 * (1.0f - band);
                        hi = rl2 * (1.0f + band);
                        }
                    if (rsq <= hi && hi > 0.0f && j != i)
                        {
                        if (rsq < lo)
                            hit = true;
                        else
                            {
                            // borderline: decide exactly as the fp64 CPU build does
                            float4 pj = c;
                            const double rl2d = MULTITYPE ? r_listsq[tp] : r_listsq_single;
                            hit = dist_sq_exact(cst.bd, pi, pj) <= rl2d;
                            }
                        }
                    }
                const uint32_t mask = __ballot_sync(0xffffffffu, hit);
                if (hit)
                    {
                    const uint32_t slot = n + __popc(mask & lt_mask);
                    if (slot < pitch)
                        row[slot] = j;
                    }
                n += __popc(mask);
                }
            if (lane == 0)
                n_neigh[i] = n;
            max_n = max(max_n, n);
            }
        first_round = false;
        __syncwarp();
        }

    if (last_pos)
        for (uint32_t i = h0 + lane; i < hend; i += 32)
            last_pos[i] = pos[i];
    if (lane == 0)
        atomicMax(&conditions[0], max_n);
    }

hb2_status launch_nlist_build(const NlistBuildArgs& a, const hb2_box& box, cudaStream_t stream)
    {
    if (a.N == 0)
        return HB2_OK;
    if (a.ntypes > 255)
        {
        set_error("nlist build: at most 255 particle types are supported");
        return HB2_ERR_INVALID;
        }
    NlConst cst;
    cst.bf = boxf(box);
    cst.bd = boxd(box);
    // fp32 error model of rsq (DESIGN.md "guard band"): |d| carries <= 2^-23 Lmax absolute error,
    // so rsq carries <= 2 sqrt(3) r 2^-23 Lmax + 2^-22 rsq; the band is twice that, relative to the
    // smallest r_list^2 in use.
    const double rl_min = a.r_list_min > 0 ? a.r_list_min : 1.0;
    double band = 2.0 * (2.0 * 1.7320508 * ldexp(1.0, -23) * (double)a.Lmax / rl_min + ldexp(1.0, -22));
    if (band > 0.25)
        band = 0.25;
    cst.band = (float)band;

    const uint32_t Nc = a.dim.x * a.dim.y * a.dim.z;
    const uint32_t blocks = (Nc + NL_WARPS - 1) / NL_WARPS;
    const size_t smem = sizeof(NlWarpSmem) * NL_WARPS;
    if (a.ntypes > 1)
        {
        static bool attr = false;
        if (!attr)
            {
            HB2_CUDA(cudaFuncSetAttribute(nlist_build_kernel<true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
            }
        nlist_build_kernel<true><<<blocks, NL_WARPS * 32, smem, stream>>>(
            a.pos, a.N, a.dim, a.cell_start, a.r_listsq, a.ntypes, a.r_listsq_single, cst, a.nlist,
            a.pitch, a.n_neigh, a.last_pos, a.conditions);
        }
    else
        {
        static bool attr = false;
        if (!attr)
            {
            HB2_CUDA(cudaFuncSetAttribute(nlist_build_kernel<false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
            }
        nlist_build_kernel<false><<<blocks, NL_WARPS * 32, smem, stream>>>(
            a.pos, a.N, a.dim, a.cell_start, a.r_listsq, a.ntypes, a.r_listsq_single, cst, a.nlist,
            a.pitch, a.n_neigh, a.last_pos, a.conditions);
        }
    HB2_CHECK_LAUNCH("nlist_build_kernel");
    return HB2_OK;
    }
    } // namespace hb2
