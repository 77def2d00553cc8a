// binning.cu — deterministic cell binning by key sort (no atomics) and the physical re-ordering of
// the particle arrays into cell order.
//
// Replaces CellList::computeCellList (hoomd/CellList.cc:404-547), gpu_compute_cell_list_kernel
// (hoomd/CellListGPU.cu:42-171, atomicInc slots) and the optional thrust sort that makes it
// deterministic (hoomd/CellListGPU.cu:439-565). Because the particles themselves are permuted into
// cell order at every rebuild, this also subsumes SFCPackTuner (hoomd/SFCPackTuner.cc:81-117): the
// sorted order is a space-filling order with x-rows of cells contiguous in memory, and the "cell
// list" is just cell_start[] into the particle arrays (no Nc x Nmax padded xyzf copy).
#include "common.cuh"
#include "kernels.cuh"
#include "comm.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <stdlib.h>

namespace hb2
    {
// Sub-cell resolution of the sort key: particles of one cell are ordered along a 4x4x4 Morton curve,
// so that 8 consecutive particles (one 128-byte line of float4 positions) form a compact blob and the
// force kernel's position gathers share lines (DESIGN.md "gather locality").
constexpr int SUB_BITS = 6;

// Cell order along a Morton (Z-order) curve over the cell coordinates, up to 8 bits per axis: spatially
// close cells are close in memory, where the row-major order puts the y neighbours of a cell a whole x-row
// and the z neighbours a whole xy-plane away (SFCPackTuner's role, hoomd/SFCPackTuner.cc:541-649, at
// cell granularity; DESIGN.md "cell order").
HB2_HOSTDEVICE uint32_t spread3(uint32_t v) // 8 bits -> every third bit
    {
    v &= 0xffu;
    v = (v | (v << 8)) & 0x00f00fu;
    v = (v | (v << 4)) & 0x0c30c3u;
    v = (v | (v << 2)) & 0x249249u;
    return v;
    }
HB2_HOSTDEVICE uint32_t compact3(uint32_t v)
    {
    v &= 0x249249u;
    v = (v | (v >> 2)) & 0x0c30c3u;
    v = (v | (v >> 4)) & 0x00f00fu;
    v = (v | (v >> 8)) & 0xffu;
    return v;
    }
HB2_HOSTDEVICE uint32_t morton_cell(uint32_t x, uint32_t y, uint32_t z)
    {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
    }

HB2_DEVICE uint32_t morton2x3(uint32_t x, uint32_t y, uint32_t z)
    {
    // interleave three 2-bit coordinates: z1 y1 x1 z0 y0 x0
    return (x & 1u) | ((y & 1u) << 1) | ((z & 1u) << 2) | ((x & 2u) << 2) | ((y & 2u) << 3)
           | ((z & 2u) << 4);
    }

// One thread per particle: cell id with the reference's rule in fp64.
// f = (p - lo) * Linv (hoomd/BoxDim.h:276-284, zero tilt, no ghost width);
// ib = (int)(f * dim), ib == dim -> 0 when periodic (hoomd/CellList.cc:455-470).
// Along a decomposed dimension (geom.n > 1) the grid is local: g = domain_g(f) in units of the
// sub-domain width, interior cells 1..dimi from floor(g * dimi), cell 0 / dimi+1 for the ghost layers
// below / above. Ghost particles (index >= N_owned) carry bit 31 of the key so that the sort keeps
// owned particles in [0, N_owned) and ghosts behind them, both in cell order.
__global__ void __launch_bounds__(256)
    bin_kernel(const float4* __restrict__ pos, uint32_t N, uint32_t N_owned, BoxD box, DomainGeom geom,
               int curve, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
               uint32_t* __restrict__ conditions)
    {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    float4 p = pos[i];
    uint32_t key = 0;
    if (isnan(p.x) || isnan(p.y) || isnan(p.z))
        {
        atomicMax(&conditions[1], i + 1);
        }
    else
        {
        double f[3];
        f[0] = __dmul_rn(__dadd_rn((double)p.x, -box.lo[0]), box.Linv[0]);
        f[1] = __dmul_rn(__dadd_rn((double)p.y, -box.lo[1]), box.Linv[1]);
        f[2] = __dmul_rn(__dadd_rn((double)p.z, -box.lo[2]), box.Linv[2]);
        int ib[3];
        uint32_t sb[3];
        bool out = false;
#pragma unroll
        for (int d = 0; d < 3; d++)
            {
            const int dm = (int)geom.dimi[d];
            if (f[d] < -0.00001 || f[d] >= 1.00001)
                out = true;
            if (geom.n[d] == 1)
                {
                const double fc = __dmul_rn(f[d], (double)dm);
                ib[d] = (int)fc;
                // position inside the cell in quarters (ordering only; never decides membership)
                const int q = (int)((fc - (double)ib[d]) * 4.0);
                sb[d] = (uint32_t)min(max(q, 0), 3);
                if (ib[d] == dm)
                    {
                    ib[d] = 0;
                    sb[d] = 0;
                    }
                if (ib[d] < 0 || ib[d] >= dm)
                    out = true;
                }
            else
                {
                const double g = domain_g(f[d], geom.n[d], geom.coord[d]);
                const double fc = g * (double)dm;
                int cell, q;
                if (g >= 1.0)
                    {
                    cell = dm + 1;
                    q = (int)((fc - (double)dm) * 4.0);
                    }
                else if (g < 0.0)
                    {
                    cell = 0;
                    q = (int)((fc + 1.0) * 4.0);
                    }
                else
                    {
                    const int c0 = min((int)fc, dm - 1);
                    cell = c0 + 1;
                    q = (int)((fc - (double)c0) * 4.0);
                    }
                ib[d] = cell;
                sb[d] = (uint32_t)min(max(q, 0), 3);
                // an owned particle must lie inside its domain (migration runs before binning)
                if (i < N_owned && (g < 0.0 || g >= 1.0))
                    out = true;
                }
            }
        if (out)
            atomicMax(&conditions[2], i + 1);
        else
            {
            const uint32_t cell
                = curve ? morton_cell((uint32_t)ib[0], (uint32_t)ib[1], (uint32_t)ib[2])
                        : ((uint32_t)ib[2] * geom.dimt[1] + (uint32_t)ib[1]) * geom.dimt[0] + (uint32_t)ib[0];
            key = (cell << SUB_BITS) | morton2x3(sb[0], sb[1], sb[2]);
            }
        }
    if (i >= N_owned)
        key |= 0x80000000u;
    keys[i] = key;
    vals[i] = i;
    }

// [cell_start, cell_end) from the sorted keys (cell = key bits 6..30): slot k opens its cell when the
// previous slot holds another cell (or the other side of the owned / ghost split) and closes the
// previous one. Both arrays are zeroed first, so empty cells read as [0, 0).
__global__ void __launch_bounds__(256)
    cell_range_kernel(const uint32_t* __restrict__ keys_sorted, uint32_t N, int curve, uint32_t dimx,
                      uint32_t dimy, uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_end)
    {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N)
        return;
    const uint32_t key = keys_sorted[k] >> SUB_BITS;
    uint32_t cell = key & ((1u << (31 - SUB_BITS)) - 1u);
    if (curve) // the tables stay indexed by the row-major cell id
        cell = (compact3(cell >> 2) * dimy + compact3(cell >> 1)) * dimx + compact3(cell);
    if (k == 0 || (keys_sorted[k - 1] >> SUB_BITS) != key)
        cell_start[cell] = k;
    if (k == N - 1 || (keys_sorted[k + 1] >> SUB_BITS) != key)
        cell_end[cell] = k + 1;
    }

// Gather every per-particle array into cell order; refresh rtag.
__global__ void __launch_bounds__(256)
    reorder_kernel(const uint32_t* __restrict__ order, uint32_t N, const float4* __restrict__ pos_in,
                   const float4* __restrict__ vel_in, const int* __restrict__ img_in,
                   const uint32_t* __restrict__ tag_in, const float4* __restrict__ force_in,
                   float4* __restrict__ pos_out, float4* __restrict__ vel_out,
                   int* __restrict__ img_out, uint32_t* __restrict__ tag_out,
                   float4* __restrict__ force_out, uint32_t* __restrict__ rtag)
    {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N)
        return;
    uint32_t o = order[k];
    pos_out[k] = pos_in[o];
    vel_out[k] = vel_in[o];
    img_out[3 * k + 0] = img_in[3 * o + 0];
    img_out[3 * k + 1] = img_in[3 * o + 1];
    img_out[3 * k + 2] = img_in[3 * o + 2];
    uint32_t t = tag_in[o];
    tag_out[k] = t;
    if (force_in)
        force_out[k] = force_in[o];
    if (rtag)
        rtag[t] = k;
    }

// Snapshot <-> device layout on the device (the host only moves the caller's raw arrays):
// pack: (pos3, vel3, mass, typeid, image3[, tag]) -> pos4 / vel4 / image / tag (= index when no tags are
// given). With a domain decomposition (geom.n > 1 somewhere) every particle is also checked for
// ownership with domain_g, the same function the host filter and the migration use.
__global__ void __launch_bounds__(256)
    snapshot_pack_kernel(uint32_t N, const float* __restrict__ pos3, const float* __restrict__ vel3,
                         const float* __restrict__ mass, const uint32_t* __restrict__ typeid_,
                         const int* __restrict__ image3, const uint32_t* __restrict__ tag_in,
                         float4* __restrict__ pos, float4* __restrict__ vel, int* __restrict__ image,
                         uint32_t* __restrict__ tag, uint32_t* __restrict__ rtag, uint32_t ntypes,
                         BoxD box, DomainGeom geom, OwnMargin margin, uint32_t* __restrict__ flags)
    {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const uint32_t t = typeid_ ? typeid_[i] : 0u;
    if (t >= ntypes)
        atomicMax(&flags[0], i + 1);
    const float3 p = make_float3(pos3[3 * (size_t)i], pos3[3 * (size_t)i + 1], pos3[3 * (size_t)i + 2]);
    const float pc[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int d = 0; d < 3; d++)
        if (geom.n[d] > 1)
            {
            // plain fp64 arithmetic in the host filter's order (no contraction: separate roundings)
            const double f = __dmul_rn(__dadd_rn((double)pc[d], -box.lo[d]), box.Linv[d]);
            // a particle a rank hands back may have drifted just past a face since the last migration
            // (the next rebuild moves it); anything farther out means the caller passed particles of
            // other domains
            const double g = domain_g(f, geom.n[d], geom.coord[d]);
            if (!(g >= -margin.frac[d] && g < 1.0 + margin.frac[d]))
                atomicMax(&flags[1], i + 1);
            }
    pos[i] = make_float4(p.x, p.y, p.z, __uint_as_float(t));
    vel[i] = make_float4(vel3 ? vel3[3 * (size_t)i] : 0.f, vel3 ? vel3[3 * (size_t)i + 1] : 0.f,
                         vel3 ? vel3[3 * (size_t)i + 2] : 0.f, mass ? mass[i] : 1.0f);
    image[3 * (size_t)i] = image3 ? image3[3 * (size_t)i] : 0;
    image[3 * (size_t)i + 1] = image3 ? image3[3 * (size_t)i + 1] : 0;
    image[3 * (size_t)i + 2] = image3 ? image3[3 * (size_t)i + 2] : 0;
    const uint32_t tg = tag_in ? tag_in[i] : i;
    tag[i] = tg;
    if (rtag)
        rtag[tg] = i;
    }

// unpack: particle k (cell order) -> row tag[k] of the tag-ordered raw arrays (by_tag), or row k with
// the tags written alongside (local order, domain-decomposed systems)
__global__ void __launch_bounds__(256)
    snapshot_unpack_kernel(uint32_t N, const float4* __restrict__ pos, const float4* __restrict__ vel,
                           const int* __restrict__ image, const uint32_t* __restrict__ tag,
                           float* __restrict__ pos3, float* __restrict__ vel3,
                           float* __restrict__ mass, uint32_t* __restrict__ typeid_,
                           int* __restrict__ image3, uint32_t* __restrict__ tag_out, bool by_tag)
    {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N)
        return;
    const size_t t = by_tag ? tag[k] : k;
    const float4 p = pos[k], v = vel[k];
    pos3[3 * t] = p.x;
    pos3[3 * t + 1] = p.y;
    pos3[3 * t + 2] = p.z;
    typeid_[t] = __float_as_uint(p.w);
    vel3[3 * t] = v.x;
    vel3[3 * t + 1] = v.y;
    vel3[3 * t + 2] = v.z;
    mass[t] = v.w;
    image3[3 * t] = image[3 * (size_t)k];
    image3[3 * t + 1] = image[3 * (size_t)k + 1];
    image3[3 * t + 2] = image[3 * (size_t)k + 2];
    if (tag_out)
        tag_out[t] = tag[k];
    }

hb2_status snapshot_pack(uint32_t N, const float* pos3, const float* vel3, const float* mass,
                         const uint32_t* typeid_, const int* image3, const uint32_t* tag_in,
                         float4* pos, float4* vel, int* image, uint32_t* tag, uint32_t* rtag,
                         uint32_t ntypes, const hb2_box& box, const DomainGeom& geom,
                         const OwnMargin& margin, uint32_t* flags, cudaStream_t stream)
    {
    if (N == 0)
        return HB2_OK;
    snapshot_pack_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, pos3, vel3, mass, typeid_, image3,
                                                              tag_in, pos, vel, image, tag, rtag,
                                                              ntypes, boxd(box), geom, margin, flags);
    HB2_CHECK_LAUNCH("snapshot_pack_kernel");
    return HB2_OK;
    }

hb2_status snapshot_unpack(uint32_t N, const float4* pos, const float4* vel, const int* image,
                           const uint32_t* tag, float* pos3, float* vel3, float* mass,
                           uint32_t* typeid_, int* image3, uint32_t* tag_out, bool by_tag,
                           cudaStream_t stream)
    {
    if (N == 0)
        return HB2_OK;
    snapshot_unpack_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, pos, vel, image, tag, pos3, vel3,
                                                                mass, typeid_, image3, tag_out, by_tag);
    HB2_CHECK_LAUNCH("snapshot_unpack_kernel");
    return HB2_OK;
    }

static inline int key_bits(uint32_t max_cell_code, bool ghosts)
    {
    if (ghosts)
        return 32; // bit 31 splits owned / ghost particles
    int b = 1;
    while (b < 32 && (1ull << b) <= (unsigned long long)max_cell_code)
        b++;
    return b + SUB_BITS;
    }

// HB2_CELL_ORDER=rowmajor|morton. Default row-major: x-neighbouring cells are contiguous in memory, so the
// tile kernels stage a stencil with one bulk copy per x-row of cells (Morton measured level for the
// per-particle-row kernels, profiles/r02_cell_order_ab.txt, and needs at most 256 cells per axis)
static int cell_curve(const DomainGeom& geom, int requested)
    {
    static int env = -1;
    if (env < 0)
        {
        const char* e = getenv("HB2_CELL_ORDER");
        env = (e && e[0] == 'm') ? 1 : 0;
        }
    int curve = requested >= 0 ? requested : env;
    for (int d = 0; d < 3; d++)
        if (geom.dimt[d] > 256)
            curve = 0;
    return curve;
    }

size_t cell_order_tmp_bytes(uint32_t N, uint32_t Nc)
    {
    size_t sort_bytes = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k, v, (int)N, 0, 32);
    sort_bytes = (sort_bytes + 255) & ~size_t(255);
    size_t arr = ((size_t)N * sizeof(uint32_t) + 255) & ~size_t(255);
    return sort_bytes + 3 * arr; // keys, keys_alt, vals_alt (+ the caller's order[] as vals)
    }

hb2_status compute_cell_order(const float4* d_pos, uint32_t N, uint32_t N_owned, const hb2_box& box,
                              const DomainGeom& geom, uint32_t* d_order, uint32_t* d_cell_start,
                              uint32_t* d_cell_end, uint32_t* d_conditions, void* d_tmp,
                              size_t tmp_bytes, cudaStream_t stream, int curve_request)
    {
    const uint32_t Nc = geom.dimt[0] * geom.dimt[1] * geom.dimt[2];
    const int curve = cell_curve(geom, curve_request);
    const uint32_t max_code = curve ? morton_cell(geom.dimt[0] - 1, geom.dimt[1] - 1, geom.dimt[2] - 1) : Nc - 1;
    if (tmp_bytes < cell_order_tmp_bytes(N, Nc))
        {
        set_error("compute_cell_order: scratch buffer too small");
        return HB2_ERR_INVALID;
        }
    size_t arr = ((size_t)N * sizeof(uint32_t) + 255) & ~size_t(255);
    char* base = (char*)d_tmp;
    uint32_t* keys = (uint32_t*)base;
    uint32_t* keys_alt = (uint32_t*)(base + arr);
    uint32_t* vals_alt = (uint32_t*)(base + 2 * arr);
    void* sort_tmp = base + 3 * arr;
    size_t sort_bytes = tmp_bytes - 3 * arr;

    HB2_CUDA(cudaMemsetAsync(d_cell_start, 0, sizeof(uint32_t) * (size_t)Nc, stream));
    HB2_CUDA(cudaMemsetAsync(d_cell_end, 0, sizeof(uint32_t) * (size_t)Nc, stream));
    if (N > 0)
        {
        bin_kernel<<<(N + 255) / 256, 256, 0, stream>>>(d_pos, N, N_owned, boxd(box), geom, curve, keys,
                                                        d_order, d_conditions);
        HB2_CHECK_LAUNCH("bin_kernel");
        cub::DoubleBuffer<uint32_t> k(keys, keys_alt), v(d_order, vals_alt);
        HB2_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, k, v, (int)N, 0,
                                                 key_bits(max_code, N_owned < N), stream));
        count_launch(3);
        if (v.Current() != d_order)
            HB2_CUDA(cudaMemcpyAsync(d_order, v.Current(), (size_t)N * sizeof(uint32_t),
                                     cudaMemcpyDeviceToDevice, stream));
        cell_range_kernel<<<(N + 255) / 256, 256, 0, stream>>>(k.Current(), N, curve, geom.dimt[0],
                                                               geom.dimt[1], d_cell_start, d_cell_end);
        HB2_CHECK_LAUNCH("cell_range_kernel");
        }
    return HB2_OK;
    }

hb2_status reorder_particles(const uint32_t* d_order, uint32_t N, const float4* pos_in,
                             const float4* vel_in, const int* img_in, const uint32_t* tag_in,
                             const float4* force_in, float4* pos_out, float4* vel_out, int* img_out,
                             uint32_t* tag_out, float4* force_out, uint32_t* rtag,
                             cudaStream_t stream)
    {
    if (N == 0)
        return HB2_OK;
    reorder_kernel<<<(N + 255) / 256, 256, 0, stream>>>(d_order, N, pos_in, vel_in, img_in, tag_in,
                                                        force_in, pos_out, vel_out, img_out, tag_out,
                                                        force_out, rtag);
    HB2_CHECK_LAUNCH("reorder_kernel");
    return HB2_OK;
    }

    } // namespace hb2
