// comm.cu — domain decomposition over NCCL (see comm.cuh for the geometry and the reference map).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded by torch.distributed, else the
// system one): the library has no link-time dependency on it and loads on boxes without NCCL. The
// communicator is bootstrapped by the Python layer: rank 0 creates the unique id here, broadcasts it
// with torch.distributed, and every rank calls hb2_system_set_domain.
#include "engine.cuh"

#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <dlfcn.h>
#include <nccl.h>

namespace hb2
    {
// ---------------------------------------------------------------------------------------------
// run-time NCCL binding
// ---------------------------------------------------------------------------------------------
struct NcclApi
    {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t)
        = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    };
static NcclApi g_nccl;

static hb2_status nccl_load()
    {
    if (g_nccl.handle)
        return HB2_OK;
    const char* names[] = {getenv("HB2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
        {
        if (!n || !*n)
            continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h)
            break;
        }
    if (!h)
        {
        set_error("domain decomposition: libnccl.so.2 could not be loaded (set HB2_NCCL_LIB)");
        return HB2_ERR_COMM;
        }
#define HB2_NCCL_SYM(field, name)                                          \
    *(void**)(&g_nccl.field) = dlsym(h, name);                             \
    if (!g_nccl.field)                                                     \
        {                                                                  \
        set_error(std::string("domain decomposition: missing symbol ") + name); \
        return HB2_ERR_COMM;                                               \
        }
    HB2_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    HB2_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    HB2_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    HB2_NCCL_SYM(Send, "ncclSend")
    HB2_NCCL_SYM(Recv, "ncclRecv")
    HB2_NCCL_SYM(AllReduce, "ncclAllReduce")
    HB2_NCCL_SYM(GroupStart, "ncclGroupStart")
    HB2_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    HB2_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef HB2_NCCL_SYM
    g_nccl.handle = h;
    return HB2_OK;
    }

static hb2_status nccl_fail(ncclResult_t r, const char* what)
    {
    set_error(std::string("NCCL error: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?")
              + " in " + what);
    return HB2_ERR_COMM;
    }
#define HB2_NCCL(call)                           \
    do                                           \
        {                                        \
        ncclResult_t _r = (call);                \
        if (_r != ncclSuccess)                   \
            return nccl_fail(_r, #call);         \
        } while (0)

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
constexpr int CB = 256;

HB2_DEVICE double frac_d(const BoxD& b, float p, int d)
    {
    // hoomd/BoxDim.h:276-284 (zero tilt, no ghost width), the operation order bin_kernel uses
    return __dmul_rn(__dadd_rn((double)p, -b.lo[d]), b.Linv[d]);
    }

// migration class along dimension d: 0 stay, 1 to above, 2 to below
// (hoomd/Communicator.cc:1600-1626: a particle leaves when its fractional coordinate is outside)
__global__ void __launch_bounds__(CB)
    classify_migrate_kernel(const float4* __restrict__ pos, uint32_t N, BoxD box, int d, uint32_t n,
                            uint32_t coord, uint8_t* __restrict__ cls)
    {
    const uint32_t i = blockIdx.x * CB + threadIdx.x;
    if (i >= N)
        return;
    const float4 p = pos[i];
    const float c = d == 0 ? p.x : (d == 1 ? p.y : p.z);
    const double g = domain_g(frac_d(box, c, d), n, coord);
    cls[i] = g >= 1.0 ? 1 : (g < 0.0 ? 2 : 0);
    }

// ghost class along dimension d over owned particles and the ghosts of earlier stages:
// bit 0: within r_ghost of the upper face (goes to the rank above), bit 1: of the lower face
// (hoomd/CommunicatorGPU.cu:306-357 builds the same plan as a 6-bit mask in one pass)
__global__ void __launch_bounds__(CB)
    classify_ghost_kernel(const float4* __restrict__ pos, uint32_t Ntot, BoxD box, int d, uint32_t n,
                          uint32_t coord, double gfrac, uint8_t* __restrict__ cls)
    {
    const uint32_t i = blockIdx.x * CB + threadIdx.x;
    if (i >= Ntot)
        return;
    const float4 p = pos[i];
    const float c = d == 0 ? p.x : (d == 1 ? p.y : p.z);
    const double g = domain_g(frac_d(box, c, d), n, coord);
    uint8_t m = 0;
    if (g >= 0.0 && g < 1.0)
        {
        if (g >= 1.0 - gfrac)
            m |= 1;
        if (g < gfrac)
            m |= 2;
        }
    cls[i] = m;
    }

struct ClsEq
    {
    const uint8_t* cls;
    uint8_t want;
    bool mask;
    __host__ __device__ uint8_t operator()(uint32_t i) const
        {
        return mask ? ((cls[i] & want) != 0) : (cls[i] == want);
        }
    };

__global__ void __launch_bounds__(CB)
    gather_particles_kernel(const uint32_t* __restrict__ idx, uint32_t n,
                            const float4* __restrict__ pos, const float4* __restrict__ vel,
                            const int* __restrict__ img, const uint32_t* __restrict__ tag,
                            float4* __restrict__ pos_o, float4* __restrict__ vel_o,
                            int* __restrict__ img_o, uint32_t* __restrict__ tag_o)
    {
    const uint32_t k = blockIdx.x * CB + threadIdx.x;
    if (k >= n)
        return;
    const uint32_t i = idx[k];
    pos_o[k] = pos[i];
    if (vel_o)
        vel_o[k] = vel[i];
    if (img_o)
        {
        img_o[3 * k] = img[3 * i];
        img_o[3 * k + 1] = img[3 * i + 1];
        img_o[3 * k + 2] = img[3 * i + 2];
        }
    if (tag_o)
        tag_o[k] = tag[i];
    }

// In-place compaction after the leavers of a migration stage were copied out. `leave` holds the m
// leaving indices in ascending order; the owned range shrinks to [0, N - m). Every stayer in the tail
// [N - m, N) moves into a hole below N - m: the r-th tail stayer (ascending) takes the r-th hole
// (ascending) — as many of one as of the other. The order of the survivors is a pure function of the
// leaver set, so the run stays reproducible; the cell sort that follows re-orders everything anyway.
__global__ void __launch_bounds__(CB)
    fill_holes_kernel(const uint32_t* __restrict__ leave, uint32_t m, uint32_t N,
                      const uint8_t* __restrict__ cls, float4* __restrict__ pos,
                      float4* __restrict__ vel, int* __restrict__ img, uint32_t* __restrict__ tag)
    {
    const uint32_t k = blockIdx.x * CB + threadIdx.x;
    if (k >= m)
        return;
    const uint32_t n_stay = N - m;
    const uint32_t t = n_stay + k;
    if (cls[t] != 0)
        return;
    auto lower_bound = [&](uint32_t v)
    {
        uint32_t lo = 0, hi = m;
        while (lo < hi)
            {
            const uint32_t mid = (lo + hi) >> 1;
            if (leave[mid] < v)
                lo = mid + 1;
            else
                hi = mid;
            }
        return lo;
    };
    const uint32_t h = lower_bound(n_stay);          // holes below the new end
    const uint32_t r = k - (lower_bound(t) - h);     // stayers in [n_stay, t)
    const uint32_t dst = leave[r];
    pos[dst] = pos[t];
    vel[dst] = vel[t];
    img[3 * dst] = img[3 * t];
    img[3 * dst + 1] = img[3 * t + 1];
    img[3 * dst + 2] = img[3 * t + 2];
    tag[dst] = tag[t];
    }

// per-step ghost refresh: sendbuf[k] = pos[idx[k]] (and vel behind it)
__global__ void __launch_bounds__(CB)
    pack_ghost_kernel(const uint32_t* __restrict__ idx, uint32_t n, const float4* __restrict__ pos,
                      const float4* __restrict__ vel, float4* __restrict__ out_pos,
                      float4* __restrict__ out_vel)
    {
    const uint32_t k = blockIdx.x * CB + threadIdx.x;
    if (k >= n)
        return;
    const uint32_t i = idx[k];
    out_pos[k] = pos[i];
    if (out_vel)
        out_vel[k] = vel[i];
    }

__global__ void __launch_bounds__(CB)
    unpack_ghost_kernel(const uint32_t* __restrict__ slot, uint32_t n,
                        const float4* __restrict__ in_pos, const float4* __restrict__ in_vel,
                        float4* __restrict__ pos, float4* __restrict__ vel)
    {
    const uint32_t k = blockIdx.x * CB + threadIdx.x;
    if (k >= n)
        return;
    const uint32_t i = slot[k];
    pos[i] = in_pos[k];
    if (in_vel)
        vel[i] = in_vel[k];
    }

__global__ void __launch_bounds__(CB)
    invert_perm_kernel(const uint32_t* __restrict__ order, uint32_t n, uint32_t* __restrict__ inv)
    {
    const uint32_t k = blockIdx.x * CB + threadIdx.x;
    if (k < n)
        inv[order[k]] = k;
    }

__global__ void __launch_bounds__(CB)
    remap_kernel(uint32_t* __restrict__ idx, uint32_t n, const uint32_t* __restrict__ inv)
    {
    const uint32_t k = blockIdx.x * CB + threadIdx.x;
    if (k < n)
        idx[k] = inv[idx[k]];
    }

__global__ void __launch_bounds__(CB)
    ghost_slot_kernel(uint32_t* __restrict__ slot, uint32_t Ng, uint32_t N,
                      const uint32_t* __restrict__ inv)
    {
    const uint32_t k = blockIdx.x * CB + threadIdx.x;
    if (k < Ng)
        slot[k] = inv[N + k];
    }

// One warp. op 0: max of one uint32 at ptr; op 1: sum of one double at ptr (rank order).
// Replaces, for the two one-word reductions of every step (the rebuild vote, hoomd/md/NeighborList.cc:1271-1285,
// and the thermostat's kinetic energy), an ncclAllReduce launch each: the exchange is n stores over NVLink and a
// spin on local HBM.
__global__ void __launch_bounds__(32)
    p2p_allreduce_kernel(P2PBox* const* __restrict__ peers, P2PBox* __restrict__ mine, int rank, int n,
                         unsigned long long epoch, int op, void* ptr)
    {
    const int lane = threadIdx.x;
    const int par = (int)(epoch & 1ull);
    const double v = op == 0 ? (double)*reinterpret_cast<const uint32_t*>(ptr) : *reinterpret_cast<const double*>(ptr);
    double got = 0.0;
    if (lane < n)
        {
        P2PBox* dst = peers[lane] + par * P2P_MAX_RANKS + rank;
        *reinterpret_cast<volatile double*>(&dst->value) = v;
        __threadfence_system(); // the value is visible system-wide before the epoch that announces it
        *reinterpret_cast<volatile unsigned long long*>(&dst->epoch) = epoch;
        const volatile P2PBox* src = mine + par * P2P_MAX_RANKS + lane;
        const long long t0 = clock64();
        bool ok = true;
        while (src->epoch != epoch)
            if (clock64() - t0 > 8000000000ll) // ~4 s: a peer died; poison the result instead of hanging
                {
                ok = false;
                break;
                }
        __threadfence_system();
        got = ok ? src->value : __longlong_as_double(0x7ff8000000000000ll);
        }
    double acc = 0.0;
    for (int r = 0; r < n; r++)
        {
        const double x = __shfl_sync(0xffffffffu, got, r);
        acc = op == 0 ? fmax(acc, x) : acc + x;
        }
    if (lane == 0)
        {
        if (op == 0)
            *reinterpret_cast<uint32_t*>(ptr) = (uint32_t)acc;
        else
            *reinterpret_cast<double*>(ptr) = acc;
        }
    }

static inline uint32_t nblk(uint32_t n)
    {
    return (n + CB - 1) / CB;
    }

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
hb2_status domain_setup(hb2_system* s, const uint32_t grid[3], int rank, int nranks,
                        const void* unique_id)
    {
    Domain& D = s->dom;
    if ((uint64_t)grid[0] * grid[1] * grid[2] != (uint64_t)nranks || rank < 0 || rank >= nranks)
        {
        set_error("hb2_system_set_domain: grid does not match the number of ranks");
        return HB2_ERR_INVALID;
        }
    HB2_TRY(nccl_load());
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm;
    HB2_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    D.comm = comm;
    D.active = true;
    D.rank = rank;
    D.nranks = nranks;
    // x fastest: rank = (cz * ny + cy) * nx + cx
    uint32_t c[3] = {(uint32_t)rank % grid[0], ((uint32_t)rank / grid[0]) % grid[1],
                     (uint32_t)rank / (grid[0] * grid[1])};
    for (int d = 0; d < 3; d++)
        {
        D.grid[d] = grid[d];
        D.coord[d] = c[d];
        }
    for (int d = 0; d < 3; d++)
        for (int dir = 0; dir < 2; dir++)
            {
            uint32_t q[3] = {c[0], c[1], c[2]};
            q[d] = (c[d] + (dir ? 1u : grid[d] - 1u)) % grid[d];
            D.nbr[d][dir] = (int)((q[2] * grid[1] + q[1]) * grid[0] + q[0]);
            }
    HB2_TRY(dev_alloc(&D.d_count, 16));
    HB2_CUDA(cudaMallocHost((void**)&D.h_count, 16 * sizeof(uint32_t)));
    return HB2_OK;
    }

void domain_destroy(hb2_system* s)
    {
    Domain& D = s->dom;
    if (!D.active)
        return;
    if (D.comm && g_nccl.CommDestroy)
        g_nccl.CommDestroy((ncclComm_t)D.comm);
    for (int r = 0; r < D.nranks && r < P2P_MAX_RANKS; r++)
        if (r != D.rank && D.peer_mail[r])
            cudaIpcCloseMemHandle(D.peer_mail[r]);
    dev_free(D.d_mail);
    dev_free(D.d_peer_mail);
    for (int d = 0; d < 3; d++)
        dev_free(D.d_send_idx[d]);
    dev_free(D.d_ghost_slot);
    dev_free(D.d_sendbuf);
    dev_free(D.d_recvbuf);
    dev_free(D.d_cls);
    dev_free(D.d_sel);
    dev_free(D.d_count);
    dev_free(D.d_inv);
    if (D.d_cub)
        cudaFree(D.d_cub);
    if (D.h_count)
        cudaFreeHost(D.h_count);
    D = Domain();
    }

hb2_status domain_ensure_scratch(hb2_system* s)
    {
    Domain& D = s->dom;
    const uint32_t cap = s->cap;
    if (D.scratch_cap >= cap)
        return HB2_OK;
    dev_free(D.d_cls);
    dev_free(D.d_sel);
    dev_free(D.d_inv);
    dev_free(D.d_ghost_slot);
    dev_free(D.d_sendbuf);
    dev_free(D.d_recvbuf);
    if (D.d_cub)
        cudaFree(D.d_cub);
    D.d_cub = nullptr;
    HB2_TRY(dev_alloc(&D.d_cls, cap));
    HB2_TRY(dev_alloc(&D.d_sel, 3 * (size_t)cap));
    HB2_TRY(dev_alloc(&D.d_inv, cap));
    HB2_TRY(dev_alloc(&D.d_ghost_slot, cap));
    D.buf_cap = cap;
    HB2_TRY(dev_alloc(&D.d_sendbuf, 2 * D.buf_cap));
    HB2_TRY(dev_alloc(&D.d_recvbuf, 2 * D.buf_cap));
    for (int d = 0; d < 3; d++)
        {
        dev_free(D.d_send_idx[d]);
        HB2_TRY(dev_alloc(&D.d_send_idx[d], cap));
        D.send_cap[d] = cap;
        }
    cub::CountingInputIterator<uint32_t> it(0);
    ClsEq f {D.d_cls, 0, false};
    cub::TransformInputIterator<uint8_t, ClsEq, cub::CountingInputIterator<uint32_t>> fl(it, f);
    size_t bytes = 0;
    cub::DeviceSelect::Flagged(nullptr, bytes, it, fl, D.d_sel, D.d_count, (int)cap);
    D.cub_bytes = bytes + 256;
    HB2_CUDA(cudaMalloc(&D.d_cub, D.cub_bytes));
    D.scratch_cap = cap;
    return HB2_OK;
    }

// indices i in [0, n) with (mask ? cls[i] & want : cls[i] == want), ascending, into out; count into
// d_count[slot]
static hb2_status select_class(hb2_system* s, uint32_t n, uint8_t want, bool mask, uint32_t* out,
                               int slot)
    {
    Domain& D = s->dom;
    cub::CountingInputIterator<uint32_t> it(0);
    ClsEq f {D.d_cls, want, mask};
    cub::TransformInputIterator<uint8_t, ClsEq, cub::CountingInputIterator<uint32_t>> fl(it, f);
    size_t bytes = D.cub_bytes;
    HB2_CUDA(cub::DeviceSelect::Flagged(D.d_cub, bytes, it, fl, out, D.d_count + slot, (int)n,
                                        s->stream));
    count_launch(2);
    return HB2_OK;
    }

// exchange the two send counts of a stage with the neighbours and bring all four numbers to the host
static hb2_status exchange_counts(hb2_system* s, int d, int slot_up, int slot_down, uint32_t& n_up,
                                  uint32_t& n_down, uint32_t& from_below, uint32_t& from_above)
    {
    Domain& D = s->dom;
    ncclComm_t comm = (ncclComm_t)D.comm;
    cudaStream_t st = s->stream;
    HB2_NCCL(g_nccl.GroupStart());
    HB2_NCCL(g_nccl.Send(D.d_count + slot_up, 1, ncclUint32, D.nbr[d][1], comm, st));
    HB2_NCCL(g_nccl.Send(D.d_count + slot_down, 1, ncclUint32, D.nbr[d][0], comm, st));
    HB2_NCCL(g_nccl.Recv(D.d_count + 8, 1, ncclUint32, D.nbr[d][0], comm, st));
    HB2_NCCL(g_nccl.Recv(D.d_count + 9, 1, ncclUint32, D.nbr[d][1], comm, st));
    HB2_NCCL(g_nccl.GroupEnd());
    HB2_CUDA(cudaMemcpyAsync(D.h_count, D.d_count, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    HB2_CUDA(cudaStreamSynchronize(st));
    n_up = D.h_count[slot_up];
    n_down = D.h_count[slot_down];
    from_below = D.h_count[8];
    from_above = D.h_count[9];
    return HB2_OK;
    }

// one grouped exchange: [up part | down part] of `send` go above / below, the part from below then the
// part from above land back to back in `recv`. Order of the calls matters when both neighbours are
// the same rank (n = 2): NCCL matches sends and receives to one peer in posting order.
static hb2_status exchange(hb2_system* s, int d, const void* send, void* recv, size_t elem_bytes,
                           uint32_t n_up, uint32_t n_down, uint32_t from_below, uint32_t from_above)
    {
    Domain& D = s->dom;
    ncclComm_t comm = (ncclComm_t)D.comm;
    cudaStream_t st = s->stream;
    const char* sp = (const char*)send;
    char* rp = (char*)recv;
    HB2_NCCL(g_nccl.GroupStart());
    if (n_up)
        HB2_NCCL(g_nccl.Send(sp, n_up * elem_bytes, ncclUint8, D.nbr[d][1], comm, st));
    if (n_down)
        HB2_NCCL(g_nccl.Send(sp + n_up * elem_bytes, n_down * elem_bytes, ncclUint8, D.nbr[d][0], comm, st));
    if (from_below)
        HB2_NCCL(g_nccl.Recv(rp, from_below * elem_bytes, ncclUint8, D.nbr[d][0], comm, st));
    if (from_above)
        HB2_NCCL(g_nccl.Recv(rp + from_below * elem_bytes, from_above * elem_bytes, ncclUint8,
                             D.nbr[d][1], comm, st));
    HB2_NCCL(g_nccl.GroupEnd());
    return HB2_OK;
    }

static hb2_status domain_migrate_copying(hb2_system* s)
    {
    Domain& D = s->dom;
    HB2_TRY(domain_ensure_scratch(s));
    cudaStream_t st = s->stream;
    const BoxD bd = boxd(s->box);
    const uint32_t cap = s->cap;
    uint32_t* sel_stay = D.d_sel;
    uint32_t* sel_up = D.d_sel + cap;
    uint32_t* sel_down = D.d_sel + 2 * (size_t)cap;
    // packed migrants: pos | vel | image | tag, up part then down part
    float4* sb_pos = D.d_sendbuf;
    float4* sb_vel = D.d_sendbuf + D.buf_cap;
    int* sb_img = (int*)D.d_recvbuf; // the receive staging is unused while migrating (direct receive)
    uint32_t* sb_tag = (uint32_t*)(D.d_recvbuf + D.buf_cap);
    for (int d = 0; d < 3; d++)
        {
        if (D.grid[d] == 1)
            continue;
        const uint32_t N = s->N;
        const int c = s->cur, o = 1 - c;
        if (N)
            {
            classify_migrate_kernel<<<nblk(N), CB, 0, st>>>(s->pos[c], N, bd, d, D.grid[d], D.coord[d],
                                                            D.d_cls);
            HB2_CHECK_LAUNCH("classify_migrate_kernel");
            }
        HB2_TRY(select_class(s, N, 0, false, sel_stay, 0));
        HB2_TRY(select_class(s, N, 1, false, sel_up, 1));
        HB2_TRY(select_class(s, N, 2, false, sel_down, 2));
        uint32_t n_up, n_down, from_below, from_above;
        HB2_TRY(exchange_counts(s, d, 1, 2, n_up, n_down, from_below, from_above));
        const uint32_t n_stay = D.h_count[0];
        const uint64_t newN = (uint64_t)n_stay + from_below + from_above;
        if (newN + s->Ng + 1 > cap || (size_t)(n_up + n_down) * 3 > D.buf_cap * 4)
            {
            set_error("domain decomposition: local particle capacity exceeded during migration");
            return HB2_ERR_INVALID;
            }
        if (n_stay)
            {
            gather_particles_kernel<<<nblk(n_stay), CB, 0, st>>>(
                sel_stay, n_stay, s->pos[c], s->vel[c], s->image[c], s->tag[c], s->pos[o], s->vel[o],
                s->image[o], s->tag[o]);
            HB2_CHECK_LAUNCH("gather_particles_kernel");
            }
        if (n_up)
            {
            gather_particles_kernel<<<nblk(n_up), CB, 0, st>>>(sel_up, n_up, s->pos[c], s->vel[c],
                                                               s->image[c], s->tag[c], sb_pos, sb_vel,
                                                               sb_img, sb_tag);
            HB2_CHECK_LAUNCH("gather_particles_kernel");
            }
        if (n_down)
            {
            gather_particles_kernel<<<nblk(n_down), CB, 0, st>>>(
                sel_down, n_down, s->pos[c], s->vel[c], s->image[c], s->tag[c], sb_pos + n_up,
                sb_vel + n_up, sb_img + 3 * (size_t)n_up, sb_tag + n_up);
            HB2_CHECK_LAUNCH("gather_particles_kernel");
            }
        HB2_TRY(exchange(s, d, sb_pos, s->pos[o] + n_stay, sizeof(float4), n_up, n_down, from_below,
                         from_above));
        HB2_TRY(exchange(s, d, sb_vel, s->vel[o] + n_stay, sizeof(float4), n_up, n_down, from_below,
                         from_above));
        HB2_TRY(exchange(s, d, sb_img, s->image[o] + 3 * (size_t)n_stay, 3 * sizeof(int), n_up, n_down,
                         from_below, from_above));
        HB2_TRY(exchange(s, d, sb_tag, s->tag[o] + n_stay, sizeof(uint32_t), n_up, n_down, from_below,
                         from_above));
        s->N = (uint32_t)newN;
        s->cur = o;
        s->order_epoch++;
        }
    return HB2_OK;
    }

// HB2_COMM_LEGACY=1 keeps the round-1 exchange (copying compaction, one message per field) for A/B runs
static bool comm_legacy()
    {
    static const bool v = getenv("HB2_COMM_LEGACY") && atoi(getenv("HB2_COMM_LEGACY")) != 0;
    return v;
    }

// Migration, in place. Per decomposed dimension: classify, three stream compactions (to above, to below,
// either — the leavers are a per-mille of the domain between two rebuilds), one host read of the four
// counts, the leavers' records copied out, ONE grouped NCCL exchange for all four fields (received
// straight behind the stayers), and the holes of the leavers filled from the tail. Nothing of the size of
// the domain is copied (the round-1 path gathered every staying particle into the other buffer, three
// times per rebuild on a 2x2x2 grid).
hb2_status domain_migrate(hb2_system* s)
    {
    if (comm_legacy())
        return domain_migrate_copying(s);
    Domain& D = s->dom;
    HB2_TRY(domain_ensure_scratch(s));
    cudaStream_t st = s->stream;
    const BoxD bd = boxd(s->box);
    const uint32_t cap = s->cap;
    uint32_t* sel_leave = D.d_sel;
    uint32_t* sel_up = D.d_sel + cap;
    uint32_t* sel_down = D.d_sel + 2 * (size_t)cap;
    float4* sb_pos = D.d_sendbuf;
    float4* sb_vel = D.d_sendbuf + D.buf_cap;
    int* sb_img = (int*)D.d_recvbuf; // the receive staging is unused while migrating (direct receive)
    uint32_t* sb_tag = (uint32_t*)(D.d_recvbuf + D.buf_cap);
    for (int d = 0; d < 3; d++)
        {
        if (D.grid[d] == 1)
            continue;
        const uint32_t N = s->N;
        const int c = s->cur;
        if (N)
            {
            classify_migrate_kernel<<<nblk(N), CB, 0, st>>>(s->pos[c], N, bd, d, D.grid[d], D.coord[d],
                                                            D.d_cls);
            HB2_CHECK_LAUNCH("classify_migrate_kernel");
            }
        HB2_TRY(select_class(s, N, 3, true, sel_leave, 0));
        HB2_TRY(select_class(s, N, 1, false, sel_up, 1));
        HB2_TRY(select_class(s, N, 2, false, sel_down, 2));
        uint32_t n_up, n_down, from_below, from_above;
        HB2_TRY(exchange_counts(s, d, 1, 2, n_up, n_down, from_below, from_above));
        const uint32_t n_leave = n_up + n_down;
        if (D.h_count[0] != n_leave || n_leave > N)
            {
            set_error("domain decomposition: inconsistent migration counts");
            return HB2_ERR_INVALID;
            }
        const uint32_t n_stay = N - n_leave;
        const uint64_t newN = (uint64_t)n_stay + from_below + from_above;
        if (newN + s->Ng + 1 > cap || (size_t)n_leave * 3 > D.buf_cap * 4)
            {
            set_error("domain decomposition: local particle capacity exceeded during migration");
            return HB2_ERR_INVALID;
            }
        if (n_up)
            {
            gather_particles_kernel<<<nblk(n_up), CB, 0, st>>>(sel_up, n_up, s->pos[c], s->vel[c],
                                                               s->image[c], s->tag[c], sb_pos, sb_vel,
                                                               sb_img, sb_tag);
            HB2_CHECK_LAUNCH("gather_particles_kernel");
            }
        if (n_down)
            {
            gather_particles_kernel<<<nblk(n_down), CB, 0, st>>>(
                sel_down, n_down, s->pos[c], s->vel[c], s->image[c], s->tag[c], sb_pos + n_up,
                sb_vel + n_up, sb_img + 3 * (size_t)n_up, sb_tag + n_up);
            HB2_CHECK_LAUNCH("gather_particles_kernel");
            }
        if (n_leave)
            {
            fill_holes_kernel<<<nblk(n_leave), CB, 0, st>>>(sel_leave, n_leave, N, D.d_cls, s->pos[c],
                                                            s->vel[c], s->image[c], s->tag[c]);
            HB2_CHECK_LAUNCH("fill_holes_kernel");
            }
        if (n_leave || from_below || from_above)
            {
            HB2_NCCL(g_nccl.GroupStart());
            HB2_TRY(exchange(s, d, sb_pos, s->pos[c] + n_stay, sizeof(float4), n_up, n_down, from_below,
                             from_above));
            HB2_TRY(exchange(s, d, sb_vel, s->vel[c] + n_stay, sizeof(float4), n_up, n_down, from_below,
                             from_above));
            HB2_TRY(exchange(s, d, sb_img, s->image[c] + 3 * (size_t)n_stay, 3 * sizeof(int), n_up, n_down,
                             from_below, from_above));
            HB2_TRY(exchange(s, d, sb_tag, s->tag[c] + n_stay, sizeof(uint32_t), n_up, n_down, from_below,
                             from_above));
            HB2_NCCL(g_nccl.GroupEnd());
            s->order_epoch++;
            }
        s->N = (uint32_t)newN;
        }
    return HB2_OK;
    }

hb2_status domain_exchange_ghosts(hb2_system* s, double r_ghost)
    {
    Domain& D = s->dom;
    HB2_TRY(domain_ensure_scratch(s));
    cudaStream_t st = s->stream;
    const BoxD bd = boxd(s->box);
    const uint32_t cap = s->cap;
    const int c = s->cur;
    s->Ng = 0;
    float4* sb_pos = D.d_sendbuf;
    float4* sb_vel = D.d_sendbuf + D.buf_cap;
    uint32_t* sb_tag = (uint32_t*)D.d_recvbuf;
    for (int d = 0; d < 3; d++)
        {
        D.n_send[d][0] = D.n_send[d][1] = D.n_recv[d][0] = D.n_recv[d][1] = 0;
        D.ghost_off[d] = s->Ng;
        if (D.grid[d] == 1)
            continue;
        const uint32_t Ntot = s->N + s->Ng;
        const double W = s->box.L[d] / (double)D.grid[d];
        // a hair wider than r_ghost: the list test itself is exact, the layer only has to be a superset
        const double gfrac = (r_ghost * (1.0 + 1e-6) + 1e-6 * s->box.L[d]) / W;
        if (gfrac >= 0.5)
            {
            set_error("domain decomposition: sub-domain narrower than twice the ghost layer "
                      "(hoomd/Communicator.cc:1370-1395 raises the same condition)");
            return HB2_ERR_BOX_TOO_SMALL;
            }
        if (Ntot)
            {
            classify_ghost_kernel<<<nblk(Ntot), CB, 0, st>>>(s->pos[c], Ntot, bd, d, D.grid[d],
                                                             D.coord[d], gfrac, D.d_cls);
            HB2_CHECK_LAUNCH("classify_ghost_kernel");
            }
        uint32_t* list = D.d_send_idx[d];
        HB2_TRY(select_class(s, Ntot, 1, true, D.d_sel, 1)); // to above
        HB2_TRY(select_class(s, Ntot, 2, true, D.d_sel + cap, 2)); // to below
        uint32_t n_up, n_down, from_below, from_above;
        HB2_TRY(exchange_counts(s, d, 1, 2, n_up, n_down, from_below, from_above));
        if ((uint64_t)Ntot + from_below + from_above + 1 > cap || n_up + n_down > D.send_cap[d]
            || n_up + n_down > D.buf_cap)
            {
            set_error("domain decomposition: local particle capacity exceeded by the ghost layer");
            return HB2_ERR_INVALID;
            }
        HB2_CUDA(cudaMemcpyAsync(list, D.d_sel, sizeof(uint32_t) * n_up, cudaMemcpyDeviceToDevice, st));
        HB2_CUDA(cudaMemcpyAsync(list + n_up, D.d_sel + cap, sizeof(uint32_t) * n_down,
                                 cudaMemcpyDeviceToDevice, st));
        const uint32_t ns = n_up + n_down;
        if (ns)
            {
            gather_particles_kernel<<<nblk(ns), CB, 0, st>>>(list, ns, s->pos[c], s->vel[c], nullptr,
                                                             s->tag[c], sb_pos, sb_vel, nullptr,
                                                             sb_tag);
            HB2_CHECK_LAUNCH("gather_particles_kernel");
            }
        // one NCCL launch for the three fields of the stage (groups nest; posting order is kept)
        if (!comm_legacy())
            HB2_NCCL(g_nccl.GroupStart());
        HB2_TRY(exchange(s, d, sb_pos, s->pos[c] + Ntot, sizeof(float4), n_up, n_down, from_below,
                         from_above));
        HB2_TRY(exchange(s, d, sb_vel, s->vel[c] + Ntot, sizeof(float4), n_up, n_down, from_below,
                         from_above));
        HB2_TRY(exchange(s, d, sb_tag, s->tag[c] + Ntot, sizeof(uint32_t), n_up, n_down, from_below,
                         from_above));
        if (!comm_legacy())
            HB2_NCCL(g_nccl.GroupEnd());
        D.n_send[d][1] = n_up;
        D.n_send[d][0] = n_down;
        D.n_recv[d][0] = from_below;
        D.n_recv[d][1] = from_above;
        s->Ng += from_below + from_above;
        }
    // ghosts carry no image; keep the array defined for the reorder pass
    if (s->Ng)
        HB2_CUDA(cudaMemsetAsync(s->image[c] + 3 * (size_t)s->N, 0, sizeof(int) * 3 * (size_t)s->Ng, st));
    return HB2_OK;
    }

hb2_status domain_remap(hb2_system* s, const uint32_t* d_order, uint32_t Ntot)
    {
    Domain& D = s->dom;
    cudaStream_t st = s->stream;
    if (!Ntot)
        return HB2_OK;
    invert_perm_kernel<<<nblk(Ntot), CB, 0, st>>>(d_order, Ntot, D.d_inv);
    HB2_CHECK_LAUNCH("invert_perm_kernel");
    for (int d = 0; d < 3; d++)
        {
        const uint32_t ns = D.n_send[d][0] + D.n_send[d][1];
        if (ns)
            {
            remap_kernel<<<nblk(ns), CB, 0, st>>>(D.d_send_idx[d], ns, D.d_inv);
            HB2_CHECK_LAUNCH("remap_kernel");
            }
        }
    if (s->Ng)
        {
        ghost_slot_kernel<<<nblk(s->Ng), CB, 0, st>>>(D.d_ghost_slot, s->Ng, s->N, D.d_inv);
        HB2_CHECK_LAUNCH("ghost_slot_kernel");
        }
    return HB2_OK;
    }

hb2_status domain_update_ghosts(hb2_system* s)
    {
    Domain& D = s->dom;
    cudaStream_t st = s->stream;
    const int c = s->cur;
    float4* sb_pos = D.d_sendbuf;
    float4* sb_vel = D.d_sendbuf + D.buf_cap;
    float4* rb_pos = D.d_recvbuf;
    float4* rb_vel = D.d_recvbuf + D.buf_cap;
    for (int d = 0; d < 3; d++)
        {
        if (D.grid[d] == 1)
            continue;
        const uint32_t n_up = D.n_send[d][1], n_down = D.n_send[d][0];
        const uint32_t fb = D.n_recv[d][0], fa = D.n_recv[d][1];
        const uint32_t ns = n_up + n_down, nr = fb + fa;
        if (ns)
            {
            pack_ghost_kernel<<<nblk(ns), CB, 0, st>>>(D.d_send_idx[d], ns, s->pos[c], s->vel[c],
                                                       sb_pos, D.need_vel ? sb_vel : nullptr);
            HB2_CHECK_LAUNCH("pack_ghost_kernel");
            }
        if (D.need_vel)
            HB2_NCCL(g_nccl.GroupStart());
        HB2_TRY(exchange(s, d, sb_pos, rb_pos, sizeof(float4), n_up, n_down, fb, fa));
        if (D.need_vel)
            {
            HB2_TRY(exchange(s, d, sb_vel, rb_vel, sizeof(float4), n_up, n_down, fb, fa));
            HB2_NCCL(g_nccl.GroupEnd());
            }
        if (nr)
            {
            unpack_ghost_kernel<<<nblk(nr), CB, 0, st>>>(D.d_ghost_slot + D.ghost_off[d], nr, rb_pos,
                                                         D.need_vel ? rb_vel : nullptr, s->pos[c],
                                                         s->vel[c]);
            HB2_CHECK_LAUNCH("unpack_ghost_kernel");
            }
        }
    return HB2_OK;
    }

static hb2_status p2p_allreduce(hb2_system* s, void* d_val, int op)
    {
    Domain& D = s->dom;
    D.p2p_epoch++;
    p2p_allreduce_kernel<<<1, 32, 0, s->stream>>>(D.d_peer_mail, D.d_mail, D.rank, D.nranks, D.p2p_epoch, op, d_val);
    HB2_CHECK_LAUNCH("p2p_allreduce_kernel");
    return HB2_OK;
    }

hb2_status domain_allreduce_max_u32(hb2_system* s, uint32_t* d_val, uint32_t n)
    {
    if (s->dom.p2p && n == 1)
        return p2p_allreduce(s, d_val, 0);
    HB2_NCCL(g_nccl.AllReduce(d_val, d_val, n, ncclUint32, ncclMax, (ncclComm_t)s->dom.comm, s->stream));
    return HB2_OK;
    }

hb2_status domain_allreduce_sum_f64(hb2_system* s, double* d_val, uint32_t n)
    {
    if (s->dom.p2p && n == 1)
        return p2p_allreduce(s, d_val, 1);
    HB2_NCCL(g_nccl.AllReduce(d_val, d_val, n, ncclDouble, ncclSum, (ncclComm_t)s->dom.comm, s->stream));
    return HB2_OK;
    }
    } // namespace hb2

// =================================================================================================
// C ABI
// =================================================================================================
// ---- peer-memory mailboxes: bootstrap. Every rank allocates its mailbox and exports it
// (hb2_domain_ipc_handle), the handles are gathered by the caller's side channel (torch.distributed, like the
// NCCL unique id), and hb2_domain_set_peers maps them. The switch is collective: if the mapping fails on any rank
// (or HB2_P2P=0), every rank keeps the NCCL all-reduces.
extern "C" hb2_status hb2_domain_ipc_handle(hb2_system* s, void* out64)
    {
    if (!s || !out64 || !s->dom.active)
        return HB2_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    Domain& D = s->dom;
    HB2_CUDA(cudaSetDevice(s->device));
    memset(out64, 0, 64);
    if (D.nranks > P2P_MAX_RANKS)
        return HB2_OK; // too many ranks for the mailbox: the zero handle switches the path off everywhere
    if (!D.d_mail)
        HB2_TRY(dev_alloc(&D.d_mail, 2 * (size_t)P2P_MAX_RANKS));
    HB2_CUDA(cudaMemset(D.d_mail, 0, sizeof(P2PBox) * 2 * P2P_MAX_RANKS));
    HB2_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, D.d_mail) != cudaSuccess)
        {
        cudaGetLastError();
        return HB2_OK; // zero handle
        }
    memcpy(out64, &h, 64);
    return HB2_OK;
    }

extern "C" hb2_status hb2_domain_set_peers(hb2_system* s, const void* handles, uint32_t nranks)
    {
    if (!s || !handles || !s->dom.active || (int)nranks != s->dom.nranks)
        return HB2_ERR_INVALID;
    Domain& D = s->dom;
    HB2_CUDA(cudaSetDevice(s->device));
    static const bool enabled = !(getenv("HB2_P2P") && atoi(getenv("HB2_P2P")) == 0);
    uint32_t failed = (enabled && D.d_mail && (int)nranks <= P2P_MAX_RANKS) ? 0u : 1u;
    const unsigned char zero[64] = {0};
    for (uint32_t r = 0; r < nranks && !failed; r++)
        {
        const unsigned char* hr = (const unsigned char*)handles + 64 * (size_t)r;
        if (memcmp(hr, zero, 64) == 0)
            {
            failed = 1u;
            break;
            }
        if ((int)r == D.rank)
            {
            D.peer_mail[r] = D.d_mail;
            continue;
            }
        cudaIpcMemHandle_t h;
        memcpy(&h, hr, 64);
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
            {
            cudaGetLastError();
            failed = 1u;
            break;
            }
        D.peer_mail[r] = (P2PBox*)p;
        }
    // agree: one failure anywhere keeps NCCL everywhere
    uint32_t* d_f = nullptr;
    HB2_TRY(dev_alloc(&d_f, 1));
    HB2_CUDA(cudaMemcpyAsync(d_f, &failed, sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
    HB2_NCCL(g_nccl.AllReduce(d_f, d_f, 1, ncclUint32, ncclMax, (ncclComm_t)D.comm, s->stream));
    HB2_CUDA(cudaMemcpyAsync(&failed, d_f, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    dev_free(d_f);
    if (failed)
        {
        D.p2p = false;
        return HB2_OK;
        }
    if (!D.d_peer_mail)
        HB2_TRY(dev_alloc(&D.d_peer_mail, (size_t)P2P_MAX_RANKS));
    HB2_CUDA(cudaMemcpy(D.d_peer_mail, D.peer_mail, sizeof(P2PBox*) * P2P_MAX_RANKS, cudaMemcpyHostToDevice));
    D.p2p_epoch = 0;
    D.p2p = true;
    return HB2_OK;
    }

extern "C" int hb2_domain_p2p_active(const hb2_system* s)
    {
    return s && s->dom.p2p ? 1 : 0;
    }

extern "C" hb2_status hb2_comm_unique_id(void* out128)
    {
    if (!out128)
        return HB2_ERR_INVALID;
    HB2_TRY(nccl_load());
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    HB2_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return HB2_OK;
    }

extern "C" hb2_status hb2_domain_find_decomposition(uint32_t nranks, double Lx, double Ly, double Lz,
                                                    uint32_t grid[3])
    {
    // hoomd/DomainDecomposition.cc:327-409: the factorisation nx*ny*nz = nranks with the smallest
    // total interface area; ties keep the first candidate in (nx, ny, nz) lexicographic order, and the
    // starting guess is (1, 1, nranks).
    if (!grid || nranks == 0 || !(Lx > 0) || !(Ly > 0) || !(Lz > 0))
        {
        set_error("hb2_domain_find_decomposition: bad arguments");
        return HB2_ERR_INVALID;
        }
    double best = Lx * Lz * (double)(nranks - 1);
    grid[0] = 1;
    grid[1] = 1;
    grid[2] = nranks;
    bool found = true; // (no user constraints on this path)
    for (uint32_t nx = 1; nx <= nranks; nx++)
        for (uint32_t ny = 1; nx * ny <= nranks; ny++)
            for (uint32_t nz = 1; nx * ny * nz <= nranks; nz++)
                {
                if (nx * ny * nz != nranks)
                    continue;
                const double area = Lx * Ly * (double)(nz - 1) + Lx * Lz * (double)(ny - 1)
                                    + Ly * Lz * (double)(nx - 1);
                if (area < best || !found)
                    {
                    grid[0] = nx;
                    grid[1] = ny;
                    grid[2] = nz;
                    best = area;
                    found = true;
                    }
                }
    return HB2_OK;
    }

extern "C" hb2_status hb2_domain_place(const hb2_box* box, const uint32_t grid[3], const float* pos3,
                                       uint32_t n, uint32_t* rank_out)
    {
    // hoomd/DomainDecomposition.cc:604-660 (placeParticle) with uniform cumulative fractions
    if (!box || !grid || !pos3 || !rank_out)
        return HB2_ERR_INVALID;
    for (uint32_t i = 0; i < n; i++)
        {
        uint32_t c[3];
        for (int d = 0; d < 3; d++)
            {
            const double f = ((double)pos3[3 * (size_t)i + d] - box->lo[d]) * box->Linv[d];
            if (f < -1e-5 || f >= 1.0 + 1e-5)
                {
                set_error("Particle coordinates outside global box.");
                return HB2_ERR_OUT_OF_BOX;
                }
            // the coordinate whose domain_g lands in [0, 1)
            uint32_t k = 0;
            for (; k < grid[d]; k++)
                {
                const double g = domain_g(f, grid[d], k);
                if (g >= 0.0 && g < 1.0)
                    break;
                }
            c[d] = k < grid[d] ? k : 0;
            }
        rank_out[i] = (c[2] * grid[1] + c[1]) * grid[0] + c[0];
        }
    return HB2_OK;
    }

extern "C" hb2_status hb2_domain_ghost_mask(const hb2_box* box, const uint32_t grid[3],
                                            const uint32_t coord[3], double r_ghost,
                                            const float* pos3, uint32_t n, uint8_t* mask_out)
    {
    // the ghost plan of classify_ghost_kernel evaluated on the host, all three dimensions at once:
    // bit 2d: within the layer of the lower face of dimension d, bit 2d+1: of the upper face (valid
    // for particles inside [0,1) of that dimension; 0 otherwise)
    if (!box || !grid || !coord || !pos3 || !mask_out)
        return HB2_ERR_INVALID;
    for (uint32_t i = 0; i < n; i++)
        {
        uint8_t m = 0;
        for (int d = 0; d < 3; d++)
            {
            if (grid[d] == 1)
                continue;
            const double W = box->L[d] / (double)grid[d];
            const double gfrac = (r_ghost * (1.0 + 1e-6) + 1e-6 * box->L[d]) / W;
            const double f = ((double)pos3[3 * (size_t)i + d] - box->lo[d]) * box->Linv[d];
            const double g = domain_g(f, grid[d], coord[d]);
            if (g >= 0.0 && g < 1.0)
                {
                if (g < gfrac)
                    m |= (uint8_t)(1u << (2 * d));
                if (g >= 1.0 - gfrac)
                    m |= (uint8_t)(1u << (2 * d + 1));
                }
            }
        mask_out[i] = m;
        }
    return HB2_OK;
    }
