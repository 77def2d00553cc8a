// engine.cu — host side of the path in C++: particle state resident in HBM, the neighbour list
// object, the pair force object and the two-step integrator, exported through the C ABI of
// include/hoomd_b200.h (layer 2).  Mirrors, for this path only:
//   ParticleData            hoomd/ParticleData.h:1285-1334
//   NeighborList(+Binned)   hoomd/md/NeighborList.cc:363-435 (compute), :452-529 (r_list),
//                           :535-561 (box check), :1325-1418 (needsUpdating), :1601-1617 (overflow)
//   PotentialPair           hoomd/md/PotentialPair.h:369-371, :467-580, :587-794
//   IntegratorTwoStep       hoomd/md/IntegratorTwoStep.cc:51-123 (update), :226-281 (prepRun)
//   TwoStepConstantVolume   hoomd/md/TwoStepConstantVolume.cc
#include "common.cuh"
#include "evaluators.cuh"
#include "kernels.cuh"
#include "pair_kernel.cuh"
#include "evaluators_more.cuh"
#include "engine.cuh"
#include <cstdio>
#include <cstdlib>
#include <ctime>

namespace hb2
    {
static thread_local std::string t_error;
uint64_t g_launches = 0;

void set_error(const std::string& msg)
    {
    t_error = msg;
    }
hb2_status cuda_fail(cudaError_t e, const char* what)
    {
    t_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return HB2_ERR_CUDA;
    }

    } // namespace hb2


// =================================================================================================
// system
// =================================================================================================
extern "C" const char* hb2_last_error(void)
    {
    return t_error.c_str();
    }
extern "C" int hb2_abi_version(void)
    {
    return HB2_ABI_VERSION;
    }
extern "C" uint64_t hb2_launch_count(void)
    {
    return g_launches;
    }

extern "C" hb2_status hb2_box_make(hb2_box* box, double Lx, double Ly, double Lz)
    {
    if (!box || !(Lx > 0) || !(Ly > 0) || !(Lz > 0))
        {
        set_error("hb2_box_make: box lengths must be positive");
        return HB2_ERR_INVALID;
        }
    const double L[3] = {Lx, Ly, Lz};
    for (int d = 0; d < 3; d++)
        {
        // hoomd/BoxDim.h:183-200
        box->L[d] = L[d];
        box->hi[d] = L[d] / 2.0;
        box->lo[d] = -box->hi[d];
        box->Linv[d] = 1.0 / L[d];
        box->Lf[d] = (float)box->L[d];
        box->hif[d] = (float)box->hi[d];
        box->lof[d] = (float)box->lo[d];
        box->Linvf[d] = (float)box->Linv[d];
        box->periodic[d] = 1;
        }
    box->_pad = 0;
    return HB2_OK;
    }

extern "C" hb2_status hb2_system_create(hb2_system** out, int device, uint32_t ntypes, uint32_t seed)
    {
    if (!out || ntypes == 0)
        {
        set_error("hb2_system_create: bad arguments");
        return HB2_ERR_INVALID;
        }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        {
        cudaGetLastError();
        set_error("hb2_system_create: no CUDA device is visible and this library has no CPU path");
        return HB2_ERR_NO_DEVICE;
        }
    if (device < 0 || device >= ndev)
        {
        set_error("hb2_system_create: device ordinal out of range");
        return HB2_ERR_INVALID;
        }
    HB2_CUDA(cudaSetDevice(device));
    hb2_system* s = new hb2_system();
    s->device = device;
    s->ntypes = ntypes;
    s->seed = seed & 0xffffu;
    hb2_status st = HB2_OK;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreateWithFlags(&s->flag_ready, cudaEventDisableTiming) != cudaSuccess
        || cudaEventCreateWithFlags(&s->flag_copied, cudaEventDisableTiming) != cudaSuccess)
        st = HB2_ERR_CUDA;
    if (st == HB2_OK)
        st = dev_alloc(&s->d_timestep, 1);
    if (st == HB2_OK)
        st = dev_alloc(&s->d_ticket, 4);
    if (st == HB2_OK)
        st = dev_alloc(&s->d_sums, 16);
    if (st == HB2_OK)
        {
        cudaMemsetAsync(s->d_ticket, 0, 4 * sizeof(uint32_t), s->stream);
        cudaMemsetAsync(s->d_sums, 0, 16 * sizeof(double), s->stream);
        st = s->set_timestep(0);
        }
    if (st != HB2_OK)
        {
        delete s;
        return st;
        }
    *out = s;
    return HB2_OK;
    }

extern "C" void hb2_system_destroy(hb2_system* s)
    {
    if (!s)
        return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    domain_destroy(s);
    for (int b = 0; b < 2; b++)
        {
        dev_free(s->pos[b]);
        dev_free(s->vel[b]);
        dev_free(s->image[b]);
        dev_free(s->tag[b]);
        }
    dev_free(s->rtag);
    dev_free(s->d_bond_head);
    dev_free(s->d_bond_partner);
    dev_free(s->d_bond_type);
    dev_free(s->d_angle_head);
    dev_free(s->d_angle_members);
    dev_free(s->d_dihedral_head);
    dev_free(s->d_dihedral_members);
    dev_free(s->d_dihedral_tr);
    dev_free(s->d_stage);
    if (s->h_stage)
        cudaFreeHost(s->h_stage);
    dev_free(s->d_timestep);
    dev_free(s->d_partial);
    dev_free(s->d_ticket);
    dev_free(s->d_sums);
    cudaStreamDestroy(s->stream);
    if (s->side_stream)
        cudaStreamDestroy(s->side_stream);
    if (s->flag_ready)
        cudaEventDestroy(s->flag_ready);
    if (s->flag_copied)
        cudaEventDestroy(s->flag_copied);
    delete s;
    }

extern "C" hb2_status hb2_system_set_box(hb2_system* s, double Lx, double Ly, double Lz, double xy,
                                         double xz, double yz)
    {
    if (!s)
        return HB2_ERR_INVALID;
    if (xy != 0.0 || xz != 0.0 || yz != 0.0)
        {
        set_error("hb2_system_set_box: triclinic boxes are outside this path's scope (tilt must be 0)");
        return HB2_ERR_INVALID;
        }
    HB2_TRY(hb2_box_make(&s->box, Lx, Ly, Lz));
    s->has_box = true;
    for (auto* nl : s->nlists)
        {
        nl->force_update = true;
        nl->params_dirty = true;
        }
    return HB2_OK;
    }

// (Re)allocates every per-particle array of the system for `cap` slots.
static hb2_status system_allocate(hb2_system* s, uint32_t cap, bool with_rtag, uint64_t n_tags)
    {
    for (int b = 0; b < 2; b++)
        {
        dev_free(s->pos[b]);
        dev_free(s->vel[b]);
        dev_free(s->image[b]);
        dev_free(s->tag[b]);
        HB2_TRY(dev_alloc(&s->pos[b], (size_t)cap));
        HB2_TRY(dev_alloc(&s->vel[b], (size_t)cap));
        HB2_TRY(dev_alloc(&s->image[b], 3 * (size_t)cap));
        HB2_TRY(dev_alloc(&s->tag[b], (size_t)cap));
        }
    // (a domain-decomposed system carries an rtag table only when it has bonds; it is indexed by the
    // global tag, so it survives a change of the local capacity)
    if (!(s->dom.active && s->rtag))
        {
        dev_free(s->rtag);
        s->rtag = nullptr;
        if (with_rtag)
            HB2_TRY(dev_alloc(&s->rtag, n_tags));
        }
    dev_free(s->d_partial);
    HB2_TRY(dev_alloc(&s->d_partial, 8 * (size_t)(reduce_blocks(cap) + 1)));
    s->cap = cap;
    return HB2_OK;
    }

// the far-away dummy particle that pads neighbour-list rows lives right behind the last ghost
static hb2_status system_place_dummy(hb2_system* s)
    {
    // static: the sources must outlive the asynchronous copies
    static const float4 far_pos = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.f);
    static const float4 far_vel = make_float4(0.f, 0.f, 0.f, 1.f);
    static const uint32_t far_tag = 0xffffffffu;
    const size_t at = (size_t)s->N + s->Ng;
    const int b = s->cur;
    HB2_CUDA(cudaMemcpyAsync(s->pos[b] + at, &far_pos, sizeof(float4), cudaMemcpyHostToDevice, s->stream));
    HB2_CUDA(cudaMemcpyAsync(s->vel[b] + at, &far_vel, sizeof(float4), cudaMemcpyHostToDevice, s->stream));
    HB2_CUDA(cudaMemcpyAsync(s->tag[b] + at, &far_tag, sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
    return HB2_OK;
    }

// Common body of set_particles / set_particles_tagged. With a domain decomposition only the
// particles this rank owns are kept (hoomd/ParticleData.cc initializeFromSnapshot scatters the
// snapshot with DomainDecomposition::placeParticle in the same way).
static hb2_status system_set(hb2_system* s, uint32_t n_in, uint64_t N_global, const uint32_t* tag,
                             const float* pos3, const float* vel3, const float* mass,
                             const uint32_t* typeid_, const int32_t* image3)
    {
    if (!s || !pos3)
        {
        set_error("hb2_system_set_particles: positions are required");
        return HB2_ERR_INVALID;
        }
    if (!s->has_box)
        {
        set_error("hb2_system_set_particles: set the box first");
        return HB2_ERR_INVALID;
        }
    HB2_CUDA(cudaSetDevice(s->device));
    Domain& D = s->dom;
    std::vector<float4> hp, hv;
    std::vector<int> hi;
    std::vector<uint32_t> ht;
    hp.reserve(D.active ? n_in / D.nranks + 16 : n_in);
    for (uint32_t i = 0; i < n_in; i++)
        {
        const uint32_t t = typeid_ ? typeid_[i] : 0u;
        if (t >= s->ntypes)
            {
            set_error("hb2_system_set_particles: type id out of range");
            return HB2_ERR_INVALID;
            }
        if (D.active)
            {
            bool mine = true;
            for (int d = 0; d < 3 && mine; d++)
                {
                if (D.grid[d] == 1)
                    continue;
                const double f = ((double)pos3[3 * (size_t)i + d] - s->box.lo[d]) * s->box.Linv[d];
                const double g = domain_g(f, D.grid[d], D.coord[d]);
                mine = (g >= 0.0 && g < 1.0);
                }
            if (!mine)
                continue;
            }
        float tw;
        memcpy(&tw, &t, 4);
        hp.push_back(make_float4(pos3[3 * (size_t)i], pos3[3 * (size_t)i + 1], pos3[3 * (size_t)i + 2], tw));
        hv.push_back(make_float4(vel3 ? vel3[3 * (size_t)i] : 0.f, vel3 ? vel3[3 * (size_t)i + 1] : 0.f,
                                 vel3 ? vel3[3 * (size_t)i + 2] : 0.f, mass ? mass[i] : 1.0f));
        for (int d = 0; d < 3; d++)
            hi.push_back(image3 ? image3[3 * (size_t)i + d] : 0);
        ht.push_back(tag ? tag[i] : i);
        }
    const uint32_t N = (uint32_t)hp.size();
    uint32_t need = N + 1;
    if (D.active)
        {
        // room for density fluctuations, migration and the ghost layer (estimated from the mean
        // density with a generous margin; exceeding it is reported, not silently truncated)
        const double n0 = (double)N_global / D.nranks;
        double shell = 1.0;
        for (int d = 0; d < 3; d++)
            if (D.grid[d] > 1)
                shell *= 1.0 + 2.0 * std::min(0.5, 4.0 / (s->box.L[d] / D.grid[d]));
        need = (uint32_t)std::min<double>(4.0e9, std::max(1.25 * n0, 1.1 * (double)N) * shell + 4096.0);
        }
    if (need > s->cap || !s->pos[0])
        HB2_TRY(system_allocate(s, need, !D.active, N_global));
    s->N = N;
    s->Ng = 0;
    s->cur = 0;
    D.N_global = N_global;
    cudaStream_t st = s->stream;
    HB2_CUDA(cudaMemcpyAsync(s->pos[0], hp.data(), sizeof(float4) * N, cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaMemcpyAsync(s->vel[0], hv.data(), sizeof(float4) * N, cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaMemcpyAsync(s->image[0], hi.data(), sizeof(int) * 3 * (size_t)N, cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaMemcpyAsync(s->tag[0], ht.data(), sizeof(uint32_t) * N, cudaMemcpyHostToDevice, st));
    if (s->rtag)
        HB2_CUDA(cudaMemcpyAsync(s->rtag, ht.data(), sizeof(uint32_t) * N, cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaStreamSynchronize(st));
    HB2_TRY(system_place_dummy(s));
    s->order_epoch++;
    for (auto* nl : s->nlists)
        {
        nl->force_update = true;
        nl->has_list = false;
        }
    for (auto* p : s->pairs)
        p->computed = false;
    return HB2_OK;
    }

static hb2_status system_stage(hb2_system* s, uint32_t N)
    {
    size_t words = 12 * (size_t)N; // pos3 vel3 mass typeid image3 tag
    if (words <= s->stage_cap)
        return HB2_OK;
    // With a domain decomposition the owned count drifts with every migration: size the staging for the
    // particle capacity once (page-locked allocations cost ~100 ms per 200 MB, and were being repeated
    // whenever a rank had gained a few particles).
    if (s->dom.active)
        words = std::max(words, 12 * (size_t)s->cap);
    dev_free(s->d_stage);
    if (s->h_stage)
        cudaFreeHost(s->h_stage);
    s->h_stage = nullptr;
    HB2_TRY(dev_alloc(&s->d_stage, words));
    HB2_CUDA(cudaMallocHost((void**)&s->h_stage, words * sizeof(float)));
    s->stage_cap = words;
    return HB2_OK;
    }

static bool host_pinned(const void* p)
    {
    if (!p)
        return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
        {
        cudaGetLastError();
        return false;
        }
    return at.type == cudaMemoryTypeHost;
    }

// Device path of set_particles[_tagged]: the caller's raw arrays cross PCIe once — by direct DMA when
// they are page-locked (cudaHostAlloc / cudaHostRegister, e.g. torch pinned tensors), through the
// pinned bounce buffer otherwise — and are packed into the float4 layout on the device. With a domain
// decomposition every particle must be owned by this rank (checked on the device with domain_g);
// `*not_owned` reports when that is not so and the caller falls back to the host filter.
static hb2_status system_set_device(hb2_system* s, uint32_t N, uint64_t N_global, const uint32_t* tag,
                                    const float* pos3, const float* vel3, const float* mass,
                                    const uint32_t* typeid_, const int32_t* image3, bool* not_owned)
    {
    HB2_TRY(system_stage(s, N));
    cudaStream_t st = s->stream;
    const size_t n = N;
    float* h = s->h_stage;
    float* d = s->d_stage;
    const bool direct = host_pinned(pos3) && host_pinned(vel3) && host_pinned(mass)
                        && host_pinned(typeid_) && host_pinned(image3) && host_pinned(tag);
    auto put = [&](size_t off_words, const void* src, size_t bytes, int fill) -> hb2_status
    {
        if (src && direct)
            {
            HB2_CUDA(cudaMemcpyAsync(d + off_words, src, bytes, cudaMemcpyHostToDevice, st));
            return HB2_OK;
            }
        if (src)
            memcpy(h + off_words, src, bytes);
        else if (fill == 1)
            std::fill(h + off_words, h + off_words + bytes / 4, 1.0f);
        else
            memset(h + off_words, 0, bytes);
        if (direct) // a missing array's default still comes from the bounce buffer
            HB2_CUDA(cudaMemcpyAsync(d + off_words, h + off_words, bytes, cudaMemcpyHostToDevice, st));
        return HB2_OK;
    };
    HB2_TRY(put(0, pos3, 12 * n, 0));
    HB2_TRY(put(3 * n, vel3, 12 * n, 0));
    HB2_TRY(put(6 * n, mass, 4 * n, 1));
    HB2_TRY(put(7 * n, typeid_, 4 * n, 0));
    HB2_TRY(put(8 * n, image3, 12 * n, 0));
    if (tag)
        HB2_TRY(put(11 * n, tag, 4 * n, 0));
    if (!direct)
        HB2_CUDA(cudaMemcpyAsync(d, h, (tag ? 48 : 44) * n, cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaMemsetAsync(s->d_ticket + 2, 0, 2 * sizeof(uint32_t), st));
    DomainGeom geom;
    for (int k = 0; k < 3; k++)
        {
        geom.n[k] = s->dom.active ? s->dom.grid[k] : 1u;
        geom.coord[k] = s->dom.active ? s->dom.coord[k] : 0u;
        geom.dimi[k] = geom.dimt[k] = 1;
        }
    // Particles a rank hands back (hb2_system_get_local -> ... -> set_particles_tagged) may have drifted
    // past a face since the last migration — at most half the list buffer before a rebuild is forced;
    // one length unit of slack covers every list in use. The rebuild this call forces migrates them.
    OwnMargin margin;
    if (s->dom.active && tag)
        for (int k = 0; k < 3; k++)
            margin.frac[k] = std::min(0.25, 1.0 / (s->box.L[k] / (double)s->dom.grid[k]));
    HB2_TRY(snapshot_pack(N, d, d + 3 * n, d + 6 * n, (const uint32_t*)(d + 7 * n),
                          (const int*)(d + 8 * n), tag ? (const uint32_t*)(d + 11 * n) : nullptr,
                          s->pos[0], s->vel[0], s->image[0], s->tag[0], s->rtag, s->ntypes, s->box, geom,
                          margin, s->d_ticket + 2, st));
    uint32_t flags[2] = {0, 0};
    HB2_CUDA(cudaMemcpyAsync(flags, s->d_ticket + 2, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    HB2_CUDA(cudaMemsetAsync(s->d_ticket + 2, 0, 2 * sizeof(uint32_t), st)); // tickets rest at zero
    HB2_CUDA(cudaStreamSynchronize(st));
    if (flags[0])
        {
        s->N = 0;
        set_error("hb2_system_set_particles: type id out of range");
        return HB2_ERR_INVALID;
        }
    *not_owned = flags[1] != 0;
    if (*not_owned)
        return HB2_OK;
    s->N = N;
    s->Ng = 0;
    s->cur = 0;
    s->dom.N_global = N_global;
    HB2_TRY(system_place_dummy(s));
    s->order_epoch++;
    for (auto* nl : s->nlists)
        {
        nl->force_update = true;
        nl->has_list = false;
        }
    for (auto* p : s->pairs)
        p->computed = false;
    return HB2_OK;
    }

extern "C" hb2_status hb2_system_set_particles(hb2_system* s, uint32_t N, const float* pos3,
                                               const float* vel3, const float* mass,
                                               const uint32_t* typeid_, const int32_t* image3)
    {
    if (!s || !pos3)
        {
        set_error("hb2_system_set_particles: positions are required");
        return HB2_ERR_INVALID;
        }
    if (s->dom.active)
        return system_set(s, N, N, nullptr, pos3, vel3, mass, typeid_, image3);
    if (!s->has_box)
        {
        set_error("hb2_system_set_particles: set the box first");
        return HB2_ERR_INVALID;
        }
    HB2_CUDA(cudaSetDevice(s->device));
    if (N + 1 > s->cap || !s->pos[0])
        HB2_TRY(system_allocate(s, N + 1, true, N));
    bool not_owned = false;
    return system_set_device(s, N, N, nullptr, pos3, vel3, mass, typeid_, image3, &not_owned);
    }

extern "C" hb2_status hb2_system_set_particles_tagged(hb2_system* s, uint32_t n, uint64_t N_global,
                                                      const uint32_t* tag, const float* pos3,
                                                      const float* vel3, const float* mass,
                                                      const uint32_t* typeid_, const int32_t* image3)
    {
    if (!tag || N_global > 0xfffffffeull)
        {
        set_error("hb2_system_set_particles_tagged: tags are required and must fit 32 bits");
        return HB2_ERR_INVALID;
        }
    if (s && pos3 && s->has_box && s->dom.active && s->pos[0] && (uint64_t)n + 1 <= s->cap
        && N_global == s->dom.N_global)
        {
        // a rank handing back (only) its own particles, e.g. after hb2_system_get_local: no host
        // filtering, the ownership check runs on the device
        HB2_CUDA(cudaSetDevice(s->device));
        bool not_owned = false;
        HB2_TRY(system_set_device(s, n, N_global, tag, pos3, vel3, mass, typeid_, image3, &not_owned));
        if (!not_owned)
            return HB2_OK;
        }
    return system_set(s, n, N_global, tag, pos3, vel3, mass, typeid_, image3);
    }

extern "C" hb2_status hb2_system_set_domain(hb2_system* s, const uint32_t grid[3], uint32_t rank,
                                            uint32_t nranks, const void* nccl_unique_id)
    {
    if (!s || !grid || !nccl_unique_id)
        return HB2_ERR_INVALID;
    if (s->N || s->dom.active)
        {
        set_error("hb2_system_set_domain: must be called once, before particles are set");
        return HB2_ERR_INVALID;
        }
    HB2_CUDA(cudaSetDevice(s->device));
    if (nranks == 1)
        return HB2_OK;
    return domain_setup(s, grid, (int)rank, (int)nranks, nccl_unique_id);
    }

extern "C" hb2_status hb2_system_get_local(hb2_system* s, uint32_t* tag, float* pos3, float* vel3,
                                           float* mass, uint32_t* typeid_, int32_t* image3)
    {
    if (!s)
        return HB2_ERR_INVALID;
    HB2_CUDA(cudaSetDevice(s->device));
    const uint32_t N = s->N;
    if (N == 0)
        return HB2_OK;
    HB2_TRY(system_stage(s, N));
    const size_t n = N;
    const int c = s->cur;
    cudaStream_t st = s->stream;
    float* d = s->d_stage;
    float* h = s->h_stage;
    // local (cell) order; the layout conversion runs on the device, one copy brings the arrays back
    HB2_TRY(snapshot_unpack(N, s->pos[c], s->vel[c], s->image[c], s->tag[c], d, d + 3 * n, d + 6 * n,
                            (uint32_t*)(d + 7 * n), (int*)(d + 8 * n), (uint32_t*)(d + 11 * n), false,
                            st));
    const bool direct = host_pinned(pos3) && host_pinned(vel3) && host_pinned(mass)
                        && host_pinned(typeid_) && host_pinned(image3) && host_pinned(tag);
    if (!direct)
        {
        HB2_CUDA(cudaMemcpyAsync(h, d, 48 * n, cudaMemcpyDeviceToHost, st));
        HB2_CUDA(cudaStreamSynchronize(st));
        }
    auto get = [&](void* dst, size_t off_words, size_t bytes) -> hb2_status
    {
        if (!dst)
            return HB2_OK;
        if (direct)
            HB2_CUDA(cudaMemcpyAsync(dst, d + off_words, bytes, cudaMemcpyDeviceToHost, st));
        else
            memcpy(dst, h + off_words, bytes);
        return HB2_OK;
    };
    HB2_TRY(get(pos3, 0, 12 * n));
    HB2_TRY(get(vel3, 3 * n, 12 * n));
    HB2_TRY(get(mass, 6 * n, 4 * n));
    HB2_TRY(get(typeid_, 7 * n, 4 * n));
    HB2_TRY(get(image3, 8 * n, 12 * n));
    HB2_TRY(get(tag, 11 * n, 4 * n));
    if (direct)
        HB2_CUDA(cudaStreamSynchronize(st));
    return HB2_OK;
    }

extern "C" uint64_t hb2_system_n_global(const hb2_system* s)
    {
    return s ? (s->dom.active ? s->dom.N_global : s->N) : 0;
    }
extern "C" uint32_t hb2_system_n_ghost(const hb2_system* s)
    {
    return s ? s->Ng : 0;
    }

extern "C" hb2_status hb2_system_get_particles(hb2_system* s, float* pos3, float* vel3, float* mass,
                                               uint32_t* typeid_, int32_t* image3)
    {
    if (!s)
        return HB2_ERR_INVALID;
    if (s->dom.active)
        {
        set_error("hb2_system_get_particles: the system is domain decomposed; use "
                  "hb2_system_get_local and gather by tag");
        return HB2_ERR_INVALID;
        }
    HB2_CUDA(cudaSetDevice(s->device));
    const uint32_t N = s->N;
    if (N == 0)
        return HB2_OK;
    HB2_TRY(system_stage(s, N));
    const size_t n = N;
    const int c = s->cur;
    cudaStream_t st = s->stream;
    float* d = s->d_stage;
    float* h = s->h_stage;
    // tag order is restored on the device; one pinned D2H copy brings all five arrays back
    HB2_TRY(snapshot_unpack(N, s->pos[c], s->vel[c], s->image[c], s->tag[c], d, d + 3 * n, d + 6 * n,
                            (uint32_t*)(d + 7 * n), (int*)(d + 8 * n), nullptr, true, st));
    if (host_pinned(pos3) && host_pinned(vel3) && host_pinned(mass) && host_pinned(typeid_)
        && host_pinned(image3))
        {
        // page-locked destinations: DMA straight into the caller's arrays
        if (pos3)
            HB2_CUDA(cudaMemcpyAsync(pos3, d, 12 * n, cudaMemcpyDeviceToHost, st));
        if (vel3)
            HB2_CUDA(cudaMemcpyAsync(vel3, d + 3 * n, 12 * n, cudaMemcpyDeviceToHost, st));
        if (mass)
            HB2_CUDA(cudaMemcpyAsync(mass, d + 6 * n, 4 * n, cudaMemcpyDeviceToHost, st));
        if (typeid_)
            HB2_CUDA(cudaMemcpyAsync(typeid_, d + 7 * n, 4 * n, cudaMemcpyDeviceToHost, st));
        if (image3)
            HB2_CUDA(cudaMemcpyAsync(image3, d + 8 * n, 12 * n, cudaMemcpyDeviceToHost, st));
        HB2_CUDA(cudaStreamSynchronize(st));
        return HB2_OK;
        }
    HB2_CUDA(cudaMemcpyAsync(h, d, 44 * n, cudaMemcpyDeviceToHost, st));
    HB2_CUDA(cudaStreamSynchronize(st));
    if (pos3)
        memcpy(pos3, h, 12 * n);
    if (vel3)
        memcpy(vel3, h + 3 * n, 12 * n);
    if (mass)
        memcpy(mass, h + 6 * n, 4 * n);
    if (typeid_)
        memcpy(typeid_, h + 7 * n, 4 * n);
    if (image3)
        memcpy(image3, h + 8 * n, 12 * n);
    return HB2_OK;
    }

extern "C" uint32_t hb2_system_n(const hb2_system* s)
    {
    return s ? s->N : 0;
    }
extern "C" uint64_t hb2_system_timestep(const hb2_system* s)
    {
    return s ? s->timestep : 0;
    }
extern "C" hb2_status hb2_system_set_timestep(hb2_system* s, uint64_t t)
    {
    if (!s)
        return HB2_ERR_INVALID;
    HB2_CUDA(cudaSetDevice(s->device));
    return s->set_timestep(t);
    }
extern "C" hb2_status hb2_system_set_compute_virial(hb2_system* s, int on)
    {
    if (!s)
        return HB2_ERR_INVALID;
    // flags changed -> the next getter recomputes (hoomd/ForceCompute.cc:562-574): pair_ensure sees
    // compute_virial && !virial_valid and relaunches the force pass over the CURRENT list, so switching
    // the flag never triggers a rebuild that an uninterrupted run would not do
    s->compute_virial = on != 0;
    return HB2_OK;
    }
extern "C" hb2_status hb2_system_synchronize(hb2_system* s)
    {
    if (!s)
        return HB2_ERR_INVALID;
    HB2_CUDA(cudaSetDevice(s->device));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    return HB2_OK;
    }

// =================================================================================================
// neighbour list
// =================================================================================================
hb2_status hb2_nlist::update_r_list()
    {
    // hoomd/md/NeighborList.cc:452-529: r_cut = max over consumers; r_listsq = (r_cut + buff)^2 or 0
    const uint32_t nt = sys->ntypes;
    // the list's own r_cut matrix (hoomd.md.nlist.NeighborList.r_cut, hoomd/md/nlist.py:127-131) is
    // one more consumer matrix
    r_cut.assign((size_t)nt * nt, 0.0);
    if (base_r_cut.size() == r_cut.size())
        r_cut = base_r_cut;
    for (auto* p : consumers)
        for (size_t q = 0; q < r_cut.size(); q++)
            r_cut[q] = std::max(r_cut[q], p->r_cut[q]);
    r_listsq.assign((size_t)nt * nt, 0.0);
    rcut_max.assign(nt, 0.0);
    rcut_max_max = 0.0;
    r_list_min = 0.0;
    for (uint32_t i = 0; i < nt; i++)
        {
        double mi = 0.0;
        for (uint32_t j = 0; j < nt; j++)
            {
            const double rc = r_cut[(size_t)j * nt + i];
            mi = std::max(mi, rc);
            const double r_list = (rc > 0.0) ? rc + buffer : 0.0;
            r_listsq[(size_t)j * nt + i] = r_list * r_list;
            if (r_list > 0.0 && (r_list_min == 0.0 || r_list < r_list_min))
                r_list_min = r_list;
            }
        rcut_max[i] = mi;
        rcut_max_max = std::max(rcut_max_max, mi);
        }
    if (!d_r_listsq)
        HB2_TRY(dev_alloc(&d_r_listsq, (size_t)nt * nt));
    HB2_CUDA(cudaMemcpyAsync(d_r_listsq, r_listsq.data(), sizeof(double) * nt * nt,
                             cudaMemcpyHostToDevice, sys->stream));
    HB2_CUDA(cudaStreamSynchronize(sys->stream));
    params_dirty = false;
    return HB2_OK;
    }

hb2_status hb2_nlist::ensure_capacity()
    {
    const uint32_t cap = sys->cap;
    const Domain& D = sys->dom;
    // cell grid: dim = floor(width / (max r_cut + buffer)), at least 1 (hoomd/CellList.cc:83-131 with
    // the nominal width of hoomd/md/NeighborListBinned.cc:39-45); along a decomposed dimension the
    // width is the sub-domain's and one ghost-cell layer is added on each side (hoomd/CellList.cc
    // does the same through ghost_width, :95-116)
    const double width = rcut_max_max + buffer;
    uint64_t Nc64 = 1;
    for (int k = 0; k < 3; k++)
        {
        const uint32_t n = D.active ? D.grid[k] : 1u;
        uint32_t d = (uint32_t)(sys->box.L[k] / (double)n / width);
        if (d == 0)
            d = 1;
        geom.n[k] = n;
        geom.coord[k] = D.active ? D.coord[k] : 0u;
        geom.dimi[k] = d;
        geom.dimt[k] = n > 1 ? d + 2 : d;
        Nc64 *= geom.dimt[k];
        }
    if (Nc64 >= (1ull << 25)) // the sort key holds the cell id above 6 sub-cell bits, below bit 31
        {
        set_error("nlist: cell grid too large");
        return HB2_ERR_INVALID;
        }
    const uint32_t Nc = (uint32_t)Nc64;
    if (Nc > cell_cap)
        {
        dev_free(d_cell_start);
        dev_free(d_cell_end);
        HB2_TRY(dev_alloc(&d_cell_start, (size_t)Nc));
        HB2_TRY(dev_alloc(&d_cell_end, (size_t)Nc));
        cell_cap = Nc;
        }
    const size_t need_tmp = cell_order_tmp_bytes(cap, Nc);
    if (need_tmp > tmp_bytes)
        {
        if (d_tmp)
            cudaFree(d_tmp);
        d_tmp = nullptr;
        HB2_CUDA(cudaMalloc(&d_tmp, need_tmp));
        tmp_bytes = need_tmp;
        }
    if (cap > cap_N)
        {
        dev_free(d_order);
        dev_free(d_n_neigh);
        dev_free(d_last_pos);
        HB2_TRY(dev_alloc(&d_order, cap));
        HB2_TRY(dev_alloc(&d_n_neigh, cap));
        HB2_TRY(dev_alloc(&d_last_pos, cap));
        cap_N = cap;
        }
    if (pitch == 0)
        {
        // first guess from the mean density; grows on overflow like hoomd/md/NeighborList.cc:1601-1617
        const double vol = sys->box.L[0] * sys->box.L[1] * sys->box.L[2];
        const double rl = rcut_max_max + buffer;
        const double Ng = D.active ? (double)D.N_global : (double)sys->N;
        const double est = Ng / vol * 4.18879020478639 * rl * rl * rl;
        // (the tile kernels hold rows in shared memory: a tight pitch is worth the occasional regrow)
        pitch = (uint32_t)std::min<double>(std::max(16.0, 1.15 * est + 12.0), std::max(Ng, 8.0));
        pitch = (pitch + 31u) & ~31u; // rows are padded to whole 32-entry chunks
        }
    // rows exist for owned particles only; allocate with room for the owned count to drift between
    // builds, and only reallocate when the rows in use no longer fit
    tg = tile_geom(geom);
    const size_t rows = D.active ? std::min<size_t>(cap, (size_t)(sys->N * 1.1) + 1024) : sys->N;
    if (want_tiles())
        {
        if (stage_cap == 0)
            {
            // first guess: 64 stencil cells at the mean occupancy, a third of head-room; grows on overflow
            const double mean = (double)(sys->N + sys->Ng) / (double)Nc;
            stage_cap = ((uint32_t)(64.0 * mean * 1.2) + 32u + 63u) & ~63u;
            }
        if (!d_tile_ticket)
            {
            HB2_TRY(dev_alloc(&d_tile_ticket, 2));
            HB2_CUDA(cudaMemsetAsync(d_tile_ticket, 0, 2 * sizeof(uint32_t), sys->stream));
            }
        if ((size_t)sys->N * pitch > rows16_cap)
            {
            dev_free(d_rows16);
            HB2_TRY(dev_alloc(&d_rows16, rows * pitch));
            rows16_cap = rows * pitch;
            }
        return HB2_OK;
        }
    if ((size_t)sys->N * pitch > nlist_cap)
        {
        dev_free(d_nlist);
        HB2_TRY(dev_alloc(&d_nlist, rows * pitch));
        nlist_cap = rows * pitch;
        }
    return HB2_OK;
    }

// The tile path covers grids with at least three cells per periodic ring and lists without exclusions
// (the exclusion filter edits rows of global indices).
bool hb2_nlist::want_tiles() const
    {
    if (!tile_enabled || !tile_capable(geom))
        return false;
    if ((exclude_bonds && sys->n_bonds) || n_ex_pairs)
        return false;
    // consumers that read rows of global indices on every step (thermostatted evaluators gather
    // velocities and tags; an explicit threads_per_particle selects the per-particle-row kernel)
    for (auto* p : consumers)
        if (p->ev == HB2_EVAL_DPD || p->ev == HB2_EVAL_DPDLJ || p->tpp != 8)
            return false;
    return sys->ntypes <= 16;
    }

TileBuildArgs hb2_nlist::tile_args() const
    {
    hb2_system* s = sys;
    TileBuildArgs a;
    a.pos = s->pos[s->cur];
    a.N = s->N;
    a.geom = geom;
    a.tg = tg;
    a.cell_start = d_cell_start;
    a.cell_end = d_cell_end;
    a.r_listsq = d_r_listsq;
    a.ntypes = s->ntypes;
    a.r_listsq_single = r_listsq[0];
    a.r_listsq_max = *std::max_element(r_listsq.begin(), r_listsq.end());
    a.r_list_min = r_list_min;
    a.Lmax = (float)std::max(s->box.L[0], std::max(s->box.L[1], s->box.L[2]));
    a.rows16 = d_rows16;
    a.pitch = pitch;
    a.n_neigh = d_n_neigh;
    a.last_pos = d_last_pos;
    a.conditions = d_conditions;
    a.stage_cap = stage_cap;
    a.desc = d_tile_desc;
    return a;
    }

// Rows of global indices for the consumers that need them while the tile layout is the live one.
hb2_status hb2_nlist::global_rows()
    {
    if (global_valid || !has_list)
        return HB2_OK;
    hb2_system* s = sys;
    if ((size_t)s->N * pitch > nlist_cap)
        {
        const size_t rows = s->dom.active ? std::min<size_t>(s->cap, (size_t)(s->N * 1.1) + 1024) : s->N;
        dev_free(d_nlist);
        HB2_TRY(dev_alloc(&d_nlist, rows * pitch));
        nlist_cap = rows * pitch;
        }
    TileBuildArgs a = tile_args();
    a.last_pos = nullptr;
    HB2_TRY(launch_tile_rows_to_global(a, d_nlist, s->N + s->Ng, s->stream));
    global_valid = true;
    return HB2_OK;
    }

hb2_status hb2_nlist::rebuild(uint64_t timestep)
    {
    hb2_system* s = sys;
    if (!s->has_box || !s->pos[0] || (s->N == 0 && !s->dom.active))
        {
        set_error("nlist: the system has no box or no particles");
        return HB2_ERR_INVALID;
        }
    if (params_dirty)
        HB2_TRY(update_r_list());
    // hoomd/md/NeighborList.cc:535-561
    const double rmax = rcut_max_max + buffer;
    for (int d = 0; d < 3; d++)
        if (s->box.L[d] <= rmax * 2.0)
            {
            set_error("nlist: Simulation box is too small, the neighbor list is searching beyond "
                      "the minimum image: rmax=" + std::to_string(rmax));
            return HB2_ERR_BOX_TOO_SMALL;
            }
    cudaStream_t st = s->stream;
    // HB2_TRACE_REBUILD=1: host-timed (synchronising) breakdown of every rebuild on stderr — a
    // diagnostic for the multi-GPU path, never on in measurements
    static const bool trace = getenv("HB2_TRACE_REBUILD") != nullptr;
    auto tick = [&](const char* what)
    {
        static double t_prev = 0.0;
        if (!trace)
            return;
        cudaStreamSynchronize(st);
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        const double t = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
        if (what && s->dom.rank == 0)
            fprintf(stderr, "[rebuild] %-10s %8.3f ms\n", what, t - t_prev);
        t_prev = t;
    };
    tick(nullptr);
    if (s->dom.active)
        {
        // hoomd/Communicator.cc:1510-1570: migrate, then rebuild the ghost layer, on rebuild steps
        HB2_TRY(domain_migrate(s));
        tick("migrate");
        HB2_TRY(domain_exchange_ghosts(s, rmax));
        tick("ghosts");
        }
    HB2_TRY(ensure_capacity());
    const uint32_t N = s->N, Ntot = s->N + s->Ng;

    HB2_CUDA(cudaMemsetAsync(d_conditions, 0, 8 * sizeof(uint32_t), st));
    if (s->dom.active && s->rtag) // tags that are neither owned nor ghosts here must read "absent"
        HB2_CUDA(cudaMemsetAsync(s->rtag, 0xff, sizeof(uint32_t) * (size_t)s->dom.N_global, st));
    const int c = s->cur, o = 1 - c;
    HB2_TRY(compute_cell_order(s->pos[c], Ntot, N, s->box, geom, d_order, d_cell_start, d_cell_end,
                               d_conditions, d_tmp, tmp_bytes, st));
    HB2_TRY(reorder_particles(d_order, Ntot, s->pos[c], s->vel[c], s->image[c], s->tag[c], nullptr,
                              s->pos[o], s->vel[o], s->image[o], s->tag[o], nullptr, s->rtag, st));
    s->cur = o;
    s->order_epoch++;
    tick("sort");
    if (s->dom.active)
        HB2_TRY(domain_remap(s, d_order, Ntot));
    HB2_TRY(system_place_dummy(s));
    tick("remap");

    bool tiles = false;
    for (;;)
        {
        tiles = want_tiles();
        if (tiles && tile_build_smem(stage_cap, pitch, s->ntypes) == 0)
            {
            // the stencil of a tile no longer fits in shared memory: per-particle rows from here on
            tile_enabled = false;
            nlist_cap = 0;
            HB2_TRY(ensure_capacity());
            continue;
            }
        if (tiles)
            {
            const size_t need = (size_t)tg.nt[0] * tg.nt[1] * tg.nt[2] * tile_desc_stride(tile_hcap(stage_cap));
            if (need > tile_desc_cap)
                {
                dev_free(d_tile_desc);
                d_tile_desc = nullptr;
                HB2_TRY(dev_alloc(&d_tile_desc, need + need / 8));
                tile_desc_cap = need + need / 8;
                }
            HB2_TRY(launch_nlist_build_tiles(tile_args(), s->box, st));
            }
        else
            {
            NlistBuildArgs a;
            a.pos = s->pos[s->cur];
            a.N = N;
            a.geom = geom;
            a.cell_start = d_cell_start;
            a.cell_end = d_cell_end;
            a.r_listsq = d_r_listsq;
            a.ntypes = s->ntypes;
            a.r_listsq_single = r_listsq[0];
            a.r_list_min = r_list_min;
            a.Lmax = (float)std::max(s->box.L[0], std::max(s->box.L[1], s->box.L[2]));
            a.nlist = d_nlist;
            a.pitch = pitch;
            a.n_neigh = d_n_neigh;
            a.last_pos = d_last_pos;
            a.conditions = d_conditions;
            a.dummy_index = Ntot;
            HB2_TRY(launch_nlist_build(a, s->box, st));
            }
        if (s->dom.active) // every rank must take the same branch below
            HB2_TRY(domain_allreduce_max_u32(s, d_conditions, 6));
        HB2_CUDA(cudaMemcpyAsync(h_conditions, d_conditions, 8 * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, st));
        HB2_CUDA(cudaStreamSynchronize(st));
        if (h_conditions[1])
            {
            set_error("Particle with index " + std::to_string(h_conditions[1] - 1)
                      + " has NaN for its position.");
            return HB2_ERR_NAN_POSITION;
            }
        if (h_conditions[2])
            {
            set_error("Particle with index " + std::to_string(h_conditions[2] - 1)
                      + " is no longer in the simulation box.");
            return HB2_ERR_OUT_OF_BOX;
            }
        if (tiles && (h_conditions[4] || h_conditions[5]))
            {
            // a tile's stencil (or one home cell's 27 cells) holds more particles than the staging
            // capacity: grow it and build again
            const uint32_t need = std::max(h_conditions[4], (h_conditions[5] * 16u + 8u) / 9u);
            stage_cap = ((uint32_t)(need * 1.125) + 64u + 63u) & ~63u;
            HB2_CUDA(cudaMemsetAsync(d_conditions, 0, 8 * sizeof(uint32_t), st));
            continue;
            }
        max_neigh = h_conditions[0];
        tick("build");
        if (max_neigh <= pitch)
            break;
        // overflow: regrow to a multiple of 8 with headroom and rebuild
        // (hoomd/md/NeighborList.cc:411-427, :1601-1617)
        pitch = ((uint32_t)(max_neigh * 1.125) + 8u + 31u) & ~31u;
        nlist_cap = 0;
        rows16_cap = 0;
        HB2_TRY(ensure_capacity());
        HB2_CUDA(cudaMemsetAsync(d_conditions, 0, 8 * sizeof(uint32_t), st));
        }
    tile_active = tiles;
    global_valid = !tiles;
    // exclusions=('bond',): drop bonded pairs from the rows (hoomd/md/NeighborList.cc:1487-1536)
    if (exclude_bonds && s->n_bonds)
        HB2_TRY(launch_nlist_filter_bonds(N, s->tag[s->cur], s->rtag, s->d_bond_head, s->d_bond_partner,
                                          d_nlist, pitch, d_n_neigh, Ntot, st));
    if (n_ex_pairs)
        HB2_TRY(launch_nlist_filter_bonds(N, s->tag[s->cur], s->rtag, d_ex_head, d_ex_partner, d_nlist,
                                          pitch, d_n_neigh, Ntot, st));
    if (has_list && timestep > last_update_step)
        shortest_rebuild = std::min<uint64_t>(shortest_rebuild, timestep - last_update_step);
    has_list = true;
    force_update = false;
    for (int d = 0; d < 3; d++)
        last_L[d] = s->box.L[d];
    num_builds++;
    last_update_step = timestep;
    return HB2_OK;
    }

hb2_status hb2_nlist::compute(uint64_t timestep, bool flag_known, bool flag_value)
    {
    // hoomd/md/NeighborList.cc:363-435 + needsUpdating :1325-1418
    if (params_dirty)
        {
        HB2_TRY(update_r_list());
        force_update = true;
        }
    bool need = force_update || !has_list;
    if (!need)
        {
        if (buffer < 1e-6)
            need = true;
        else if (timestep >= last_update_step + rebuild_check_delay)
            {
            if (!check_dist)
                need = true;
            else if (flag_known)
                need = flag_value;
            else
                {
                // standalone distance check against last_pos (positions may have been set outside)
                need = true;
                }
            }
        }
    if (need)
        return rebuild(timestep);
    return HB2_OK;
    }

extern "C" hb2_status hb2_nlist_create(hb2_nlist** out, hb2_system* s, double buffer,
                                       uint32_t rebuild_check_delay, int check_dist)
    {
    if (!out || !s || buffer < 0)
        {
        set_error("hb2_nlist_create: bad arguments");
        return HB2_ERR_INVALID;
        }
    HB2_CUDA(cudaSetDevice(s->device));
    hb2_nlist* nl = new hb2_nlist();
    nl->sys = s;
    nl->buffer = buffer;
    nl->rebuild_check_delay = rebuild_check_delay;
    nl->check_dist = check_dist != 0;
    if (const char* e = getenv("HB2_TILE"))
        nl->tile_enabled = atoi(e) != 0;
    hb2_status st = dev_alloc(&nl->d_conditions, 8);
    if (st == HB2_OK && cudaMallocHost((void**)&nl->h_conditions, 8 * sizeof(uint32_t)) != cudaSuccess)
        st = cuda_fail(cudaGetLastError(), "cudaMallocHost");
    if (st != HB2_OK)
        {
        delete nl;
        return st;
        }
    cudaMemsetAsync(nl->d_conditions, 0, 8 * sizeof(uint32_t), s->stream);
    s->nlists.push_back(nl);
    *out = nl;
    return HB2_OK;
    }

extern "C" void hb2_nlist_destroy(hb2_nlist* nl)
    {
    if (!nl)
        return;
    hb2_system* s = nl->sys;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    s->nlists.erase(std::remove(s->nlists.begin(), s->nlists.end(), nl), s->nlists.end());
    dev_free(nl->d_cell_start);
    dev_free(nl->d_cell_end);
    dev_free(nl->d_order);
    if (nl->d_tmp)
        cudaFree(nl->d_tmp);
    dev_free(nl->d_r_listsq);
    dev_free(nl->d_nlist);
    dev_free(nl->d_rows16);
    dev_free(nl->d_tile_ticket);
    dev_free(nl->d_tile_desc);
    dev_free(nl->d_n_neigh);
    dev_free(nl->d_last_pos);
    dev_free(nl->d_ex_head);
    dev_free(nl->d_ex_partner);
    dev_free(nl->d_conditions);
    if (nl->h_conditions)
        cudaFreeHost(nl->h_conditions);
    delete nl;
    }

extern "C" hb2_status hb2_nlist_set_buffer(hb2_nlist* nl, double buffer)
    {
    if (!nl || buffer < 0)
        return HB2_ERR_INVALID;
    nl->buffer = buffer;
    nl->params_dirty = true;
    nl->force_update = true;
    nl->pitch = 0; // re-estimate
    return HB2_OK;
    }

extern "C" hb2_status hb2_nlist_set_r_cut(hb2_nlist* nl, uint32_t ti, uint32_t tj, double r_cut)
    {
    if (!nl || ti >= nl->sys->ntypes || tj >= nl->sys->ntypes || r_cut < 0)
        {
        set_error("hb2_nlist_set_r_cut: bad arguments");
        return HB2_ERR_INVALID;
        }
    const uint32_t nt = nl->sys->ntypes;
    if (nl->base_r_cut.size() != (size_t)nt * nt)
        nl->base_r_cut.assign((size_t)nt * nt, 0.0);
    nl->base_r_cut[(size_t)tj * nt + ti] = nl->base_r_cut[(size_t)ti * nt + tj] = r_cut;
    nl->params_dirty = true;
    nl->force_update = true;
    return HB2_OK;
    }

// The list of the last build as the reference lays it out (hoomd/md/NeighborList.h:662-722
// LocalNeighborListData: head_list, n_neigh, nlist): device pointers valid until the next rebuild. Row i
// starts at i * pitch (the head list is implicit), entries in local particle indices.
extern "C" hb2_status hb2_nlist_device_arrays(hb2_nlist* nl, const uint32_t** d_n_neigh,
                                              const uint32_t** d_nlist, uint32_t* pitch, uint32_t* n_rows)
    {
    if (!nl || !d_n_neigh || !d_nlist || !pitch || !n_rows)
        return HB2_ERR_INVALID;
    if (!nl->has_list)
        {
        set_error("nlist: no neighbour list has been built yet");
        return HB2_ERR_INVALID;
        }
    HB2_CUDA(cudaSetDevice(nl->sys->device));
    HB2_TRY(nl->global_rows());
    HB2_CUDA(cudaStreamSynchronize(nl->sys->stream));
    *d_n_neigh = nl->d_n_neigh;
    *d_nlist = nl->d_nlist;
    *pitch = nl->pitch;
    *n_rows = nl->sys->N;
    return HB2_OK;
    }

extern "C" hb2_status hb2_nlist_compute(hb2_nlist* nl)
    {
    if (!nl)
        return HB2_ERR_INVALID;
    HB2_CUDA(cudaSetDevice(nl->sys->device));
    nl->force_update = true;
    return nl->compute(nl->sys->timestep, false, false);
    }
extern "C" uint64_t hb2_nlist_num_builds(const hb2_nlist* nl)
    {
    return nl ? nl->num_builds : 0;
    }
extern "C" uint64_t hb2_nlist_shortest_rebuild(const hb2_nlist* nl)
    {
    return nl ? (nl->shortest_rebuild == ~0ull ? 0 : nl->shortest_rebuild) : 0;
    }
extern "C" hb2_status hb2_nlist_dimensions(const hb2_nlist* nl, uint32_t dim[3])
    {
    if (!nl || !dim)
        return HB2_ERR_INVALID;
    for (int d = 0; d < 3; d++)
        dim[d] = nl->geom.dimi[d];
    return HB2_OK;
    }

static hb2_status nlist_download(hb2_nlist* nl, std::vector<uint32_t>& n_neigh,
                                 std::vector<uint32_t>& list, std::vector<uint32_t>& tag)
    {
    hb2_system* s = nl->sys;
    if (!nl->has_list)
        {
        set_error("nlist: no neighbour list has been built yet");
        return HB2_ERR_INVALID;
        }
    HB2_TRY(nl->global_rows());
    const uint32_t N = s->N;
    n_neigh.resize(N);
    tag.resize((size_t)N + s->Ng);
    list.resize((size_t)N * nl->pitch);
    HB2_CUDA(cudaMemcpyAsync(n_neigh.data(), nl->d_n_neigh, sizeof(uint32_t) * N,
                             cudaMemcpyDeviceToHost, s->stream));
    HB2_CUDA(cudaMemcpyAsync(tag.data(), s->tag[s->cur], sizeof(uint32_t) * tag.size(),
                             cudaMemcpyDeviceToHost, s->stream));
    HB2_CUDA(cudaMemcpyAsync(list.data(), nl->d_nlist, sizeof(uint32_t) * (size_t)N * nl->pitch,
                             cudaMemcpyDeviceToHost, s->stream));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    return HB2_OK;
    }

extern "C" hb2_status hb2_nlist_pair_count(hb2_nlist* nl, uint64_t* n_pairs)
    {
    if (!nl || !n_pairs)
        return HB2_ERR_INVALID;
    HB2_CUDA(cudaSetDevice(nl->sys->device));
    if (!nl->has_list)
        {
        set_error("nlist: no neighbour list has been built yet");
        return HB2_ERR_INVALID;
        }
    std::vector<uint32_t> nn(nl->sys->N);
    HB2_CUDA(cudaMemcpyAsync(nn.data(), nl->d_n_neigh, sizeof(uint32_t) * nn.size(),
                             cudaMemcpyDeviceToHost, nl->sys->stream));
    HB2_CUDA(cudaStreamSynchronize(nl->sys->stream));
    uint64_t t = 0;
    for (uint32_t v : nn)
        t += v;
    *n_pairs = t / 2; // full list: every unordered pair is stored twice
    if (nl->sys->dom.active)
        *n_pairs = t; // upper bound: pairs with a ghost are stored once on each of the two owners
    return HB2_OK;
    }

extern "C" hb2_status hb2_nlist_pair_list(hb2_nlist* nl, uint32_t* pairs, uint64_t cap,
                                          uint64_t* n_pairs)
    {
    if (!nl || !pairs || !n_pairs)
        return HB2_ERR_INVALID;
    HB2_CUDA(cudaSetDevice(nl->sys->device));
    std::vector<uint32_t> nn, list, tag;
    HB2_TRY(nlist_download(nl, nn, list, tag));
    uint64_t n = 0;
    const uint32_t N = nl->sys->N;
    for (uint32_t i = 0; i < N; i++)
        for (uint32_t k = 0; k < nn[i]; k++)
            {
            const uint32_t j = list[(size_t)i * nl->pitch + k];
            // half-list view: each unordered pair once (hoomd/md/NeighborList.cc:1814-1828). A pair
            // with a ghost is also listed by the ghost's owner: the smaller tag keeps it.
            if (j < N ? i < j : tag[i] < tag[j])
                {
                if (n < cap)
                    {
                    pairs[2 * n] = tag[i];
                    pairs[2 * n + 1] = tag[j];
                    }
                n++;
                }
            }
    *n_pairs = n;
    if (n > cap)
        {
        set_error("hb2_nlist_pair_list: output buffer too small");
        return HB2_ERR_INVALID;
        }
    return HB2_OK;
    }

extern "C" hb2_status hb2_nlist_stats(hb2_nlist* nl, double* mean_neigh, uint32_t* max_neigh,
                                      uint32_t* pitch)
    {
    if (!nl)
        return HB2_ERR_INVALID;
    uint64_t np = 0;
    HB2_TRY(hb2_nlist_pair_count(nl, &np));
    if (nl->sys->dom.active)
        np /= 2;
    if (mean_neigh)
        *mean_neigh = nl->sys->N ? 2.0 * (double)np / nl->sys->N : 0.0;
    if (max_neigh)
        *max_neigh = nl->max_neigh;
    if (pitch)
        *pitch = nl->pitch;
    return HB2_OK;
    }

// =================================================================================================
// pair force
// =================================================================================================
// one entry per evaluator id: user-parameter count and the host-side pack
template<class Ev> static void pack_as(const double* user, float* packed)
    {
    typename Ev::param_type p;
    Ev::pack(user, p);
    memset(packed, 0, HB2_PARAM_FLOATS * sizeof(float));
    memcpy(packed, &p, sizeof(p));
    }
struct EvalInfo
    {
    unsigned n_user;
    void (*pack)(const double*, float*);
    };
static void pack_dpd_conservative(const double* user, float* packed)
    {
    memset(packed, 0, HB2_PARAM_FLOATS * sizeof(float));
    packed[0] = (float)user[0];
    }
static const EvalInfo* eval_info(hb2_evaluator ev)
    {
    static const EvalInfo table[] = {
        {EvaluatorPairLJ::n_user_params, pack_as<EvaluatorPairLJ>},
        {EvaluatorPairDPDThermoDPD::n_user_params, pack_as<EvaluatorPairDPDThermoDPD>},
        {1, pack_dpd_conservative},
        {EvaluatorPairExample::n_user_params, pack_as<EvaluatorPairExample>},
        {EvaluatorPairGauss::n_user_params, pack_as<EvaluatorPairGauss>},
        {EvaluatorPairYukawa::n_user_params, pack_as<EvaluatorPairYukawa>},
        {EvaluatorPairLJ1208::n_user_params, pack_as<EvaluatorPairLJ1208>},
        {EvaluatorPairLJ0804::n_user_params, pack_as<EvaluatorPairLJ0804>},
        {EvaluatorPairForceShiftedLJ::n_user_params, pack_as<EvaluatorPairForceShiftedLJ>},
        {EvaluatorPairMie::n_user_params, pack_as<EvaluatorPairMie>},
        {EvaluatorPairMorse::n_user_params, pack_as<EvaluatorPairMorse>},
        {EvaluatorPairBuckingham::n_user_params, pack_as<EvaluatorPairBuckingham>},
        {EvaluatorPairExpandedLJ::n_user_params, pack_as<EvaluatorPairExpandedLJ>},
        {EvaluatorPairExpandedGaussian::n_user_params, pack_as<EvaluatorPairExpandedGaussian>},
        {EvaluatorPairLJGauss::n_user_params, pack_as<EvaluatorPairLJGauss>},
        {EvaluatorPairTWF::n_user_params, pack_as<EvaluatorPairTWF>},
        {EvaluatorPairDPDThermoLJ::n_user_params, pack_as<EvaluatorPairDPDThermoLJ>},
        {0, nullptr}, // Table: variable-size parameters, hb2_pair_set_table
        {EvaluatorPairExpandedMie::n_user_params, pack_as<EvaluatorPairExpandedMie>},
        {EvaluatorPairOPP::n_user_params, pack_as<EvaluatorPairOPP>},
        {EvaluatorPairFourier::n_user_params, pack_as<EvaluatorPairFourier>},
    };
    if ((unsigned)ev >= sizeof(table) / sizeof(table[0]))
        return nullptr;
    return &table[(unsigned)ev];
    }
static unsigned n_user_params(hb2_evaluator ev)
    {
    const EvalInfo* e = eval_info(ev);
    return e ? e->n_user : 0;
    }

extern "C" hb2_status hb2_pack_params(hb2_evaluator ev, const double* user, uint32_t n_user,
                                      float packed[HB2_PARAM_FLOATS])
    {
    const EvalInfo* e = eval_info(ev);
    if (!e)
        {
        set_error("hb2_pack_params: unknown evaluator");
        return HB2_ERR_INVALID;
        }
    if (!e->pack)
        {
        set_error("hb2_pack_params: this evaluator's parameters are not a fixed list (Table: hb2_pair_set_table)");
        return HB2_ERR_INVALID;
        }
    if (!user || !packed || n_user < e->n_user)
        {
        set_error("hb2_pack_params: not enough user parameters for this evaluator");
        return HB2_ERR_INVALID;
        }
    e->pack(user, packed);
    return HB2_OK;
    }

namespace hb2
    {
hb2_status launch_pair_forces(hb2_evaluator ev, hb2_shift_mode mode, const PairArgs& a,
                              const hb2_box& box, cudaStream_t stream)
    {
    switch (ev)
        {
    case HB2_EVAL_LJ:
        return pair_forces_lj(mode, a, box, stream);
    case HB2_EVAL_DPD:
        return pair_forces_dpd(mode, a, box, stream);
    case HB2_EVAL_DPD_CONSERVATIVE:
        return pair_forces_dpd_conservative(mode, a, box, stream);
    case HB2_EVAL_EXAMPLE:
        return pair_forces_example(mode, a, box, stream);
    case HB2_EVAL_GAUSS:
        return pair_forces_gauss(mode, a, box, stream);
    case HB2_EVAL_YUKAWA:
        return pair_forces_yukawa(mode, a, box, stream);
    case HB2_EVAL_LJ1208:
        return pair_forces_lj1208(mode, a, box, stream);
    case HB2_EVAL_LJ0804:
        return pair_forces_lj0804(mode, a, box, stream);
    case HB2_EVAL_FORCE_SHIFTED_LJ:
        return pair_forces_force_shifted_lj(mode, a, box, stream);
    case HB2_EVAL_MIE:
        return pair_forces_mie(mode, a, box, stream);
    case HB2_EVAL_MORSE:
        return pair_forces_morse(mode, a, box, stream);
    case HB2_EVAL_BUCKINGHAM:
        return pair_forces_buckingham(mode, a, box, stream);
    case HB2_EVAL_EXPANDED_LJ:
        return pair_forces_expanded_lj(mode, a, box, stream);
    case HB2_EVAL_EXPANDED_GAUSSIAN:
        return pair_forces_expanded_gaussian(mode, a, box, stream);
    case HB2_EVAL_LJ_GAUSS:
        return pair_forces_lj_gauss(mode, a, box, stream);
    case HB2_EVAL_TWF:
        return pair_forces_twf(mode, a, box, stream);
    case HB2_EVAL_DPDLJ:
        return pair_forces_dpdlj(mode, a, box, stream);
    case HB2_EVAL_TABLE:
        return pair_forces_table(mode, a, box, stream);
    case HB2_EVAL_EXPANDED_MIE:
        return pair_forces_expanded_mie(mode, a, box, stream);
    case HB2_EVAL_OPP:
        return pair_forces_opp(mode, a, box, stream);
    case HB2_EVAL_FOURIER:
        return pair_forces_fourier(mode, a, box, stream);
        }
    set_error("pair forces: unknown evaluator");
    return HB2_ERR_INVALID;
    }
    } // namespace hb2

hb2_status hb2_pair::upload_params()
    {
    const uint32_t nt = sys->ntypes;
    const size_t nt2 = (size_t)nt * nt;
    packed.assign(HB2_PARAM_FLOATS * nt2, 0.f);
    rcutsq.assign(nt2, 0.f);
    ronsq.assign(nt2, 0.f);
    rcutsq_d.assign(nt2, 0.0);
    for (size_t q = 0; q < nt2; q++)
        {
        if (!have_params[q])
            {
            set_error("pair: parameters are not set for every type pair");
            return HB2_ERR_INVALID;
            }
        if (ev != HB2_EVAL_TABLE)
            HB2_TRY(hb2_pack_params(ev, &user[HB2_PARAM_FLOATS * q], HB2_PARAM_FLOATS, &packed[HB2_PARAM_FLOATS * q]));
        rcutsq_d[q] = r_cut[q] * r_cut[q]; // hoomd/md/PotentialPair.h:498
        rcutsq[q] = (float)rcutsq_d[q];
        ronsq[q] = (float)(r_on[q] * r_on[q]);
        }
    if (ev == HB2_EVAL_TABLE)
        {
        // all tables back to back in one device buffer (V then F per type pair); param_type points into it
        size_t total = 0;
        for (size_t q = 0; q < nt2; q++)
            total += tables[q].size();
        std::vector<float> flat;
        flat.reserve(total);
        std::vector<size_t> off(nt2);
        for (size_t q = 0; q < nt2; q++)
            {
            off[q] = flat.size();
            flat.insert(flat.end(), tables[q].begin(), tables[q].end());
            }
        if (total > table_cap)
            {
            dev_free(d_tables);
            HB2_TRY(dev_alloc(&d_tables, total));
            table_cap = total;
            }
        if (total)
            HB2_CUDA(cudaMemcpyAsync(d_tables, flat.data(), sizeof(float) * total, cudaMemcpyHostToDevice,
                                     sys->stream));
        HB2_CUDA(cudaStreamSynchronize(sys->stream));
        for (size_t q = 0; q < nt2; q++)
            {
            EvaluatorPairTable::param_type pt;
            pt.rmin = (float)table_rmin[q];
            pt.width = (uint32_t)(tables[q].size() / 2);
            pt.table = d_tables + off[q];
            memcpy(&packed[HB2_PARAM_FLOATS * q], &pt, sizeof(pt));
            }
        }
    if (!d_params)
        {
        HB2_TRY(dev_alloc(&d_params, HB2_PARAM_FLOATS * nt2));
        HB2_TRY(dev_alloc(&d_rcutsq, nt2));
        HB2_TRY(dev_alloc(&d_ronsq, nt2));
        HB2_TRY(dev_alloc(&d_rcutsq_d, nt2));
        }
    cudaStream_t st = sys->stream;
    HB2_CUDA(cudaMemcpyAsync(d_params, packed.data(), sizeof(float) * HB2_PARAM_FLOATS * nt2, cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaMemcpyAsync(d_rcutsq, rcutsq.data(), sizeof(float) * nt2, cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaMemcpyAsync(d_ronsq, ronsq.data(), sizeof(float) * nt2, cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaMemcpyAsync(d_rcutsq_d, rcutsq_d.data(), sizeof(double) * nt2, cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaStreamSynchronize(st));
    params_dirty = false;
    return HB2_OK;
    }

hb2_status hb2_pair::compute(uint64_t timestep, bool flag_known, bool flag_value, bool want_energy)
    {
    hb2_system* s = sys;
    if (params_dirty)
        HB2_TRY(upload_params());
    const uint64_t builds_before = nl->num_builds;
    HB2_TRY(nl->compute(timestep, flag_known, flag_value));
    if (s->dom.active)
        {
        // DPD reads the velocities of its neighbours (hoomd/md/PotentialPairDPDThermo.h:274-286
        // requests the ghost velocity flag for the same reason)
        if ((ev == HB2_EVAL_DPD || ev == HB2_EVAL_DPDLJ) && !s->dom.need_vel)
            s->dom.need_vel = true;
        // a rebuild has just exchanged the ghost layer; otherwise refresh it
        // (Communicator::beginUpdateGhosts, hoomd/Communicator.cc:2630)
        if (nl->num_builds == builds_before)
            HB2_TRY(domain_update_ghosts(s));
        }
    return launch(timestep, want_energy, nullptr);
    }

// The force pass over the current list. skip_flag (device word) makes the kernel a no-op when set:
// the step loop launches it before the host knows whether this step rebuilds (see hb2_integrator_run).
hb2_status hb2_pair::launch(uint64_t timestep, bool want_energy, const uint32_t* skip_flag)
    {
    hb2_system* s = sys;
    const uint32_t N = s->N;
    extra_added = false;
    if (s->cap > cap_N)
        {
        dev_free(d_force);
        dev_free(d_virial);
        HB2_TRY(dev_alloc(&d_force, s->cap));
        HB2_TRY(dev_alloc(&d_virial, 6 * (size_t)s->cap));
        cap_N = s->cap;
        computed = false;
        }
    PairArgs a;
    a.pos = s->pos[s->cur];
    a.vel = s->vel[s->cur];
    a.tag = s->tag[s->cur];
    a.N = N;
    a.n_neigh = nl->d_n_neigh;
    a.nlist = nl->d_nlist;
    a.pitch = nl->pitch;
    if (nl->tile_active)
        {
        // thermostatted evaluators gather velocities and tags by global index; a non-default
        // threads_per_particle asks for the per-particle-row kernel explicitly
        if (ev == HB2_EVAL_DPD || ev == HB2_EVAL_DPDLJ || tpp != 8)
            {
            HB2_TRY(nl->global_rows());
            a.nlist = nl->d_nlist;
            }
        else
            {
            a.rows16 = nl->d_rows16;
            a.tg = nl->tg;
            a.geom = nl->geom;
            a.cell_start = nl->d_cell_start;
            a.cell_end = nl->d_cell_end;
            a.stage_cap = nl->stage_cap;
            a.tile_ticket = nl->d_tile_ticket;
            a.tile_desc = nl->d_tile_desc;
            }
        }
    a.params = d_params;
    a.rcutsq = d_rcutsq;
    a.ronsq = d_ronsq;
    for (int q = 0; q < HB2_PARAM_FLOATS; q++)
        a.params0[q] = packed[q];
    a.rcutsq0 = rcutsq[0];
    a.ronsq0 = ronsq[0];
    a.rcutsq_d = d_rcutsq_d;
    a.rcutsq0_d = rcutsq_d[0];
    a.ntypes = s->ntypes;
    a.force = d_force;
    a.virial = s->compute_virial ? d_virial : nullptr;
    a.virial_pitch = cap_N;
    a.seed = s->seed;
    a.d_timestep = nullptr;
    a.timestep = timestep;
    a.dt = (float)dt;
    a.kT = (float)kT;
    // A pair across a periodic face with a separation below r_cut needs its home particle within r_cut of that
    // face. Farther away, a listed neighbour may still sit in another image (the list reaches r_cut + buffer and
    // both particles have moved since), but then its true separation and the plain difference (>= L - r_list
    // >= r_list) are both beyond the cut-off: no contribution either way, so the plain loop serves.
    a.face_margin = (float)nl->rcut_max_max * 1.0001f;
    a.tpp = tpp;
    a.energy = want_energy || energy_every_step || tpp != 8;
    a.skip_flag = skip_flag;
    if (s->mid_event)
        HB2_CUDA(cudaEventRecord(s->mid_event, s->stream));
    HB2_TRY(launch_pair_forces(ev, mode, a, s->box, s->stream));
    virial_valid = s->compute_virial;
    energy_valid = a.energy;
    computed = true;
    computed_step = timestep;
    computed_epoch = s->order_epoch;
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_create(hb2_pair** out, hb2_system* s, hb2_nlist* nl, hb2_evaluator ev,
                                      hb2_shift_mode mode)
    {
    if (!out || !s || !nl || nl->sys != s)
        {
        set_error("hb2_pair_create: bad arguments");
        return HB2_ERR_INVALID;
        }
    hb2_pair* p = new hb2_pair();
    p->sys = s;
    p->nl = nl;
    p->ev = ev;
    p->mode = mode;
    const size_t nt2 = (size_t)s->ntypes * s->ntypes;
    p->user.assign(HB2_PARAM_FLOATS * nt2, 0.0);
    p->have_params.assign(nt2, 0);
    p->r_cut.assign(nt2, 0.0);
    p->r_on.assign(nt2, 0.0);
    nl->consumers.push_back(p); // addRCutMatrix, hoomd/md/PotentialPair.h:369-371
    nl->params_dirty = true;
    s->pairs.push_back(p);
    *out = p;
    return HB2_OK;
    }

extern "C" void hb2_pair_destroy(hb2_pair* p)
    {
    if (!p)
        return;
    hb2_system* s = p->sys;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    auto& c = p->nl->consumers;
    c.erase(std::remove(c.begin(), c.end(), p), c.end());
    p->nl->params_dirty = true;
    s->pairs.erase(std::remove(s->pairs.begin(), s->pairs.end(), p), s->pairs.end());
    dev_free(p->d_params);
    dev_free(p->d_rcutsq);
    dev_free(p->d_ronsq);
    dev_free(p->d_rcutsq_d);
    dev_free(p->d_tables);
    dev_free(p->d_force);
    dev_free(p->d_virial);
    delete p;
    }

static bool pair_index(hb2_pair* p, uint32_t ti, uint32_t tj, size_t& a, size_t& b)
    {
    const uint32_t nt = p->sys->ntypes;
    if (ti >= nt || tj >= nt)
        {
        set_error("pair: type index out of range");
        return false;
        }
    a = (size_t)tj * nt + ti;
    b = (size_t)ti * nt + tj;
    return true;
    }

extern "C" hb2_status hb2_pair_set_params(hb2_pair* p, uint32_t ti, uint32_t tj,
                                          const double* user_params, uint32_t n_user)
    {
    size_t a, b;
    if (!p || !user_params || !pair_index(p, ti, tj, a, b))
        return HB2_ERR_INVALID;
    if (n_user < n_user_params(p->ev) || n_user > HB2_PARAM_FLOATS)
        {
        set_error("hb2_pair_set_params: wrong number of parameters for this evaluator");
        return HB2_ERR_INVALID;
        }
    for (size_t q : {a, b})
        {
        for (uint32_t k = 0; k < HB2_PARAM_FLOATS; k++)
            p->user[HB2_PARAM_FLOATS * q + k] = k < n_user ? user_params[k] : 0.0;
        p->have_params[q] = 1;
        }
    p->params_dirty = true;
    p->computed = false;
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_set_table(hb2_pair* p, uint32_t ti, uint32_t tj, double r_min, uint32_t width,
                                         const double* U, const double* F)
    {
    size_t a, b;
    if (!p || !U || !F || !pair_index(p, ti, tj, a, b))
        return HB2_ERR_INVALID;
    if (p->ev != HB2_EVAL_TABLE)
        {
        set_error("hb2_pair_set_table: not a Table pair force");
        return HB2_ERR_INVALID;
        }
    if (width == 0)
        {
        set_error("The length of V and F must not be zero."); // hoomd/md/EvaluatorPairTable.h:85-88
        return HB2_ERR_INVALID;
        }
    const size_t nt2 = (size_t)p->sys->ntypes * p->sys->ntypes;
    p->tables.resize(nt2);
    p->table_rmin.resize(nt2, 0.0);
    for (size_t q : {a, b})
        {
        p->tables[q].resize(2 * (size_t)width);
        for (uint32_t k = 0; k < width; k++)
            {
            p->tables[q][k] = (float)U[k];
            p->tables[q][width + k] = (float)F[k];
            }
        p->table_rmin[q] = r_min;
        p->have_params[q] = 1;
        }
    p->params_dirty = true;
    p->computed = false;
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_set_r_cut(hb2_pair* p, uint32_t ti, uint32_t tj, double r_cut)
    {
    size_t a, b;
    if (!p || !pair_index(p, ti, tj, a, b))
        return HB2_ERR_INVALID;
    p->r_cut[a] = p->r_cut[b] = r_cut;
    p->params_dirty = true;
    p->computed = false;
    p->nl->params_dirty = true; // notifyRCutMatrixChange
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_set_r_on(hb2_pair* p, uint32_t ti, uint32_t tj, double r_on)
    {
    size_t a, b;
    if (!p || !pair_index(p, ti, tj, a, b))
        return HB2_ERR_INVALID;
    p->r_on[a] = p->r_on[b] = r_on;
    p->params_dirty = true;
    p->computed = false;
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_set_kT(hb2_pair* p, double kT)
    {
    if (!p)
        return HB2_ERR_INVALID;
    p->kT = kT;
    p->computed = false;
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_set_threads_per_particle(hb2_pair* p, uint32_t tpp)
    {
    if (!p || (tpp != 4 && tpp != 8 && tpp != 16 && tpp != 32))
        {
        set_error("threads_per_particle must be 4, 8, 16 or 32");
        return HB2_ERR_INVALID;
        }
    p->tpp = tpp;
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_compute(hb2_pair* p, uint64_t timestep)
    {
    if (!p)
        return HB2_ERR_INVALID;
    HB2_CUDA(cudaSetDevice(p->sys->device));
    return p->compute(timestep);
    }

static hb2_status pair_ensure(hb2_pair* p, bool need_energy = true)
    {
    // recompute when the particle order changed or nothing was computed yet
    // (hoomd/ForceCompute.cc:562-574: "sorted / new timestep / flags changed")
    if (!p->computed || p->computed_epoch != p->sys->order_epoch)
        return p->compute(p->sys->timestep);
    // Same particles, same positions, same list: only the array's content is not what the getter wants
    // (bond forces were added into it, or the energy / virial variant did not run). Launch the force
    // pass again over the current list — going through the neighbour list's update logic would
    // rebuild (and re-order the particles) where an uninterrupted run would not.
    if (p->extra_added || (p->sys->compute_virial && !p->virial_valid) || (need_energy && !p->energy_valid))
        return p->launch(p->computed_step, true, nullptr);
    return HB2_OK;
    }

static hb2_status pair_download(hb2_pair* p, std::vector<float4>& f, std::vector<float>* vir,
                                std::vector<uint32_t>& tag)
    {
    hb2_system* s = p->sys;
    HB2_CUDA(cudaSetDevice(s->device));
    HB2_TRY(pair_ensure(p));
    const uint32_t N = s->N;
    f.resize(N);
    tag.resize(N);
    HB2_CUDA(cudaMemcpyAsync(f.data(), p->d_force, sizeof(float4) * N, cudaMemcpyDeviceToHost,
                             s->stream));
    HB2_CUDA(cudaMemcpyAsync(tag.data(), s->tag[s->cur], sizeof(uint32_t) * N, cudaMemcpyDeviceToHost,
                             s->stream));
    if (vir)
        {
        vir->resize(6 * (size_t)p->cap_N);
        HB2_CUDA(cudaMemcpyAsync(vir->data(), p->d_virial, sizeof(float) * 6 * (size_t)p->cap_N,
                                 cudaMemcpyDeviceToHost, s->stream));
        }
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_get_forces(hb2_pair* p, double* forces3)
    {
    if (!p || !forces3)
        return HB2_ERR_INVALID;
    if (p->sys->dom.active)
        {
        set_error("the system is domain decomposed: use hb2_pair_get_local and gather by tag");
        return HB2_ERR_INVALID;
        }
    std::vector<float4> f;
    std::vector<uint32_t> tag;
    HB2_TRY(pair_download(p, f, nullptr, tag));
    for (size_t k = 0; k < f.size(); k++)
        {
        forces3[3 * (size_t)tag[k]] = f[k].x;
        forces3[3 * (size_t)tag[k] + 1] = f[k].y;
        forces3[3 * (size_t)tag[k] + 2] = f[k].z;
        }
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_get_energies(hb2_pair* p, double* energies)
    {
    if (!p || !energies)
        return HB2_ERR_INVALID;
    if (p->sys->dom.active)
        {
        set_error("the system is domain decomposed: use hb2_pair_get_local and gather by tag");
        return HB2_ERR_INVALID;
        }
    std::vector<float4> f;
    std::vector<uint32_t> tag;
    HB2_TRY(pair_download(p, f, nullptr, tag));
    for (size_t k = 0; k < f.size(); k++)
        energies[tag[k]] = f[k].w;
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_get_virials(hb2_pair* p, double* virials6)
    {
    if (!p || !virials6)
        return HB2_ERR_INVALID;
    if (p->sys->dom.active)
        {
        set_error("the system is domain decomposed: use hb2_pair_get_local and gather by tag");
        return HB2_ERR_INVALID;
        }
    if (!p->sys->compute_virial)
        {
        set_error("virials are not computed unless the pressure flag is set "
                  "(always_compute_pressure)");
        return HB2_ERR_INVALID;
        }
    std::vector<float4> f;
    std::vector<float> v;
    std::vector<uint32_t> tag;
    HB2_TRY(pair_download(p, f, &v, tag));
    const size_t N = f.size();
    for (size_t k = 0; k < N; k++)
        for (int q = 0; q < 6; q++)
            virials6[6 * (size_t)tag[k] + q] = v[q * (size_t)p->cap_N + k];
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_get_local(hb2_pair* p, uint32_t* tag_out, double* forces3,
                                         double* energies, double* virials6)
    {
    if (!p)
        return HB2_ERR_INVALID;
    if (virials6 && !p->sys->compute_virial)
        {
        set_error("virials are not computed unless the pressure flag is set "
                  "(always_compute_pressure)");
        return HB2_ERR_INVALID;
        }
    std::vector<float4> f;
    std::vector<float> v;
    std::vector<uint32_t> tag;
    HB2_TRY(pair_download(p, f, virials6 ? &v : nullptr, tag));
    for (size_t k = 0; k < f.size(); k++)
        {
        if (tag_out)
            tag_out[k] = tag[k];
        if (forces3)
            {
            forces3[3 * k] = f[k].x;
            forces3[3 * k + 1] = f[k].y;
            forces3[3 * k + 2] = f[k].z;
            }
        if (energies)
            energies[k] = f[k].w;
        if (virials6)
            for (int q = 0; q < 6; q++)
                virials6[6 * k + q] = v[q * (size_t)p->cap_N + k];
        }
    return HB2_OK;
    }

static hb2_status sums7(hb2_pair* p, double out[7])
    {
    hb2_system* s = p->sys;
    HB2_CUDA(cudaSetDevice(s->device));
    HB2_TRY(pair_ensure(p));
    HB2_TRY(launch_reduce_energy(p->d_force, p->virial_valid ? p->d_virial : nullptr, p->cap_N, s->N,
                                 s->d_partial, s->d_ticket + 1, s->d_sums + 1, s->stream));
    if (s->dom.active) // collective: every rank asks for the same quantity (as with MPI in the reference)
        HB2_TRY(domain_allreduce_sum_f64(s, s->d_sums + 1, 7));
    HB2_CUDA(cudaMemcpyAsync(out, s->d_sums + 1, 7 * sizeof(double), cudaMemcpyDeviceToHost,
                             s->stream));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    return HB2_OK;
    }

// n_per_type_in: particle counts per type already known (the ConstantPressure loop evaluates the correction at
// every step's volume without downloading the positions again); n_per_type_out: hand the counts back
static void tail_correction(hb2_pair* p, double& e, double& w, const std::vector<uint32_t>* n_per_type_in = nullptr,
                            std::vector<uint32_t>* n_per_type_out = nullptr)
    {
    // hoomd/md/PotentialPair.h:283-337 with hoomd/md/EvaluatorPairLJ.h:213-248
    e = 0.0;
    w = 0.0;
    if (!p->tail_correction || p->ev != HB2_EVAL_LJ)
        return;
    hb2_system* s = p->sys;
    const uint32_t nt = s->ntypes;
    std::vector<uint32_t> counted;
    if (!n_per_type_in)
        {
        std::vector<float4> hp(s->N);
        cudaMemcpyAsync(hp.data(), s->pos[s->cur], sizeof(float4) * s->N, cudaMemcpyDeviceToHost, s->stream);
        cudaStreamSynchronize(s->stream);
        counted.assign(nt, 0);
        for (auto& q : hp)
            {
            uint32_t t;
            memcpy(&t, &q.w, 4);
            counted[t]++;
            }
        if (s->dom.active)
            {
            // the densities are the global ones: sum the counts over the domains (collective, like every
            // getter of a decomposed system; counts travel as doubles, exact below 2^53)
            std::vector<double> c(nt);
            for (uint32_t q = 0; q < nt; q++)
                c[q] = (double)counted[q];
            double* d_c = nullptr;
            if (dev_alloc(&d_c, nt) == HB2_OK)
                {
                cudaMemcpyAsync(d_c, c.data(), sizeof(double) * nt, cudaMemcpyHostToDevice, s->stream);
                domain_allreduce_sum_f64(s, d_c, nt);
                cudaMemcpyAsync(c.data(), d_c, sizeof(double) * nt, cudaMemcpyDeviceToHost, s->stream);
                cudaStreamSynchronize(s->stream);
                dev_free(d_c);
                for (uint32_t q = 0; q < nt; q++)
                    counted[q] = (uint32_t)c[q];
                }
            }
        if (n_per_type_out)
            *n_per_type_out = counted;
        }
    const std::vector<uint32_t>& n_per_type = n_per_type_in ? *n_per_type_in : counted;
    const double volume = s->box.L[0] * s->box.L[1] * s->box.L[2];
    const double pi = 3.14159265358979323846;
    for (uint32_t ti = 0; ti < nt; ti++)
        {
        const double rho_i = n_per_type[ti] / volume;
        for (uint32_t tj = 0; tj < nt; tj++)
            {
            const double rho_j = n_per_type[tj] / volume;
            const size_t tp = (size_t)tj * nt + ti;
            const double eps = p->user[HB2_PARAM_FLOATS * tp], sig = p->user[HB2_PARAM_FLOATS * tp + 1];
            const double s6 = sig * sig * sig * sig * sig * sig;
            const double lj1 = 4.0 * eps * s6 * s6, lj2 = 4.0 * eps * s6;
            const double rcutsq = p->r_cut[tp] * p->r_cut[tp];
            double ie = 0.0, ip = 0.0;
            if (rcutsq != 0)
                {
                const double rc3inv = 1.0 / pow(rcutsq, 1.5);
                const double rc9inv = rc3inv * rc3inv * rc3inv;
                ie = lj1 / 9.0 * rc9inv - lj2 / 3.0 * rc3inv;
                ip = lj1 * 4.0 / 3.0 * rc9inv - lj2 * 2.0 * rc3inv;
                }
            e += 2.0 * n_per_type[ti] * pi * rho_j * ie;
            const double dp = 4.0 / 6.0 * rho_i * rho_j * pi * ip;
            w += 3 * volume * dp / 3.0;
            }
        }
    }

extern "C" hb2_status hb2_pair_get_energy(hb2_pair* p, double* energy)
    {
    if (!p || !energy)
        return HB2_ERR_INVALID;
    double v[7];
    HB2_TRY(sums7(p, v));
    double e, w;
    tail_correction(p, e, w);
    *energy = v[0] + e; // hoomd/ForceCompute.cc:213-234: sum + external energy
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_set_tail_correction(hb2_pair* p, int on)
    {
    if (!p)
        return HB2_ERR_INVALID;
    if (on && p->ev != HB2_EVAL_LJ)
        {
        set_error("tail_correction is only available for the LJ evaluator");
        return HB2_ERR_INVALID;
        }
    // (with a domain decomposition the densities come from the all-reduced type counts: tail_correction())
    p->tail_correction = on != 0;
    return HB2_OK;
    }

extern "C" hb2_status hb2_pair_get_additional(hb2_pair* p, double* energy, double virial6[6])
    {
    if (!p)
        return HB2_ERR_INVALID;
    HB2_CUDA(cudaSetDevice(p->sys->device));
    double e, w;
    tail_correction(p, e, w);
    if (energy)
        *energy = e;
    if (virial6)
        {
        for (int q = 0; q < 6; q++)
            virial6[q] = 0.0;
        // delta_virial = D V delta_P / 3 goes on EACH diagonal element (hoomd/md/PotentialPair.h:318-337),
        // so that P = (2 K / 3 + (W_xx + W_yy + W_zz) / 3) / V moves by delta_P
        virial6[0] = virial6[3] = virial6[5] = w;
        }
    return HB2_OK;
    }

// =================================================================================================
// integrator
// =================================================================================================
// ---------------------------------------------------------------------------------- bonds ----
static hb2_status ensure_domain_rtag(hb2_system* s);

extern "C" hb2_status hb2_system_set_bonds(hb2_system* s, uint32_t n_bonds, const uint32_t* group,
                                           const uint32_t* typeid_, uint32_t n_bond_types)
    {
    if (!s || (n_bonds && (!group || !typeid_)))
        return HB2_ERR_INVALID;
    // With a domain decomposition every rank holds the whole bond table (keyed by global tag) and an
    // rtag table over the global tags that the rebuild refills for its owned + ghost particles: bond
    // data does not have to migrate, and a bond is complete when its partner is owned or a ghost
    // (bond length < r_cut + buffer, as in the reference).
    if (!s->pos[0] || (!s->dom.active && !s->rtag))
        {
        set_error("hb2_system_set_bonds: set the particles first");
        return HB2_ERR_INVALID;
        }
    if (s->dom.active && s->dom.N_global >= 0xffffffffull)
        return HB2_ERR_INVALID;
    const uint32_t N = s->dom.active ? (uint32_t)s->dom.N_global : s->N;
    HB2_TRY(ensure_domain_rtag(s));
    std::vector<uint32_t> head((size_t)N + 1, 0u), partner(2 * (size_t)n_bonds), type(2 * (size_t)n_bonds);
    for (uint32_t k = 0; k < n_bonds; k++)
        {
        const uint32_t a = group[2 * k], b = group[2 * k + 1];
        if (a >= N || b >= N || a == b || typeid_[k] >= n_bond_types)
            {
            set_error("hb2_system_set_bonds: bond " + std::to_string(k)
                      + " names a particle tag or a bond type out of range");
            return HB2_ERR_INVALID;
            }
        head[a + 1]++;
        head[b + 1]++;
        }
    for (uint32_t t = 0; t < N; t++)
        head[t + 1] += head[t];
    std::vector<uint32_t> fill(N, 0u);
    for (uint32_t k = 0; k < n_bonds; k++)
        {
        const uint32_t a = group[2 * k], b = group[2 * k + 1];
        partner[head[a] + fill[a]] = b;
        type[head[a] + fill[a]++] = typeid_[k];
        partner[head[b] + fill[b]] = a;
        type[head[b] + fill[b]++] = typeid_[k];
        }
    HB2_CUDA(cudaSetDevice(s->device));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    dev_free(s->d_bond_head);
    dev_free(s->d_bond_partner);
    dev_free(s->d_bond_type);
    s->d_bond_head = s->d_bond_partner = s->d_bond_type = nullptr;
    HB2_TRY(dev_alloc(&s->d_bond_head, head.size()));
    HB2_TRY(dev_alloc(&s->d_bond_partner, std::max<size_t>(1, partner.size())));
    HB2_TRY(dev_alloc(&s->d_bond_type, std::max<size_t>(1, type.size())));
    HB2_CUDA(cudaMemcpy(s->d_bond_head, head.data(), sizeof(uint32_t) * head.size(), cudaMemcpyHostToDevice));
    if (n_bonds)
        {
        HB2_CUDA(cudaMemcpy(s->d_bond_partner, partner.data(), sizeof(uint32_t) * partner.size(),
                            cudaMemcpyHostToDevice));
        HB2_CUDA(cudaMemcpy(s->d_bond_type, type.data(), sizeof(uint32_t) * type.size(),
                            cudaMemcpyHostToDevice));
        }
    s->n_bonds = n_bonds;
    s->n_bond_types = n_bond_types;
    s->bond_epoch++;
    for (auto* nl : s->nlists)
        nl->force_update = true; // the exclusions changed
    for (auto* pr : s->pairs)
        pr->computed = false;
    return HB2_OK;
    }

static hb2_status ensure_domain_rtag(hb2_system* s)
    {
    if (s->dom.active && !s->rtag)
        {
        HB2_CUDA(cudaSetDevice(s->device));
        HB2_TRY(dev_alloc(&s->rtag, (size_t)s->dom.N_global));
        HB2_CUDA(cudaMemset(s->rtag, 0xff, sizeof(uint32_t) * (size_t)s->dom.N_global));
        for (auto* nl : s->nlists)
            nl->force_update = true; // the next rebuild fills the table
        }
    return HB2_OK;
    }

extern "C" hb2_status hb2_system_set_angles(hb2_system* s, uint32_t n_angles, const uint32_t* group,
                                            const uint32_t* typeid_, uint32_t n_angle_types)
    {
    if (!s || (n_angles && (!group || !typeid_)))
        return HB2_ERR_INVALID;
    if (!s->pos[0] || (!s->dom.active && !s->rtag))
        {
        set_error("hb2_system_set_angles: set the particles first");
        return HB2_ERR_INVALID;
        }
    if (n_angle_types >= (1u << 30))
        return HB2_ERR_INVALID;
    const uint32_t N = s->dom.active ? (uint32_t)s->dom.N_global : s->N;
    HB2_TRY(ensure_domain_rtag(s));
    std::vector<uint32_t> head((size_t)N + 1, 0u);
    for (uint32_t k = 0; k < n_angles; k++)
        {
        const uint32_t* g = group + 3 * (size_t)k;
        if (g[0] >= N || g[1] >= N || g[2] >= N || g[0] == g[1] || g[1] == g[2] || g[0] == g[2]
            || typeid_[k] >= n_angle_types)
            {
            set_error("hb2_system_set_angles: angle " + std::to_string(k)
                      + " names a particle tag or an angle type out of range");
            return HB2_ERR_INVALID;
            }
        for (int m = 0; m < 3; m++)
            head[g[m] + 1]++;
        }
    for (uint32_t t = 0; t < N; t++)
        head[t + 1] += head[t];
    std::vector<uint4> members(std::max<size_t>(1, 3 * (size_t)n_angles));
    std::vector<uint32_t> fill(N, 0u);
    for (uint32_t k = 0; k < n_angles; k++)
        {
        const uint32_t* g = group + 3 * (size_t)k;
        for (uint32_t m = 0; m < 3; m++)
            members[head[g[m]] + fill[g[m]]++] = make_uint4(g[0], g[1], g[2], typeid_[k] | (m << 30));
        }
    HB2_CUDA(cudaSetDevice(s->device));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    dev_free(s->d_angle_head);
    dev_free(s->d_angle_members);
    s->d_angle_head = nullptr;
    s->d_angle_members = nullptr;
    HB2_TRY(dev_alloc(&s->d_angle_head, head.size()));
    HB2_TRY(dev_alloc(&s->d_angle_members, members.size()));
    HB2_CUDA(cudaMemcpy(s->d_angle_head, head.data(), sizeof(uint32_t) * head.size(), cudaMemcpyHostToDevice));
    HB2_CUDA(cudaMemcpy(s->d_angle_members, members.data(), sizeof(uint4) * members.size(),
                        cudaMemcpyHostToDevice));
    s->n_angles = n_angles;
    s->n_angle_types = n_angle_types;
    return HB2_OK;
    }

extern "C" hb2_status hb2_system_set_dihedrals(hb2_system* s, uint32_t n_dihedrals, const uint32_t* group,
                                               const uint32_t* typeid_, uint32_t n_dihedral_types)
    {
    if (!s || (n_dihedrals && (!group || !typeid_)))
        return HB2_ERR_INVALID;
    if (!s->pos[0] || (!s->dom.active && !s->rtag))
        {
        set_error("hb2_system_set_dihedrals: set the particles first");
        return HB2_ERR_INVALID;
        }
    if (n_dihedral_types >= (1u << 30))
        return HB2_ERR_INVALID;
    const uint32_t N = s->dom.active ? (uint32_t)s->dom.N_global : s->N;
    HB2_TRY(ensure_domain_rtag(s));
    std::vector<uint32_t> head((size_t)N + 1, 0u);
    for (uint32_t k = 0; k < n_dihedrals; k++)
        {
        const uint32_t* g = group + 4 * (size_t)k;
        bool ok = typeid_[k] < n_dihedral_types;
        for (int m = 0; m < 4 && ok; m++)
            {
            ok = g[m] < N;
            for (int q = 0; q < m && ok; q++)
                ok = g[q] != g[m];
            }
        if (!ok)
            {
            set_error("hb2_system_set_dihedrals: dihedral " + std::to_string(k)
                      + " names a particle tag or a dihedral type out of range");
            return HB2_ERR_INVALID;
            }
        for (int m = 0; m < 4; m++)
            head[g[m] + 1]++;
        }
    for (uint32_t t = 0; t < N; t++)
        head[t + 1] += head[t];
    std::vector<uint4> members(std::max<size_t>(1, 4 * (size_t)n_dihedrals));
    std::vector<uint32_t> tr(members.size(), 0u), fill(N, 0u);
    for (uint32_t k = 0; k < n_dihedrals; k++)
        {
        const uint32_t* g = group + 4 * (size_t)k;
        for (uint32_t m = 0; m < 4; m++)
            {
            const uint32_t at = head[g[m]] + fill[g[m]]++;
            members[at] = make_uint4(g[0], g[1], g[2], g[3]);
            tr[at] = typeid_[k] | (m << 30);
            }
        }
    HB2_CUDA(cudaSetDevice(s->device));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    dev_free(s->d_dihedral_head);
    dev_free(s->d_dihedral_members);
    dev_free(s->d_dihedral_tr);
    s->d_dihedral_head = nullptr;
    s->d_dihedral_members = nullptr;
    s->d_dihedral_tr = nullptr;
    HB2_TRY(dev_alloc(&s->d_dihedral_head, head.size()));
    HB2_TRY(dev_alloc(&s->d_dihedral_members, members.size()));
    HB2_TRY(dev_alloc(&s->d_dihedral_tr, tr.size()));
    HB2_CUDA(cudaMemcpy(s->d_dihedral_head, head.data(), sizeof(uint32_t) * head.size(), cudaMemcpyHostToDevice));
    HB2_CUDA(cudaMemcpy(s->d_dihedral_members, members.data(), sizeof(uint4) * members.size(),
                        cudaMemcpyHostToDevice));
    HB2_CUDA(cudaMemcpy(s->d_dihedral_tr, tr.data(), sizeof(uint32_t) * tr.size(), cudaMemcpyHostToDevice));
    s->n_dihedrals = n_dihedrals;
    s->n_dihedral_types = n_dihedral_types;
    return HB2_OK;
    }

extern "C" hb2_status hb2_nlist_set_exclusion_pairs(hb2_nlist* nl, uint32_t n_pairs, const uint32_t* pairs)
    {
    if (!nl || (n_pairs && !pairs))
        return HB2_ERR_INVALID;
    hb2_system* s = nl->sys;
    if (!s->pos[0] || (!s->dom.active && !s->rtag))
        {
        set_error("hb2_nlist_set_exclusion_pairs: set the particles first");
        return HB2_ERR_INVALID;
        }
    const uint32_t N = s->dom.active ? (uint32_t)s->dom.N_global : s->N;
    if (n_pairs)
        HB2_TRY(ensure_domain_rtag(s));
    std::vector<uint32_t> head((size_t)N + 1, 0u), partner(std::max<size_t>(1, 2 * (size_t)n_pairs));
    for (uint32_t k = 0; k < n_pairs; k++)
        {
        const uint32_t a = pairs[2 * k], b = pairs[2 * k + 1];
        if (a >= N || b >= N || a == b)
            {
            set_error("hb2_nlist_set_exclusion_pairs: pair " + std::to_string(k) + " is out of range");
            return HB2_ERR_INVALID;
            }
        head[a + 1]++;
        head[b + 1]++;
        }
    for (uint32_t t = 0; t < N; t++)
        head[t + 1] += head[t];
    std::vector<uint32_t> fill(N, 0u);
    for (uint32_t k = 0; k < n_pairs; k++)
        {
        const uint32_t a = pairs[2 * k], b = pairs[2 * k + 1];
        partner[head[a] + fill[a]++] = b;
        partner[head[b] + fill[b]++] = a;
        }
    HB2_CUDA(cudaSetDevice(s->device));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    dev_free(nl->d_ex_head);
    dev_free(nl->d_ex_partner);
    nl->d_ex_head = nl->d_ex_partner = nullptr;
    nl->n_ex_pairs = 0;
    if (n_pairs)
        {
        HB2_TRY(dev_alloc(&nl->d_ex_head, head.size()));
        HB2_TRY(dev_alloc(&nl->d_ex_partner, partner.size()));
        HB2_CUDA(cudaMemcpy(nl->d_ex_head, head.data(), sizeof(uint32_t) * head.size(), cudaMemcpyHostToDevice));
        HB2_CUDA(cudaMemcpy(nl->d_ex_partner, partner.data(), sizeof(uint32_t) * partner.size(),
                            cudaMemcpyHostToDevice));
        nl->n_ex_pairs = n_pairs;
        }
    nl->force_update = true;
    for (auto* pr : nl->consumers)
        pr->computed = false;
    return HB2_OK;
    }

extern "C" hb2_status hb2_nlist_set_exclusions(hb2_nlist* nl, int exclude_bonds)
    {
    if (!nl)
        return HB2_ERR_INVALID;
    nl->exclude_bonds = exclude_bonds != 0;
    nl->force_update = true;
    for (auto* pr : nl->consumers)
        pr->computed = false;
    return HB2_OK;
    }

extern "C" hb2_status hb2_bond_create(hb2_bond** out, hb2_system* s, hb2_bond_kind kind)
    {
    if (!out || !s || (unsigned)kind > (unsigned)HB2_DIHEDRAL_OPLS)
        return HB2_ERR_INVALID;
    hb2_bond* b = new hb2_bond();
    b->sys = s;
    b->kind = kind;
    HB2_CUDA(cudaSetDevice(s->device));
    hb2_status st = dev_alloc(&b->d_err, 2);
    if (st != HB2_OK)
        {
        delete b;
        return st;
        }
    cudaMemset(b->d_err, 0, 2 * sizeof(uint32_t));
    *out = b;
    return HB2_OK;
    }

extern "C" void hb2_bond_destroy(hb2_bond* b)
    {
    if (!b)
        return;
    cudaSetDevice(b->sys->device);
    cudaStreamSynchronize(b->sys->stream);
    dev_free(b->d_params);
    dev_free(b->d_force);
    dev_free(b->d_virial);
    dev_free(b->d_err);
    delete b;
    }

extern "C" hb2_status hb2_bond_set_params(hb2_bond* b, uint32_t type, const double* user, uint32_t n_user)
    {
    const uint32_t need = !b ? 2u : (b->kind == HB2_BOND_FENEWCA ? 5u : ((b->kind == HB2_DIHEDRAL_PERIODIC || b->kind == HB2_DIHEDRAL_OPLS) ? 4u : 2u));
    if (!b || !user || n_user < need)
        {
        set_error("hb2_bond_set_params: harmonic takes {k, r0}, FENEWCA {k, r0, epsilon, sigma, delta}");
        return HB2_ERR_INVALID;
        }
    if (type >= b->have.size())
        {
        b->packed.resize(HB2_BOND_PARAM_FLOATS * (size_t)(type + 1), 0.f);
        b->have.resize(type + 1, 0);
        }
    float* p = &b->packed[HB2_BOND_PARAM_FLOATS * (size_t)type];
    p[0] = (float)user[0];
    p[1] = (float)user[1];
    if (b->kind == HB2_BOND_FENEWCA)
        {
        // hoomd/md/EvaluatorBondFENE.h:44-53 keeps sigma^6 and 4 epsilon as Scalars; lj1, lj2 from them
        const float sigma = (float)user[3];
        const float sigma_6 = sigma * sigma * sigma * sigma * sigma * sigma;
        const float epsilon_x_4 = 4.0f * (float)user[2];
        p[2] = epsilon_x_4 * sigma_6 * sigma_6;
        p[3] = epsilon_x_4 * sigma_6;
        p[4] = (float)user[4];
        }
    if (b->kind == HB2_DIHEDRAL_PERIODIC)
        {
        p[2] = (float)user[2]; // {k, d (sign), n (multiplicity), phi0}
        p[3] = (float)user[3];
        }
    if (b->kind == HB2_DIHEDRAL_OPLS)
        for (int q = 0; q < 4; q++) // OPLSDihedralForceCompute::setParams stores k_i / 2
            p[q] = (float)(user[q] / 2.0);
    b->have[type] = 1;
    b->params_dirty = true;
    return HB2_OK;
    }

hb2_status hb2_bond::upload_params()
    {
    const bool is_angle = kind == HB2_ANGLE_HARMONIC || kind == HB2_ANGLE_COSINESQ;
    const uint32_t nt = (kind == HB2_DIHEDRAL_PERIODIC || kind == HB2_DIHEDRAL_OPLS) ? sys->n_dihedral_types
                                                       : (is_angle ? sys->n_angle_types : sys->n_bond_types);
    for (uint32_t t = 0; t < nt; t++)
        if (t >= have.size() || !have[t])
            {
            set_error(std::string(is_angle ? "angle" : "bond") + ": parameters of type " + std::to_string(t)
                      + " are not set");
            return HB2_ERR_INVALID;
            }
    std::vector<float> h(HB2_BOND_PARAM_FLOATS * (size_t)std::max<uint32_t>(1, nt), 0.f);
    std::copy(packed.begin(), packed.begin() + std::min(packed.size(), h.size()), h.begin());
    dev_free(d_params);
    d_params = nullptr;
    HB2_TRY(dev_alloc(&d_params, h.size()));
    HB2_CUDA(cudaMemcpyAsync(d_params, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice, sys->stream));
    HB2_CUDA(cudaStreamSynchronize(sys->stream));
    params_dirty = false;
    return HB2_OK;
    }

hb2_status hb2_bond::launch(float4* force, float* virial, size_t virial_pitch, bool accumulate,
                            const uint32_t* skip_flag)
    {
    hb2_system* s = sys;
    if (params_dirty)
        HB2_TRY(upload_params());
    const bool is_angle = kind == HB2_ANGLE_HARMONIC || kind == HB2_ANGLE_COSINESQ;
    const bool is_dihedral = kind == HB2_DIHEDRAL_PERIODIC || kind == HB2_DIHEDRAL_OPLS;
    if ((is_dihedral ? s->n_dihedrals : (is_angle ? s->n_angles : s->n_bonds)) == 0 && accumulate)
        return HB2_OK;
    if (is_dihedral && !s->d_dihedral_head)
        HB2_TRY(hb2_system_set_dihedrals(s, 0, nullptr, nullptr, 0));
    if (!is_angle && !is_dihedral && !s->d_bond_head) // no table was ever set: an empty one
        HB2_TRY(hb2_system_set_bonds(s, 0, nullptr, nullptr, 0));
    if (is_angle && !s->d_angle_head)
        HB2_TRY(hb2_system_set_angles(s, 0, nullptr, nullptr, 0));
    BondArgs a;
    a.pos = s->pos[s->cur];
    a.tag = s->tag[s->cur];
    a.rtag = s->rtag;
    a.N = s->N;
    a.Ntot = s->N + s->Ng;
    a.head = is_dihedral ? s->d_dihedral_head : (is_angle ? s->d_angle_head : s->d_bond_head);
    a.partner = s->d_bond_partner;
    a.btype = s->d_bond_type;
    a.params = d_params;
    a.force = force;
    a.virial = s->compute_virial ? virial : nullptr;
    a.virial_pitch = virial_pitch;
    a.skip_flag = skip_flag;
    a.err = d_err;
    if (is_dihedral)
        return launch_dihedral_forces(kind, a, s->d_dihedral_members, s->d_dihedral_tr, accumulate, s->box,
                                      s->stream);
    if (is_angle)
        return launch_angle_forces(kind, a, s->d_angle_members, accumulate, s->box, s->stream);
    return launch_bond_forces(kind, a, accumulate, s->box, s->stream);
    }

hb2_status hb2_bond::compute_own()
    {
    hb2_system* s = sys;
    HB2_CUDA(cudaSetDevice(s->device));
    if (s->cap > cap_N)
        {
        dev_free(d_force);
        dev_free(d_virial);
        d_force = nullptr;
        d_virial = nullptr;
        HB2_TRY(dev_alloc(&d_force, s->cap));
        HB2_TRY(dev_alloc(&d_virial, 6 * (size_t)s->cap));
        cap_N = s->cap;
        }
    HB2_TRY(launch(d_force, d_virial, cap_N, false, nullptr));
    return check_complete();
    }

hb2_status hb2_bond::check_complete()
    {
    uint32_t e[2] = {0, 0};
    HB2_CUDA(cudaMemcpyAsync(e, d_err, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, sys->stream));
    HB2_CUDA(cudaStreamSynchronize(sys->stream));
    if (e[0] || e[1])
        {
        HB2_CUDA(cudaMemsetAsync(d_err, 0, 2 * sizeof(uint32_t), sys->stream));
        if (e[0])
            set_error("Error: a bond with particle " + std::to_string(e[0] - 1) + " is incomplete.");
        else
            set_error("bond.fene: bond out of bounds (particle " + std::to_string(e[1] - 1) + ")");
        return HB2_ERR_INVALID;
        }
    return HB2_OK;
    }

// per-particle results in tag order: forces N x 3, energies N, virials N x 6 (any may be null)
extern "C" hb2_status hb2_bond_get(hb2_bond* b, double* forces3, double* energies, double* virials6,
                                   double* energy_total)
    {
    if (!b)
        return HB2_ERR_INVALID;
    hb2_system* s = b->sys;
    if (s->dom.active)
        {
        set_error("the system is domain decomposed: per-particle bond results are not gathered "
                  "(thermodynamic sums include the bonds)");
        return HB2_ERR_INVALID;
        }
    HB2_TRY(b->compute_own());
    const uint32_t N = s->N;
    std::vector<float4> f(N);
    std::vector<uint32_t> tag(N);
    std::vector<float> vir;
    HB2_CUDA(cudaMemcpyAsync(f.data(), b->d_force, sizeof(float4) * N, cudaMemcpyDeviceToHost, s->stream));
    HB2_CUDA(cudaMemcpyAsync(tag.data(), s->tag[s->cur], sizeof(uint32_t) * N, cudaMemcpyDeviceToHost,
                             s->stream));
    if (virials6)
        {
        if (!s->compute_virial)
            {
            set_error("virials are not computed unless the pressure flag is set "
                      "(Simulation.always_compute_pressure)");
            return HB2_ERR_INVALID;
            }
        vir.resize(6 * (size_t)b->cap_N);
        HB2_CUDA(cudaMemcpyAsync(vir.data(), b->d_virial, sizeof(float) * vir.size(), cudaMemcpyDeviceToHost,
                                 s->stream));
        }
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    double tot = 0.0;
    for (uint32_t i = 0; i < N; i++)
        {
        const uint32_t t = tag[i];
        if (forces3)
            {
            forces3[3 * (size_t)t] = f[i].x;
            forces3[3 * (size_t)t + 1] = f[i].y;
            forces3[3 * (size_t)t + 2] = f[i].z;
            }
        if (energies)
            energies[t] = f[i].w;
        if (virials6)
            for (int q = 0; q < 6; q++)
                virials6[6 * (size_t)t + q] = vir[(size_t)q * b->cap_N + i];
        tot += f[i].w;
        }
    if (energy_total)
        *energy_total = tot;
    return HB2_OK;
    }

// Integrator::computeNetForce (hoomd/Integrator.cc:166-246): the net force is the sum of every force.
// The first pair force's arrays ARE the net-force arrays; every further pair force runs its own pass
// over the shared neighbour list (whose r_cut matrix is the maximum over its consumers,
// hoomd/md/NeighborList.cc:452-529) into its own arrays and is then added in, and the bonded forces
// accumulate directly.
static hb2_status integrator_add_extra(hb2_integrator* it, hb2_pair* pf, uint64_t timestep, bool want_energy,
                                       const uint32_t* skip_flag)
    {
    hb2_system* s = it->sys;
    for (size_t k = 1; k < it->forces.size(); k++)
        {
        hb2_pair* p = it->forces[k];
        if (p->params_dirty)
            HB2_TRY(p->upload_params());
        HB2_TRY(p->launch(timestep, want_energy, skip_flag));
        HB2_TRY(launch_accumulate_force(pf->d_force, s->compute_virial ? pf->d_virial : nullptr, pf->cap_N,
                                        p->d_force, p->d_virial, p->cap_N, s->N, skip_flag, s->stream));
        pf->extra_added = true;
        }
    for (auto* b : it->bonds)
        HB2_TRY(b->launch(pf->d_force, pf->d_virial, pf->cap_N, true, skip_flag));
    if (!it->bonds.empty())
        pf->extra_added = true;
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_add_bond(hb2_integrator* it, hb2_bond* b)
    {
    if (!it || !b || b->sys != it->sys)
        return HB2_ERR_INVALID;
    it->bonds.push_back(b);
    it->prepared = false;
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_create(hb2_integrator** out, hb2_system* s, double dt)
    {
    if (!out || !s)
        return HB2_ERR_INVALID;
    HB2_CUDA(cudaSetDevice(s->device));
    hb2_integrator* it = new hb2_integrator();
    it->sys = s;
    it->dt = dt;
    cudaEventCreate(&it->ev0);
    cudaEventCreate(&it->ev1);
    *out = it;
    return HB2_OK;
    }

extern "C" void hb2_integrator_destroy(hb2_integrator* it)
    {
    if (!it)
        return;
    cudaSetDevice(it->sys->device);
    cudaEventDestroy(it->ev0);
    cudaEventDestroy(it->ev1);
    for (auto e : it->pool)
        cudaEventDestroy(e);
    dev_free(it->d_gamma_coeff);
    dev_free(it->d_total);
    dev_free(it->d_gen);
    dev_free(it->d_member);
    if (it->npt.h_thermo)
        cudaFreeHost(it->npt.h_thermo);
    delete it;
    }

extern "C" hb2_status hb2_integrator_add_force(hb2_integrator* it, hb2_pair* p)
    {
    if (!it || !p || p->sys != it->sys)
        return HB2_ERR_INVALID;
    if (!it->forces.empty() && it->forces[0]->nl != p->nl)
        {
        set_error("integrator: the pair forces of one integrator share one neighbour list on this path "
                  "(its r_cut matrix is the maximum over them, hoomd/md/NeighborList.cc:452-529)");
        return HB2_ERR_INVALID;
        }
    for (auto* q : it->forces)
        if (q == p)
            {
            set_error("integrator: this force is already attached");
            return HB2_ERR_INVALID;
            }
    it->forces.push_back(p);
    it->prepared = false;
    return HB2_OK;
    }

// System::run resets the per-run statistics of every compute once per run; a caller that splits one run
// into several hb2_integrator_run segments (writers between them) switches the reset off for all but the
// first segment.
extern "C" hb2_status hb2_integrator_set_reset_stats(hb2_integrator* it, int on)
    {
    if (!it)
        return HB2_ERR_INVALID;
    it->reset_stats = on != 0;
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_set_constant_volume(hb2_integrator* it, int thermostat,
                                                         double kT, double tau)
    {
    if (!it || (thermostat != 0 && thermostat != 1))
        return HB2_ERR_INVALID;
    it->thermostat = thermostat;
    it->kT = kT;
    it->tau = tau;
    it->npt.on = false;
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_set_filter_tags(hb2_integrator* it, const uint32_t* tags, uint64_t n)
    {
    if (!it || (n && !tags))
        return HB2_ERR_INVALID;
    hb2_system* s = it->sys;
    HB2_CUDA(cudaSetDevice(s->device));
    const uint64_t Ng = s->dom.active ? s->dom.N_global : s->N;
    dev_free(it->d_member);
    it->d_member = nullptr;
    it->n_member = Ng;
    it->prepared = false;
    it->total_valid = false;
    if (!tags && n == 0)
        return HB2_OK; // filter.All
    std::vector<uint8_t> m((size_t)Ng, 0);
    uint64_t cnt = 0;
    for (uint64_t k = 0; k < n; k++)
        {
        if (tags[k] >= Ng)
            {
            set_error("hb2_integrator_set_filter_tags: tag out of range");
            return HB2_ERR_INVALID;
            }
        cnt += m[tags[k]] ? 0 : 1;
        m[tags[k]] = 1;
        }
    it->n_member = cnt;
    if (cnt == Ng)
        return HB2_OK; // every particle: no mask
    HB2_TRY(dev_alloc(&it->d_member, (size_t)std::max<uint64_t>(Ng, 1)));
    HB2_CUDA(cudaMemcpyAsync(it->d_member, m.data(), (size_t)Ng, cudaMemcpyHostToDevice, s->stream));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_kinetic_of_tags(hb2_integrator* it, const uint32_t* tags, uint64_t n,
                                                     double* kinetic)
    {
    if (!it || !kinetic || (n && !tags))
        return HB2_ERR_INVALID;
    hb2_system* s = it->sys;
    HB2_CUDA(cudaSetDevice(s->device));
    const uint64_t Ng = s->dom.active ? s->dom.N_global : s->N;
    uint8_t* d_m = nullptr;
    if (tags)
        {
        std::vector<uint8_t> m((size_t)Ng, 0);
        for (uint64_t k = 0; k < n; k++)
            {
            if (tags[k] >= Ng)
                {
                set_error("hb2_integrator_kinetic_of_tags: tag out of range");
                return HB2_ERR_INVALID;
                }
            m[tags[k]] = 1;
            }
        HB2_TRY(dev_alloc(&d_m, (size_t)std::max<uint64_t>(Ng, 1)));
        HB2_CUDA(cudaMemcpyAsync(d_m, m.data(), (size_t)Ng, cudaMemcpyHostToDevice, s->stream));
        HB2_CUDA(cudaStreamSynchronize(s->stream)); // m goes out of scope
        }
    double ke = 0.0;
    hb2_status st = launch_kinetic(s->vel[s->cur], s->N, s->d_partial, s->d_ticket + 2, s->d_sums + 9, s->stream,
                                   GroupMask {d_m, s->tag[s->cur]});
    if (st == HB2_OK && s->dom.active)
        st = domain_allreduce_sum_f64(s, s->d_sums + 9, 1);
    if (st == HB2_OK)
        {
        cudaMemcpyAsync(&ke, s->d_sums + 9, sizeof(double), cudaMemcpyDeviceToHost, s->stream);
        cudaStreamSynchronize(s->stream);
        }
    dev_free(d_m);
    if (st != HB2_OK)
        return st;
    *kinetic = ke;
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_set_langevin(hb2_integrator* it, double kT, const double* gamma,
                                                  uint32_t ntypes)
    {
    if (!it || !gamma || ntypes != it->sys->ntypes || !(it->dt > 0.0) || kT < 0.0)
        {
        set_error("hb2_integrator_set_langevin: one gamma per particle type, dt > 0 and kT >= 0 are required");
        return HB2_ERR_INVALID;
        }
    std::vector<float> gc(2 * (size_t)ntypes);
    for (uint32_t t = 0; t < ntypes; t++)
        {
        if (gamma[t] < 0.0)
            {
            set_error("hb2_integrator_set_langevin: gamma must be >= 0");
            return HB2_ERR_INVALID;
            }
        gc[2 * t] = (float)gamma[t];
        gc[2 * t + 1] = (float)std::sqrt(6.0 * gamma[t] * kT / it->dt); // hoomd/md/TwoStepLangevin.cc:269
        }
    HB2_CUDA(cudaSetDevice(it->sys->device));
    if (!it->d_gamma_coeff)
        HB2_TRY(dev_alloc(&it->d_gamma_coeff, gc.size()));
    HB2_CUDA(cudaMemcpyAsync(it->d_gamma_coeff, gc.data(), sizeof(float) * gc.size(), cudaMemcpyHostToDevice,
                             it->sys->stream));
    HB2_CUDA(cudaStreamSynchronize(it->sys->stream));
    it->thermostat = 2;
    it->npt.on = false;
    it->kT = kT;
    it->tau = 0.0;
    return HB2_OK;
    }

// =================================================================================================
// ConstantPressure: the barostat lives on the host in fp64 (six numbers), the particles on the device.
// hoomd/md/TwoStepConstantPressure.cc: advanceBarostat :860-1000, updatePropagator :716-858,
// getRelevantCouplings :128-170, thermalizeBarostatDOF :501-580, getBarostatEnergy :693-714.
// Orthorhombic box degrees of freedom only (x, y, z): the engine has no tilt.
// =================================================================================================
static int npt_relevant_couplings(int couple, int flags)
    {
    // 0 none, 1 xy, 2 xz, 3 yz, 4 xyz
    if (!(flags & 1))
        {
        if (couple == 4)
            couple = 3;
        if (couple == 1 || couple == 2)
            couple = 0;
        }
    if (!(flags & 2))
        {
        if (couple == 4)
            couple = 2;
        if (couple == 3 || couple == 1)
            couple = 0;
        }
    if (!(flags & 4))
        {
        if (couple == 4)
            couple = 1;
        if (couple == 3 || couple == 2)
            couple = 0;
        }
    return couple;
    }

static double npt_W(const hb2_integrator* it)
    {
    const auto& n = it->npt;
    const double T = it->thermostat == 1 ? it->kT : 1.0;
    return (n.ndof + 3.0) / 3.0 * T * n.tauS * n.tauS;
    }

static void npt_advance_barostat(hb2_integrator* it, uint64_t timestep)
    {
    auto& n = it->npt;
    hb2_system* s = it->sys;
    const double dt = it->dt;
    const double V = s->box.L[0] * s->box.L[1] * s->box.L[2];
    const double T = it->thermostat == 1 ? it->kT : 1.0;
    const double W = npt_W(it);
    double mtk_term = 2.0 * n.ke;
    mtk_term *= 0.5 * dt / n.ndof / W;
    const int couple = npt_relevant_couplings(n.couple, n.flags);
    const double Pxx = n.P[0], Pyy = n.P[3], Pzz = n.P[5];
    double Pd[3] = {0, 0, 0}, Rd[3] = {0, 0, 0};
    double z[3];
    rng_normals_host(46 /* RNGIdentifier::ConstantPressure */, timestep, s->seed, 0u, 3, z);
    switch (couple)
        {
    case 0:
        Pd[0] = Pxx;
        Rd[0] = z[0];
        Pd[1] = Pyy;
        Rd[1] = z[1];
        Pd[2] = Pzz;
        Rd[2] = z[2];
        break;
    case 1:
        Pd[0] = Pd[1] = 0.5 * (Pxx + Pyy);
        Rd[0] = Rd[1] = z[0];
        Pd[2] = Pzz;
        Rd[2] = z[1];
        break;
    case 2:
        Pd[0] = Pd[2] = 0.5 * (Pxx + Pzz);
        Rd[0] = Rd[2] = z[0];
        Pd[1] = Pyy;
        Rd[1] = z[1];
        break;
    case 3:
        Pd[0] = Pxx;
        Rd[0] = z[0];
        Pd[1] = Pd[2] = 0.5 * (Pyy + Pzz);
        Rd[1] = Rd[2] = z[1];
        break;
    default:
        Pd[0] = Pd[1] = Pd[2] = (1.0 / 3.0) * (Pxx + Pyy + Pzz);
        Rd[0] = Rd[1] = Rd[2] = z[0];
        break;
        }
    const double e = std::exp(-n.gamma * dt / 2.0);
    const double coeff = std::sqrt(T * (1.0 - e * e) / W);
    const int nu_idx[3] = {0, 3, 5}; // nu_xx, nu_yy, nu_zz
    for (int d = 0; d < 3; d++)
        if (n.flags & (1 << d))
            {
            double& nu = n.nu[nu_idx[d]];
            nu = nu * e + coeff * Rd[d];
            nu += 0.5 * dt * V / W * (Pd[d] - n.S[d]) + mtk_term;
            }
    }

static void npt_update_propagator(hb2_integrator* it)
    {
    auto& n = it->npt;
    const double dt = it->dt;
    static const double f_coeff[6] = {1.0, 1.0 / 6.0, 1.0 / 120.0, 1.0 / 5040.0, 1.0 / 362880.0, 1.0 / 39916800.0};
    const int nu_idx[3] = {0, 3, 5};
    for (int d = 0; d < 3; d++)
        {
        const double nu = n.nu[nu_idx[d]];
        const double v_fac = -0.25 * nu;
        const double arg_r = 0.5 * nu * dt;
        double f_r = 0.0, term = 1.0;
        for (int i = 0; i < 6; i++)
            {
            f_r += f_coeff[i] * term;
            term = term * arg_r * arg_r;
            }
        n.exp_v[d] = std::exp(2.0 * v_fac * dt);
        n.exp_r[d] = std::exp(nu * dt);
        n.exp_r_int[d] = dt * std::exp(0.5 * nu * dt) * f_r;
        }
    }

// the thermodynamic quantities the barostat reads (ComputeThermo of the method's group, every particle on
// this path): kinetic tensor of the current velocities (k.dt == 0) or left in d_sums + 10 by step two,
// virial sums of the net force plus the forces' external virial; one host read.
static hb2_status npt_thermo(hb2_integrator* it, hb2_pair* pf, bool kinetic_from_velocities)
    {
    hb2_system* s = it->sys;
    auto& n = it->npt;
    cudaStream_t st = s->stream;
    if (kinetic_from_velocities)
        {
        NptFactors k0;
        HB2_TRY(launch_step_two_npt(s->vel[s->cur], pf->d_force, s->N, k0, s->d_partial, s->d_ticket + 2,
                                    s->d_sums + 10, st, GroupMask {it->d_member, s->tag[s->cur]}));
        }
    HB2_TRY(launch_reduce_energy(pf->d_force, pf->d_virial, pf->cap_N, s->N, s->d_partial, s->d_ticket + 1,
                                 s->d_sums + 1, st));
    if (s->dom.active)
        {
        // every rank holds the same barostat: the sums over the owned particles of all domains, in one
        // all-reduce each (identical bits on every rank, so the boxes never drift apart)
        HB2_TRY(domain_allreduce_sum_f64(s, s->d_sums + 1, 7));
        HB2_TRY(domain_allreduce_sum_f64(s, s->d_sums + 10, 6));
        }
    HB2_CUDA(cudaMemcpyAsync(n.h_thermo, s->d_sums + 1, 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
    HB2_CUDA(cudaMemcpyAsync(n.h_thermo + 7, s->d_sums + 10, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
    HB2_CUDA(cudaStreamSynchronize(st));
    const double V = s->box.L[0] * s->box.L[1] * s->box.L[2];
    // external virial: the LJ tail correction at the current volume (hoomd/md/PotentialPair.h:283-337),
    // a function of the volume and the type counts alone, so a run cut into segments sees the same numbers
    double wext = 0.0;
    if (n.tail_valid)
        for (auto* p : it->forces)
            {
            double te, tw;
            tail_correction(p, te, tw, &n.n_per_type);
            wext += tw;
            }
    const double* w = n.h_thermo + 1; // xx xy xz yy yz zz
    const double* kt = n.h_thermo + 7;
    for (int q = 0; q < 6; q++)
        {
        const bool diag = q == 0 || q == 3 || q == 5;
        n.P[q] = (kt[q] + w[q] + (diag ? wext : 0.0)) / V;
        }
    n.ke = 0.5 * (kt[0] + kt[3] + kt[5]);
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_set_constant_pressure(hb2_integrator* it, const double S[6], double tauS,
                                                           int couple, const int box_dof[6], int rescale_all,
                                                           double gamma, int thermostat, double kT, double tau)
    {
    if (!it || !S || !box_dof || couple < 0 || couple > 4 || (thermostat != 0 && thermostat != 1) || !(tauS > 0.0))
        {
        set_error("hb2_integrator_set_constant_pressure: bad arguments");
        return HB2_ERR_INVALID;
        }
    if (box_dof[3] || box_dof[4] || box_dof[5])
        {
        set_error("ConstantPressure: the shear degrees of freedom (xy, xz, yz) need a triclinic box, which is "
                  "outside this path's scope");
        return HB2_ERR_INVALID;
        }
    HB2_CUDA(cudaSetDevice(it->sys->device));
    auto& n = it->npt;
    n.on = true;
    for (int q = 0; q < 6; q++)
        n.S[q] = S[q];
    n.tauS = tauS;
    n.couple = couple;
    n.flags = (box_dof[0] ? 1 : 0) | (box_dof[1] ? 2 : 0) | (box_dof[2] ? 4 : 0);
    n.rescale_all = rescale_all != 0;
    n.gamma = gamma;
    if (!n.h_thermo)
        HB2_CUDA(cudaMallocHost((void**)&n.h_thermo, 16 * sizeof(double)));
    it->thermostat = thermostat;
    it->kT = kT;
    it->tau = tau;
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_get_barostat_dof(hb2_integrator* it, double nu[6])
    {
    if (!it || !nu)
        return HB2_ERR_INVALID;
    for (int q = 0; q < 6; q++)
        nu[q] = it->npt.nu[q];
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_set_barostat_dof(hb2_integrator* it, const double nu[6])
    {
    if (!it || !nu)
        return HB2_ERR_INVALID;
    for (int q = 0; q < 6; q++)
        it->npt.nu[q] = nu[q];
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_thermalize_barostat_dof(hb2_integrator* it, uint64_t timestep)
    {
    if (!it || !it->npt.on)
        return HB2_ERR_INVALID;
    auto& n = it->npt;
    const double sigmasq = 1.0 / ((n.ndof + 3.0) / 3.0 * n.tauS * n.tauS);
    // one draw per enabled degree of freedom in the order xx, xy, xz, yy, yz, zz
    // (RNGIdentifier::TwoStepConstantPressureThermalizeBarostat = 27)
    double z[3];
    int nd = 0;
    for (int d = 0; d < 3; d++)
        nd += (n.flags >> d) & 1;
    rng_normals_host(27, timestep, it->sys->seed, 0u, nd, z);
    const int nu_idx[3] = {0, 3, 5};
    int k = 0;
    for (int d = 0; d < 3; d++)
        if (n.flags & (1 << d))
            n.nu[nu_idx[d]] = std::sqrt(sigmasq) * z[k++];
    switch (npt_relevant_couplings(n.couple, n.flags))
        {
    case 1:
        n.nu[3] = n.nu[0];
        break;
    case 2:
        n.nu[5] = n.nu[0];
        break;
    case 3:
        n.nu[3] = n.nu[5];
        break;
    case 4:
        n.nu[0] = n.nu[5];
        n.nu[3] = n.nu[5];
        break;
    default:
        break;
        }
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_barostat_energy(hb2_integrator* it, double* energy)
    {
    if (!it || !energy || !it->npt.on)
        return HB2_ERR_INVALID;
    const auto& n = it->npt;
    double s2 = 0.0;
    for (int q = 0; q < 6; q++)
        s2 += n.nu[q] * n.nu[q];
    *energy = npt_W(it) * s2 / 2.0;
    return HB2_OK;
    }

extern "C" hb2_status hb2_system_get_box(hb2_system* s, double L[3])
    {
    if (!s || !L || !s->has_box)
        return HB2_ERR_INVALID;
    for (int d = 0; d < 3; d++)
        L[d] = s->box.L[d];
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_run(hb2_integrator* it, uint64_t nsteps)
    {
    if (!it || it->forces.empty())
        {
        set_error("integrator: no force attached");
        return HB2_ERR_INVALID;
        }
    hb2_system* s = it->sys;
    hb2_pair* pf = it->forces[0];
    hb2_nlist* nl = pf->nl;
    HB2_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    bool any_dirty = false;
    for (auto* p : it->forces)
        {
        p->dt = it->dt;
        // DPD reads the velocities of its neighbours, ghosts included
        if (s->dom.active && (p->ev == HB2_EVAL_DPD || p->ev == HB2_EVAL_DPDLJ))
            s->dom.need_vel = true;
        }
    double* d_ke = s->d_sums;
    double* d_rescale = s->d_sums + 8;
    auto& npt = it->npt;
    // the barostat reads the pressure tensor every step: the pressure flag is up for the whole run
    // (TwoStepConstantPressure::getRequestedPDataFlags) and handed back afterwards
    struct VirialFlag
        {
        hb2_system* s;
        bool saved, active;
        ~VirialFlag()
            {
            if (active)
                s->compute_virial = saved;
            }
        } virial_flag {s, s->compute_virial, npt.on};
    if (npt.on)
        s->compute_virial = true;
    // translational degrees of freedom of filter All, momentum conserving, 3D
    // (hoomd/md/IntegratorTwoStep.cc:154-189)
    // (a method on a proper subset of the particles keeps all 3 n: the centre of mass of the group is not
    // conserved, hoomd/md/IntegratorTwoStep.cc:172-178)
    const uint64_t n_all = s->dom.active ? s->dom.N_global : s->N;
    const uint64_t n_grp = it->d_member ? it->n_member : n_all;
    const double dof = 3.0 * (double)n_grp - (n_grp == n_all ? 3.0 : 0.0);

    // System::run resets the per-run statistics of every compute (hoomd/System.cc:103-110 ->
    // NeighborList::resetStats, hoomd/md/NeighborList.cc:1420-1427): num_builds counts this run only
    if (it->reset_stats)
        {
        nl->num_builds = 0;
        nl->shortest_rebuild = ~0ull;
        }
    HB2_CUDA(cudaEventRecord(it->ev0, st));
    // prepRun (hoomd/md/IntegratorTwoStep.cc:226-281): net force at the current timestep
    if (!it->prepared || !pf->computed || pf->computed_epoch != s->order_epoch
        || pf->computed_step != s->timestep)
        {
        HB2_TRY(pf->compute(s->timestep));
        HB2_TRY(integrator_add_extra(it, pf, s->timestep, true, nullptr));
        it->prepared = true;
        }
    else if ((!it->bonds.empty() || it->forces.size() > 1) && !pf->extra_added)
        {
        // a getter left the pure pair force in the array: same state, same list — put the net force
        // back without touching the neighbour list (an uninterrupted run would not rebuild here)
        HB2_TRY(pf->launch(s->timestep, true, nullptr));
        HB2_TRY(integrator_add_extra(it, pf, s->timestep, true, nullptr));
        }
    if (it->thermostat == 1)
        {
        HB2_TRY(launch_kinetic(s->vel[s->cur], s->N, s->d_partial, s->d_ticket, d_ke, st,
                               GroupMask {it->d_member, s->tag[s->cur]}));
        if (s->dom.active)
            HB2_TRY(domain_allreduce_sum_f64(s, d_ke, 1));
        }
    if (npt.on)
        {
        if (!pf->virial_valid)
            {
            // the net force of this timestep exists but without its virial: same list, one more pass
            HB2_TRY(pf->launch(s->timestep, true, nullptr));
            HB2_TRY(integrator_add_extra(it, pf, s->timestep, true, nullptr));
            }
        npt.tail_valid = false;
        npt.n_per_type.clear();
        for (auto* p : it->forces)
            if (p->tail_correction && !npt.tail_valid)
                {
                double te, tw;
                tail_correction(p, te, tw, nullptr, &npt.n_per_type);
                npt.tail_valid = true;
                }
        // thermo of the current timestep (the reference's advanceBarostat(timestep) at the top of step one
        // computes it when no step two left it behind)
        HB2_TRY(npt_thermo(it, pf, true));
        }

    it->pool_used = 0;
    it->marks.clear();
    // ConstantVolume on one GPU: step two of a timestep and step one of the next share a launch
    static const bool fuse_enabled = !(getenv("HB2_FUSE_INTEGRATE") && atoi(getenv("HB2_FUSE_INTEGRATE")) == 0);
    bool fuse = fuse_enabled && it->thermostat != 2 && !s->dom.active && !npt.on;
    if (fuse && !it->d_gen)
        {
        HB2_TRY(dev_alloc(&it->d_gen, 1));
        HB2_CUDA(cudaMemsetAsync(it->d_gen, 0, sizeof(uint32_t), st));
        }
    bool one_done = false; // the previous iteration's fused launch already did this timestep's step one
    for (uint64_t step = 0; step < nsteps; step++)
        {
        const uint64_t t = s->timestep;
        const double* rescale = nullptr;
        cudaEvent_t pe[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        if (it->profiling)
            {
            for (int q = 0; q < 5; q++)
                {
                it->marks.push_back(it->pool_used);
                pe[q] = it->next_event();
                }
            HB2_CUDA(cudaEventRecord(pe[0], st));
            }
        if (it->thermostat == 1 && !npt.on)
            {
            // the factor of step t comes from the KE left by step t-1: drawn inside step two's
            // reduction on a single GPU; after the KE all-reduce with a domain decomposition, and for
            // the first step of a run (KE from prepRun)
            if (step == 0 || s->dom.active)
                HB2_TRY(launch_bussi(d_ke, dof, it->dt, it->kT, it->tau, t, s->seed, 0u, d_rescale, st));
            rescale = d_rescale;
            }
        const bool check = nl->has_list && nl->check_dist;
        NptFactors nk;
        if (npt.on)
            {
            // TwoStepConstantPressure::integrateStepOne, host part: half a barostat step with the thermo of
            // this timestep, MTK and thermostat factors, propagators, the box; then the particles
            npt.ndof = dof;
            npt_advance_barostat(it, t);
            const double mtk = std::exp(-0.5 * it->dt * (npt.nu[0] + npt.nu[3] + npt.nu[5]) / npt.ndof);
            const double th = it->thermostat == 1
                                  ? bussi_factor_host(npt.ke, dof, it->dt, it->kT, it->tau, t, s->seed, 0u)
                                  : 1.0;
            npt_update_propagator(it);
            HB2_TRY(hb2_box_make(&s->box, npt.exp_r[0] * s->box.L[0], npt.exp_r[1] * s->box.L[1],
                                 npt.exp_r[2] * s->box.L[2]));
            nk.dt = (float)it->dt;
            nk.rescale_all = npt.rescale_all ? 1 : 0;
            double lambda_min = 1e300;
            for (int d = 0; d < 3; d++)
                {
                nk.exp_v[d] = (float)npt.exp_v[d];
                nk.exp_r[d] = (float)npt.exp_r[d];
                nk.exp_r_int[d] = (float)npt.exp_r_int[d];
                const double lam = check ? s->box.L[d] / nl->last_L[d] : 1.0;
                nk.lambda[d] = (float)lam;
                lambda_min = std::min(lambda_min, lam);
                }
            // hoomd/md/NeighborList.cc:1236-1260: the displacement left after the homogeneous dilation is
            // taken out, against ((r_cut_max + buffer) lambda_min - r_cut_max) / 2 per type
            double worst = 1e300;
            for (uint32_t ty = 0; ty < s->ntypes; ty++)
                {
                const double rmin = nl->rcut_max.size() > ty ? nl->rcut_max[ty] : nl->rcut_max_max;
                const double dmax = ((rmin + nl->buffer) * lambda_min - rmin) / 2.0;
                const double msq = dmax > 0.0 ? dmax * dmax : 0.0;
                worst = std::min(worst, msq);
                if (ty < (uint32_t)NPT_MAX_TYPES)
                    nk.maxsq[ty] = (float)msq;
                }
            if (s->ntypes > (uint32_t)NPT_MAX_TYPES) // more types than slots: the strictest threshold for the rest
                nk.maxsq[NPT_MAX_TYPES - 1] = (float)worst;
            nk.mtk = (float)(mtk * th);
            HB2_TRY(launch_step_one_npt(s->pos[s->cur], s->vel[s->cur], pf->d_force, s->image[s->cur], s->N, s->box,
                                        nk, nullptr, check ? nl->d_last_pos : nullptr, nl->d_conditions + 3, st,
                                        GroupMask {it->d_member, s->tag[s->cur]}));
            nk.mtk = (float)mtk; // step two: the MTK factor alone (Bussi rescales in step one only)
            }
        else if (!one_done)
            {
        // (Langevin: the previous step two left F + F_bd for this half step)
        // (valid for this particle order and this timestep only: a snapshot set in between, or a
        // rebuild triggered by a getter, falls back to the net force like computeAccelerations)
        const float4* f_one = (it->thermostat == 2 && it->total_valid && it->total_epoch == s->order_epoch
                               && it->total_step == t)
                                  ? it->d_total
                                  : pf->d_force;
        HB2_TRY(launch_step_one(s->pos[s->cur], s->vel[s->cur], f_one, s->image[s->cur], s->N,
                                s->box, (float)it->dt, rescale, check ? nl->d_last_pos : nullptr,
                                nl->maxshift(), nl->d_conditions + 3, st, GroupMask {it->d_member, s->tag[s->cur]}));
            }
        one_done = false;
        bool flag = false;
        const uint64_t builds_before = nl->num_builds;
        // Speculative force launch. On most steps the list is still valid, so the force pass is
        // queued right behind step one with the rebuild flag as its skip word, and the flag travels
        // to the host on a side stream meanwhile: the GPU never idles waiting for the host's
        // decision. When the flag turns out set, the skipped launch cost ~2 us; rebuild and launch
        // again. (The reference reads the flag with a blocking copy every step,
        // hoomd/md/NeighborListGPU.cc:78-82.)
        any_dirty = false;
        for (auto* p : it->forces)
            any_dirty = any_dirty || p->params_dirty;
        const bool speculate = check && !it->profiling && !nl->force_update && !nl->params_dirty
                               && !any_dirty && nl->buffer >= 1e-6
                               && t + 1 >= nl->last_update_step + nl->rebuild_check_delay;
        if (check)
            {
            // all ranks rebuild on the same step (hoomd/md/NeighborList.cc:1271-1285)
            if (s->dom.active)
                HB2_TRY(domain_allreduce_max_u32(s, nl->d_conditions + 3, 1));
            }
        if (speculate)
            {
            HB2_CUDA(cudaEventRecord(s->flag_ready, st));
            HB2_CUDA(cudaStreamWaitEvent(s->side_stream, s->flag_ready, 0));
            HB2_CUDA(cudaMemcpyAsync(nl->h_conditions + 3, nl->d_conditions + 3, sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, s->side_stream));
            HB2_CUDA(cudaEventRecord(s->flag_copied, s->side_stream));
            if (s->dom.active)
                HB2_TRY(domain_update_ghosts(s));
            HB2_TRY(pf->launch(t + 1, step + 1 == nsteps, nl->d_conditions + 3));
            HB2_TRY(integrator_add_extra(it, pf, t + 1, step + 1 == nsteps, nl->d_conditions + 3));
            HB2_CUDA(cudaEventSynchronize(s->flag_copied));
            flag = nl->h_conditions[3] != 0;
            if (flag)
                {
                HB2_TRY(nl->rebuild(t + 1));
                HB2_TRY(pf->launch(t + 1, step + 1 == nsteps, nullptr));
                HB2_TRY(integrator_add_extra(it, pf, t + 1, step + 1 == nsteps, nullptr));
                }
            }
        else
            {
            if (check)
                {
                // the one host read per step of the reference's loop (hoomd/md/NeighborListGPU.cc:78-82)
                HB2_CUDA(cudaMemcpyAsync(nl->h_conditions + 3, nl->d_conditions + 3, sizeof(uint32_t),
                                         cudaMemcpyDeviceToHost, st));
                HB2_CUDA(cudaStreamSynchronize(st));
                flag = nl->h_conditions[3] != 0;
                }
            if (it->profiling)
                {
                HB2_CUDA(cudaEventRecord(pe[1], st));
                s->mid_event = pe[2];
                }
            hb2_status cst = pf->compute(t + 1, check, flag, step + 1 == nsteps);
            s->mid_event = nullptr;
            HB2_TRY(cst);
            HB2_TRY(integrator_add_extra(it, pf, t + 1, step + 1 == nsteps, nullptr));
            }
        if (it->profiling)
            {
            HB2_CUDA(cudaEventRecord(pe[3], st));
            it->phase_n[1] += nl->num_builds - builds_before;
            }
        BussiNext bn;
        if (it->thermostat == 1 && !s->dom.active && step + 1 < nsteps)
            {
            bn.rescale = d_rescale;
            bn.dof = dof;
            bn.dt = it->dt;
            bn.kT = it->kT;
            bn.tau = it->tau;
            bn.timestep = t + 1;
            bn.seed = s->seed;
            }
        if (npt.on)
            {
            // TwoStepConstantPressure::integrateStepTwo: kick with the same propagator, then the thermo of
            // timestep t + 1 and the second half of the barostat step
            HB2_TRY(launch_step_two_npt(s->vel[s->cur], pf->d_force, s->N, nk, s->d_partial, s->d_ticket + 2,
                                        s->d_sums + 10, st, GroupMask {it->d_member, s->tag[s->cur]}));
            HB2_TRY(npt_thermo(it, pf, false));
            npt_advance_barostat(it, t + 1);
            }
        else if (it->thermostat == 2)
            {
            if (s->N > it->total_cap)
                {
                dev_free(it->d_total);
                it->d_total = nullptr;
                HB2_TRY(dev_alloc(&it->d_total, (size_t)s->cap));
                it->total_cap = s->cap;
                }
            HB2_TRY(launch_step_two_langevin(s->vel[s->cur], pf->d_force, it->d_total, s->pos[s->cur],
                                             s->tag[s->cur], s->N, (float)it->dt, it->d_gamma_coeff,
                                             s->seed, t, st, it->d_member));
            it->total_valid = true;
            it->total_epoch = s->order_epoch;
            it->total_step = t + 1;
            }
        else if (fuse && step + 1 < nsteps)
            {
            // step two of this timestep and step one of the next in one launch (integrate.cu); the Bussi
            // factor of the next step is drawn between the halves
            HB2_TRY(launch_step_two_one(s->pos[s->cur], s->vel[s->cur], pf->d_force, s->image[s->cur], s->N, s->box,
                                        (float)it->dt, check ? nl->d_last_pos : nullptr, nl->maxshift(),
                                        nl->d_conditions + 3, s->d_partial, s->d_ticket,
                                        it->thermostat == 1 ? d_ke : nullptr, it->thermostat == 1 ? &bn : nullptr,
                                        it->d_gen, ++it->gen_epoch, st, &one_done,
                                        GroupMask {it->d_member, s->tag[s->cur]}));
            if (!one_done) // the cooperative launch was refused: the stand-alone kernels from here on
                {
                fuse = false;
                HB2_TRY(launch_step_two(s->vel[s->cur], pf->d_force, s->N, (float)it->dt, nullptr, s->d_partial,
                                        s->d_ticket, it->thermostat == 1 ? d_ke : nullptr, st, &bn,
                                        GroupMask {it->d_member, s->tag[s->cur]}));
                }
            }
        else
            HB2_TRY(launch_step_two(s->vel[s->cur], pf->d_force, s->N, (float)it->dt, nullptr, s->d_partial,
                                    s->d_ticket, it->thermostat == 1 ? d_ke : nullptr, st, &bn,
                                    GroupMask {it->d_member, s->tag[s->cur]}));
        if (it->thermostat == 1 && s->dom.active)
            HB2_TRY(domain_allreduce_sum_f64(s, d_ke, 1));
        if (it->profiling)
            HB2_CUDA(cudaEventRecord(pe[4], st));
        s->timestep = t + 1;
        }
    HB2_CUDA(cudaMemcpyAsync(s->d_timestep, &s->timestep, sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    HB2_CUDA(cudaEventRecord(it->ev1, st));
    HB2_CUDA(cudaStreamSynchronize(st));
    for (auto* b : it->bonds)
        HB2_TRY(b->check_complete());
    HB2_CUDA(cudaEventElapsedTime(&it->last_ms, it->ev0, it->ev1));
    if (it->profiling)
        {
        for (size_t k = 0; k + 4 < it->marks.size() + 0 && k < it->marks.size(); k += 5)
            {
            cudaEvent_t* e = &it->pool[it->marks[k]];
            float ab, bc, cd, de;
            HB2_CUDA(cudaEventElapsedTime(&ab, e[0], e[1]));
            HB2_CUDA(cudaEventElapsedTime(&bc, e[1], e[2]));
            HB2_CUDA(cudaEventElapsedTime(&cd, e[2], e[3]));
            HB2_CUDA(cudaEventElapsedTime(&de, e[3], e[4]));
            it->phase_ms[0] += ab + de;
            it->phase_ms[1] += bc;
            it->phase_ms[2] += cd;
            it->phase_n[0] += 1;
            it->phase_n[2] += 1;
            }
        }
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_set_profiling(hb2_integrator* it, int on)
    {
    if (!it)
        return HB2_ERR_INVALID;
    it->profiling = on != 0;
    for (int q = 0; q < 3; q++)
        {
        it->phase_ms[q] = 0;
        it->phase_n[q] = 0;
        }
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_get_profile(hb2_integrator* it, double phase_ms[3],
                                                 uint64_t phase_n[3])
    {
    if (!it || !phase_ms || !phase_n)
        return HB2_ERR_INVALID;
    for (int q = 0; q < 3; q++)
        {
        phase_ms[q] = it->phase_ms[q];
        phase_n[q] = it->phase_n[q];
        }
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_thermo(hb2_integrator* it, double* kinetic, double* potential,
                                            double virial6[6])
    {
    if (!it || it->forces.empty())
        return HB2_ERR_INVALID;
    hb2_system* s = it->sys;
    HB2_CUDA(cudaSetDevice(s->device));
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    for (auto* p : it->forces)
        {
        double w[7];
        HB2_TRY(sums7(p, w));
        for (int q = 0; q < 7; q++)
            v[q] += w[q];
        // external energy / virial of the force (LJ tail correction): ComputeThermo adds
        // getExternalEnergy() and getExternalVirial() (hoomd/md/ComputeThermo.cc:212-220)
        double te, tw;
        tail_correction(p, te, tw);
        v[0] += te;
        if (s->compute_virial)
            {
            v[1] += tw;
            v[4] += tw;
            v[6] += tw;
            }
        }
    for (auto* b : it->bonds)
        {
        // the bond force's own arrays, reduced like the pair force's
        HB2_TRY(b->compute_own());
        double w[7];
        HB2_TRY(launch_reduce_energy(b->d_force, s->compute_virial ? b->d_virial : nullptr, b->cap_N, s->N,
                                     s->d_partial, s->d_ticket + 1, s->d_sums + 1, s->stream));
        if (s->dom.active)
            HB2_TRY(domain_allreduce_sum_f64(s, s->d_sums + 1, 7));
        HB2_CUDA(cudaMemcpyAsync(w, s->d_sums + 1, sizeof(double) * 7, cudaMemcpyDeviceToHost, s->stream));
        HB2_CUDA(cudaStreamSynchronize(s->stream));
        for (int q = 0; q < 7; q++)
            v[q] += w[q];
        }
    double ke = 0.0;
    HB2_TRY(launch_kinetic(s->vel[s->cur], s->N, s->d_partial, s->d_ticket + 2, s->d_sums + 9,
                           s->stream));
    if (s->dom.active)
        HB2_TRY(domain_allreduce_sum_f64(s, s->d_sums + 9, 1));
    HB2_CUDA(cudaMemcpyAsync(&ke, s->d_sums + 9, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    if (kinetic)
        *kinetic = ke;
    if (potential)
        *potential = v[0];
    if (virial6)
        for (int q = 0; q < 6; q++)
            virial6[q] = v[1 + q];
    return HB2_OK;
    }

// Kinetic part of the pressure tensor, sum m v_a v_b over every particle (hoomd/md/ComputeThermo.cc:104-123):
// xx, xy, xz, yy, yz, zz. With the virial sums of hb2_integrator_thermo: P_ab = (K_ab + W_ab) / V.
extern "C" hb2_status hb2_integrator_kinetic_tensor(hb2_integrator* it, double kt[6])
    {
    if (!it || !kt)
        return HB2_ERR_INVALID;
    hb2_system* s = it->sys;
    HB2_CUDA(cudaSetDevice(s->device));
    NptFactors k0; // dt = 0: the velocities as they are
    HB2_TRY(launch_step_two_npt(s->vel[s->cur], nullptr, s->N, k0, s->d_partial, s->d_ticket + 2, s->d_sums + 10,
                                s->stream));
    if (s->dom.active)
        HB2_TRY(domain_allreduce_sum_f64(s, s->d_sums + 10, 6));
    HB2_CUDA(cudaMemcpyAsync(kt, s->d_sums + 10, 6 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    HB2_CUDA(cudaStreamSynchronize(s->stream));
    return HB2_OK;
    }

extern "C" hb2_status hb2_integrator_last_run_ms(hb2_integrator* it, float* ms)
    {
    if (!it || !ms)
        return HB2_ERR_INVALID;
    *ms = it->last_ms;
    return HB2_OK;
    }
