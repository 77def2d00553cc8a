// comm.cuh — spatial domain decomposition of the pair-force path over the GPUs of one NVSwitch box:
// one process per GPU, a 3D grid of orthorhombic sub-domains, NCCL send/recv for particle migration
// and the ghost layer.
//
// Replaces, for this path: DomainDecomposition (hoomd/DomainDecomposition.cc:327-409 grid choice,
// :604-660 placeParticle), Communicator::migrateParticles / exchangeGhosts / beginUpdateGhosts
// (hoomd/Communicator.cc:1572,1795,2630; GPU plan kernel hoomd/CommunicatorGPU.cu:306-357) and the
// global rebuild vote (hoomd/md/NeighborList.cc:1271-1285). The reference stages every message through
// host memory around MPI_Isend/Irecv; here the buffers stay in HBM and move over NVLink.
//
// Geometry. Along a dimension with n > 1 ranks a domain owns fractional coordinates
// [coord/n, (coord+1)/n) of the GLOBAL box. `domain_g` maps a particle's global fractional coordinate
// to g = (f n - coord) mod n: g in [0,1) is owned, [1, 1 + (n-1)/2) lies above the domain, the rest
// below (returned negative). Ownership, migration, ghost selection and cell binning all use this one
// function (fp64, same operation order on host and device), so they can never disagree.
// Ghost particles keep their exact global coordinates (no +-L shift is ever applied to stored
// positions: an fp32 shift would round and break the bit-exact neighbour test); the cell grid is local
// (interior cells plus one ghost-cell layer per decomposed side) and the neighbour-list build applies
// the periodic image per STENCIL CELL, exactly as it does for the periodic ring of a single GPU.
#pragma once
#include "common.cuh"

struct hb2_system;

namespace hb2
    {
struct DomainGeom // passed by value to kernels
    {
    uint32_t n[3];     // ranks along x, y, z (1: not decomposed, periodic ring of cells)
    uint32_t coord[3]; // this rank's grid coordinate
    uint32_t dimi[3];  // interior cells
    uint32_t dimt[3];  // total cells = dimi + 2 where n > 1
    };

inline DomainGeom periodic_geom(const uint32_t dim[3])
    {
    DomainGeom g;
    for (int d = 0; d < 3; d++)
        {
        g.n[d] = 1;
        g.coord[d] = 0;
        g.dimi[d] = g.dimt[d] = dim[d];
        }
    return g;
    }

HB2_HOSTDEVICE double domain_g(double f, uint32_t n, uint32_t coord)
    {
    double g = f * (double)n - (double)coord;
    if (g < 0.0)
        g += (double)n;
    if (g >= (double)n)
        g -= (double)n;
    if (g < 1.0)
        return g < 0.0 ? 0.0 : g;
    if (g < 1.0 + 0.5 * (double)(n - 1))
        return g;
    return g - (double)n;
    }

// One-word collectives over NVLink peer memory (comm.cu: p2p_allreduce_kernel). Every rank owns a mailbox of
// [2 epochs parity][P2P_MAX_RANKS] boxes in its HBM, mapped into every other process with CUDA IPC; a reduction
// is ONE small kernel: store my value into slot [rank] of every peer's mailbox, wait until my own mailbox holds
// this epoch's value of every rank, reduce in rank order (bit-identical on all ranks).
constexpr int P2P_MAX_RANKS = 16;
struct P2PBox
    {
    unsigned long long epoch;
    double value;
    };

struct Domain
    {
    bool active = false;
    int rank = 0, nranks = 1;
    uint32_t grid[3] = {1, 1, 1};
    uint32_t coord[3] = {0, 0, 0};
    int nbr[3][2] = {{0, 0}, {0, 0}, {0, 0}}; // [dim][0 = below, 1 = above]
    void* comm = nullptr;                     // ncclComm_t
    uint64_t N_global = 0;
    bool need_vel = false; // ghost velocities are refreshed every step (DPD)

    // ghost plan of the current neighbour list (indices are post-sort particle indices)
    uint32_t n_send[3][2] = {{0, 0}, {0, 0}, {0, 0}}; // [dim][0 = to below, 1 = to above]
    uint32_t n_recv[3][2] = {{0, 0}, {0, 0}, {0, 0}}; // [dim][0 = from below, 1 = from above]
    uint32_t ghost_off[3] = {0, 0, 0};                // arrival ordinal of the first ghost of stage d
    uint32_t* d_send_idx[3] = {nullptr, nullptr, nullptr}; // [to above | to below]
    uint32_t send_cap[3] = {0, 0, 0};
    uint32_t* d_ghost_slot = nullptr; // arrival ordinal -> particle index
    float4* d_sendbuf = nullptr;      // 2 float4 per entry (pos, vel)
    float4* d_recvbuf = nullptr;
    size_t buf_cap = 0; // entries

    // selection scratch
    uint8_t* d_cls = nullptr;
    uint32_t* d_sel = nullptr;   // 3 lists of cap entries
    uint32_t* d_count = nullptr; // [0..2] selected counts, [4..7] exchanged counts
    uint32_t* h_count = nullptr; // pinned mirror
    void* d_cub = nullptr;
    size_t cub_bytes = 0;
    uint32_t* d_inv = nullptr; // inverse of the cell-sort permutation
    uint32_t scratch_cap = 0;

    // peer-memory mailboxes (null / false: the NCCL all-reduces serve)
    bool p2p = false;
    P2PBox* d_mail = nullptr;                       // this rank's mailbox
    P2PBox* peer_mail[P2P_MAX_RANKS] = {};          // every rank's mailbox as mapped here (own included)
    P2PBox** d_peer_mail = nullptr;                 // the same table on the device
    unsigned long long p2p_epoch = 0;
    };

// host entry points (comm.cu)
hb2_status domain_setup(hb2_system* s, const uint32_t grid[3], int rank, int nranks,
                        const void* unique_id);
void domain_destroy(hb2_system* s);
hb2_status domain_ensure_scratch(hb2_system* s);
// rebuild steps: move particles that left the domain, then (re)build the ghost layer of width r_ghost
hb2_status domain_migrate(hb2_system* s);
hb2_status domain_exchange_ghosts(hb2_system* s, double r_ghost);
// after the cell sort permuted the particle arrays: d_order[k] = old index of new slot k
hb2_status domain_remap(hb2_system* s, const uint32_t* d_order, uint32_t Ntot);
// every other step: refresh ghost positions (and velocities when need_vel)
hb2_status domain_update_ghosts(hb2_system* s);
hb2_status domain_allreduce_max_u32(hb2_system* s, uint32_t* d_val, uint32_t n);
hb2_status domain_allreduce_sum_f64(hb2_system* s, double* d_val, uint32_t n);
    } // namespace hb2
