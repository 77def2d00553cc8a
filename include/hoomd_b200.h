/* hoomd_b200.h — C ABI of the B200-native short-range MD pair-force step.
 *
 * This is the drop-in boundary for the path BASELINE.json:north_star names: cell binning ->
 * neighbour-list build -> PotentialPair<Evaluator> force/energy/virial -> ConstantVolume half steps.
 * Plain pointers and sizes only; no C++/torch types cross this boundary; no exceptions either: every
 * entry point returns an hb2_status and hb2_last_error() holds the message for the calling thread.
 *
 * Two layers:
 *   (1) hb2_gpu_*      stateless kernel drivers over caller-owned DEVICE arrays. They replace the
 *                      reference's GPU driver functions one for one (cited per function) and are what
 *                      a HOOMD `...GPU` class would call from computeForces/buildNlist/integrateStep*.
 *   (2) hb2_system_* … opaque-handle engine over HOST buffers (ParticleData + NeighborList +
 *                      ForceCompute + IntegratorTwoStep in one place, state resident in HBM). The
 *                      Python operator surface (hoomd_blue_b200/) binds this layer with ctypes.
 *
 * Citations are to /root/reference (HOOMD-blue v4.8.2) file:line.
 * Precision contract: particle state is fp32 (HOOMD_LONGREAL_SIZE=32 layout: Scalar4 = float4),
 * neighbour-list membership is decided in fp64 with the CPU build's operation order (bit-exact pair
 * set), per-particle sums are fp32, global sums fp64.
 * Scope: 3D orthorhombic periodic boxes (tilt factors must be 0).
 */
#ifndef HOOMD_B200_H
#define HOOMD_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define HB2_ABI_VERSION 2

typedef enum
    {
    HB2_OK = 0,
    HB2_ERR_INVALID = 1,        /* bad argument / wrong state (std::invalid_argument in the reference) */
    HB2_ERR_CUDA = 2,           /* CUDA runtime error (CHECK_CUDA_ERROR, hoomd/md/PotentialPairGPU.h:163) */
    HB2_ERR_NAN_POSITION = 3,   /* hoomd/CellList.cc:575-590 "Particle with unique tag ... has NaN" */
    HB2_ERR_OUT_OF_BOX = 4,     /* hoomd/CellList.cc:592-604 "is no longer in the simulation box" */
    HB2_ERR_BOX_TOO_SMALL = 5,  /* hoomd/md/NeighborList.cc:544-560 "box too small for r_cut+buffer" */
    HB2_ERR_NLIST_OVERFLOW = 6, /* hoomd/md/NeighborList.cc:1601-1617 (reference regrows; see below) */
    HB2_ERR_NO_DEVICE = 7,      /* no CUDA device: there is NO CPU fallback in this library */
    HB2_ERR_COMM = 8            /* NCCL error on the domain-decomposed path */
    } hb2_status;

const char* hb2_last_error(void);
int hb2_abi_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t hb2_launch_count(void);

/* ------------------------------------------------------------------------------------------ *
 * POD box (replaces BoxDim in kernel signatures; hoomd/BoxDim.h:183-200: hi = L/2, lo = -hi).
 * The fp64 copy feeds the bit-exact membership test and the binning, the fp32 copy the hot loops.
 * ------------------------------------------------------------------------------------------ */
typedef struct
    {
    double L[3], lo[3], hi[3], Linv[3];
    float Lf[3], lof[3], hif[3], Linvf[3];
    int32_t periodic[3];
    int32_t _pad;
    } hb2_box;

hb2_status hb2_box_make(hb2_box* box, double Lx, double Ly, double Lz);

/* Evaluators compiled into the library (the pair_plugin template adds more, see
 * hoomd_blue_b200/csrc/evaluators.cuh and INTEGRATION.md). */
typedef enum
    {
    HB2_EVAL_LJ = 0,  /* hoomd/md/EvaluatorPairLJ.h:95-270; params {epsilon, sigma} */
    HB2_EVAL_DPD = 1, /* hoomd/md/EvaluatorPairDPDThermoDPD.h:144-277; params {A, gamma} (+kT, seed) */
    HB2_EVAL_DPD_CONSERVATIVE = 2, /* same header :186-203; params {A} */
    HB2_EVAL_EXAMPLE = 3, /* example_plugins/pair_plugin/EvaluatorPairExample.h; params {k, sigma} */
    /* more evaluators through the same template (hoomd/md/CMakeLists.txt:505-527): */
    HB2_EVAL_GAUSS = 4,            /* hoomd/md/EvaluatorPairGauss.h; {epsilon, sigma} */
    HB2_EVAL_YUKAWA = 5,           /* hoomd/md/EvaluatorPairYukawa.h; {epsilon, kappa} */
    HB2_EVAL_LJ1208 = 6,           /* hoomd/md/EvaluatorPairLJ1208.h; {epsilon, sigma} */
    HB2_EVAL_LJ0804 = 7,           /* hoomd/md/EvaluatorPairLJ0804.h; {epsilon, sigma} */
    HB2_EVAL_FORCE_SHIFTED_LJ = 8, /* hoomd/md/EvaluatorPairForceShiftedLJ.h; {epsilon, sigma} */
    HB2_EVAL_MIE = 9,              /* hoomd/md/EvaluatorPairMie.h; {epsilon, sigma, n, m} */
    HB2_EVAL_MORSE = 10,           /* hoomd/md/EvaluatorPairMorse.h; {D0, alpha, r0} */
    HB2_EVAL_BUCKINGHAM = 11,      /* hoomd/md/EvaluatorPairBuckingham.h; {A, rho, C} */
    HB2_EVAL_EXPANDED_LJ = 12,     /* hoomd/md/EvaluatorPairExpandedLJ.h; {epsilon, sigma, delta} */
    HB2_EVAL_EXPANDED_GAUSSIAN = 13, /* hoomd/md/EvaluatorPairExpandedGaussian.h; {epsilon, sigma, delta} */
    HB2_EVAL_LJ_GAUSS = 14,        /* hoomd/md/EvaluatorPairLJGauss.h; {epsilon, sigma, r0} */
    HB2_EVAL_TWF = 15,             /* hoomd/md/EvaluatorPairTWF.h; {epsilon, sigma, alpha} */
    HB2_EVAL_DPDLJ = 16,           /* hoomd/md/EvaluatorPairDPDThermoLJ.h; {epsilon, sigma, gamma} (+kT, seed) */
    HB2_EVAL_TABLE = 17,           /* hoomd/md/EvaluatorPairTable.h; tabulated U(r), F(r): hb2_pair_set_table */
    HB2_EVAL_EXPANDED_MIE = 18,    /* hoomd/md/EvaluatorPairExpandedMie.h; {epsilon, sigma, n, m, delta} */
    HB2_EVAL_OPP = 19,             /* hoomd/md/EvaluatorPairOPP.h; {C1, C2, eta1, eta2, k, phi} */
    HB2_EVAL_FOURIER = 20          /* hoomd/md/EvaluatorPairFourier.h; {a2, a3, a4, b2, b3, b4}; modes none / xplor */
    } hb2_evaluator;
/* Stride, in floats, of one type pair's packed evaluator parameters (param_type is at most 32 bytes;
 * evaluators with variable-size data, like Table, keep a device pointer in it) and the most user
 * parameters hb2_pair_set_params takes. */
#define HB2_PARAM_FLOATS 8

/* hoomd/md/PotentialPair.h:104-109 */
typedef enum
    {
    HB2_SHIFT_NONE = 0,
    HB2_SHIFT_SHIFT = 1,
    HB2_SHIFT_XPLOR = 2
    } hb2_shift_mode;

/* ========================================================================================== *
 * Layer 1: stateless kernel drivers (device pointers, caller's stream).
 * `stream` is a cudaStream_t passed as void*.
 * ========================================================================================== */

/* Replaces gpu_compute_cell_list (hoomd/CellListGPU.cu:173-194) + the deterministic sort pass
 * (:439-565). Bins N particles with the reference's rule (hoomd/CellList.cc:455-470, fp64
 * makeFraction hoomd/BoxDim.h:276-284) on a dim[0] x dim[1] x dim[2] grid, then key-sorts (cell,
 * index) — no atomics, order is deterministic. Outputs: order[k] = old index of the k-th particle in
 * cell order; cell c owns sorted slots [cell_start[c], cell_end[c]) (Nc entries each, an empty cell
 * reads [0, 0)). conditions[3] = {0, nan idx+1, out-of-box idx+1} as hoomd/CellList.cc:549-607.
 * tmp/tmp_bytes: scratch; call with tmp = NULL to get the size. */
hb2_status hb2_gpu_compute_cell_order(const float* d_pos4, uint32_t N, const hb2_box* box,
                                      const uint32_t dim[3], uint32_t* d_order,
                                      uint32_t* d_cell_start, uint32_t* d_cell_end,
                                      uint32_t* d_conditions, void* d_tmp, size_t* tmp_bytes,
                                      void* stream);

/* Replaces gpu_compute_nlist_binned (hoomd/md/NeighborListGPUBinned.cu:756-824) for particles that
 * are already in cell order (d_pos4 sorted so that cell c owns [cell_start[c], cell_end[c])).
 * Membership rule hoomd/md/NeighborListBinned.cc:144-157: j != i, r_cut(ti,tj) > 0,
 * |minImage(pi - pj)|^2 <= (r_cut(ti,tj) + r_buff)^2 decided in fp64. Full list, row i at
 * d_nlist[i * pitch], ascending j. d_r_listsq: ntypes^2 doubles indexed tj*ntypes+ti
 * (hoomd/md/NeighborList.cc:505-506). d_conditions[0] receives max(n_neigh) (overflow if > pitch).
 * Also writes d_last_pos4 = d_pos4 (hoomd/md/NeighborList.cc:1293-1311).
 * Row padding: pitch must be a multiple of 32 and every row is padded to a multiple of 32 entries
 * with the index N_total of a far-away dummy particle, which this call writes to d_pos4[N_total]
 * (so d_pos4 — and d_vel4 / d_tag for DPD — need N_total + 1 entries); the force driver relies on it. */
hb2_status hb2_gpu_compute_nlist_binned(float* d_pos4, uint32_t N, uint32_t N_total,
                                        const hb2_box* box, const uint32_t dim[3],
                                        const uint32_t* d_cell_start, const uint32_t* d_cell_end,
                                        const double* d_r_listsq, uint32_t ntypes,
                                        uint32_t* d_nlist, uint32_t pitch,
                                        uint32_t* d_n_neigh, float* d_last_pos4,
                                        uint32_t* d_conditions, void* stream);

/* Replaces gpu_compute_pair_forces<evaluator> (hoomd/md/PotentialPairGPU.cuh:579-656; args
 * pair_args_t :44-98). d_params: ntypes^2 x HB2_PARAM_FLOATS floats of evaluator-packed
 * parameters (hb2_pack_params), d_rcutsq / d_ronsq: ntypes^2 floats, all indexed tj*ntypes+ti. d_force4 =
 * (fx,fy,fz,pe) per particle, d_virial = 6 x virial_pitch floats (xx,xy,xz,yy,yz,zz) or NULL
 * (pressure flag off, hoomd/md/PotentialPair.h:739-752). DPD-only inputs may be NULL otherwise.
 * face_margin: a particle farther than this from every box face takes the loop without image shifts: no pair
 * closer than r_cut can cross a face from there, and a listed neighbour in another image lies beyond the cut-off
 * both truly and by the plain difference (pass r_cut_max or more; <= 0 disables the shortcut). */
hb2_status hb2_gpu_compute_pair_forces(hb2_evaluator ev, hb2_shift_mode mode, const float* d_pos4,
                                       const float* d_vel4, const uint32_t* d_tag, uint32_t N,
                                       const hb2_box* box, const uint32_t* d_n_neigh,
                                       const uint32_t* d_nlist, uint32_t pitch,
                                       const float* d_params, const float* d_rcutsq,
                                       const float* d_ronsq, uint32_t ntypes, float* d_force4,
                                       float* d_virial, size_t virial_pitch, uint32_t seed,
                                       uint64_t timestep, float dt, float kT, float face_margin,
                                       uint32_t threads_per_particle, void* stream);

/* Host helper: user parameters -> packed evaluator parameters (the param_type constructor of the
 * evaluator, e.g. hoomd/md/EvaluatorPairLJ.h:119-140: lj1 = 4 eps sigma^12, lj2 = 4 eps sigma^6). */
hb2_status hb2_pack_params(hb2_evaluator ev, const double* user_params, uint32_t n_user,
                           float packed[HB2_PARAM_FLOATS]);

/* Replace gpu_nvt_rescale_step_one / _step_two (hoomd/md/TwoStepConstantVolumeGPU.cu:102-114,
 * :215-223; arithmetic hoomd/md/TwoStepConstantVolume.cc:33-79, :234-260). Step one also performs
 * the neighbour-list distance check against d_last_pos4 (hoomd/md/NeighborList.cc:1244-1269,
 * replacing gpu_nlist_needs_update_check_new, hoomd/md/NeighborListGPU.cu:42-84): *d_flag is set to 1
 * when any particle moved >= maxshift. d_rescale: device scalar (thermostat factor) or NULL = 1. */
hb2_status hb2_gpu_nve_step_one(float* d_pos4, float* d_vel4, const float* d_force4,
                                int32_t* d_image3, uint32_t N, const hb2_box* box, float dt,
                                const double* d_rescale, const float* d_last_pos4, float maxshift,
                                uint32_t* d_flag, void* stream);
hb2_status hb2_gpu_nve_step_two(float* d_vel4, const float* d_force4, uint32_t N, float dt,
                                const double* d_rescale, double* d_ke_out, void* stream);

/* ========================================================================================== *
 * Layer 2: engine with opaque handles (state resident in HBM; host buffers in/out).
 * ========================================================================================== */
typedef struct hb2_system hb2_system;         /* ~ SystemDefinition + ParticleData (hoomd/ParticleData.h:1285) */
typedef struct hb2_nlist hb2_nlist;           /* ~ NeighborListGPUBinned (hoomd/md/NeighborList.h:121) */
typedef struct hb2_pair hb2_pair;             /* ~ PotentialPairGPU<evaluator> (hoomd/md/PotentialPairGPU.h:89) */
#define HB2_BOND_PARAM_FLOATS 8
typedef enum hb2_bond_kind
    {
    HB2_BOND_HARMONIC = 0, /* hoomd/md/EvaluatorBondHarmonic.h */
    HB2_BOND_FENEWCA = 1,  /* hoomd/md/EvaluatorBondFENE.h */
    HB2_ANGLE_HARMONIC = 2, /* hoomd/md/HarmonicAngleForceCompute.cc; {k, t0} per ANGLE type */
    HB2_ANGLE_COSINESQ = 3, /* hoomd/md/CosineSqAngleForceCompute.cc; {k, t0} per ANGLE type */
    HB2_DIHEDRAL_PERIODIC = 4, /* hoomd/md/HarmonicDihedralForceCompute.cc; {k, d, n, phi0} per DIHEDRAL type */
    HB2_DIHEDRAL_OPLS = 5      /* hoomd/md/OPLSDihedralForceCompute.cc; {k1, k2, k3, k4} per DIHEDRAL type */
    } hb2_bond_kind;
typedef struct hb2_bond hb2_bond;             /* ~ PotentialBond<EvaluatorBondHarmonic> */
typedef struct hb2_integrator hb2_integrator; /* ~ IntegratorTwoStep + TwoStepConstantVolume */

/* device: CUDA ordinal. ntypes: number of particle types. seed: hoomd.Simulation(seed) (16 bit used,
 * hoomd/RandomNumbers.h:136-171). */
hb2_status hb2_system_create(hb2_system** out, int device, uint32_t ntypes, uint32_t seed);
void hb2_system_destroy(hb2_system* sys);
hb2_status hb2_system_set_box(hb2_system* sys, double Lx, double Ly, double Lz, double xy, double xz,
                              double yz);
/* Snapshot in (hoomd/Snapshot / GSD frame layout): N x 3 float positions and velocities, N masses,
 * N type ids, N x 3 images (may be NULL = 0), tags are 0..N-1 in array order. */
hb2_status hb2_system_set_particles(hb2_system* sys, uint32_t N, const float* pos3, const float* vel3,
                                    const float* mass, const uint32_t* typeid_,
                                    const int32_t* image3);
/* Snapshot out, tag order. Any pointer may be NULL. */
hb2_status hb2_system_get_particles(hb2_system* sys, float* pos3, float* vel3, float* mass,
                                    uint32_t* typeid_, int32_t* image3);
uint32_t hb2_system_n(const hb2_system* sys); /* owned particles on this rank */

/* ---- domain decomposition (one process per GPU; replaces DomainDecomposition + Communicator:
 * hoomd/DomainDecomposition.cc:327-409,604-660, hoomd/Communicator.cc:1510-1570). Bootstrap: rank 0
 * calls hb2_comm_unique_id and broadcasts the 128 bytes (torch.distributed / MPI / any side channel),
 * then EVERY rank calls hb2_system_set_domain (collective: ncclCommInitRank) before setting particles.
 * grid = ranks along x, y, z with grid[0]*grid[1]*grid[2] == nranks; rank = (cz*ny + cy)*nx + cx.
 * With a domain set, set_particles[_tagged] keeps only the particles this rank owns, every run /
 * compute / energy call is collective, and per-particle results are read with the *_local getters
 * (local order, with tags) and gathered by the caller. */
hb2_status hb2_comm_unique_id(void* out128);
hb2_status hb2_system_set_domain(hb2_system* sys, const uint32_t grid[3], uint32_t rank,
                                 uint32_t nranks, const void* nccl_unique_id);
/* Optional, after hb2_system_set_domain: one-word reductions of every step (the rebuild vote of
 * hoomd/md/NeighborList.cc:1271-1285 — MPI_Allreduce there — and the thermostat's kinetic energy) over NVLink
 * peer memory instead of NCCL launches. Every rank exports a 64-byte CUDA IPC handle of its mailbox, the caller
 * gathers the nranks handles in rank order (any side channel) and hands them to every rank; the switch is
 * collective (one rank that cannot map a peer, or HB2_P2P=0, keeps NCCL everywhere). */
hb2_status hb2_domain_ipc_handle(hb2_system* sys, void* out64);
hb2_status hb2_domain_set_peers(hb2_system* sys, const void* handles, uint32_t nranks);
int hb2_domain_p2p_active(const hb2_system* sys);
/* hoomd/DomainDecomposition.cc:327-409 (findDecomposition): the grid with the least interface area. */
hb2_status hb2_domain_find_decomposition(uint32_t nranks, double Lx, double Ly, double Lz,
                                         uint32_t grid[3]);
/* Host helpers (no GPU needed): owner rank of each position (placeParticle,
 * hoomd/DomainDecomposition.cc:604-660) and the ghost-layer plan of one rank (bit 2d: within r_ghost of
 * the lower face of dimension d, bit 2d+1: of the upper face; hoomd/CommunicatorGPU.cu:306-357). */
hb2_status hb2_domain_place(const hb2_box* box, const uint32_t grid[3], const float* pos3, uint32_t n,
                            uint32_t* rank_out);
hb2_status hb2_domain_ghost_mask(const hb2_box* box, const uint32_t grid[3], const uint32_t coord[3],
                                 double r_ghost, const float* pos3, uint32_t n, uint8_t* mask_out);
/* Candidates with explicit global tags (any subset of the N_global particles; with a domain only the
 * owned ones are kept), e.g. each rank passing just its own slab of a large lattice. */
hb2_status hb2_system_set_particles_tagged(hb2_system* sys, uint32_t n, uint64_t N_global,
                                           const uint32_t* tag, const float* pos3, const float* vel3,
                                           const float* mass, const uint32_t* typeid_,
                                           const int32_t* image3);
/* Owned particles in local (cell) order with their tags; arrays of hb2_system_n entries. */
hb2_status hb2_system_get_local(hb2_system* sys, uint32_t* tag, float* pos3, float* vel3, float* mass,
                                uint32_t* typeid_, int32_t* image3);
uint64_t hb2_system_n_global(const hb2_system* sys);
uint32_t hb2_system_n_ghost(const hb2_system* sys);
uint64_t hb2_system_timestep(const hb2_system* sys);
hb2_status hb2_system_set_timestep(hb2_system* sys, uint64_t timestep);
/* hoomd.Simulation.always_compute_pressure (hoomd/simulation.py:447-483): virials are only computed
 * when set (PDataFlags pressure_tensor, hoomd/ParticleData.h:70-84). */
hb2_status hb2_system_set_compute_virial(hb2_system* sys, int on);
hb2_status hb2_system_synchronize(hb2_system* sys);

/* hoomd.md.nlist.Cell(buffer, rebuild_check_delay, check_dist, deterministic) hoomd/md/nlist.py:375 */
hb2_status hb2_nlist_create(hb2_nlist** out, hb2_system* sys, double buffer,
                            uint32_t rebuild_check_delay, int check_dist);
void hb2_nlist_destroy(hb2_nlist* nl);
hb2_status hb2_nlist_set_buffer(hb2_nlist* nl, double buffer);
/* hoomd.md.nlist.NeighborList.r_cut (hoomd/md/nlist.py:127-131): the list's own cut-off for the unordered
 * type pair; the effective r_cut is the maximum over this matrix and every pair force on the list
 * (hoomd/md/NeighborList.cc:452-529). */
hb2_status hb2_nlist_set_r_cut(hb2_nlist* nl, uint32_t ti, uint32_t tj, double r_cut);
/* cpu/gpu_local_nlist_arrays (hoomd/md/nlist.py:154-268, hoomd/md/NeighborList.h:662-722): device pointers
 * to n_neigh (n_rows entries) and nlist (row i at i * pitch, local particle indices) of the last build,
 * valid until the next rebuild; the head list is i * pitch. */
hb2_status hb2_nlist_device_arrays(hb2_nlist* nl, const uint32_t** d_n_neigh, const uint32_t** d_nlist,
                                   uint32_t* pitch, uint32_t* n_rows);
/* Force a (re)build now at the current positions (NeighborList::compute with m_force_update,
 * hoomd/md/NeighborList.cc:363-435). */
hb2_status hb2_nlist_compute(hb2_nlist* nl);
uint64_t hb2_nlist_num_builds(const hb2_nlist* nl);       /* hoomd/md/nlist.py:307 */
uint64_t hb2_nlist_shortest_rebuild(const hb2_nlist* nl); /* hoomd/md/nlist.py:315 */
hb2_status hb2_nlist_dimensions(const hb2_nlist* nl, uint32_t dim[3]); /* hoomd/md/nlist.py:399 */
/* Number of (i<j) pairs currently listed and the pairs themselves as tags, like
 * NeighborList::getPairListPython in half mode (hoomd/md/NeighborList.cc:1787-1942): each unordered
 * pair once. pairs: n x 2 uint32 host buffer of capacity cap pairs. */
hb2_status hb2_nlist_pair_count(hb2_nlist* nl, uint64_t* n_pairs);
hb2_status hb2_nlist_pair_list(hb2_nlist* nl, uint32_t* pairs, uint64_t cap, uint64_t* n_pairs);
/* Mean / max full neighbours per particle at the last build. */
hb2_status hb2_nlist_stats(hb2_nlist* nl, double* mean_neigh, uint32_t* max_neigh,
                           uint32_t* pitch);

/* hoomd.md.pair.LJ / DPD / plugin pair (hoomd/md/pair/pair.py:163,676); registers its r_cut matrix
 * with the nlist (hoomd/md/PotentialPair.h:369-371). */
hb2_status hb2_pair_create(hb2_pair** out, hb2_system* sys, hb2_nlist* nl, hb2_evaluator ev,
                           hb2_shift_mode mode);
void hb2_pair_destroy(hb2_pair* p);
/* setParams / setRCut / setROn for the unordered type pair (ti,tj) (hoomd/md/PotentialPair.h:467-580).
 * user_params: evaluator's user parameters in declaration order, e.g. LJ {epsilon, sigma}. */
hb2_status hb2_pair_set_params(hb2_pair* p, uint32_t ti, uint32_t tj, const double* user_params,
                               uint32_t n_user);
/* hoomd.md.pair.Table params {r_min, U, F} (hoomd/md/EvaluatorPairTable.h:42-110): `width` samples of
 * U(r) and F(r) = -dU/dr at r_min + i (r_cut - r_min) / width, linearly interpolated; HB2_EVAL_TABLE only. */
hb2_status hb2_pair_set_table(hb2_pair* p, uint32_t ti, uint32_t tj, double r_min, uint32_t width,
                              const double* U, const double* F);
hb2_status hb2_pair_set_r_cut(hb2_pair* p, uint32_t ti, uint32_t tj, double r_cut);
hb2_status hb2_pair_set_r_on(hb2_pair* p, uint32_t ti, uint32_t tj, double r_on);
hb2_status hb2_pair_set_kT(hb2_pair* p, double kT); /* DPD thermostat temperature */
hb2_status hb2_pair_set_threads_per_particle(hb2_pair* p, uint32_t tpp); /* Autotuner<2> knob */
/* ForceCompute::compute(timestep) (hoomd/ForceCompute.cc:562-574): builds the nlist if needed. */
hb2_status hb2_pair_compute(hb2_pair* p, uint64_t timestep);
/* Tag-ordered results (hoomd/md/force.py:77-156): forces N x 3, energies N, virials N x 6 (error
 * HB2_ERR_INVALID if the pressure flag was off), total energy. */
hb2_status hb2_pair_get_forces(hb2_pair* p, double* forces3);
hb2_status hb2_pair_get_energies(hb2_pair* p, double* energies);
hb2_status hb2_pair_get_virials(hb2_pair* p, double* virials6);
hb2_status hb2_pair_get_energy(hb2_pair* p, double* energy); /* global sum (collective) */
/* Local-order results with tags (any output may be NULL): the domain-decomposed counterpart of the
 * three getters above, hb2_system_n entries each. */
hb2_status hb2_pair_get_local(hb2_pair* p, uint32_t* tag, double* forces3, double* energies,
                              double* virials6);
/* hoomd/md/PotentialPair.h:255-341 (LJ only): isotropic tail correction -> additional energy and
 * the diagonal virial term; off by default. */
hb2_status hb2_pair_set_tail_correction(hb2_pair* p, int on);
hb2_status hb2_pair_get_additional(hb2_pair* p, double* energy, double virial6[6]);

/* hoomd.md.Integrator(dt, forces=[pair], methods=[ConstantVolume(filter=All, thermostat)])
 * hoomd/md/integrate.py:267-293, hoomd/md/methods/methods.py:145-175. */
hb2_status hb2_integrator_create(hb2_integrator** out, hb2_system* sys, double dt);
void hb2_integrator_destroy(hb2_integrator* it);
/* Every pair force added contributes to the net force (Integrator::computeNetForce,
 * hoomd/Integrator.cc:166-246); they must share one neighbour list, whose r_cut matrix is the maximum
 * over them (hoomd/md/NeighborList.cc:452-529). */
hb2_status hb2_integrator_add_force(hb2_integrator* it, hb2_pair* p);
/* System::run resets num_builds / shortest_rebuild once per run (hoomd/System.cc:103-110); a caller that
 * splits one run into several hb2_integrator_run calls switches the reset off after the first. */
hb2_status hb2_integrator_set_reset_stats(hb2_integrator* it, int on);
/* thermostat: 0 none (NVE), 1 Bussi(kT, tau) (hoomd/md/Thermostat.h:337-447; tau = 0 is allowed). */
hb2_status hb2_integrator_set_constant_volume(hb2_integrator* it, int thermostat, double kT,
                                              double tau);
/* Bonds. With a domain decomposition every rank passes the same global table.
 * hb2_system_set_bonds: BondData of the snapshot
 * (hoomd/BondedGroupData.h) — `group` holds 2 member TAGS per bond, `typeid_` the bond types; call
 * after the particles are set. hb2_nlist_set_exclusions: exclusions=('bond',) (the reference's default,
 * hoomd/md/nlist.py:87-102; hoomd/md/NeighborList.cc:569-640,1487-1536) drops bonded pairs from the
 * list. hb2_bond_*: hoomd.md.bond.Harmonic / FENEWCA (hoomd/md/PotentialBond.h:176-312,
 * hoomd/md/EvaluatorBondHarmonic.h {k, r0}, hoomd/md/EvaluatorBondFENE.h {k, r0, epsilon, sigma, delta}); hb2_bond_get returns per-particle
 * forces (N x 3), energies (N), virials (N x 6) in tag order and the total energy (any may be null).
 * hb2_integrator_add_bond: the bond forces are added to the pair force (net force) every step. */
hb2_status hb2_system_set_bonds(hb2_system* sys, uint32_t n_bonds, const uint32_t* group,
                                const uint32_t* typeid_, uint32_t n_bond_types);
/* AngleData of the snapshot: `group` holds 3 member TAGS per angle (a, b, c; b is the vertex). The angle
 * kinds of hb2_bond_create take their parameters per angle type. */
hb2_status hb2_system_set_angles(hb2_system* sys, uint32_t n_angles, const uint32_t* group,
                                 const uint32_t* typeid_, uint32_t n_angle_types);
/* DihedralData of the snapshot: 4 member TAGS per dihedral. */
hb2_status hb2_system_set_dihedrals(hb2_system* sys, uint32_t n_dihedrals, const uint32_t* group,
                                    const uint32_t* typeid_, uint32_t n_dihedral_types);
hb2_status hb2_nlist_set_exclusions(hb2_nlist* nl, int exclude_bonds);
/* Further excluded pairs by TAG (what exclusions=('angle', '1-3', ...) add through
 * NeighborList::addExclusion, hoomd/md/NeighborList.cc:881-1120); replaces the previous set. */
hb2_status hb2_nlist_set_exclusion_pairs(hb2_nlist* nl, uint32_t n_pairs, const uint32_t* pairs);
hb2_status hb2_bond_create(hb2_bond** out, hb2_system* sys, hb2_bond_kind kind);
void hb2_bond_destroy(hb2_bond* b);
hb2_status hb2_bond_set_params(hb2_bond* b, uint32_t type, const double* user_params, uint32_t n_user);
hb2_status hb2_bond_get(hb2_bond* b, double* forces3, double* energies, double* virials6,
                        double* energy_total);
hb2_status hb2_integrator_add_bond(hb2_integrator* it, hb2_bond* b);
/* hoomd.md.methods.Langevin(filter=All, kT) (hoomd/md/methods/methods.py, hoomd/md/TwoStepLangevin.cc:
 * 47-82 step one = the NVE half step, :213-292 step two with the BD force): translational degrees of
 * freedom, 3D, one drag coefficient per particle type (`gamma[ntypes]`), no reservoir tally. Replaces
 * the ConstantVolume method of this integrator. */
/* The method's particle group (the ParticleGroup of the method's filter: hoomd/md/IntegrationMethodTwoStep.h,
 * hoomd/filter/ParticleFilter*.h; hoomd.md.methods.*(filter=...) hoomd/md/methods/methods.py:145-175): the
 * tags the filter selects. Particles outside the group are not integrated (they keep position and
 * velocity), the thermostat's kinetic energy and degrees of freedom are the group's (3 n, less 3 only when
 * the group is every particle: hoomd/md/IntegratorTwoStep.cc:154-189). tags == NULL, n == 0: filter.All. */
hb2_status hb2_integrator_set_filter_tags(hb2_integrator* it, const uint32_t* tags, uint64_t n);
/* Kinetic energy of the particles with the given tags (ComputeThermo over a group,
 * hoomd/md/ComputeThermo.cc:139-153); tags == NULL: every particle. */
hb2_status hb2_integrator_kinetic_of_tags(hb2_integrator* it, const uint32_t* tags, uint64_t n,
                                          double* kinetic);
hb2_status hb2_integrator_set_langevin(hb2_integrator* it, double kT, const double* gamma,
                                       uint32_t ntypes);
/* hoomd.md.methods.ConstantPressure(filter=All, S, tauS, couple, thermostat=None|Bussi, box_dof, rescale_all,
 * gamma) (hoomd/md/methods/methods.py:177-580; hoomd/md/TwoStepConstantPressure.cc: integrateStepOne
 * :213-499, integrateStepTwo :583-690, advanceBarostat :860-1000, updatePropagator :716-858; GPU kernels
 * hoomd/md/TwoStepConstantPressureGPU.cu). Martyna-Tobias-Klein barostat on the box lengths (box_dof[0..2];
 * the shear degrees of freedom box_dof[3..5] need a triclinic box and are refused), couplings
 * couple = 0 none, 1 xy, 2 xz, 3 yz, 4 xyz, optional Langevin piston (gamma), thermostat 0 none / 1 Bussi.
 * S[6] is the stress set point in the reference's Voigt order xx yy zz yz xz xy. The barostat state is six
 * fp64 numbers on the host; every step reads the pressure tensor (kinetic tensor + net virial + external
 * virial) once, rescales the box, and the neighbour list's distance check takes the homogeneous dilation
 * out (hoomd/md/NeighborList.cc:1236-1260). Replaces the integrator's method. Under a domain decomposition
 * the sums are all-reduced and every rank keeps the same barostat (collective call, like every run). */
hb2_status hb2_integrator_set_constant_pressure(hb2_integrator* it, const double S[6], double tauS,
                                                int couple, const int box_dof[6], int rescale_all,
                                                double gamma, int thermostat, double kT, double tau);
/* barostat_dof (nu_xx, nu_xy, nu_xz, nu_yy, nu_yz, nu_zz), thermalize_barostat_dof and barostat_energy
 * (TwoStepConstantPressure::getBarostatDOF / setBarostatDOF / thermalizeBarostatDOF / getBarostatEnergy) */
hb2_status hb2_integrator_get_barostat_dof(hb2_integrator* it, double nu[6]);
hb2_status hb2_integrator_set_barostat_dof(hb2_integrator* it, const double nu[6]);
hb2_status hb2_integrator_thermalize_barostat_dof(hb2_integrator* it, uint64_t timestep);
hb2_status hb2_integrator_barostat_energy(hb2_integrator* it, double* energy);
/* Kinetic part of the pressure tensor, sum m v_a v_b (xx, xy, xz, yy, yz, zz; hoomd/md/ComputeThermo.cc:
 * 104-123): with hb2_integrator_thermo's virial sums, ThermodynamicQuantities.pressure_tensor. */
hb2_status hb2_integrator_kinetic_tensor(hb2_integrator* it, double kt[6]);
/* the current global box lengths (State.box after a ConstantPressure run: ParticleData::getGlobalBox) */
hb2_status hb2_system_get_box(hb2_system* s, double L[3]);
/* System::run(nsteps) (hoomd/System.cc:80-181) restricted to this path: prepRun on the first call
 * (hoomd/md/IntegratorTwoStep.cc:226-281), then per step: step one -> nlist check/rebuild -> forces
 * -> step two (step two of a timestep and step one of the next share one launch for ConstantVolume on
 * one GPU). The host reads one word per step — the rebuild flag, copied on a side stream while the
 * force pass it may invalidate is already queued behind step one (the reference blocks on the same flag,
 * hoomd/md/NeighborListGPU.cc:78-82) — and synchronises fully only on rebuild steps (capacity checks) and
 * at the end of the run. */
hb2_status hb2_integrator_run(hb2_integrator* it, uint64_t nsteps);
/* ComputeThermo quantities of the current state (hoomd/md/ComputeThermo.cc:76-…): kinetic energy,
 * potential energy, and the 6 virial sums (virial valid only with the pressure flag). */
hb2_status hb2_integrator_thermo(hb2_integrator* it, double* kinetic, double* potential,
                                 double virial6[6]);
/* Device time (ms, CUDA events) spent in the last hb2_integrator_run, and per-phase kernel counts. */
hb2_status hb2_integrator_last_run_ms(hb2_integrator* it, float* ms);
/* Per-phase device timing of the following runs (CUDA events on the launching stream, the
 * counterpart of the reference's Autotuner event sampling, hoomd/Autotuner.h:509-511): phase 0 =
 * integrate (step one + thermostat + step two), 1 = neighbour list (binning, sort, reorder, build and
 * the rebuild-flag read-back), 2 = pair-force kernel. phase_n = steps, builds, force launches.
 * Profiling adds five event records per step; leave it off for throughput measurements. */
hb2_status hb2_integrator_set_profiling(hb2_integrator* it, int on);
hb2_status hb2_integrator_get_profile(hb2_integrator* it, double phase_ms[3], uint64_t phase_n[3]);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HOOMD_B200_H */
