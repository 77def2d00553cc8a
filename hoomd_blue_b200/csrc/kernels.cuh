// kernels.cuh — internal launcher interface shared by the C-ABI layers (capi.cu, engine.cu).
#pragma once
#include "common.cuh"
#include "comm.cuh"
#include "tile.cuh"

namespace hb2
    {
// ---- binning.cu
size_t cell_order_tmp_bytes(uint32_t N, uint32_t Nc);
// N particles of which the first N_owned are owned (the rest are ghosts of a domain decomposition);
// geom.n = {1,1,1} is the single-GPU periodic grid. d_cell_start / d_cell_end: Nc entries each.
hb2_status compute_cell_order(const float4* d_pos, uint32_t N, uint32_t N_owned, const hb2_box& box,
                              const DomainGeom& geom, uint32_t* d_order, uint32_t* d_cell_start,
                              uint32_t* d_cell_end, uint32_t* d_conditions, void* d_tmp,
                              size_t tmp_bytes, cudaStream_t stream, int curve_request = -1);
// curve_request: -1 the engine's choice (HB2_CELL_ORDER, default a Morton curve over the cells), 0 cells
// in row-major order, 1 Morton
hb2_status reorder_particles(const uint32_t* d_order, uint32_t N, const float4* pos_in,
                             const float4* vel_in, const int* img_in, const uint32_t* tag_in,
                             const float4* force_in, float4* pos_out, float4* vel_out, int* img_out,
                             uint32_t* tag_out, float4* force_out, uint32_t* rtag,
                             cudaStream_t stream);

// how far past a face of its domain (in units of the domain width) a particle handed to snapshot_pack
// may lie and still count as this rank's own
struct OwnMargin
    {
    double frac[3] = {0.0, 0.0, 0.0};
    };
// snapshot (raw arrays, all DEVICE pointers) <-> engine layout. flags[0] <- index+1 of a particle with
// a bad type id, flags[1] <- index+1 of a particle this rank does not own (domain decompositions).
hb2_status snapshot_pack(uint32_t N, const float* pos3, const float* vel3, const float* mass,
                         const uint32_t* typeid_, const int* image3, const uint32_t* tag_in,
                         float4* pos, float4* vel, int* image, uint32_t* tag, uint32_t* rtag,
                         uint32_t ntypes, const hb2_box& box, const DomainGeom& geom,
                         const OwnMargin& margin, uint32_t* flags, cudaStream_t stream);
hb2_status snapshot_unpack(uint32_t N, const float4* pos, const float4* vel, const int* image,
                           const uint32_t* tag, float* pos3, float* vel3, float* mass,
                           uint32_t* typeid_, int* image3, uint32_t* tag_out, bool by_tag,
                           cudaStream_t stream);

// ---- nlist.cu
struct NlistBuildArgs
    {
    const float4* pos;          // cell-ordered positions, N_local first-class rows
    uint32_t N;                 // particles that get a row
    DomainGeom geom;            // cell grid (dimt) and decomposition (n = 1: periodic ring)
    const uint32_t* cell_start; // Nc
    const uint32_t* cell_end;   // Nc
    const double* r_listsq;     // ntypes^2 doubles (device), tj*ntypes+ti
    uint32_t ntypes;
    double r_listsq_single;     // copy of r_listsq[0] for the one-type fast path
    double r_list_min;          // smallest positive r_list (sets the fp32 guard band)
    float Lmax;                 // longest box edge (sets the fp32 guard band)
    uint32_t* nlist;
    uint32_t pitch;
    uint32_t* n_neigh;
    float4* last_pos;           // may be null
    uint32_t* conditions;       // [0] <- max n_neigh
    uint32_t dummy_index;       // index of the far-away padding particle (pos[N_total])
    };
hb2_status launch_nlist_build(const NlistBuildArgs& a, const hb2_box& box, cudaStream_t stream);

// ---- bond.cu
struct BondArgs
    {
    const float4* pos;
    const uint32_t* tag;
    const uint32_t* rtag;
    uint32_t N, Ntot;
    const uint32_t* head;    // bond table by tag (CSR), partner tags, bond types
    const uint32_t* partner;
    const uint32_t* btype;
    const float* params;     // HB2_BOND_PARAM_FLOATS floats per bond type
    float4* force;
    float* virial;           // null: not computed
    size_t virial_pitch;
    const uint32_t* skip_flag;
    uint32_t* err;           // [0] <- 1 + tag of a partner that is not on this device,
                             // [1] <- 1 + tag of a particle with a bond out of bounds
    };
hb2_status launch_bond_forces(hb2_bond_kind kind, const BondArgs& a, bool accumulate, const hb2_box& box,
                              cudaStream_t stream);
// angles: `members` = (tag a, tag b, tag c, type | role << 30) per table entry; a.head indexes it by tag
hb2_status launch_angle_forces(hb2_bond_kind kind, const BondArgs& a, const uint4* members, bool accumulate,
                               const hb2_box& box, cudaStream_t stream);
// dihedrals: members = the four tags, type_role = type | role << 30 per table entry
hb2_status launch_dihedral_forces(hb2_bond_kind kind, const BondArgs& a, const uint4* members,
                                  const uint32_t* type_role, bool accumulate, const hb2_box& box,
                                  cudaStream_t stream);
hb2_status launch_nlist_filter_bonds(uint32_t N, const uint32_t* tag, const uint32_t* rtag,
                                     const uint32_t* head, const uint32_t* partner, uint32_t* nlist,
                                     uint32_t pitch, uint32_t* n_neigh, uint32_t dummy_index,
                                     cudaStream_t stream);

// ---- pair.cu
struct PairArgs
    {
    const float4* pos;
    const float4* vel;   // DPD only
    const uint32_t* tag; // DPD only
    uint32_t N;
    const uint32_t* n_neigh;
    const uint32_t* nlist;
    uint32_t pitch;
    const float* params; // ntypes^2 x HB2_PARAM_FLOATS floats (packed by the evaluator's param_type)
    const float* rcutsq; // ntypes^2
    const float* ronsq;  // ntypes^2
    // r_cut^2 in fp64 (hoomd/md/PotentialPair.h:498 computes it in Scalar = double): decides pairs
    // whose fp32 rsq lies within rounding of the cut-off exactly as the reference does. Null: the
    // float values are widened (layer-1 callers that only have floats).
    const double* rcutsq_d = nullptr;
    double rcutsq0_d = 0.0;
    // host copies of entry [0] for the one-type fast path (passed in the constant bank)
    float params0[HB2_PARAM_FLOATS];
    float rcutsq0, ronsq0;
    uint32_t ntypes;
    float4* force;
    float* virial; // 6 x virial_pitch or null
    size_t virial_pitch;
    uint32_t seed;
    const uint64_t* d_timestep; // device timestep (graph-replay safe) or null -> timestep
    uint64_t timestep;
    float dt, kT;
    float face_margin; // >= r_cut_max: farther than this from every face, no pair inside the cut-off crosses a face
    uint32_t tpp;
    bool energy = true; // false: skip the per-particle potential energy (force.w = 0)
    const uint32_t* skip_flag = nullptr; // device word; non-zero => the launch does nothing
    // tile-shared list (tile.cuh): when rows16 is set the pass runs pair_force_tile_kernel over it and
    // n_neigh; nlist / tpp / face_margin are not used
    const uint16_t* rows16 = nullptr;
    TileGeom tg;
    DomainGeom geom;
    const uint32_t* cell_start = nullptr;
    const uint32_t* cell_end = nullptr;
    uint32_t stage_cap = 0;
    const unsigned char* tile_desc = nullptr; // per-tile descriptors of the last list build
    uint32_t* tile_ticket = nullptr; // two zeroed words (the persistent tile kernel's work ticket)
    };
hb2_status launch_pair_forces(hb2_evaluator ev, hb2_shift_mode mode, const PairArgs& a,
                              const hb2_box& box, cudaStream_t stream);

// ---- integrate.cu
// The particles an integration method acts on (hoomd ParticleGroup of the method's filter,
// hoomd/md/IntegrationMethodTwoStep.h): a byte per particle TAG; member == nullptr: every particle.
struct GroupMask
    {
    const uint8_t* member = nullptr;
    const uint32_t* tag = nullptr;
    };
hb2_status launch_step_one(float4* pos, float4* vel, const float4* force, int* image, uint32_t N,
                           const hb2_box& box, float dt, const double* d_rescale,
                           const float4* last_pos, float maxshift, uint32_t* d_flag,
                           cudaStream_t stream, GroupMask group = GroupMask());
// Bussi factor of the next step, drawn by the thread that completes step two's KE sum
struct BussiNext
    {
    double* rescale = nullptr; // null: off
    double dof = 0, dt = 0, kT = 0, tau = 0;
    uint64_t timestep = 0; // the step whose step-one will apply the factor
    uint32_t seed = 0, first_tag = 0;
    };
// d_ke_partial: >= step_two_blocks(N) doubles of scratch; d_ke_out: 1 double (deterministic sum).
uint32_t step_two_blocks(uint32_t N);
hb2_status launch_step_two(float4* vel, const float4* force, uint32_t N, float dt,
                           const double* d_rescale, double* d_ke_partial, uint32_t* d_ticket,
                           double* d_ke_out, cudaStream_t stream, const BussiNext* bussi = nullptr,
                           GroupMask group = GroupMask());
// step two of timestep t fused with step one of timestep t + 1 (integrate.cu); bussi with rescale set:
// persistent kernel with the grid-wide KE sum and the factor of step t + 1 between the halves
hb2_status launch_step_two_one(float4* pos, float4* vel, const float4* force, int* image, uint32_t N,
                               const hb2_box& box, float dt, const float4* last_pos, float maxshift,
                               uint32_t* d_flag, double* d_ke_partial, uint32_t* d_ticket, double* d_ke_out,
                               const BussiNext* bussi, uint32_t* d_gen, uint32_t epoch, cudaStream_t stream,
                               bool* fused = nullptr, // *fused = false: not launched (no co-resident grid); use the two kernels
                               GroupMask group = GroupMask());
// TwoStepLangevin step two; d_gamma_coeff = (gamma, sqrt(6 gamma kT / dt)) per type; total <- F + F_bd
// (what the next step one integrates with)
hb2_status launch_step_two_langevin(float4* vel, const float4* force, float4* total, const float4* pos,
                                    const uint32_t* tag, uint32_t N, float dt, const float* d_gamma_coeff,
                                    uint32_t seed, uint64_t timestep, cudaStream_t stream,
                                    const uint8_t* member = nullptr);
// ConstantPressure half steps (integrate.cu): the host advances the barostat and hands the propagators
constexpr int NPT_MAX_TYPES = 8;
struct NptFactors
    {
    float dt = 0.f;
    float mtk = 1.f;                          // exp(-dt/2 (nu_xx + nu_yy + nu_zz) / ndof)
    float exp_v[3] = {1.f, 1.f, 1.f};         // velocity propagator, diagonal
    float exp_r[3] = {1.f, 1.f, 1.f};         // position (and box) propagator, diagonal
    float exp_r_int[3] = {0.f, 0.f, 0.f};     // integrated position propagator (multiplies v), diagonal
    float lambda[3] = {1.f, 1.f, 1.f};        // L_now / L_at_last_build
    float maxsq[NPT_MAX_TYPES] = {0, 0, 0, 0, 0, 0, 0, 0}; // distance-check threshold per type
    int rescale_all = 0;
    };
hb2_status launch_step_one_npt(float4* pos, float4* vel, const float4* force, int* image, uint32_t N,
                               const hb2_box& box, const NptFactors& k, const double* d_rescale,
                               const float4* last_pos, uint32_t* d_flag, cudaStream_t stream,
                               GroupMask group = GroupMask());
// k.dt = 0: only the kinetic tensor of the current velocities
hb2_status launch_step_two_npt(float4* vel, const float4* force, uint32_t N, const NptFactors& k,
                               double* d_partial, uint32_t* d_ticket, double* d_out6, cudaStream_t stream,
                               GroupMask group = GroupMask());
hb2_status launch_kinetic(const float4* vel, uint32_t N, double* d_partial, uint32_t* d_ticket,
                          double* d_ke_out, cudaStream_t stream, GroupMask group = GroupMask());
// Bussi factor on device (hoomd/md/Thermostat.h:357-447): *d_rescale <- factor for the step-one of
// timestep *d_timestep; reads *d_ke.
hb2_status launch_bussi(const double* d_ke, double dof, double dt, double kT, double tau,
                        uint64_t timestep, uint32_t seed, uint32_t first_tag, double* d_rescale,
                        cudaStream_t stream);
hb2_status launch_advance_timestep(uint64_t* d_timestep, cudaStream_t stream);
// dst += src over N particles (force float4 and, when both virial arrays are given, the six virial rows);
// skip_flag as in PairArgs
hb2_status launch_accumulate_force(float4* dst, float* dst_virial, size_t dst_pitch, const float4* src,
                                   const float* src_virial, size_t src_pitch, uint32_t N,
                                   const uint32_t* skip_flag, cudaStream_t stream);
// sum of force.w (potential energy), deterministic; and 6 virial sums
hb2_status launch_reduce_energy(const float4* force, const float* virial, size_t virial_pitch,
                                uint32_t N, double* d_partial, uint32_t* d_ticket, double* d_out7,
                                cudaStream_t stream);
uint32_t reduce_blocks(uint32_t N);

void rng_normals_host(uint8_t id, uint64_t timestep, uint32_t seed, uint32_t first_tag, int n, double* out);
double bussi_factor_host(double K, double dof, double dt, double kT, double tau, uint64_t timestep,
                         uint32_t seed, uint32_t first_tag);
    } // namespace hb2
