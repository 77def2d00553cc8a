// engine.cuh — the engine's host-side objects (shared by engine.cu and comm.cu). See engine.cu for
// the reference classes each one mirrors.
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "comm.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace hb2
    {
template<class T> inline hb2_status dev_alloc(T** p, size_t n)
    {
    *p = nullptr;
    if (n == 0)
        n = 1;
    HB2_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    return HB2_OK;
    }
template<class T> inline void dev_free(T*& p)
    {
    if (p)
        cudaFree(p);
    p = nullptr;
    }
    } // namespace hb2

using namespace hb2;

#define HB2_TRY(expr)              \
    do                             \
        {                          \
        hb2_status _s = (expr);    \
        if (_s != HB2_OK)          \
            return _s;             \
        } while (0)

// =================================================================================================
struct hb2_system
    {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t ntypes = 1;
    uint32_t seed = 0;
    hb2_box box;
    bool has_box = false;
    uint32_t N = 0;   // owned particles, indices [0, N)
    uint32_t Ng = 0;  // ghost particles of the domain decomposition, indices [N, N + Ng)
    uint32_t cap = 0; // capacity of every per-particle array (N + Ng + 1 <= cap; the slot after the
                      // last ghost holds the far-away dummy that pads neighbour-list rows)
    Domain dom;
    uint64_t timestep = 0;
    uint64_t* d_timestep = nullptr;
    bool compute_virial = false;

    // particle arrays, double buffered: every nlist rebuild gathers them into cell order
    float4* pos[2] = {nullptr, nullptr};
    float4* vel[2] = {nullptr, nullptr};
    int* image[2] = {nullptr, nullptr};
    uint32_t* tag[2] = {nullptr, nullptr};
    uint32_t* rtag = nullptr;
    // snapshot staging: the caller's raw tag-ordered arrays as they cross PCIe (11 words per particle)
    // and the pinned bounce buffer behind them
    float* d_stage = nullptr;
    float* h_stage = nullptr;
    size_t stage_cap = 0;
    int cur = 0;
    uint64_t order_epoch = 0; // bumped whenever the particle order changes

    // reduction scratch
    double* d_partial = nullptr;
    uint32_t* d_ticket = nullptr;
    double* d_sums = nullptr; // [0] KE, [1..7] PE + virial, [8] rescale factor

    // bond table keyed by particle tag (CSR): partners' tags and bond types (csrc/bond.cu)
    uint32_t n_bonds = 0, n_bond_types = 0;
    uint32_t* d_bond_head = nullptr;
    uint32_t* d_bond_partner = nullptr;
    uint32_t* d_bond_type = nullptr;
    uint64_t bond_epoch = 0; // bumped by hb2_system_set_bonds
    // angle table keyed by particle tag: (tag a, tag b, tag c, type | role << 30) per membership
    uint32_t n_angles = 0, n_angle_types = 0;
    uint32_t* d_angle_head = nullptr;
    uint4* d_angle_members = nullptr;
    // dihedral table keyed by particle tag: four member tags + (type | role << 30) per membership
    uint32_t n_dihedrals = 0, n_dihedral_types = 0;
    uint32_t* d_dihedral_head = nullptr;
    uint4* d_dihedral_members = nullptr;
    uint32_t* d_dihedral_tr = nullptr;

    std::vector<hb2_nlist*> nlists;
    std::vector<hb2_pair*> pairs;
    cudaEvent_t mid_event = nullptr; // recorded between the nlist update and the force launch
    // the rebuild flag travels to the host on a side stream while the speculative force pass runs
    cudaStream_t side_stream = nullptr;
    cudaEvent_t flag_ready = nullptr, flag_copied = nullptr;

    hb2_status set_timestep(uint64_t t)
        {
        timestep = t;
        HB2_CUDA(cudaMemcpyAsync(d_timestep, &timestep, sizeof(uint64_t), cudaMemcpyHostToDevice,
                                 stream));
        HB2_CUDA(cudaStreamSynchronize(stream));
        return HB2_OK;
        }
    };

struct hb2_nlist
    {
    hb2_system* sys = nullptr;
    double buffer = 0.4;
    uint32_t rebuild_check_delay = 1;
    bool check_dist = true;

    std::vector<hb2_pair*> consumers;
    std::vector<double> r_cut, r_listsq, rcut_max;
    std::vector<double> base_r_cut; // the list's own r_cut matrix (empty: all zero)
    double rcut_max_max = 0.0, r_list_min = 0.0;
    bool params_dirty = true;

    DomainGeom geom;  // cell grid of the last build (dimi = dimensions reported to Python)
    uint32_t cell_cap = 0;
    uint32_t* d_cell_start = nullptr;
    uint32_t* d_cell_end = nullptr;
    uint32_t* d_order = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    double* d_r_listsq = nullptr;

    // Two layouts of the same full list. The tile-shared layout (tile.cuh: 16-bit staged slots, built by
    // nlist_tile.cu, consumed by pair_tile.cuh) is the step loop's; the per-particle rows of global
    // indices (nlist.cu) serve the thermostatted evaluators, the exclusion filter, the exported arrays and
    // grids the tile path does not cover. With the tile layout active the global rows are produced on
    // demand from it (global_rows()).
    bool tile_enabled = true; // HB2_TILE=0 switches the tile path off
    bool tile_active = false; // the last build produced rows16
    bool global_valid = false; // d_nlist matches the last build
    uint16_t* d_rows16 = nullptr;
    unsigned char* d_tile_desc = nullptr; // per-tile descriptors (tile.cuh), rewritten by every build
    size_t tile_desc_cap = 0;
    uint32_t* d_tile_ticket = nullptr; // work ticket of the persistent tile force kernel (pair_tile.cuh)
    size_t rows16_cap = 0;
    uint32_t stage_cap = 0; // staged slots per tile (0: not yet estimated)
    TileGeom tg;
    uint32_t* d_nlist = nullptr;
    uint32_t pitch = 0;
    size_t nlist_cap = 0;
    uint32_t* d_n_neigh = nullptr;
    float4* d_last_pos = nullptr;
    uint32_t cap_N = 0;
    uint32_t* d_conditions = nullptr; // [0] max n_neigh, [1] nan idx+1, [2] oob idx+1, [3] rebuild flag,
                                      // [4] / [5] tile build: staged slots / candidates beyond capacity
    uint32_t* h_conditions = nullptr; // pinned mirror

    bool exclude_bonds = true; // exclusions=('bond',), the reference's default (hoomd/md/nlist.py:87-102)
    // further excluded pairs ('angle', '1-3', ...): CSR keyed by particle tag, like the bond table
    uint32_t n_ex_pairs = 0;
    uint32_t* d_ex_head = nullptr;
    uint32_t* d_ex_partner = nullptr;
    bool force_update = true;
    bool has_list = false;
    uint64_t num_builds = 0;
    uint64_t last_update_step = 0;
    uint64_t shortest_rebuild = ~0ull;
    uint32_t max_neigh = 0;

    hb2_status update_r_list();
    hb2_status ensure_capacity();
    hb2_status rebuild(uint64_t timestep);
    hb2_status compute(uint64_t timestep, bool flag_known, bool flag_value);
    bool want_tiles() const;
    TileBuildArgs tile_args() const;
    hb2_status global_rows(); // make d_nlist valid for the current build
    double last_L[3] = {0, 0, 0}; // box lengths at the last build (hoomd/md/NeighborList.cc m_last_L)
    float maxshift() const
        {
        // hoomd/md/NeighborList.cc:1244-1260 with a fixed box (lambda = 1): (r_max - r_min)/2
        return (float)(buffer / 2.0);
        }
    };

struct hb2_pair
    {
    hb2_system* sys = nullptr;
    hb2_nlist* nl = nullptr;
    hb2_evaluator ev = HB2_EVAL_LJ;
    hb2_shift_mode mode = HB2_SHIFT_NONE;
    std::vector<double> user; // ntypes^2 x 2 user parameters
    std::vector<uint8_t> have_params;
    std::vector<double> r_cut, r_on;
    std::vector<float> packed, rcutsq, ronsq;
    float* d_params = nullptr;
    float* d_rcutsq = nullptr;
    float* d_ronsq = nullptr;
    std::vector<double> rcutsq_d;
    double* d_rcutsq_d = nullptr;
    // Table evaluator: per type pair V[width] then F[width], and r_min
    std::vector<std::vector<float>> tables;
    std::vector<double> table_rmin;
    float* d_tables = nullptr;
    size_t table_cap = 0;
    bool params_dirty = true;
    double kT = 1.0;
    double dt = 0.005;
    uint32_t tpp = 8;
    bool tail_correction = false;

    float4* d_force = nullptr;
    float* d_virial = nullptr;
    uint32_t cap_N = 0;
    bool virial_valid = false;
    bool energy_valid = false; // the last force pass also accumulated per-particle energies
    bool energy_every_step = false;
    bool computed = false;
    bool extra_added = false; // the integrator added the bond forces into d_force (net force)
    uint64_t computed_step = 0, computed_epoch = 0;

    hb2_status upload_params();
    hb2_status compute(uint64_t timestep, bool flag_known = false, bool flag_value = false,
                       bool want_energy = true);
    hb2_status launch(uint64_t timestep, bool want_energy, const uint32_t* skip_flag);
    };

// PotentialBond<EvaluatorBondHarmonic> (hoomd/md/PotentialBond.h): parameters per bond type, own
// force / virial arrays for its getters; inside the step loop its forces are added to the pair
// force's array (the net force).
struct hb2_bond
    {
    hb2_system* sys = nullptr;
    hb2_bond_kind kind = HB2_BOND_HARMONIC;
    std::vector<float> packed; // HB2_BOND_PARAM_FLOATS per bond type
    std::vector<char> have;
    bool params_dirty = true;
    float* d_params = nullptr;
    float4* d_force = nullptr;
    float* d_virial = nullptr;
    uint32_t cap_N = 0;
    uint32_t* d_err = nullptr;
    hb2_status upload_params();
    hb2_status launch(float4* force, float* virial, size_t virial_pitch, bool accumulate,
                      const uint32_t* skip_flag);
    hb2_status compute_own();
    hb2_status check_complete();
    };

struct hb2_integrator
    {
    hb2_system* sys = nullptr;
    double dt = 0.005;
    std::vector<hb2_pair*> forces;
    std::vector<hb2_bond*> bonds;
    int thermostat = 0; // 0 NVE, 1 ConstantVolume + Bussi, 2 Langevin method
    double kT = 1.0, tau = 0.0;
    float* d_gamma_coeff = nullptr; // Langevin: (gamma, sqrt(6 gamma kT / dt)) per type
    float4* d_total = nullptr;      // Langevin: F + F_bd of the last step two (next step one's force)
    size_t total_cap = 0;
    bool total_valid = false;
    uint8_t* d_member = nullptr; // the method's group: a byte per particle tag (null: every particle)
    uint64_t n_member = 0;       // group size (global)
    uint32_t* d_gen = nullptr; // generation word of the fused step-two/step-one kernel (integrate.cu)
    uint32_t gen_epoch = 0;
    uint64_t total_epoch = 0, total_step = 0; // particle order and timestep d_total belongs to
    // ConstantPressure (hoomd/md/TwoStepConstantPressure.h): barostat state, host side, fp64
    struct Npt
        {
        bool on = false;
        double S[6] = {0, 0, 0, 0, 0, 0}; // set point, Voigt order xx yy zz yz xz xy (the reference's m_S)
        double tauS = 1.0, gamma = 0.0;
        int couple = 0;                   // 0 none, 1 xy, 2 xz, 3 yz, 4 xyz
        int flags = 7;                    // bit 0 x, bit 1 y, bit 2 z
        bool rescale_all = false;
        double nu[6] = {0, 0, 0, 0, 0, 0}; // nu_xx, nu_xy, nu_xz, nu_yy, nu_yz, nu_zz
        double ndof = 0.0;                 // set by step one, as the reference's m_ndof
        double P[6] = {0, 0, 0, 0, 0, 0};  // pressure tensor xx xy xz yy yz zz of the last thermo
        double ke = 0.0;
        double exp_v[3] = {1, 1, 1}, exp_r[3] = {1, 1, 1}, exp_r_int[3] = {0, 0, 0};
        std::vector<uint32_t> n_per_type; // particles per type (the tail correction's densities)
        bool tail_valid = false;          // some force carries a tail correction
        double* h_thermo = nullptr; // pinned: 7 energy/virial sums + 6 kinetic sums
        } npt;
    bool prepared = false;
    bool reset_stats = true; // reset the neighbour list's num_builds / shortest_rebuild at the start of run
    uint64_t prepared_epoch = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    // optional per-phase device timing (CUDA events on the launching stream)
    bool profiling = false;
    std::vector<cudaEvent_t> pool;
    size_t pool_used = 0;
    std::vector<size_t> marks; // five events per step: a b c d e
    double phase_ms[3] = {0, 0, 0}; // integrate, nlist, force
    uint64_t phase_n[3] = {0, 0, 0};
    cudaEvent_t next_event()
        {
        if (pool_used == pool.size())
            {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
            }
        return pool[pool_used++];
        }
    };

