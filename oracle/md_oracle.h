/* md_oracle.h — CPU restatement (fp64, FMA-free) of HOOMD-blue v4.8.2's short-range MD pair-force
 * step.  TEST INFRASTRUCTURE ONLY: nothing under hoomd_blue_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 * Compile with -O2 -ffp-contract=off (see oracle/Makefile): the reference CPU build is x86-64
 * baseline (SSE2 doubles, no FMA), and bit-exact neighbour-pair membership depends on that.
 *
 * Scope: orthorhombic boxes (tilt factors xy = xz = yz = 0), 3D. With zero tilt every tilt term of
 * hoomd/BoxDim.h multiplies by 0 and adds/subtracts 0, so the restated formulas produce identical
 * doubles.
 */
#ifndef MD_ORACLE_H
#define MD_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
    {
    double L[3], lo[3], hi[3], Linv[3];
    int periodic[3];
    } orc_box;

/* hoomd/BoxDim.h:183-200 (setL) */
void orc_box_set(orc_box* b, double Lx, double Ly, double Lz);
/* hoomd/BoxDim.h:276-284 */
void orc_make_fraction(const orc_box* b, const double p[3], const double ghost_width[3], double f[3]);
/* hoomd/BoxDim.h:343-389 (CPU branch form) */
void orc_min_image(const orc_box* b, double w[3]);
/* hoomd/BoxDim.h:407-474 */
void orc_wrap(const orc_box* b, double w[3], int img[3]);

/* hoomd/CellList.cc:83-131 */
void orc_cell_dims(const orc_box* b, double nominal_width, const double ghost_width[3], uint32_t dim[3]);
/* hoomd/CellList.cc:346-402 + :252-273 (row width). Returns the row width n_adj; adj must hold
 * n_adj * Nc entries laid out adj[cell * n_adj + k] (hoomd/Index1D.h Index2D(w,h)(i,j)=j*w+i). */
uint32_t orc_cell_adj_width(const uint32_t dim[3]);
void orc_cell_adj(const uint32_t dim[3], uint32_t* adj);
/* hoomd/CellList.cc:404-547. pos is N x 3. cell_size[Nc]; cell_idx[Nc*Nmax] (particle index per
 * slot, cli(offset,bin) = bin*Nmax + offset). Returns 0 ok; conditions[3] = {overflow size, nan idx+1,
 * out-of-box idx+1} like hoomd/CellList.cc:549-607. */
int orc_cell_list(const orc_box* b, const uint32_t dim[3], const double ghost_width[3],
                  const double* pos, uint32_t n_local, uint32_t n_total, uint32_t Nmax,
                  uint32_t* cell_size, uint32_t* cell_idx, uint32_t conditions[3]);

/* hoomd/md/NeighborList.cc:452-529: r_listsq[tj*nt+ti] = (r_cut>0 ? r_cut+buff : 0)^2 */
void orc_r_listsq(const double* r_cut, uint32_t ntypes, double r_buff, double* r_listsq,
                  double* rcut_max_per_type, double* rcut_max_max);

/* hoomd/md/NeighborListBinned.cc:36-179 (+ the overflow/retry loop of
 * hoomd/md/NeighborList.cc:411-427 folded into a count pass). full != 0 selects storageMode full.
 * Returns a malloc'ed nlist (caller frees with orc_free); head[i] is the exclusive scan of n_neigh.
 * n_local particles get lists; candidates run over n_total (ghosts appended). */
uint32_t* orc_nlist_binned(const orc_box* b, const double* pos, const uint32_t* type,
                           uint32_t n_local, uint32_t n_total, const double* r_cut, uint32_t ntypes,
                           double r_buff, int full, const double ghost_width[3],
                           uint32_t* n_neigh, uint64_t* head, uint64_t* total, uint32_t dim_out[3]);
/* O(N^2) statement of the same membership rule (no cells) — checks the binned walk itself. */
uint32_t* orc_nlist_brute(const orc_box* b, const double* pos, const uint32_t* type, uint32_t n_local,
                          uint32_t n_total, const double* r_cut, uint32_t ntypes, double r_buff,
                          int full, uint32_t* n_neigh, uint64_t* head, uint64_t* total);
void orc_free(void* p);

/* hoomd/md/NeighborList.cc:1220-1289 with a fixed box (lambda = 1). Returns 1 if a rebuild is due. */
int orc_distance_check(const orc_box* b, const double* pos, const double* last_pos,
                       const uint32_t* type, uint32_t n, const double* rcut_max_per_type,
                       double r_buff);

/* Pair potentials. shift_mode: 0 none, 1 shift, 2 xplor (hoomd/md/PotentialPair.h:104-109).
 * params: LJ -> {epsilon, sigma} per type pair [tj*nt+ti] (hoomd/md/EvaluatorPairLJ.h:119-140)
 *         DPD -> {A, gamma}                          (hoomd/md/EvaluatorPairDPDThermoDPD.h:80-120)
 * force is N x 4 (fx,fy,fz,pe), virial is 6 x N (xx,xy,xz,yy,yz,zz; pitch N).
 * hoomd/md/PotentialPair.h:587-794 */
void orc_pair_lj(const orc_box* b, const double* pos, const uint32_t* type, uint32_t n_local,
                 const uint32_t* n_neigh, const uint64_t* head, const uint32_t* nlist, int half,
                 const double* params, const double* r_cut, const double* r_on, uint32_t ntypes,
                 int shift_mode, int compute_virial, double* force, double* virial);
/* hoomd/md/PotentialPairDPDThermo.h:102-268 + hoomd/md/EvaluatorPairDPDThermoDPD.h:222-277 */
void orc_pair_dpd(const orc_box* b, const double* pos, const double* vel, const uint32_t* type,
                  const uint32_t* tag, uint32_t n_local, const uint32_t* n_neigh,
                  const uint64_t* head, const uint32_t* nlist, int half, const double* params,
                  const double* r_cut, uint32_t ntypes, uint16_t seed, uint64_t timestep, double dt,
                  double kT, double* force, double* virial);
/* single-pair evaluators for the golden tables (forces_and_energies.json) */
int orc_eval_lj(double rsq, double rcutsq, double epsilon, double sigma, int energy_shift,
                double* force_divr, double* pair_eng);
int orc_eval_dpd_conservative(double rsq, double rcutsq, double A, double* force_divr,
                              double* pair_eng);
/* hoomd/md/PotentialPair.h:255-341 (isotropic LJ tail correction; returns energy, writes the
 * pressure-virial diagonal term) */
double orc_lj_tail_correction(const double* params, const double* r_cut, uint32_t ntypes,
                              const uint32_t* n_per_type, double volume, double* virial_diag);

/* Exclusions from bonds and the filter pass (hoomd/md/NeighborList.cc:569-640 addExclusionsFromBonds,
 * :1448-1485 updateExListIdx, :1487-1536 filterNlist): every bonded pair (tags a, b) is removed from
 * both rows; rows are compacted in place (head stays, n_neigh shrinks). rtag[tag] = index. */
void orc_filter_bond_exclusions(uint32_t n, uint32_t* n_neigh, const uint64_t* head, uint32_t* nlist,
                                const uint32_t* tag, const uint32_t* rtag, const uint32_t* bond_group,
                                uint32_t n_bonds);

/* PotentialBond<EvaluatorBondHarmonic> (hoomd/md/PotentialBond.h:176-312,
 * hoomd/md/EvaluatorBondHarmonic.h:104-124): V = k/2 (r - r0)^2, half the energy and half the virial
 * to each member. params = (k, r0) per bond type; bond_group holds member TAGS. force N x 4 and virial
 * 6 x N are overwritten. Returns 0, or 1 + the index of an incomplete bond. */
int orc_bond_harmonic(const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
                      const uint32_t* bond_group, const uint32_t* bond_type, uint32_t n_bonds,
                      const double* params, double* force, double* virial);

/* PotentialBond<EvaluatorBondFENE> (hoomd.md.bond.FENEWCA, hoomd/md/EvaluatorBondFENE.h): params =
 * (k, r0, epsilon, sigma, delta) per bond type. Returns 0, 1 + index of an incomplete bond, or
 * -1 - index of a bond beyond r0 ("bond out of bounds"). */
int orc_bond_fenewca(const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
                     const uint32_t* bond_group, const uint32_t* bond_type, uint32_t n_bonds,
                     const double* params, double* force, double* virial);

/* Angle forces: kind 0 harmonic (hoomd/md/HarmonicAngleForceCompute.cc), 1 cosine squared
 * (hoomd/md/CosineSqAngleForceCompute.cc); params = (k, t0) per angle type; a third of the energy and
 * of the virial to each member. Returns 0 or 1 + the index of an incomplete angle. */
int orc_angle(int kind, const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
              const uint32_t* angle_group, const uint32_t* angle_type, uint32_t n_angles,
              const double* params, double* force, double* virial);

/* hoomd.md.dihedral.Periodic (hoomd/md/HarmonicDihedralForceCompute.cc): params = (k, d, n, phi0) per
 * dihedral type; group = 4 member tags. Returns 0 or 1 + the index of an incomplete dihedral. */
int orc_dihedral_periodic(const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
                          const uint32_t* group, const uint32_t* type, uint32_t n_dihedrals,
                          const double* params, double* force, double* virial);

/* kind 0: Periodic as above; kind 1: hoomd.md.dihedral.OPLS (hoomd/md/OPLSDihedralForceCompute.cc),
 * params = (k1, k2, k3, k4) per dihedral type */
int orc_dihedral(int kind, const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
                 const uint32_t* group, const uint32_t* type, uint32_t n_dihedrals, const double* params,
                 double* force, double* virial);

/* RNG: Philox4x32-10 (Salmon et al., SC'11; hoomd/extern/random123 commit a2c4f04) and HOOMD's
 * Seed/Counter packing hoomd/RandomNumbers.h:136-199 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_seed_key(uint8_t id, uint64_t timestep, uint16_t seed, uint32_t key[2]);
/* uniform in [-1,1] for the DPD pair (tag_i, tag_j) — EvaluatorPairDPDThermoDPD.h:250-254 */
double orc_dpd_alpha(uint16_t seed, uint32_t tag_i, uint32_t tag_j, uint64_t timestep);

/* DPDLJ: LJ conservative force + the DPD thermostat (hoomd/md/EvaluatorPairDPDThermoLJ.h); full list;
 * params = (epsilon, sigma, gamma) per type pair */
void orc_pair_dpdlj(const orc_box* b, const double* pos, const double* vel, const uint32_t* type,
                    const uint32_t* tag, uint32_t N, const uint32_t* n_neigh, const uint64_t* head,
                    const uint32_t* nlist, const double* params, const double* r_cut, uint32_t nt,
                    int energy_shift, uint16_t seed, uint64_t timestep, double dt, double kT, double* force,
                    double* virial);

/* hoomd/md/TwoStepConstantVolume.cc:7-80 / :213-261. vel N x 3, mass N, accel N x 3, image N x 3 */
void orc_step_one(const orc_box* b, double* pos, double* vel, const double* accel, int* image,
                  uint32_t n, double dt, double rescale);
void orc_step_two(double* vel, double* accel, const double* mass, const double* force4, uint32_t n,
                  double dt, double rescale);
/* hoomd/md/TwoStepLangevin.cc:213-292 (translational part, 3D, no reservoir tally): BD force
 * r * sqrt(6 gamma kT / dt) - gamma v with r uniform in [-1,1] drawn per particle from
 * RandomGenerator(Seed(TwoStepLangevin = 25, timestep, seed), Counter(tag)); gamma per type. */
void orc_step_two_langevin(double* vel, double* accel, const double* mass, const double* force4,
                           const uint32_t* type, const uint32_t* tag, uint32_t n, double dt, double kT,
                           const double* gamma, uint16_t seed, uint64_t timestep);
/* hoomd/md/ComputeThermo.cc:139-153 */
double orc_kinetic_energy(const double* vel, const double* mass, uint32_t n);
/* hoomd/md/Thermostat.h:357-447 (Bussi) with hoomd/RandomNumbers.h Normal/Gamma */
double orc_bussi_factor(double K, double dof, double dt, double kT, double tau, uint64_t timestep,
                        uint16_t seed, uint32_t first_tag);

/* Whole step loop, hoomd/md/IntegratorTwoStep.cc:51-123 + hoomd/md/NeighborList.cc:363-435 (half
 * list like hoomd.device.CPU). Used as the CPU baseline ("port") and for trajectory parity.
 * thermostat: 0 none, 1 Bussi, 2 Langevin method (TwoStepLangevin, one gamma for every type).
 * potential: 0 LJ, 1 DPD. Returns number of nlist builds. */
typedef struct
    {
    uint32_t n, ntypes;
    double dt, r_buff, kT, tau;
    int thermostat, potential, shift_mode;
    uint16_t seed;
    uint64_t timestep;
    double gamma; /* Langevin drag coefficient (thermostat == 2) */
    } orc_run_cfg;
int orc_run(const orc_box* b, orc_run_cfg* cfg, double* pos, double* vel, double* accel, int* image,
            const double* mass, const uint32_t* type, const uint32_t* tag, const double* params,
            const double* r_cut, uint64_t nsteps, double* force_out, double* n_neigh_mean);

#ifdef __cplusplus
}
#endif
#endif
