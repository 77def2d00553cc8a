/* md_oracle.c — CPU restatement (fp64, FMA-free) of HOOMD-blue v4.8.2's pair-force step.
 * TEST INFRASTRUCTURE ONLY (see md_oracle.h). Parity pinning: tests/test_oracle_golden.py checks this
 * file against the reference's own golden vectors (true_pair_list.json, forces_and_energies.json,
 * Philox KATs, Seed/Counter KATs) and against outputs of the reference itself run in the build
 * container (tests/golden npz files, made by tests/golden/make_golden.py).
 *
 * Written from the algorithm description in SURVEY.md §8 / Appendix A; citations are to
 * /root/reference paths. Nothing here is copied from the reference sources.
 */
#include "md_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- box ---- */

void orc_box_set(orc_box* b, double Lx, double Ly, double Lz)
    {
    /* hoomd/BoxDim.h:183-200: hi = L/2, lo = -hi, Linv = 1/L */
    double L[3] = {Lx, Ly, Lz};
    for (int d = 0; d < 3; d++)
        {
        b->L[d] = L[d];
        b->hi[d] = L[d] / 2.0;
        b->lo[d] = -b->hi[d];
        b->Linv[d] = (L[d] == 0.0) ? 0.0 : 1.0 / L[d];
        b->periodic[d] = 1;
        }
    }

void orc_make_fraction(const orc_box* b, const double p[3], const double gw[3], double f[3])
    {
    /* hoomd/BoxDim.h:276-284 with zero tilt: ((p-lo)*Linv + gf) / (1 + 2 gf), gf = gw / L */
    for (int d = 0; d < 3; d++)
        {
        double delta = p[d] - b->lo[d];
        double gf = gw ? gw[d] / b->L[d] : 0.0;
        f[d] = (delta * b->Linv[d] + gf) / (1.0 + 2.0 * gf);
        }
    }

static inline void min_image_inl(const orc_box* b, double w[3])
    {
    /* hoomd/BoxDim.h:343-389. z: one +-L shift; y, x: int(w*Linv + 0.5) multiples of L. */
    if (b->periodic[2])
        {
        if (w[2] >= b->hi[2])
            w[2] -= b->L[2];
        else if (w[2] < b->lo[2])
            w[2] += b->L[2];
        }
    for (int d = 1; d >= 0; d--)
        {
        if (!b->periodic[d])
            continue;
        if (w[d] >= b->hi[d])
            {
            int i = (int)(w[d] * b->Linv[d] + 0.5);
            w[d] -= (double)i * b->L[d];
            }
        else if (w[d] < b->lo[d])
            {
            int i = (int)(-w[d] * b->Linv[d] + 0.5);
            w[d] += (double)i * b->L[d];
            }
        }
    }

void orc_min_image(const orc_box* b, double w[3])
    {
    min_image_inl(b, w);
    }

void orc_wrap(const orc_box* b, double w[3], int img[3])
    {
    /* hoomd/BoxDim.h:407-474, flags = 0, zero tilt: x then y then z; >= hi -> -L, < lo -> +L */
    for (int d = 0; d < 3; d++)
        {
        if (!b->periodic[d])
            continue;
        if (w[d] >= b->hi[d])
            {
            w[d] -= b->L[d];
            img[d]++;
            }
        else if (w[d] < b->lo[d])
            {
            w[d] += b->L[d];
            img[d]--;
            }
        }
    }

/* ---------------------------------------------------------- cell list ---- */

void orc_cell_dims(const orc_box* b, double width, const double gw[3], uint32_t dim[3])
    {
    /* hoomd/CellList.cc:83-131 (m_multiple = 1) */
    for (int d = 0; d < 3; d++)
        {
        double g = gw ? gw[d] : 0.0;
        dim[d] = (uint32_t)((b->L[d] + 2.0 * g) / width);
        if (dim[d] == 0)
            dim[d] = 1;
        }
    }

uint32_t orc_cell_adj_width(const uint32_t dim[3])
    {
    /* hoomd/CellList.cc:252-273, radius 1 */
    uint32_t w = 1;
    for (int d = 0; d < 3; d++)
        w *= (dim[d] > 3) ? 3 : dim[d];
    return w;
    }

static int cmp_u32(const void* a, const void* b)
    {
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return (x > y) - (x < y);
    }

void orc_cell_adj(const uint32_t dim[3], uint32_t* adj)
    {
    /* hoomd/CellList.cc:346-402: the 27 periodic-wrapped neighbours, sorted, de-duplicated */
    const uint32_t width = orc_cell_adj_width(dim);
    const int mx = (int)dim[0], my = (int)dim[1], mz = (int)dim[2];
    for (int k = 0; k < mz; k++)
        for (int j = 0; j < my; j++)
            for (int i = 0; i < mx; i++)
                {
                uint32_t tmp[27];
                int n = 0;
                for (int nk = k - 1; nk <= k + 1; nk++)
                    for (int nj = j - 1; nj <= j + 1; nj++)
                        for (int ni = i - 1; ni <= i + 1; ni++)
                            {
                            int wi = ni % mx, wj = nj % my, wk = nk % mz;
                            if (wi < 0)
                                wi += mx;
                            if (wj < 0)
                                wj += my;
                            if (wk < 0)
                                wk += mz;
                            tmp[n++] = (uint32_t)((wk * my + wj) * mx + wi);
                            }
                qsort(tmp, 27, sizeof(uint32_t), cmp_u32);
                int m = 0;
                for (int q = 0; q < 27; q++)
                    if (q == 0 || tmp[q] != tmp[q - 1])
                        tmp[m++] = tmp[q];
                uint32_t cell = (uint32_t)((k * my + j) * mx + i);
                for (int q = 0; q < m && q < (int)width; q++)
                    adj[(size_t)cell * width + q] = tmp[q];
                }
    }

static int bin_of(const orc_box* b, const uint32_t dim[3], const double gw[3], const double p[3],
                  int ijk[3], double f[3])
    {
    /* hoomd/CellList.cc:455-470 and hoomd/md/NeighborListBinned.cc:107-121 */
    orc_make_fraction(b, p, gw, f);
    for (int d = 0; d < 3; d++)
        {
        ijk[d] = (int)(f[d] * dim[d]);
        if (ijk[d] == (int)dim[d] && b->periodic[d])
            ijk[d] = 0;
        }
    return (ijk[2] * (int)dim[1] + ijk[1]) * (int)dim[0] + ijk[0];
    }

int orc_cell_list(const orc_box* b, const uint32_t dim[3], const double gw[3], const double* pos,
                  uint32_t n_local, uint32_t n_total, uint32_t Nmax, uint32_t* cell_size,
                  uint32_t* cell_idx, uint32_t conditions[3])
    {
    /* hoomd/CellList.cc:404-547 */
    const size_t Nc = (size_t)dim[0] * dim[1] * dim[2];
    memset(cell_size, 0, sizeof(uint32_t) * Nc);
    conditions[0] = conditions[1] = conditions[2] = 0;
    for (uint32_t n = 0; n < n_total; n++)
        {
        const double* p = pos + 3 * (size_t)n;
        if (isnan(p[0]) || isnan(p[1]) || isnan(p[2]))
            {
            conditions[1] = n + 1;
            continue;
            }
        int ijk[3];
        double f[3];
        int bin = bin_of(b, dim, gw, p, ijk, f);
        int out = 0;
        for (int d = 0; d < 3; d++)
            if (f[d] < -0.00001 || f[d] >= 1.00001 || ijk[d] < 0 || ijk[d] >= (int)dim[d])
                out = 1;
        if (out)
            {
            if (n < n_local)
                conditions[2] = n + 1;
            continue;
            }
        uint32_t offset = cell_size[bin];
        if (offset < Nmax)
            cell_idx[(size_t)bin * Nmax + offset] = n;
        else if (offset + 1 > conditions[0])
            conditions[0] = offset + 1;
        cell_size[bin]++;
        }
    return (conditions[0] || conditions[1] || conditions[2]) ? 1 : 0;
    }

/* --------------------------------------------------------- neighbour list ---- */

void orc_r_listsq(const double* r_cut, uint32_t nt, double r_buff, double* r_listsq,
                  double* rcut_max, double* rcut_max_max)
    {
    /* hoomd/md/NeighborList.cc:490-512 */
    double mm = 0.0;
    for (uint32_t i = 0; i < nt; i++)
        {
        double mi = 0.0;
        for (uint32_t j = 0; j < nt; j++)
            {
            double rc = r_cut[j * nt + i];
            if (rc > mi)
                mi = rc;
            double r_list = (rc > 0.0) ? rc + r_buff : 0.0;
            r_listsq[j * nt + i] = r_list * r_list;
            }
        if (rcut_max)
            rcut_max[i] = mi;
        if (mi > mm)
            mm = mi;
        }
    if (rcut_max_max)
        *rcut_max_max = mm;
    }

static inline double dist_sq(const orc_box* b, const double* pi, const double* pj)
    {
    /* dx = pi - pj; minImage; dot = x*x + y*y + z*z left to right
     * (hoomd/md/NeighborListBinned.cc:139-143, hoomd/VectorMath.h dot) */
    double dx[3] = {pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]};
    min_image_inl(b, dx);
    return dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    }

uint32_t* orc_nlist_binned(const orc_box* b, const double* pos, const uint32_t* type,
                           uint32_t n_local, uint32_t n_total, const double* r_cut, uint32_t nt,
                           double r_buff, int full, const double gw[3], uint32_t* n_neigh,
                           uint64_t* head, uint64_t* total, uint32_t dim_out[3])
    {
    double* r_listsq = (double*)malloc(sizeof(double) * nt * nt);
    double rmaxmax;
    orc_r_listsq(r_cut, nt, r_buff, r_listsq, NULL, &rmaxmax);

    /* hoomd/md/NeighborListBinned.cc:39-45: nominal width = max r_cut + buffer */
    uint32_t dim[3];
    orc_cell_dims(b, rmaxmax + r_buff, gw, dim);
    if (dim_out)
        memcpy(dim_out, dim, sizeof(dim));
    const size_t Nc = (size_t)dim[0] * dim[1] * dim[2];
    const uint32_t aw = orc_cell_adj_width(dim);
    uint32_t* adj = (uint32_t*)malloc(sizeof(uint32_t) * Nc * aw);
    orc_cell_adj(dim, adj);

    /* cell list with the overflow/regrow loop of hoomd/CellList.cc:186-198 */
    uint32_t Nmax = (uint32_t)ceil((double)n_total / (double)Nc);
    if (Nmax == 0)
        Nmax = 1;
    uint32_t* cell_size = (uint32_t*)malloc(sizeof(uint32_t) * Nc);
    uint32_t* cell_idx = NULL;
    for (;;)
        {
        cell_idx = (uint32_t*)realloc(cell_idx, sizeof(uint32_t) * Nc * Nmax);
        uint32_t cond[3];
        orc_cell_list(b, dim, gw, pos, n_local, n_total, Nmax, cell_size, cell_idx, cond);
        if (cond[0] > Nmax)
            {
            Nmax = cond[0];
            continue;
            }
        break;
        }

    /* Cell contents as contiguous (x, y, z) beside the index, as the reference's xyzf array
       (hoomd/CellList.cc:513-521): the inner loop streams them instead of gathering pos[j]. */
    double* cell_xyz = (double*)malloc(sizeof(double) * 3 * Nc * Nmax);
    for (size_t c = 0; c < Nc; c++)
        for (uint32_t o = 0; o < cell_size[c]; o++)
            {
            const uint32_t j = cell_idx[c * Nmax + o];
            memcpy(cell_xyz + 3 * (c * Nmax + o), pos + 3 * (size_t)j, sizeof(double) * 3);
            }

    /* One pass appending rows in particle order: head[i] is the running total, so no count pass is
       needed; the buffer doubles on overflow (stands in for the Nmax overflow/retry loop of
       hoomd/md/NeighborList.cc:411-427 — same rows, same order). */
    uint64_t cap = (uint64_t)n_local * 64 + 1024, s_tot = 0;
    uint32_t* nlist = (uint32_t*)malloc(sizeof(uint32_t) * cap);
    for (uint32_t i = 0; i < n_local; i++)
        {
        const double* pi = pos + 3 * (size_t)i;
        const uint32_t ti = type[i];
        int ijk[3];
        double f[3];
        int my_cell = bin_of(b, dim, gw, pi, ijk, f);
        uint32_t cnt = 0;
        head[i] = s_tot;
        for (uint32_t a = 0; a < aw; a++)
            {
            const uint32_t nc = adj[(size_t)my_cell * aw + a];
            const uint32_t sz = cell_size[nc];
            if (s_tot + cnt + sz > cap)
                {
                cap = 2 * cap + sz;
                nlist = (uint32_t*)realloc(nlist, sizeof(uint32_t) * cap);
                }
            const uint32_t* idx = cell_idx + (size_t)nc * Nmax;
            const double* xyz = cell_xyz + 3 * (size_t)nc * Nmax;
            for (uint32_t o = 0; o < sz; o++)
                {
                const uint32_t j = idx[o];
                const uint32_t tp = type[j] * nt + ti; /* Index2D(ti,tj) = tj*nt + ti */
                if (j == i || r_cut[tp] <= 0.0)
                    continue;
                const double drsq = dist_sq(b, pi, xyz + 3 * (size_t)o);
                if (drsq <= r_listsq[tp] && (full || i < j))
                    nlist[s_tot + cnt++] = j;
                }
            }
        n_neigh[i] = cnt;
        s_tot += cnt;
        }
    *total = s_tot;
    free(cell_xyz);
    free(cell_idx);
    free(cell_size);
    free(adj);
    free(r_listsq);
    return nlist;
    }

uint32_t* orc_nlist_brute(const orc_box* b, const double* pos, const uint32_t* type, uint32_t n_local,
                          uint32_t n_total, const double* r_cut, uint32_t nt, double r_buff, int full,
                          uint32_t* n_neigh, uint64_t* head, uint64_t* total)
    {
    double* r_listsq = (double*)malloc(sizeof(double) * nt * nt);
    orc_r_listsq(r_cut, nt, r_buff, r_listsq, NULL, NULL);
    uint32_t* nlist = NULL;
    for (int pass = 0; pass < 2; pass++)
        {
        for (uint32_t i = 0; i < n_local; i++)
            {
            uint32_t cnt = 0;
            for (uint32_t j = 0; j < n_total; j++)
                {
                uint32_t tp = type[j] * nt + type[i];
                if (j == i || r_cut[tp] <= 0.0)
                    continue;
                double drsq = dist_sq(b, pos + 3 * (size_t)i, pos + 3 * (size_t)j);
                if (drsq <= r_listsq[tp] && (full || i < j))
                    {
                    if (pass)
                        nlist[head[i] + cnt] = j;
                    cnt++;
                    }
                }
            if (!pass)
                n_neigh[i] = cnt;
            }
        if (!pass)
            {
            uint64_t s = 0;
            for (uint32_t i = 0; i < n_local; i++)
                {
                head[i] = s;
                s += n_neigh[i];
                }
            *total = s;
            nlist = (uint32_t*)malloc(sizeof(uint32_t) * (s ? s : 1));
            }
        }
    free(r_listsq);
    return nlist;
    }

void orc_free(void* p)
    {
    free(p);
    }

int orc_distance_check(const orc_box* b, const double* pos, const double* last_pos,
                       const uint32_t* type, uint32_t n, const double* rcut_max, double r_buff)
    {
    /* hoomd/md/NeighborList.cc:1244-1269 with lambda = (1,1,1), lambda_min = 1 */
    for (uint32_t i = 0; i < n; i++)
        {
        double old_rmin = rcut_max[type[i]];
        double rmax = old_rmin + r_buff;
        double delta_max = (rmax * 1.0 - old_rmin) / 2.0;
        double maxsq = (delta_max > 0) ? delta_max * delta_max : 0;
        double dx[3];
        for (int d = 0; d < 3; d++)
            dx[d] = pos[3 * (size_t)i + d] - 1.0 * last_pos[3 * (size_t)i + d];
        min_image_inl(b, dx);
        if (dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2] >= maxsq)
            return 1;
        }
    return 0;
    }

/* -------------------------------------------------------------- potentials ---- */

int orc_eval_lj(double rsq, double rcutsq, double epsilon, double sigma, int energy_shift,
                double* force_divr, double* pair_eng)
    {
    /* params: hoomd/md/EvaluatorPairLJ.h:119-140; evaluation :162-211 */
    double sigma_6 = sigma * sigma * sigma * sigma * sigma * sigma;
    double epsilon_x_4 = 4.0 * epsilon;
    double lj1 = epsilon_x_4 * sigma_6 * sigma_6;
    double lj2 = epsilon_x_4 * sigma_6;
    if (rsq < rcutsq && lj1 != 0)
        {
        double r2inv = 1.0 / rsq;
        double r6inv = r2inv * r2inv * r2inv;
        *force_divr = r2inv * r6inv * (12.0 * lj1 * r6inv - 6.0 * lj2);
        *pair_eng = r6inv * (lj1 * r6inv - lj2);
        if (energy_shift)
            {
            double rcut2inv = 1.0 / rcutsq;
            double rcut6inv = rcut2inv * rcut2inv * rcut2inv;
            *pair_eng -= rcut6inv * (lj1 * rcut6inv - lj2);
            }
        return 1;
        }
    return 0;
    }

int orc_eval_dpd_conservative(double rsq, double rcutsq, double A, double* force_divr,
                              double* pair_eng)
    {
    /* hoomd/md/EvaluatorPairDPDThermoDPD.h:186-203 */
    if (rsq < rcutsq)
        {
        double rinv = 1.0 / sqrt(rsq);
        double r = 1.0 / rinv;
        double rcutinv = 1.0 / sqrt(rcutsq);
        double rcut = 1.0 / rcutinv;
        *force_divr = A * (rinv - rcutinv);
        *pair_eng = A * (rcut - r) - (1.0 / 2.0) * A * rcutinv * (rcutsq - rsq);
        return 1;
        }
    return 0;
    }

static void accumulate_pair(double* force, double* virial, uint32_t N, uint32_t i, uint32_t j,
                            const double dx[3], double fdivr, double fdivr_virial, double eng,
                            int third_law_j, int compute_virial, double acc[10])
    {
    /* hoomd/md/PotentialPair.h:739-772: acc = {fx,fy,fz,pe,vxx,vxy,vxz,vyy,vyz,vzz} of i */
    double f2 = fdivr_virial * 0.5;
    acc[0] += dx[0] * fdivr;
    acc[1] += dx[1] * fdivr;
    acc[2] += dx[2] * fdivr;
    acc[3] += eng * 0.5;
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (compute_virial)
        {
        v[0] = f2 * dx[0] * dx[0];
        v[1] = f2 * dx[0] * dx[1];
        v[2] = f2 * dx[0] * dx[2];
        v[3] = f2 * dx[1] * dx[1];
        v[4] = f2 * dx[1] * dx[2];
        v[5] = f2 * dx[2] * dx[2];
        for (int l = 0; l < 6; l++)
            acc[4 + l] += v[l];
        }
    if (third_law_j)
        {
        force[4 * (size_t)j + 0] -= dx[0] * fdivr;
        force[4 * (size_t)j + 1] -= dx[1] * fdivr;
        force[4 * (size_t)j + 2] -= dx[2] * fdivr;
        force[4 * (size_t)j + 3] += eng * 0.5;
        if (compute_virial)
            for (int l = 0; l < 6; l++)
                virial[(size_t)l * N + j] += v[l];
        }
    (void)i;
    }

void orc_pair_lj(const orc_box* b, const double* pos, const uint32_t* type, uint32_t N,
                 const uint32_t* n_neigh, const uint64_t* head, const uint32_t* nlist, int half,
                 const double* params, const double* r_cut, const double* r_on, uint32_t nt,
                 int shift_mode, int compute_virial, double* force, double* virial)
    {
    /* hoomd/md/PotentialPair.h:587-794 */
    memset(force, 0, sizeof(double) * 4 * (size_t)N);
    memset(virial, 0, sizeof(double) * 6 * (size_t)N);
    for (uint32_t i = 0; i < N; i++)
        {
        const double* pi = pos + 3 * (size_t)i;
        double acc[10] = {0};
        for (uint32_t k = 0; k < n_neigh[i]; k++)
            {
            uint32_t j = nlist[head[i] + k];
            const double* pj = pos + 3 * (size_t)j;
            double dx[3] = {pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]};
            min_image_inl(b, dx);
            double rsq = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
            uint32_t tp = type[j] * nt + type[i];
            double rcutsq = r_cut[tp] * r_cut[tp]; /* hoomd/md/PotentialPair.h:498 */
            double ronsq = 0.0;
            if (shift_mode == 2)
                ronsq = r_on[tp] * r_on[tp];
            int energy_shift = (shift_mode == 1) || (shift_mode == 2 && ronsq > rcutsq);
            double fdivr = 0.0, eng = 0.0;
            if (!orc_eval_lj(rsq, rcutsq, params[2 * tp], params[2 * tp + 1], energy_shift, &fdivr,
                             &eng))
                continue;
            if (shift_mode == 2 && rsq >= ronsq && rsq < rcutsq)
                {
                /* hoomd/md/PotentialPair.h:711-737 */
                double old_eng = eng, old_f = fdivr;
                double denom_inv = 1.0 / ((rcutsq - ronsq) * (rcutsq - ronsq) * (rcutsq - ronsq));
                double m = rsq - rcutsq;
                double s = m * m * (rcutsq + 2.0 * rsq - 3.0 * ronsq) * denom_inv;
                double ds = 12.0 * (rsq - ronsq) * m * denom_inv;
                eng = old_eng * s;
                fdivr = s * old_f - ds * old_eng;
                }
            accumulate_pair(force, virial, N, i, j, dx, fdivr, fdivr, eng, half && j < N,
                            compute_virial, acc);
            }
        for (int c = 0; c < 4; c++)
            force[4 * (size_t)i + c] += acc[c];
        if (compute_virial)
            for (int l = 0; l < 6; l++)
                virial[(size_t)l * N + i] += acc[4 + l];
        }
    }

double orc_lj_tail_correction(const double* params, const double* r_cut, uint32_t nt,
                              const uint32_t* n_per_type, double volume, double* virial_diag)
    {
    /* hoomd/md/PotentialPair.h:283-337 with hoomd/md/EvaluatorPairLJ.h:213-248 */
    double e = 0.0, w = 0.0;
    for (uint32_t ti = 0; ti < nt; ti++)
        {
        double rho_i = n_per_type[ti] / volume;
        for (uint32_t tj = 0; tj < nt; tj++)
            {
            double rho_j = n_per_type[tj] / volume;
            uint32_t tp = tj * nt + ti;
            double eps = params[2 * tp], sig = params[2 * tp + 1];
            double s6 = sig * sig * sig * sig * sig * sig;
            double lj1 = 4.0 * eps * s6 * s6, lj2 = 4.0 * eps * s6;
            double rcutsq = r_cut[tp] * r_cut[tp];
            double ie = 0.0, ip = 0.0;
            if (rcutsq != 0)
                {
                double rc3inv = 1.0 / pow(rcutsq, 1.5);
                double rc9inv = rc3inv * rc3inv * rc3inv;
                ie = lj1 / 9.0 * rc9inv - lj2 / 3.0 * rc3inv;
                ip = lj1 * 4.0 / 3.0 * rc9inv - lj2 * 2.0 * rc3inv;
                }
            e += 2.0 * n_per_type[ti] * M_PI * rho_j * ie;
            double dp = 4.0 / 6.0 * rho_i * rho_j * M_PI * ip;
            w += 3 * volume * dp / 3.0;
            }
        }
    if (virial_diag)
        *virial_diag = w;
    return e;
    }

/* ----------------------------------------------------------- bonds, exclusions ---- */

void orc_filter_bond_exclusions(uint32_t n, uint32_t* n_neigh, const uint64_t* head, uint32_t* nlist,
                                const uint32_t* tag, const uint32_t* rtag, const uint32_t* bond_group,
                                uint32_t n_bonds)
    {
    /* addExclusionsFromBonds: addExclusion(tag_a, tag_b) for every bond (duplicates are ignored by
       isExcluded, hoomd/md/NeighborList.cc:569-640); exclusion lists by tag, then by index, then the
       filter of :1487-1536 */
    uint32_t max_tag = 0;
    for (uint32_t i = 0; i < n; i++)
        if (tag[i] > max_tag)
            max_tag = tag[i];
    uint32_t* cnt = (uint32_t*)calloc((size_t)max_tag + 2, sizeof(uint32_t));
    for (uint32_t k = 0; k < n_bonds; k++)
        {
        cnt[bond_group[2 * k] + 1]++;
        cnt[bond_group[2 * k + 1] + 1]++;
        }
    for (uint32_t t = 0; t <= max_tag; t++)
        cnt[t + 1] += cnt[t];
    uint32_t* ex = (uint32_t*)malloc(sizeof(uint32_t) * (2 * (size_t)n_bonds + 1));
    uint32_t* fill = (uint32_t*)calloc((size_t)max_tag + 1, sizeof(uint32_t));
    for (uint32_t k = 0; k < n_bonds; k++)
        {
        const uint32_t a = bond_group[2 * k], c = bond_group[2 * k + 1];
        ex[cnt[a] + fill[a]++] = c;
        ex[cnt[c] + fill[c]++] = a;
        }
    for (uint32_t idx = 0; idx < n; idx++)
        {
        const uint32_t t = tag[idx];
        const uint32_t n_ex = cnt[t + 1] - cnt[t];
        uint32_t new_n = 0;
        for (uint32_t q = 0; q < n_neigh[idx]; q++)
            {
            const uint32_t cur = nlist[head[idx] + q];
            int excluded = 0;
            for (uint32_t e = 0; e < n_ex; e++)
                if (rtag[ex[cnt[t] + e]] == cur)
                    {
                    excluded = 1;
                    break;
                    }
            if (!excluded)
                nlist[head[idx] + new_n++] = cur;
            }
        n_neigh[idx] = new_n;
        }
    free(cnt);
    free(ex);
    free(fill);
    }

static int bond_forces(int kind, const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
                       const uint32_t* bond_group, const uint32_t* bond_type, uint32_t n_bonds,
                       const double* params, uint32_t stride, double* force, double* virial)
    {
    /* hoomd/md/PotentialBond.h:176-312 */
    memset(force, 0, sizeof(double) * 4 * (size_t)N);
    memset(virial, 0, sizeof(double) * 6 * (size_t)N);
    for (uint32_t i = 0; i < n_bonds; i++)
        {
        const uint32_t ia = rtag[bond_group[2 * i]], ib = rtag[bond_group[2 * i + 1]];
        if (ia >= N || ib >= N)
            return 1 + (int)i;
        double dx[3];
        for (int d = 0; d < 3; d++)
            dx[d] = pos[3 * (size_t)ib + d] - pos[3 * (size_t)ia + d];
        orc_min_image(b, dx);
        double rsq = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
        const double* prm = params + (size_t)stride * bond_type[i];
        double force_divr = 0.0, bond_eng = 0.0;
        if (kind == 0)
            {
            /* hoomd/md/EvaluatorBondHarmonic.h:104-124 */
            const double K = prm[0], r_0 = prm[1];
            const double r = sqrt(rsq);
            force_divr = K * (r_0 / r - 1.0);
            if (!isfinite(force_divr))
                force_divr = 0.0;
            bond_eng = 0.5 * K * (r_0 - r) * (r_0 - r);
            }
        else
            {
            /* hoomd/md/EvaluatorBondFENE.h:44-53 parameters, :117-158 evaluation */
            const double K = prm[0], r_0 = prm[1], epsilon_in = prm[2], sigma = prm[3], delta = prm[4];
            const double sigma_6 = sigma * sigma * sigma * sigma * sigma * sigma;
            const double lj1 = 4.0 * epsilon_in * sigma_6 * sigma_6, lj2 = 4.0 * epsilon_in * sigma_6;
            const double rtemp = sqrt(rsq) - delta;
            const double rmdoverr = rtemp / sqrt(rsq);
            rsq = rtemp * rtemp;
            const double r2inv = 1.0 / rsq;
            const double r6inv = r2inv * r2inv * r2inv;
            double wca = 0.0, pair_eng = 0.0;
            const double sigma6inv = lj2 / lj1;
            const double epsilon = lj2 * lj2 / 4.0 / lj1;
            if (lj1 != 0 && r6inv > sigma6inv / 2.0)
                {
                wca = r2inv * r6inv * (12.0 * lj1 * r6inv - 6.0 * lj2);
                pair_eng = (r6inv * (lj1 * r6inv - lj2) + epsilon);
                }
            force_divr = wca * rmdoverr;
            bond_eng = pair_eng;
            if (K != 0.0)
                {
                if (rsq >= r_0 * r_0)
                    return -1 - (int)i; /* "bond out of bounds" */
                force_divr += -K / (1.0 - rsq / (r_0 * r_0)) * rmdoverr;
                bond_eng += -0.5 * K * (r_0 * r_0) * log(1.0 - rsq / (r_0 * r_0));
                }
            }
        bond_eng *= 0.5;
        const double f2 = 0.5 * force_divr;
        const double v[6] = {dx[0] * dx[0] * f2, dx[0] * dx[1] * f2, dx[0] * dx[2] * f2,
                             dx[1] * dx[1] * f2, dx[1] * dx[2] * f2, dx[2] * dx[2] * f2};
        for (int d = 0; d < 3; d++)
            {
            force[4 * (size_t)ib + d] += force_divr * dx[d];
            force[4 * (size_t)ia + d] -= force_divr * dx[d];
            }
        force[4 * (size_t)ib + 3] += bond_eng;
        force[4 * (size_t)ia + 3] += bond_eng;
        for (int q = 0; q < 6; q++)
            {
            virial[(size_t)q * N + ib] += v[q];
            virial[(size_t)q * N + ia] += v[q];
            }
        }
    return 0;
    }

int orc_bond_harmonic(const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
                      const uint32_t* bond_group, const uint32_t* bond_type, uint32_t n_bonds,
                      const double* params, double* force, double* virial)
    {
    return bond_forces(0, b, pos, N, rtag, bond_group, bond_type, n_bonds, params, 2, force, virial);
    }

int orc_bond_fenewca(const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
                     const uint32_t* bond_group, const uint32_t* bond_type, uint32_t n_bonds,
                     const double* params, double* force, double* virial)
    {
    return bond_forces(1, b, pos, N, rtag, bond_group, bond_type, n_bonds, params, 5, force, virial);
    }

int orc_angle(int kind, const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
              const uint32_t* angle_group, const uint32_t* angle_type, uint32_t n_angles,
              const double* params, double* force, double* virial)
    {
    /* kind 0: hoomd/md/HarmonicAngleForceCompute.cc:104-258 (E = k/2 (theta - t0)^2);
       kind 1: hoomd/md/CosineSqAngleForceCompute.cc:104-262 (E = k/2 (cos theta - cos t0)^2).
       params = (k, t0) per angle type; angle_group holds 3 member tags (a, b, c), b the vertex. */
    const double SMALL = 0.001;
    memset(force, 0, sizeof(double) * 4 * (size_t)N);
    memset(virial, 0, sizeof(double) * 6 * (size_t)N);
    for (uint32_t i = 0; i < n_angles; i++)
        {
        const uint32_t ia = rtag[angle_group[3 * i]], ib = rtag[angle_group[3 * i + 1]],
                       ic = rtag[angle_group[3 * i + 2]];
        if (ia >= N || ib >= N || ic >= N)
            return 1 + (int)i;
        double dab[3], dcb[3];
        for (int d = 0; d < 3; d++)
            {
            dab[d] = pos[3 * (size_t)ia + d] - pos[3 * (size_t)ib + d];
            dcb[d] = pos[3 * (size_t)ic + d] - pos[3 * (size_t)ib + d];
            }
        orc_min_image(b, dab);
        orc_min_image(b, dcb);
        const double rsqab = dab[0] * dab[0] + dab[1] * dab[1] + dab[2] * dab[2];
        const double rab = sqrt(rsqab);
        const double rsqcb = dcb[0] * dcb[0] + dcb[1] * dcb[1] + dcb[2] * dcb[2];
        const double rcb = sqrt(rsqcb);
        double c_abbc = dab[0] * dcb[0] + dab[1] * dcb[1] + dab[2] * dcb[2];
        c_abbc /= rab * rcb;
        if (c_abbc > 1.0)
            c_abbc = 1.0;
        if (c_abbc < -1.0)
            c_abbc = -1.0;
        const double K = params[2 * angle_type[i]], t_0 = params[2 * angle_type[i] + 1];
        double a, eng;
        if (kind == 0)
            {
            double s_abbc = sqrt(1.0 - c_abbc * c_abbc);
            if (s_abbc < SMALL)
                s_abbc = SMALL;
            s_abbc = 1.0 / s_abbc;
            const double dth = acos(c_abbc) - t_0;
            const double tk = K * dth;
            a = -1.0 * tk * s_abbc;
            eng = (tk * dth) * (1.0 / 6.0);
            }
        else
            {
            const double dcosth = c_abbc - cos(t_0);
            const double tk = K * dcosth;
            a = 1.0 * tk;
            eng = (tk * dcosth) * (1.0 / 6.0);
            }
        const double a11 = a * c_abbc / rsqab;
        const double a12 = -a / (rab * rcb);
        const double a22 = a * c_abbc / rsqcb;
        double fab[3], fcb[3];
        for (int d = 0; d < 3; d++)
            {
            fab[d] = a11 * dab[d] + a12 * dcb[d];
            fcb[d] = a22 * dcb[d] + a12 * dab[d];
            }
        const double v[6] = {(1. / 3.) * (dab[0] * fab[0] + dcb[0] * fcb[0]),
                             (1. / 3.) * (dab[1] * fab[0] + dcb[1] * fcb[0]),
                             (1. / 3.) * (dab[2] * fab[0] + dcb[2] * fcb[0]),
                             (1. / 3.) * (dab[1] * fab[1] + dcb[1] * fcb[1]),
                             (1. / 3.) * (dab[2] * fab[1] + dcb[2] * fcb[1]),
                             (1. / 3.) * (dab[2] * fab[2] + dcb[2] * fcb[2])};
        for (int d = 0; d < 3; d++)
            {
            force[4 * (size_t)ia + d] += fab[d];
            force[4 * (size_t)ib + d] -= fab[d] + fcb[d];
            force[4 * (size_t)ic + d] += fcb[d];
            }
        const uint32_t idx[3] = {ia, ib, ic};
        for (int m = 0; m < 3; m++)
            {
            force[4 * (size_t)idx[m] + 3] += eng;
            for (int q = 0; q < 6; q++)
                virial[(size_t)q * N + idx[m]] += v[q];
            }
        }
    return 0;
    }

int orc_dihedral(int kind, const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
                 const uint32_t* group, const uint32_t* type, uint32_t n_dihedrals, const double* params,
                 double* force, double* virial)
    {
    /* kind 1: hoomd.md.dihedral.OPLS (hoomd/md/OPLSDihedralForceCompute.cc:100-352), params =
       (k1, k2, k3, k4) per type (halved on entry like setParams :57-72). kind 0: */
    /* hoomd/md/HarmonicDihedralForceCompute.cc:104-356 (hoomd.md.dihedral.Periodic):
       E = k/2 (1 + d cos(n phi - phi0)); params = (k, d, n, phi0) per dihedral type; a quarter of the
       energy and of the virial to each member. */
    memset(force, 0, sizeof(double) * 4 * (size_t)N);
    memset(virial, 0, sizeof(double) * 6 * (size_t)N);
    for (uint32_t i = 0; i < n_dihedrals; i++)
        {
        uint32_t idx[4];
        for (int m = 0; m < 4; m++)
            {
            idx[m] = rtag[group[4 * i + m]];
            if (idx[m] >= N)
                return 1 + (int)i;
            }
        double dab[3], dcb[3], ddc[3];
        for (int d = 0; d < 3; d++)
            {
            dab[d] = pos[3 * (size_t)idx[0] + d] - pos[3 * (size_t)idx[1] + d];
            dcb[d] = pos[3 * (size_t)idx[2] + d] - pos[3 * (size_t)idx[1] + d];
            ddc[d] = pos[3 * (size_t)idx[3] + d] - pos[3 * (size_t)idx[2] + d];
            }
        orc_min_image(b, dab);
        orc_min_image(b, dcb);
        orc_min_image(b, ddc);
        const double* prm = params + 4 * (size_t)type[i];
        const double half = kind == 1 ? 0.5 : 1.0;
        const double prm_d[4] = {prm[0] * half, prm[1] * half, prm[2] * half, prm[3] * half};
#define MINIMAGE(v) orc_min_image(b, v)

        double dcbm[3] = {-dcb[0], -dcb[1], -dcb[2]};
        MINIMAGE(dcbm);
        const double aax = dab[1] * dcbm[2] - dab[2] * dcbm[1];
        const double aay = dab[2] * dcbm[0] - dab[0] * dcbm[2];
        const double aaz = dab[0] * dcbm[1] - dab[1] * dcbm[0];
        const double bbx = ddc[1] * dcbm[2] - ddc[2] * dcbm[1];
        const double bby = ddc[2] * dcbm[0] - ddc[0] * dcbm[2];
        const double bbz = ddc[0] * dcbm[1] - ddc[1] * dcbm[0];
        const double raasq = aax * aax + aay * aay + aaz * aaz;
        const double rbbsq = bbx * bbx + bby * bby + bbz * bbz;
        const double rgsq = dcbm[0] * dcbm[0] + dcbm[1] * dcbm[1] + dcbm[2] * dcbm[2];
        const double rg = sqrt(rgsq);
        double rginv = 0.0, raa2inv = 0.0, rbb2inv = 0.0;
        if (rg > 0.0)
            rginv = 1.0 / rg;
        if (raasq > 0.0)
            raa2inv = 1.0 / raasq;
        if (rbbsq > 0.0)
            rbb2inv = 1.0 / rbbsq;
        const double rabinv = sqrt(raa2inv * rbb2inv);
        double c_abcd = (aax * bbx + aay * bby + aaz * bbz) * rabinv;
        const double s_abcd = rg * rabinv * (aax * ddc[0] + aay * ddc[1] + aaz * ddc[2]);
        if (c_abcd > 1.0)
            c_abcd = 1.0;
        if (c_abcd < -1.0)
            c_abcd = -1.0;
        double df, eng;
        if (kind == 1)
            {
            // hoomd/md/OPLSDihedralForceCompute.cc:236-279: p = sum_i k_i/2 (1 + (-1)^(i+1) cos(i phi)),
            // df = dp/dphi through the Chebyshev-like recursion; the halved k_i are the parameters
            const double k1 = prm_d[0], k2 = prm_d[1], k3 = prm_d[2], k4 = prm_d[3];
            double ddf1 = c_abcd, df1 = s_abcd, cos_term = ddf1;
            double p = k1 * (1.0 + cos_term);
            df = k1 * df1;
            ddf1 = cos_term * c_abcd - df1 * s_abcd;
            df1 = cos_term * s_abcd + df1 * c_abcd;
            cos_term = ddf1;
            p += k2 * (1.0 - cos_term);
            df += -2.0 * k2 * df1;
            ddf1 = cos_term * c_abcd - df1 * s_abcd;
            df1 = cos_term * s_abcd + df1 * c_abcd;
            cos_term = ddf1;
            p += k3 * (1.0 + cos_term);
            df += 3.0 * k3 * df1;
            ddf1 = cos_term * c_abcd - df1 * s_abcd;
            df1 = cos_term * s_abcd + df1 * c_abcd;
            cos_term = ddf1;
            p += k4 * (1.0 - cos_term);
            df += -4.0 * k4 * df1;
            eng = 0.25 * p;
            }
        else
            {
            const double K = prm_d[0], sign = prm_d[1], phi_0 = prm_d[3];
            const int multi = (int)prm_d[2];
            double p = 1.0, dfab = 0.0, ddfab = 0.0;
            for (int j = 0; j < multi; j++)
                {
                ddfab = p * c_abcd - dfab * s_abcd;
                dfab = p * s_abcd + dfab * c_abcd;
                p = ddfab;
                }
            const double sin_phi_0 = sin(phi_0), cos_phi_0 = cos(phi_0);
            p = p * cos_phi_0 + dfab * sin_phi_0;
            p = p * sign;
            dfab = dfab * cos_phi_0 - ddfab * sin_phi_0;
            dfab = dfab * sign;
            dfab *= (double)-multi;
            p += 1.0;
            if (multi == 0)
                {
                p = 1.0 + sign;
                dfab = 0.0;
                }
            df = -K * dfab * 0.500;
            eng = p * K * 0.125;
            }
        const double fg = dab[0] * dcbm[0] + dab[1] * dcbm[1] + dab[2] * dcbm[2];
        const double hg = ddc[0] * dcbm[0] + ddc[1] * dcbm[1] + ddc[2] * dcbm[2];
        const double fga = fg * raa2inv * rginv, hgb = hg * rbb2inv * rginv;
        const double gaa = -raa2inv * rg, gbb = rbb2inv * rg;
        const double dtf[3] = {gaa * aax, gaa * aay, gaa * aaz};
        const double dtg[3] = {fga * aax - hgb * bbx, fga * aay - hgb * bby, fga * aaz - hgb * bbz};
        const double dth[3] = {gbb * bbx, gbb * bby, gbb * bbz};
        double ffa[3], ffb[3], ffc[3], ffd[3];
        for (int q = 0; q < 3; q++)
            {
            const double s2 = df * dtg[q];
            ffa[q] = df * dtf[q];
            ffb[q] = s2 - ffa[q];
            ffd[q] = df * dth[q];
            ffc[q] = -s2 - ffd[q];
            }
        const double vir[6] = {(1. / 4.) * (dab[0] * ffa[0] + dcb[0] * ffc[0] + (ddc[0] + dcb[0]) * ffd[0]),
                               (1. / 4.) * (dab[1] * ffa[0] + dcb[1] * ffc[0] + (ddc[1] + dcb[1]) * ffd[0]),
                               (1. / 4.) * (dab[2] * ffa[0] + dcb[2] * ffc[0] + (ddc[2] + dcb[2]) * ffd[0]),
                               (1. / 4.) * (dab[1] * ffa[1] + dcb[1] * ffc[1] + (ddc[1] + dcb[1]) * ffd[1]),
                               (1. / 4.) * (dab[2] * ffa[1] + dcb[2] * ffc[1] + (ddc[2] + dcb[2]) * ffd[1]),
                               (1. / 4.) * (dab[2] * ffa[2] + dcb[2] * ffc[2] + (ddc[2] + dcb[2]) * ffd[2])};
#undef MINIMAGE
        const double* ff[4] = {ffa, ffb, ffc, ffd};
        for (int m = 0; m < 4; m++)
            {
            for (int d = 0; d < 3; d++)
                force[4 * (size_t)idx[m] + d] += ff[m][d];
            force[4 * (size_t)idx[m] + 3] += eng;
            for (int q = 0; q < 6; q++)
                virial[(size_t)q * N + idx[m]] += vir[q];
            }
        }
    return 0;
    }

int orc_dihedral_periodic(const orc_box* b, const double* pos, uint32_t N, const uint32_t* rtag,
                          const uint32_t* group, const uint32_t* type, uint32_t n_dihedrals,
                          const double* params, double* force, double* virial)
    {
    return orc_dihedral(0, b, pos, N, rtag, group, type, n_dihedrals, params, force, virial);
    }

/* ---------------------------------------------------------------------- RNG ---- */

void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
    {
    /* Philox4x32-10, Salmon/Moraes/Dror/Shaw SC'11 (random123 philox.h): multipliers
       0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85 */
    uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++)
        {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0;
        c[1] = n1;
        c[2] = n2;
        c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
        }
    memcpy(out, c, sizeof(c));
    }

void orc_seed_key(uint8_t id, uint64_t timestep, uint16_t seed, uint32_t key[2])
    {
    /* hoomd/RandomNumbers.h:153-158 */
    key[0] = (uint32_t)id << 24 | (uint32_t)seed << 8
             | (uint32_t)((timestep & 0x000000ff00000000ull) >> 32);
    key[1] = (uint32_t)(timestep & 0x00000000ffffffffull);
    }

typedef struct
    {
    uint32_t key[2], ctr[4];
    } orc_rng;

static void rng_init(orc_rng* g, uint8_t id, uint64_t timestep, uint16_t seed, uint32_t a,
                     uint32_t b, uint32_t c, uint16_t d)
    {
    /* hoomd/RandomNumbers.h:185-187: ctr = {d<<16, c, b, a} */
    orc_seed_key(id, timestep, seed, g->key);
    g->ctr[0] = (uint32_t)d << 16;
    g->ctr[1] = c;
    g->ctr[2] = b;
    g->ctr[3] = a;
    }

static void rng_next(orc_rng* g, uint32_t u[4])
    {
    /* hoomd/RandomNumbers.h:264-270 */
    orc_philox4x32_10(g->ctr, g->key, u);
    g->ctr[0] += 1;
    }

static double u01_u64(uint64_t in)
    {
    /* hoomd/RandomNumbers.h:63-74: in * 2^-64 + 2^-65 */
    const double factor = 1.0 / (18446744073709551615.0 + 1.0);
    const double halffactor = 0.5 * factor;
    return (double)in * factor + halffactor;
    }

static double uneg11_u64(uint64_t in)
    {
    /* hoomd/RandomNumbers.h:76-87: int64(in) * 2^-63 + 2^-64 */
    const double factor = 1.0 / (9223372036854775807.0 + 1.0);
    const double halffactor = 0.5 * factor;
    return (double)(int64_t)in * factor + halffactor;
    }

static double rng_canonical(orc_rng* g)
    {
    uint32_t u[4];
    rng_next(g, u);
    return u01_u64((uint64_t)u[0] << 32 | u[1]);
    }

static double rng_normal(orc_rng* g)
    {
    /* hoomd/RandomNumbers.h:361-374 (Box-Muller, sigma 1, mu 0); CPU sincospi = sincos(M_PI*x)
       hoomd/HOOMDMath.h:330-339 */
    uint32_t u[4];
    rng_next(g, u);
    uint64_t u0 = (uint64_t)u[0] << 32 | u[1], u1 = (uint64_t)u[2] << 32 | u[3];
    double a = M_PI * uneg11_u64(u0);
    double x = sin(a);
    double r = sqrt(-2.0 * log(u01_u64(u1)));
    x *= r;
    return x * 1.0 + 0.0;
    }

static double rng_gamma(orc_rng* g, double alpha_in, double bparam)
    {
    /* hoomd/RandomNumbers.h:461-505 (Marsaglia-Tsang) */
    double alpha = alpha_in;
    if (alpha_in < 1.0)
        alpha += 1.0;
    double d = alpha - (1. / 3.);
    double c = (1.0 / sqrt(d)) / 3.;
    double v;
    for (;;)
        {
        double x;
        do
            {
            x = rng_normal(g);
            v = 1.0 + c * x;
            } while (v <= 0.);
        v = v * v * v;
        const double x2 = x * x;
        double u = rng_canonical(g);
        if (u < 1.0 - 0.0331 * x2 * x2)
            break;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v)))
            break;
        }
    v *= d * bparam;
    if (alpha_in < 1.0)
        v *= pow(rng_canonical(g), 1.0 / alpha_in);
    return v;
    }

double orc_dpd_alpha(uint16_t seed, uint32_t tag_i, uint32_t tag_j, uint64_t timestep)
    {
    /* hoomd/md/EvaluatorPairDPDThermoDPD.h:237-254; id 22 = hoomd/RNGIdentifiers.h
       EvaluatorPairDPDThermo; Counter(min, max) */
    uint32_t oi = tag_i, oj = tag_j;
    if (tag_i > tag_j)
        {
        oi = tag_j;
        oj = tag_i;
        }
    orc_rng g;
    rng_init(&g, 22, timestep, seed, oi, oj, 0, 0);
    /* UniformDistribution(-1,1): a + width*u, hoomd/RandomNumbers.h:325-335 */
    return -1.0 + 2.0 * rng_canonical(&g);
    }

void orc_pair_dpd(const orc_box* b, const double* pos, const double* vel, const uint32_t* type,
                  const uint32_t* tag, uint32_t N, const uint32_t* n_neigh, const uint64_t* head,
                  const uint32_t* nlist, int half, const double* params, const double* r_cut,
                  uint32_t nt, uint16_t seed, uint64_t timestep, double dt, double kT, double* force,
                  double* virial)
    {
    /* hoomd/md/PotentialPairDPDThermo.h:102-268. Virial is accumulated unconditionally there. */
    memset(force, 0, sizeof(double) * 4 * (size_t)N);
    memset(virial, 0, sizeof(double) * 6 * (size_t)N);
    for (uint32_t i = 0; i < N; i++)
        {
        const double* pi = pos + 3 * (size_t)i;
        const double* vi = vel + 3 * (size_t)i;
        double acc[10] = {0};
        for (uint32_t k = 0; k < n_neigh[i]; k++)
            {
            uint32_t j = nlist[head[i] + k];
            const double* pj = pos + 3 * (size_t)j;
            const double* vj = vel + 3 * (size_t)j;
            double dx[3] = {pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]};
            double dv[3] = {vi[0] - vj[0], vi[1] - vj[1], vi[2] - vj[2]};
            min_image_inl(b, dx);
            double rsq = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
            double rdotv = dx[0] * dv[0] + dx[1] * dv[1] + dx[2] * dv[2];
            uint32_t tp = type[j] * nt + type[i];
            double rcutsq = r_cut[tp] * r_cut[tp];
            double A = params[2 * tp], gamma = params[2 * tp + 1];
            if (!(rsq < rcutsq))
                continue;
            /* hoomd/md/EvaluatorPairDPDThermoDPD.h:222-277 */
            double rinv = 1.0 / sqrt(rsq);
            double r = 1.0 / rinv;
            double rcutinv = 1.0 / sqrt(rcutsq);
            double rcut = 1.0 / rcutinv;
            double alpha = orc_dpd_alpha(seed, tag[i], tag[j], timestep);
            double fdivr = A * (rinv - rcutinv);
            double fcons = fdivr;
            fdivr -= gamma * rdotv * (rinv - rcutinv) * (rinv - rcutinv);
            fdivr += (1.0 / sqrt(dt / (kT * gamma * 6.0))) * (rinv - rcutinv) * alpha;
            double eng = A * (rcut - r) - (1.0 / 2.0) * A * rcutinv * (rcutsq - rsq);
            /* third-law update has no j < N guard here (PotentialPairDPDThermo.h:238) */
            accumulate_pair(force, virial, N, i, j, dx, fdivr, fcons, eng, half && j < N, 1, acc);
            }
        for (int c = 0; c < 4; c++)
            force[4 * (size_t)i + c] += acc[c];
        for (int l = 0; l < 6; l++)
            virial[(size_t)l * N + i] += acc[4 + l];
        }
    }

void orc_pair_dpdlj(const orc_box* b, const double* pos, const double* vel, const uint32_t* type,
                    const uint32_t* tag, uint32_t N, const uint32_t* n_neigh, const uint64_t* head,
                    const uint32_t* nlist, const double* params, const double* r_cut, uint32_t nt,
                    int energy_shift, uint16_t seed, uint64_t timestep, double dt, double kT, double* force,
                    double* virial)
    {
    /* PotentialPairDPDThermo<EvaluatorPairDPDThermoLJ> (hoomd/md/PotentialPairDPDThermo.h:102-268 with
       hoomd/md/EvaluatorPairDPDThermoLJ.h:66-76 parameters, :236-296 evaluation), full list.
       params = (epsilon, sigma, gamma) per type pair. */
    memset(force, 0, sizeof(double) * 4 * (size_t)N);
    memset(virial, 0, sizeof(double) * 6 * (size_t)N);
    for (uint32_t i = 0; i < N; i++)
        {
        const double* pi = pos + 3 * (size_t)i;
        const double* vi = vel + 3 * (size_t)i;
        double acc[10] = {0};
        for (uint32_t k = 0; k < n_neigh[i]; k++)
            {
            uint32_t j = nlist[head[i] + k];
            const double* pj = pos + 3 * (size_t)j;
            const double* vj = vel + 3 * (size_t)j;
            double dx[3] = {pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]};
            double dv[3] = {vi[0] - vj[0], vi[1] - vj[1], vi[2] - vj[2]};
            min_image_inl(b, dx);
            double rsq = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
            double rdotv = dx[0] * dv[0] + dx[1] * dv[1] + dx[2] * dv[2];
            uint32_t tp = type[j] * nt + type[i];
            double rcutsq = r_cut[tp] * r_cut[tp];
            const double epsilon = params[3 * tp], sigma = params[3 * tp + 1], gamma = params[3 * tp + 2];
            const double sigma_6 = sigma * sigma * sigma * sigma * sigma * sigma;
            const double lj1 = 4.0 * epsilon * sigma_6 * sigma_6, lj2 = 4.0 * epsilon * sigma_6;
            if (!(rsq < rcutsq && lj1 != 0))
                continue;
            double rinv = 1.0 / sqrt(rsq);
            double r2inv = 1.0 / rsq;
            double r6inv = r2inv * r2inv * r2inv;
            double rcutinv = 1.0 / sqrt(rcutsq);
            double alpha = orc_dpd_alpha(seed, tag[i], tag[j], timestep);
            double fdivr = r2inv * r6inv * (12.0 * lj1 * r6inv - 6.0 * lj2);
            double fcons = fdivr;
            fdivr -= gamma * rdotv * (rinv - rcutinv) * (rinv - rcutinv);
            fdivr += (1.0 / sqrt(dt / (kT * gamma * 6.0))) * (rinv - rcutinv) * alpha;
            double eng = r6inv * (lj1 * r6inv - lj2);
            if (energy_shift)
                {
                double rcut2inv = 1.0 / rcutsq;
                double rcut6inv = rcut2inv * rcut2inv * rcut2inv;
                eng -= rcut6inv * (lj1 * rcut6inv - lj2);
                }
            accumulate_pair(force, virial, N, i, j, dx, fdivr, fcons, eng, 0, 1, acc);
            }
        for (int c = 0; c < 4; c++)
            force[4 * (size_t)i + c] += acc[c];
        for (int l = 0; l < 6; l++)
            virial[(size_t)l * N + i] += acc[4 + l];
        }
    }

/* ---------------------------------------------------------------- integration ---- */

void orc_step_one(const orc_box* b, double* pos, double* vel, const double* accel, int* image,
                  uint32_t n, double dt, double rescale)
    {
    /* hoomd/md/TwoStepConstantVolume.cc:33-79: v += 1/2 a dt; v *= s; r += dt v; then wrap */
    for (uint32_t j = 0; j < n; j++)
        for (int d = 0; d < 3; d++)
            {
            double v = vel[3 * (size_t)j + d] + (1.0 / 2.0) * accel[3 * (size_t)j + d] * dt;
            v *= rescale;
            pos[3 * (size_t)j + d] += dt * v;
            vel[3 * (size_t)j + d] = v;
            }
    for (uint32_t j = 0; j < n; j++)
        orc_wrap(b, pos + 3 * (size_t)j, image + 3 * (size_t)j);
    }

void orc_step_two(double* vel, double* accel, const double* mass, const double* force4, uint32_t n,
                  double dt, double rescale)
    {
    /* hoomd/md/TwoStepConstantVolume.cc:234-260: a = F * (1/m); v *= s; v += 1/2 dt a */
    for (uint32_t j = 0; j < n; j++)
        {
        double minv = 1.0 / mass[j];
        for (int d = 0; d < 3; d++)
            {
            double a = force4[4 * (size_t)j + d] * minv;
            double v = vel[3 * (size_t)j + d] * rescale;
            v += (1.0 / 2.0) * dt * a;
            vel[3 * (size_t)j + d] = v;
            accel[3 * (size_t)j + d] = a;
            }
        }
    }

void orc_step_two_langevin(double* vel, double* accel, const double* mass, const double* force4,
                           const uint32_t* type, const uint32_t* tag, uint32_t n, double dt, double kT,
                           const double* gamma, uint16_t seed, uint64_t timestep)
    {
    /* hoomd/md/TwoStepLangevin.cc:245-292; id 25 = TwoStepLangevin (hoomd/RNGIdentifiers.h:51);
       UniformDistribution(-1, 1): a + width * canonical, hoomd/RandomNumbers.h:325-335 */
    for (uint32_t j = 0; j < n; j++)
        {
        orc_rng g;
        rng_init(&g, 25, timestep, seed, tag[j], 0, 0, 0);
        double r[3];
        for (int d = 0; d < 3; d++)
            r[d] = -1.0 + 2.0 * rng_canonical(&g);
        const double gm = gamma[type[j]];
        const double coeff = sqrt(6.0 * gm * kT / dt);
        const double minv = 1.0 / mass[j];
        for (int d = 0; d < 3; d++)
            {
            const double bd = r[d] * coeff - gm * vel[3 * (size_t)j + d];
            const double a = (force4[4 * (size_t)j + d] + bd) * minv;
            accel[3 * (size_t)j + d] = a;
            vel[3 * (size_t)j + d] += (1.0 / 2.0) * a * dt;
            }
        }
    }

double orc_kinetic_energy(const double* vel, const double* mass, uint32_t n)
    {
    /* hoomd/md/ComputeThermo.cc:139-153 */
    double ke = 0.0;
    for (uint32_t j = 0; j < n; j++)
        ke += mass[j]
              * (vel[3 * (size_t)j] * vel[3 * (size_t)j] + vel[3 * (size_t)j + 1] * vel[3 * (size_t)j + 1]
                 + vel[3 * (size_t)j + 2] * vel[3 * (size_t)j + 2]);
    return ke * 0.5;
    }

double orc_bussi_factor(double K, double dof, double dt, double kT, double tau, uint64_t timestep,
                        uint16_t seed, uint32_t first_tag)
    {
    /* hoomd/md/Thermostat.h:357-447; id 45 = BussiThermostat (hoomd/RNGIdentifiers.h);
       RandomGenerator(Seed, Counter(instance_id)) */
    if (dt == 0.0 || dof == 0)
        return 1.0;
    orc_rng g;
    rng_init(&g, 45, timestep, seed, first_tag, 0, 0, 0);
    double decay = 0.0;
    if (tau != 0.0)
        decay = exp(-dt / tau);
    double rn = rng_normal(&g);
    double r_gamma = 0.0;
    if (dof > 1.0)
        r_gamma = 2.0 * rng_gamma(&g, (dof - 1.0) / 2.0, 1.0);
    double v = kT / 2.0 / K;
    double term1 = v * (1.0 - decay) * (r_gamma + rn * rn);
    double term2 = 2.0 * rn * sqrt(v * (1.0 - decay) * decay);
    return sqrt(decay + term1 + term2);
    }

/* ----------------------------------------------------------------- step loop ---- */

int orc_run(const orc_box* b, orc_run_cfg* cfg, double* pos, double* vel, double* accel, int* image,
            const double* mass, const uint32_t* type, const uint32_t* tag, const double* params,
            const double* r_cut, uint64_t nsteps, double* force_out, double* n_neigh_mean)
    {
    /* hoomd/md/IntegratorTwoStep.cc:51-123 (update) and :226-281 (prepRun): forces at the current
       timestep first (accelerations are taken as given), then per step: step one, forces at t+1,
       step two. Half neighbour list as hoomd.device.CPU selects (hoomd/md/pair/pair.py:135-142).
       Rebuild policy hoomd/md/NeighborList.cc:1325-1418 with rebuild_check_delay=1, check_dist. */
    const uint32_t N = cfg->n, nt = cfg->ntypes;
    double* force = (double*)malloc(sizeof(double) * 4 * (size_t)N);
    double* virial = (double*)malloc(sizeof(double) * 6 * (size_t)N);
    double* last_pos = (double*)malloc(sizeof(double) * 3 * (size_t)N);
    uint32_t* n_neigh = (uint32_t*)malloc(sizeof(uint32_t) * N);
    uint64_t* head = (uint64_t*)malloc(sizeof(uint64_t) * N);
    double* rcut_max = (double*)malloc(sizeof(double) * nt);
    double* r_listsq = (double*)malloc(sizeof(double) * nt * nt);
    orc_r_listsq(r_cut, nt, cfg->r_buff, r_listsq, rcut_max, NULL);
    uint32_t* nlist = NULL;
    uint64_t total = 0;
    int builds = 0, have_list = 0;
    double nn_acc = 0.0;
    const double dof = 3.0 * N - 3.0; /* hoomd/md/IntegratorTwoStep.cc:154-189 (3D, momentum conserved) */

    for (uint64_t s = 0; s <= nsteps; s++)
        {
        const uint64_t t_force = cfg->timestep + s; /* forces evaluated at this timestep */
        if (s > 0)
            {
            double scale = 1.0;
            if (cfg->thermostat == 1)
                scale = orc_bussi_factor(orc_kinetic_energy(vel, mass, N), dof, cfg->dt, cfg->kT,
                                         cfg->tau, t_force - 1, cfg->seed, tag[0]);
            orc_step_one(b, pos, vel, accel, image, N, cfg->dt, scale);
            }
        int rebuild = !have_list || cfg->r_buff < 1e-6
                      || orc_distance_check(b, pos, last_pos, type, N, rcut_max, cfg->r_buff);
        if (rebuild)
            {
            orc_free(nlist);
            nlist = orc_nlist_binned(b, pos, type, N, N, r_cut, nt, cfg->r_buff, 0, NULL, n_neigh,
                                     head, &total, NULL);
            memcpy(last_pos, pos, sizeof(double) * 3 * (size_t)N);
            have_list = 1;
            builds++;
            nn_acc += (double)total / N;
            }
        if (cfg->potential == 1)
            orc_pair_dpd(b, pos, vel, type, tag, N, n_neigh, head, nlist, 1, params, r_cut, nt,
                         cfg->seed, t_force, cfg->dt, cfg->kT, force, virial);
        else
            orc_pair_lj(b, pos, type, N, n_neigh, head, nlist, 1, params, r_cut, r_cut, nt,
                        cfg->shift_mode, 0, force, virial);
        if (s > 0 && cfg->thermostat == 2)
            {
            /* integrateStepTwo(timestep) of the step that started at t_force - 1 */
            double* gm = (double*)malloc(sizeof(double) * nt);
            for (uint32_t q = 0; q < nt; q++)
                gm[q] = cfg->gamma;
            orc_step_two_langevin(vel, accel, mass, force, type, tag, N, cfg->dt, cfg->kT, gm, cfg->seed,
                                  t_force - 1);
            free(gm);
            }
        else if (s > 0)
            orc_step_two(vel, accel, mass, force, N, cfg->dt, 1.0);
        }
    cfg->timestep += nsteps;
    if (force_out)
        memcpy(force_out, force, sizeof(double) * 4 * (size_t)N);
    if (n_neigh_mean)
        *n_neigh_mean = builds ? nn_acc / builds : 0.0;
    orc_free(nlist);
    free(force);
    free(virial);
    free(last_pos);
    free(n_neigh);
    free(head);
    free(rcut_max);
    free(r_listsq);
    return builds;
    }
