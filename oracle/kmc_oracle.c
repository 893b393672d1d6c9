/*
 * kmc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See kmc_oracle.h.
 *
 * A from-scratch C restatement of the reference algorithm; it deliberately keeps the
 * reference's order of floating-point operations (pow() for the LJ powers, left-to-right sums,
 * Python float modulo for the periodic wrap) so that it agrees with the reference to a few ulp.
 */
#include "kmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct OrcBins {
    int32_t nbx, nby;
    int64_t n;          /* atoms binned */
    int64_t* start;     /* [nbx*nby+2]; cell nbx*nby is the overflow cell (never addressed by a stencil) */
    double* xy;         /* cell-sorted interleaved positions, input order kept inside a cell */
    int64_t* idx;       /* original atom index of every sorted slot */
    int32_t* cell;      /* per input atom: flat cell id or -1 */
};

/* ------------------------------------------------------------------ Periodic.py */

/* Python's float `%` (sign follows the divisor). */
static double py_mod(double x, double m) {
    double r = fmod(x, m);
    if (r != 0.0) {
        if ((m < 0.0) != (r < 0.0)) r += m;
    } else {
        r = copysign(0.0, m);
    }
    return r;
}

/* Periodic.py:8-20 */
double orc_put_in_box(const OrcParams* p, double x) {
    x = py_mod(x, p->L);
    if (fabs(x) > p->L / 2) x += (p->L < 0) ? p->L : -p->L;
    return x;
}

/* Periodic.py:22-25 (even entries only) */
void orc_put_all_in_box(const OrcParams* p, double* R, int64_t n) {
    for (int64_t i = 0; i < n; ++i) R[2 * i] = orc_put_in_box(p, R[2 * i]);
}

/* Periodic.py:27-36: Rj-Ri, x entries wrapped */
void orc_displacement(const OrcParams* p, const double* Ri, const double* Rj, int64_t k, double* out) {
    for (int64_t i = 0; i < k; ++i) {
        out[2 * i] = orc_put_in_box(p, Rj[2 * i] - Ri[2 * i]);
        out[2 * i + 1] = Rj[2 * i + 1] - Ri[2 * i + 1];
    }
}

static inline double dist_xy(const OrcParams* p, double xi, double yi, double xj, double yj) {
    double dx = orc_put_in_box(p, xj - xi);
    double dy = yj - yi;
    return sqrt(dx * dx + dy * dy);   /* Periodic.py:58-59: d*=d ; sqrt(d[2j]+d[2j+1]) */
}

/* Periodic.py:38-48 */
double orc_distance(const OrcParams* p, const double* Ri, const double* Rj) {
    return dist_xy(p, Ri[0], Ri[1], Rj[0], Rj[1]);
}

/* Periodic.py:54-62 */
void orc_distances(const OrcParams* p, const double* a, int64_t na, const double* b, int64_t nb, double* out) {
    for (int64_t i = 0; i < na; ++i)
        for (int64_t j = 0; j < nb; ++j)
            out[i * nb + j] = dist_xy(p, a[2 * i], a[2 * i + 1], b[2 * j], b[2 * j + 1]);
}

/* ------------------------------------------------------------------ Bins.py */

/* Bins.py:9-22; int() truncates toward zero exactly like the C cast */
void orc_bin_index(const OrcParams* p, double x, double y, int32_t* nx, int32_t* ny) {
    *nx = (int32_t)((x + p->L / 2) / p->bin_size);
    *ny = (int32_t)((y - p->Dmin) / p->bin_size);
}

/* Bins.py:24-42.  The reference grows ragged python lists; a cell that was never grown behaves
 * like an empty one (IndexError is swallowed, Bins.py:80-83), so a dense CSR grid is equivalent.
 * Atoms whose (nx,ny) fall outside [0,nbins_x)x[0,nbins_y) are stored by the reference in cells
 * no stencil ever addresses (for -nbins <= n < 2 nbins); they go to the overflow cell here. */
OrcBins* orc_put_in_bins(const OrcParams* p, const double* R, int64_t n) {
    OrcBins* b = (OrcBins*)calloc(1, sizeof(OrcBins));
    b->nbx = p->nbins_x; b->nby = p->nbins_y; b->n = n;
    int64_t nc = (int64_t)b->nbx * b->nby;
    b->start = (int64_t*)calloc((size_t)nc + 2, sizeof(int64_t));
    b->xy = (double*)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
    b->idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    b->cell = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) {
        int32_t nx, ny;
        orc_bin_index(p, R[2 * i], R[2 * i + 1], &nx, &ny);
        int64_t c = (nx >= 0 && nx < b->nbx && ny >= 0 && ny < b->nby) ? (int64_t)nx * b->nby + ny : nc;
        b->cell[i] = (c == nc) ? -1 : (int32_t)c;
        b->start[c + 1]++;
    }
    for (int64_t c = 0; c <= nc; ++c) b->start[c + 1] += b->start[c];
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * ((size_t)nc + 1));
    memcpy(fill, b->start, sizeof(int64_t) * ((size_t)nc + 1));
    for (int64_t i = 0; i < n; ++i) {
        int64_t c = b->cell[i] < 0 ? nc : b->cell[i];
        int64_t s = fill[c]++;
        b->xy[2 * s] = R[2 * i]; b->xy[2 * s + 1] = R[2 * i + 1]; b->idx[s] = i;
    }
    free(fill);
    return b;
}

void orc_bins_free(OrcBins* b) {
    if (!b) return;
    free(b->start); free(b->xy); free(b->idx); free(b->cell); free(b);
}

int64_t orc_bins_count(const OrcBins* b) { return b->n; }

void orc_bins_cell_of_atom(const OrcBins* b, int32_t* cell_out) {
    memcpy(cell_out, b->cell, sizeof(int32_t) * (size_t)b->n);
}

void orc_bins_cell_counts(const OrcBins* b, int32_t* counts_out) {
    int64_t nc = (int64_t)b->nbx * b->nby;
    for (int64_t c = 0; c < nc; ++c) counts_out[c] = (int32_t)(b->start[c + 1] - b->start[c]);
}

/* Bins.py:44-70.  Each of the three indices is wrapped ONCE (both axes, Bins.py:57-65); the seam
 * hack appends a 4th column for nx==0 / nx==nbins_x-2 (Bins.py:66-69, tested on the raw nx). */
int orc_nearby_bins(const OrcParams* p, double x, double y, int32_t* xs, int32_t* ys) {
    int32_t nx, ny;
    orc_bin_index(p, x, y, &nx, &ny);
    int k = 3;
    for (int i = 0; i < 3; ++i) {
        xs[i] = nx - 1 + i;
        ys[i] = ny - 1 + i;
        if (xs[i] < 0) xs[i] += p->nbins_x;
        if (xs[i] >= p->nbins_x) xs[i] -= p->nbins_x;
        if (ys[i] < 0) ys[i] += p->nbins_y;
        if (ys[i] >= p->nbins_y) ys[i] -= p->nbins_y;
    }
    if (nx == 0) xs[k++] = p->nbins_x - 2;
    if (nx == p->nbins_x - 2) xs[k++] = 0;
    /* nbins_x==2 can make both conditions fire in the reference (5 columns); not representable
     * with 4 slots and never produced by Constants.py (W>=6 => nbins_x>=2; W=6 => nbins_x=2).
     * The caller guarantees nbins_x >= 3. */
    return k > 4 ? 4 : k;
}

/* iterate the stencil of (x,y) over cell list b in NearbyAtoms order (Bins.py:76-83) */
#define STENCIL_BEGIN(p, b, x, y)                                                   \
    {                                                                               \
        int32_t _xs[5], _ys[3];                                                     \
        int _nxs = orc_nearby_bins((p), (x), (y), _xs, _ys);                        \
        for (int _a = 0; _a < _nxs; ++_a) {                                         \
            if (_xs[_a] < 0 || _xs[_a] >= (b)->nbx) continue;                       \
            for (int _c = 0; _c < 3; ++_c) {                                        \
                if (_ys[_c] < 0 || _ys[_c] >= (b)->nby) continue;                   \
                int64_t _cell = (int64_t)_xs[_a] * (b)->nby + _ys[_c];              \
                for (int64_t _s = (b)->start[_cell]; _s < (b)->start[_cell + 1]; ++_s) {
#define STENCIL_END }}}}

/* Bins.py:72-84 */
int64_t orc_nearby_atoms(const OrcParams* p, const OrcBins* b, double x, double y, double* out) {
    int64_t m = 0;
    STENCIL_BEGIN(p, b, x, y)
        if (out) { out[2 * m] = b->xy[2 * _s]; out[2 * m + 1] = b->xy[2 * _s + 1]; }
        ++m;
    STENCIL_END
    return m;
}

/* Bins.py:86-111.  The adatoms are re-binned on every call (Bins.py:91); an adatom is its own
 * neighbour (d = 0 < 1.2 r_a, Bins.py:100). */
void orc_nearest_neighbors(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins,
                           int32_t* n_a, int32_t* n_s,
                           double* a_vals, int64_t* a_offs, double* s_vals, int64_t* s_offs) {
    OrcBins* ab = orc_put_in_bins(p, ad, n);
    const double ca = 1.2 * p->r_a, cs = 1.2 * p->r_as;
    int64_t oa = 0, os = 0;
    if (a_offs) a_offs[0] = 0;
    if (s_offs) s_offs[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
        double x = ad[2 * i], y = ad[2 * i + 1];
        int32_t ka = 0, ks = 0;
        STENCIL_BEGIN(p, ab, x, y)
            double d = dist_xy(p, x, y, ab->xy[2 * _s], ab->xy[2 * _s + 1]);
            if (d < ca) {
                if (a_vals) { a_vals[oa] = ab->xy[2 * _s]; a_vals[oa + 1] = ab->xy[2 * _s + 1]; }
                oa += 2; ++ka;
            }
        STENCIL_END
        STENCIL_BEGIN(p, sbins, x, y)
            double d = dist_xy(p, x, y, sbins->xy[2 * _s], sbins->xy[2 * _s + 1]);
            if (d < cs) {
                if (s_vals) { s_vals[os] = sbins->xy[2 * _s]; s_vals[os + 1] = sbins->xy[2 * _s + 1]; }
                os += 2; ++ks;
            }
        STENCIL_END
        if (n_a) n_a[i] = ka;
        if (n_s) n_s[i] = ks;
        if (a_offs) a_offs[i + 1] = oa;
        if (s_offs) s_offs[i + 1] = os;
    }
    orc_bins_free(ab);
}

/* ------------------------------------------------------------------ Energy.py */

/* Energy.py:24,28: E*((r0/r)**12 - 2*(r0/r)**6), 0 where r <= 0 */
double orc_pair_potential(double r, double E, double r0) {
    if (!(r > 0)) return 0.0;
    double q = r0 / r;
    return E * (pow(q, 12) - 2 * pow(q, 6));
}

/* Energy.py:49-60: one term of PairwiseForces; pairs with r < 1e-2 are skipped (:55) */
static inline void pair_force_acc(const OrcParams* p, double xi, double yi, double xj, double yj,
                                  double E, double r0, double* fx, double* fy) {
    double dx = orc_put_in_box(p, xj - xi);
    double dy = yj - yi;
    double r = sqrt(dx * dx + dy * dy);
    if (r < 1e-2) return;
    double f = E * (-12 * pow(r0, 12) / pow(r, 13) + 12 * pow(r0, 6) / pow(r, 7));
    *fx += dx / r * f;
    *fy += dy / r * f;
}

/* Energy.py:62-67 (all pairs, no cutoff) / bins3x3 stencil in aa_mode 1 */
void orc_adatom_adatom_forces(const OrcParams* p, const double* ad, int64_t n, double* F) {
    if (p->aa_mode == 0) {
#pragma omp parallel for schedule(static) if (n >= 512)
        for (int64_t i = 0; i < n; ++i) {
            double fx = 0, fy = 0;
            for (int64_t j = 0; j < n; ++j)
                pair_force_acc(p, ad[2 * i], ad[2 * i + 1], ad[2 * j], ad[2 * j + 1], p->E_a, p->r_a, &fx, &fy);
            F[2 * i] = fx; F[2 * i + 1] = fy;
        }
    } else {
        OrcBins* ab = orc_put_in_bins(p, ad, n);
#pragma omp parallel for schedule(static) if (n >= 512)
        for (int64_t i = 0; i < n; ++i) {
            double fx = 0, fy = 0;
            STENCIL_BEGIN(p, ab, ad[2 * i], ad[2 * i + 1])
                pair_force_acc(p, ad[2 * i], ad[2 * i + 1], ab->xy[2 * _s], ab->xy[2 * _s + 1], p->E_a, p->r_a, &fx, &fy);
            STENCIL_END
            F[2 * i] = fx; F[2 * i + 1] = fy;
        }
        orc_bins_free(ab);
    }
}

/* Energy.py:69-75 */
void orc_adatom_substrate_forces(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins, double* F) {
#pragma omp parallel for schedule(static) if (n >= 512)
    for (int64_t i = 0; i < n; ++i) {
        double fx = 0, fy = 0;
        STENCIL_BEGIN(p, sbins, ad[2 * i], ad[2 * i + 1])
            pair_force_acc(p, ad[2 * i], ad[2 * i + 1], sbins->xy[2 * _s], sbins->xy[2 * _s + 1], p->E_as, p->r_as, &fx, &fy);
        STENCIL_END
        F[2 * i] = fx; F[2 * i + 1] = fy;
    }
}

/* Energy.py:77-95 */
double orc_adatom_surface_energy(const OrcParams* p, const double* Ri, const OrcBins* sbins) {
    double u = 0;
    STENCIL_BEGIN(p, sbins, Ri[0], Ri[1])
        u += orc_pair_potential(dist_xy(p, Ri[0], Ri[1], sbins->xy[2 * _s], sbins->xy[2 * _s + 1]), p->E_as, p->r_as);
    STENCIL_END
    return u;
}

/* sum of LJ_aa between a point and a set of adatoms: all of them (aa_mode 0; Energy.py:115-117,
 * 172,181) or those in the stencil of an adatom cell list (aa_mode 1); only_after >= 0 keeps the
 * pairs with original index > only_after (the pop(0) pattern of RelaxEnergy, Energy.py:144-146). */
static double aa_energy_point(const OrcParams* p, double x, double y, const double* ad, int64_t n,
                              const OrcBins* ab, int64_t only_after) {
    double u = 0;
    if (p->aa_mode == 0) {
        for (int64_t j = only_after + 1; j < n; ++j)
            u += orc_pair_potential(dist_xy(p, x, y, ad[2 * j], ad[2 * j + 1]), p->E_a, p->r_a);
    } else {
        STENCIL_BEGIN(p, ab, x, y)
            if (ab->idx[_s] > only_after)
                u += orc_pair_potential(dist_xy(p, x, y, ab->xy[2 * _s], ab->xy[2 * _s + 1]), p->E_a, p->r_a);
        STENCIL_END
    }
    return u;
}

/* Energy.py:133-148: sum_{i<j} LJ_aa + sum_i substrate-stencil energy */
double orc_relax_energy(const OrcParams* p, const double* xn, int64_t n, const OrcBins* sbins) {
    OrcBins* ab = p->aa_mode ? orc_put_in_bins(p, xn, n) : NULL;
    double U = 0;
#pragma omp parallel for schedule(static) reduction(+ : U) if (n >= 512)
    for (int64_t i = 0; i < n; ++i) {
        double u = aa_energy_point(p, xn[2 * i], xn[2 * i + 1], xn, n, ab, i);
        u += orc_adatom_surface_energy(p, xn + 2 * i, sbins);
        U += u;
    }
    orc_bins_free(ab);
    return U;
}

/* Energy.py:121-131 (quirk kept: the loop stops at N-2, atom N-2 contributes nothing, the last
 * atom only its substrate term).  Result is unused by the driver (2DGrowth.py:116). */
double orc_total_energy(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins) {
    double U = 0;
    for (int64_t i = 0; i < n - 2; ++i) {
        double s = 0;
        for (int64_t j = i + 1; j < n; ++j)
            s += orc_pair_potential(dist_xy(p, ad[2 * i], ad[2 * i + 1], ad[2 * j], ad[2 * j + 1]), p->E_a, p->r_a);
        U += s;
        U += orc_adatom_surface_energy(p, ad + 2 * i, sbins);
    }
    if (n >= 1) U += orc_adatom_surface_energy(p, ad + 2 * (n - 1), sbins);
    return U;
}

/* Energy.py:170-172 */
double orc_hopping_energy(const OrcParams* p, const double* pos, const double* ad, int64_t n, const OrcBins* sbins) {
    OrcBins* ab = p->aa_mode ? orc_put_in_bins(p, ad, n) : NULL;
    double u = aa_energy_point(p, pos[0], pos[1], ad, n, ab, -1);
    u += orc_adatom_surface_energy(p, pos, sbins);
    orc_bins_free(ab);
    return u;
}

/* Energy.py:174-217 for all atoms.  U_loc sums LJ over the 1.2 r bond lists (self term is 0),
 * U_loc_ideal = -(n_a E_a + n_s E_as) with n_a counting the atom itself, C() = 1.0. */
void orc_barriers(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins,
                  double* u_appx, double* u_loc, double* u_ideal, double* delta_u, int32_t* n_a, int32_t* n_s) {
    OrcBins* ab = orc_put_in_bins(p, ad, n);
    const double ca = 1.2 * p->r_a, cs = 1.2 * p->r_as;
#pragma omp parallel for schedule(static) if (n >= 512)
    for (int64_t i = 0; i < n; ++i) {
        double x = ad[2 * i], y = ad[2 * i + 1];
        double ua = aa_energy_point(p, x, y, ad, n, ab, -1);        /* Energy.py:181 first sum */
        double us = 0, la = 0, ls = 0;
        int32_t ka = 0, ks = 0;
        STENCIL_BEGIN(p, ab, x, y)
            double d = dist_xy(p, x, y, ab->xy[2 * _s], ab->xy[2 * _s + 1]);
            if (d < ca) { la += orc_pair_potential(d, p->E_a, p->r_a); ++ka; }
        STENCIL_END
        STENCIL_BEGIN(p, sbins, x, y)
            double d = dist_xy(p, x, y, sbins->xy[2 * _s], sbins->xy[2 * _s + 1]);
            us += orc_pair_potential(d, p->E_as, p->r_as);
            if (d < cs) { ls += orc_pair_potential(d, p->E_as, p->r_as); ++ks; }
        STENCIL_END
        double appx = ua + us;
        double loc = la + ls;
        double ideal = -(ka * p->E_a + ks * p->E_as);
        if (u_appx) u_appx[i] = appx;
        if (u_loc) u_loc[i] = loc;
        if (u_ideal) u_ideal[i] = ideal;
        if (delta_u) delta_u[i] = 1.0 * (loc - ideal) + appx;
        if (n_a) n_a[i] = ka;
        if (n_s) n_s[i] = ks;
    }
    orc_bins_free(ab);
}

/* ------------------------------------------------------------------ 2DGrowth.py: rate table */

/* 2DGrowth.py:46-62 (surface iff n_a + n_s < 6) and :201-215 (rate = 1e12*exp(DeltaU*beta)) */
int64_t orc_hopping_rates(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins,
                          double* rate_dense, uint8_t* is_surface, double* delta_u) {
    double* du = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int32_t* na = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t* ns = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    orc_barriers(p, ad, n, sbins, NULL, NULL, NULL, du, na, ns);
    int64_t S = 0;
    for (int64_t i = 0; i < n; ++i) {
        int surf = (na[i] + ns[i] < 6);
        if (is_surface) is_surface[i] = (uint8_t)surf;
        if (delta_u) delta_u[i] = du[i];
        if (rate_dense) rate_dense[i] = surf ? p->omega * exp(du[i] * p->beta) : 0.0;
        S += surf;
    }
    free(du); free(na); free(ns);
    return S;
}

/* 2DGrowth.py:217-227 and :322-329.  pj = [0, cumsum(Rk)] left to right; Rtot = Rd + sum(Rk);
 * r = u*Rtot; first pj > r.  The dense array holds exact zeros for non-surface atoms, which
 * leaves every partial sum and the answer identical to the compacted list. */
int64_t orc_select(const OrcParams* p, const double* rate_dense, int64_t n, double u, int64_t* dense_idx, double* rtot) {
    double sum = 0;
    for (int64_t i = 0; i < n; ++i) sum += rate_dense[i];
    double Rtot = p->Rd + sum;
    if (rtot) *rtot = Rtot;
    double r = u * Rtot;
    double pj = 0;
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (rate_dense[i] == 0.0) continue;          /* not in the compacted list (rates are > 0 for surface atoms) */
        pj += rate_dense[i];
        if (pj > r) { if (dense_idx) *dense_idx = i; return k; }
        ++k;
    }
    if (dense_idx) *dense_idx = -1;
    return -1;
}

/* ------------------------------------------------------------------ 2DGrowth.py: events */

static void total_forces(const OrcParams* p, const double* x, int64_t n, const OrcBins* sbins, double* F, double* tmp) {
    orc_adatom_adatom_forces(p, x, n, F);
    orc_adatom_substrate_forces(p, x, n, sbins, tmp);
    for (int64_t i = 0; i < 2 * n; ++i) F[i] = F[i] + tmp[i];      /* 2DGrowth.py:120 */
}

static double max_force_sq(const double* F, int64_t n) {
    double m = 0;   /* 2DGrowth.py:129: the comprehension also visits empty slices whose dot is 0 */
    for (int64_t i = 0; i < n; ++i) {
        double v = F[2 * i] * F[2 * i] + F[2 * i + 1] * F[2 * i + 1];
        if (v > m) m = v;
    }
    return m;
}

static int64_t count_pairs_energy(const OrcParams* p, const double* x, int64_t n, const OrcBins* sbins) {
    int64_t c = 0;
    if (p->aa_mode == 0) c += n * (n - 1) / 2;
    for (int64_t i = 0; i < n; ++i) c += orc_nearby_atoms(p, sbins, x[2 * i], x[2 * i + 1], NULL);
    if (p->aa_mode) {
        OrcBins* ab = orc_put_in_bins(p, x, n);
        for (int64_t i = 0; i < n; ++i) c += orc_nearby_atoms(p, ab, x[2 * i], x[2 * i + 1], NULL);
        orc_bins_free(ab);
    }
    return c;
}

/* line search of 2DGrowth.py:121-123 / :152-156: 20 candidates x + a*s/20/scale, first argmin */
static int line_search(const OrcParams* p, const double* x, const double* s, int64_t n, int scale,
                       const OrcBins* sbins, double* cand, double* best, OrcRelaxStats* st) {
    double emin = 0; int amin = 0;
    for (int a = 0; a < 20; ++a) {
        for (int64_t i = 0; i < 2 * n; ++i) cand[i] = x[i] + a * s[i] / 20 / scale;
        double e = orc_relax_energy(p, cand, n, sbins);
        if (a == 0 || e < emin) { emin = e; amin = a; memcpy(best, cand, sizeof(double) * 2 * (size_t)n); }
    }
    if (st) st->line_searches++;
    return amin;
}

/* 2DGrowth.py:97-166.  Restarts (scale*2, or threshold*10 past scale 1024) start again from the
 * ORIGINAL positions handed to the call, exactly like the recursion in the reference. */
int orc_relaxation(const OrcParams* p, double* x, int64_t n, const OrcBins* sbins, OrcRelaxStats* st, int64_t max_ls) {
    return orc_relaxation_ex(p, x, n, sbins, st, max_ls, 16, 1e-4);       /* the defaults of 2DGrowth.py:97 */
}

/* Relaxation(adatoms, substrate_bins, scale, threshold) with explicit start values (LocalRelaxation calls it with
 * scale=4, 2DGrowth.py:192,198) */
int orc_relaxation_ex(const OrcParams* p, double* x, int64_t n, const OrcBins* sbins, OrcRelaxStats* st, int64_t max_ls,
                      int scale0, double threshold0) {
    size_t bytes = sizeof(double) * 2 * (size_t)(n > 0 ? n : 1);
    double* x0 = (double*)malloc(bytes); double* xn = (double*)malloc(bytes); double* xnp = (double*)malloc(bytes);
    double* sn = (double*)malloc(bytes); double* snp = (double*)malloc(bytes); double* dxnp = (double*)malloc(bytes);
    double* tmp = (double*)malloc(bytes); double* cand = (double*)malloc(bytes);
    memcpy(x0, x, bytes);
    int scale = scale0; double threshold = threshold0; int rc = 0;
    OrcRelaxStats local; memset(&local, 0, sizeof(local));
    if (!st) st = &local; else memset(st, 0, sizeof(*st));
    int64_t pairs_e = (n > 0) ? count_pairs_energy(p, x, n, sbins) : 0;
    for (;;) {                                           /* one pass == one (recursive) call */
        if (scale > 1024) { scale = 16; threshold *= 10; }               /* :111-113 */
        memcpy(xn, x0, bytes);
        total_forces(p, xn, n, sbins, sn, tmp);                           /* dx0, :120 ; sn = dx0, :125 */
        line_search(p, x0, sn, n, scale, sbins, cand, xnp, st);           /* :121-123 */
        total_forces(p, xnp, n, sbins, dxnp, tmp);                        /* :128 */
        double maxF = max_force_sq(dxnp, n), lastmaxF = maxF;             /* :129-130 */
        int restart = 0;
        while (maxF > threshold) {                                        /* :131 */
            if (max_ls > 0 && st->line_searches >= max_ls) { rc = -1; break; }
            double miny = xnp[1];
            for (int64_t i = 1; i < n; ++i) if (xnp[2 * i + 1] < miny) miny = xnp[2 * i + 1];
            if (miny > p->Dmax) { restart = 1; break; }                   /* :137-140 */
            double top = 0, bot = 0;
            for (int64_t i = 0; i < 2 * n; ++i) { top += xnp[i] * (xnp[i] - xn[i]); bot += xn[i] * xn[i]; }
            double beta = top / bot;                                      /* :143-145 */
            for (int64_t i = 0; i < 2 * n; ++i) snp[i] = dxnp[i] + beta * sn[i];   /* :147 */
            memcpy(xn, xnp, bytes); memcpy(sn, snp, bytes);               /* :149-150 */
            line_search(p, xn, sn, n, scale, sbins, cand, xnp, st);       /* :152-156 */
            total_forces(p, xnp, n, sbins, dxnp, tmp);                    /* :159 */
            maxF = max_force_sq(dxnp, n);                                 /* :160 */
            if (fabs(lastmaxF - maxF) < 1e-8) { restart = 1; break; }     /* :162-164 */
            lastmaxF = maxF;
        }
        st->maxF = maxF; st->final_scale = scale; st->final_threshold = threshold;
        if (rc < 0 || !restart) break;
        st->restarts++;
        scale *= 2;
    }
    st->pair_evals = st->line_searches * 20 * pairs_e;
    memcpy(x, xnp, bytes);
    orc_put_all_in_box(p, x, n);                                          /* :166 */
    free(x0); free(xn); free(xnp); free(sn); free(snp); free(dxnp); free(tmp); free(cand);
    return rc;
}

/* 2DGrowth.py:64-93 (placement only) */
int orc_deposit_place(const OrcParams* p, double* ad, int64_t n, const OrcBins* sbins, double u) {
    double ref_y = 0;                                                    /* :73-75 */
    if (n > 0) {
        ref_y = ad[1];
        for (int64_t i = 1; i < n; ++i) {
            double y = ad[2 * i + 1];
            if (p->deposit_mode == 0 ? (y < ref_y) : (y > ref_y)) ref_y = y;
        }
    }
    double new_x = orc_put_in_box(p, u * p->L);                           /* :76 */
    double new_y = 2 * p->r_as + ref_y;                                   /* :77 */
    OrcBins* ab = orc_put_in_bins(p, ad, n);                              /* :78 */
    int closeby = 0; int64_t guard = 0;
    while (!closeby) {                                                    /* :80-92 */
        STENCIL_BEGIN(p, ab, new_x, new_y)
            if (dist_xy(p, new_x, new_y, ab->xy[2 * _s], ab->xy[2 * _s + 1]) < 2 * p->r_a) closeby = 1;
        STENCIL_END
        STENCIL_BEGIN(p, sbins, new_x, new_y)
            if (dist_xy(p, new_x, new_y, sbins->xy[2 * _s], sbins->xy[2 * _s + 1]) < 2 * p->r_as) closeby = 1;
        STENCIL_END
        if (!closeby) new_y -= p->r_a / 2;
        if (++guard > 1000000) { orc_bins_free(ab); return -2; }
    }
    orc_bins_free(ab);
    ad[2 * n] = new_x; ad[2 * n + 1] = new_y;                             /* :93 */
    return 0;
}

/* 2DGrowth.py:229-255 (placement only).  k indexes the ADATOM array although the caller derived
 * it from the surface list (:327) -- hop_index_mode 0 keeps that; mode 1 expects the caller to
 * pass the adatom index of the selected surface atom. */
int orc_hop_place(const OrcParams* p, double* ad, int64_t n, const OrcBins* sbins, int64_t k, double u) {
    if (k < 0 || k >= n) return -3;
    int32_t* na = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* ns = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    orc_nearest_neighbors(p, ad, n, sbins, na, ns, NULL, NULL, NULL, NULL);      /* :239 SurfaceAtoms */
    double jx = ad[2 * k], jy = ad[2 * k + 1];                            /* :240 */
    /* candidates: surface atoms (of the list computed BEFORE removal) within 3 r_a, :242-246 */
    int64_t ncand = 0;
    int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i)
        if (na[i] + ns[i] < 6 && dist_xy(p, jx, jy, ad[2 * i], ad[2 * i + 1]) < 3 * p->r_a) cand[ncand++] = i;
    free(na); free(ns);
    if (ncand == 0) { free(cand); return -4; }                            /* reference: randint(0,-1) raises */
    int64_t pick = (int64_t)(u * (double)ncand);                          /* randint(0, ncand-1) -> int(u*n) */
    if (pick >= ncand) pick = ncand - 1;
    double tx = ad[2 * cand[pick]], ty = ad[2 * cand[pick] + 1];
    free(cand);
    memmove(ad + 2 * k, ad + 2 * k + 2, sizeof(double) * 2 * (size_t)(n - 1 - k));   /* :241 */
    double best_e = 0, bx = 0, by = 0;
    OrcParams q = *p;
    for (int i = 0; i < 6; ++i) {                                         /* :249-253 */
        double t = i * M_PI / 3;
        double pos[2] = { tx + cos(t) * p->r_a, ty + sin(t) * p->r_a };
        double e = orc_hopping_energy(&q, pos, ad, n - 1, sbins);
        if (i == 0 || e < best_e) { best_e = e; bx = pos[0]; by = pos[1]; }
    }
    ad[2 * (n - 1)] = orc_put_in_box(p, bx);                              /* :254-255 */
    ad[2 * (n - 1) + 1] = by;
    return 0;
}

/* one iteration of 2DGrowth.py:317-329 */
int orc_event(const OrcParams* p, double* ad, int64_t* n, const OrcBins* sbins, const double* us,
              int32_t* kind_out, int64_t* hop_k_out, double* rtot_out, OrcRelaxStats* st, int64_t max_ls) {
    int64_t N = *n;
    double* rate = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    orc_hopping_rates(p, ad, N, sbins, rate, NULL, NULL);                 /* :318 */
    int64_t dense = -1;
    int64_t k = orc_select(p, rate, N, us[0], &dense, rtot_out);          /* :319-324 */
    free(rate);
    int rc;
    if (k >= 0) {
        int64_t who = p->hop_index_mode == 0 ? k : dense;                /* :327 */
        rc = orc_hop_place(p, ad, N, sbins, who, us[1]);
        if (kind_out) *kind_out = 1;
        if (hop_k_out) *hop_k_out = who;
    } else {
        rc = orc_deposit_place(p, ad, N, sbins, us[1]);                   /* :329 */
        if (rc == 0) { N += 1; *n = N; }
        if (kind_out) *kind_out = 0;
        if (hop_k_out) *hop_k_out = -1;
    }
    if (rc != 0) return rc;
    return orc_relaxation(p, ad, N, sbins, st, max_ls);                   /* :94 / :256 */
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
