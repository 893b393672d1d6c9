/*
 * kmc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C fp64 restatement of the hot path of YouCantCodeWithUs/KMCIslandGrowth
 * (Periodic.py, Bins.py, Energy.py and the rate-table / selection / event functions of
 * 2DGrowth.py).  Every function cites the reference file:line it follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Parity status: PINNED -- the reference ships no golden vectors of its own, so
 * the oracle is pinned against outputs of the reference itself, produced by
 * oracle/make_golden.py (mechanical py2->py3 conversion run in the build container) and
 * committed under tests/golden/ (tests/test_oracle_golden.py).
 *
 * aa_mode: 0 = all-pairs adatom sums exactly as the reference computes them for < 1000 adatoms
 *              (Energy.py:62-67,115-117,172,181);
 *          1 = "bins3x3": the adatom sums run over the reference's own 3x3(+seam) bin stencil
 *              of an adatom cell list (the evident intent of the crashing branch Energy.py:118),
 *              the only mode usable for the >=1000-adatom configs where the reference cannot run.
 */
#ifndef KMC_ORACLE_H
#define KMC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcParams {
    double L;          /* Constants.py:35  L = W*r_s                      */
    double bin_size;   /* Constants.py:39  3*r_s                          */
    double Dmin;       /* Constants.py:36                                 */
    double Dmax;       /* Constants.py:37                                 */
    double E_a, r_a;   /* Constants.py:16,20                              */
    double E_as, r_as; /* Constants.py:28,29                              */
    double beta;       /* Constants.py:45                                 */
    double omega;      /* 2DGrowth.py:210  1.0e12                         */
    double Rd;         /* 2DGrowth.py:43   6.0*18/W*deposition_rate       */
    int32_t nbins_x;   /* Constants.py:40                                 */
    int32_t nbins_y;   /* Constants.py:41                                 */
    int32_t aa_mode;   /* 0 all-pairs (reference), 1 bins3x3              */
    int32_t hop_index_mode;  /* 0 reference quirk (2DGrowth.py:327 -> :240), 1 mapped to the surface atom */
    int32_t deposit_mode;    /* 0 reference (2DGrowth.py:73-77, min_y), 1 start above the highest adatom  */
    int32_t reserved;
} OrcParams;

typedef struct OrcBins OrcBins;   /* cell list in the reference's column-major [nx][ny] order */

/* Periodic.py */
double orc_put_in_box(const OrcParams* p, double x);                                  /* :8-20  */
void   orc_put_all_in_box(const OrcParams* p, double* R, int64_t n);                  /* :22-25 */
void   orc_displacement(const OrcParams* p, const double* Ri, const double* Rj, int64_t k, double* out); /* :27-36 */
double orc_distance(const OrcParams* p, const double* Ri, const double* Rj);          /* :38-48 */
void   orc_distances(const OrcParams* p, const double* a, int64_t na, const double* b, int64_t nb, double* out); /* :54-62 */

/* Bins.py */
void     orc_bin_index(const OrcParams* p, double x, double y, int32_t* nx, int32_t* ny);        /* :9-22  */
OrcBins* orc_put_in_bins(const OrcParams* p, const double* R, int64_t n);                        /* :24-42 */
void     orc_bins_free(OrcBins* b);
int64_t  orc_bins_count(const OrcBins* b);
/* flat cell id nx*nbins_y+ny of every atom (or -1 when outside the addressable grid) */
void     orc_bins_cell_of_atom(const OrcBins* b, int32_t* cell_out);
void     orc_bins_cell_counts(const OrcBins* b, int32_t* counts_out /* nbins_x*nbins_y */);
int      orc_nearby_bins(const OrcParams* p, double x, double y, int32_t* xs /*[4]*/, int32_t* ys /*[3]*/); /* :44-70, returns len(xs) */
int64_t  orc_nearby_atoms(const OrcParams* p, const OrcBins* b, double x, double y, double* out /* may be NULL */); /* :72-84, returns atom count */
/* Bins.py:86-111. counts always; CSR position lists when the pointers are non-NULL. */
void     orc_nearest_neighbors(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins,
                               int32_t* n_a, int32_t* n_s,
                               double* a_vals, int64_t* a_offs, double* s_vals, int64_t* s_offs);

/* Energy.py */
double orc_pair_potential(double r, double E, double r0);                                         /* :10-29 */
void   orc_adatom_adatom_forces(const OrcParams* p, const double* ad, int64_t n, double* F);      /* :62-67 */
void   orc_adatom_substrate_forces(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins, double* F); /* :69-75 */
double orc_adatom_surface_energy(const OrcParams* p, const double* Ri, const OrcBins* sbins);     /* :77-95 */
double orc_relax_energy(const OrcParams* p, const double* xn, int64_t n, const OrcBins* sbins);   /* :133-148 */
double orc_total_energy(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins);   /* :121-131 */
double orc_hopping_energy(const OrcParams* p, const double* pos, const double* ad, int64_t n, const OrcBins* sbins); /* :170-172 */
/* Energy.py:174-217 for every adatom at once; any output may be NULL */
void   orc_barriers(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins,
                    double* u_appx, double* u_loc, double* u_ideal, double* delta_u, int32_t* n_a, int32_t* n_s);

/* 2DGrowth.py rate table + selection */
/* :46-62,201-215. rate_dense[i] = 0 for non-surface atoms; returns S = number of surface atoms */
int64_t orc_hopping_rates(const OrcParams* p, const double* ad, int64_t n, const OrcBins* sbins,
                          double* rate_dense, uint8_t* is_surface, double* delta_u);
/* :217-227,322-329. returns the reference's compacted hop index k (position in the surface list)
 * or -1 for a deposit; *dense_idx = adatom index of that surface atom; *rtot = Rd + sum(Rk). */
int64_t orc_select(const OrcParams* p, const double* rate_dense, int64_t n, double u, int64_t* dense_idx, double* rtot);

/* 2DGrowth.py events (callers of the hot path) */
typedef struct OrcRelaxStats { int64_t line_searches; int64_t restarts; int64_t pair_evals; double maxF; int32_t final_scale; double final_threshold; } OrcRelaxStats;
/* :97-166. x is relaxed in place. returns 0 ok, <0 on iteration cap */
int  orc_relaxation(const OrcParams* p, double* x, int64_t n, const OrcBins* sbins, OrcRelaxStats* st, int64_t max_line_searches);
int  orc_relaxation_ex(const OrcParams* p, double* x, int64_t n, const OrcBins* sbins, OrcRelaxStats* st, int64_t max_line_searches,
                       int scale0, double threshold0);
/* :64-95 without the trailing Relaxation: appends the new atom to ad (capacity n+1). */
int  orc_deposit_place(const OrcParams* p, double* ad, int64_t n, const OrcBins* sbins, double u);
/* :229-255 without the trailing Relaxation: removes atom k and appends it at the chosen site. */
int  orc_hop_place(const OrcParams* p, double* ad, int64_t n, const OrcBins* sbins, int64_t k, double u);
/* one iteration of the loop :317-329 (rates -> select -> hop/deposit -> relaxation).
 * us: uniform stream (u for selection, then one more for the hop/deposit). *n is updated.
 * kind_out: 0 deposit, 1 hop. returns 0 ok. */
int  orc_event(const OrcParams* p, double* ad, int64_t* n, const OrcBins* sbins, const double* us,
               int32_t* kind_out, int64_t* hop_k_out, double* rtot_out, OrcRelaxStats* st, int64_t max_line_searches);

/* bench helpers: thread count used by the OpenMP sweeps (1 when built without OpenMP) */
int  orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
