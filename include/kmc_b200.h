/*
 * kmc_b200.h -- C-ABI of libkmcb200.so: the B200 (sm_100a) implementation of the data-parallel
 * hot path of YouCantCodeWithUs/KMCIslandGrowth.
 *
 * The reference has no FFI: its boundary is the module-level Python API of Periodic.py, Bins.py,
 * Energy.py (plus the rate-table / selection / event functions of 2DGrowth.py).  Each entry point
 * below names the reference function(s) it replaces (file:line in the reference repository).
 * The Python drop-in modules kmcislandgrowth_b200/{Periodic,Bins,Energy,Growth}.py bind these
 * symbols through ctypes.
 *
 * Conventions
 *  - every function returns 0 on success, a negative KMC_E_* code otherwise; kmc_last_error()
 *    returns a thread-local message.  No exception crosses the boundary.
 *  - positions are the reference's layout: interleaved fp64 [x0,y0,x1,y1,...] (2DGrowth.py:36,93).
 *  - `mem` says where EVERY array argument of that call lives: KMC_HOST (plain or pinned host
 *    memory; the call copies in/out and synchronises before returning) or KMC_DEVICE (device
 *    pointers, e.g. torch.Tensor.data_ptr(); the call only enqueues work on the context's stream).
 *  - a context is bound to one device, owns one stream (kmc_set_stream) and is not thread-safe.
 *  - there is no CPU fallback anywhere in this library.
 */
#ifndef KMC_B200_H
#define KMC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMC_HOST 0
#define KMC_DEVICE 1

#define KMC_OK 0
#define KMC_E_CUDA (-1)        /* a CUDA runtime call failed                                  */
#define KMC_E_ARG (-2)         /* bad argument                                                */
#define KMC_E_NODEVICE (-3)    /* no CUDA device / wrong architecture                         */
#define KMC_E_CAPACITY (-4)    /* output buffer too small                                     */
#define KMC_E_NOCANDIDATE (-5) /* HopAtom found no surface atom within 3 r_a (reference: randint(0,-1) raises, 2DGrowth.py:248) */
#define KMC_E_ITERCAP (-6)     /* relaxation hit the caller's line-search cap                 */
#define KMC_E_TIMEOUT (-8)     /* a device-side wait of the persistent relaxation kernel timed out (watchdog) */
#define KMC_E_DEPOSIT (-7)     /* deposition probe never came close to anything (2DGrowth.py:80-92 would loop forever) */

/* Values of the reference's Constants.py (and the two literals of 2DGrowth.py), computed on the
 * host with the reference's own expressions and passed down verbatim; nothing is re-derived on
 * the device (nbins_x = int(ceil(L/bin_size)) is floating-point fragile). */
typedef struct KmcParams {
    double L;          /* Constants.py:35 */
    double bin_size;   /* Constants.py:39 */
    double Dmin;       /* Constants.py:36 */
    double Dmax;       /* Constants.py:37 */
    double E_a;        /* Constants.py:16 */
    double r_a;        /* Constants.py:20 */
    double E_as;       /* Constants.py:28 */
    double r_as;       /* Constants.py:29 */
    double beta;       /* Constants.py:45 */
    double omega;      /* 2DGrowth.py:210 */
    double Rd;         /* 2DGrowth.py:43  */
    int32_t nbins_x;   /* Constants.py:40 */
    int32_t nbins_y;   /* Constants.py:41 */
    int32_t aa_mode;   /* 0: adatom sums over ALL adatoms (reference, Energy.py:62-67,115-117,172,181);
                          1: over the 3x3(+seam) stencil of an adatom cell list (intent of Energy.py:118) */
    int32_t precision; /* 0: fp64 everywhere; 1: fp32 pair terms on fp64-formed displacements, fp64 accumulation */
    int32_t hop_index_mode; /* 0: reference quirk, the surface-list index is used as adatom index (2DGrowth.py:327,240); 1: mapped */
    int32_t deposit_mode;   /* 0: reference (start at 2 r_as + min_y, 2DGrowth.py:73-77); 1: start above the highest adatom */
} KmcParams;

typedef struct KmcCtx KmcCtx;
typedef struct KmcBins KmcBins;   /* a device-resident cell list == the object Bins.PutInBins returns (Bins.py:24-42) */

/* per-relaxation counters (2DGrowth.py:97-166) */
typedef struct KmcRelaxStats {
    int64_t line_searches;   /* 20-candidate line searches executed                        */
    int64_t restarts;        /* recursive restarts (stall / out of bounds)                 */
    int64_t pair_evals;      /* ordered pair visits, SURVEY.md §8d definition              */
    double  maxF;            /* last max_i |F_i|^2                                         */
    double  final_threshold;
    int32_t final_scale;
    int32_t status;          /* 0 converged, KMC_E_ITERCAP                                 */
} KmcRelaxStats;

const char* kmc_last_error(void);
int kmc_version(void);

/* ---- context ---- */
int kmc_create(const KmcParams* p, int device, KmcCtx** out);
int kmc_destroy(KmcCtx* ctx);
int kmc_set_params(KmcCtx* ctx, const KmcParams* p);          /* Constants.py values changed */
int kmc_get_params(KmcCtx* ctx, KmcParams* p);
int kmc_set_stream(KmcCtx* ctx, void* cuda_stream);            /* default: a stream created by the context */
int kmc_synchronize(KmcCtx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t kmc_launch_count(KmcCtx* ctx);
/* CUDA-event timing of this library's kernels on its own stream: enable, run, read back
 * (name filter: substring of the kernel family, "" = all) -> total ms and launches */
int kmc_timing_enable(KmcCtx* ctx, int on);
int kmc_timing_read(KmcCtx* ctx, const char* family, double* total_ms, int64_t* launches);
int kmc_timing_reset(KmcCtx* ctx);

/* ---- Periodic.py ---- */
/* Periodic.py:22-25 PutAllInBox (in place, x entries only; bit-exact Python float modulo) */
int kmc_put_all_in_box(KmcCtx* ctx, double* xy, int64_t n, int mem);
/* Periodic.py:54-62 Distances: out[na*nb] row-major */
int kmc_distances(KmcCtx* ctx, const double* a, int64_t na, const double* b, int64_t nb, double* out, int mem);
/* Periodic.py:27-36 Displacement for k pairs: out = wrap_x(Rj - Ri) */
int kmc_displacement(KmcCtx* ctx, const double* Ri, const double* Rj, int64_t k, double* out, int mem);

/* ---- Bins.py ---- */
/* Bins.py:9-22 BinIndex for n points: nxny_out[2n] (bit-exact: IEEE add, IEEE divide, truncate) */
int kmc_bin_index(KmcCtx* ctx, const double* xy, int64_t n, int32_t* nxny_out, int mem);
/* Bins.py:24-42 PutInBins: builds a device cell list (column-major [nx][ny], input order kept inside a cell) */
int kmc_bins_create(KmcCtx* ctx, const double* xy, int64_t n, int mem, KmcBins** out);
int kmc_bins_destroy(KmcCtx* ctx, KmcBins* bins);
int64_t kmc_bins_count(const KmcBins* bins);
/* read a cell list back: cell_of_atom[n] (flat nx*nbins_y+ny, -1 outside the grid), cell_start[nbins_x*nbins_y+2]
 * (last cell = atoms outside the grid), sorted_xy[2n], sorted_idx[n]; any pointer may be NULL */
int kmc_bins_export(KmcCtx* ctx, const KmcBins* bins, int32_t* cell_of_atom, int32_t* cell_start,
                    double* sorted_xy, int32_t* sorted_idx, int mem);
/* Bins.py:72-84 NearbyAtoms for nq query points: atoms of the 3x3(+seam) stencil in the reference's order.
 * offsets[nq+1] in atoms; out_xy capacity cap_atoms (KMC_E_CAPACITY if too small; offsets are still valid) */
int kmc_nearby_atoms(KmcCtx* ctx, const KmcBins* bins, const double* qxy, int64_t nq,
                     double* out_xy, int64_t cap_atoms, int64_t* offsets, int mem);
/* Bins.py:86-111 NearestNeighbors: bond counts n_a (INCLUDING the atom itself, Bins.py:100) and n_s;
 * optional position lists (CSR offsets in atoms; capacity in atoms; pass NULL to skip) */
int kmc_neighbors(KmcCtx* ctx, const double* xy, int64_t n, const KmcBins* sbins, int32_t* n_a, int32_t* n_s,
                  double* a_xy, int64_t a_cap, int64_t* a_offs, double* s_xy, int64_t s_cap, int64_t* s_offs, int mem);

/* ---- Energy.py ---- */
#define KMC_F_AA 1   /* Energy.py:62-67 AdatomAdatomForces     */
#define KMC_F_AS 2   /* Energy.py:69-75 AdatomSubstrateForces  */
/* which = KMC_F_AA | KMC_F_AS -> their elementwise sum as formed at 2DGrowth.py:120 */
int kmc_forces(KmcCtx* ctx, const double* xy, int64_t n, const KmcBins* sbins, int which, double* f_out, int mem);
/* Energy.py:133-148 RelaxEnergy for B configurations of n adatoms each: xy[B][2n] -> e_out[B].
 * (one launch replaces one pool.map(Energy.RelaxEnergy, xs), 2DGrowth.py:122,153) */
int kmc_relax_energy(KmcCtx* ctx, const double* xy, int64_t n, int64_t B, const KmcBins* sbins, double* e_out, int mem);
/* the 20 line-search candidates of 2DGrowth.py:121,152 formed on the device: x + a*s/20/scale, a=0..K-1 */
int kmc_line_energies(KmcCtx* ctx, const double* xy, const double* dir, int64_t n, int32_t scale, int32_t K,
                      const KmcBins* sbins, double* e_out, int mem);
/* Energy.py:121-131 TotalEnergy (quirk kept: loop stops at N-2) */
int kmc_total_energy(KmcCtx* ctx, const double* xy, int64_t n, const KmcBins* sbins, double* e_out, int mem);
/* Energy.py:170-172 HoppingEnergy / :77-95 AdatomSurfaceEnergy for m probe points against n adatoms (n may be 0):
 * e_aa[m], e_as[m] (sum = HoppingEnergy); min_d_a[m], min_d_s[m] = min distance to stencil adatoms / substrate
 * (the closeness test of DepositAdatom, 2DGrowth.py:82-90; +inf when the stencil is empty). Outputs may be NULL. */
int kmc_probe(KmcCtx* ctx, const double* probes_xy, int64_t m, const double* xy, int64_t n, const KmcBins* sbins,
              double* e_aa, double* e_as, double* min_d_a, double* min_d_s, int mem);
/* Energy.py:174-217 U_appx, U_loc, U_loc_ideal, DeltaU + 2DGrowth.py:46-62,201-215 surface flag and
 * rate = omega*exp(DeltaU*beta) for EVERY adatom (rate 0 for non-surface atoms). Outputs may be NULL. */
int kmc_rates(KmcCtx* ctx, const double* xy, int64_t n, const KmcBins* sbins, double* rate_dense, uint8_t* is_surface,
              double* delta_u, double* u_appx, double* u_loc, int32_t* n_a, int32_t* n_s, int mem);

/* ---- 2DGrowth.py: selection ---- */
/* 2DGrowth.py:217-227,322-329: pj = [0,cumsum(Rk)] summed left to right exactly like the reference,
 * Rtot = Rd + sum(Rk), r = u*Rtot, first pj > r.  result[0] = compacted hop index k (position in the surface
 * list) or -1 for a deposit, result[1] = adatom index of that surface atom or -1; rtot_out[0] = Rtot.
 * pj_out (optional, n+1 entries, dense) for HoppingPartialSums. */
int kmc_select(KmcCtx* ctx, const double* rate_dense, const uint8_t* is_surface, int64_t n, double u,
               int64_t* result, double* rtot_out, double* pj_out, int mem);

/* 0 (default): sequential-order sums (bit-exact index) up to 65536 adatoms, block scan + search beyond;
 * 1: always sequential order; 2: always block scan (partial sums associated differently: the index can differ
 * from the reference's only when r lies within rounding of a partial sum). */
int kmc_set_select_mode(KmcCtx* ctx, int mode);

/* ---- 2DGrowth.py: events on a device-resident trajectory ---- */
/* upload / download the adatom array of the context's trajectory (capacity grows on demand) */
int kmc_state_set(KmcCtx* ctx, const double* xy, int64_t n, int mem);
int kmc_state_get(KmcCtx* ctx, double* xy, int64_t cap_atoms, int64_t* n_out, int mem);
/* new positions for the SAME atoms (n must equal the state's count): unlike kmc_state_set this keeps the history of the
 * incremental rate table, i.e. the next kmc_rates_update redoes only what these moves touched */
int kmc_state_move(KmcCtx* ctx, const double* xy, int64_t n, int mem);
int64_t kmc_state_count(KmcCtx* ctx);
/* Incremental rate table of the resident trajectory (north star item 3; the reference rebuilds the whole list on every
 * event, 2DGrowth.py:201-215,318).  The context remembers, per adatom, the rate / surface flag / barrier and the position
 * they were computed at; an update recomputes only the adatoms of cells whose 3x3(+seam) stencil (Bins.py:44-70) holds an
 * atom that moved by more than `threshold` Angstrom since then, was added (DepositAdatom :93, HopAtom :255) or removed
 * (HopAtom :241 -- the cached entries follow the index shift).  threshold = 0 reproduces the full sweep bit for bit;
 * all-pairs adatom sums (aa_mode 0) always fall back to the full sweep.  Outputs are dense, like kmc_rates; any may be NULL.
 * kmc_set_rate_mode(ctx, 1, threshold) makes kmc_event use the incremental table (default 0: full sweep). */
int kmc_rates_update(KmcCtx* ctx, const KmcBins* sbins, double threshold, double* rate_dense, uint8_t* is_surface, double* delta_u,
                     int64_t* n_recomputed, int mem);
int kmc_set_rate_mode(KmcCtx* ctx, int mode, double threshold);
/* 2DGrowth.py:97-166 Relaxation(adatoms, substrate_bins, scale=16, threshold=1e-4) of the trajectory state:
 * 20-point line search + pseudo-CG + the reference's restart logic.  max_line_searches <= 0: no cap. */
int kmc_relax(KmcCtx* ctx, const KmcBins* sbins, int32_t scale0, double threshold0, int64_t max_line_searches, KmcRelaxStats* stats);
/* 2 (default): the relaxation runs as one persistent cooperative kernel on cell-sorted state; 1: first-generation
 * persistent kernel; 0: host-driven loop of per-phase launches (same arithmetic, kept as cross-checks).
 * Environment KMC_RELAX_MODE=host|persistent1 selects 0|1 at context creation. */
int kmc_set_relax_mode(KmcCtx* ctx, int mode);
/* 2DGrowth.py:64-93 DepositAdatom placement (u = the uniform of :76), without the relaxation */
int kmc_deposit_place(KmcCtx* ctx, const KmcBins* sbins, double u);
/* 2DGrowth.py:229-255 HopAtom placement (u -> randint of :248 as int(u*n)), without the relaxation */
int kmc_hop_place(KmcCtx* ctx, const KmcBins* sbins, int64_t k, double u);
/* one iteration of the loop 2DGrowth.py:317-329: rates -> select(u0) -> hop/deposit(u1) -> relaxation.
 * kind_out: 0 deposit, 1 hop; hop_k_out: index handed to HopAtom. */
int kmc_event(KmcCtx* ctx, const KmcBins* sbins, double u0, double u1, int64_t max_line_searches,
              int32_t* kind_out, int64_t* hop_k_out, double* rtot_out, KmcRelaxStats* stats);

/* The event above is queued on the context's stream as ONE chain of kernels (rate table, selection, placement decided
 * on the device, relaxation) with a single host synchronisation at its end.  kmc_set_event_mode(ctx, 0) selects the
 * first-generation host-stepped placement instead (same results; cross-check).  The split form lets one host thread
 * advance several replicas (one context each) concurrently: enqueue all, then finish all. */
int kmc_set_event_mode(KmcCtx* ctx, int mode);
int kmc_event_enqueue(KmcCtx* ctx, const KmcBins* sbins, double u0, double u1, int64_t max_line_searches);
int kmc_event_finish(KmcCtx* ctx, int32_t* kind_out, int64_t* hop_k_out, double* rtot_out, KmcRelaxStats* stats);
/* rate table -> selection (u0) -> placement (u1) without the relaxation (the caller relaxes, e.g. with a domain-decomposed loop) */
int kmc_event_place(KmcCtx* ctx, const KmcBins* sbins, double u0, double u1, int32_t* kind_out, int64_t* hop_k_out, double* rtot_out);
/* 1: the enqueued event has completed (kmc_event_finish returns at once), 0: still running, < 0: error */
int kmc_event_ready(KmcCtx* ctx);
/* temperature x flux sweeps (BASELINE configs[3]; per-replica constants: Constants.py:45-46, 2DGrowth.py:43): one event
 * of each of R replicas, all queued before the first is waited for.  Arrays have R entries; rc_out[r] = status of
 * replica r.  Returns the first non-zero status other than KMC_E_ITERCAP, else 0. */
int kmc_events_batch(KmcCtx** ctxs, KmcBins** sbins, int32_t R, const double* u0, const double* u1, int64_t max_line_searches,
                     int32_t* kind_out, int64_t* hop_k_out, double* rtot_out, KmcRelaxStats* stats, int32_t* rc_out);

/* ---- spatial domain decomposition over bin columns (BASELINE configs[4]; the reference has no counterpart) ----
 * A rank's LOCAL configuration: xy[0 .. n_own) are the adatoms it owns, xy[n_own .. n) halo copies of the neighbouring
 * ranks' boundary columns (Bins.py:54-70: one column each side, the seam pair across the periodic boundary).  Only own
 * atoms are evaluated; the cell grid is the global one.  bins3x3 adatom sums only (aa_mode 1).
 * kmc_domain_sweep: forces F_aa + F_as (Energy.py:62-75), site energies e_aa (stencil sum, self term 0) / e_as
 * (Energy.py:174-181), rate and surface flag (2DGrowth.py:46-62,201-215) of the own atoms from ONE cell-list build;
 * any output may be NULL.  kmc_domain_line_energies: this rank's share of RelaxEnergy (Energy.py:133-148) for the K
 * candidates of a line search (2DGrowth.py:121,152); the all-reduce over the ranks gives the candidate energies. */
int kmc_domain_sweep(KmcCtx* ctx, const double* xy, int64_t n, int64_t n_own, const KmcBins* sbins, double* f_out, double* e_aa,
                     double* e_as, double* rate, uint8_t* is_surface, int mem);
int kmc_domain_line_energies(KmcCtx* ctx, const double* xy, const double* dir, int64_t n, int64_t n_own, int32_t scale, int32_t K,
                             const KmcBins* sbins, double* e_out, int mem);

/* Measurement helpers (bench.py).  kmc_fp64_peak: non-tensor FP64 FMA rate of the device in TFLOP/s (register-resident
 * DFMA chains at full occupancy, best of `repeats` runs, CUDA events) -- the denominator of the pair kernels' FP-pipe
 * roofline.  kmc_relax_diag: counters of the context's last relaxation: out[0] re-sorts of the
 * cell-sorted state, out[1] movers summed over its line searches, out[2] work items; out[3] = rate-table entries
 * recomputed by the last kmc_rates_update / thresholded kmc_event. */
int kmc_fp64_peak(KmcCtx* ctx, int32_t repeats, double* tflops_out, double* ms_out);
int kmc_relax_diag(KmcCtx* ctx, int64_t out[4]);
/* kmc_barrier_bench: microseconds per device-wide barrier of one 512-thread block per SM that publishes one word per block
 * (k_relax2 pays three per line search).  variant 0 = the barrier k_relax2 uses (red.release.gpu + ld.acquire.gpu poll by
 * thread 0 between two bar.sync), variant 1 = hierarchical arrival through thread-block clusters of `cluster` blocks.
 * stale_out != 0 would mean a block read another block's word of an earlier round (never observed). */
int kmc_barrier_bench(KmcCtx* ctx, int32_t variant, int32_t cluster, int32_t iters, double* us_out, int32_t* stale_out);

/* The persistent adatom cell list of the trajectory state (replaces the reference's PutInBins of the whole adatom array
 * per neighbour query, Bins.py:91): kmc_rates_update and the thresholded kmc_event edit it in place -- only atoms whose
 * cell id changed, the atom HopAtom deleted (2DGrowth.py:241) and appended atoms (:93,:255) are re-filed -- and fall
 * back to a rebuild beyond 64 edits.  kmc_cells_stats: out[0] full builds, out[1] in-place updates, out[2] edits of the
 * last update (atoms, after a full build), out[3] result of the last kmc_cells_check.  kmc_cells_check brings the list
 * up to date, builds the same list from scratch and returns the number of words (starts, indices, position bit
 * patterns) in which the two differ: 0 by construction (tests). */
int kmc_cells_stats(KmcCtx* ctx, int64_t out[4]);
int kmc_cells_check(KmcCtx* ctx, int64_t* differing_words);

#ifdef __cplusplus
}
#endif
#endif
