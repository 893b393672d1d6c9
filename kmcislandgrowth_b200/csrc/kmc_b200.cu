// kmc_b200.cu -- host side of libkmcb200.so: context, device buffers, kernel launches, C-ABI.
// See include/kmc_b200.h for the contract of every entry point and the reference lines it replaces.
// (The persistent relaxation kernels live in kmc_relax1.cu / kmc_relax2.cu, the on-device event path in kmc_event.cu.)
#include "kmc_host.h"
#include "kmc_kernels.cuh"
#include "kmc_line_kernels.cuh"

#include <cmath>
#include <algorithm>
#include <cstring>
#include <map>

using namespace kmc;

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

int kmc_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define fail kmc_fail

int kmc_ensure(KmcCtx* c, DBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return 0;
    if (b.p) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    // Growing a slot frees and reallocates it, and cudaFree waits for EVERY stream of the device: with a dozen replicas
    // in flight each growth step stalled the submitting thread for milliseconds (measured: 95 % of the host time of a
    // 64-replica sweep).  Slots therefore start at 64 KB -- small systems never grow them -- and grow by 2x.
    size_t want = std::max<size_t>(2 * bytes + 256, 65536);
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}

// per-call staging of host arguments
struct Pending {
    void* host;
    const void* dev;
    size_t bytes;
};
struct Call {
    KmcCtx* c;
    int mem;
    DevGuard guard;
    std::vector<Pending> outs;
    Call(KmcCtx* ctx, int m) : c(ctx), mem(m), guard(ctx) {}
    template <class T>
    int in(const T* p, size_t count, int slot_id, const T** dev) {
        if (mem == KMC_DEVICE || p == nullptr) { *dev = p; return 0; }
        T* d;
        TRY(slot(c, slot_id, count, &d));
        if (count) CK(cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
        *dev = d;
        return 0;
    }
    template <class T>
    int out(T* p, size_t count, int slot_id, T** dev) {
        if (mem == KMC_DEVICE || p == nullptr) { *dev = p; return 0; }
        T* d;
        TRY(slot(c, slot_id, count, &d));
        outs.push_back({(void*)p, d, count * sizeof(T)});
        *dev = d;
        return 0;
    }
    int finish() {
        if (mem == KMC_DEVICE) return 0;
        for (auto& o : outs)
            if (o.bytes) CK(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
};

// ------------------------------------------------------------------------------------ launches
int kmc_family_id(KmcCtx* c, const char* name) {
    for (size_t i = 0; i < c->families.size(); ++i)
        if (c->families[i] == name) return (int)i;
    c->families.push_back(name);
    return (int)c->families.size() - 1;
}

// timing events are recycled: kmc_timing_reset returns them to the pool instead of destroying them
cudaEvent_t kmc_event_from_pool(KmcCtx* c) {
    cudaEvent_t e = nullptr;
    if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}

static DevParams to_dev(const KmcParams& p) {
    DevParams d{};
    d.L = p.L; d.halfL = p.L / 2; d.bin_size = p.bin_size; d.Dmin = p.Dmin; d.Dmax = p.Dmax;
    d.E_a = p.E_a; d.r_a = p.r_a; d.E_as = p.E_as; d.r_as = p.r_as;
    d.ra2 = p.r_a * p.r_a; d.ras2 = p.r_as * p.r_as;
    d.inv_ra2 = 1.0 / d.ra2; d.inv_ras2 = 1.0 / d.ras2;
    d.bond_a = 1.2 * p.r_a; d.bond_s = 1.2 * p.r_as;
    d.beta = p.beta; d.omega = p.omega; d.Rd = p.Rd;
    d.nbx = p.nbins_x; d.nby = p.nbins_y; d.ncell = p.nbins_x * p.nbins_y;
    d.aa_mode = p.aa_mode;
    d.precision = p.precision;
    return d;
}

static int check_params(const KmcParams* p) {
    if (!p) return fail(KMC_E_ARG, "params is NULL");
    if (!(p->L > 0) || !(p->bin_size > 0)) return fail(KMC_E_ARG, "L and bin_size must be positive");
    if (p->nbins_x < 3 || p->nbins_y < 1) return fail(KMC_E_ARG, "need nbins_x >= 3 and nbins_y >= 1 (got %d x %d)", p->nbins_x, p->nbins_y);
    if ((int64_t)p->nbins_x * p->nbins_y > (1 << 28)) return fail(KMC_E_ARG, "cell grid too large");
    if (p->aa_mode != 0 && p->aa_mode != 1) return fail(KMC_E_ARG, "aa_mode must be 0 or 1");
    if (p->precision != 0 && p->precision != 1) return fail(KMC_E_ARG, "precision must be 0 (fp64) or 1 (fp32 pair terms)");
    return 0;
}

// ------------------------------------------------------------------------------------ cell lists
// B configurations of n atoms each (d_xy[b*n+i]).
int kmc_build_cells(KmcCtx* c, const double2* d_xy, int n, int B, int* cell_of, int* start, int* cursor, int* sidx, double2* sxy) {
    const int ncell2 = c->dp.ncell + 2;
    TRY(kmc_zero_async(c, start, sizeof(int) * (size_t)ncell2 * B));
    if (n > 0) {
        dim3 g(std::min(cdiv(n, 256), 1024u), B);
        LAUNCH(c, "bin_count", k_bin_count, g, 256, c->dp, d_xy, n, B, cell_of, start);
    }
    if (B == 1 && ncell2 > 16384) {      // whole-device tiled scan for large grids
        const int m = ncell2 - 1;
        const int tiles = (m + SCAN_TILE - 1) / SCAN_TILE;
        int* tile_tot;
        TRY(slot(c, S_AMIN, (size_t)tiles, &tile_tot));
        LAUNCH(c, "bin_scan", k_bin_scan_tiles, tiles, 1024, m, start, tile_tot);
        LAUNCH(c, "bin_scan", k_bin_scan_add, tiles, 1024, m, start, tile_tot, cursor);
    } else {
        LAUNCH(c, "bin_scan", k_bin_scan, B, 1024, ncell2, start, cursor);
    }
    if (n > 0) {
        dim3 g(std::min(cdiv(n, 256), 1024u), B);
        LAUNCH(c, "bin_scatter", k_bin_scatter, g, 256, n, B, ncell2, cell_of, cursor, sidx);
        dim3 g2(std::min(cdiv(ncell2 - 1, 128), 4096u), B);
        LAUNCH(c, "bin_order", k_bin_order, g2, 128, n, B, ncell2, start, sidx, d_xy, sxy);
    }
    CK(cudaGetLastError());
    return 0;
}

// scratch cell lists for B configurations of the adatoms
int kmc_scratch_cells(KmcCtx* c, const double2* d_xy, int n, int B, CellView* out) {
    const int ncell2 = c->dp.ncell + 2;
    int *cell_of, *start, *cursor, *sidx;
    double2* sxy;
    TRY(slot(c, S_CELLOF, (size_t)n * B, &cell_of));
    TRY(slot(c, S_START, (size_t)ncell2 * B, &start));
    TRY(slot(c, S_CURSOR, (size_t)ncell2 * B, &cursor));
    TRY(slot(c, S_SIDX, (size_t)n * B, &sidx));
    TRY(slot(c, S_SXY, (size_t)n * B, &sxy));
    TRY(kmc_build_cells(c, d_xy, n, B, cell_of, start, cursor, sidx, sxy));
    out->start = start; out->sxy = sxy; out->sidx = sidx; out->n = n;
    return 0;
}

// identity view (aa_mode 0 needs no adatom cell list for forces / energies)
static CellView identity_view(const double2* d_xy, int n) {
    CellView v;
    v.start = nullptr; v.sxy = d_xy; v.sidx = nullptr; v.n = n;
    return v;
}

static int pick_group(size_t atoms) {
    // enough lanes to fill 148 SMs x 2048 threads when the problem is small
    if (atoms * 32 <= 400000) return 32;
    if (atoms * 16 <= 400000) return 16;
    if (atoms * 8 <= 400000) return 8;
    return 4;
}

// ------------------------------------------------------------------------------------ internals (device pointers)
static int forces_dev(KmcCtx* c, CellView ad, const int* which_cfg, const KmcBins* sb, int n, double2* faa, double2* fas,
                      double2* fsum, int n_compute = -1) {
    if (n == 0) return 0;
    const int G = pick_group(n);
    const int ncell2 = c->dp.ncell + 2;
    const unsigned grid = cdiv((size_t)n * G, 256);
    CellView sub = view_of(sb);
    switch (G) {
        case 32: LAUNCH(c, "sweep_forces", k_sweep_forces<32>, grid, 256, c->dp, ad, sub, n, which_cfg, ncell2, faa, fas, fsum, n_compute); break;
        case 16: LAUNCH(c, "sweep_forces", k_sweep_forces<16>, grid, 256, c->dp, ad, sub, n, which_cfg, ncell2, faa, fas, fsum, n_compute); break;
        case 8: LAUNCH(c, "sweep_forces", k_sweep_forces<8>, grid, 256, c->dp, ad, sub, n, which_cfg, ncell2, faa, fas, fsum, n_compute); break;
        default: LAUNCH(c, "sweep_forces", k_sweep_forces<4>, grid, 256, c->dp, ad, sub, n, which_cfg, ncell2, faa, fas, fsum, n_compute); break;
    }
    CK(cudaGetLastError());
    return 0;
}

static int energy_dev(KmcCtx* c, CellView ad, const KmcBins* sb, int n, int B, double* d_e, int n_compute = -1) {
    if (n == 0) { CK(cudaMemsetAsync(d_e, 0, sizeof(double) * B, c->stream)); return 0; }
    const int G = pick_group((size_t)n * B);
    const int ncell2 = c->dp.ncell + 2;
    const unsigned gx = cdiv((size_t)n * G, 256);
    double* partial;
    unsigned int* done;
    TRY(slot(c, S_PARTIAL, (size_t)gx * B, &partial));
    done = (unsigned int*)c->slots[S_DONE].p;      // 65536 zeroed counters allocated by kmc_create
    dim3 grid(gx, B);
    CellView sub = view_of(sb);
    switch (G) {
        case 32: LAUNCH(c, "sweep_energy", k_sweep_energy<32>, grid, 256, c->dp, ad, sub, n, ncell2, partial, done, d_e, n_compute); break;
        case 16: LAUNCH(c, "sweep_energy", k_sweep_energy<16>, grid, 256, c->dp, ad, sub, n, ncell2, partial, done, d_e, n_compute); break;
        case 8: LAUNCH(c, "sweep_energy", k_sweep_energy<8>, grid, 256, c->dp, ad, sub, n, ncell2, partial, done, d_e, n_compute); break;
        default: LAUNCH(c, "sweep_energy", k_sweep_energy<4>, grid, 256, c->dp, ad, sub, n, ncell2, partial, done, d_e, n_compute); break;
    }
    CK(cudaGetLastError());
    return 0;
}

// 2DGrowth.py:121-122,152-153: energies of the KC = 20 candidates x + a*s/20/scale from one base cell list
// (`ad` = cell list of x).  d_e[KC] and *d_amin (first argmin) are device pointers; d_amin may be null.
static const int LINE_BLOCKS = 148 * 8;
static int line_energies_dev(KmcCtx* c, const double2* d_x, const double2* d_s, int n, double scale, CellView ad, const KmcBins* sb,
                             double* d_e, int* d_amin, int n_compute = -1) {
    LineArgs A{};
    A.n_compute = n_compute;
    A.P = c->dp; A.ad = ad; A.sub = view_of(sb); A.x = d_x; A.s = d_s; A.scale = scale; A.inv_step = 1.0 / (20.0 * scale); A.n = n;
    TRY(slot(c, S_FLAG, (size_t)n, &A.flag));
    TRY(slot(c, S_MOVERS, (size_t)n, &A.movers));
    TRY(slot(c, S_NMOV, 4, &A.n_movers));
    TRY(slot(c, S_LPART, (size_t)LINE_BLOCKS * KC, &A.partial));
    TRY(slot(c, S_FIX, (size_t)n * KC, &A.fix));
    A.e_out = d_e; A.amin_out = d_amin;
    A.colrange = nullptr;
    // sparse tall grid (configs[4]: 198 cells per adatom) and the scratch cell list of this call: occupied rows per column.  The
    // home cells are then dealt out by COLUMN, which starves the grid on a dense film (86 columns for 1184 blocks on configs[1]).
    if (c->dp.nby > 64 && (size_t)c->dp.ncell > 16 * (size_t)n && ad.start == (const int*)c->slots[S_START].p) {
        int2* cr;
        TRY(slot(c, S_COLR, (size_t)c->dp.nbx, &cr));
        LAUNCH(c, "line_mark", k_colrange_init, cdiv(c->dp.nbx, 256), 256, cr, c->dp.nbx);
        LAUNCH(c, "line_mark", k_colrange, cdiv(n, 256), 256, c->dp, (const int*)c->slots[S_CELLOF].p, n, cr);
        A.colrange = cr;
    }
    LAUNCH(c, "line_mark", k_line_mark, std::min(cdiv(n, 256), 1184u), 256, A);
    LAUNCH(c, "line_compact", k_line_compact, 1, 1024, A);
    LAUNCH(c, "line_energy", k_line_energy, LINE_BLOCKS, LT_THREADS, A);
    LAUNCH(c, "line_fix", k_line_fix, 296, 128, A);
    LAUNCH(c, "line_reduce", k_line_reduce, 1, KC * 32, A, LINE_BLOCKS);
    CK(cudaGetLastError());
    return 0;
}

static int adatom_view(KmcCtx* c, const double2* d_xy, int n, int B, bool need_cells, CellView* v) {
    if (c->dp.aa_mode == 0 && !need_cells) { *v = identity_view(d_xy, n); return 0; }
    return kmc_scratch_cells(c, d_xy, n, B, v);
}

int kmc_rates_dev(KmcCtx* c, const double2* d_xy, int n, const KmcBins* sb, double* rate, uint8_t* surf, double* du,
                     double* ua, double* ul, int* na, int* ns) {
    if (n == 0) return 0;
    CellView ad;
    TRY(kmc_scratch_cells(c, d_xy, n, 1, &ad));
    const int G = pick_group(n);
    const unsigned grid = cdiv((size_t)n * G, 256);
    CellView sub = view_of(sb);
    switch (G) {
        case 32: LAUNCH(c, "sweep_rates", k_sweep_rates<32>, grid, 256, c->dp, ad, sub, n, rate, surf, du, ua, ul, na, ns); break;
        case 16: LAUNCH(c, "sweep_rates", k_sweep_rates<16>, grid, 256, c->dp, ad, sub, n, rate, surf, du, ua, ul, na, ns); break;
        case 8: LAUNCH(c, "sweep_rates", k_sweep_rates<8>, grid, 256, c->dp, ad, sub, n, rate, surf, du, ua, ul, na, ns); break;
        default: LAUNCH(c, "sweep_rates", k_sweep_rates<4>, grid, 256, c->dp, ad, sub, n, rate, surf, du, ua, ul, na, ns); break;
    }
    CK(cudaGetLastError());
    return 0;
}

// ---- incremental rate table ------------------------------------------------------------------------------
constexpr int RC_PENDING = 64;
static int rc_reserve(KmcCtx* c, int64_t n) {
    if (n <= c->rc_cap && c->rc_xy) return 0;
    const int64_t cap = std::max<int64_t>(n + n / 2, 1024);
    double2* xy = nullptr; double* rate = nullptr; double* du = nullptr; uint8_t* surf = nullptr;
    CK(cudaStreamSynchronize(c->stream));
    cudaError_t e = cudaMalloc(&xy, sizeof(double2) * (size_t)cap);
    if (e == cudaSuccess) e = cudaMalloc(&rate, sizeof(double) * (size_t)cap);
    if (e == cudaSuccess) e = cudaMalloc(&du, sizeof(double) * (size_t)cap);
    if (e == cudaSuccess) e = cudaMalloc(&surf, (size_t)cap);
    if (e == cudaSuccess && !c->rc_pending) e = cudaMalloc(&c->rc_pending, sizeof(double2) * RC_PENDING);
    if (e == cudaSuccess && c->rc_xy) {
        e = cudaMemcpy(xy, c->rc_xy, sizeof(double2) * (size_t)c->rc_n, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(rate, c->rc_rate, sizeof(double) * (size_t)c->rc_n, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(du, c->rc_du, sizeof(double) * (size_t)c->rc_n, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(surf, c->rc_surf, (size_t)c->rc_n, cudaMemcpyDeviceToDevice);
    }
    if (e != cudaSuccess) {       // nothing half-allocated survives a failure; the old tables stay valid
        cudaFree(xy); cudaFree(rate); cudaFree(du); cudaFree(surf);
        return fail(KMC_E_CUDA, "rate-table cache of %lld atoms: %s", (long long)cap, cudaGetErrorString(e));
    }
    if (c->rc_xy) { cudaFree(c->rc_xy); cudaFree(c->rc_rate); cudaFree(c->rc_du); cudaFree(c->rc_surf); }
    c->rc_xy = xy; c->rc_rate = rate; c->rc_du = du; c->rc_surf = surf; c->rc_cap = cap;
    return 0;
}

// the cached tables follow the index changes of the trajectory state: HopAtom deletes entry k and appends the moved atom
// (2DGrowth.py:241,255), DepositAdatom appends (:93).  A new atom has no cached entry (NaN reference position).
int kmc_rc_on_remove(KmcCtx* c, int k) {
    kmc_cells_on_remove(c, k);
    if (!c->rc_valid) return 0;
    const int n = (int)c->rc_n;
    if (c->rc_npending >= RC_PENDING) { c->rc_valid = false; return 0; }
    CK(cudaMemcpyAsync(c->rc_pending + c->rc_npending, c->st_xy + 2 * (size_t)k, sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
    c->rc_npending++;
    double2* t2; double* t1; uint8_t* t0;
    TRY(slot(c, S_IN1, (size_t)n, &t2));
    LAUNCH(c, "misc_erase", k_erase_at<double2>, cdiv(n, 256), 256, (const double2*)c->rc_xy, n, k, t2);
    CK(cudaMemcpyAsync(c->rc_xy, t2, sizeof(double2) * (size_t)(n - 1), cudaMemcpyDeviceToDevice, c->stream));
    TRY(slot(c, S_IN2, (size_t)n, &t1));
    LAUNCH(c, "misc_erase", k_erase_at<double>, cdiv(n, 256), 256, (const double*)c->rc_rate, n, k, t1);
    CK(cudaMemcpyAsync(c->rc_rate, t1, sizeof(double) * (size_t)(n - 1), cudaMemcpyDeviceToDevice, c->stream));
    LAUNCH(c, "misc_erase", k_erase_at<double>, cdiv(n, 256), 256, (const double*)c->rc_du, n, k, t1);
    CK(cudaMemcpyAsync(c->rc_du, t1, sizeof(double) * (size_t)(n - 1), cudaMemcpyDeviceToDevice, c->stream));
    TRY(slot(c, S_IN3, (size_t)n, &t0));
    LAUNCH(c, "misc_erase", k_erase_at<uint8_t>, cdiv(n, 256), 256, (const uint8_t*)c->rc_surf, n, k, t0);
    CK(cudaMemcpyAsync(c->rc_surf, t0, (size_t)(n - 1), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaGetLastError());
    c->rc_n = n - 1;
    return 0;
}
int kmc_rc_on_append(KmcCtx* c) {
    kmc_cells_on_append(c);
    if (!c->rc_valid) return 0;
    TRY(rc_reserve(c, c->rc_n + 1));
    CK(cudaMemsetAsync(c->rc_xy + c->rc_n, 0xFF, sizeof(double2), c->stream));      // all-ones bit pattern == NaN: "no cached entry"
    c->rc_n++;
    return 0;
}

// bring the cached rate table up to date with the trajectory state; returns the number of atoms recomputed
int kmc_rates_update_impl(KmcCtx* c, const KmcBins* sb, double threshold, int64_t* recomputed) {
    const int n = (int)c->st_n;
    *recomputed = 0;
    if (n == 0) { c->rc_n = 0; c->rc_valid = true; c->rc_npending = 0; return 0; }
    TRY(rc_reserve(c, n));
    const bool full = !c->rc_valid || c->rc_n != n || c->dp.aa_mode == 0;      // all-pairs U_appx: every rate depends on every atom
    if (full) {
        TRY(kmc_rates_dev(c, (const double2*)c->st_xy, n, sb, c->rc_rate, c->rc_surf, c->rc_du, nullptr, nullptr, nullptr, nullptr));
        CK(cudaMemcpyAsync(c->rc_xy, c->st_xy, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        c->rc_n = n; c->rc_valid = true; c->rc_npending = 0;
        *recomputed = n;
        c->rc_last = n;
        return 0;
    }
    uint8_t *dirty_cell, *self_dirty;
    unsigned long long* counter;
    TRY(slot(c, S_OUT4, (size_t)c->dp.ncell + 16, &dirty_cell));
    TRY(slot(c, S_OUT5, (size_t)n, &self_dirty));
    TRY(slot(c, S_OUT6, 1, &counter));
    CK(cudaMemsetAsync(dirty_cell, 0, (size_t)c->dp.ncell + 16, c->stream));
    CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), c->stream));
    LAUNCH(c, "rates_mark", k_rates_mark, cdiv(std::max(n, c->rc_npending), 256), 256, c->dp, (const double2*)c->st_xy, (const double2*)c->rc_xy, n,
           threshold * threshold, (const double2*)c->rc_pending, c->rc_npending, dirty_cell, self_dirty);
    CellView ad;
    TRY(kmc_cells_update(c, &ad));          // the persistent list, edited in place (kmc_cells.cu)
    const int G = pick_group(n);
    const unsigned grid = cdiv((size_t)n * G, 256);
    CellView sub = view_of(sb);
    double* nd = nullptr; int* ni = nullptr;
#define RATES_MASKED(GG) LAUNCH(c, "sweep_rates", k_sweep_rates<GG>, grid, 256, c->dp, ad, sub, n, c->rc_rate, c->rc_surf, c->rc_du, nd, nd, ni, ni, \
                                (const uint8_t*)dirty_cell, (const uint8_t*)self_dirty, c->rc_xy, counter)
    switch (G) {
        case 32: RATES_MASKED(32); break;
        case 16: RATES_MASKED(16); break;
        case 8: RATES_MASKED(8); break;
        default: RATES_MASKED(4); break;
    }
#undef RATES_MASKED
    CK(cudaGetLastError());
    TRY(kmc_mail(c, counter, sizeof(unsigned long long)));
    *recomputed = (int64_t) * (const unsigned long long*)c->h_mail;
    c->rc_last = *recomputed;
    c->rc_npending = 0;
    return 0;
}

int kmc_state_reserve(KmcCtx* c, int64_t n) {
    if (n <= c->st_cap) return 0;
    int64_t cap = std::max<int64_t>(n + n / 2, 1024);
    double* p;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMalloc(&p, sizeof(double) * 2 * (size_t)cap));
    if (c->st_xy) {
        CK(cudaMemcpy(p, c->st_xy, sizeof(double) * 2 * (size_t)c->st_n, cudaMemcpyDeviceToDevice));
        CK(cudaFree(c->st_xy));
    }
    c->st_xy = p;
    c->st_cap = cap;
    return 0;
}

// read `bytes` from device memory into the pinned mailbox and wait
int kmc_mail(KmcCtx* c, const void* dev, size_t bytes) {
    CK(cudaMemcpyAsync(c->h_mail, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// 2DGrowth.py:319-324 on device pointers; the mode switch of kmc_set_select_mode applies to kmc_select AND kmc_event
int kmc_select_dev(KmcCtx* c, const double* dr, const uint8_t* dsf, int n, double u, long long* dres, double* drt, double* dpj) {
    const bool parallel = !dpj && (c->select_mode == 2 || (c->select_mode == 0 && n > 65536));
    if (parallel) {
        const int nblocks = (int)std::min<int64_t>(1024, ((int64_t)n + 4095) / 4096);
        const int chunk = (int)(((int64_t)n + nblocks - 1) / nblocks);
        double* bsum;
        int* bcount;
        TRY(slot(c, S_TMP2, (size_t)nblocks, &bsum));
        TRY(slot(c, S_TMP3, (size_t)nblocks, &bcount));
        LAUNCH(c, "select", k_select_block_sums, nblocks, 1024, dr, dsf, n, chunk, bsum, bcount);
        LAUNCH(c, "select", k_select_final, 1, 1024, c->dp, dr, dsf, n, chunk, nblocks, bsum, bcount, u, dres, drt);
    } else {
        LAUNCH(c, "select", k_select, 1, 32, c->dp, dr, dsf, n, u, dres, drt, dpj);
    }
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------ C-ABI
extern "C" {

const char* kmc_last_error(void) { return g_err; }
int kmc_version(void) { return 100; }

int kmc_create(const KmcParams* p, int device, KmcCtx** out) {
    if (!out) return fail(KMC_E_ARG, "out is NULL");
    TRY(check_params(p));
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(KMC_E_NODEVICE, "no CUDA device available (%s); libkmcb200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(KMC_E_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    CK(cudaSetDevice(device));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int sm_count = prop.multiProcessorCount;
    if (prop.major < 10) return fail(KMC_E_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    KmcCtx* c = new KmcCtx();
    c->device = device;
    c->hp = *p;
    c->dp = to_dev(*p);
    c->slots.resize(S_COUNT);
    c->sm_count = sm_count;
    if (const char* e = getenv("KMC_RELAX_MODE")) c->relax_mode = (strcmp(e, "host") == 0) ? 0 : (strcmp(e, "persistent1") == 0 ? 1 : 2);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail(KMC_E_CUDA, "cudaStreamCreate failed"); }
    c->own_stream = true;
    if (cudaHostAlloc(&c->h_mail, 4096, cudaHostAllocMapped) != cudaSuccess) { kmc_destroy(c); return fail(KMC_E_CUDA, "cudaHostAlloc failed"); }
    if (cudaHostGetDevicePointer(&c->h_mail_dev, c->h_mail, 0) != cudaSuccess) { kmc_destroy(c); return fail(KMC_E_CUDA, "cudaHostGetDevicePointer failed"); }
    memset(c->h_mail, 0, 4096);
    if (cudaMalloc(&c->slots[S_DONE].p, 65536 * sizeof(unsigned)) != cudaSuccess) { kmc_destroy(c); return fail(KMC_E_CUDA, "cudaMalloc failed"); }
    c->slots[S_DONE].cap = 65536 * sizeof(unsigned);
    cudaMemset(c->slots[S_DONE].p, 0, c->slots[S_DONE].cap);
    if (cudaMalloc(&c->st_n_dev, 64) != cudaSuccess) { kmc_destroy(c); return fail(KMC_E_CUDA, "cudaMalloc failed"); }
    cudaMemset(c->st_n_dev, 0, 64);
    if (const char* e = getenv("KMC_EVENT_MODE")) c->event_mode = atoi(e) ? 1 : 0;
    c->debug_sync = getenv("KMC_SYNC_DEBUG") != nullptr;
    *out = c;
    return 0;
}

int kmc_destroy(KmcCtx* c) {
    if (!c) return 0;
    DevGuard guard(c);
    if (getenv("KMC_ENQ_TIMING") && c->enq_s[0] > 0)
        fprintf(stderr, "[kmc enqueue host s] placement %.4f | relax2 total %.4f (buffers+memset %.4f, launch %.4f) | placement parts: slots %.4f init %.4f rates %.4f select %.4f args %.4f deposit %.4f hop+apply %.4f mailbox %.4f\n",
                c->enq_s[0], c->enq_s[1], c->enq_s[2], c->enq_s[3], c->enq_s[4], c->enq_s[5], c->enq_s[6], c->enq_s[7], c->enq_s[8], c->enq_s[9], c->enq_s[10], c->enq_s[11]);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& b : c->slots) if (b.p) cudaFree(b.p);
    for (auto& r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto& e : c->event_pool) cudaEventDestroy(e);
    if (c->st_xy) cudaFree(c->st_xy);
    if (c->st_n_dev) cudaFree(c->st_n_dev);
    if (c->rc_xy) { cudaFree(c->rc_xy); cudaFree(c->rc_rate); cudaFree(c->rc_du); cudaFree(c->rc_surf); }
    if (c->rc_pending) cudaFree(c->rc_pending);
    if (c->h_mail) cudaFreeHost(c->h_mail);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int kmc_set_params(KmcCtx* c, const KmcParams* p) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    TRY(check_params(p));
    c->hp = *p;
    c->dp = to_dev(*p);
    return 0;
}

int kmc_get_params(KmcCtx* c, KmcParams* p) {
    if (!c || !p) return fail(KMC_E_ARG, "NULL argument");
    *p = c->hp;
    return 0;
}

int kmc_set_stream(KmcCtx* c, void* s) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    DevGuard guard(c);
    CK(cudaStreamSynchronize(c->stream));
    if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = (cudaStream_t)s;
    return 0;
}

int kmc_synchronize(KmcCtx* c) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    DevGuard guard(c);
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int64_t kmc_launch_count(KmcCtx* c) { return c ? c->launches : 0; }

int kmc_timing_enable(KmcCtx* c, int on) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    c->timing = on != 0;
    return 0;
}

int kmc_timing_reset(KmcCtx* c) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    DevGuard guard(c);
    CK(cudaStreamSynchronize(c->stream));
    for (auto& r : c->recs) { c->event_pool.push_back(r.a); c->event_pool.push_back(r.b); }      // recycled, not destroyed
    c->recs.clear();
    return 0;
}

int kmc_timing_read(KmcCtx* c, const char* family, double* total_ms, int64_t* launches) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    DevGuard guard(c);
    CK(cudaStreamSynchronize(c->stream));
    double tot = 0;
    int64_t cnt = 0;
    for (auto& r : c->recs) {
        if (family && family[0] && c->families[r.family].find(family) == std::string::npos) continue;
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, r.a, r.b));
        tot += ms;
        cnt++;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = cnt;
    return 0;
}

// ---------------------------------------------------------------- Periodic.py
int kmc_put_all_in_box(KmcCtx* c, double* xy, int64_t n, int mem) {
    if (!c || (n > 0 && !xy)) return fail(KMC_E_ARG, "NULL argument");
    if (n == 0) return 0;
    Call call(c, mem);
    double* d = xy;
    if (mem == KMC_HOST) {
        const double* din;
        TRY(call.in(xy, (size_t)2 * n, S_IN0, &din));
        d = (double*)din;
        call.outs.push_back({xy, d, sizeof(double) * 2 * (size_t)n});
    }
    LAUNCH(c, "misc_wrap", k_put_all_in_box, cdiv(n, 256), 256, c->dp, (double2*)d, (int)n);
    CK(cudaGetLastError());
    return call.finish();
}

int kmc_distances(KmcCtx* c, const double* a, int64_t na, const double* b, int64_t nb, double* out, int mem) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    if (na == 0 || nb == 0) return 0;
    Call call(c, mem);
    const double *da, *db;
    double* dout;
    TRY(call.in(a, (size_t)2 * na, S_IN0, &da));
    TRY(call.in(b, (size_t)2 * nb, S_IN1, &db));
    TRY(call.out(out, (size_t)na * nb, S_OUT0, &dout));
    LAUNCH(c, "misc_distances", k_distances, cdiv((size_t)na * nb, 256), 256, c->dp, (const double2*)da, (int)na, (const double2*)db, (int)nb, dout);
    CK(cudaGetLastError());
    return call.finish();
}

int kmc_displacement(KmcCtx* c, const double* Ri, const double* Rj, int64_t k, double* out, int mem) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    if (k == 0) return 0;
    Call call(c, mem);
    const double *da, *db;
    double* dout;
    TRY(call.in(Ri, (size_t)2 * k, S_IN0, &da));
    TRY(call.in(Rj, (size_t)2 * k, S_IN1, &db));
    TRY(call.out(out, (size_t)2 * k, S_OUT0, &dout));
    LAUNCH(c, "misc_displacement", k_displacement, cdiv(k, 256), 256, c->dp, (const double2*)da, (const double2*)db, (int)k, (double2*)dout);
    CK(cudaGetLastError());
    return call.finish();
}

// ---------------------------------------------------------------- Bins.py
int kmc_bin_index(KmcCtx* c, const double* xy, int64_t n, int32_t* out, int mem) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    if (n == 0) return 0;
    Call call(c, mem);
    const double* d;
    int32_t* dout;
    TRY(call.in(xy, (size_t)2 * n, S_IN0, &d));
    TRY(call.out(out, (size_t)2 * n, S_OUT0, &dout));
    LAUNCH(c, "bin_index", k_bin_index, cdiv(n, 256), 256, c->dp, (const double2*)d, (int)n, (int2*)dout);
    CK(cudaGetLastError());
    return call.finish();
}

int kmc_bins_create(KmcCtx* c, const double* xy, int64_t n, int mem, KmcBins** out) {
    if (!c || !out) return fail(KMC_E_ARG, "NULL argument");
    if (n < 0 || n > (1ll << 30)) return fail(KMC_E_ARG, "bad atom count");
    DevGuard guard(c);
    KmcBins* b = new KmcBins();
    b->n = n;
    b->ncell2 = c->dp.ncell + 2;
    size_t na = (size_t)(n > 0 ? n : 1);
    int* cursor;
    if (cudaMalloc(&b->xy, sizeof(double2) * na) != cudaSuccess || cudaMalloc(&b->cell_of, sizeof(int) * na) != cudaSuccess ||
        cudaMalloc(&b->start, sizeof(int) * b->ncell2) != cudaSuccess || cudaMalloc(&b->sxy, sizeof(double2) * na) != cudaSuccess ||
        cudaMalloc(&b->sidx, sizeof(int) * na) != cudaSuccess) {
        kmc_bins_destroy(c, b);
        return fail(KMC_E_CUDA, "cudaMalloc failed for a cell list of %lld atoms", (long long)n);
    }
    auto body = [&]() -> int {
        if (n > 0)
            CK(cudaMemcpyAsync(b->xy, xy, sizeof(double2) * (size_t)n, mem == KMC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
        TRY(slot(c, S_CURSOR, (size_t)b->ncell2, &cursor));
        TRY(kmc_build_cells(c, b->xy, (int)n, 1, b->cell_of, b->start, cursor, b->sidx, b->sxy));
        if (mem == KMC_HOST) CK(cudaStreamSynchronize(c->stream));
        return 0;
    };
    const int rc = body();
    if (rc != 0) { kmc_bins_destroy(c, b); return rc; }       // no half-built cell list leaks
    *out = b;
    return 0;
}

int kmc_bins_destroy(KmcCtx* c, KmcBins* b) {
    if (!b) return 0;
    DevGuard guard(c);
    if (c) cudaStreamSynchronize(c->stream);
    cudaFree(b->xy); cudaFree(b->cell_of); cudaFree(b->start); cudaFree(b->sxy); cudaFree(b->sidx);
    delete b;
    return 0;
}

int64_t kmc_bins_count(const KmcBins* b) { return b ? b->n : 0; }

int kmc_bins_export(KmcCtx* c, const KmcBins* b, int32_t* cell_of, int32_t* cell_start, double* sorted_xy, int32_t* sorted_idx, int mem) {
    if (!c || !b) return fail(KMC_E_ARG, "NULL argument");
    DevGuard guard(c);
    cudaMemcpyKind k = mem == KMC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (cell_of && b->n) CK(cudaMemcpyAsync(cell_of, b->cell_of, sizeof(int) * (size_t)b->n, k, c->stream));
    if (cell_start) CK(cudaMemcpyAsync(cell_start, b->start, sizeof(int) * (size_t)b->ncell2, k, c->stream));
    if (sorted_xy && b->n) CK(cudaMemcpyAsync(sorted_xy, b->sxy, sizeof(double2) * (size_t)b->n, k, c->stream));
    if (sorted_idx && b->n) CK(cudaMemcpyAsync(sorted_idx, b->sidx, sizeof(int) * (size_t)b->n, k, c->stream));
    if (mem == KMC_HOST) CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int kmc_nearby_atoms(KmcCtx* c, const KmcBins* b, const double* qxy, int64_t nq, double* out_xy, int64_t cap, int64_t* offsets, int mem) {
    if (!c || !b || !offsets) return fail(KMC_E_ARG, "NULL argument");
    if (nq > (1 << 24)) return fail(KMC_E_ARG, "too many query points");
    Call call(c, mem);
    const double* dq;
    long long* doff;
    double* dout;
    TRY(call.in(qxy, (size_t)2 * nq, S_IN0, &dq));
    TRY(call.out((long long*)offsets, (size_t)nq + 1, S_OUT0, &doff));
    TRY(call.out(out_xy, (size_t)2 * cap, S_OUT1, &dout));
    CellView cv = view_of(b);
    if (nq == 0) { CK(cudaMemsetAsync(doff, 0, sizeof(long long), c->stream)); return call.finish(); }
    LAUNCH(c, "list_count", k_nearby_count, cdiv(nq, 128), 128, c->dp, cv, (const double2*)dq, (int)nq, doff);
    LAUNCH(c, "list_scan", k_scan_i64, 1, 1024, doff, (int)nq);
    if (dout) LAUNCH(c, "list_fill", k_nearby_fill, cdiv(nq, 128), 128, c->dp, cv, (const double2*)dq, (int)nq, doff, (long long)cap, (double2*)dout);
    CK(cudaGetLastError());
    TRY(call.finish());
    if (mem == KMC_HOST && out_xy && offsets[nq] > cap) return fail(KMC_E_CAPACITY, "nearby_atoms needs %lld atoms, capacity %lld", (long long)offsets[nq], (long long)cap);
    return 0;
}

int kmc_neighbors(KmcCtx* c, const double* xy, int64_t n, const KmcBins* sb, int32_t* n_a, int32_t* n_s, double* a_xy, int64_t a_cap,
                  int64_t* a_offs, double* s_xy, int64_t s_cap, int64_t* s_offs, int mem) {
    if (!c || !sb) return fail(KMC_E_ARG, "NULL argument");
    Call call(c, mem);
    const double* d;
    int *dna, *dns;
    TRY(call.in(xy, (size_t)2 * n, S_IN0, &d));
    // counts are always computed (scratch when the caller does not want them)
    if (n_a) TRY(call.out(n_a, (size_t)n, S_OUT0, &dna)); else TRY(slot(c, S_OUT0, (size_t)n, &dna));
    if (n_s) TRY(call.out(n_s, (size_t)n, S_OUT1, &dns)); else TRY(slot(c, S_OUT1, (size_t)n, &dns));
    long long *dao = nullptr, *dso = nullptr;
    double *dax = nullptr, *dsx = nullptr;
    if (a_offs) TRY(call.out((long long*)a_offs, (size_t)n + 1, S_OUT2, &dao));
    if (s_offs) TRY(call.out((long long*)s_offs, (size_t)n + 1, S_OUT3, &dso));
    if (a_xy) TRY(call.out(a_xy, (size_t)2 * a_cap, S_OUT4, &dax));
    if (s_xy) TRY(call.out(s_xy, (size_t)2 * s_cap, S_OUT5, &dsx));
    if (n == 0) {
        if (dao) CK(cudaMemsetAsync(dao, 0, sizeof(long long), c->stream));
        if (dso) CK(cudaMemsetAsync(dso, 0, sizeof(long long), c->stream));
        return call.finish();
    }
    TRY(kmc_rates_dev(c, (const double2*)d, (int)n, sb, nullptr, nullptr, nullptr, nullptr, nullptr, dna, dns));
    CellView ad;   // cell list built by rates_dev lives in the scratch slots
    ad.start = (int*)c->slots[S_START].p; ad.sxy = (double2*)c->slots[S_SXY].p; ad.sidx = (int*)c->slots[S_SIDX].p; ad.n = (int)n;
    if (dao) {
        LAUNCH(c, "list_count", k_i32_to_offsets, cdiv(n, 256), 256, dna, (int)n, dao);
        LAUNCH(c, "list_scan", k_scan_i64, 1, 1024, dao, (int)n);
        if (dax) LAUNCH(c, "list_fill", k_neighbor_fill, cdiv(n, 128), 128, c->dp, (const double2*)d, (int)n, ad, c->dp.bond_a, dao, (long long)a_cap, (double2*)dax);
    }
    if (dso) {
        LAUNCH(c, "list_count", k_i32_to_offsets, cdiv(n, 256), 256, dns, (int)n, dso);
        LAUNCH(c, "list_scan", k_scan_i64, 1, 1024, dso, (int)n);
        if (dsx) LAUNCH(c, "list_fill", k_neighbor_fill, cdiv(n, 128), 128, c->dp, (const double2*)d, (int)n, view_of(sb), c->dp.bond_s, dso, (long long)s_cap, (double2*)dsx);
    }
    CK(cudaGetLastError());
    TRY(call.finish());
    if (mem == KMC_HOST) {
        if (a_xy && a_offs && a_offs[n] > a_cap) return fail(KMC_E_CAPACITY, "adatom bond list needs %lld atoms", (long long)a_offs[n]);
        if (s_xy && s_offs && s_offs[n] > s_cap) return fail(KMC_E_CAPACITY, "substrate bond list needs %lld atoms", (long long)s_offs[n]);
    }
    return 0;
}

// ---------------------------------------------------------------- Energy.py
int kmc_forces(KmcCtx* c, const double* xy, int64_t n, const KmcBins* sb, int which, double* f_out, int mem) {
    if (!c || !sb || !f_out) return fail(KMC_E_ARG, "NULL argument");
    if (!(which & 3)) return fail(KMC_E_ARG, "which must include KMC_F_AA and/or KMC_F_AS");
    if (n == 0) return 0;
    Call call(c, mem);
    const double* d;
    double* dout;
    TRY(call.in(xy, (size_t)2 * n, S_IN0, &d));
    TRY(call.out(f_out, (size_t)2 * n, S_OUT0, &dout));
    CellView ad;
    TRY(adatom_view(c, (const double2*)d, (int)n, 1, false, &ad));
    double2 *faa = nullptr, *fas = nullptr, *fsum = nullptr;
    if (which == KMC_F_AA) faa = (double2*)dout;
    else if (which == KMC_F_AS) fas = (double2*)dout;
    else fsum = (double2*)dout;
    TRY(forces_dev(c, ad, nullptr, sb, (int)n, faa, fas, fsum));
    return call.finish();
}

int kmc_relax_energy(KmcCtx* c, const double* xy, int64_t n, int64_t B, const KmcBins* sb, double* e_out, int mem) {
    if (!c || !sb || !e_out) return fail(KMC_E_ARG, "NULL argument");
    if (B < 1 || B > 65535) return fail(KMC_E_ARG, "B must be in 1..65535");
    Call call(c, mem);
    const double* d;
    double* de;
    TRY(call.in(xy, (size_t)2 * n * B, S_IN0, &d));
    TRY(call.out(e_out, (size_t)B, S_OUT0, &de));
    CellView ad;
    TRY(adatom_view(c, (const double2*)d, (int)n, (int)B, false, &ad));
    TRY(energy_dev(c, ad, sb, (int)n, (int)B, de));
    return call.finish();
}

int kmc_line_energies(KmcCtx* c, const double* xy, const double* dir, int64_t n, int32_t scale, int32_t K, const KmcBins* sb, double* e_out, int mem) {
    if (!c || !sb || !e_out) return fail(KMC_E_ARG, "NULL argument");
    if (K < 1 || K > 1024 || scale < 1) return fail(KMC_E_ARG, "bad K or scale");
    Call call(c, mem);
    const double *dx, *ds;
    double *de, *cand;
    TRY(call.in(xy, (size_t)2 * n, S_IN0, &dx));
    TRY(call.in(dir, (size_t)2 * n, S_IN1, &ds));
    TRY(call.out(e_out, (size_t)K, S_OUT0, &de));
    if (K == KC && n > 0) {      // the line-search engine: one cell list, 20 candidates per pair in registers
        CellView ad;
        TRY(kmc_scratch_cells(c, (const double2*)dx, (int)n, 1, &ad));
        TRY(line_energies_dev(c, (const double2*)dx, (const double2*)ds, (int)n, (double)scale, ad, sb, de, nullptr));
        return call.finish();
    }
    TRY(slot(c, S_CAND, (size_t)2 * n * K, &cand));
    if (n > 0) {
        dim3 g(cdiv(2 * n, 256), K);
        LAUNCH(c, "misc_candidates", k_make_candidates, g, 256, dx, ds, (int)(2 * n), (int)K, (double)scale, cand);
    }
    CellView ad;
    TRY(adatom_view(c, (const double2*)cand, (int)n, K, false, &ad));
    TRY(energy_dev(c, ad, sb, (int)n, K, de));
    return call.finish();
}

int kmc_total_energy(KmcCtx* c, const double* xy, int64_t n, const KmcBins* sb, double* e_out, int mem) {
    if (!c || !sb || !e_out) return fail(KMC_E_ARG, "NULL argument");
    Call call(c, mem);
    const double* d;
    double* de;
    TRY(call.in(xy, (size_t)2 * n, S_IN0, &d));
    TRY(call.out(e_out, 1, S_OUT0, &de));
    LAUNCH(c, "misc_total_energy", k_total_energy, 1, 1024, c->dp, (const double2*)d, view_of(sb), (int)n, de);
    CK(cudaGetLastError());
    return call.finish();
}

int kmc_probe(KmcCtx* c, const double* probes, int64_t m, const double* xy, int64_t n, const KmcBins* sb, double* e_aa, double* e_as,
              double* min_d_a, double* min_d_s, int mem) {
    if (!c || !sb) return fail(KMC_E_ARG, "NULL argument");
    if (m == 0) return 0;
    Call call(c, mem);
    const double *dp, *dx;
    double *d0, *d1, *d2, *d3;
    TRY(call.in(probes, (size_t)2 * m, S_IN0, &dp));
    TRY(call.in(xy, (size_t)2 * n, S_IN1, &dx));
    TRY(call.out(e_aa, (size_t)m, S_OUT0, &d0));
    TRY(call.out(e_as, (size_t)m, S_OUT1, &d1));
    TRY(call.out(min_d_a, (size_t)m, S_OUT2, &d2));
    TRY(call.out(min_d_s, (size_t)m, S_OUT3, &d3));
    CellView ad = identity_view(nullptr, 0);
    if (n > 0) TRY(kmc_scratch_cells(c, (const double2*)dx, (int)n, 1, &ad));
    LAUNCH(c, "probe", k_probe, cdiv((size_t)m * 32, 128), 128, c->dp, (const double2*)dp, (int)m, ad, view_of(sb), d0, d1, d2, d3);
    CK(cudaGetLastError());
    return call.finish();
}

int kmc_rates(KmcCtx* c, const double* xy, int64_t n, const KmcBins* sb, double* rate, uint8_t* surf, double* du, double* ua, double* ul,
              int32_t* na, int32_t* ns, int mem) {
    if (!c || !sb) return fail(KMC_E_ARG, "NULL argument");
    if (n == 0) return 0;
    Call call(c, mem);
    const double* d;
    double *d_rate, *d_du, *d_ua, *d_ul;
    uint8_t* d_surf;
    int *d_na, *d_ns;
    TRY(call.in(xy, (size_t)2 * n, S_IN0, &d));
    TRY(call.out(rate, (size_t)n, S_OUT0, &d_rate));
    TRY(call.out(surf, (size_t)n, S_OUT1, &d_surf));
    TRY(call.out(du, (size_t)n, S_OUT2, &d_du));
    TRY(call.out(ua, (size_t)n, S_OUT3, &d_ua));
    TRY(call.out(ul, (size_t)n, S_OUT4, &d_ul));
    TRY(call.out(na, (size_t)n, S_OUT5, &d_na));
    TRY(call.out(ns, (size_t)n, S_OUT6, &d_ns));
    TRY(kmc_rates_dev(c, (const double2*)d, (int)n, sb, d_rate, d_surf, d_du, d_ua, d_ul, d_na, d_ns));
    return call.finish();
}

int kmc_select(KmcCtx* c, const double* rate, const uint8_t* surf, int64_t n, double u, int64_t* result, double* rtot, double* pj, int mem) {
    if (!c || !result || !rtot) return fail(KMC_E_ARG, "NULL argument");
    if (n > 0 && (!rate || !surf)) return fail(KMC_E_ARG, "NULL rate table");
    Call call(c, mem);
    const double* dr;
    const uint8_t* dsf;
    long long* dres;
    double *drt, *dpj;
    TRY(call.in(rate, (size_t)n, S_IN0, &dr));
    TRY(call.in(surf, (size_t)n, S_IN1, &dsf));
    TRY(call.out((long long*)result, 2, S_OUT0, &dres));
    TRY(call.out(rtot, 1, S_OUT1, &drt));
    TRY(call.out(pj, (size_t)n + 1, S_OUT2, &dpj));
    TRY(kmc_select_dev(c, dr, dsf, (int)n, u, dres, drt, dpj));
    return call.finish();
}

// ---------------------------------------------------------------- spatial domain decomposition (BASELINE configs[4])
// One sweep of a rank's LOCAL configuration: the first n_own atoms of xy are the rank's own, the rest are halo copies
// received from the neighbouring ranks (or far-away dummies).  ONE cell-list build serves the force sweep and the rate
// table; only own atoms are evaluated.  Everything stays on the device (mem = KMC_DEVICE: no copy, no synchronisation).
int kmc_domain_sweep(KmcCtx* c, const double* xy, int64_t n, int64_t n_own, const KmcBins* sb, double* f_out, double* e_aa, double* e_as,
                     double* rate, uint8_t* surf, int mem) {
    if (!c || !sb) return fail(KMC_E_ARG, "NULL argument");
    if (n_own < 0 || n_own > n) return fail(KMC_E_ARG, "n_own must be in 0..n");
    if (c->dp.aa_mode != 1) return fail(KMC_E_ARG, "domain decomposition needs aa_mode == 1 (bins3x3 adatom sums)");
    if (n == 0) return 0;
    Call call(c, mem);
    const double* d;
    double *df, *da, *ds, *dr;
    uint8_t* dsf;
    TRY(call.in(xy, (size_t)2 * n, S_IN0, &d));
    TRY(call.out(f_out, (size_t)2 * n_own, S_OUT0, &df));
    TRY(call.out(e_aa, (size_t)n_own, S_OUT1, &da));
    TRY(call.out(e_as, (size_t)n_own, S_OUT2, &ds));
    TRY(call.out(rate, (size_t)n_own, S_OUT3, &dr));
    TRY(call.out(surf, (size_t)n_own, S_OUT4, &dsf));
    CellView ad;
    TRY(kmc_scratch_cells(c, (const double2*)d, (int)n, 1, &ad));
    if (df) TRY(forces_dev(c, ad, nullptr, sb, (int)n, nullptr, nullptr, (double2*)df, (int)n_own));
    if (da || ds || dr || dsf) {
        const int G = pick_group((size_t)n);
        const unsigned grid = cdiv((size_t)n * G, 256);
        CellView sub = view_of(sb);
        double* nd = nullptr; int* ni = nullptr; const uint8_t* nu = nullptr; double2* n2 = nullptr; unsigned long long* nc = nullptr;
#define DD_RATES(GG) LAUNCH(c, "sweep_rates", k_sweep_rates<GG>, grid, 256, c->dp, ad, sub, (int)n, dr, dsf, nd, nd, nd, ni, ni, nu, nu, n2, nc, (int)n_own, da, ds)
        switch (G) {
            case 32: DD_RATES(32); break;
            case 16: DD_RATES(16); break;
            case 8: DD_RATES(8); break;
            default: DD_RATES(4); break;
        }
#undef DD_RATES
        CK(cudaGetLastError());
    }
    return call.finish();
}

// This rank's share of RelaxEnergy (Energy.py:133-148) for the K line-search candidates x + a*s/20/scale (2DGrowth.py:121,152)
// of its local configuration: own atoms take half of each adatom pair term and all of their substrate terms; the sum
// over the ranks (one all-reduce of K doubles) is the configuration energy of every candidate.
int kmc_domain_line_energies(KmcCtx* c, const double* xy, const double* dir, int64_t n, int64_t n_own, int32_t scale, int32_t K,
                             const KmcBins* sb, double* e_out, int mem) {
    if (!c || !sb || !e_out) return fail(KMC_E_ARG, "NULL argument");
    if (K < 1 || K > 1024 || scale < 1) return fail(KMC_E_ARG, "bad K or scale");
    if (n_own < 0 || n_own > n) return fail(KMC_E_ARG, "n_own must be in 0..n");
    if (c->dp.aa_mode != 1) return fail(KMC_E_ARG, "domain decomposition needs aa_mode == 1 (bins3x3 adatom sums)");
    Call call(c, mem);
    const double *dx, *ds;
    double *de, *cand;
    TRY(call.in(xy, (size_t)2 * n, S_IN0, &dx));
    TRY(call.in(dir, (size_t)2 * n, S_IN1, &ds));
    TRY(call.out(e_out, (size_t)K, S_OUT0, &de));
    if (K == KC && n > 0) {      // the line-search engine: ONE cell list of the base positions, 20 candidates per pair in registers
        CellView ad;
        TRY(kmc_scratch_cells(c, (const double2*)dx, (int)n, 1, &ad));
        TRY(line_energies_dev(c, (const double2*)dx, (const double2*)ds, (int)n, (double)scale, ad, sb, de, nullptr, (int)n_own));
        return call.finish();
    }
    TRY(slot(c, S_CAND, (size_t)2 * n * K, &cand));
    if (n > 0) {
        dim3 g(cdiv(2 * n, 256), K);
        LAUNCH(c, "misc_candidates", k_make_candidates, g, 256, dx, ds, (int)(2 * n), (int)K, (double)scale, cand);
    }
    CellView ad;
    TRY(kmc_scratch_cells(c, (const double2*)cand, (int)n, K, &ad));
    TRY(energy_dev(c, ad, sb, (int)n, K, de, (int)n_own));
    return call.finish();
}

// ---------------------------------------------------------------- trajectory state + events
int kmc_state_set(KmcCtx* c, const double* xy, int64_t n, int mem) {
    if (!c || (n > 0 && !xy)) return fail(KMC_E_ARG, "NULL argument");
    if (c->ev_pending) return fail(KMC_E_ARG, "an enqueued event has not been finished (kmc_event_finish)");
    DevGuard guard(c);
    TRY(kmc_state_reserve(c, n + 1));
    if (n > 0)
        CK(cudaMemcpyAsync(c->st_xy, xy, sizeof(double) * 2 * (size_t)n, mem == KMC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    c->st_n = n;
    c->rc_valid = false;            // an uploaded state has no history: the next rate table is a full sweep
    c->pl_valid = false;
    if (mem == KMC_HOST) CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int kmc_state_move(KmcCtx* c, const double* xy, int64_t n, int mem) {
    if (!c || (n > 0 && !xy)) return fail(KMC_E_ARG, "NULL argument");
    if (c->ev_pending) return fail(KMC_E_ARG, "an enqueued event has not been finished (kmc_event_finish)");
    DevGuard guard(c);
    if (n != c->st_n) return fail(KMC_E_ARG, "kmc_state_move: %lld positions for a state of %lld atoms (same atoms, new positions)", (long long)n, (long long)c->st_n);
    if (n > 0)
        CK(cudaMemcpyAsync(c->st_xy, xy, sizeof(double) * 2 * (size_t)n, mem == KMC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    if (mem == KMC_HOST) CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int kmc_state_get(KmcCtx* c, double* xy, int64_t cap, int64_t* n_out, int mem) {
    if (!c) return fail(KMC_E_ARG, "ctx is NULL");
    if (c->ev_pending) return fail(KMC_E_ARG, "an enqueued event has not been finished (kmc_event_finish)");
    DevGuard guard(c);
    if (n_out) *n_out = c->st_n;
    if (!xy) return 0;
    if (cap < c->st_n) return fail(KMC_E_CAPACITY, "state holds %lld atoms, capacity %lld", (long long)c->st_n, (long long)cap);
    if (c->st_n > 0)
        CK(cudaMemcpyAsync(xy, c->st_xy, sizeof(double) * 2 * (size_t)c->st_n, mem == KMC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    if (mem == KMC_HOST) CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int64_t kmc_state_count(KmcCtx* c) { return c ? c->st_n : 0; }

// 2DGrowth.py:97-166, host-driven: every line search is a handful of launches plus one 48-byte
// readback (maxF, miny, top, bot) for the reference's control flow.
static int relax_host_driven(KmcCtx* c, const KmcBins* sb, int32_t scale0, double threshold0, int64_t max_ls, KmcRelaxStats* stats) {
    if (scale0 < 1 || !(threshold0 > 0)) return fail(KMC_E_ARG, "bad scale0 / threshold0");
    KmcRelaxStats st{};
    const int n = (int)c->st_n;
    if (n == 0) { if (stats) *stats = st; return 0; }
    const int n2 = 2 * n, K = 20;
    double *x0, *xn, *xnp, *sn, *F, *cand, *e;
    RelaxCtrl* ctrl;
    TRY(slot(c, S_X0, (size_t)n2, &x0));
    TRY(slot(c, S_XN, (size_t)n2, &xn));
    TRY(slot(c, S_XNP, (size_t)n2, &xnp));
    TRY(slot(c, S_SN, (size_t)n2, &sn));
    TRY(slot(c, S_F, (size_t)n2, &F));
    TRY(slot(c, S_CAND, (size_t)n2 * K, &cand));
    TRY(slot(c, S_E, (size_t)K, &e));
    TRY(slot(c, S_CTRL, 1, &ctrl));
    CK(cudaMemcpyAsync(x0, c->st_xy, sizeof(double) * n2, cudaMemcpyDeviceToDevice, c->stream));
    int scale = scale0;
    double threshold = threshold0;
    int rc = 0;
    const RelaxCtrl* hm = (const RelaxCtrl*)c->h_mail;
    const bool cells = c->dp.aa_mode != 0;
    double maxF = 0;

    // one line search from (xbase, sn): candidates -> (cells) -> energies -> argmin -> forces at the winner -> reductions
    auto line_search = [&](const double* xbase) -> int {
        dim3 g(cdiv(n2, 256), K);
        LAUNCH(c, "misc_candidates", k_make_candidates, g, 256, xbase, sn, n2, K, (double)scale, cand);
        CellView ad;
        TRY(adatom_view(c, (const double2*)cand, n, K, false, &ad));
        TRY(energy_dev(c, ad, sb, n, K, e));
        LAUNCH(c, "relax_argmin", k_argmin, 1, 32, e, K, ctrl);
        if (!cells) ad.sxy = (const double2*)cand;   // identity view: which_cfg offsets into the candidate batch
        TRY(forces_dev(c, ad, &ctrl->amin, sb, n, nullptr, nullptr, (double2*)F));
        LAUNCH(c, "relax_reduce", k_relax_reduce, 1, 1024, cand, xbase, (const double2*)F, n, ctrl, xnp);
        CK(cudaGetLastError());
        TRY(kmc_mail(c, ctrl, sizeof(RelaxCtrl)));
        st.line_searches++;
        return 0;
    };

    for (;;) {
        if (scale > 1024) { scale = 16; threshold *= 10; }                       // :111-113
        CK(cudaMemcpyAsync(xn, x0, sizeof(double) * n2, cudaMemcpyDeviceToDevice, c->stream));
        {   // dx0 = forces(x0) ; sn = dx0                                        // :120,125
            CellView ad;
            TRY(adatom_view(c, (const double2*)xn, n, 1, false, &ad));
            TRY(forces_dev(c, ad, nullptr, sb, n, nullptr, nullptr, (double2*)sn));
        }
        TRY(line_search(xn));                                                     // :121-129
        maxF = hm->maxF;
        double lastmaxF = maxF;
        bool restart = false;
        while (maxF > threshold) {                                                // :131
            if (max_ls > 0 && st.line_searches >= max_ls) { rc = KMC_E_ITERCAP; break; }
            if (hm->miny > c->hp.Dmax) { restart = true; break; }                 // :137-140
            const double beta = hm->top / hm->bot;                                // :143-145
            LAUNCH(c, "relax_update", k_cg_update, cdiv(n2, 256), 256, F, beta, sn, xn, xnp, n2);   // :147-150
            TRY(line_search(xn));                                                 // :152-160
            maxF = hm->maxF;
            if (std::fabs(lastmaxF - maxF) < 1e-8) { restart = true; break; }     // :162-164
            lastmaxF = maxF;
        }
        if (rc != 0 || !restart) break;
        st.restarts++;
        scale *= 2;
    }
    st.maxF = maxF;
    st.final_scale = scale;
    st.final_threshold = threshold;
    st.status = rc;
    CK(cudaMemcpyAsync(c->st_xy, xnp, sizeof(double) * n2, cudaMemcpyDeviceToDevice, c->stream));
    LAUNCH(c, "misc_wrap", k_put_all_in_box, cdiv(n, 256), 256, c->dp, (double2*)c->st_xy, n);       // :166
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    if (stats) *stats = st;
    if (rc != 0) return fail(rc, "relaxation stopped at the line-search cap (%lld)", (long long)max_ls);
    return 0;
}


int kmc_relax(KmcCtx* c, const KmcBins* sb, int32_t scale0, double threshold0, int64_t max_ls, KmcRelaxStats* stats) {
    if (!c || !sb) return fail(KMC_E_ARG, "NULL argument");
    if (scale0 < 1 || !(threshold0 > 0)) return fail(KMC_E_ARG, "bad scale0 / threshold0");
    if (c->ev_pending) return fail(KMC_E_ARG, "an enqueued event has not been finished (kmc_event_finish)");
    DevGuard guard(c);
    TRY(kmc_state_reserve(c, c->st_n + 1));
    if (c->relax_mode == 0) return relax_host_driven(c, sb, scale0, threshold0, max_ls, stats);
    if (c->relax_mode == 2) return kmc_relax2_persistent(c, sb, scale0, threshold0, max_ls, stats);
    return kmc_relax_persistent1(c, sb, scale0, threshold0, max_ls, stats);
}

int kmc_set_relax_mode(KmcCtx* c, int mode) {
    if (!c || mode < 0 || mode > 2) return fail(KMC_E_ARG, "relax mode must be 0 (host-driven), 1 (persistent kernel v1) or 2 (persistent kernel on sorted state)");
    c->relax_mode = mode;
    return 0;
}

int kmc_set_select_mode(KmcCtx* c, int mode) {
    if (!c || mode < 0 || mode > 2) return fail(KMC_E_ARG, "select mode must be 0 (auto), 1 (sequential order) or 2 (block scan)");
    c->select_mode = mode;
    return 0;
}

int kmc_set_rate_mode(KmcCtx* c, int mode, double threshold) {
    if (!c || mode < 0 || mode > 1 || !(threshold >= 0.0)) return fail(KMC_E_ARG, "rate mode must be 0 (full sweep) or 1 (incremental), threshold >= 0");
    c->rate_mode = mode;
    c->rate_threshold = threshold;
    c->rc_valid = false;
    return 0;
}

int kmc_rates_update(KmcCtx* c, const KmcBins* sb, double threshold, double* rate, uint8_t* surf, double* du, int64_t* recomputed, int mem) {
    if (!c || !sb) return fail(KMC_E_ARG, "NULL argument");
    if (!(threshold >= 0.0)) return fail(KMC_E_ARG, "threshold must be >= 0");
    DevGuard guard(c);
    int64_t redone = 0;
    TRY(kmc_rates_update_impl(c, sb, threshold, &redone));
    if (recomputed) *recomputed = redone;
    const size_t n = (size_t)c->st_n;
    const cudaMemcpyKind kind = mem == KMC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (n) {
        if (rate) CK(cudaMemcpyAsync(rate, c->rc_rate, sizeof(double) * n, kind, c->stream));
        if (surf) CK(cudaMemcpyAsync(surf, c->rc_surf, n, kind, c->stream));
        if (du) CK(cudaMemcpyAsync(du, c->rc_du, sizeof(double) * n, kind, c->stream));
    }
    if (mem == KMC_HOST) CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// host-stepped placement (event mode 0): the first-generation path, kept as a cross-check of the device-resident one
static int deposit_place_host(KmcCtx* c, const KmcBins* sb, double u) {
    const int n = (int)c->st_n;
    TRY(kmc_state_reserve(c, n + 1));
    double ref_y = 0.0;                                                           // :73-75
    double* tmp;
    TRY(slot(c, S_TMP0, 8, &tmp));
    if (n > 0) {
        LAUNCH(c, "misc_min_y", k_min_y, 1, 1024, (const double2*)c->st_xy, n, c->hp.deposit_mode ? 1 : 0, tmp);
        TRY(kmc_mail(c, tmp, sizeof(double)));
        ref_y = *(const double*)c->h_mail;
    }
    // :76  new_x = PutInBox(u*L) (host arithmetic identical to the reference's, Periodic.py:8-20)
    double new_x = u * c->hp.L;
    {
        double r = std::fmod(new_x, c->hp.L);
        if (r != 0.0) { if (r < 0.0) r += c->hp.L; } else r = 0.0;
        if (std::fabs(r) > c->hp.L / 2) r -= c->hp.L;
        new_x = r;
    }
    double new_y = 2 * c->hp.r_as + ref_y;                                        // :77
    CellView ad = identity_view(nullptr, 0);
    if (n > 0) TRY(kmc_scratch_cells(c, (const double2*)c->st_xy, n, 1, &ad));        // :78
    // :80-92 -- the probe descends by r_a/2 until it is close to something; 64 heights per launch
    const int M = 64;
    double *probes, *dmin_a, *dmin_s;
    TRY(slot(c, S_TMP1, (size_t)2 * M, &probes));
    TRY(slot(c, S_TMP2, (size_t)M, &dmin_a));
    TRY(slot(c, S_TMP3, (size_t)M, &dmin_s));
    std::vector<double> hp(2 * M), hd(2 * M);
    for (int round = 0; round < 4096; ++round) {
        double y = new_y;
        for (int k = 0; k < M; ++k) { hp[2 * k] = new_x; hp[2 * k + 1] = y; y -= c->hp.r_a / 2; }
        CK(cudaMemcpyAsync(probes, hp.data(), sizeof(double) * 2 * M, cudaMemcpyHostToDevice, c->stream));
        LAUNCH(c, "probe", k_probe, cdiv((size_t)M * 32, 128), 128, c->dp, (const double2*)probes, M, ad, view_of(sb), nullptr, nullptr, dmin_a, dmin_s);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(hd.data(), dmin_a, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(hd.data() + M, dmin_s, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < M; ++k) {
            if (hd[k] < 2 * c->hp.r_a || hd[M + k] < 2 * c->hp.r_as) {
                double pos[2] = {new_x, hp[2 * k + 1]};
                CK(cudaMemcpyAsync(c->st_xy + 2 * (size_t)n, pos, sizeof(pos), cudaMemcpyHostToDevice, c->stream));   // :93
                CK(cudaStreamSynchronize(c->stream));
                c->st_n = n + 1;
                TRY(kmc_rc_on_append(c));
                return 0;
            }
        }
        new_y = y;
    }
    return fail(KMC_E_DEPOSIT, "deposition probe never came within 2 r of an atom");
}

// device helper for HopAtom: ordered list of surface atoms within 3 r_a of the jumper, pick int(u*ncand)
__global__ void __launch_bounds__(1024) k_hop_target(DevParams P, const double2* __restrict__ xy, const uint8_t* __restrict__ surf, int n, int k, double u,
                                                     double* __restrict__ out /* [0]=ncand [1]=tx [2]=ty */) {
    // one block: every thread owns a contiguous slice (keeps the candidates in adatom order, 2DGrowth.py:244-246)
    __shared__ int s_scan[1024];
    const int T = blockDim.x, t = threadIdx.x;
    const double2 j = xy[k];
    const double cut = 3 * P.r_a;
    const int per = (n + T - 1) / T;
    const int lo = t * per, hi = min(lo + per, n);
    int cnt = 0;
    for (int i = lo; i < hi; ++i)
        if (surf[i] && sqrt(dist2_exact(P, j.x, j.y, xy[i].x, xy[i].y)) < cut) ++cnt;
    s_scan[t] = cnt;
    __syncthreads();
    for (int off = 1; off < T; off <<= 1) {
        int v = (t >= off) ? s_scan[t - off] : 0;
        __syncthreads();
        s_scan[t] += v;
        __syncthreads();
    }
    const int ncand = s_scan[T - 1];
    if (t == 0) out[0] = ncand;
    if (ncand == 0) return;
    int pick = (int)(u * (double)ncand);                                  // randint(0, ncand-1) -> int(u*ncand)
    if (pick >= ncand) pick = ncand - 1;
    const int before = s_scan[t] - cnt;
    if (pick >= before && pick < before + cnt) {
        int seen = before;
        for (int i = lo; i < hi; ++i)
            if (surf[i] && sqrt(dist2_exact(P, j.x, j.y, xy[i].x, xy[i].y)) < cut) {
                if (seen == pick) { out[1] = xy[i].x; out[2] = xy[i].y; }
                ++seen;
            }
    }
}

__global__ void k_remove_atom(const double2* __restrict__ in, int n, int k, double2* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) out[i] = in[i < k ? i : i + 1];
}

static int hop_place_impl(KmcCtx* c, const KmcBins* sb, int64_t k, double u, const uint8_t* surface_flags) {
    const int n = (int)c->st_n;
    if (k < 0 || k >= n) return fail(KMC_E_ARG, "hop index %lld out of range (n=%d)", (long long)k, n);
    uint8_t* own;
    double* tmp;
    TRY(slot(c, S_OUT1, (size_t)n, &own));
    TRY(slot(c, S_TMP0, 8, &tmp));
    // :239 SurfaceAtoms of the current state; kmc_event has just computed the same flags for the rate table
    const uint8_t* surf = surface_flags;
    if (!surf) {
        TRY(kmc_rates_dev(c, (const double2*)c->st_xy, n, sb, nullptr, own, nullptr, nullptr, nullptr, nullptr, nullptr));
        surf = own;
    }
    LAUNCH(c, "hop_target", k_hop_target, 1, 1024, c->dp, (const double2*)c->st_xy, surf, n, (int)k, u, tmp);          // :240-248
    CK(cudaGetLastError());
    TRY(kmc_mail(c, tmp, 3 * sizeof(double)));
    const double* hm = (const double*)c->h_mail;
    if (hm[0] < 1) return fail(KMC_E_NOCANDIDATE, "HopAtom: no surface atom within 3 r_a of adatom %lld", (long long)k);
    const double tx = hm[1], ty = hm[2];
    double* rest;
    TRY(slot(c, S_XN, (size_t)2 * n, &rest));
    LAUNCH(c, "misc_remove", k_remove_atom, cdiv(n, 256), 256, (const double2*)c->st_xy, n, (int)k, (double2*)rest);   // :241
    TRY(kmc_rc_on_remove(c, (int)k));
    double hp[12];
    for (int i = 0; i < 6; ++i) {                                                 // :249-251
        const double t = i * M_PI / 3;
        hp[2 * i] = tx + std::cos(t) * c->hp.r_a;
        hp[2 * i + 1] = ty + std::sin(t) * c->hp.r_a;
    }
    double *probes, *eaa, *eas;
    TRY(slot(c, S_TMP1, 128, &probes));
    TRY(slot(c, S_TMP2, 64, &eaa));
    TRY(slot(c, S_TMP3, 64, &eas));
    CK(cudaMemcpyAsync(probes, hp, sizeof(hp), cudaMemcpyHostToDevice, c->stream));
    CellView ad = identity_view(nullptr, 0);
    if (n - 1 > 0) TRY(kmc_scratch_cells(c, (const double2*)rest, n - 1, 1, &ad));
    LAUNCH(c, "probe", k_probe, cdiv(6 * 32, 128), 128, c->dp, (const double2*)probes, 6, ad, view_of(sb), eaa, eas, nullptr, nullptr);   // :252
    CK(cudaGetLastError());
    double he[12];
    CK(cudaMemcpyAsync(he, eaa, sizeof(double) * 6, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(he + 6, eas, sizeof(double) * 6, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int best = 0;
    for (int i = 1; i < 6; ++i)
        if (he[i] + he[6 + i] < he[best] + he[6 + best]) best = i;                // :253
    double pos[2] = {hp[2 * best], hp[2 * best + 1]};
    {   // :254 PutAllInBox
        double r = std::fmod(pos[0], c->hp.L);
        if (r != 0.0) { if (r < 0.0) r += c->hp.L; } else r = 0.0;
        if (std::fabs(r) > c->hp.L / 2) r -= c->hp.L;
        pos[0] = r;
    }
    CK(cudaMemcpyAsync(c->st_xy, rest, sizeof(double) * 2 * (size_t)(n - 1), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->st_xy + 2 * (size_t)(n - 1), pos, sizeof(pos), cudaMemcpyHostToDevice, c->stream));   // :255
    CK(cudaStreamSynchronize(c->stream));
    TRY(kmc_rc_on_append(c));
    return 0;
}

int kmc_deposit_place(KmcCtx* c, const KmcBins* sb, double u) {
    if (!c || !sb) return fail(KMC_E_ARG, "NULL argument");
    if (c->ev_pending) return fail(KMC_E_ARG, "an enqueued event has not been finished (kmc_event_finish)");
    DevGuard guard(c);
    if (c->event_mode == 0 || c->rate_mode == 1) return deposit_place_host(c, sb, u);
    return kmc_deposit_place_dev(c, sb, u);
}

int kmc_hop_place(KmcCtx* c, const KmcBins* sb, int64_t k, double u) {
    if (!c || !sb) return fail(KMC_E_ARG, "NULL argument");
    if (c->ev_pending) return fail(KMC_E_ARG, "an enqueued event has not been finished (kmc_event_finish)");
    DevGuard guard(c);
    if (c->event_mode == 0 || c->rate_mode == 1) return hop_place_impl(c, sb, k, u, nullptr);
    return kmc_hop_place_dev(c, sb, k, u);
}

int kmc_event(KmcCtx* c, const KmcBins* sb, double u0, double u1, int64_t max_ls, int32_t* kind_out, int64_t* hop_k_out, double* rtot_out,
              KmcRelaxStats* stats) {
    if (!c || !sb) return fail(KMC_E_ARG, "NULL argument");
    if (c->ev_pending) return fail(KMC_E_ARG, "an enqueued event has not been finished (kmc_event_finish)");
    if (c->event_mode == 1 && c->rate_mode == 0) {      // default: the whole event is queued, ONE host synchronisation
        TRY(kmc_event_enqueue(c, sb, u0, u1, max_ls));
        return kmc_event_finish(c, kind_out, hop_k_out, rtot_out, stats);
    }
    DevGuard guard(c);
    const int n = (int)c->st_n;
    double *rate, *rt;
    uint8_t* surf;
    long long* res;
    TRY(slot(c, S_OUT0, (size_t)n, &rate));
    TRY(slot(c, S_OUT1, (size_t)n, &surf));
    TRY(slot(c, S_OUT2, 2, &res));
    TRY(slot(c, S_OUT3, 1, &rt));
    if (c->rate_mode == 1) {        // incremental table: only the neighbourhoods that changed since the last event are redone
        int64_t redone = 0;
        TRY(kmc_rates_update_impl(c, sb, c->rate_threshold, &redone));
        rate = c->rc_rate; surf = c->rc_surf;
    } else {
        TRY(kmc_rates_dev(c, (const double2*)c->st_xy, n, sb, rate, surf, nullptr, nullptr, nullptr, nullptr, nullptr));   // :318
    }
    TRY(kmc_select_dev(c, rate, surf, n, u0, res, rt, nullptr));                                                   // :319-324
    long long hres[2];
    double hrt;
    CK(cudaMemcpyAsync(hres, res, sizeof(hres), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&hrt, rt, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (rtot_out) *rtot_out = hrt;
    if (hres[0] >= 0) {
        const int64_t who = c->hp.hop_index_mode == 0 ? hres[0] : hres[1];       // :327
        if (kind_out) *kind_out = 1;
        if (hop_k_out) *hop_k_out = who;
        TRY(hop_place_impl(c, sb, who, u1, surf));      // the surface flags of this state, just computed
    } else {
        if (kind_out) *kind_out = 0;
        if (hop_k_out) *hop_k_out = -1;
        TRY(deposit_place_host(c, sb, u1));                                       // :329
    }
    return kmc_relax(c, sb, 16, 1e-4, max_ls, stats);                                       // :94 / :256
}

}  // extern "C"
