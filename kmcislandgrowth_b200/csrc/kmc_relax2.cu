// kmc_relax2.cu -- launcher of k_relax2 (kmc_relax2.cuh), the product path of Relaxation (2DGrowth.py:97-166):
// one persistent cooperative kernel per relaxation on cell-sorted state.
//
// The launch is split into an enqueue and a collect step so that a whole KMC event (rate table -> selection ->
// placement -> relaxation, kmc_event.cu) can be queued on the stream without a host round trip in the middle: the
// adatom count may be read by the kernel from device memory (it depends on the event kind chosen on the device).
#include "kmc_host.h"
#include "kmc_relax2.cuh"

#include <algorithm>
#include <cstdlib>

using namespace kmc;

static const int LINE_BLOCKS = 148 * 8;
static_assert(R2_MAXBLK >= 148, "the max|F|^2 partials of every block must fit the per-lane registers");

static int launch_relax2(KmcCtx* c, Relax2Args& R, int blocks) {
    void* args[] = {(void*)&R};
    c->launches++;
    cudaEvent_t ea = nullptr, eb = nullptr;
    if (c->timing) { ea = kmc_event_from_pool(c); eb = kmc_event_from_pool(c); cudaEventRecord(ea, c->stream); }
    if (!c->relax2_smem_set) {          // per context (= per device): the staged operands exceed the 48 KB static limit
        CK(cudaFuncSetAttribute((const void*)k_relax2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Relax2Smem)));
        CK(cudaFuncSetAttribute((const void*)k_relax2<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Relax2Smem)));
        CK(cudaFuncSetAttribute((const void*)k_relax2<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Relax2Smem)));
        CK(cudaFuncSetAttribute((const void*)k_relax2<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Relax2Smem)));
        c->relax2_smem_set = true;
    }
    // tall cell grids (configs[4]: 1.6 M cells for 8192 adatoms) take the instantiation whose re-sort walks occupied rows only
    const bool tall = R.ncell2 > R2_TALL && blocks > 1;
    // the diagnostics (KMC_PHASE_TIMING, KMC_DEBUG_E=8) live in instantiations of their own: their timer reads are scheduling
    // fences that cost the production kernel 4 % even when they never execute
    const bool timing = R.phase_block >= 0 || R.debug == 8 || R.debug == 7;
    void (*kernel)(Relax2Args) = tall ? (timing ? k_relax2<true, true> : k_relax2<true, false>) : (timing ? k_relax2<false, true> : k_relax2<false, false>);
    if (R.sync_mode == 1 && !getenv("KMC_CLUSTER1")) {          // one block: a plain launch
        kernel<<<1, R2_THREADS, sizeof(Relax2Smem), c->stream>>>(R);
        CK(cudaGetLastError());
    } else if (R.sync_mode == 0) {
        CK(cudaLaunchCooperativeKernel((const void*)kernel, dim3(blocks), dim3(R2_THREADS), args, sizeof(Relax2Smem), c->stream));
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(R2_THREADS); cfg.dynamicSmemBytes = sizeof(Relax2Smem); cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = blocks; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, kernel, R));
    }
    if (c->timing) { cudaEventRecord(eb, c->stream); c->recs.push_back({kmc_family_id(c, "relax_persistent"), ea, eb}); }
    kmc_debug_sync(c, "k_relax2");
    return 0;
}

// launch shape: one block (<= 16 adatoms), one cluster of <= 8 blocks (<= 256 adatoms), else one block per SM
static void relax2_shape(const KmcCtx* c, int n, int* blocks, int* sync_mode) {
    int b = std::min(std::min(c->sm_count, R2_MAXBLK), std::max(1, (n + 15) / 16));
    int m = 0;
    if (b == 1) m = 1;
    else if (n <= 256 && !getenv("KMC_NO_CLUSTER")) { b = std::min(b, 8); m = 2; }
    *blocks = b;
    *sync_mode = m;
}

int kmc_relax2_enqueue(KmcCtx* c, const KmcBins* sb, int32_t scale0, double threshold0, int64_t max_ls, int n, const int* n_dev,
                       const int* status_dev) {
    int blocks, sync_mode;
    relax2_shape(c, n, &blocks, &sync_mode);
    const double t_a = kmc_now();
    if (c->item_cap == 0) c->item_cap = (size_t)n * 4 + (size_t)c->dp.ncell * 2 + 1024;
    if (const char* e = getenv("KMC_ITEM_CAP")) {      // tests: force the grow-and-relaunch path (once per context)
        if (!c->item_cap_forced) { c->item_cap = (size_t)std::max(1, atoi(e)); c->item_cap_forced = true; }
    }
    Relax2Args R{};
    R.P = c->dp; R.sub = view_of(sb); R.n = n; R.ncell2 = c->dp.ncell + 2;
    R.n_dev = n_dev; R.status_dev = status_dev;
    R.x0 = (double2*)c->st_xy;
    TRY(slot(c, S_POS0, (size_t)n, &R.pos[0]));
    TRY(slot(c, S_POS1, (size_t)n, &R.pos[1]));
    TRY(slot(c, S_SD0, (size_t)n, &R.sdir[0]));
    TRY(slot(c, S_SD1, (size_t)n, &R.sdir[1]));
    TRY(slot(c, S_OI0, (size_t)n, &R.oidx[0]));
    TRY(slot(c, S_OI1, (size_t)n, &R.oidx[1]));
    TRY(slot(c, S_CELLOF, (size_t)n, &R.cellof));
    TRY(slot(c, S_NBR, (size_t)n * 4, &R.nbr));
    TRY(slot(c, S_FLAG, (size_t)n + 32, &R.flag));        // the staged flag copy reads whole 16-byte groups
    TRY(slot(c, S_START, (size_t)R.ncell2, &R.start));
    TRY(slot(c, S_CURSOR, (size_t)R.ncell2, &R.cursor));
    TRY(slot(c, S_KEY, (size_t)n, &R.key));
    TRY(slot(c, S_PERM, (size_t)n, &R.perm));
    TRY(slot(c, S_CHUNKS, (size_t)c->dp.ncell * R2_PARTS + 2, &R.chunks));
    TRY(slot(c, S_COLR, (size_t)c->dp.nbx, &R.colr));
    TRY(slot(c, S_BX, (size_t)R2_MAXBLK, &R.tile_tot));
    TRY(slot(c, S_ITEMS, c->item_cap, &R.items));
    R.item_cap = (int)c->item_cap;
    TRY(slot(c, S_MOVERS, (size_t)n, &R.movers));
    TRY(slot(c, S_LPART, (size_t)std::max(blocks, LINE_BLOCKS) * R2_NE, &R.partial));
    TRY(slot(c, S_FIX, (size_t)n * KC, &R.fix));
    TRY(slot(c, S_RED, (size_t)blocks * 8, &R.red));
    unsigned char* sync;
    TRY(slot(c, S_SYNC, 512, &sync));
    TRY(kmc_zero_async(c, sync, 512));
    R.barrier = (unsigned int*)sync;
    R.abort = (int*)(sync + 12);
    R.need_rebuild = (int*)(sync + 16);
    R.n_items = (int*)(sync + 20);
    R.n_movers = (int*)(sync + 24);
    R.pair_counter = (unsigned long long*)(sync + 32);
    R.stats = (KmcRelaxStatsDev*)(sync + 64);
    R.phase_ns = (unsigned long long*)(sync + 128);
    R.hist = (unsigned long long*)(sync + 256);          // KMC_DEBUG_E=7 (shares the words of the unused phase slots 16-31)
    R.scale0 = scale0; R.threshold0 = threshold0; R.max_ls = max_ls;
    R.sync_mode = sync_mode;
    R.debug = getenv("KMC_DEBUG_E") ? atoi(getenv("KMC_DEBUG_E")) : 0;
    R.phase_block = !getenv("KMC_PHASE_TIMING") ? -1 : (getenv("KMC_PHASE_BLOCK") ? std::min(atoi(getenv("KMC_PHASE_BLOCK")), blocks - 1) : 0);
    if (R.debug == 8) {      // diagnostics: per-block phase times
        TRY(slot(c, S_TMP3, (size_t)3 * blocks, &R.blk_ns));
        CK(cudaMemsetAsync(R.blk_ns, 0, sizeof(unsigned long long) * 3 * blocks, c->stream));
    }
    const double t_b = kmc_now();
    TRY(launch_relax2(c, R, blocks));
    const double t_c = kmc_now();
    c->enq_s[2] += t_b - t_a;
    c->enq_s[3] += t_c - t_b;
    if (R.debug == 8) {
        std::vector<unsigned long long> h(3 * blocks);
        CK(cudaMemcpyAsync(h.data(), R.blk_ns, sizeof(unsigned long long) * 3 * blocks, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int ph = 0; ph < 2; ++ph) {
            double mn = 1e30, mx = 0, sum = 0; int arg = 0;
            for (int b = (blocks > 1 ? 1 : 0); b < blocks; ++b) { const double v = (double)h[2 * b + ph]; sum += v; if (v < mn) mn = v; if (v > mx) { mx = v; arg = b; } }
            fprintf(stderr, "[kmc2 per-block %s ns total] min %.0f mean %.0f max %.0f (block %d) block0 %.0f :", ph ? "G" : "E", mn, sum / std::max(1, blocks - 1), mx, arg, (double)h[ph]);
            for (int b = 0; b < blocks; b += 7) fprintf(stderr, " %.0f", (double)h[2 * b + ph] * 1e-3);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "[kmc2 per-block G us / smid]");
        for (int b = 0; b < blocks; ++b) fprintf(stderr, " %d:%.0f/%llu", b, (double)h[2 * b + 1] * 1e-3, h[2 * blocks + b]);
        fprintf(stderr, "\n");
    }
    // the 512-byte result block travels to the pinned mailbox behind the kernel; the caller synchronises
    TRY(kmc_to_mailbox_async(c, 0, sync, 512));
    CK(cudaGetLastError());
    return 0;
}

// after the stream has been synchronised: decode the mailbox of the last kmc_relax2_enqueue
int kmc_relax2_collect(KmcCtx* c, int n, KmcRelaxStats* stats, bool* again) {
    KmcRelaxStats st{};
    *again = false;
    const char* hm = (const char*)c->h_mail;
    const int debug = getenv("KMC_DEBUG_E") ? atoi(getenv("KMC_DEBUG_E")) : 0;
    if (debug == 7) {     // diagnostics: per-atom displacement over a whole line (19/20/scale*|s|), decades from 1e-6 A
        const unsigned long long* h = (const unsigned long long*)(hm + 256);
        fprintf(stderr, "[kmc2 disp hist]");
        for (int k = 0; k < 16; ++k) fprintf(stderr, " %llu", h[k]);
        fprintf(stderr, "\n");
    }
    const KmcRelaxStatsDev* ds = (const KmcRelaxStatsDev*)(hm + 64);
    if (ds->status == -4) {            // work list overflow: grow and run again (x0 is untouched)
        const int need = *(const int*)(hm + 20);
        c->item_cap = (size_t)need + need / 2 + 1024;
        *again = true;
        return 0;
    }
    const unsigned long long visits = *(const unsigned long long*)(hm + 32);
    {
        const unsigned long long* ph = (const unsigned long long*)(hm + 128);
        c->diag[0] = (int64_t)ph[8]; c->diag[1] = (int64_t)ph[9]; c->diag[2] = *(const int*)(hm + 20); c->diag[3] = 0;
    }
    if (getenv("KMC_PHASE_TIMING")) {
        const unsigned long long* ph = (const unsigned long long*)(hm + 128);
        const double nls = ds->line_searches > 0 ? (double)ds->line_searches : 1.0;
        fprintf(stderr, "[kmc2 E detail us/ls] control %.2f | items %.2f | reduce+store %.2f\n", ph[32] / nls * 1e-3, ph[33] / nls * 1e-3, ph[0] / nls * 1e-3);
        fprintf(stderr, "[kmc2 G detail us/ls] beta-reduce %.2f | rebuild-check %.2f | ranges+copy wait %.2f | adatom sweep %.2f | substrate sweep %.2f | group sums %.2f | update+flags %.2f | block max+store %.2f\n", ph[10] / nls * 1e-3, ph[11] / nls * 1e-3, ph[34] / nls * 1e-3, ph[12] / nls * 1e-3, ph[13] / nls * 1e-3, ph[14] / nls * 1e-3, ph[15] / nls * 1e-3, ph[5] / nls * 1e-3);
        fprintf(stderr, "[kmc2 phase us/ls] E %.2f | bar %.2f | fix+bar %.2f | S %.2f | bar %.2f | G %.2f | bar %.2f | rebuild %.2f (%.1f us each, %lld) movers/ls %.1f (ls=%lld restarts=%lld n=%d items=%d)\n",
                (ph[0] + ph[32] + ph[33]) / nls * 1e-3, ph[1] / nls * 1e-3, ph[2] / nls * 1e-3, ph[3] / nls * 1e-3, ph[4] / nls * 1e-3, (ph[5] + ph[34] + ph[10] + ph[11] + ph[12] + ph[13] + ph[14] + ph[15]) / nls * 1e-3,
                ph[6] / nls * 1e-3, ph[7] / nls * 1e-3, ph[8] ? ph[7] * 1e-3 / ph[8] : 0.0, (long long)ph[8], ph[9] / nls, (long long)ds->line_searches, (long long)ds->restarts, n, *(const int*)(hm + 20));
    }
    if (ds->status == KMC_E_TIMEOUT)
        return kmc_fail(KMC_E_TIMEOUT, "relaxation kernel: a device-side wait timed out (watchdog word %d: 1 grid barrier, 3 bulk copy)", *(const int*)(hm + 12));
    st.line_searches = ds->line_searches; st.restarts = ds->restarts; st.maxF = ds->maxF;
    st.final_threshold = ds->final_threshold; st.final_scale = ds->final_scale; st.status = ds->status;
    const long long per_sweep = (long long)visits + (c->dp.aa_mode == 0 ? (long long)n * (n - 1) / 2 : 0);
    st.pair_evals = st.line_searches * KC * per_sweep;
    if (stats) *stats = st;
    return 0;
}

int kmc_relax2_persistent(KmcCtx* c, const KmcBins* sb, int32_t scale0, double threshold0, int64_t max_ls, KmcRelaxStats* stats) {
    KmcRelaxStats st{};
    const int n = (int)c->st_n;
    if (n == 0) { if (stats) *stats = st; return 0; }
    for (int attempt = 0; attempt < 8; ++attempt) {
        TRY(kmc_relax2_enqueue(c, sb, scale0, threshold0, max_ls, n, nullptr, nullptr));
        CK(cudaStreamSynchronize(c->stream));
        bool again = false;
        TRY(kmc_relax2_collect(c, n, &st, &again));
        if (again) continue;
        if (stats) *stats = st;
        if (st.status != 0) return kmc_fail(st.status, "relaxation stopped at the line-search cap (%lld)", (long long)max_ls);
        return 0;
    }
    return kmc_fail(KMC_E_CAPACITY, "relaxation work list kept overflowing");
}

extern "C" int kmc_relax_diag(KmcCtx* c, int64_t out[4]) {
    if (!c || !out) return kmc_fail(KMC_E_ARG, "NULL argument");
    for (int k = 0; k < 3; ++k) out[k] = c->diag[k];
    out[3] = c->rc_last;
    return 0;
}
