// kmc_relax1.cu -- launcher of the first-generation persistent relaxation kernel (kmc_relax.cuh), kept as a
// cross-check engine (kmc_set_relax_mode 1); see kmc_relax2.cu for the product path.
#include "kmc_host.h"
#include "kmc_relax.cuh"

#include <algorithm>
#include <cstdlib>

using namespace kmc;

static const int LINE_BLOCKS = 148 * 8;

// 2DGrowth.py:97-166 as one persistent cooperative kernel (kmc_relax.cuh)
template <int G>
static int launch_relax(KmcCtx* c, RelaxArgs& R, int blocks) {
    void* args[] = {(void*)&R};
    c->launches++;
    cudaEvent_t ea = nullptr, eb = nullptr;
    if (c->timing) { ea = kmc_event_from_pool(c); eb = kmc_event_from_pool(c); cudaEventRecord(ea, c->stream); }
    CK(cudaLaunchCooperativeKernel((const void*)k_relax_persistent<G>, dim3(blocks), dim3(256), args, 0, c->stream));
    if (c->timing) { cudaEventRecord(eb, c->stream); c->recs.push_back({kmc_family_id(c, "relax_persistent"), ea, eb}); }
    return 0;
}

int kmc_relax_persistent1(KmcCtx* c, const KmcBins* sb, int32_t scale0, double threshold0, int64_t max_ls, KmcRelaxStats* stats) {
    KmcRelaxStats st{};
    const int n = (int)c->st_n;
    if (n == 0) { if (stats) *stats = st; return 0; }
    RelaxArgs R{};
    R.P = c->dp; R.sub = view_of(sb); R.n = n; R.ncell2 = c->dp.ncell + 2;
    // grid: all SMs x 2 blocks for large systems, fewer blocks (cheaper barriers) for small ones
    int blocks = std::min(2 * c->sm_count, std::max(1, (n * 32 + 255) / 256));
    const int gthreads = blocks * 256;
    int G = 32;
    while (G > 1 && (long long)n * G > gthreads) G >>= 1;
    R.x0 = (double2*)c->st_xy;
    TRY(slot(c, S_XN, (size_t)n, &R.xn));
    TRY(slot(c, S_XNP, (size_t)n, &R.xnp));
    TRY(slot(c, S_SN, (size_t)n, &R.sn));
    TRY(slot(c, S_F, (size_t)n, &R.F));
    TRY(slot(c, S_CELLOF, (size_t)n, &R.cell_of));
    TRY(slot(c, S_START, (size_t)R.ncell2, &R.start));
    TRY(slot(c, S_CURSOR, (size_t)R.ncell2, &R.cursor));
    TRY(slot(c, S_SIDX, (size_t)n, &R.sidx));
    TRY(slot(c, S_SXY, (size_t)n, &R.sxy));
    TRY(slot(c, S_FLAG, (size_t)n, &R.flag));
    TRY(slot(c, S_MOVERS, (size_t)n, &R.movers));
    TRY(slot(c, S_NMOV, 4, &R.n_movers));
    TRY(slot(c, S_LPART, (size_t)std::max(blocks, LINE_BLOCKS) * KC, &R.partial));
    TRY(slot(c, S_FIX, (size_t)n * KC, &R.fix));
    TRY(slot(c, S_RED, (size_t)blocks * 4, &R.red));
    unsigned char* sync;
    TRY(slot(c, S_SYNC, 256, &sync));
    CK(cudaMemsetAsync(sync, 0, 256, c->stream));
    R.barrier = (unsigned int*)sync;
    R.need_rebuild = (int*)(sync + 16);
    R.pair_counter = (unsigned long long*)(sync + 32);
    R.stats = (KmcRelaxStatsDev*)(sync + 64);
    R.phase_ns = (unsigned long long*)(sync + 128);
    R.scale0 = scale0; R.threshold0 = threshold0; R.max_ls = max_ls;
    switch (G) {
        case 32: TRY(launch_relax<32>(c, R, blocks)); break;
        case 16: TRY(launch_relax<16>(c, R, blocks)); break;
        case 8: TRY(launch_relax<8>(c, R, blocks)); break;
        case 4: TRY(launch_relax<4>(c, R, blocks)); break;
        case 2: TRY(launch_relax<2>(c, R, blocks)); break;
        default: TRY(launch_relax<1>(c, R, blocks)); break;
    }
    TRY(kmc_mail(c, sync, 256));
    if (getenv("KMC_PHASE_TIMING")) {
        const unsigned long long* ph = (const unsigned long long*)((const char*)c->h_mail + 128);
        const KmcRelaxStatsDev* d0 = (const KmcRelaxStatsDev*)((const char*)c->h_mail + 64);
        const double nls = d0->line_searches > 0 ? (double)d0->line_searches : 1.0;
        fprintf(stderr, "[kmc phase us/ls] energy %.2f | bar %.2f | fix+bar %.2f | sums+xnp %.2f | bar %.2f | forces %.2f | bar %.2f | ctl+update+bar %.2f  (ls=%lld blocks=%d G=%d)\n",
                ph[0] / nls * 1e-3, ph[1] / nls * 1e-3, ph[2] / nls * 1e-3, ph[3] / nls * 1e-3, ph[4] / nls * 1e-3, ph[5] / nls * 1e-3, ph[6] / nls * 1e-3, ph[7] / nls * 1e-3,
                (long long)d0->line_searches, blocks, G);
    }
    const unsigned long long visits = *(const unsigned long long*)((const char*)c->h_mail + 32);
    const KmcRelaxStatsDev* ds = (const KmcRelaxStatsDev*)((const char*)c->h_mail + 64);
    st.line_searches = ds->line_searches; st.restarts = ds->restarts; st.maxF = ds->maxF;
    st.final_threshold = ds->final_threshold; st.final_scale = ds->final_scale; st.status = ds->status;
    // ordered pair visits of one energy sweep x 20 candidates x line searches
    const long long per_sweep = (long long)visits + (c->dp.aa_mode == 0 ? (long long)n * (n - 1) / 2 : 0);
    st.pair_evals = st.line_searches * KC * per_sweep;
    if (stats) *stats = st;
    if (st.status != 0) return kmc_fail(st.status, "relaxation stopped at the line-search cap (%lld)", (long long)max_ls);
    return 0;
}

