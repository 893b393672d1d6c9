// kmc_peak.cu -- measured denominators for the pair kernels' roofline: the non-tensor FP64 (DFMA) issue rate of the
// device, measured with register-resident FMA chains at full occupancy (BASELINE.md section 3 asks for 24*pairs/time
// "against the non-tensor FP64 peak"; MEASURED_PEAKS.json only holds the HBM copy rate and the bf16 GEMM rate).
#include "kmc_host.h"

#include <algorithm>

namespace {

constexpr int PEAK_CHAINS = 8;        // independent accumulators per thread (hides the DFMA latency at 64 warps / SM)

__global__ void __launch_bounds__(256) k_fp64_peak(double* __restrict__ out, int iters, double a, double b) {
    double acc[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) acc[k] = (double)(threadIdx.x + k) * 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < PEAK_CHAINS; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) s += acc[k];
    if (s == 12345.678) out[0] = s;       // never true: keeps the chains alive
}

}  // namespace

extern "C" int kmc_fp64_peak(KmcCtx* c, int32_t repeats, double* tflops_out, double* ms_out) {
    if (!c || !tflops_out) return kmc_fail(KMC_E_ARG, "NULL argument");
    DevGuard guard(c);
    double* d;
    TRY(slot(c, S_TMP0, 8, &d));
    const int blocks = c->sm_count * 8, iters = 4096;
    const double flops = 2.0 * 4 * PEAK_CHAINS * (double)iters * 256.0 * blocks;
    cudaEvent_t e0 = kmc_event_from_pool(c), e1 = kmc_event_from_pool(c);
    double best = 1e30;
    for (int r = 0; r < std::max(repeats, 1) + 1; ++r) {            // the first run warms up
        cudaEventRecord(e0, c->stream);
        k_fp64_peak<<<blocks, 256, 0, c->stream>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1, c->stream);
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) best = std::min(best, (double)ms);
    }
    c->event_pool.push_back(e0);
    c->event_pool.push_back(e1);
    c->launches += std::max(repeats, 1) + 1;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}
