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

// ---- grid-barrier cost (the other denominator of k_relax2: three of these per line search) -------------------------
// variant 0: the barrier of k_relax2 -- bar.sync, thread 0 arrives with red.release.gpu and polls with ld.acquire.gpu, bar.sync;
// variant 1: hierarchical -- barrier.cluster inside a thread-block cluster, ONE arrival and ONE poller per cluster, then
//            barrier.cluster again to release the other blocks of the cluster.
// Every block also writes and reads one word of data per round (what a real phase boundary has to publish).
__global__ void __launch_bounds__(512, 1) k_barrier_bench(unsigned int* __restrict__ counter, unsigned int* __restrict__ data, int iters,
                                                          int variant, int cluster) {
    unsigned int check = 0;
    const unsigned int nb = gridDim.x;
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x == 0) data[(it & 1) * 160 + blockIdx.x] = (unsigned int)it;      // double buffered: a block can be one round ahead
        if (variant == 0) {
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned int target = (unsigned int)(it + 1) * nb;
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
                unsigned int v;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
            }
            __syncthreads();
        } else if (variant == 2) {          // relaxed polls, one acquire fence after the last one
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned int target = (unsigned int)(it + 1) * nb;
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
                unsigned int v;
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
            __syncthreads();
        } else if (variant == 3) {          // arrivals spread over four counters in four L2 lines, relaxed polls of all four
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned int target = (unsigned int)(it + 1) * nb;
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter + 64 * (blockIdx.x & 3)) : "memory");
                unsigned int v0, v1, v2, v3;
                do {
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(counter) : "memory");
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(counter + 64) : "memory");
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v2) : "l"(counter + 128) : "memory");
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v3) : "l"(counter + 192) : "memory");
                } while ((int)(v0 + v1 + v2 + v3 - target) < 0);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
            __syncthreads();
        } else if (variant == 4) {          // every warp's lane 0 polls for itself: no second bar.sync
            __syncthreads();
            const unsigned int target = (unsigned int)(it + 1) * nb;
            if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
            if ((threadIdx.x & 31) == 0) {
                unsigned int v;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
            }
            __syncwarp();
        } else {
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
            unsigned int rank;
            asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
            if (rank == 0 && threadIdx.x == 0) {
                const unsigned int target = (unsigned int)(it + 1) * (nb / cluster);
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
                unsigned int v;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
            }
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
        if (threadIdx.x == 0) check += data[(it & 1) * 160 + (blockIdx.x + 1 + it) % nb] - (unsigned int)it;      // another block's word of this round
    }
    if (threadIdx.x == 0 && check != 0) data[320] = check;                                           // stale read detected
}

}  // namespace

// us per grid barrier of `blocks` x 512 threads (one block per SM); variant 1 uses clusters of `cluster` blocks
extern "C" int kmc_barrier_bench(KmcCtx* c, int32_t variant, int32_t cluster, int32_t iters, double* us_out, int32_t* stale_out) {
    if (!c || !us_out || iters < 1 || cluster < 1) return kmc_fail(KMC_E_ARG, "bad argument");
    DevGuard guard(c);
    int blocks = c->sm_count;
    if (variant == 1) blocks -= blocks % cluster;
    unsigned int* buf;
    TRY(slot(c, S_TMP0, 2048, &buf));
    TRY(kmc_zero_async(c, buf, 8192));
    unsigned int* counter = buf;
    unsigned int* data = buf + 512;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = variant == 1 ? cluster : 1; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = variant == 1 ? 2 : 1;
    cudaEvent_t e0 = kmc_event_from_pool(c), e1 = kmc_event_from_pool(c);
    int it = iters, var = variant, cl = cluster;
    cudaEventRecord(e0, c->stream);
    CK(cudaLaunchKernelEx(&cfg, k_barrier_bench, counter, data, it, var, cl));
    cudaEventRecord(e1, c->stream);
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    c->event_pool.push_back(e0);
    c->event_pool.push_back(e1);
    c->launches++;
    unsigned int stale = 0;
    CK(cudaMemcpy(&stale, data + 320, sizeof(unsigned int), cudaMemcpyDeviceToHost));
    if (stale_out) *stale_out = (int32_t)stale;
    *us_out = (double)ms * 1e3 / iters;
    return 0;
}

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
