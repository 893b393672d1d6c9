// kmc_line_kernels.cuh -- the standalone kernels of the host-driven line-search engine (kmc_line.cuh holds the
// device functions they share with the persistent relaxation kernels).  Included by kmc_b200.cu only.
#pragma once
#include "kmc_line.cuh"

namespace kmc {

// ================================================================================ standalone kernels
// lowest / highest occupied cell row of every column (tall, mostly empty grids: the home-cell loop visits only these rows)
static __global__ void __launch_bounds__(256) k_colrange_init(int2* __restrict__ cr, int nbx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nbx) cr[i] = make_int2(0x7fffffff, -1);
}
static __global__ void __launch_bounds__(256) k_colrange(DevParams P, const int* __restrict__ cell_of, int n, int2* __restrict__ cr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_of[i];
    if (c >= P.ncell) return;
    const int cx = c / P.nby, cy = c - cx * P.nby;
    atomicMin(&cr[cx].x, cy);
    atomicMax(&cr[cx].y, cy);
}

static __global__ void __launch_bounds__(256) k_line_mark(LineArgs A) {
    line_mark_movers(A, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

static __global__ void __launch_bounds__(1024) k_line_compact(LineArgs A) {
    __shared__ int s_scan[1024];
    line_compact_movers(A, s_scan);
}

static __global__ void __launch_bounds__(LT_THREADS) k_line_energy(LineArgs A) {
    __shared__ LineSmem S;
    double acc[KC];
#pragma unroll
    for (int a = 0; a < KC; ++a) acc[a] = 0.0;
    line_energy_cells(A, S, blockIdx.x, gridDim.x, acc);
    if (A.P.aa_mode == 0) {
        const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        line_energy_allpairs(A, gw, (gridDim.x * blockDim.x) >> 5, threadIdx.x & 31, acc);
    }
    __syncthreads();
    line_store_partial(acc, S.red, A.partial + (size_t)blockIdx.x * KC);
}

static __global__ void __launch_bounds__(128) k_line_fix(LineArgs A) {
    const int nm = *A.n_movers;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int w = gw; w < nm * KC; w += nw) {
        const double e = line_fix_one(A, w / KC, w % KC, lane);
        if (lane == 0) A.fix[w] = e;
    }
}

static __global__ void __launch_bounds__(KC * 32) k_line_reduce(LineArgs A, int nblocks) {
    __shared__ double s_e[KC];
    line_reduce(A, nblocks, s_e);
}

}  // namespace kmc
