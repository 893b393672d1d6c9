// kmc_relax.cuh -- the reference's Relaxation (2DGrowth.py:97-166) as ONE persistent cooperative kernel:
// 20-point line search + pseudo-CG + the stall / out-of-bounds / threshold restart rules, with the
// whole trajectory state resident in HBM/L2 and grid-wide barriers between the phases of a line search.
// One launch replaces ~10 launches and one host round trip per line search.
//
// Phases of one line search (barriers marked ||):
//   energy over home cells (+ ordered mover compaction by block 0) || mover fix-up (only if any) ||
//   candidate sums, first argmin, x_np, cell-change check, sorted copy || forces at x_np + reductions ||
//   control (every block, identical arithmetic), CG update, mover flags of the next line ||
#pragma once
#include "kmc_line.cuh"

namespace kmc {

struct KmcRelaxStatsDev {
    long long line_searches, restarts;
    double maxF, final_threshold;
    int final_scale, status;
};

struct RelaxArgs {
    DevParams P;
    CellView sub;
    int n;
    int ncell2;
    // state, original order
    double2* x0;      // positions handed to the call (restart point)
    double2* xn;
    double2* xnp;
    double2* sn;
    double2* F;
    // base cell list of the current iterate (mutable)
    int* cell_of;     // [n]
    int* start;       // [ncell+2]
    int* cursor;      // [ncell+2]
    int* sidx;        // [n]
    double2* sxy;     // [n]
    // line search scratch
    uint8_t* flag;
    int* movers;
    int* n_movers;
    double* partial;  // [gridDim][KC]
    double* fix;      // [n][KC]
    double* red;      // [gridDim][4]  maxF, miny, top, bot
    unsigned int* barrier;     // zeroed before launch
    int* need_rebuild;         // zeroed before launch
    unsigned long long* pair_counter;   // zeroed before launch
    // parameters
    int scale0;
    double threshold0;
    long long max_ls;
    // results
    KmcRelaxStatsDev* stats;
    unsigned long long* phase_ns;   // [8] block 0's time per phase (diagnostics)
};

// grid-wide barrier on a monotonically increasing 32-bit counter (all blocks are co-resident: cooperative launch);
// arrival counts are compared modulo 2^32, so the counter may wrap in a very long relaxation
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // arrive with release semantics (cumulative over the block's writes ordered by the bar.sync above), then poll
        // with acquire loads: no separate membar.gl on either side of the atomic
        const unsigned int target = (epoch + 1u) * gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while ((int)(v - target) < 0);          // wrap-safe: the counter only ever runs a few arrivals ahead of the target
    }
    ++epoch;
    __syncthreads();
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define PHASE_MARK(k)                                              \
    if (gtid == 0) {                                               \
        const unsigned long long _t = gtimer();                    \
        R.phase_ns[k] += _t - t_mark;                              \
        t_mark = _t;                                               \
    }

// ---- cell list rebuild from positions `pos` (original order), 4 internal barriers ----------------
__device__ __forceinline__ void relax_rebuild_cells(const RelaxArgs& R, const double2* pos, unsigned int& epoch, int* s_scan) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int c = gtid; c < R.ncell2; c += gthreads) R.start[c] = 0;
    grid_barrier(R.barrier, epoch);
    for (int i = gtid; i < R.n; i += gthreads) {
        int nx, ny;
        bin_index(R.P, pos[i].x, pos[i].y, nx, ny);
        const int c = flat_cell(R.P, nx, ny);
        R.cell_of[i] = c;
        atomicAdd(&R.start[1 + c], 1);
    }
    grid_barrier(R.barrier, epoch);
    if (blockIdx.x == 0) {      // inclusive scan of start[1..ncell+1] by one block
        const int T = blockDim.x, t = threadIdx.x;
        const int m = R.ncell2 - 1;
        const int per = (m + T - 1) / T;
        const int lo = 1 + t * per, hi = min(lo + per, R.ncell2);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += R.start[i];
        s_scan[t] = sum;
        __syncthreads();
        for (int off = 1; off < T; off <<= 1) {
            int v = (t >= off) ? s_scan[t - off] : 0;
            __syncthreads();
            s_scan[t] += v;
            __syncthreads();
        }
        int run = s_scan[t] - sum;
        for (int i = lo; i < hi; ++i) { run += R.start[i]; R.start[i] = run; }
        __syncthreads();
        for (int i = t; i < R.ncell2; i += T) R.cursor[i] = R.start[i];
    }
    grid_barrier(R.barrier, epoch);
    for (int i = gtid; i < R.n; i += gthreads) {
        const int slot = atomicAdd(&R.cursor[R.cell_of[i]], 1);
        R.sidx[slot] = i;
    }
    grid_barrier(R.barrier, epoch);
    for (int c = gtid; c < R.ncell2 - 1; c += gthreads) {
        const int lo = R.start[c], hi = R.start[c + 1];
        sort_indices(R.sidx + lo, hi - lo);
        for (int s = lo; s < hi; ++s) R.sxy[s] = pos[R.sidx[s]];
    }
    grid_barrier(R.barrier, epoch);
}

// ---- forces at the sorted positions (cell list valid for them), G lanes per atom ------------------
// out[idx] = F_aa + F_as (2DGrowth.py:120); optional reductions against xn: maxF, miny, top, bot
template <int G>
__device__ __forceinline__ void relax_forces(const RelaxArgs& R, CellView ad, double2* out, const double2* xn, bool reduce, bool count_pairs,
                             double* s_red) {
    const int lane = threadIdx.x & (G - 1);
    const int ggroups = (gridDim.x * blockDim.x) / G;
    double mf = 0.0, my = INFINITY, top = 0.0, bot = 0.0;
    unsigned long long visits = 0;
    const int rounds = (R.n + ggroups - 1) / ggroups;
    for (int rd = 0; rd < rounds; ++rd) {
        const int gid = rd * ggroups + (blockIdx.x * blockDim.x + threadIdx.x) / G;
        const bool live = gid < R.n;
        double ax = 0, ay = 0, sx = 0, sy = 0;
        double2 pi = make_double2(0, 0);
        int idx = 0;
        if (live) {
            pi = ad.sxy[gid];
            idx = ad.sidx[gid];
            const Stencil st = make_stencil(R.P, pi.x, pi.y);
            if (R.P.aa_mode == 0) {
                for (int s = lane; s < R.n; s += G) acc_force(R.P, pi.x, pi.y, ad.sxy[s], R.P.E_a, R.P.ra2, ax, ay);
            } else {
                visit_stencil(R.P, ad, st, lane, G, [&](int s) {
                    acc_force(R.P, pi.x, pi.y, ad.sxy[s], R.P.E_a, R.P.ra2, ax, ay);
                    ++visits;
                });
            }
            visit_stencil(R.P, R.sub, st, lane, G, [&](int s) {
                acc_force(R.P, pi.x, pi.y, R.sub.sxy[s], R.P.E_as, R.P.ras2, sx, sy);
                ++visits;
            });
        }
        ax = group_sum<G>(ax); ay = group_sum<G>(ay);
        sx = group_sum<G>(sx); sy = group_sum<G>(sy);
        if (live && lane == 0) {
            const double fx = ax + sx, fy = ay + sy;
            out[idx] = make_double2(fx, fy);
            if (reduce) {
                mf = fmax(mf, fx * fx + fy * fy);                      // :129,160
                my = fmin(my, pi.y);                                   // :137
                const double2 p0 = xn[idx];
                top += pi.x * (pi.x - p0.x) + pi.y * (pi.y - p0.y);    // :143
                bot += p0.x * p0.x + p0.y * p0.y;                      // :144
            }
        }
    }
    if (reduce) {
        const double r_mf = block_max(mf, s_red);
        const double r_my = block_min(my, s_red);
        const double r_top = block_sum(top, s_red);
        const double r_bot = block_sum(bot, s_red);
        if (threadIdx.x == 0) {
            double* o = R.red + (size_t)blockIdx.x * 4;
            o[0] = r_mf; o[1] = r_my; o[2] = r_top; o[3] = r_bot;
        }
    }
    if (count_pairs) {
        // ordered stencil visits of one sweep; all-pairs adatom part added by the host
        for (int o = 16; o > 0; o >>= 1) visits += __shfl_xor_sync(0xffffffffu, visits, o);
        if ((threadIdx.x & 31) == 0 && visits) atomicAdd(R.pair_counter, visits);
    }
}

struct RelaxSmem {
    LineSmem line;
    int scan[1024];
    double e[KC];
    double ctl[4];
};

template <int G>
static __global__ void __launch_bounds__(256, 2) k_relax_persistent(RelaxArgs R) {
    __shared__ RelaxSmem S;
    unsigned int epoch = 0;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int gwarp = gtid >> 5, nwarps = gthreads >> 5;

    LineArgs A{};
    A.n_compute = -1;
    A.colrange = nullptr;
    A.P = R.P; A.sub = R.sub; A.n = R.n;
    A.ad.start = R.start; A.ad.sxy = R.sxy; A.ad.sidx = R.sidx; A.ad.n = R.n;
    A.x = R.xn; A.s = R.sn;
    A.flag = R.flag; A.movers = R.movers; A.n_movers = R.n_movers; A.partial = R.partial; A.fix = R.fix;
    A.e_out = nullptr; A.amin_out = nullptr;

    int scale = R.scale0;
    double threshold = R.threshold0;
    long long ls = 0, restarts = 0;
    int status = 0;
    double maxF = 0.0, lastmaxF = 0.0;
    bool first_pass = true;
    unsigned long long t_mark = gtimer();

    // one line search from (xn, sn) with the current scale; leaves xnp, F(xnp) and the reductions in R.red
    auto line_search = [&]() {
        A.scale = (double)scale;
        A.inv_step = 1.0 / (20.0 * (double)scale);
        double acc[KC];
#pragma unroll
        for (int a = 0; a < KC; ++a) acc[a] = 0.0;
        if (blockIdx.x == 0) line_compact_movers(A, S.scan);
        line_energy_cells(A, S.line, blockIdx.x, gridDim.x, acc);
        if (R.P.aa_mode == 0) line_energy_allpairs(A, gwarp, nwarps, lane, acc);
        __syncthreads();
        line_store_partial(acc, S.line.red, R.partial + (size_t)blockIdx.x * KC);
        PHASE_MARK(0)
        grid_barrier(R.barrier, epoch);
        PHASE_MARK(1)
        const int nm = *R.n_movers;
        if (nm > 0) {
            for (int w = gwarp; w < nm * KC; w += nwarps) {
                const double e = line_fix_one(A, w / KC, w % KC, lane);
                if (lane == 0) R.fix[w] = e;
            }
            grid_barrier(R.barrier, epoch);
            PHASE_MARK(2)
        }
        // candidate sums: every block forms them in the same order -> identical argmin everywhere
        {
            const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
            for (int a = w; a < KC; a += nw) {
                double v = 0.0;
                for (int b = lane; b < (int)gridDim.x; b += 32) v += R.partial[(size_t)b * KC + a];
                for (int r = lane; r < nm; r += 32) v += R.fix[(size_t)r * KC + a];
                v = group_sum<32>(v);
                if (lane == 0) S.e[a] = v;
            }
            __syncthreads();
        }
        int amin = 0;
        {
            double em = S.e[0];
            for (int a = 1; a < KC; ++a)
                if (S.e[a] < em) { em = S.e[a]; amin = a; }       // first argmin, 2DGrowth.py:123,155
        }
        // x_np = candidate amin; sorted copy; does the base cell list still describe it?
        for (int slot = gtid; slot < R.n; slot += gthreads) {
            const int idx = R.sidx[slot];
            const double2 p = cand2(R.xn[idx], R.sn[idx], amin, (double)scale);
            R.xnp[idx] = p;
            R.sxy[slot] = p;
            if (R.flag[idx]) {        // only movers can have left their base cell
                int nx, ny;
                bin_index(R.P, p.x, p.y, nx, ny);
                if (flat_cell(R.P, nx, ny) != R.cell_of[idx]) *R.need_rebuild = 1;
            }
        }
        PHASE_MARK(3)
        grid_barrier(R.barrier, epoch);
        PHASE_MARK(4)
        if (*R.need_rebuild) {
            relax_rebuild_cells(R, R.xnp, epoch, S.scan);
            if (gtid == 0) *R.need_rebuild = 0;
            // (the next barrier orders this reset before any later write)
        }
        relax_forces<G>(R, A.ad, R.F, R.xn, true, false, S.line.red);
        PHASE_MARK(5)
        grid_barrier(R.barrier, epoch);
        PHASE_MARK(6)
        // every block reduces the per-block results in the same order
        {
            double mf = 0.0, my = INFINITY, top = 0.0, bot = 0.0;
            if (threadIdx.x < 32) {
                for (int b = lane; b < (int)gridDim.x; b += 32) {
                    const double* o = R.red + (size_t)b * 4;
                    mf = fmax(mf, o[0]); my = fmin(my, o[1]); top += o[2]; bot += o[3];
                }
                for (int o = 16; o > 0; o >>= 1) {
                    mf = fmax(mf, __shfl_xor_sync(0xffffffffu, mf, o));
                    my = fmin(my, __shfl_xor_sync(0xffffffffu, my, o));
                }
                top = group_sum<32>(top); bot = group_sum<32>(bot);
                if (threadIdx.x == 0) { S.ctl[0] = mf; S.ctl[1] = my; S.ctl[2] = top; S.ctl[3] = bot; }
            }
            __syncthreads();
        }
        ++ls;
    };

    for (;;) {                                                     // one pass == one (recursive) call of Relaxation
        if (scale > 1024) { scale = 16; threshold *= 10; }          // :111-113
        for (int i = gtid; i < R.n; i += gthreads) R.xn[i] = R.x0[i];
        grid_barrier(R.barrier, epoch);
        relax_rebuild_cells(R, R.xn, epoch, S.scan);
        relax_forces<G>(R, A.ad, R.sn, R.xn, false, first_pass, S.line.red);      // dx0 ; sn = dx0   :120,125
        first_pass = false;
        grid_barrier(R.barrier, epoch);
        A.scale = (double)scale;
        line_mark_movers(A, gtid, gthreads);
        grid_barrier(R.barrier, epoch);
        line_search();                                              // :121-129
        maxF = S.ctl[0];
        lastmaxF = maxF;
        bool restart = false;
        while (maxF > threshold) {                                  // :131
            if (R.max_ls > 0 && ls >= R.max_ls) { status = -6; break; }
            if (S.ctl[1] > R.P.Dmax) { restart = true; break; }     // :137-140
            const double beta = S.ctl[2] / S.ctl[3];                // :143-145
            __syncthreads();
            A.scale = (double)scale;
            for (int i = gtid; i < R.n; i += gthreads) {            // :147-150 + mover flags of the new line
                const double2 f = R.F[i], s = R.sn[i];
                const double2 sn = make_double2(__dadd_rn(f.x, __dmul_rn(beta, s.x)), __dadd_rn(f.y, __dmul_rn(beta, s.y)));
                R.sn[i] = sn;
                const double2 p0 = R.xnp[i];
                R.xn[i] = p0;
                const double2 pe = cand2(p0, sn, KC - 1, (double)scale);
                int nx0, ny0, nx1, ny1;
                bin_index(R.P, p0.x, p0.y, nx0, ny0);
                bin_index(R.P, pe.x, pe.y, nx1, ny1);
                const bool in0 = nx0 >= 0 && nx0 < R.P.nbx && ny0 >= 0 && ny0 < R.P.nby;
                R.flag[i] = (in0 && nx0 == nx1 && ny0 == ny1) ? 0 : 1;
            }
            grid_barrier(R.barrier, epoch);
            PHASE_MARK(7)
            line_search();                                          // :152-160
            maxF = S.ctl[0];
            if (fabs(lastmaxF - maxF) < 1e-8) { restart = true; break; }   // :162-164
            lastmaxF = maxF;
        }
        if (status != 0 || !restart) break;
        ++restarts;
        scale *= 2;
    }
    // result: PutAllInBox(xnp)   :166   (written to x0, the trajectory state)
    grid_barrier(R.barrier, epoch);
    for (int i = gtid; i < R.n; i += gthreads) {
        double2 p = R.xnp[i];
        p.x = put_in_box(p.x, R.P.L, R.P.halfL);
        R.x0[i] = p;
    }
    if (gtid == 0) {
        R.stats->line_searches = ls;
        R.stats->restarts = restarts;
        R.stats->maxF = maxF;
        R.stats->final_threshold = threshold;
        R.stats->final_scale = scale;
        R.stats->status = status;
    }
}

}  // namespace kmc
