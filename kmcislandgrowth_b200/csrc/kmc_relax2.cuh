// kmc_relax2.cuh -- Relaxation (2DGrowth.py:97-166) as one persistent cooperative kernel, second
// generation: the trajectory lives in CELL-SORTED order for the whole relaxation (positions,
// directions, original indices), so every pass is coalesced and free of gathers; the line-search
// energy is split into work items (<= 64 pairs) of which every block owns a contiguous range (its
// operands are staged in shared memory by TMA bulk copies requested one phase ahead), every sum is
// formed in a fixed order (bit-reproducible runs); the 20 candidate energies of a pair are ONE
// Taylor polynomial (kmc_line.cuh: pair_line_poly) whose coefficients are summed over the pairs;
// the CG update and the mover flags of the next line are fused into the force phase, leaving three
// grid barriers per line search:
//
//   E: polynomial / per-candidate sums over the work items; block 0: ordered mover list and, for
//      up to R2_FIX_INLINE movers, their per-candidate evaluation                             ||
//      [F: mover fix-up phase, only for more movers                                           ||]
//   S: sums over blocks, candidate energies, first argmin, x_np, top/bot/min-y partials, cell check ||
//      [rebuild of the sorted state when a mover really changed cell at the winner]
//   G: forces at x_np (one pass, T/atoms lanes per atom), max|F|^2 partials, s <- F + beta s,
//      mover flags of the next line                                                           ||
//   C: control (every block, identical arithmetic): converged / stalled / out of bounds / continue;
//      a restart from the original positions reuses the sorted layout and F(x0) when it can
//
// Measured in round 2 and NOT adopted (git history, commit 18b40f1): a two-barrier variant -- every block forms the x_np of its
// own AND its halo slots from staged x_n / s_n (two-segment stage at the periodic seam), so the force sweep needs no
// published x_np; the beta partials travel through a split barrier whose release is issued by a control warp.  Steady
// state 25.3 us per line search against 26.5 us here, but 100.5 against 101.5 events/s on the bench (restarts and re-sorts
// take its slower fall-back path), and on configs[4] the first line search after a restart formed inconsistent halo copies
// (found by a cross-check of the local copies against the published ones; cause not established).  Also measured and
// rejected: exchanging the 27 energy partials "data as flag" (relaxed 8-byte stores polled against a sentinel, no fence):
// 148 blocks polling the same cache lines serialise at the L2 slices -- 34.9 us per line search.
#pragma once
#include "kmc_line.cuh"
#include "kmc_relax.cuh"

namespace kmc {

#ifndef KMC_R2_ILP2
#define KMC_R2_ILP2 1
#endif
#ifndef KMC_R2_THREADS
#define KMC_R2_THREADS 512
#endif
constexpr int R2_THREADS = KMC_R2_THREADS;
constexpr int R2_NE = KC + PD + 1;  // values reduced per line search: KC per-candidate sums (pairs that are not expanded) + the polynomial
#ifndef KMC_R2_FB
#define KMC_R2_FB 4
#endif
constexpr int R2_FB = KMC_R2_FB;       // neighbours fetched and evaluated per batch and lane in the force sweep (measured G phase,
                                       // us per line search: 8 -> 8.5, 7 -> 8.4, 6 -> 8.2, 5 -> 8.2, 4 -> 8.0)
#ifndef KMC_R2_CHUNK
#define KMC_R2_CHUNK 64
#endif
constexpr int R2_CHUNK = KMC_R2_CHUNK;        // pairs per work item
constexpr int R2_WI = 12;          // work items of a warp cached in shared memory between rebuilds
constexpr int R2_PARTS = 8;         // per home cell: adatom columns 0..3, substrate columns 0..3
constexpr int R2_TALL = 16384;      // cell grids above this size (configs[4]: 1.6 M cells, a few thousand occupied) take the rebuild
                                    // path whose scans run on the whole grid and whose per-cell work visits occupied rows only

struct WorkItem {
    int hb, nh;            // home atoms: slots [hb, hb+nh)
    int b0, n0, b1, n1, b2; // partner slot ranges [b0,b0+n0) [b1,b1+n1) [b2,...) in stencil order
    int kind;              // bit 0..2: range k is the home cell itself (pair only if slot_i < slot_j); bit 3: substrate partners
    int begin;             // first pair of the chunk in the (home x partner) pair space, pair q = i + nh*j
    int count;             // pairs in the chunk
    float inv_nh;          // 1/nh for the index split
    int n2;                // length of the third partner range
};

constexpr int R2_MAXBLK = 160;      // grids have at most this many blocks (the control step keeps 5 partials per lane)

struct Relax2Args {
    DevParams P;
    CellView sub;
    int n, ncell2;          // n: adatoms (an upper bound when n_dev is set)
    const int* n_dev;       // optional: the adatom count lives in device memory (on-device event path, kmc_event.cu)
    const int* status_dev;  // optional: the kernel does nothing when *status_dev != 0 (the placement before it failed)
    double2* x0;            // original order: input and output of the relaxation
    double2* pos[2];        // sorted: base iterate / candidate winner (ping-pong)
    double2* sdir[2];       // sorted: search direction (ping-pong across rebuilds)
    int* oidx[2];           // sorted: original index
    int* cellof;            // sorted: cell of the slot in the current cell list
    int4* nbr;              // sorted, 4 x int4 per slot: adatom partner ranges b[4], e[4], substrate ranges b[4], e[4] of the
                            // slot's stencil columns (valid until the next rebuild; b[0] = -1: use the generic stencil walk)
    uint8_t* flag;          // sorted: 1 = mover on the current line
    int* start;             // [ncell+2]
    int* cursor;            // [ncell+2]
    int* key;               // [n] scratch: cell of source element
    int* perm;              // [n] scratch: source element of a slot
    int2* colr;             // tall grids only (ncell2 > R2_TALL): lowest / highest occupied cell row of every column
    int* tile_tot;          // tall grids only: per-block totals of the grid-wide scans
    int* chunks;            // [ncell*R2_PARTS + 1] items per (cell, part), then their exclusive scan
    WorkItem* items;        // work list
    int* n_items;
    int item_cap;
    int* movers;            // slots of the movers, ascending
    int* n_movers;
    double* partial;        // [R2_NE][gridDim] partial sums (transposed: coalesced reads): KC candidate sums of the pairs evaluated
                            // per candidate, then the PD+1 polynomial coefficients of all other pairs (pair_line_poly)
    double* fix;            // [n][KC]
    double* red;            // [gridDim][4] top, bot, miny, maxF
    unsigned int* barrier;
    int* abort;                 // watchdog word: != 0 once a spin loop timed out
    int* need_rebuild;
    unsigned long long* pair_counter;
    int scale0;
    double threshold0;
    long long max_ls;
    KmcRelaxStatsDev* stats;
    unsigned long long* phase_ns;
    unsigned long long* hist;   // diagnostics (debug == 7): histogram of per-atom line displacements
    unsigned long long* blk_ns; // diagnostics (debug == 8): [gridDim][2] time every block spends in the energy / force phase
    int phase_block;        // which block's thread 0 keeps the phase timers (diagnostics)
    int sync_mode;          // 0: global atomic barrier (cooperative launch), 1: one block (__syncthreads), 2: one thread-block cluster (hardware barrier)
    int debug;              // diagnostics only: 1 = energy phase skips the pair arithmetic, 2 = skips the work items
};

// barrier between the phases of a line search, by launch shape:
//  - many blocks: grid-wide barrier on a global counter (cooperative launch), ~2 us;
//  - up to 8 blocks (small systems: the replica sweeps): ONE thread-block cluster, hardware barrier.cluster
//    with release/acquire semantics (~0.2 us);
//  - one block: __syncthreads (global writes of a block are visible to the block after it; L1 stays warm).
// Watchdog: no spin loop of this kernel waits forever.  A poll that has not succeeded after R2_SPIN_NS raises the abort
// word (first writer wins: 1 = grid barrier, 3 = bulk-copy mbarrier) and falls through; every later poll sees the word
// and falls through as well, the line search returns, and the launch ends with status KMC_E_TIMEOUT instead of hanging
// the device (a hung cooperative kernel can only be removed by resetting the GPU).
constexpr unsigned long long R2_SPIN_NS = 4000000000ull;
// (the timer is first read after the first batch of failed polls, t0 == 0 meaning "not started": a timer read on the fast
// path of every barrier and copy wait is a volatile instruction the scheduler cannot move work across)
__device__ __forceinline__ bool relax2_spin_expired(const Relax2Args& R, unsigned long long& t0, int who) {
    if (*(volatile int*)R.abort != 0) return true;
    const unsigned long long now = gtimer();
    if (t0 == 0ull) { t0 = now; return false; }
    if (now - t0 > R2_SPIN_NS) { atomicCAS(R.abort, 0, who); return true; }
    return false;
}
// returns true (block-uniform) when the watchdog fired
__device__ __forceinline__ bool relax2_barrier(const Relax2Args& R, unsigned int& epoch) {
    if (R.sync_mode == 1) {
        __syncthreads();
        return false;
    } else if (R.sync_mode == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        return false;
    } else {
        bool dead = false;
        __syncthreads();
        if (threadIdx.x == 0) {
            // arrive with release semantics (cumulative over the block's writes ordered by the bar.sync above), then poll
            // with acquire loads: no separate membar.gl on either side of the atomic
            const unsigned int target = (epoch + 1u) * gridDim.x;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(R.barrier) : "memory");
            unsigned int v;
            unsigned long long t0 = 0ull;
            unsigned int spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(R.barrier) : "memory");
                if ((++spins & 1023u) == 0 && relax2_spin_expired(R, t0, 1)) { dead = true; break; }
            } while ((int)(v - target) < 0);
        }
        ++epoch;
        return __syncthreads_or(dead) != 0;
    }
}

constexpr int R2_STAGE = 1024;       // atoms of the neighbourhood staged in shared memory by the force phase (16 KB)

constexpr int R2_FIX_INLINE = 2;     // up to this many movers are evaluated by block 0 inside the energy phase
constexpr int R2_NBRC = 80;          // atoms of the block's chunk whose stencil ranges are cached in shared memory

constexpr int R2_SUBSTAGE = 512;     // substrate atoms cached per block (static between rebuilds)

struct Relax2Smem {
    alignas(16) double2 stage[R2_STAGE];     // positions of slots [smin, smax): one TMA 1-D bulk copy per block and line search
    // energy phase: operands of this block's work items, slots [emin, emax): positions, directions, mover flags (three
    // bulk copies completing on mbar_e, requested as soon as the previous line search has published them)
    alignas(16) double2 epos[R2_STAGE];
    alignas(16) double2 edir[R2_STAGE];
    alignas(16) uint8_t eflag[R2_STAGE + 16];
    alignas(16) double2 esub[R2_SUBSTAGE];   // substrate atoms [submin, submax) of the block's substrate items
    union {                                  // phase-exclusive scratch
        struct {                             // force phase: lane partials (adatom / substrate part), summed per atom in lane order
            double2 fpa[R2_THREADS];
            double2 fps[R2_THREADS];
        } g;
        struct {                             // energy / sums phases, rebuild
            double wred[R2_THREADS / 32][R2_NE];
            int scan[R2_THREADS];
        } e;
    } u;
    alignas(16) int4 nbrc[R2_NBRC][4];       // stencil ranges (R.nbr) of the first R2_NBRC atoms of the chunk, valid until the next rebuild
    WorkItem wit[R2_THREADS / 32][R2_WI];    // this block's warps' first work items (static dealing: valid until the next rebuild)
    alignas(8) unsigned long long mbar;      // mbarrier the bulk copy completes on
    alignas(8) unsigned long long mbar_e;    // mbarrier of the energy-phase copies
    unsigned long long ph[48];               // KMC_PHASE_TIMING: phase times of the observed block, flushed to R.phase_ns at the end
    int smin, smax;
    int emin, emax, submin, submax;
    double e[R2_NE];
    int amin;
    double ctl[4];
    double red[32];
};


// stencil of an in-grid cell (same rule as make_stencil, driven by cell coordinates)
__device__ __forceinline__ Stencil cell_stencil(const DevParams& P, int cx, int cy) {
    Stencil st;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int vx = cx - 1 + i, vy = cy - 1 + i;
        if (vx < 0) vx += P.nbx;
        if (vx >= P.nbx) vx -= P.nbx;
        if (vy < 0) vy += P.nby;
        if (vy >= P.nby) vy -= P.nby;
        st.xs[i] = vx;
        st.ys[i] = vy;
    }
    st.nxs = 3;
    st.xs[3] = -1;
    if (cx == 0) { st.xs[3] = P.nbx - 2; st.nxs = 4; }
    else if (cx == P.nbx - 2) { st.xs[3] = 0; st.nxs = 4; }
    return st;
}

// partner ranges of (home cell c, part): up to three (begin, end) slot ranges, in stencil order.
// Adatom parts keep only cells c2 >= c (every unordered pair once, with the stencil's multiplicity).
struct Ranges {
    int b[3], e[3];
    int self[3];      // 1 when the range is the home cell itself (pair only if slot_i < slot_j)
    int total;
};
__device__ __forceinline__ Ranges partner_ranges(const DevParams& P, const int* ad_start, const int* sub_start, int c, int part) {
    Ranges r;
    r.total = 0;
    const int cx = c / P.nby, cy = c % P.nby;
    const Stencil st = cell_stencil(P, cx, cy);
    const int a = part & 3;
    const bool is_sub = part >= 4;
#pragma unroll
    for (int k = 0; k < 3; ++k) { r.b[k] = r.e[k] = 0; r.self[k] = 0; }
    if (a >= st.nxs) return r;
    const int nx = st.xs[a];
    if (nx < 0 || nx >= P.nbx) return r;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int ny = st.ys[k];
        if (ny < 0 || ny >= P.nby) continue;
        const int c2 = nx * P.nby + ny;
        if (!is_sub && c2 < c) continue;
        const int* s = is_sub ? sub_start : ad_start;
        r.b[k] = s[c2];
        r.e[k] = s[c2 + 1];
        r.self[k] = (!is_sub && c2 == c) ? 1 : 0;
        r.total += r.e[k] - r.b[k];
    }
    return r;
}

// In-place inclusive scan of a[1..m] over the WHOLE grid (a[0] is left alone), optional mirror copy: every block scans its
// tile and publishes the tile total, one grid barrier, every block adds the totals of the tiles before it.  (The one-block
// scan of the small-grid path walks 1.6 M cells with 512 threads on configs[4]: milliseconds per rebuild.)  A real function
// (rare path, plain arguments): it must not bloat the persistent kernel, whose line-search body has to stay inlined.
// `epoch` is the caller's barrier epoch: the function executes ONE grid barrier, the caller increments its counter.
// Returns true when the watchdog fired.
__device__ __noinline__ bool relax2_scan_tall(int* a, int m, int* mirror, int* total_out, int* tile_tot, int* s_scan,
                                              unsigned int* bar, int* abortw, unsigned int epoch) {
    const int T = blockDim.x, t = threadIdx.x, nb = gridDim.x;
    const int tile = (m + nb - 1) / nb;
    const int tlo = 1 + blockIdx.x * tile, thi = min(tlo + tile, m + 1);
    const int per = (tile + T - 1) / T;
    const int lo = tlo + t * per, hi = min(lo + per, thi);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += a[i];
    s_scan[t] = sum;
    __syncthreads();
    for (int off = 1; off < T; off <<= 1) {
        int v = (t >= off) ? s_scan[t - off] : 0;
        __syncthreads();
        s_scan[t] += v;
        __syncthreads();
    }
    const int before = s_scan[t] - sum;
    if (t == T - 1) tile_tot[blockIdx.x] = s_scan[t];
    bool dead = false;
    __syncthreads();
    if (t == 0) {
        const unsigned int target = (epoch + 1u) * gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
        unsigned int v;
        unsigned long long t0 = 0ull;
        unsigned int spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((++spins & 1023u) == 0) {
                if (*(volatile int*)abortw != 0) { dead = true; break; }
                const unsigned long long now = gtimer();
                if (t0 == 0ull) t0 = now;
                else if (now - t0 > R2_SPIN_NS) { atomicCAS(abortw, 0, 1); dead = true; break; }
            }
        } while ((int)(v - target) < 0);
    }
    if (__syncthreads_or(dead)) return true;
    int offset = 0;
    for (int b = t; b < (int)blockIdx.x; b += T) offset += tile_tot[b];
    s_scan[t] = offset;
    __syncthreads();
    for (int off = T >> 1; off > 0; off >>= 1) {
        if (t < off) s_scan[t] += s_scan[t + off];
        __syncthreads();
    }
    int run = s_scan[0] + before;
    __syncthreads();
    for (int i = lo; i < hi; ++i) { run += a[i]; a[i] = run; if (mirror) mirror[i] = run; }
    if (mirror && blockIdx.x == 0 && t == 0) mirror[0] = a[0];
    if (total_out && blockIdx.x == nb - 1 && t == 0) {
        int tot = 0;
        for (int b = 0; b < nb; ++b) tot += tile_tot[b];
        *total_out = tot;
    }
    return false;
}

// ---- rebuild of the sorted state from `n` source elements (7 barriers; 9 on tall grids) ----------------------
// src_s / src_oidx may be null (direction 0 / identity index).
// TALL is a compile-time property of the launch (kmc_relax2.cu picks the instantiation): the common kernel carries none of
// the tall-grid code (it cost 1 % of the headline when it was a run-time branch).
template <bool TALL>
__device__ __forceinline__ bool relax2_rebuild(const Relax2Args& R, const double2* src_pos, const double2* src_s, const int* src_oidx,
                                               double2* dst_pos, double2* dst_s, int* dst_oidx, unsigned int& epoch, Relax2Smem& S, int nA) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const DevParams& P = R.P;
    constexpr bool tall = TALL;
    for (int c = gtid; c < R.ncell2; c += gthreads) R.start[c] = 0;
    if (tall) for (int c = gtid; c < P.nbx; c += gthreads) R.colr[c] = make_int2(0x7fffffff, -1);
    if (relax2_barrier(R, epoch)) return true;
    for (int k = gtid; k < nA; k += gthreads) {
        int nx, ny;
        bin_index(P, src_pos[k].x, src_pos[k].y, nx, ny);
        const int c = flat_cell(P, nx, ny);
        R.key[k] = c;
        atomicAdd(&R.start[1 + c], 1);
        if (tall && c < P.ncell) { atomicMin(&R.colr[nx].x, ny); atomicMax(&R.colr[nx].y, ny); }
    }
    if (relax2_barrier(R, epoch)) return true;
    if (tall) {
        if (relax2_scan_tall(R.start, R.ncell2 - 1, R.cursor, nullptr, R.tile_tot, S.u.e.scan, R.barrier, R.abort, epoch)) return true;
        ++epoch;
    } else if (blockIdx.x == 0) {
        const int T = blockDim.x, t = threadIdx.x;
        const int m = R.ncell2 - 1;
        const int per = (m + T - 1) / T;
        const int lo = 1 + t * per, hi = min(lo + per, R.ncell2);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += R.start[i];
        S.u.e.scan[t] = sum;
        __syncthreads();
        for (int off = 1; off < T; off <<= 1) {
            int v = (t >= off) ? S.u.e.scan[t - off] : 0;
            __syncthreads();
            S.u.e.scan[t] += v;
            __syncthreads();
        }
        int run = S.u.e.scan[t] - sum;
        for (int i = lo; i < hi; ++i) { run += R.start[i]; R.start[i] = run; }
        __syncthreads();
        for (int i = t; i < R.ncell2; i += T) R.cursor[i] = R.start[i];
    }
    if (relax2_barrier(R, epoch)) return true;
    for (int k = gtid; k < nA; k += gthreads) {
        const int slot = atomicAdd(&R.cursor[R.key[k]], 1);
        R.perm[slot] = k;
    }
    if (relax2_barrier(R, epoch)) return true;
    // per cell: order by original index (deterministic), move the payload, count the work items.
    // One WARP per cell: cells of up to 32 atoms are ranked with shuffles (three dependent fetches in all); the
    // rest of the per-cell work below is done by lane 0.
    const int gwarp_ = gtid >> 5, nwarps_ = gthreads >> 5, lane_ = threadIdx.x & 31;
    // cells visited by this warp: every nwarps-th cell of the grid, or -- tall grid -- the occupied rows of every nwarps-th
    // column plus the overflow cell (the counters of the many empty cells are zeroed by a streaming pass: every cell has
    // exactly one writer).  ONE instance of the per-cell body.
    if (tall)
        for (int c = gtid; c < P.ncell; c += gthreads)
            if (R.start[c + 1] == R.start[c]) R.cursor[1 + c] = 0;
    int it_col = gwarp_, it_row = 0, it_hi = -1, it_c = gwarp_ - nwarps_;
    bool it_over = false;
    for (;;) {
        int c;
        if (!tall) {
            it_c += nwarps_;
            if (it_c >= R.ncell2 - 1) break;
            c = it_c;
        } else if (it_row <= it_hi) {
            c = (it_col - nwarps_) * P.nby + it_row;
            ++it_row;
        } else {
            bool found = false;
            for (; it_col < P.nbx; it_col += nwarps_) {
                const int2 r = R.colr[it_col];
                if (r.x <= r.y) { it_row = r.x; it_hi = r.y; found = true; break; }
            }
            if (found) {
                c = it_col * P.nby + it_row;
                ++it_row;
                it_col += nwarps_;
            } else {
                if (it_over || gwarp_ != 0) break;
                it_over = true;
                c = P.ncell;                  // the overflow cell (atoms outside the grid)
            }
        }
        const int lo = R.start[c], hi = R.start[c + 1];
        const int m = hi - lo;
        if (m == 0) {               // empty cell: no work items (the counters still have to be defined)
            if (lane_ < R2_PARTS && c < P.ncell) R.chunks[c * R2_PARTS + lane_] = 0;
            if (lane_ == 0 && c < P.ncell) R.cursor[1 + c] = 0;
            continue;
        }
        if (m <= 32) {
            int k = 0, key = 0x7fffffff;
            if (lane_ < m) { k = R.perm[lo + lane_]; key = src_oidx ? src_oidx[k] : k; }
            int rank = 0;
            for (int j = 0; j < m; ++j) rank += (__shfl_sync(0xffffffffu, key, j) < key);
            if (lane_ < m) {
                const int s = lo + rank;
                dst_pos[s] = src_pos[k];
                dst_s[s] = src_s ? src_s[k] : make_double2(0.0, 0.0);
                dst_oidx[s] = key;
                R.cellof[s] = c;
            }
        } else if (lane_ == 0) {
            for (int i = lo + 1; i < hi; ++i) {          // insertion sort of perm[lo..hi) keyed on the original index
                const int v = R.perm[i];
                const int kv = src_oidx ? src_oidx[v] : v;
                int j = i - 1;
                while (j >= lo) {
                    const int w = R.perm[j];
                    if ((src_oidx ? src_oidx[w] : w) <= kv) break;
                    R.perm[j + 1] = w;
                    --j;
                }
                R.perm[j + 1] = v;
            }
            for (int s = lo; s < hi; ++s) {
                const int k = R.perm[s];
                dst_pos[s] = src_pos[k];
                dst_s[s] = src_s ? src_s[k] : make_double2(0.0, 0.0);
                dst_oidx[s] = src_oidx ? src_oidx[k] : k;
                R.cellof[s] = c;
            }
        }
        if (lane_ != 0) continue;
        if (hi > lo) {      // stencil ranges of this cell's atoms: one dependent fetch less in every force sweep
            int4 ab = make_int4(-1, 0, 0, 0), ae = make_int4(0, 0, 0, 0), sb = ab, se = ae;
            if (c < P.ncell) {
                const Stencil st = cell_stencil(P, c / P.nby, c % P.nby);
                const bool ycontig = (st.ys[0] >= 0) && (st.ys[2] < P.nby) && (st.ys[1] == st.ys[0] + 1) && (st.ys[2] == st.ys[1] + 1);
                if (ycontig) {
                    int b1[4], e1[4], b2[4], e2[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int cx = st.xs[a];
                        const bool ok = a < st.nxs && cx >= 0 && cx < P.nbx;
                        b1[a] = ok ? R.start[cx * P.nby + st.ys[0]] : 0;
                        e1[a] = ok ? R.start[cx * P.nby + st.ys[2] + 1] : 0;
                        b2[a] = ok ? R.sub.start[cx * P.nby + st.ys[0]] : 0;
                        e2[a] = ok ? R.sub.start[cx * P.nby + st.ys[2] + 1] : 0;
                    }
                    ab = make_int4(b1[0], b1[1], b1[2], b1[3]); ae = make_int4(e1[0], e1[1], e1[2], e1[3]);
                    sb = make_int4(b2[0], b2[1], b2[2], b2[3]); se = make_int4(e2[0], e2[1], e2[2], e2[3]);
                }
            }
            for (int s2 = lo; s2 < hi; ++s2) {
                R.nbr[4 * s2] = ab; R.nbr[4 * s2 + 1] = ae; R.nbr[4 * s2 + 2] = sb; R.nbr[4 * s2 + 3] = se;
            }
        }
        if (c < P.ncell) {
            const int nh = hi - lo;
            int tot = 0;
#pragma unroll
            for (int part = 0; part < R2_PARTS; ++part) {
                int cnt = 0;
                if (nh > 0 && !(part < 4 && P.aa_mode == 0)) {
                    const Ranges r = partner_ranges(P, R.start, R.sub.start, c, part);
                    cnt = (nh * r.total + R2_CHUNK - 1) / R2_CHUNK;
                }
                R.chunks[c * R2_PARTS + part] = cnt;
                tot += cnt;
            }
            R.cursor[1 + c] = tot;          // items of this home cell (the scatter cursor is free again)
        }
        }
    if (gtid == 0) R.cursor[0] = 0;
    if (relax2_barrier(R, epoch)) return true;
    if (tall) {
        if (relax2_scan_tall(R.cursor, P.ncell, nullptr, R.n_items, R.tile_tot, S.u.e.scan, R.barrier, R.abort, epoch)) return true;
        ++epoch;
    } else if (blockIdx.x == 0) {      // inclusive scan of the per-cell item counts cursor[1..ncell]
        const int T = blockDim.x, t = threadIdx.x;
        const int m = P.ncell;
        const int per = (m + T - 1) / T;
        const int lo = 1 + t * per, hi = min(lo + per, m + 1);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += R.cursor[i];
        S.u.e.scan[t] = sum;
        __syncthreads();
        for (int off = 1; off < T; off <<= 1) {
            int v = (t >= off) ? S.u.e.scan[t - off] : 0;
            __syncthreads();
            S.u.e.scan[t] += v;
            __syncthreads();
        }
        int run = S.u.e.scan[t] - sum;
        for (int i = lo; i < hi; ++i) { run += R.cursor[i]; R.cursor[i] = run; }
        if (t == T - 1) *R.n_items = S.u.e.scan[t];      // true total; the caller checks it against item_cap
    }
    if (relax2_barrier(R, epoch)) return true;
    for (int c = gtid; c < P.ncell; c += gthreads) {
        int k = R.cursor[c];
        const int kend = R.cursor[c + 1];
        if (k == kend) continue;
        const int nh = R.start[c + 1] - R.start[c];
        for (int part = 0; part < R2_PARTS; ++part) {
            const int cnt = R.chunks[c * R2_PARTS + part];
            if (cnt == 0) continue;
            const Ranges r = partner_ranges(P, R.start, R.sub.start, c, part);
            const int pairs = nh * r.total;
            WorkItem it;
            it.hb = R.start[c]; it.nh = nh;
            it.b0 = r.b[0]; it.n0 = r.e[0] - r.b[0]; it.b1 = r.b[1]; it.n1 = r.e[1] - r.b[1]; it.b2 = r.b[2];
            it.kind = r.self[0] | (r.self[1] << 1) | (r.self[2] << 2) | (part >= 4 ? 8 : 0);
            it.inv_nh = 1.0f / (float)nh;
            it.n2 = r.e[2] - r.b[2];
            for (int j = 0; j < cnt; ++j, ++k) {
                if (k >= R.item_cap) break;
                it.begin = j * R2_CHUNK;
                it.count = min(R2_CHUNK, pairs - it.begin);
                R.items[k] = it;
            }
        }
    }
    if (relax2_barrier(R, epoch)) return true;
    return false;
}

// ---- E: work items --------------------------------------------------------------------------------
// One pair per lane and round.  Quiet pairs add their Taylor coefficients to acc (pair_line_poly).  Pairs that cannot be
// expanded are evaluated per candidate by the WARP: each one is broadcast and lane a (< KC) adds candidate a to its private
// sum `eslow` (fixed order: deterministic; one register instead of KC accumulators per lane).
__device__ __forceinline__ void relax2_slow_pairs(const DevParams& P, bool slow, double dx0, double dy0, double ddx, double ddy, bool is_sub,
                                                  int lane, double& eslow) {
    unsigned int m = __ballot_sync(0xffffffffu, slow);
    if (m) {
        const double E = is_sub ? P.E_as : P.E_a;
        const double r0sq = is_sub ? P.ras2 : P.ra2;
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const double bx = __shfl_sync(0xffffffffu, dx0, src), by = __shfl_sync(0xffffffffu, dy0, src);
            const double vx = __shfl_sync(0xffffffffu, ddx, src), vy = __shfl_sync(0xffffffffu, ddy, src);
            if (lane < KC) eslow += pair_line_one(P, bx, by, vx, vy, E, r0sq, lane);
        }
    }
}
__device__ __forceinline__ void relax2_pair_round(const DevParams& P, bool valid, double dx0, double dy0, double ddx, double ddy, bool is_sub,
                                                  int lane, double (&acc)[PD + 1], double& eslow) {
    const bool ok = pair_line_poly(P, valid, dx0, dy0, ddx, ddy, is_sub ? P.E_as : P.E_a, is_sub ? P.inv_ras2 : P.inv_ra2, acc);
    relax2_slow_pairs(P, valid && !ok, dx0, dy0, ddx, ddy, is_sub, lane, eslow);
}

// operands of one round of a work item: every load is issued before any of them is used (one L2 round trip), and the
// caller keeps two rounds in flight (after a grid barrier every access is an L2 hit of ~0.45 us; items come from shared memory)
struct PairLoad {
    double2 pi, si, pj, sj;
    int hs, js;
    unsigned int bits;      // 0: in range, 1: same-cell rule applies, 2: substrate partner, 8..: flag[hs], 16..: flag[js]
};
struct EOperands {          // where the energy phase reads from: staged copies (rebased to slot indices) or global memory
    const double2* pos;
    const double2* dir;
    const uint8_t* flag;
    const double2* sub;
};
__device__ __forceinline__ PairLoad relax2_round_load(const EOperands& O, const WorkItem& it, int q0, int lane) {
    PairLoad L;
    const bool is_sub = (it.kind & 8) != 0;
    const int end = it.begin + it.count;
    const int q = min(q0 + lane, end - 1);
    int j = (int)((float)q * it.inv_nh);                    // q / nh with one fix-up step (pair spaces stay far below 2^23)
    int i = q - j * it.nh;
    if (i < 0) { --j; i += it.nh; } else if (i >= it.nh) { ++j; i -= it.nh; }
    const int hs = it.hb + i;
    int js, self;
    if (j < it.n0) { js = it.b0 + j; self = it.kind & 1; }
    else if (j < it.n0 + it.n1) { js = it.b1 + (j - it.n0); self = it.kind & 2; }
    else { js = it.b2 + (j - it.n0 - it.n1); self = it.kind & 4; }
    L.hs = hs; L.js = js;
    const unsigned int fh = O.flag[hs];
    const unsigned int fj = is_sub ? 0u : (unsigned int)O.flag[js];
    L.pi = O.pos[hs];
    L.si = O.dir[hs];
    L.pj = is_sub ? O.sub[js] : O.pos[js];
    L.sj = is_sub ? make_double2(0.0, 0.0) : O.dir[js];
    L.bits = ((q0 + lane < end) ? 1u : 0u) | (self ? 2u : 0u) | (is_sub ? 4u : 0u) | (fh << 8) | (fj << 16);
    return L;
}
struct PairGeo { double dx0, dy0, ddx, ddy; bool valid, is_sub; };
__device__ __forceinline__ PairGeo relax2_round_geo(const DevParams& P, const PairLoad& L, double inv_step) {
    PairGeo g;
    g.is_sub = (L.bits & 4u) != 0;       // warp uniform (a property of the item)
    // movers are handled by the fix-up phase; pairs inside the home cell count once (slot_i < slot_j)
    g.valid = (L.bits & 1u) && !(L.bits >> 8) && !((L.bits & 2u) && L.js <= L.hs);
    g.dx0 = put_in_box(L.pj.x - L.pi.x, P.L, P.halfL);
    g.dy0 = L.pj.y - L.pi.y;
    g.ddx = (L.sj.x - L.si.x) * inv_step;
    g.ddy = (L.sj.y - L.si.y) * inv_step;
    return g;
}
__device__ __forceinline__ void relax2_round_compute(const Relax2Args& R, const PairLoad& L, double inv_step, int lane, double (&acc)[PD + 1], double& eslow) {
    const DevParams& P = R.P;
    const PairGeo g = relax2_round_geo(P, L, inv_step);
    relax2_pair_round(P, g.valid, g.dx0, g.dy0, g.ddx, g.ddy, g.is_sub, lane, acc, eslow);
}
// two rounds at once: their polynomial chains are independent and interleave
__device__ __forceinline__ void relax2_round_compute2(const Relax2Args& R, const PairLoad& LA, const PairLoad& LB, double inv_step, int lane,
                                                      double (&acc)[PD + 1], double& eslow) {
    const DevParams& P = R.P;
    const PairGeo ga = relax2_round_geo(P, LA, inv_step), gb = relax2_round_geo(P, LB, inv_step);
    const PolyPrep ra = pair_line_prep(P, ga.valid, ga.dx0, ga.dy0, ga.ddx, ga.ddy, ga.is_sub ? P.E_as : P.E_a, ga.is_sub ? P.inv_ras2 : P.inv_ra2);
    const PolyPrep rb = pair_line_prep(P, gb.valid, gb.dx0, gb.dy0, gb.ddx, gb.ddy, gb.is_sub ? P.E_as : P.E_a, gb.is_sub ? P.inv_ras2 : P.inv_ra2);
    if (__all_sync(0xffffffffu, (!ra.ok || ra.small) && (!rb.ok || rb.small))) {       // warp-uniform choice of the degree
        pair_line_acc<3>(ra, acc);
        pair_line_acc<3>(rb, acc);
    } else {
        pair_line_acc<PD>(ra, acc);
        pair_line_acc<PD>(rb, acc);
    }
    relax2_slow_pairs(P, ga.valid && !ra.ok, ga.dx0, ga.dy0, ga.ddx, ga.ddy, ga.is_sub, lane, eslow);
    relax2_slow_pairs(P, gb.valid && !rb.ok, gb.dx0, gb.dy0, gb.ddx, gb.ddy, gb.is_sub, lane, eslow);
}

// all items k = w0, w0 + nw, ... of this warp; `wit` = the warp's cached copies of the first R2_WI of them
// items k = kfirst, kfirst + kstride, ... < kend of this warp; `wit` = the warp's cached copies of the first R2_WI of them
__device__ __forceinline__ void relax2_items(const Relax2Args& R, const WorkItem* wit, int kfirst, int kstride, int kend, const EOperands& O,
                                             double inv_step, int lane, double (&acc)[PD + 1], double& eslow) {
    if (kfirst >= kend) return;
    int idx = 0;                            // ordinal of the current item of this warp
    WorkItem it = wit[0];
    int q0 = it.begin;
    // advance (it, q0) to the next round of this warp; false when there is none
    auto advance = [&]() -> bool {
        q0 += 32;
        if (q0 >= it.begin + it.count) {
            ++idx;
            const int k = kfirst + idx * kstride;
            if (k >= kend) return false;
            it = (idx < R2_WI) ? wit[idx] : R.items[k];
            q0 = it.begin;
        }
        return true;
    };
    // two rounds per step: both sets of operands are requested together and their polynomial chains interleave
    bool more = true;
    while (more) {
        const PairLoad a = relax2_round_load(O, it, q0, lane);
        more = advance();
        if (more) {
            const PairLoad b = relax2_round_load(O, it, q0, lane);
            more = advance();
            relax2_round_compute2(R, a, b, inv_step, lane, acc, eslow);
        } else {
            relax2_round_compute(R, a, inv_step, lane, acc, eslow);
        }
    }
}

// ---- F: one (mover, candidate) by one warp; sorted-state version of line_fix_one ------------------
__device__ __forceinline__ double relax2_fix_one(const Relax2Args& R, const double2* pos, const double2* sdir, const int* oidx,
                                                 double scale, int r, int a, int lane, int nA) {
    const DevParams& P = R.P;
    const int nm = *R.n_movers;
    const int m = R.movers[r];
    const int om = oidx[m];
    const double2 pm = cand2(pos[m], sdir[m], a, scale);
    int rx, ry;
    bin_index(P, pm.x, pm.y, rx, ry);
    const bool in_grid = rx >= 0 && rx < P.nbx && ry >= 0 && ry < P.nby;
    const Stencil st = make_stencil(P, pm.x, pm.y);
    CellView ad;
    ad.start = R.start; ad.sxy = pos; ad.sidx = oidx; ad.n = nA;
    double e = 0.0;
    if (P.aa_mode != 0) {
        // the partners' candidate positions only enter the pair distance: formed with one FMA (a few 1e-16 relative
        // from the reference's (a*s)/20/scale); the mover's own position decides its cell and stays exact
        const double ak = (double)a / (20.0 * scale);
        visit_stencil(P, ad, st, lane, 32, [&](int slot) {
            if (R.flag[slot] || (!in_grid && oidx[slot] < om)) return;
            const double2 p0 = pos[slot], s0 = sdir[slot];
            const double2 pj = make_double2(fma(s0.x, ak, p0.x), fma(s0.y, ak, p0.y));
            e += pair_energy(P, pm.x, pm.y, pj, P.E_a, P.ra2);
        });
        for (int r2 = lane; r2 < nm; r2 += 32) {               // mover - mover: the lower index sees the higher one
            const int m2 = R.movers[r2];
            if (oidx[m2] <= om) continue;
            const double2 p2 = cand2(pos[m2], sdir[m2], a, scale);
            int cx, cy;
            bin_index(P, p2.x, p2.y, cx, cy);
            if (cx < 0 || cx >= P.nbx || cy < 0 || cy >= P.nby) continue;
            const int mult = stencil_multiplicity(P, st, cx, cy);
            if (mult) e += mult * pair_energy(P, pm.x, pm.y, p2, P.E_a, P.ra2);
        }
    }
    visit_stencil(P, R.sub, st, lane, 32, [&](int slot) { e += pair_energy(P, pm.x, pm.y, R.sub.sxy[slot], P.E_as, P.ras2); });
    return group_sum<32>(e);
}

// mover test of the next line: 0 only when both ends of the line (a = 0 and a = KC-1; coordinates are monotone in a)
// lie in the same in-grid cell.  Conservative and division-free: cell coordinates are formed with a reciprocal
// multiply, and an end closer than 1e-9 cells to a cell wall (or in a negative cell coordinate, where the reference's
// int() truncates toward zero) makes the atom a mover -- movers are evaluated exactly, per candidate, by the fix-up
// phase, so a false positive costs time, never accuracy.
__device__ __forceinline__ uint8_t mover_flag(const DevParams& P, double2 p0, double2 s, double scale) {
    const double inv_bin = 1.0 / P.bin_size;            // loop invariant: hoisted by the compiler
    const double k = (double)(KC - 1) / (20.0 * scale) * inv_bin;
    const double ux0 = (p0.x + P.halfL) * inv_bin, uy0 = (p0.y - P.Dmin) * inv_bin;
    const double ux1 = fma(s.x, k, ux0), uy1 = fma(s.y, k, uy0);
    const double fx = floor(ux0), fy = floor(uy0);
    const double m = 1e-9;
    const bool same = fmin(ux0, ux1) > fx + m && fmax(ux0, ux1) < fx + 1.0 - m && fmin(uy0, uy1) > fy + m && fmax(uy0, uy1) < fy + 1.0 - m;
    const bool in0 = fx >= 0.0 && fx < (double)P.nbx && fy >= 0.0 && fy < (double)P.nby;
    return (same && in0) ? 0 : 1;
}

// ---- G: forces at `pos` (+ optional fused CG update / mover flags) --------------------------------
// mode 0: sdir <- F (the first direction dx0, 2DGrowth.py:120,125), count pairs when asked
// mode 1: max|F|^2 partial, sdir <- F + beta*sdir (2DGrowth.py:147), flag <- mover test of the next line
// forces of the stencil atoms of list `cv` on (xi, yi): same visiting order as visit_stencil, but four
// neighbour loads are issued before their arithmetic (after a grid barrier every first touch is an L2 hit,
// and one load in flight per lane made this phase latency bound)
__device__ __forceinline__ void stencil_forces(const DevParams& P, const CellView& cv, const Stencil& st, int lane, int G, double xi, double yi,
                                               double E, double r0sq, double& fx, double& fy, unsigned long long& visits) {
    const bool ycontig = (st.ys[0] >= 0) && (st.ys[2] < P.nby) && (st.ys[1] == st.ys[0] + 1) && (st.ys[2] == st.ys[1] + 1);
    for (int a = 0; a < st.nxs; ++a) {
        const int cx = st.xs[a];
        if (cx < 0 || cx >= P.nbx) continue;
        const int base = cx * P.nby;
        const int nseg = ycontig ? 1 : 3;
        for (int c = 0; c < nseg; ++c) {
            int b, e;
            if (ycontig) { b = cv.start[base + st.ys[0]]; e = cv.start[base + st.ys[2] + 1]; }
            else {
                const int cy = st.ys[c];
                if (cy < 0 || cy >= P.nby) continue;
                b = cv.start[base + cy]; e = cv.start[base + cy + 1];
            }
            for (int s = b + lane; s < e; s += 4 * G) {
                const int s1 = s + G, s2 = s + 2 * G, s3 = s + 3 * G;
                const double2 p0 = cv.sxy[s];
                const double2 p1 = cv.sxy[min(s1, e - 1)];
                const double2 p2 = cv.sxy[min(s2, e - 1)];
                const double2 p3 = cv.sxy[min(s3, e - 1)];
                acc_force(P, xi, yi, p0, E, r0sq, fx, fy);
                if (s1 < e) acc_force(P, xi, yi, p1, E, r0sq, fx, fy);
                if (s2 < e) acc_force(P, xi, yi, p2, E, r0sq, fx, fy);
                if (s3 < e) acc_force(P, xi, yi, p3, E, r0sq, fx, fy);
                visits += 1 + (s1 < e) + (s2 < e) + (s3 < e);
            }
        }
    }
}

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) of a contiguous slot range into shared memory ----------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned int bytes, unsigned long long* bar) {
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// three copies completing on one mbarrier phase (one arrive.expect_tx for their total size)
__device__ __forceinline__ void bulk_load3(void* d0, const void* g0, unsigned int n0, void* d1, const void* g1, unsigned int n1,
                                           void* d2, const void* g2, unsigned int n2, unsigned long long* bar) {
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(n0 + n1 + n2) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(d0)), "l"(g0), "r"(n0), "r"(smem_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(d1)), "l"(g1), "r"(n1), "r"(smem_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(d2)), "l"(g2), "r"(n2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned int parity) {
    unsigned int ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(const Relax2Args& R, unsigned long long* bar, unsigned int parity) {
    unsigned long long t0 = 0ull;
    unsigned int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0 && relax2_spin_expired(R, t0, 3)) break;
    }
}

// forces of up to four contiguous slot ranges (the stencil columns, in stencil order) on (xi, yi): the ranges are
// walked as ONE flat index space and eight neighbour positions are fetched before their arithmetic, so a sweep
// costs one or two L2 round trips instead of one per column (after a grid barrier every first touch misses L1;
// a dependent L2 access measures ~0.45 us here)
__device__ __forceinline__ void flat_forces(const DevParams& P, const double2* ptr, int4 b, int4 e, int lane, int G, double xi, double yi,
                                            double E, double r0sq, double& fx, double& fy, unsigned long long& visits) {
    const int c0 = e.x - b.x, c1 = c0 + (e.y - b.y), c2 = c1 + (e.z - b.z), tot = c2 + (e.w - b.w);
    const double E12 = 12.0 * E;
    for (int t = lane; t < tot; t += R2_FB * G) {
        double2 p[R2_FB];
        bool far = false;
#pragma unroll
        for (int k = 0; k < R2_FB; ++k) {
            const int tk = min(t + k * G, tot - 1);
            const int sl = tk < c0 ? b.x + tk : (tk < c1 ? b.y + (tk - c0) : (tk < c2 ? b.z + (tk - c1) : b.w + (tk - c2)));
            p[k] = ptr[sl];
            far = far || !(fabs(p[k].x - xi) < P.L);
        }
        if (far || P.precision == 1) {      // |dx| >= L (atoms that drifted by a box length) or fp32 pair terms: the general form
#pragma unroll
            for (int k = 0; k < R2_FB; ++k)
                if (t + k * G < tot) { acc_force(P, xi, yi, p[k], E, r0sq, fx, fy); ++visits; }
        } else {
            // branch-free: the eight pair chains are independent and interleave (one basic block); a slot past the end
            // holds a clamped duplicate and contributes a zero coefficient.  Same arithmetic as acc_force.
#pragma unroll
            for (int k = 0; k < R2_FB; ++k) {
                double dx = p[k].x - xi;
                dx = (dx < 0.0) ? __dadd_rn(dx, P.L) : dx;
                dx = (dx > P.halfL) ? __dadd_rn(dx, -P.L) : dx;
                const double dy = p[k].y - yi;
                const double r2 = dx * dx + dy * dy;
                const double inv = fast_rcp(r2);
                const double q2 = r0sq * inv;
                const double q6 = q2 * q2 * q2;
                const bool on = (t + k * G < tot) && !(r2 < 1e-4);       // r < 1e-2 skipped (Energy.py:55)
                const double c = on ? E12 * (q6 - q6 * q6) * inv : 0.0;
                fx += dx * c;
                fy += dy * c;
                visits += (t + k * G < tot) ? 1 : 0;
            }
        }
    }
}

// what the energy phase of a block reads: its contiguous range of work items and, when they fit, the staged operands
struct ERange {
    int kb, ke;             // work items [kb, ke) of this block
    int emin, emax;         // staged slots (emax == emin: not staged, read global memory)
    int submin, submax;     // cached substrate atoms (submax == submin: read global memory)
};
__device__ __forceinline__ void relax2_item_range(const Relax2Args& R, int n_items, int& kb, int& ke) {
    if (gridDim.x == 1) { kb = 0; ke = n_items; return; }
    if (blockIdx.x == 0) { kb = ke = 0; return; }          // block 0 builds the mover list instead
    const long long nb = gridDim.x - 1, b = blockIdx.x - 1;
    kb = (int)(b * n_items / nb);
    ke = (int)((b + 1) * n_items / nb);
}

// Slot range covered by the stencils of this block's chunk of atoms.  The atoms are cell-sorted column-major and only
// a few cells of a column are occupied, so the stencils of a contiguous chunk cover ONE contiguous slot range (three
// adjacent columns) except at the periodic seam.  Valid until the next rebuild.
__device__ __forceinline__ void relax2_stage_range(const Relax2Args& R, Relax2Smem& S, int& smin, int& smax, ERange& er, int nA) {
    const int T = blockDim.x;
    {   // the block's work items: a contiguous range of the (home-cell ordered) list, dealt to its warps round robin; the
        // first R2_WI of every warp are cached.  Contiguous items = adjacent home cells = ONE contiguous slot range of
        // operands (the same argument as for the force phase), which is what makes the staging below possible.
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
        const int n_items = min(*R.n_items, R.item_cap);
        relax2_item_range(R, n_items, er.kb, er.ke);
        if (lane < R2_WI) {
            const int k = er.kb + warp + lane * wpb;
            if (k < er.ke) S.wit[warp][lane] = R.items[k];
        }
        if (threadIdx.x == 0) { S.emin = 0x7fffffff; S.emax = -1; S.submin = 0x7fffffff; S.submax = -1; }
        __syncthreads();
        int mn = 0x7fffffff, mx = -1, smn = 0x7fffffff, smx = -1;
        for (int k = er.kb + threadIdx.x; k < er.ke; k += T) {
            const WorkItem it = R.items[k];
            mn = min(mn, it.hb); mx = max(mx, it.hb + it.nh);
            int& lo2 = (it.kind & 8) ? smn : mn;
            int& hi2 = (it.kind & 8) ? smx : mx;
            if (it.n0 > 0) { lo2 = min(lo2, it.b0); hi2 = max(hi2, it.b0 + it.n0); }
            if (it.n1 > 0) { lo2 = min(lo2, it.b1); hi2 = max(hi2, it.b1 + it.n1); }
            if (it.n2 > 0) { lo2 = min(lo2, it.b2); hi2 = max(hi2, it.b2 + it.n2); }
        }
        if (mx > mn) { atomicMin(&S.emin, mn); atomicMax(&S.emax, mx); }
        if (smx > smn) { atomicMin(&S.submin, smn); atomicMax(&S.submax, smx); }
        __syncthreads();
        er.emin = S.emin & ~15;               // 16-slot aligned: the flag copy needs a 16-byte aligned source
        er.emax = S.emax;
        if (!(er.emax > er.emin && er.emax - er.emin <= R2_STAGE) || R.debug == 6) er.emin = er.emax = 0;
        er.submin = S.submin; er.submax = S.submax;
        if (!(er.submax > er.submin && er.submax - er.submin <= R2_SUBSTAGE) || R.debug == 6) er.submin = er.submax = 0;
        for (int i = er.submin + threadIdx.x; i < er.submax; i += T) S.esub[i - er.submin] = R.sub.sxy[i];      // static: loaded once
    }
    {   // stencil ranges of this block's atoms
        const int per = (nA + gridDim.x - 1) / gridDim.x;
        const int lo = blockIdx.x * per, hi = min(lo + per, nA);
        for (int i = threadIdx.x; i < 4 * min(hi - lo, R2_NBRC); i += T) S.nbrc[i >> 2][i & 3] = R.nbr[4 * lo + i];
    }
    const int per = (nA + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(lo + per, nA);
    __syncthreads();
    if (threadIdx.x == 0) { S.smin = 0x7fffffff; S.smax = -1; }
    __syncthreads();
    if (R.P.aa_mode != 0) {
        for (int slot = lo + threadIdx.x; slot < hi; slot += T) {
            const int4 ab = R.nbr[4 * slot], ae = R.nbr[4 * slot + 1];
            if (ab.x < 0) continue;
            int mn = 0x7fffffff, mx = -1;
            if (ae.x > ab.x) { mn = min(mn, ab.x); mx = max(mx, ae.x); }
            if (ae.y > ab.y) { mn = min(mn, ab.y); mx = max(mx, ae.y); }
            if (ae.z > ab.z) { mn = min(mn, ab.z); mx = max(mx, ae.z); }
            if (ae.w > ab.w) { mn = min(mn, ab.w); mx = max(mx, ae.w); }
            if (mx > mn) { atomicMin(&S.smin, mn); atomicMax(&S.smax, mx); }
        }
    } else if (threadIdx.x == 0 && hi > lo) { S.smin = 0; S.smax = nA; }
    __syncthreads();
    smin = S.smin; smax = S.smax;
    if (!(smax > smin && (smax - smin) <= R2_STAGE)) { smin = 0; smax = 0; }
}

__device__ __forceinline__ double group_sum_rt(double v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Every block owns a contiguous chunk of slots (cell-sorted, so its atoms share one neighbourhood, kept in shared memory).
// The chunk is processed in ONE pass with T/atoms lanes per atom (see below); larger chunks take several passes.
template <bool TIMING>
__device__ __forceinline__ void relax2_forces(const Relax2Args& R, const double2* pos, double2* sdir, const int* oidx, int mode, double beta,
                                              double scale, bool count_pairs, Relax2Smem& S, unsigned int& parity, int smin, int smax,
                                              bool issue_here, int nA, unsigned long long* tm = nullptr) {
// (TIMING is a template parameter: the marks are volatile timer reads and act as scheduling fences; compiled into the
// production kernel -- even never executed -- they cost 4 % of the line search)
#define GMARK(k)                                                                          \
    if (TIMING && tm && threadIdx.x == 0 && blockIdx.x == R.phase_block) {                \
        const unsigned long long _t = gtimer();                                           \
        S.ph[k] += _t - *tm;      /* shared memory: a global read-modify-write per mark cost ~0.7 us */ \
        *tm = _t;                                                                         \
    }
    const DevParams& P = R.P;
    const int T = blockDim.x;
    const int per = (nA + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(lo + per, nA);
    CellView ad;
    ad.start = R.start; ad.sxy = pos; ad.sidx = oidx; ad.n = nA;
    double mf = 0.0;
    unsigned long long visits = 0;
    // The neighbourhood of this block's atoms [smin, smax) was determined after the last rebuild (relax2_stage_range);
    // its positions are fetched by ONE TMA bulk copy (issued here, or earlier by the caller so that the copy overlaps
    // the control arithmetic) and the sweep reads shared memory.
    const bool staged = smax > smin && (smax - smin) <= R2_STAGE && R.debug != 6;
    if (staged && issue_here) {          // not loaded ahead by the caller (first sweep of a pass, after a rebuild)
        __syncthreads();
        for (int i = threadIdx.x; i < smax - smin; i += T) S.stage[i] = pos[smin + i];
        __syncthreads();
    }
    const double2* npos = staged ? (S.stage - smin) : pos;       // neighbour positions by slot
    // ONE pass for the block's chunk: every atom gets LPA = T / atoms lanes (7 for 68 atoms on 512 threads; not a power of
    // two, so the lane partials are summed through shared memory, in lane order: deterministic).  A second, smaller
    // round for the atoms a power-of-two split leaves over cost a full latency chain on nearly idle SMs.
    const int cnt = hi - lo;
    const int LPA = max(1, min(32, T / max(cnt, 1)));
    const int app = T / LPA;                                  // atoms per pass
    for (int r0 = lo; r0 < hi; r0 += app) {
        const int G = LPA;
        const int grp = threadIdx.x / LPA;
        const int lane = threadIdx.x - grp * LPA;
        const int slot = r0 + grp;
        const bool live = grp < app && slot < hi;
        double ax = 0, ay = 0, sx = 0, sy = 0;
        double2 pi = make_double2(0, 0), so = make_double2(0, 0);
        int4 ab = make_int4(0, 0, 0, 0), ae = ab, sb = ab, se = ab;
        if (live) {
            if (lane == 0 && mode == 1) so = sdir[slot];       // requested now, used by the fused update at the end
            const int li = slot - lo;
            if (li < R2_NBRC) { ab = S.nbrc[li][0]; ae = S.nbrc[li][1]; sb = S.nbrc[li][2]; se = S.nbrc[li][3]; }
            else { ab = R.nbr[4 * slot]; ae = R.nbr[4 * slot + 1]; sb = R.nbr[4 * slot + 2]; se = R.nbr[4 * slot + 3]; }
        }
        GMARK(34)
        if (live) {
            pi = (staged && slot >= smin && slot < smax) ? npos[slot] : pos[slot];
            if (ab.x >= 0) {        // stencil ranges cached at the last rebuild
                if (P.aa_mode == 0) {
                    for (int s = lane; s < nA; s += G) acc_force(P, pi.x, pi.y, npos[s], P.E_a, P.ra2, ax, ay);
                } else {
                    flat_forces(P, npos, ab, ae, lane, G, pi.x, pi.y, P.E_a, P.ra2, ax, ay, visits);
                }
                GMARK(12)
                flat_forces(P, R.sub.sxy, sb, se, lane, G, pi.x, pi.y, P.E_as, P.ras2, sx, sy, visits);
                GMARK(13)
            } else {                // atoms outside the grid / y-wrapped stencils: generic walk from the raw position
                const Stencil st = make_stencil(P, pi.x, pi.y);
                if (P.aa_mode == 0) {
                    for (int s = lane; s < nA; s += G) acc_force(P, pi.x, pi.y, pos[s], P.E_a, P.ra2, ax, ay);
                } else {
                    stencil_forces(P, ad, st, lane, G, pi.x, pi.y, P.E_a, P.ra2, ax, ay, visits);
                }
                stencil_forces(P, R.sub, st, lane, G, pi.x, pi.y, P.E_as, P.ras2, sx, sy, visits);
            }
        }
        if (r0 != lo) __syncthreads();                          // the previous pass has read its partials
        S.u.g.fpa[threadIdx.x] = make_double2(ax, ay);
        S.u.g.fps[threadIdx.x] = make_double2(sx, sy);
        __syncthreads();
        if (live && lane == 0) {
            ax = ay = sx = sy = 0.0;
            for (int l = 0; l < LPA; ++l) {
                const double2 a2 = S.u.g.fpa[threadIdx.x + l], s2 = S.u.g.fps[threadIdx.x + l];
                ax += a2.x; ay += a2.y; sx += s2.x; sy += s2.y;
            }
        }
        GMARK(14)
        if (live && lane == 0) {
            const double fx = ax + sx, fy = ay + sy;                // 2DGrowth.py:120
            double2 sn;
            if (mode == 0) {
                sn = make_double2(fx, fy);
            } else {
                mf = fmax(mf, fx * fx + fy * fy);                   // :129,160
                sn = make_double2(__dadd_rn(fx, __dmul_rn(beta, so.x)), __dadd_rn(fy, __dmul_rn(beta, so.y)));   // :147
            }
            sdir[slot] = sn;
            R.flag[slot] = mover_flag(P, pi, sn, scale);
            if (TIMING && R.debug == 7) {
                const double d = sqrt(sn.x * sn.x + sn.y * sn.y) * (double)(KC - 1) / (20.0 * scale);
                int b = (int)floor(2.0 * (log10(fmax(d, 1e-30)) + 6.0));       // half decades from 1e-6
                b = max(0, min(15, b));
                atomicAdd(&R.hist[b], 1ull);
            }
        }
        GMARK(15)
    }
    if (mode == 1) {
        const double r_mf = block_max(mf, S.red);
        if (threadIdx.x == 0) R.red[(size_t)blockIdx.x * 4 + 3] = r_mf;
    }
    if (count_pairs) {
        for (int o = 16; o > 0; o >>= 1) visits += __shfl_xor_sync(0xffffffffu, visits, o);
        if ((threadIdx.x & 31) == 0 && visits) atomicAdd(R.pair_counter, visits);
    }
}

#define PHASE2_MARK(k)                                             \
    if (TIMING && threadIdx.x == 0 && blockIdx.x == R.phase_block) {                                     \
        const unsigned long long _t = gtimer();                    \
        S.ph[k] += _t - t_mark;                                    \
        t_mark = _t;                                               \
    }

template <bool TALL, bool TIMING>
static __global__ void __launch_bounds__(R2_THREADS, 1) k_relax2(Relax2Args R) {
    extern __shared__ __align__(16) unsigned char r2_smem_raw[];
    Relax2Smem& S = *reinterpret_cast<Relax2Smem*>(r2_smem_raw);
    if (R.status_dev && *R.status_dev != 0) return;            // the placement of this event failed: nothing to relax (uniform exit)
    const int nA = R.n_dev ? *R.n_dev : R.n;                   // adatoms of this relaxation
    if (nA <= 0) return;
    unsigned int epoch = 0;
    unsigned int tma_parity = 0;
    int st_min = 0, st_max = 0;      // staged slot range of this block (relax2_stage_range)
    int n_items_reg = 0;             // *R.n_items of the last rebuild
    ERange er{};                     // work items and staged operands of the energy phase (relax2_stage_range)
    unsigned int e_parity = 0;
    bool e_inflight = false;         // copies of the energy-phase operands requested and not yet waited for
    bool e_fresh = false;            // ... and they are the operands of the NEXT energy phase
    bool stall_check = false;        // the stall test applies from the second line search of a pass on
    bool layout_x0 = false;          // the sorted layout is the one built from x0 and sdir[cs ^ 1] holds F(x0)
    if (threadIdx.x == 0) { mbar_init(&S.mbar, 1); mbar_init(&S.mbar_e, 1); }
    if (TIMING && threadIdx.x < 48) S.ph[threadIdx.x] = 0ull;
    if (TIMING && R.debug == 8 && threadIdx.x == 0) {        // diagnostics: which SM runs this block
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        R.blk_ns[2 * gridDim.x + blockIdx.x] = smid;
    }
    __syncthreads();
    const DevParams& P = R.P;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gwarp = gtid >> 5, nwarps = gthreads >> 5;
    const int nblk = gridDim.x;

    int cur = 0;          // pos[cur] = base iterate x_n, pos[cur^1] = x_np
    int cs = 0;           // sdir[cs], oidx[cs] current
    int scale = R.scale0;
    double threshold = R.threshold0;
    long long ls = 0, restarts = 0;
    int status = 0;
    double maxF = 0.0, lastmaxF = 0.0;
    bool first_pass = true;
    bool overflow = false;
    long long rebuilds = 0, mover_sum = 0;
    unsigned long long t_mark = TIMING ? gtimer() : 0ull;

    // request the operands of an energy phase: positions, directions and mover flags of slots [emin, emax)
    auto issue_e = [&](const double2* p, const double2* d) {
        if (threadIdx.x == 0) {
            const unsigned int ne = (unsigned int)(er.emax - er.emin);
            bulk_load3(S.epos, p + er.emin, ne * 16u, S.edir, d + er.emin, ne * 16u, S.eflag, R.flag + er.emin, (ne + 15u) & ~15u, &S.mbar_e);
        }
        e_inflight = true;
        e_fresh = true;
    };
    // One line search.  With `decide` the call first settles, from the max|F|^2 of the PREVIOUS line search, whether there is a
    // line search at all (2DGrowth.py:131-140,162-164): the partials are requested at the start of the energy phase and read
    // at its end, so the control step costs no L2 round trip and no block-wide synchronisation of its own -- the energy
    // phase runs speculatively (it only writes scratch) and the call returns 1 (converged), 2 (restart: stalled or out of
    // bounds) or 3 (line-search cap) BEFORE anything is published.  Returns 0 after a complete line search.
    // (always_inline: with two call sites the compiler has outlined this lambda in some instantiations, which puts the whole
    // line-search state on the stack)
    auto line_search = [&](bool decide) __attribute__((always_inline)) -> int {
        const double dscale = (double)scale;
        const double inv_step = 1.0 / (20.0 * dscale);
        const double2* pos = R.pos[cur];
        double2* posn = R.pos[cur ^ 1];
        double2* sdir = R.sdir[cs];
        const int* oidx = R.oidx[cs];
        // ---------------- E
        unsigned long long t_blk = (TIMING && R.debug == 8 && threadIdx.x == 0) ? gtimer() : 0ull;
        double mfv[5];                          // this lane's share of the max|F|^2 partials (grids have at most 160 blocks)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int b = lane + 32 * i;
            mfv[i] = (decide && warp == 0 && b < nblk) ? R.red[(size_t)b * 4 + 3] : 0.0;
        }
        if (er.emax > er.emin && !e_fresh) {            // first line of a pass: nothing was requested ahead
            if (e_inflight) { mbar_wait(R, &S.mbar_e, e_parity); e_parity ^= 1u; }
            issue_e(pos, sdir);
        }
        if (blockIdx.x == 0) {          // ordered list of mover slots
            const int T = blockDim.x, t = threadIdx.x;
            const int per = (nA + T - 1) / T;
            const int lo = t * per, hi = min(lo + per, nA);
            int cnt = 0;
            for (int i = lo; i < hi; ++i) cnt += R.flag[i];
            S.u.e.scan[t] = cnt;
            __syncthreads();
            for (int off = 1; off < T; off <<= 1) {
                int v = (t >= off) ? S.u.e.scan[t - off] : 0;
                __syncthreads();
                S.u.e.scan[t] += v;
                __syncthreads();
            }
            int p = S.u.e.scan[t] - cnt;
            for (int i = lo; i < hi; ++i)
                if (R.flag[i]) R.movers[p++] = i;
            if (t == T - 1) *R.n_movers = S.u.e.scan[t];
            // a handful of movers: this block (it has no work items) evaluates them right here, which saves the
            // separate fix-up phase and its grid barrier
            __syncthreads();
            const int nm0 = S.u.e.scan[T - 1];
            if (nblk > 1 && nm0 > 0 && nm0 <= R2_FIX_INLINE)
                for (int w = warp; w < nm0 * KC; w += R2_THREADS / 32) {
                    const double e = relax2_fix_one(R, pos, sdir, oidx, dscale, w / KC, w % KC, lane, nA);
                    if (lane == 0) R.fix[w] = e;
                }
        }
        double acc[PD + 1];
#pragma unroll
        for (int a = 0; a <= PD; ++a) acc[a] = 0.0;
        double eslow = 0.0;                     // lane a < KC: candidate a of the pairs evaluated per candidate
        const int n_items = n_items_reg;
        PHASE2_MARK(32)
        {   // block 0 builds the mover list; when other blocks exist they share the items among themselves
            const int wpb = R2_THREADS / 32;
            const int w0 = nblk > 1 ? gwarp - wpb : gwarp, nw = nblk > 1 ? nwarps - wpb : nwarps;
            (void)w0; (void)nw; (void)n_items;
            EOperands O;
            O.pos = pos; O.dir = sdir; O.flag = R.flag; O.sub = R.sub.sxy;
            if (er.emax > er.emin) {
                mbar_wait(R, &S.mbar_e, e_parity);
                e_parity ^= 1u;
                e_inflight = false;
                e_fresh = false;
                O.pos = S.epos - er.emin; O.dir = S.edir - er.emin; O.flag = S.eflag - er.emin;
            }
            if (er.submax > er.submin) O.sub = S.esub - er.submin;
            relax2_items(R, S.wit[warp], er.kb + warp, wpb, er.ke, O, inv_step, lane, acc, eslow);
        }
        if (P.aa_mode == 0) {
            for (int i = gwarp; i < nA; i += nwarps) {
                const double2 pi = pos[i], si = sdir[i];
                for (int j0 = i + 1; j0 < nA; j0 += 32) {
                    const int j = j0 + lane;
                    const bool valid = j < nA;
                    double dx0 = 0.0, dy0 = 0.0, ddx = 0.0, ddy = 0.0;
                    if (valid) {
                        const double2 pj = pos[j], sj = sdir[j];
                        dx0 = put_in_box(pj.x - pi.x, P.L, P.halfL);
                        dy0 = pj.y - pi.y;
                        ddx = (sj.x - si.x) * inv_step; ddy = (sj.y - si.y) * inv_step;
                    }
                    relax2_pair_round(P, valid, dx0, dy0, ddx, ddy, false, lane, acc, eslow);
                }
            }
        }
        PHASE2_MARK(33)
        double* wrow = S.u.e.wred[warp];            // this warp's row: [0,KC) per-candidate sums, [KC, R2_NE) polynomial
        if (lane < KC) wrow[lane] = eslow;
#pragma unroll
        for (int a = 0; a <= PD; ++a) {
            const double v = group_sum<32>(acc[a]);
            if (lane == 0) wrow[KC + a] = v;
        }
        if (decide && warp == 0) {
            double mf = fmax(fmax(fmax(mfv[0], mfv[1]), fmax(mfv[2], mfv[3])), mfv[4]);
            for (int o = 16; o > 0; o >>= 1) mf = fmax(mf, __shfl_xor_sync(0xffffffffu, mf, o));
            if (lane == 0) S.ctl[3] = mf;
        }
        __syncthreads();
        if (decide) {       // every block, identical arithmetic: the order of the tests is the reference's
            const double mF = S.ctl[3];
            maxF = mF;
            int code = 0;
            if (stall_check && fabs(lastmaxF - mF) < 1e-8) code = 2;         // :162-164 relaxation stalled
            else {
                lastmaxF = mF;
                stall_check = true;
                if (!(mF > threshold)) code = 1;                             // :131 converged
                else if (R.max_ls > 0 && ls >= R.max_ls) code = 3;
                else if (S.ctl[2] > P.Dmax) code = 2;                        // :137-140 atom out of bounds
            }
            if (code) return code;
        }
        if (threadIdx.x < R2_NE) {
            double v = 0.0;
            for (int w = 0; w < R2_THREADS / 32; ++w) v += S.u.e.wred[w][threadIdx.x];
            R.partial[(size_t)threadIdx.x * nblk + blockIdx.x] = v;
        }
        PHASE2_MARK(0)
        if (TIMING && R.debug == 8 && threadIdx.x == 0) R.blk_ns[2 * blockIdx.x] += gtimer() - t_blk;
        if (relax2_barrier(R, epoch)) return 4;
        PHASE2_MARK(1)
        // ---------------- F
        // everything the S phase needs is requested here, in one L2 round trip: the mover count, this thread's atom
        // (the block's chunk of slots, as in the force phase) and this warp's share of the partial sums
        const int nm = *R.n_movers;
        const int per = (nA + nblk - 1) / nblk;
        const int lo = blockIdx.x * per, hi = min(lo + per, nA);
        const int myslot = lo + threadIdx.x;
        double2 pre_p = make_double2(0, 0), pre_s = make_double2(0, 0);
        unsigned int pre_f = 0;
        if (myslot < hi) { pre_p = pos[myslot]; pre_s = sdir[myslot]; pre_f = R.flag[myslot]; }
        constexpr int WPB = R2_THREADS / 32, NEW = (R2_NE + WPB - 1) / WPB;
        double pv[NEW];
#pragma unroll
        for (int i = 0; i < NEW; ++i) {
            const int a = warp + i * WPB;
            double v = 0.0;
            if (a < R2_NE)
                for (int b = lane; b < nblk; b += 32) v += R.partial[(size_t)a * nblk + b];
            pv[i] = v;
        }
        mover_sum += nm;
        if (nm > (nblk > 1 ? R2_FIX_INLINE : 0)) {
            for (int w = gwarp; w < nm * KC; w += nwarps) {
                const double e = relax2_fix_one(R, pos, sdir, oidx, dscale, w / KC, w % KC, lane, nA);
                if (lane == 0) R.fix[w] = e;
            }
            if (relax2_barrier(R, epoch)) return 4;
            PHASE2_MARK(2)
        }
        // ---------------- S
#pragma unroll
        for (int i = 0; i < NEW; ++i) {
            const int a = warp + i * WPB;
            if (a < R2_NE) {
                double v = pv[i];
                if (a < KC)
                    for (int r = lane; r < nm; r += 32) v += R.fix[(size_t)r * KC + a];
                v = group_sum<32>(v);
                if (lane == 0) S.e[a] = v;
            }
        }
        __syncthreads();
        int amin;
        {   // candidate energies = polynomial + per-candidate sums; first argmin (2DGrowth.py:123,155).  EVERY warp forms it from
            // the same shared sums with the same butterfly (identical result in every lane of every warp): no second block
            // synchronisation, no warp-0 section the other fifteen warps wait for
            double ev = INFINITY;
            int ai = lane;
            if (lane < KC) ev = poly_at(S.e + KC, lane) + S.e[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, ev, o);
                const int oi = __shfl_xor_sync(0xffffffffu, ai, o);
                if (ov < ev || (ov == ev && oi < ai)) { ev = ov; ai = oi; }
            }
            amin = ai;
        }
        double top = 0.0, bot = 0.0, my = INFINITY;
        for (int slot = myslot; slot < hi; slot += R2_THREADS) {
            const bool first = slot == myslot;
            const double2 p0 = first ? pre_p : pos[slot];
            const double2 p = cand2(p0, first ? pre_s : sdir[slot], amin, dscale);
            posn[slot] = p;
            top += p.x * (p.x - p0.x) + p.y * (p.y - p0.y);                // :143
            bot += p0.x * p0.x + p0.y * p0.y;                              // :144
            my = fmin(my, p.y);                                            // :137
            if (first ? pre_f : (unsigned int)R.flag[slot]) {
                int nx, ny;
                bin_index(P, p.x, p.y, nx, ny);
                if (flat_cell(P, nx, ny) != R.cellof[slot]) *R.need_rebuild = 1;
            }
        }
        {   // one fused block reduction (fixed tree) of the three partials
            top = group_sum<32>(top); bot = group_sum<32>(bot); my = group_min<32>(my);
            if (lane == 0) { S.u.e.wred[warp][0] = top; S.u.e.wred[warp][1] = bot; S.u.e.wred[warp][2] = my; }
            __syncthreads();
            if (warp == 0) {
                double t2 = lane < WPB ? S.u.e.wred[lane][0] : 0.0, b2 = lane < WPB ? S.u.e.wred[lane][1] : 0.0, m2 = lane < WPB ? S.u.e.wred[lane][2] : INFINITY;
                t2 = group_sum<32>(t2); b2 = group_sum<32>(b2); m2 = group_min<32>(m2);
                if (lane == 0) {
                    double* o = R.red + (size_t)blockIdx.x * 4;
                    o[0] = t2; o[1] = b2; o[2] = m2;
                }
            }
        }
        PHASE2_MARK(3)
        if (relax2_barrier(R, epoch)) return 4;
        PHASE2_MARK(4)
        if (TIMING && R.debug == 8 && threadIdx.x == 0) t_blk = gtimer();
        // x_np is published: fetch this block's neighbourhood [st_min, st_max) (two elements per thread) together with
        // the beta partials -- ONE L2 round trip, one __syncthreads -- into shared memory for the force sweep.
        // (A TMA bulk copy of the same ~5 KB took ~1.9 us from issue to completion here; plain loads take ~0.7 us.)
        const bool early = st_max > st_min && R.debug != 6;
        const int rebuild_now = *R.need_rebuild;             // requested with the rest
        {
            const int scnt = st_max - st_min;
            double2 sv[R2_STAGE / R2_THREADS];
#pragma unroll
            for (int k = 0; k < R2_STAGE / R2_THREADS; ++k) {
                const int i = threadIdx.x + k * R2_THREADS;
                sv[k] = (early && i < scnt) ? posn[st_min + i] : make_double2(0.0, 0.0);
            }
            // every block: beta and min-y from the same partials in the same order
            if (threadIdx.x < 32) {
                double t2 = 0.0, b2 = 0.0, m2 = INFINITY;
                for (int b = lane; b < nblk; b += 32) {
                    const double* o = R.red + (size_t)b * 4;
                    t2 += o[0]; b2 += o[1]; m2 = fmin(m2, o[2]);
                }
                t2 = group_sum<32>(t2); b2 = group_sum<32>(b2); m2 = group_min<32>(m2);
                if (lane == 0) { S.ctl[0] = t2; S.ctl[1] = b2; S.ctl[2] = m2; }
            }
#pragma unroll
            for (int k = 0; k < R2_STAGE / R2_THREADS; ++k) {
                const int i = threadIdx.x + k * R2_THREADS;
                if (early && i < scnt) S.stage[i] = sv[k];
            }
        }
        __syncthreads();
        const double beta = S.ctl[0] / S.ctl[1];                          // :143-145
        PHASE2_MARK(10)
        // ---------------- rebuild (rare) + G
        bool issue_here = false;
        if (rebuild_now) {
            issue_here = true;
            if (relax2_rebuild<TALL>(R, posn, sdir, oidx, R.pos[cur], R.sdir[cs ^ 1], R.oidx[cs ^ 1], epoch, S, nA)) return 4;
            if (gtid == 0) *R.need_rebuild = 0;
            if (*R.n_items > R.item_cap) overflow = true;
            ++rebuilds;
            layout_x0 = false;
            relax2_stage_range(R, S, st_min, st_max, er, nA);
            n_items_reg = *R.n_items;
            PHASE2_MARK(7)
            cs ^= 1;          // the rebuilt x_np now lives in pos[cur]: make it the "next" buffer by flipping cur
            cur ^= 1;
        }
        PHASE2_MARK(11)
        relax2_forces<TIMING>(R, R.pos[cur ^ 1], R.sdir[cs], R.oidx[cs], 1, beta, dscale, false, S, tma_parity, st_min, st_max, issue_here, nA, &t_mark);
        PHASE2_MARK(5)
        if (TIMING && R.debug == 8 && threadIdx.x == 0) R.blk_ns[2 * blockIdx.x + 1] += gtimer() - t_blk;
        if (relax2_barrier(R, epoch)) return 4;
        PHASE2_MARK(6)
        // the next line search (if there is one) starts from x_np with the directions and flags just published: request
        // its operands now, the copies land while the control arithmetic runs
        if (er.emax > er.emin) issue_e(R.pos[cur ^ 1], R.sdir[cs]);
        ++ls;
        return 0;
    };

    for (;;) {                                                         // one pass == one (recursive) call of Relaxation
        if (scale > 1024) { scale = 16; threshold *= 10; }              // :111-113
        e_fresh = false;                                                // the pass rewrites positions, directions and flags
        const int per = (nA + nblk - 1) / nblk;
        const int lo = blockIdx.x * per, hi = min(lo + per, nA);
        if (layout_x0) {
            // restart from the ORIGINAL positions while the sorted layout is still the one built from them: the cell list,
            // the work items and the first direction dx0 = F(x0) are unchanged (only the mover flags depend on the new
            // scale), so the pass starts with a gather instead of a rebuild + force sweep
            const double2* f0 = R.sdir[cs ^ 1];
            double2* pd = R.pos[cur];
            double2* sd = R.sdir[cs];
            const int* oi = R.oidx[cs];
            for (int slot = lo + threadIdx.x; slot < hi; slot += R2_THREADS) {
                const double2 p = R.x0[oi[slot]], s0 = f0[slot];
                pd[slot] = p;
                sd[slot] = s0;
                R.flag[slot] = mover_flag(P, p, s0, (double)scale);
            }
        } else {
            if (relax2_rebuild<TALL>(R, R.x0, nullptr, nullptr, R.pos[cur], R.sdir[cs], R.oidx[cs], epoch, S, nA)) { status = -8; break; }
            if (*R.n_items > R.item_cap) { status = -4; break; }            // work list too small: the host grows it and relaunches
            relax2_stage_range(R, S, st_min, st_max, er, nA);
            n_items_reg = *R.n_items;
            relax2_forces<TIMING>(R, R.pos[cur], R.sdir[cs], R.oidx[cs], 0, 0.0, (double)scale, first_pass, S, tma_parity, st_min, st_max, true, nA);   // dx0, sn = dx0  :120,125
            first_pass = false;
            __syncthreads();
            for (int slot = lo + threadIdx.x; slot < hi; slot += R2_THREADS) R.sdir[cs ^ 1][slot] = R.sdir[cs][slot];     // dx0, kept for the restarts
            layout_x0 = true;
        }
        if (relax2_barrier(R, epoch)) { status = -8; break; }
        int code = line_search(false);                                  // :121-129
        if (overflow) { status = -4; break; }
        stall_check = false;                                            // lastmaxF = maxF after the first line search (:130)
        for (;;) {
            cur ^= 1;                                                   // x_n <- x_np ; s_n already updated in phase G  (:147-150)
            code = line_search(true);                                   // :131-160, the exit tests included
            if (code) { cur ^= 1; break; }                              // no line search happened: x_np stays where it was
            if (overflow) { status = -4; break; }
        }
        if (code == 3) status = -6;
        if (code == 4) status = -8;                                     // watchdog: a spin loop timed out (KMC_E_TIMEOUT)
        const bool restart = code == 2;
        if (status != 0 || !restart) break;
        ++restarts;
        scale *= 2;
    }
    if (e_inflight) mbar_wait(R, &S.mbar_e, e_parity);                     // no copy may be pending when the block exits
    // result: PutAllInBox(x_np) in the original order   (:166)
    if (status != -4 && status != -8) {
        const double2* posn = R.pos[cur ^ 1];
        const int* oidx = R.oidx[cs];
        for (int slot = gtid; slot < nA; slot += gthreads) {
            double2 p = posn[slot];
            p.x = put_in_box(p.x, P.L, P.halfL);
            R.x0[oidx[slot]] = p;
        }
    }
    if (TIMING && (int)blockIdx.x == R.phase_block && threadIdx.x == 0)
        for (int k = 0; k < 48; ++k)
            if (k != 8 && k != 9) R.phase_ns[k] = S.ph[k];
    if (gtid == 0) {
        R.stats->line_searches = ls;
        R.stats->restarts = restarts;
        R.stats->maxF = maxF;
        R.stats->final_threshold = threshold;
        R.stats->final_scale = scale;
        R.stats->status = status;
        R.phase_ns[8] = (unsigned long long)rebuilds;
        R.phase_ns[9] = (unsigned long long)mover_sum;
    }
}

}  // namespace kmc
