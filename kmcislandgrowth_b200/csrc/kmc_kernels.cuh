// kmc_kernels.cuh -- the CUDA kernels of the hot path (sm_100a).
//
// Kernel families:
//   bin_*      cell list build: cell ids (bit-exact BinIndex), counts, scan, scatter, per-cell order
//   sweep_*    pair sweeps over the cell lists: forces, configuration energy (batched), rates
//   probe      energies / closest distances of probe points (HoppingEnergy, DepositAdatom)
//   select     sequential-order partial sums + first-greater search
//   misc       wrap, distance matrix, list gathers, line-search candidates, relaxation control
#pragma once
#include "kmc_device.cuh"

namespace kmc {

// =============================================================================== cell lists

// Bins.py:9-22 for n points -> (nx, ny)
static __global__ void k_bin_index(DevParams P, const double2* __restrict__ xy, int n, int2* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 p = xy[i];
    int nx, ny;
    bin_index(P, p.x, p.y, nx, ny);
    out[i] = make_int2(nx, ny);
}

// B configurations of n atoms: xy[b*n+i].  cell_of[b*n+i] = flat cell (ncell = outside); cnt[b*(ncell+2)+1+cell]++
static __global__ void k_bin_count(DevParams P, const double2* __restrict__ xy, int n, int B, int* __restrict__ cell_of,
                            int* __restrict__ cnt) {
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double2 p = xy[(size_t)b * n + i];
        int nx, ny;
        bin_index(P, p.x, p.y, nx, ny);
        const int c = flat_cell(P, nx, ny);
        cell_of[(size_t)b * n + i] = c;
        atomicAdd(&cnt[(size_t)b * (P.ncell + 2) + 1 + c], 1);
    }
}

// in-place inclusive scan of cnt[b][1..ncell+1] (cnt[b][0] stays 0) -> start offsets; cursor = copy.
// One CTA per configuration; chunked: each thread owns a contiguous chunk.
static __global__ void k_bin_scan(int ncell2 /* ncell+2 */, int* __restrict__ start, int* __restrict__ cursor) {
    __shared__ int s_tot[1024];
    const int b = blockIdx.x;
    int* a = start + (size_t)b * ncell2;
    int* cur = cursor + (size_t)b * ncell2;
    const int m = ncell2 - 1;                         // entries 1..ncell+1
    const int per = (m + blockDim.x - 1) / blockDim.x;
    const int lo = 1 + threadIdx.x * per;
    const int hi = min(lo + per, ncell2);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += a[i];
    s_tot[threadIdx.x] = sum;
    __syncthreads();
    // exclusive scan of the per-thread totals (Hillis-Steele in shared memory)
    for (int off = 1; off < blockDim.x; off <<= 1) {
        int v = (threadIdx.x >= off) ? s_tot[threadIdx.x - off] : 0;
        __syncthreads();
        s_tot[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_tot[threadIdx.x] - sum;
    for (int i = lo; i < hi; ++i) {
        run += a[i];
        a[i] = run;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncell2; i += blockDim.x) cur[i] = a[i];
}

// Large grids (one configuration, > 16k cells): tiled scan over the whole device instead of one block.
// Pass 1: every block scans its tile of SCAN_TILE entries of a[1..m] in place (inclusive) and stores the tile total.
// Pass 2: every block adds the sum of the totals of the tiles before it, and mirrors the result into `cursor`.
constexpr int SCAN_TILE = 8192;
static __global__ void __launch_bounds__(1024) k_bin_scan_tiles(int m, int* __restrict__ a /* entries 1..m */, int* __restrict__ tile_tot) {
    __shared__ int s_tot[1024];
    const int per = SCAN_TILE / 1024;
    const int lo = 1 + blockIdx.x * SCAN_TILE + threadIdx.x * per;
    int v[SCAN_TILE / 1024];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < per; ++k) { v[k] = (lo + k <= m) ? a[lo + k] : 0; sum += v[k]; }
    s_tot[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int t = (threadIdx.x >= off) ? s_tot[threadIdx.x - off] : 0;
        __syncthreads();
        s_tot[threadIdx.x] += t;
        __syncthreads();
    }
    int run = s_tot[threadIdx.x] - sum;
#pragma unroll
    for (int k = 0; k < per; ++k) { run += v[k]; if (lo + k <= m) a[lo + k] = run; }
    if (threadIdx.x == 1023) tile_tot[blockIdx.x] = s_tot[1023];
}

static __global__ void __launch_bounds__(1024) k_bin_scan_add(int m, int* __restrict__ a, const int* __restrict__ tile_tot, int* __restrict__ cursor) {
    __shared__ int s_red[32];
    __shared__ int s_base;
    int part = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += 1024) part += tile_tot[t];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) { int b = 0; for (int w = 0; w < 32; ++w) b += s_red[w]; s_base = b; }
    __syncthreads();
    const int base = s_base;
    const int lo = 1 + blockIdx.x * SCAN_TILE;
    for (int i = lo + threadIdx.x; i < lo + SCAN_TILE && i <= m; i += 1024) { const int v = a[i] + base; a[i] = v; cursor[i] = v; }
    if (blockIdx.x == 0 && threadIdx.x == 0) cursor[0] = a[0];
}

// slot = cursor[cell]++ ; sidx[slot] = i   (order inside a cell fixed up by k_bin_order)
static __global__ void k_bin_scatter(int n, int B, int ncell2, const int* __restrict__ cell_of, int* __restrict__ cursor,
                              int* __restrict__ sidx) {
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = cell_of[(size_t)b * n + i];
        const int slot = atomicAdd(&cursor[(size_t)b * ncell2 + c], 1);
        sidx[(size_t)b * n + slot] = i;
    }
}

// restore the reference's input order inside every cell (Bins.py:40-41 appends in input order) and
// gather the positions.  One thread per cell (sort_indices: kmc_device.cuh).
static __global__ void k_bin_order(int n, int B, int ncell2, const int* __restrict__ start, int* __restrict__ sidx,
                            const double2* __restrict__ xy, double2* __restrict__ sxy) {
    const int b = blockIdx.y;
    const int* st = start + (size_t)b * ncell2;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell2 - 1; c += gridDim.x * blockDim.x) {
        const int lo = st[c], hi = st[c + 1];
        int* a = sidx + (size_t)b * n + lo;
        // (the last cell holds the atoms outside the grid: nobody's stencil addresses it, so the order inside it is
        // immaterial -- a large population there, e.g. the empty halo slots of a domain-decomposed configuration, is
        // not sorted by this one thread)
        if (!(c == ncell2 - 2 && hi - lo > 64)) sort_indices(a, hi - lo);
        for (int s = lo; s < hi; ++s) sxy[(size_t)b * n + s] = xy[(size_t)b * n + sidx[(size_t)b * n + s]];
    }
}

// =============================================================================== pair sweeps

// Energy.py:62-75.  One group of G lanes per adatom (cell-sorted order so that a warp shares its
// stencil); f_aa / f_as / f_sum are written at the atom's ORIGINAL index; any may be null.
// `which_cfg` (device pointer, may be null) selects the configuration of a batch (line search winner).
template <int G>
static __global__ void __launch_bounds__(256) k_sweep_forces(DevParams P, CellView ad_base, CellView sub, int n,
                                                      const int* __restrict__ which_cfg, int ncell2,
                                                      double2* __restrict__ f_aa, double2* __restrict__ f_as,
                                                      double2* __restrict__ f_sum, int n_compute = -1) {
    const int cfg = which_cfg ? *which_cfg : 0;
    CellView ad = ad_base;
    if (ad.start) ad.start += (size_t)cfg * ncell2;
    ad.sxy += (size_t)cfg * n;
    if (ad.sidx) ad.sidx += (size_t)cfg * n;      // null sidx == identity order (aa_mode 0 needs no cell list)
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x & (G - 1);
    bool live = gid < n;
    double ax = 0, ay = 0, sx = 0, sy = 0;
    int idx = 0;
    if (live) {
        idx = ad.sidx ? ad.sidx[gid] : gid;
        // domain decomposition: only the first n_compute atoms (the rank's own) are evaluated, the rest are halo copies
        if (n_compute >= 0 && idx >= n_compute) live = false;
    }
    if (live) {
        const double2 pi = ad.sxy[gid];
        const Stencil st = make_stencil(P, pi.x, pi.y);
        if (P.aa_mode == 0) {
            for (int s = lane; s < n; s += G) acc_force(P, pi.x, pi.y, ad.sxy[s], P.E_a, P.ra2, ax, ay);
        } else {
            stencil_forces4(P, ad, st, lane, G, pi.x, pi.y, P.E_a, P.ra2, ax, ay);
        }
        stencil_forces4(P, sub, st, lane, G, pi.x, pi.y, P.E_as, P.ras2, sx, sy);
    }
    ax = group_sum<G>(ax); ay = group_sum<G>(ay);
    sx = group_sum<G>(sx); sy = group_sum<G>(sy);
    if (live && lane == 0) {
        if (f_aa) f_aa[idx] = make_double2(ax, ay);
        if (f_as) f_as[idx] = make_double2(sx, sy);
        if (f_sum) f_sum[idx] = make_double2(ax + sx, ay + sy);     // 2DGrowth.py:120: F_aa + F_as
    }
}

// Energy.py:133-148 for B configurations: E[b] = sum_i ( sum_{j visible from i, idx_j > idx_i} LJ_aa
//                                                        + sum_{s in substrate stencil of i} LJ_as ).
// grid = (blocks, B).  Deterministic: per-block partials, the last block of a configuration adds them
// in block order.  done[b] must be zero on entry and is reset on exit.
template <int G>
static __global__ void __launch_bounds__(256) k_sweep_energy(DevParams P, CellView ad_base, CellView sub, int n, int ncell2,
                                                      double* __restrict__ partial, unsigned int* __restrict__ done,
                                                      double* __restrict__ e_out, int n_compute = -1) {
    __shared__ double s_red[32];
    __shared__ bool s_last;
    const int b = blockIdx.y;
    CellView ad = ad_base;
    if (ad.start) ad.start += (size_t)b * ncell2;
    ad.sxy += (size_t)b * n;
    if (ad.sidx) ad.sidx += (size_t)b * n;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x & (G - 1);
    double e = 0.0;
    if (gid < n) {
        const double2 pi = ad.sxy[gid];
        const int idx = ad.sidx ? ad.sidx[gid] : gid;
        const Stencil st = make_stencil(P, pi.x, pi.y);
        if (n_compute >= 0) {
            // domain decomposition (bins3x3 only): an own atom takes HALF of each of its pair terms, the other half is taken
            // by the partner -- on this rank when it is an own atom too, on its owner's rank when it is a halo copy; halo
            // copies themselves contribute nothing here.  Summed over the ranks this is the configuration energy.
            if (idx < n_compute) {
                visit_stencil(P, ad, st, lane, G, [&](int s) { e += 0.5 * pair_energy(P, pi.x, pi.y, ad.sxy[s], P.E_a, P.ra2); });
                visit_stencil(P, sub, st, lane, G, [&](int s) { e += pair_energy(P, pi.x, pi.y, sub.sxy[s], P.E_as, P.ras2); });
            }
        } else {
        if (P.aa_mode == 0) {
            for (int s = lane; s < n; s += G)
                if ((ad.sidx ? ad.sidx[s] : s) > idx) e += pair_energy(P, pi.x, pi.y, ad.sxy[s], P.E_a, P.ra2);
        } else {
            visit_stencil(P, ad, st, lane, G, [&](int s) {
                if (ad.sidx[s] > idx) e += pair_energy(P, pi.x, pi.y, ad.sxy[s], P.E_a, P.ra2);
            });
        }
        visit_stencil(P, sub, st, lane, G, [&](int s) { e += pair_energy(P, pi.x, pi.y, sub.sxy[s], P.E_as, P.ras2); });
        }
    }
    const double tot = block_sum(e, s_red);
    if (threadIdx.x == 0) {
        partial[(size_t)b * gridDim.x + blockIdx.x] = tot;
        __threadfence();
        const unsigned int t = atomicAdd(&done[b], 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double v = 0.0;
        if (threadIdx.x < 32) {
            for (int k = threadIdx.x; k < (int)gridDim.x; k += 32) v += partial[(size_t)b * gridDim.x + k];
            v = group_sum<32>(v);
            if (threadIdx.x == 0) { e_out[b] = v; done[b] = 0u; }
        }
    }
}

// Energy.py:121-131 TotalEnergy (unused by the driver, kept for API parity): quirky loop bounds.
// Small-N utility: one block, all-pairs.
static __global__ void k_total_energy(DevParams P, const double2* __restrict__ xy, CellView sub, int n, double* __restrict__ e_out) {
    __shared__ double s_red[32];
    double e = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 pi = xy[i];
        const bool in_loop = i < n - 2, last = (i == n - 1);
        if (in_loop)
            for (int j = i + 1; j < n; ++j) e += pair_energy(P, pi.x, pi.y, xy[j], P.E_a, P.ra2);
        if (in_loop || last) {
            const Stencil st = make_stencil(P, pi.x, pi.y);
            visit_stencil(P, sub, st, 0, 1, [&](int s) { e += pair_energy(P, pi.x, pi.y, sub.sxy[s], P.E_as, P.ras2); });
        }
    }
    const double tot = block_sum(e, s_red);
    if (threadIdx.x == 0) e_out[0] = tot;
}

// Energy.py:174-217 + 2DGrowth.py:46-62,201-215 for every adatom.  Bond tests use the reference's exact
// arithmetic (dist2_exact + IEEE sqrt) so that n_a / n_s / the surface flag agree bit for bit.
template <int G>
static __global__ void __launch_bounds__(256) k_sweep_rates(DevParams P, CellView ad, CellView sub, int n,
                                                     double* __restrict__ rate, uint8_t* __restrict__ is_surface,
                                                     double* __restrict__ delta_u, double* __restrict__ u_appx_o,
                                                     double* __restrict__ u_loc_o, int* __restrict__ n_a_o,
                                                     int* __restrict__ n_s_o,
                                                     // incremental update (kmc_rates_update): only atoms that moved or whose
                                                     // cell is marked are recomputed; their reference positions are refreshed
                                                     const uint8_t* __restrict__ dirty_cell = nullptr, const uint8_t* __restrict__ self_dirty = nullptr,
                                                     double2* __restrict__ ref_xy = nullptr, unsigned long long* __restrict__ recomputed = nullptr,
                                                     int n_compute = -1, double* __restrict__ e_aa_o = nullptr, double* __restrict__ e_as_o = nullptr) {
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x & (G - 1);
    bool live = gid < n;
    double ua = 0, us = 0, la = 0, ls = 0;
    int ka = 0, ks = 0, idx = 0;
    double2 pi = make_double2(0.0, 0.0);
    if (live) {
        pi = ad.sxy[gid];
        idx = ad.sidx[gid];
        if (n_compute >= 0 && idx >= n_compute) live = false;      // halo copy (domain decomposition)
        if (live && dirty_cell) {
            int nx, ny;
            bin_index(P, pi.x, pi.y, nx, ny);
            const int cell = flat_cell(P, nx, ny);
            live = self_dirty[idx] || cell == P.ncell || dirty_cell[cell];     // atoms outside the grid are always redone
        }
    }
    if (live) {
        const Stencil st = make_stencil(P, pi.x, pi.y);
        if (P.aa_mode == 0)
            for (int s = lane; s < n; s += G) ua += pair_energy(P, pi.x, pi.y, ad.sxy[s], P.E_a, P.ra2);
        visit_stencil(P, ad, st, lane, G, [&](int s) {
            const double2 pj = ad.sxy[s];
            const double r2 = dist2_exact(P, pi.x, pi.y, pj.x, pj.y);
            const double e = lj_energy(r2, P.E_a, P.ra2);
            if (P.aa_mode != 0) ua += e;
            if (sqrt(r2) < P.bond_a) { la += e; ++ka; }               // Bins.py:100 (the atom itself counts)
        });
        visit_stencil(P, sub, st, lane, G, [&](int s) {
            const double2 pj = sub.sxy[s];
            const double r2 = dist2_exact(P, pi.x, pi.y, pj.x, pj.y);
            const double e = lj_energy(r2, P.E_as, P.ras2);
            us += e;
            if (sqrt(r2) < P.bond_s) { ls += e; ++ks; }               // Bins.py:107
        });
    }
    ua = group_sum<G>(ua); us = group_sum<G>(us); la = group_sum<G>(la); ls = group_sum<G>(ls);
    ka = group_sum_i<G>(ka); ks = group_sum_i<G>(ks);
    if (live && lane == 0) {
        const double appx = ua + us;                                  // Energy.py:181
        const double loc = la + ls;                                   // Energy.py:192
        const double ideal = -((double)ka * P.E_a + (double)ks * P.E_as);   // Energy.py:202
        const double du = 1.0 * (loc - ideal) + appx;                 // Energy.py:217, C() = 1.0
        const bool surf = (ka + ks) < 6;                              // 2DGrowth.py:59
        if (rate) rate[idx] = surf ? P.omega * exp(du * P.beta) : 0.0;   // 2DGrowth.py:215
        if (is_surface) is_surface[idx] = surf ? 1 : 0;
        if (delta_u) delta_u[idx] = du;
        if (u_appx_o) u_appx_o[idx] = appx;
        if (e_aa_o) e_aa_o[idx] = ua;
        if (e_as_o) e_as_o[idx] = us;
        if (u_loc_o) u_loc_o[idx] = loc;
        if (n_a_o) n_a_o[idx] = ka;
        if (n_s_o) n_s_o[idx] = ks;
        if (ref_xy) ref_xy[idx] = pi;
        if (recomputed) atomicAdd(recomputed, 1ull);
    }
}

// ---- incremental rate table: which cells hold atoms whose rate may have changed ----------------------
// An atom's barrier depends on the atoms of its own 3x3(+seam) stencil (Bins.py:44-84, bins3x3 adatom sums), and the
// relation "cell D is in the stencil of cell C" is symmetric (both wraps and the seam column come in pairs), so an atom
// that moved by more than the threshold since its neighbours' rates were last computed dirties the stencil cells of its
// old and of its new position.  `pending` = positions of atoms that were removed (HopAtom, 2DGrowth.py:241).
static __global__ void __launch_bounds__(256) k_rates_mark(DevParams P, const double2* __restrict__ xy, const double2* __restrict__ ref, int n, double thr2,
                                                    const double2* __restrict__ pending, int npending, uint8_t* __restrict__ dirty_cell,
                                                    uint8_t* __restrict__ self_dirty) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto mark = [&](double x, double y) {
        const Stencil st = make_stencil(P, x, y);
        for (int a = 0; a < st.nxs; ++a) {
            const int cx = st.xs[a];
            if (cx < 0 || cx >= P.nbx) continue;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int cy = st.ys[b];
                if (cy < 0 || cy >= P.nby) continue;
                dirty_cell[cx * P.nby + cy] = 1;
            }
        }
    };
    if (i < n) {
        const double2 p = xy[i], q = ref[i];
        bool moved;
        if (isnan(q.x)) moved = true;                                   // no cached entry (new atom)
        else if (thr2 == 0.0) moved = (p.x != q.x) || (p.y != q.y);     // exact mode: any change at all
        else {
            const double dx = p.x - q.x, dy = p.y - q.y;
            moved = dx * dx + dy * dy > thr2;
        }
        self_dirty[i] = moved ? 1 : 0;
        if (moved) {
            mark(p.x, p.y);
            if (!isnan(q.x)) mark(q.x, q.y);
        }
    }
    if (i < npending) mark(pending[i].x, pending[i].y);
}

// erase element k of an array (the cached tables follow HopAtom's np.delete, 2DGrowth.py:241)
template <class T>
static __global__ void __launch_bounds__(256) k_erase_at(const T* __restrict__ src, int n, int k, T* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) dst[i] = src[i < k ? i : i + 1];
}

// Energy.py:170-172 / :77-95 and the closeness test of 2DGrowth.py:82-90.  One warp per probe point.
static __global__ void __launch_bounds__(128) k_probe(DevParams P, const double2* __restrict__ probes, int m, CellView ad, CellView sub,
                                               double* __restrict__ e_aa, double* __restrict__ e_as,
                                               double* __restrict__ min_d_a, double* __restrict__ min_d_s) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const bool live = w < m;
    double ea = 0, es = 0, da = INFINITY, ds = INFINITY;
    if (live) {
        const double2 pi = probes[w];
        const Stencil st = make_stencil(P, pi.x, pi.y);
        if (ad.n > 0) {
            if (P.aa_mode == 0)
                for (int s = lane; s < ad.n; s += 32) ea += pair_energy(P, pi.x, pi.y, ad.sxy[s], P.E_a, P.ra2);
            visit_stencil(P, ad, st, lane, 32, [&](int s) {
                const double2 pj = ad.sxy[s];
                const double r2 = dist2_exact(P, pi.x, pi.y, pj.x, pj.y);
                if (P.aa_mode != 0) ea += lj_energy(r2, P.E_a, P.ra2);
                da = fmin(da, sqrt(r2));
            });
        }
        visit_stencil(P, sub, st, lane, 32, [&](int s) {
            const double2 pj = sub.sxy[s];
            const double r2 = dist2_exact(P, pi.x, pi.y, pj.x, pj.y);
            es += lj_energy(r2, P.E_as, P.ras2);
            ds = fmin(ds, sqrt(r2));
        });
    }
    ea = group_sum<32>(ea); es = group_sum<32>(es);
    da = group_min<32>(da); ds = group_min<32>(ds);
    if (live && lane == 0) {
        if (e_aa) e_aa[w] = ea;
        if (e_as) e_as[w] = es;
        if (min_d_a) min_d_a[w] = da;
        if (min_d_s) min_d_s[w] = ds;
    }
}

// =============================================================================== selection

// 2DGrowth.py:217-227,322-329.  The partial sums are formed strictly left to right like the
// reference's sum(Rk[:i+1]) so that every pj -- and therefore the selected index -- is bit-identical
// for the same rates and the same u.  One warp: lanes prefetch 32 rates, lane 0 owns the running sum.
// result[0] = compacted k or -1 ; result[1] = dense index or -1 ; rtot[0] = Rd + sum.
static __global__ void k_select(DevParams P, const double* __restrict__ rate, const uint8_t* __restrict__ surf, int n, double u,
                         long long* __restrict__ result, double* __restrict__ rtot, double* __restrict__ pj_out) {
    const int lane = threadIdx.x;
    // Adding +0.0 never changes a running sum, and the dense table is zero for every non-surface atom: only the
    // non-zero (or surface-flagged) entries are visited, in index order, so every partial sum keeps the reference's
    // bits while the chain of dependent additions shrinks from n to the number of surface atoms.  Eight 32-entry
    // chunks are requested before the first one is consumed (one L2 round trip per 256 entries instead of per 32).
    constexpr int NB = 8;
    // pass 1: total in reference order
    double sum = 0.0;
    for (int base = 0; base < n; base += 32 * NB) {
        double v[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int i = base + j * 32 + lane;
            v[j] = (i < n) ? rate[i] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            unsigned int m = __ballot_sync(0xffffffffu, v[j] != 0.0);
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                sum = __dadd_rn(sum, __shfl_sync(0xffffffffu, v[j], k));      // identical in every lane
            }
        }
    }
    const double Rtot = __dadd_rn(P.Rd, sum);               // 2DGrowth.py:227
    const double r = __dmul_rn(u, Rtot);                    // 2DGrowth.py:322
    // pass 2: first partial sum > r
    double pj = 0.0;
    int found = -1, kcomp = 0, kfound = -1;
    if (pj_out) {           // the drop-in API asked for every partial sum (HoppingPartialSums): plain sequential walk
        if (lane == 0) pj_out[0] = 0.0;
        for (int base = 0; base < n; base += 32) {
            const double v = (base + lane < n) ? rate[base + lane] : 0.0;
            const int sf = (base + lane < n) ? (int)surf[base + lane] : 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const double vk = __shfl_sync(0xffffffffu, v, k);
                const int sk = __shfl_sync(0xffffffffu, sf, k);
                if (base + k < n) {
                    pj = __dadd_rn(pj, vk);
                    if (lane == 0) pj_out[base + k + 1] = pj;
                    if (found < 0 && sk) {
                        if (pj > r) { found = base + k; kfound = kcomp; }
                        ++kcomp;
                    }
                }
            }
        }
    } else {
        for (int base = 0; base < n && found < 0; base += 32 * NB) {
            double v[NB];
            int sf[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int i = base + j * 32 + lane;
                v[j] = (i < n) ? rate[i] : 0.0;
                sf[j] = (i < n) ? (int)surf[i] : 0;
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                unsigned int m = __ballot_sync(0xffffffffu, v[j] != 0.0 || sf[j] != 0);
                while (m && found < 0) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    pj = __dadd_rn(pj, __shfl_sync(0xffffffffu, v[j], k));
                    if (__shfl_sync(0xffffffffu, sf[j], k)) {
                        if (pj > r) { found = base + j * 32 + k; kfound = kcomp; }
                        ++kcomp;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        result[0] = kfound;
        result[1] = found;
        rtot[0] = Rtot;
    }
}

// ---- parallel selection for very large tables (block scan + search) -------------------------------
// Pass 1: every block sums its contiguous chunk (thread-sequential segments + fixed-tree block sum).
// Pass 2: one block forms the running sums of the block totals in index order, finds the chunk whose range
// contains r = u*Rtot, scans that chunk (thread-sequential segments, ordered prefix of the segment totals) and
// returns the first index with partial sum > r.  Same rule as k_select; the partial sums are associated
// differently (ulp-level), so the index can differ from the sequential order only when r falls within rounding of
// a partial sum.  Used by kmc_select for n > 65536 (or on request).
static __global__ void __launch_bounds__(1024) k_select_block_sums(const double* __restrict__ rate, const uint8_t* __restrict__ surf, int n, int chunk,
                                                            double* __restrict__ bsum, int* __restrict__ bcount) {
    __shared__ double s_red[32];
    __shared__ int s_cnt[32];
    const int lo = blockIdx.x * chunk, hi = min(lo + chunk, n);
    const int per = (chunk + blockDim.x - 1) / blockDim.x;
    const int a = lo + threadIdx.x * per, b = min(a + per, hi);
    double s = 0.0;
    int c = 0;
    for (int i = a; i < b; ++i) { s += rate[i]; c += surf[i]; }
    const double tot = block_sum(s, s_red);
    int cw = c;
    for (int o = 16; o > 0; o >>= 1) cw += __shfl_xor_sync(0xffffffffu, cw, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cw;
    __syncthreads();
    if (threadIdx.x == 0) {
        int ct = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) ct += s_cnt[w];
        bsum[blockIdx.x] = tot;
        bcount[blockIdx.x] = ct;
    }
}

static __global__ void __launch_bounds__(1024) k_select_final(DevParams P, const double* __restrict__ rate, const uint8_t* __restrict__ surf, int n, int chunk,
                                                       int nblocks, const double* __restrict__ bsum, const int* __restrict__ bcount, double u,
                                                       long long* __restrict__ result, double* __restrict__ rtot) {
    __shared__ double s_seg[1024];
    __shared__ int s_segc[1024];
    __shared__ double s_base;
    __shared__ int s_blk, s_kbase;
    __shared__ double s_r;
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int b = 0; b < nblocks; ++b) tot += bsum[b];
        const double Rtot = P.Rd + tot;
        const double r = u * Rtot;
        rtot[0] = Rtot;
        double run = 0.0;
        int blk = -1, kb = 0;
        for (int b = 0; b < nblocks; ++b) {
            if (run + bsum[b] > r) { blk = b; break; }
            run += bsum[b];
            kb += bcount[b];
        }
        s_blk = blk; s_base = run; s_kbase = kb; s_r = r;
        if (blk < 0) { result[0] = -1; result[1] = -1; }
    }
    __syncthreads();
    const int blk = s_blk;
    if (blk < 0) return;
    const int lo = blk * chunk, hi = min(lo + chunk, n);
    const int per = (chunk + blockDim.x - 1) / blockDim.x;
    const int a = lo + threadIdx.x * per, b = min(a + per, hi);
    double s = 0.0;
    int c = 0;
    for (int i = a; i < b; ++i) { s += rate[i]; c += surf[i]; }
    s_seg[threadIdx.x] = s;
    s_segc[threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = s_base;
        int kb = s_kbase, seg = -1;
        for (int t = 0; t < (int)blockDim.x; ++t) {
            if (run + s_seg[t] > s_r) { seg = t; break; }
            run += s_seg[t];
            kb += s_segc[t];
        }
        long long found = -1, kfound = -1;
        if (seg >= 0) {
            const int a2 = lo + seg * per, b2 = min(a2 + per, hi);
            for (int i = a2; i < b2; ++i) {
                run += rate[i];
                if (surf[i]) {
                    if (run > s_r) { found = i; kfound = kb; break; }
                    ++kb;
                }
            }
        }
        if (found < 0) {
            // rounding left r between the chunk's bound and its re-associated sums: fall back to the last surface atom of the chunk
            for (int i = hi - 1; i >= lo; --i)
                if (surf[i]) { found = i; break; }
            kfound = s_kbase;
            for (int i = lo; i < found; ++i) kfound += surf[i];
        }
        result[0] = kfound;
        result[1] = found;
    }
}

// =============================================================================== misc

// Periodic.py:22-25
static __global__ void k_put_all_in_box(DevParams P, double2* __restrict__ xy, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) xy[i].x = put_in_box(xy[i].x, P.L, P.halfL);
}

// Periodic.py:54-62
static __global__ void k_distances(DevParams P, const double2* __restrict__ a, int na, const double2* __restrict__ b, int nb,
                            double* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)na * nb) return;
    const int i = (int)(t / nb), j = (int)(t % nb);
    out[t] = sqrt(dist2_exact(P, a[i].x, a[i].y, b[j].x, b[j].y));
}

// Periodic.py:27-36
static __global__ void k_displacement(DevParams P, const double2* __restrict__ ri, const double2* __restrict__ rj, int k,
                               double2* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    out[i] = make_double2(put_in_box(__dadd_rn(rj[i].x, -ri[i].x), P.L, P.halfL), __dadd_rn(rj[i].y, -ri[i].y));
}

// Bins.py:72-84: stencil populations of nq query points
static __global__ void k_nearby_count(DevParams P, CellView cv, const double2* __restrict__ q, int nq, long long* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    counts[i + 1] = stencil_population(P, cv, make_stencil(P, q[i].x, q[i].y));
    if (i == 0) counts[0] = 0;
}

// in-place inclusive scan of a[1..m] (a[0] = 0), single block (list lengths, not a hot path)
static __global__ void k_scan_i64(long long* __restrict__ a, int m) {
    __shared__ long long s_tot[1024];
    const int per = (m + blockDim.x - 1) / blockDim.x;
    const int lo = 1 + threadIdx.x * per, hi = min(lo + per, m + 1);
    long long sum = 0;
    for (int i = lo; i < hi; ++i) sum += a[i];
    s_tot[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < blockDim.x; off <<= 1) {
        long long v = (threadIdx.x >= off) ? s_tot[threadIdx.x - off] : 0;
        __syncthreads();
        s_tot[threadIdx.x] += v;
        __syncthreads();
    }
    long long run = s_tot[threadIdx.x] - sum;
    for (int i = lo; i < hi; ++i) { run += a[i]; a[i] = run; }
}

static __global__ void k_nearby_fill(DevParams P, CellView cv, const double2* __restrict__ q, int nq,
                              const long long* __restrict__ offs, long long cap, double2* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    long long o = offs[i];
    const Stencil st = make_stencil(P, q[i].x, q[i].y);
    visit_stencil(P, cv, st, 0, 1, [&](int s) {
        if (o < cap) out[o] = cv.sxy[s];
        ++o;
    });
}

// Bins.py:86-111 position lists, reference order; which: 0 adatom bonds, 1 substrate bonds
static __global__ void k_neighbor_fill(DevParams P, const double2* __restrict__ xy, int n, CellView cv, double bond,
                                const long long* __restrict__ offs, long long cap, double2* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 pi = xy[i];
    long long o = offs[i];
    const Stencil st = make_stencil(P, pi.x, pi.y);
    visit_stencil(P, cv, st, 0, 1, [&](int s) {
        const double2 pj = cv.sxy[s];
        if (sqrt(dist2_exact(P, pi.x, pi.y, pj.x, pj.y)) < bond) {
            if (o < cap) out[o] = pj;
            ++o;
        }
    });
}

static __global__ void k_i32_to_offsets(const int* __restrict__ counts, int n, long long* __restrict__ offs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) offs[i + 1] = counts[i];
    if (i == 0) offs[0] = 0;
}

// 2DGrowth.py:121,152: cand[a][i] = x[i] + a*s[i]/20/scale (same operation order: ((a*s)/20)/scale)
static __global__ void k_make_candidates(const double* __restrict__ x, const double* __restrict__ s, int n2, int K, double scale,
                                  double* __restrict__ cand) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (i >= n2 || a >= K) return;
    cand[(size_t)a * n2 + i] = __dadd_rn(x[i], __ddiv_rn(__ddiv_rn(__dmul_rn((double)a, s[i]), 20.0), scale));
}

// ---- relaxation control (2DGrowth.py:97-166) ----------------------------------------------------
struct RelaxCtrl {
    double maxF, miny, top, bot;   // reductions of the current iterate
    int amin;                      // winning candidate of the last line search
    int pad;
};

// first argmin of K energies (ys.index(min(ys)), 2DGrowth.py:123,155)
static __global__ void k_argmin(const double* __restrict__ e, int K, RelaxCtrl* __restrict__ ctrl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int am = 0;
        double em = e[0];
        for (int a = 1; a < K; ++a)
            if (e[a] < em) { em = e[a]; am = a; }
        ctrl->amin = am;
    }
}

// after the forces at xnp = cand[amin]: maxF = max_i |F_i|^2 (:129,160), miny (:137), top = xnp.(xnp-xn),
// bot = xn.xn (:143-144).  Also copies xnp out of the candidate batch.  Single block (deterministic).
static __global__ void k_relax_reduce(const double* __restrict__ cand, const double* __restrict__ xn, const double2* __restrict__ F,
                               int n, RelaxCtrl* __restrict__ ctrl, double* __restrict__ xnp) {
    __shared__ double s_red[32];
    const double* xc = cand + (size_t)ctrl->amin * 2 * n;
    double mf = 0.0, my = INFINITY, top = 0.0, bot = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 f = F[i];
        mf = fmax(mf, f.x * f.x + f.y * f.y);
        const double x = xc[2 * i], y = xc[2 * i + 1];
        const double x0 = xn[2 * i], y0 = xn[2 * i + 1];
        my = fmin(my, y);
        top += x * (x - x0) + y * (y - y0);
        bot += x0 * x0 + y0 * y0;
        xnp[2 * i] = x;
        xnp[2 * i + 1] = y;
    }
    const double r_mf = block_max(mf, s_red);
    const double r_my = block_min(my, s_red);
    const double r_top = block_sum(top, s_red);
    const double r_bot = block_sum(bot, s_red);
    if (threadIdx.x == 0) { ctrl->maxF = r_mf; ctrl->miny = r_my; ctrl->top = r_top; ctrl->bot = r_bot; }
}

// snp = dxnp + beta*sn ; xn = xnp ; sn = snp  (2DGrowth.py:147-150)
static __global__ void k_cg_update(const double* __restrict__ F, double beta, double* __restrict__ sn, double* __restrict__ xn,
                            const double* __restrict__ xnp, int n2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    sn[i] = __dadd_rn(F[i], __dmul_rn(beta, sn[i]));
    xn[i] = xnp[i];
}

static __global__ void k_min_y(const double2* __restrict__ xy, int n, int want_max, double* __restrict__ out) {
    __shared__ double s_red[32];
    double v = want_max ? -INFINITY : INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v = want_max ? fmax(v, xy[i].y) : fmin(v, xy[i].y);
    const double r = want_max ? block_max(v, s_red) : block_min(v, s_red);
    if (threadIdx.x == 0) out[0] = r;
}

}  // namespace kmc
