// kmc_device.cuh -- device-side building blocks shared by every kernel of libkmcb200.so.
//
// Geometry follows the reference bit for bit where the reference's result is an integer or a
// comparison (bin ids, stencil columns, bond tests); pair energies / forces use the algebraically
// identical r^2 forms (no sqrt, no pow) and agree with the reference to a few ulp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kmc {

// POD handed to every kernel by value (lives in the constant bank of the launch).
struct DevParams {
    double L, halfL, bin_size, Dmin, Dmax;
    double E_a, r_a, E_as, r_as;
    double ra2, ras2;        // r0^2 for the r^2-based LJ forms
    double inv_ra2, inv_ras2;
    double bond_a, bond_s;   // 1.2*r_a, 1.2*r_as           (Bins.py:100,107)
    double beta, omega, Rd;
    int nbx, nby, ncell;     // ncell = nbx*nby ; cell id ncell = atoms outside the addressable grid
    int aa_mode;
    int precision;           // 0: fp64 pair terms; 1: fp32 pair terms on fp64-formed displacements, fp64 accumulation
};

// A cell list: atoms sorted by cell (column-major nx*nby+ny, input order kept inside a cell).
// (no __restrict__: the persistent relaxation kernel rewrites these arrays between grid barriers, so
// loads must stay on the coherent path)
struct CellView {
    const int* start;      // [ncell+2]
    const double2* sxy;    // [n] cell-sorted positions
    const int* sidx;       // [n] original index of every sorted slot
    int n;
};

// Periodic.py:8-20 applied to a displacement/position: Python float modulo, then fold to (-L/2, L/2].
// For |d| < L fmod is the identity, so the common path is two predicated adds and reproduces the
// reference's two roundings exactly.
__device__ __forceinline__ double put_in_box(double d, double L, double halfL) {
    double r = d;
    if (!(fabs(d) < L)) {
        r = fmod(d, L);
        if (r == 0.0) r = 0.0;           // copysign(0, L), L > 0
    }
    if (r < 0.0) r = __dadd_rn(r, L);
    if (r > halfL) r = __dadd_rn(r, -L);
    return r;
}

// Bins.py:9-22 -- IEEE add, IEEE divide, truncation toward zero (never a reciprocal multiply).
__device__ __forceinline__ void bin_index(const DevParams& P, double x, double y, int& nx, int& ny) {
    nx = (int)__ddiv_rn(__dadd_rn(x, P.halfL), P.bin_size);
    ny = (int)__ddiv_rn(__dadd_rn(y, -P.Dmin), P.bin_size);
}

__device__ __forceinline__ int flat_cell(const DevParams& P, int nx, int ny) {
    return (nx >= 0 && nx < P.nbx && ny >= 0 && ny < P.nby) ? nx * P.nby + ny : P.ncell;
}

// Bins.py:44-70: three columns / rows wrapped ONCE, seam column appended for nx==0 / nx==nbx-2.
struct Stencil {
    int xs[4];
    int ys[3];
    int nxs;
};

__device__ __forceinline__ Stencil make_stencil(const DevParams& P, double x, double y) {
    Stencil s;
    int nx, ny;
    bin_index(P, x, y, nx, ny);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int cx = nx - 1 + i, cy = ny - 1 + i;
        if (cx < 0) cx += P.nbx;
        if (cx >= P.nbx) cx -= P.nbx;
        if (cy < 0) cy += P.nby;
        if (cy >= P.nby) cy -= P.nby;
        s.xs[i] = cx;
        s.ys[i] = cy;
    }
    s.nxs = 3;
    s.xs[3] = -1;
    if (nx == 0) { s.xs[3] = P.nbx - 2; s.nxs = 4; }
    else if (nx == P.nbx - 2) { s.xs[3] = 0; s.nxs = 4; }
    return s;
}

// Visit every atom slot of the stencil in the reference's NearbyAtoms order (Bins.py:76-83):
// for cx in xs: for cy in ys: atoms of the cell.  `lane`/`G` split one atom's neighbours over a
// group of G lanes (G = 1 keeps the exact sequential order).  Cells of one column are adjacent in
// memory, so the usual case is ONE contiguous range per column.
template <class Fn>
__device__ __forceinline__ void visit_stencil(const DevParams& P, const CellView& cv, const Stencil& st,
                                              int lane, int G, Fn&& fn) {
    const bool ycontig = (st.ys[0] >= 0) && (st.ys[2] < P.nby) && (st.ys[1] == st.ys[0] + 1) && (st.ys[2] == st.ys[1] + 1);
    for (int a = 0; a < st.nxs; ++a) {
        const int cx = st.xs[a];
        if (cx < 0 || cx >= P.nbx) continue;
        const int base = cx * P.nby;
        if (ycontig) {
            const int b = cv.start[base + st.ys[0]], e = cv.start[base + st.ys[2] + 1];
            for (int s = b + lane; s < e; s += G) fn(s);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int cy = st.ys[c];
                if (cy < 0 || cy >= P.nby) continue;
                const int b = cv.start[base + cy], e = cv.start[base + cy + 1];
                for (int s = b + lane; s < e; s += G) fn(s);
            }
        }
    }
}

// number of atoms in the stencil (no lane split)
__device__ __forceinline__ int stencil_population(const DevParams& P, const CellView& cv, const Stencil& st) {
    int m = 0;
    for (int a = 0; a < st.nxs; ++a) {
        const int cx = st.xs[a];
        if (cx < 0 || cx >= P.nbx) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int cy = st.ys[c];
            if (cy < 0 || cy >= P.nby) continue;
            m += cv.start[cx * P.nby + cy + 1] - cv.start[cx * P.nby + cy];
        }
    }
    return m;
}

// 1/x to ~1 ulp in three dependent FMAs: MUFU.RCP64H seed y0 (relative error e0 ~ 2^-23), then the cubically
// convergent step y0*(1 + e0 + e0^2), truncation error e0^3 ~ 2^-69 (x is a squared distance: finite, > 0 where used)
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}

// ---- Lennard-Jones pair terms on r^2 (Energy.py:24,28 and :55-59) ------------------------------
// E*((r0/r)^12 - 2 (r0/r)^6) = E*q6*(q6-2) with q6 = (r0^2/r^2)^3 ; zero when r == 0.
__device__ __forceinline__ double lj_energy(double r2, double E, double r0sq) {
    if (!(r2 > 0.0)) return 0.0;
    const double q2 = r0sq / r2;
    const double q6 = q2 * q2 * q2;
    return E * (q6 * (q6 - 2.0));
}

// force on i from j: (d/r)*E*(-12 r0^12/r^13 + 12 r0^6/r^7) = d * 12 E (q6 - q12)/r^2 ; pairs with
// r < 1e-2 are skipped (Energy.py:55), i.e. r^2 < 1e-4.
__device__ __forceinline__ double lj_force_coef(double r2, double E, double r0sq) {
    if (r2 < 1e-4) return 0.0;
    const double inv = fast_rcp(r2);
    const double q2 = r0sq * inv;
    const double q6 = q2 * q2 * q2;
    return 12.0 * E * (q6 - q6 * q6) * inv;
}

// fp32 pair terms (precision 1): the min-image displacement is formed in fp64, then the LJ term is evaluated
// in single precision (relative error ~1e-6 per term) and accumulated by the caller in fp64.
__device__ __forceinline__ double lj_energy_f32(double dx, double dy, double E, double r0sq) {
    const float fx = (float)dx, fy = (float)dy;
    const float r2 = fx * fx + fy * fy;
    if (!(r2 > 0.0f)) return 0.0;
    const float q2 = __fdividef((float)r0sq, r2);
    const float q6 = q2 * q2 * q2;
    return (double)((float)E * (q6 * (q6 - 2.0f)));
}
__device__ __forceinline__ float lj_force_coef_f32(float r2, float E, float r0sq) {
    if (r2 < 1e-4f) return 0.0f;
    const float inv = __fdividef(1.0f, r2);
    const float q2 = r0sq * inv;
    const float q6 = q2 * q2 * q2;
    return 12.0f * E * (q6 - q6 * q6) * inv;
}

// exact squared distance the way the reference forms it (Periodic.py:58-59: d*=d ; d[2j]+d[2j+1]),
// without FMA contraction, so that sqrt(r2) < 1.2 r compares exactly like the reference.
__device__ __forceinline__ double dist2_exact(const DevParams& P, double xi, double yi, double xj, double yj) {
    const double dx = put_in_box(__dadd_rn(xj, -xi), P.L, P.halfL);
    const double dy = __dadd_rn(yj, -yi);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// ---- reductions ------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ int group_sum_i(int v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ double group_min(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// deterministic block-wide sum (fixed tree); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 doubles */) {
    v = group_sum<32>(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) smem[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = (l < nw) ? smem[l] : 0.0;
        r = group_sum<32>(r);
    }
    return r;
}
__device__ __forceinline__ double block_max(double v, double* smem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) smem[w] = v;
    __syncthreads();
    double r = -INFINITY;
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = (l < nw) ? smem[l] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    }
    return r;
}
__device__ __forceinline__ double block_min(double v, double* smem) { return -block_max(-v, smem); }

// ---- pair terms shared by the sweep kernels, the event kernels and the relaxation kernels ----------------
// accumulate the force of neighbour (xj,yj) on (xi,yi)
__device__ __forceinline__ void acc_force(const DevParams& P, double xi, double yi, double2 pj, double E, double r0sq,
                                          double& fx, double& fy) {
    const double dx = put_in_box(pj.x - xi, P.L, P.halfL);
    const double dy = pj.y - yi;
    if (P.precision == 1) {
        const float ex = (float)dx, ey = (float)dy;
        const float c = lj_force_coef_f32(ex * ex + ey * ey, (float)E, (float)r0sq);
        fx += (double)(ex * c);
        fy += (double)(ey * c);
        return;
    }
    const double r2 = dx * dx + dy * dy;
    const double c = lj_force_coef(r2, E, r0sq);
    fx += dx * c;
    fy += dy * c;
}

// Forces of the slots [b, e) of a cell-sorted list on (xi, yi); the lanes of a group take every G-th slot in the same
// order as a plain `for (s = b + lane; s < e; s += G) acc_force(...)` loop (identical sums), but four neighbour positions
// are requested before their arithmetic and the four pair chains are independent and branch-free (a slot past the end
// holds a clamped duplicate and contributes a zero coefficient).  |dx| >= L and fp32 pair terms take the general form.
__device__ __forceinline__ void range_forces4(const DevParams& P, const double2* __restrict__ sxy, int b, int e, int lane, int G,
                                              double xi, double yi, double E, double r0sq, double& fx, double& fy) {
    if (G > 8 || P.precision == 1) {        // wide groups (small systems): a range holds fewer slots than 4 G lanes would take, batching only
                                            // adds padding; fp32 pair terms: the general form
        for (int s = b + lane; s < e; s += G) acc_force(P, xi, yi, sxy[s], E, r0sq, fx, fy);
        return;
    }
    const double E12 = 12.0 * E;
    for (int s = b + lane; s < e; s += 4 * G) {
        double2 p[4];
        bool far = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            p[k] = sxy[min(s + k * G, e - 1)];
            far = far || !(fabs(p[k].x - xi) < P.L);
        }
        if (far || P.precision == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (s + k * G < e) acc_force(P, xi, yi, p[k], E, r0sq, fx, fy);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double dx = p[k].x - xi;
                dx = (dx < 0.0) ? __dadd_rn(dx, P.L) : dx;
                dx = (dx > P.halfL) ? __dadd_rn(dx, -P.L) : dx;
                const double dy = p[k].y - yi;
                const double r2 = dx * dx + dy * dy;
                const double inv = fast_rcp(r2);
                const double q2 = r0sq * inv;
                const double q6 = q2 * q2 * q2;
                const bool on = (s + k * G < e) && !(r2 < 1e-4);        // r < 1e-2 skipped (Energy.py:55)
                const double c = on ? E12 * (q6 - q6 * q6) * inv : 0.0;
                fx += dx * c;
                fy += dy * c;
            }
        }
    }
}

// visit_stencil order (Bins.py:76-83) with range_forces4 per contiguous range
__device__ __forceinline__ void stencil_forces4(const DevParams& P, const CellView& cv, const Stencil& st, int lane, int G,
                                                double xi, double yi, double E, double r0sq, double& fx, double& fy) {
    const bool ycontig = (st.ys[0] >= 0) && (st.ys[2] < P.nby) && (st.ys[1] == st.ys[0] + 1) && (st.ys[2] == st.ys[1] + 1);
    for (int a = 0; a < st.nxs; ++a) {
        const int cx = st.xs[a];
        if (cx < 0 || cx >= P.nbx) continue;
        const int base = cx * P.nby;
        if (ycontig) {
            range_forces4(P, cv.sxy, cv.start[base + st.ys[0]], cv.start[base + st.ys[2] + 1], lane, G, xi, yi, E, r0sq, fx, fy);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int cy = st.ys[c];
                if (cy < 0 || cy >= P.nby) continue;
                range_forces4(P, cv.sxy, cv.start[base + cy], cv.start[base + cy + 1], lane, G, xi, yi, E, r0sq, fx, fy);
            }
        }
    }
}

__device__ __forceinline__ double pair_energy(const DevParams& P, double xi, double yi, double2 pj, double E, double r0sq) {
    const double dx = put_in_box(pj.x - xi, P.L, P.halfL);
    const double dy = pj.y - yi;
    if (P.precision == 1) return lj_energy_f32(dx, dy, E, r0sq);
    return lj_energy(dx * dx + dy * dy, E, r0sq);
}

// ascending sort of a short index list (cell contents keep the reference's input order, Bins.py:40-41):
// insertion sort for the usual ~10 atoms, heapsort beyond 32
__device__ inline void sort_indices(int* a, int m) {
    if (m <= 32) {
        for (int i = 1; i < m; ++i) {
            int v = a[i], j = i - 1;
            while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
            a[j + 1] = v;
        }
        return;
    }
    for (int start = m / 2 - 1; start >= 0; --start) {           // heapify
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= m) break;
            if (child + 1 < m && a[child] < a[child + 1]) ++child;
            if (a[root] >= a[child]) break;
            int t = a[root]; a[root] = a[child]; a[child] = t;
            root = child;
        }
    }
    for (int end = m - 1; end > 0; --end) {
        int t = a[0]; a[0] = a[end]; a[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child] < a[child + 1]) ++child;
            if (a[root] >= a[child]) break;
            int t2 = a[root]; a[root] = a[child]; a[child] = t2;
            root = child;
        }
    }
}

}  // namespace kmc
