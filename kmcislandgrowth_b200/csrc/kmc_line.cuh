// kmc_line.cuh -- the line-search engine: energies of the 20 candidates x + a*s/20/scale
// (2DGrowth.py:121-122,152-153) from ONE cell list, every pair evaluated once with its 20 candidate
// energies held in registers.
//
// Exactness.  RelaxEnergy (Energy.py:133-148) is a pure function of the candidate positions: the
// substrate stencil of an adatom -- and in bins3x3 mode the adatom pairs that are counted -- depend on
// the CELLS of the candidate positions.  Along one line x + a*delta every coordinate is monotone in a,
// so an atom whose raw cell is the same (and inside the grid) at a = 0 and a = K-1 keeps it for every
// candidate.  Those atoms take the fast path below (stencil membership fixed by the base cell list);
// the others ("movers", typically none to a few dozen) are evaluated per candidate with their own
// stencils by k_line_fix, reproducing the per-candidate semantics exactly, including atoms that leave
// the addressable grid and become invisible to others (Bins.py:24-42,57-65).
#pragma once
#include "kmc_device.cuh"

namespace kmc {

constexpr int KC = 20;        // candidates per line search: range(20) at 2DGrowth.py:121,152
constexpr int LT_HOME = 64;   // home atoms staged per pass
constexpr int LT_TILE = 256;  // neighbour atoms staged per pass
constexpr int LT_THREADS = 128;

// candidate a of coordinate v moving along s: v + ((a*s)/20)/scale, the reference's operation order
// Both quotients are formed without the division subroutine and are still the correctly rounded IEEE results:
//  x/20: q0 = RN(x*y) with y = RN(1/20), r = x - 20 q0 (exact in an FMA), q = RN(q0 + r*y) -- Markstein's correction, which
//        yields the correctly rounded quotient when y is the correctly rounded reciprocal (20 = 1.25*2^4: no exceptional
//        mantissa) and no intermediate is subnormal; operands outside [1e-280, 1e280] take the real division;
//  q/scale: every scale the relaxation uses is a power of two (16 * 2^k), for which multiplying by the exact reciprocal IS the
//        quotient; any other scale takes the real division.
__device__ __forceinline__ double cand1(double v, double s, int a, double scale) {
    const double x = __dmul_rn((double)a, s);
    double q;
    const double ax = fabs(x);
    if (ax < 1e280 && (ax > 1e-280 || ax == 0.0)) {
        const double y = 0.05;
        const double q0 = __dmul_rn(x, y);
        const double r = __fma_rn(-20.0, q0, x);
        q = __fma_rn(r, y, q0);
    } else {
        q = __ddiv_rn(x, 20.0);
    }
    const long long sb = __double_as_longlong(scale);
    const bool pow2 = (sb & 0x000fffffffffffffLL) == 0 && scale > 1e-300 && scale < 1e300;
    const double t = pow2 ? __dmul_rn(q, __ddiv_rn(1.0, scale)) : __ddiv_rn(q, scale);
    return __dadd_rn(v, t);
}
__device__ __forceinline__ double2 cand2(double2 p, double2 s, int a, double scale) {
    return make_double2(cand1(p.x, s.x, a, scale), cand1(p.y, s.y, a, scale));
}

// energies of one pair for all KC candidates: displacement d0 + a*dd (d0 already min-image wrapped).
// Stable pairs (the wrap cannot change along the line) use r^2(a) = A + a(B + aC).
__device__ __forceinline__ void pair_line(const DevParams& P, double dx0, double dy0, double ddx, double ddy, double E,
                                          double r0sq, double (&acc)[KC]) {
    const double reach = fabs(dx0) + (KC - 1) * fabs(ddx);
    if (reach < P.halfL - 1e-9) {
        const double A = dx0 * dx0 + dy0 * dy0;
        const double B = 2.0 * (dx0 * ddx + dy0 * ddy);
        const double C = ddx * ddx + ddy * ddy;
#pragma unroll
        for (int a = 0; a < KC; ++a) {
            const double r2 = fma((double)a, fma((double)a, C, B), A);
            const double q2 = r0sq * fast_rcp(r2);
            const double q6 = q2 * q2 * q2;
            const double e = q6 * (q6 - 2.0);
            acc[a] = (r2 > 0.0) ? fma(E, e, acc[a]) : acc[a];
        }
    } else {        // the pair straddles +-L/2 somewhere along the line: wrap every candidate (Periodic.py:8-20)
#pragma unroll
        for (int a = 0; a < KC; ++a) {
            const double dx = put_in_box(fma((double)a, ddx, dx0), P.L, P.halfL);
            const double dy = fma((double)a, ddy, dy0);
            acc[a] += lj_energy(dx * dx + dy * dy, E, r0sq);
        }
    }
}

// the same for the NC candidates a0 .. a0+NC-1 only (two lanes share one pair in the persistent kernel:
// half the accumulators per lane, twice the warps per SM)
template <int NC>
__device__ __forceinline__ void pair_line_part(const DevParams& P, double dx0, double dy0, double ddx, double ddy, double E,
                                               double r0sq, int a0, double (&acc)[NC]) {
    const double reach = fabs(dx0) + (KC - 1) * fabs(ddx);
    if (reach < P.halfL - 1e-9) {
        const double bx = fma((double)a0, ddx, dx0), by = fma((double)a0, ddy, dy0);   // displacement at candidate a0
        const double A = bx * bx + by * by;
        const double B = 2.0 * (bx * ddx + by * ddy);
        const double C = ddx * ddx + ddy * ddy;
        // |d(a)| >= |d0| - (KC-1)|dd|: when that bound is positive no candidate can have r == 0 and the r > 0 test of
        // Energy.py:28 is decided once per pair instead of once per candidate
        if (A > (double)((KC - 1) * (KC - 1)) * C * 1.0000001 + 1e-30) {
            // (r0/r)^2 = 1 / (r^2/r0^2): scaling the three coefficients once per pair saves a multiply per candidate
            const double inv0 = (r0sq == P.ra2) ? P.inv_ra2 : P.inv_ras2;
            const double As = A * inv0, Bs = B * inv0, Cs = C * inv0;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const double q2 = fast_rcp(fma((double)k, fma((double)k, Cs, Bs), As));
                const double q6 = q2 * q2 * q2;
                acc[k] = fma(E, q6 * (q6 - 2.0), acc[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const double r2 = fma((double)k, fma((double)k, C, B), A);
                const double q2 = r0sq * fast_rcp(r2);
                const double q6 = q2 * q2 * q2;
                const double e = q6 * (q6 - 2.0);
                acc[k] = (r2 > 0.0) ? fma(E, e, acc[k]) : acc[k];
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double dx = put_in_box(fma((double)(a0 + k), ddx, dx0), P.L, P.halfL);
            const double dy = fma((double)(a0 + k), ddy, dy0);
            acc[k] += lj_energy(dx * dx + dy * dy, E, r0sq);
        }
    }
}

// ---- Taylor form of the line energy ------------------------------------------------------------------
// Along a line the squared distance of a pair is r^2(t) = A (1 + u(t)), u = b t + c t^2, with t = a - 9.5 measured
// from the MIDDLE candidate, so its energy E (q^6 (1+u)^-6 - 2 q^3 (1+u)^-3), q = r0^2/A, is a power series in u
// whose term u^k is O(w^k), w = relative change of r along half the line.  In the relaxation's steady state w is
// 1e-6..1e-3 for all but a handful of pairs, so the degree-PD polynomial in t is exact to rounding (the first
// dropped term is 792 u^7 < 2e-15 for |u| < POLY_UMAX) and the pair costs ~90 FP64 instructions instead of
// 13 per candidate x 20; the PD+1 coefficients are summed over pairs and the 20 candidates are evaluated ONCE.
// Side effect: the candidate DIFFERENCES are carried by the small coefficients 1..PD, i.e. they no longer
// drown in the rounding of the ~6e4 eV total.  Pairs that move more (|u| >= POLY_UMAX), that can cross the +-L/2
// seam along the line, or that can reach r = 0 are not expanded: the caller evaluates them per candidate
// (pair_line_one), exactly as before.
constexpr int PD = 6;
constexpr double POLY_UMAX = 3e-3;

constexpr double POLY_UMAX3 = 5e-5;    // below this a degree-3 polynomial is exact to rounding (first dropped term 126 u^4 < 1e-15)

// per-pair quantities of the expansion; ok == false: the pair is not expanded (invalid lane, or it must take the
// per-candidate path) and every coefficient below is zero, so accumulating it is harmless (branch-free: the chains of the
// two pairs a lane holds interleave, no basic-block boundary between them)
struct PolyPrep {
    double b, c, Q3, Q6;
    bool ok, small;
};
__device__ __forceinline__ PolyPrep pair_line_prep(const DevParams& P, bool valid, double dx0, double dy0, double ddx, double ddy, double E, double inv0) {
    constexpr double tm = 0.5 * (KC - 1);
    PolyPrep r;
    const double mx = fma(tm, ddx, dx0), my = fma(tm, ddy, dy0);      // displacement at the middle of the line
    const double A = fma(mx, mx, my * my);
    const double Bh = fma(mx, ddx, my * ddy);                         // B/2
    const double C = fma(ddx, ddx, ddy * ddy);
    const bool no_wrap = fabs(dx0) + (KC - 1) * fabs(ddx) < P.halfL - 1e-9;
    const double umax_A = fma(C, tm * tm, 2.0 * tm * fabs(Bh));       // max |u| * A over the line
    const bool quiet = umax_A < POLY_UMAX * A;                        // strict: A == 0 is never quiet
    r.ok = valid && no_wrap && quiet;
    r.small = umax_A < POLY_UMAX3 * A;
    const double q = fast_rcp(r.ok ? A * inv0 : 1.0);        // r0^2 / A
    const double w = q * inv0;                               // 1 / A
    r.b = r.ok ? (Bh + Bh) * w : 0.0;
    r.c = r.ok ? C * w : 0.0;
    const double q3 = q * q * q;
    r.Q3 = r.ok ? E * q3 : 0.0;
    r.Q6 = r.Q3 * q3;
    return r;
}

// e(u) = sum_k g_k u^k,  g_k = (-1)^k (C(k+5,5) Q6 - 2 C(k+2,2) Q3); composition G(u(t)) by Horner in u, every
// intermediate truncated to the degree that can still reach t^DEG: P_k = g_k + u P_{k+1}, (u p)[m] = b p[m-1] + c p[m-2]
template <int DEG>
__device__ __forceinline__ void pair_line_acc(const PolyPrep& r, double (&acc)[PD + 1]) {
    const double b = r.b, c = r.c, Q3 = r.Q3, Q6 = r.Q6;
    const double g0 = fma(1.0, Q6, -2.0 * Q3);
    const double g1 = fma(-6.0, Q6, 6.0 * Q3);
    const double g2 = fma(21.0, Q6, -12.0 * Q3);
    const double g3 = fma(-56.0, Q6, 20.0 * Q3);
    if (DEG == 3) {
        double p0 = g3, p1, p2;
        p1 = b * p0;                                  p0 = g2;       // P2: degree 1
        p2 = fma(b, p1, c * p0); p1 = b * p0;         p0 = g1;       // P1: degree 2
        acc[3] = fma(b, p2, fma(c, p1, acc[3]));
        acc[2] = fma(b, p1, fma(c, p0, acc[2]));
        acc[1] = fma(b, p0, acc[1]);
        acc[0] += g0;
        return;
    }
    const double g4 = fma(126.0, Q6, -30.0 * Q3);
    const double g5 = fma(-252.0, Q6, 42.0 * Q3);
    const double g6 = fma(462.0, Q6, -56.0 * Q3);
    double p0 = g6, p1, p2, p3, p4, p5;
    p1 = b * p0;                                                  p0 = g5;       // P5: degree 1
    p2 = fma(b, p1, c * p0); p1 = b * p0;                         p0 = g4;       // P4: degree 2
    p3 = fma(b, p2, c * p1); p2 = fma(b, p1, c * p0); p1 = b * p0; p0 = g3;      // P3
    p4 = fma(b, p3, c * p2); p3 = fma(b, p2, c * p1); p2 = fma(b, p1, c * p0); p1 = b * p0; p0 = g2;   // P2
    p5 = fma(b, p4, c * p3); p4 = fma(b, p3, c * p2); p3 = fma(b, p2, c * p1); p2 = fma(b, p1, c * p0); p1 = b * p0; p0 = g1;   // P1
    // P0 = g0 + u P1, accumulated directly
    acc[6] = fma(b, p5, fma(c, p4, acc[6]));
    acc[5] = fma(b, p4, fma(c, p3, acc[5]));
    acc[4] = fma(b, p3, fma(c, p2, acc[4]));
    acc[3] = fma(b, p2, fma(c, p1, acc[3]));
    acc[2] = fma(b, p1, fma(c, p0, acc[2]));
    acc[1] = fma(b, p0, acc[1]);
    acc[0] += g0;
}

// one pair; returns false when it was not expanded.  The degree is chosen per WARP: when every expanded pair of the round
// moves by |u| < POLY_UMAX3 (the usual case in the long tail of a relaxation) three powers of u are exact to rounding.
__device__ __forceinline__ bool pair_line_poly(const DevParams& P, bool valid, double dx0, double dy0, double ddx, double ddy, double E,
                                               double inv0, double (&acc)[PD + 1]) {
    const PolyPrep r = pair_line_prep(P, valid, dx0, dy0, ddx, ddy, E, inv0);
    if (__all_sync(0xffffffffu, !r.ok || r.small)) pair_line_acc<3>(r, acc);
    else pair_line_acc<PD>(r, acc);
    return r.ok;
}

// one candidate of one pair, general form (min-image wrap per candidate, r == 0 excluded: Periodic.py:8-20, Energy.py:28)
__device__ __forceinline__ double pair_line_one(const DevParams& P, double dx0, double dy0, double ddx, double ddy, double E,
                                                double r0sq, int a) {
    const double dx = put_in_box(fma((double)a, ddx, dx0), P.L, P.halfL);
    const double dy = fma((double)a, ddy, dy0);
    return lj_energy(dx * dx + dy * dy, E, r0sq);
}

// value of the summed polynomial at candidate a
__device__ __forceinline__ double poly_at(const double* cf, int a) {
    const double t = (double)a - 0.5 * (KC - 1);
    double v = cf[PD];
#pragma unroll
    for (int m = PD - 1; m >= 0; --m) v = fma(v, t, cf[m]);
    return v;
}

struct LineArgs {
    DevParams P;
    CellView ad;                 // base cell list of x (ad.sxy[slot] == x[ad.sidx[slot]])
    CellView sub;
    const double2* x;            // original order
    const double2* s;            // original order
    double scale;
    double inv_step;             // 1/(20*scale): delta = s * inv_step (fast path only)
    int n;
    uint8_t* flag;               // [n] 1 = mover
    int* movers;                 // [n] mover atom indices, ascending
    int* n_movers;               // [1]
    double* partial;             // [nblocks_energy][KC]
    double* fix;                 // [n_movers][KC]
    double* e_out;               // [KC]
    int* amin_out;               // first argmin (ys.index(min(ys)))
    const int2* colrange;        // optional [nbx]: lowest / highest occupied cell row of every column of `ad` (x > y: column empty); lets
                                 // the home-cell loop skip the empty part of a tall grid (configs[4]: 1186 rows, ~6 occupied)
    int n_compute;               // domain decomposition: atoms with index < n_compute are the rank's own, the rest halo copies; a pair
                                 // counts with weight (own_i + own_j)/2, a substrate term with weight own_i (-1: every atom is own)
};

// ---- phase M: movers ---------------------------------------------------------------------------
__device__ __forceinline__ void line_mark_movers(const LineArgs& A, int gtid, int gthreads) {
    for (int i = gtid; i < A.n; i += gthreads) {
        const double2 p0 = A.x[i];
        const double2 pe = cand2(p0, A.s[i], KC - 1, A.scale);
        int nx0, ny0, nx1, ny1;
        bin_index(A.P, p0.x, p0.y, nx0, ny0);
        bin_index(A.P, pe.x, pe.y, nx1, ny1);
        const bool in0 = nx0 >= 0 && nx0 < A.P.nbx && ny0 >= 0 && ny0 < A.P.nby;
        // domain decomposition: a halo copy that starts outside the grid (an empty halo slot, or an atom nobody can see)
        // takes no part at all -- flag 2: neither paired on the fast path nor evaluated as a mover
        A.flag[i] = (A.n_compute >= 0 && i >= A.n_compute && !in0) ? 2 : ((in0 && nx0 == nx1 && ny0 == ny1) ? 0 : 1);
    }
}

// ---- phase C: ordered compaction of the flags (one block) ---------------------------------------
__device__ __forceinline__ void line_compact_movers(const LineArgs& A, int* s_scan /* blockDim ints */) {
    const int T = blockDim.x, t = threadIdx.x;
    const int per = (A.n + T - 1) / T;
    const int lo = t * per, hi = min(lo + per, A.n);
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += (A.flag[i] == 1);
    s_scan[t] = cnt;
    __syncthreads();
    for (int off = 1; off < T; off <<= 1) {
        int v = (t >= off) ? s_scan[t - off] : 0;
        __syncthreads();
        s_scan[t] += v;
        __syncthreads();
    }
    int pos = s_scan[t] - cnt;
    for (int i = lo; i < hi; ++i)
        if (A.flag[i] == 1) A.movers[pos++] = i;
    if (t == T - 1) *A.n_movers = s_scan[t];
    __syncthreads();
}

// ---- phase E: fast path over home cells ----------------------------------------------------------
struct LineSmem {
    double hx[LT_HOME], hy[LT_HOME], hsx[LT_HOME], hsy[LT_HOME];
    int hslot[LT_HOME];                 // slot of the home atom, INT_MAX for movers (never paired here)
    float hown[LT_HOME], town[LT_TILE]; // ownership (1 / 0) of the staged home / partner atoms (domain decomposition)
    double tx[LT_TILE], ty[LT_TILE], tsx[LT_TILE], tsy[LT_TILE];
    int tslot[LT_TILE];                 // same-cell partner: its slot (pair iff hslot < tslot); other cells: INT_MAX; movers: -1
    double red[32];
};

// process the staged tile against the staged home atoms
__device__ __forceinline__ void line_tile_pairs(const LineArgs& A, LineSmem& S, int nh, int nt, double E, double r0sq,
                                                double (&acc)[KC], bool substrate = false) {
    const int total = nh * nt;
    for (int p = threadIdx.x; p < total; p += blockDim.x) {
        const int i = p % nh, j = p / nh;
        if (S.hslot[i] >= S.tslot[j]) continue;
        double Ew = E;
        if (A.n_compute >= 0) {          // share of this rank in the pair (own-own 1, own-halo 1/2, halo-halo 0; substrate: own 1)
            const double w = substrate ? (double)S.hown[i] : 0.5 * (double)(S.hown[i] + S.town[j]);
            if (w == 0.0) continue;
            Ew = E * w;
        }
        const double dx0 = put_in_box(S.tx[j] - S.hx[i], A.P.L, A.P.halfL);
        const double dy0 = S.ty[j] - S.hy[i];
        pair_line(A.P, dx0, dy0, S.tsx[j] - S.hsx[i], S.tsy[j] - S.hsy[i], Ew, r0sq, acc);
    }
}

__device__ __forceinline__ void line_energy_cells(const LineArgs& A, LineSmem& S, int first_cell, int cell_stride,
                                                  double (&acc)[KC]) {
    const DevParams& P = A.P;
    // home cells of this block: every cell_stride-th cell of the grid, or -- with the occupied row range of every column
    // known -- every cell_stride-th COLUMN, rows [lowest, highest] only.  Both orders are functions of the data alone.
    int col = first_cell, row = 0, row_hi = -1;
    int c = first_cell - cell_stride;
    for (;;) {
        if (A.colrange) {
            if (row > row_hi) {
                bool found = false;
                for (; col < P.nbx; col += cell_stride) {
                    const int2 r = A.colrange[col];
                    if (r.x <= r.y) { row = r.x; row_hi = r.y; found = true; break; }
                }
                if (!found) break;
                c = col * P.nby + row;
                col += cell_stride;
            } else {
                c += 1;
            }
            ++row;
        } else {
            c += cell_stride;
            if (c >= P.ncell) break;
        }
        const int hb = A.ad.start[c], he = A.ad.start[c + 1];
        if (hb == he) continue;
        const int cx = c / P.nby, cy = c % P.nby;
        // the stencil of an in-grid atom is a function of its cell: same code path as every other kernel
        Stencil st;
        {
            int ys[3], xs[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                int vx = cx - 1 + i, vy = cy - 1 + i;
                if (vx < 0) vx += P.nbx;
                if (vx >= P.nbx) vx -= P.nbx;
                if (vy < 0) vy += P.nby;
                if (vy >= P.nby) vy -= P.nby;
                xs[i] = vx; ys[i] = vy;
            }
            st.xs[0] = xs[0]; st.xs[1] = xs[1]; st.xs[2] = xs[2]; st.xs[3] = -1;
            st.ys[0] = ys[0]; st.ys[1] = ys[1]; st.ys[2] = ys[2];
            st.nxs = 3;
            if (cx == 0) { st.xs[3] = P.nbx - 2; st.nxs = 4; }
            else if (cx == P.nbx - 2) { st.xs[3] = 0; st.nxs = 4; }
        }
        for (int h0 = hb; h0 < he; h0 += LT_HOME) {
            const int nh = min(LT_HOME, he - h0);
            __syncthreads();
            for (int i = threadIdx.x; i < nh; i += blockDim.x) {
                const int slot = h0 + i, idx = A.ad.sidx[slot];
                const double2 p = A.ad.sxy[slot], s = A.s[idx];
                S.hx[i] = p.x; S.hy[i] = p.y;
                S.hsx[i] = s.x * A.inv_step; S.hsy[i] = s.y * A.inv_step;
                S.hslot[i] = A.flag[idx] ? 0x7fffffff : slot;
                S.hown[i] = (A.n_compute < 0 || idx < A.n_compute) ? 1.0f : 0.0f;
            }
            // ---- adatom partners (bins3x3 mode): cells c' >= c of the stencil, with multiplicity
            if (P.aa_mode != 0) {
                for (int a = 0; a < st.nxs; ++a) {
                    const int nx = st.xs[a];
                    if (nx < 0 || nx >= P.nbx) continue;
                    for (int b = 0; b < 3; ++b) {
                        const int ny = st.ys[b];
                        if (ny < 0 || ny >= P.nby) continue;
                        const int c2 = nx * P.nby + ny;
                        if (c2 < c) continue;
                        const int nb = A.ad.start[c2], ne = A.ad.start[c2 + 1];
                        for (int t0 = nb; t0 < ne; t0 += LT_TILE) {
                            const int nt = min(LT_TILE, ne - t0);
                            __syncthreads();
                            for (int j = threadIdx.x; j < nt; j += blockDim.x) {
                                const int slot = t0 + j, idx = A.ad.sidx[slot];
                                const double2 p = A.ad.sxy[slot], s = A.s[idx];
                                S.tx[j] = p.x; S.ty[j] = p.y;
                                S.tsx[j] = s.x * A.inv_step; S.tsy[j] = s.y * A.inv_step;
                                S.tslot[j] = A.flag[idx] ? -1 : (c2 == c ? slot : 0x7fffffff);
                                S.town[j] = (A.n_compute < 0 || idx < A.n_compute) ? 1.0f : 0.0f;
                            }
                            __syncthreads();
                            line_tile_pairs(A, S, nh, nt, P.E_a, P.ra2, acc);
                        }
                    }
                }
            }
            // ---- substrate partners: every cell of the stencil (static atoms: direction 0)
            for (int a = 0; a < st.nxs; ++a) {
                const int nx = st.xs[a];
                if (nx < 0 || nx >= P.nbx) continue;
                for (int b = 0; b < 3; ++b) {
                    const int ny = st.ys[b];
                    if (ny < 0 || ny >= P.nby) continue;
                    const int c2 = nx * P.nby + ny;
                    const int nb = A.sub.start[c2], ne = A.sub.start[c2 + 1];
                    for (int t0 = nb; t0 < ne; t0 += LT_TILE) {
                        const int nt = min(LT_TILE, ne - t0);
                        __syncthreads();
                        for (int j = threadIdx.x; j < nt; j += blockDim.x) {
                            const double2 p = A.sub.sxy[t0 + j];
                            S.tx[j] = p.x; S.ty[j] = p.y; S.tsx[j] = 0.0; S.tsy[j] = 0.0;
                            S.tslot[j] = 0x7ffffffe;       // pairs with every non-mover home atom... see below
                        }
                        __syncthreads();
                        // movers carry hslot = INT_MAX > 0x7ffffffe, so they are skipped here as well
                        line_tile_pairs(A, S, nh, nt, P.E_as, P.ras2, acc, true);
                    }
                }
            }
        }
    }
}

// all-pairs adatom sums (aa_mode 0, Energy.py:115-117): every pair i<j is always counted, only the
// min-image wrap can change along the line (handled inside pair_line)
__device__ __forceinline__ void line_energy_allpairs(const LineArgs& A, int gwarp, int nwarps, int lane, double (&acc)[KC]) {
    for (int i = gwarp; i < A.n; i += nwarps) {
        const double2 pi = A.x[i], si = A.s[i];
        for (int j = i + 1 + lane; j < A.n; j += 32) {
            const double2 pj = A.x[j], sj = A.s[j];
            const double dx0 = put_in_box(pj.x - pi.x, A.P.L, A.P.halfL);
            pair_line(A.P, dx0, pj.y - pi.y, (sj.x - si.x) * A.inv_step, (sj.y - si.y) * A.inv_step, A.P.E_a, A.P.ra2, acc);
        }
    }
}

// block-reduce the KC accumulators and store them as this block's partial
__device__ __forceinline__ void line_store_partial(double (&acc)[KC], double* red, double* out) {
#pragma unroll
    for (int a = 0; a < KC; ++a) {
        const double v = block_sum(acc[a], red);
        if (threadIdx.x == 0) out[a] = v;
    }
}

// ---- phase F: movers, one warp per (mover, candidate) ------------------------------------------
__device__ __forceinline__ int stencil_multiplicity(const DevParams& P, const Stencil& st, int cx, int cy) {
    int mx = 0, my = 0;
    for (int a = 0; a < st.nxs; ++a) mx += (st.xs[a] == cx);
#pragma unroll
    for (int b = 0; b < 3; ++b) my += (st.ys[b] == cy);
    return mx * my;
}

__device__ __forceinline__ double line_fix_one(const LineArgs& A, int r, int a, int lane) {
    const DevParams& P = A.P;
    const int nm = *A.n_movers;
    const int m = A.movers[r];
    const double2 pm = cand2(A.x[m], A.s[m], a, A.scale);
    int rx, ry;
    bin_index(P, pm.x, pm.y, rx, ry);
    const bool in_grid = rx >= 0 && rx < P.nbx && ry >= 0 && ry < P.nby;
    const Stencil st = make_stencil(P, pm.x, pm.y);
    double e = 0.0;
    const bool dd = A.n_compute >= 0;
    const double own_m = (!dd || m < A.n_compute) ? 1.0 : 0.0;
    if (P.aa_mode != 0) {
        // non-mover partners keep their base cell; an in-grid mover both sees them and is seen by them
        // (symmetric stencil relation), an out-of-grid mover only sees partners with a higher index
        visit_stencil(P, A.ad, st, lane, 32, [&](int slot) {
            const int j = A.ad.sidx[slot];
            if (A.flag[j] || (!in_grid && j < m)) return;
            const double w = dd ? 0.5 * (own_m + (j < A.n_compute ? 1.0 : 0.0)) : 1.0;
            if (w == 0.0) return;
            const double2 pj = cand2(A.x[j], A.s[j], a, A.scale);
            e += w * pair_energy(P, pm.x, pm.y, pj, P.E_a, P.ra2);
        });
        // mover - mover: the lower index sees the higher one if the latter's cell is in its stencil
        for (int r2 = r + 1 + lane; r2 < nm; r2 += 32) {
            const int m2 = A.movers[r2];
            const double2 p2 = cand2(A.x[m2], A.s[m2], a, A.scale);
            int cx, cy;
            bin_index(P, p2.x, p2.y, cx, cy);
            if (cx < 0 || cx >= P.nbx || cy < 0 || cy >= P.nby) continue;
            const int mult = stencil_multiplicity(P, st, cx, cy);
            const double w = dd ? 0.5 * (own_m + (m2 < A.n_compute ? 1.0 : 0.0)) : 1.0;
            if (mult) e += w * mult * pair_energy(P, pm.x, pm.y, p2, P.E_a, P.ra2);
        }
    }
    if (own_m != 0.0)
        visit_stencil(P, A.sub, st, lane, 32, [&](int slot) { e += pair_energy(P, pm.x, pm.y, A.sub.sxy[slot], P.E_as, P.ras2); });
    return group_sum<32>(e);
}

// ---- phase R: final sums + first argmin (one block of KC warps) ---------------------------------
__device__ __forceinline__ void line_reduce(const LineArgs& A, int nblocks, double* s_e /* KC doubles */) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < KC) {
        double v = 0.0;
        for (int b = lane; b < nblocks; b += 32) v += A.partial[(size_t)b * KC + w];
        const int nm = *A.n_movers;
        for (int r = lane; r < nm; r += 32) v += A.fix[(size_t)r * KC + w];
        v = group_sum<32>(v);
        if (lane == 0) { s_e[w] = v; if (A.e_out) A.e_out[w] = v; }
    }
    __syncthreads();
    if (threadIdx.x == 0 && A.amin_out) {
        int am = 0;
        double em = s_e[0];
        for (int a = 1; a < KC; ++a)
            if (s_e[a] < em) { em = s_e[a]; am = a; }
        *A.amin_out = am;
    }
}

}  // namespace kmc
