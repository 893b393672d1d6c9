// kmc_event.cu -- one iteration of the reference's KMC loop (2DGrowth.py:317-329) queued on the stream WITHOUT a host
// round trip in the middle:
//
//   rate table (cell list + k_sweep_rates) -> selection (k_select) -> DepositAdatom placement (k_ev_deposit, :64-93) or
//   HopAtom placement (k_ev_hop + k_ev_apply/k_ev_copyback, :229-255) -> Relaxation (k_relax2, :97-166) -> ONE readback.
//
// Which branch runs is decided on the device: both placement kernels are launched, each looks at the selection result
// in device memory and returns at once when the event is not its kind.  The adatom count after the placement (n or
// n + 1) therefore lives in device memory (KmcCtx::st_n_dev) and the relaxation kernel reads it at its start.  The two
// uniforms of an event (2DGrowth.py:322 and :76 / :248) are kernel arguments: the random stream stays with the caller.
//
// kmc_event_enqueue / kmc_event_finish split the call so that several replicas (one context and stream each) can be
// advanced by one host thread: kmc_events_batch queues all of them before it waits for the first (configs[3]).
#include "kmc_host.h"

#include <cmath>
#include <cstring>

using namespace kmc;

namespace {

constexpr int EV_THREADS = 1024;
constexpr int EV_WARPS = EV_THREADS / 32;
constexpr int EV_MAX_ROUNDS = 8192;       // 8192 x 32 probe heights x r_a/2 per step: far beyond any film

struct EvArgs {
    DevParams P;
    CellView ad;            // cell list of the current state (built for the rate table of this event)
    CellView sub;
    double2* xy;            // trajectory state, capacity >= n + 1
    double2* tmp;           // [n] scratch for the removal
    const uint8_t* surf;    // surface flags of the current state (this event's rate table)
    EventMail* mail;
    int* n_dev;
    int n;
    int hop_index_mode, deposit_mode;
    int who_override;       // >= 0: HopAtom of this adatom regardless of the selection (kmc_hop_place)
    int force_kind;         // -1: take the kind from the selection; 0 / 1: forced (kmc_deposit_place / kmc_hop_place)
    double u1;
    double new_x;           // PutInBox(u1 * L), 2DGrowth.py:76 (host arithmetic, identical to the reference's)
    double two_ra, two_ras, half_ra, three_ra;
    double hex_dx[6], hex_dy[6];    // cos/sin(i*pi/3) * r_a, 2DGrowth.py:249-250
};

// min distance of probe (px, py) to the stencil adatoms / substrate atoms: the closeness test of 2DGrowth.py:82-90
// (same arithmetic and visiting order as k_probe); one warp per probe
__device__ __forceinline__ void probe_distances(const EvArgs& A, double px, double py, int lane, double& da, double& ds) {
    const Stencil st = make_stencil(A.P, px, py);
    da = INFINITY; ds = INFINITY;
    if (A.ad.n > 0)
        visit_stencil(A.P, A.ad, st, lane, 32, [&](int s) {
            const double2 pj = A.ad.sxy[s];
            da = fmin(da, sqrt(dist2_exact(A.P, px, py, pj.x, pj.y)));
        });
    visit_stencil(A.P, A.sub, st, lane, 32, [&](int s) {
        const double2 pj = A.sub.sxy[s];
        ds = fmin(ds, sqrt(dist2_exact(A.P, px, py, pj.x, pj.y)));
    });
    da = group_min<32>(da);
    ds = group_min<32>(ds);
}

// 2DGrowth.py:64-93 DepositAdatom without the relaxation.  One block; every warp evaluates one probe height per round.
__global__ void __launch_bounds__(EV_THREADS) k_ev_deposit(EvArgs A) {
    __shared__ double s_red[32];
    __shared__ double s_y[EV_WARPS + 1];
    __shared__ int s_hit[EV_WARPS];
    EventMail* M = A.mail;
    const int kind = A.force_kind >= 0 ? A.force_kind : (M->res[0] >= 0 ? 1 : 0);
    if (kind != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // :73-75 min_y (deposit_mode 1: the film's top)
    double v = A.deposit_mode ? -INFINITY : INFINITY;
    for (int i = threadIdx.x; i < A.n; i += blockDim.x) v = A.deposit_mode ? fmax(v, A.xy[i].y) : fmin(v, A.xy[i].y);
    double r = A.deposit_mode ? block_max(v, s_red) : block_min(v, s_red);
    if (threadIdx.x == 0) s_y[EV_WARPS] = (A.n > 0) ? r : 0.0;
    __syncthreads();
    const double ref_y = s_y[EV_WARPS];
    double y0 = __dadd_rn(A.two_ras, ref_y);                           // :77
    int found = -1;
    double found_y = 0.0;
    for (int round = 0; round < EV_MAX_ROUNDS && found < 0; ++round) {
        // :92 new_y -= r_a/2, one subtraction per probe, sequentially rounded like the reference's loop
        double y = y0;
        for (int k = 0; k < warp; ++k) y = __dadd_rn(y, -A.half_ra);
        double da, ds;
        probe_distances(A, A.new_x, y, lane, da, ds);
        if (lane == 0) {
            s_hit[warp] = (da < A.two_ra || ds < A.two_ras) ? 1 : 0;  // :84,89
            s_y[warp] = y;
        }
        __syncthreads();
        for (int k = 0; k < EV_WARPS; ++k)
            if (s_hit[k]) { found = k; found_y = s_y[k]; break; }
        y0 = __dadd_rn(s_y[EV_WARPS - 1], -A.half_ra);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        M->kind = 0;
        M->hop_who = -1;
        M->ref_y = ref_y;
        if (found >= 0) {
            A.xy[A.n] = make_double2(A.new_x, found_y);                // :93
            M->new_xy[0] = A.new_x; M->new_xy[1] = found_y;
            M->n_new = A.n + 1;
            *A.n_dev = A.n + 1;
            M->status = 0;
        } else {
            M->n_new = A.n;
            *A.n_dev = A.n;
            M->status = KMC_E_DEPOSIT;
        }
    }
}

// HoppingEnergy (Energy.py:170-172) of probe (px, py) against the state WITHOUT adatom `who` (2DGrowth.py:241,252).
// The sums run over the cell list of the full state with the slot of `who` skipped, in exactly the order (slot -> lane)
// a cell list of the remaining atoms would give -- bit-identical to rebuilding the list after the removal.
__device__ __forceinline__ double hop_probe_energy(const EvArgs& A, double px, double py, int who_slot, int lane) {
    const DevParams& P = A.P;
    const Stencil st = make_stencil(P, px, py);
    double ea = 0.0, es = 0.0;
    const int nrest = A.n - 1;
    auto full_slot = [&](int s) { return s + (s >= who_slot ? 1 : 0); };      // slot of the reduced list -> slot of the full list
    if (nrest > 0) {
        if (P.aa_mode == 0) {
            for (int s = lane; s < nrest; s += 32) ea += pair_energy(P, px, py, A.ad.sxy[full_slot(s)], P.E_a, P.ra2);
        } else {
            const bool ycontig = (st.ys[0] >= 0) && (st.ys[2] < P.nby) && (st.ys[1] == st.ys[0] + 1) && (st.ys[2] == st.ys[1] + 1);
            for (int a = 0; a < st.nxs; ++a) {
                const int cx = st.xs[a];
                if (cx < 0 || cx >= P.nbx) continue;
                const int base = cx * P.nby;
                const int nseg = ycontig ? 1 : 3;
                for (int c = 0; c < nseg; ++c) {
                    int b, e;
                    if (ycontig) { b = A.ad.start[base + st.ys[0]]; e = A.ad.start[base + st.ys[2] + 1]; }
                    else {
                        const int cy = st.ys[c];
                        if (cy < 0 || cy >= P.nby) continue;
                        b = A.ad.start[base + cy]; e = A.ad.start[base + cy + 1];
                    }
                    b -= (b > who_slot) ? 1 : 0;          // the same range in the reduced list
                    e -= (e > who_slot) ? 1 : 0;
                    for (int s = b + lane; s < e; s += 32) {
                        const double2 pj = A.ad.sxy[full_slot(s)];
                        ea += lj_energy(dist2_exact(P, px, py, pj.x, pj.y), P.E_a, P.ra2);
                    }
                }
            }
        }
    }
    visit_stencil(P, A.sub, st, lane, 32, [&](int s) {
        const double2 pj = A.sub.sxy[s];
        es += lj_energy(dist2_exact(P, px, py, pj.x, pj.y), P.E_as, P.ras2);
    });
    ea = group_sum<32>(ea);
    es = group_sum<32>(es);
    return ea + es;                                        // Energy.py:172
}

// 2DGrowth.py:229-255 HopAtom without the relaxation: candidate list in adatom order, int(u*ncand), six hexagonal
// sites, first argmin of HoppingEnergy, wrap.  One block.  The removal itself (np.delete + np.append) is k_ev_apply.
__global__ void __launch_bounds__(EV_THREADS) k_ev_hop(EvArgs A) {
    __shared__ int s_scan[EV_THREADS];
    __shared__ double s_t[2];
    __shared__ double s_e[6];
    __shared__ int s_slot;
    EventMail* M = A.mail;
    const int kind = A.force_kind >= 0 ? A.force_kind : (M->res[0] >= 0 ? 1 : 0);
    if (kind != 1) return;
    const DevParams& P = A.P;
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int n = A.n;
    const long long who_ll = A.who_override >= 0 ? (long long)A.who_override : (A.hop_index_mode == 0 ? M->res[0] : M->res[1]);   // :327
    if (who_ll < 0 || who_ll >= n) {
        if (t == 0) { M->kind = 1; M->hop_who = (int)who_ll; M->status = KMC_E_ARG; M->n_new = n; *A.n_dev = n; }
        return;
    }
    const int who = (int)who_ll;
    const double2 j = A.xy[who];
    // :242-246 surface atoms within 3 r_a of the jumper, in adatom order (every thread owns a contiguous slice)
    const int per = (n + T - 1) / T;
    const int lo = t * per, hi = min(lo + per, n);
    int cnt = 0;
    for (int i = lo; i < hi; ++i)
        if (A.surf[i] && sqrt(dist2_exact(P, j.x, j.y, A.xy[i].x, A.xy[i].y)) < A.three_ra) ++cnt;
    s_scan[t] = cnt;
    if (t == 0) s_slot = -1;
    __syncthreads();
    for (int off = 1; off < T; off <<= 1) {
        int v = (t >= off) ? s_scan[t - off] : 0;
        __syncthreads();
        s_scan[t] += v;
        __syncthreads();
    }
    const int ncand = s_scan[T - 1];
    if (ncand == 0) {
        if (t == 0) { M->kind = 1; M->hop_who = who; M->ncand = 0; M->status = KMC_E_NOCANDIDATE; M->n_new = n; *A.n_dev = n; }
        return;
    }
    int pick = (int)(A.u1 * (double)ncand);                               // :248 randint(0, ncand-1) -> int(u*ncand)
    if (pick >= ncand) pick = ncand - 1;
    const int before = s_scan[t] - cnt;
    if (pick >= before && pick < before + cnt) {
        int seen = before;
        for (int i = lo; i < hi; ++i)
            if (A.surf[i] && sqrt(dist2_exact(P, j.x, j.y, A.xy[i].x, A.xy[i].y)) < A.three_ra) {
                if (seen == pick) { s_t[0] = A.xy[i].x; s_t[1] = A.xy[i].y; }
                ++seen;
            }
    }
    // slot of the jumper in the cell list (the list is a permutation: exactly one slot holds it)
    for (int s = t; s < n; s += T)
        if (A.ad.sidx[s] == who) s_slot = s;
    __syncthreads();
    const double tx = s_t[0], ty = s_t[1];
    if (warp < 6) {                                                       // :249-252
        const double px = __dadd_rn(tx, A.hex_dx[warp]), py = __dadd_rn(ty, A.hex_dy[warp]);
        const double e = hop_probe_energy(A, px, py, s_slot, lane);
        if (lane == 0) s_e[warp] = e;
    }
    __syncthreads();
    if (t == 0) {
        int best = 0;
        for (int i = 1; i < 6; ++i)
            if (s_e[i] < s_e[best]) best = i;                             // :253 first minimum
        const double px = put_in_box(__dadd_rn(tx, A.hex_dx[best]), P.L, P.halfL);      // :254
        const double py = __dadd_rn(ty, A.hex_dy[best]);
        M->kind = 1; M->hop_who = who; M->ncand = ncand; M->pick = pick;
        M->new_xy[0] = px; M->new_xy[1] = py;
        M->n_new = n;
        *A.n_dev = n;
        M->status = 0;
    }
}

// np.delete(adatoms, who) + np.append(adatoms, new position) (2DGrowth.py:241,255), through a scratch copy
__global__ void __launch_bounds__(256) k_ev_apply(EvArgs A) {
    const EventMail* M = A.mail;
    if (M->kind != 1 || M->status != 0) return;
    const int who = M->hop_who;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.n - 1) A.tmp[i] = A.xy[i < who ? i : i + 1];
    else if (i == A.n - 1) A.tmp[i] = make_double2(M->new_xy[0], M->new_xy[1]);
}
__global__ void __launch_bounds__(256) k_ev_copyback(EvArgs A) {
    const EventMail* M = A.mail;
    if (M->kind != 1 || M->status != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.n) A.xy[i] = A.tmp[i];
}

__global__ void k_ev_init(EventMail* M, int* n_dev, int n) {
    M->res[0] = -1; M->res[1] = -1; M->rtot = 0.0; M->kind = -1; M->status = 0; M->n_new = n; M->hop_who = -1;
    M->ref_y = 0.0; M->new_xy[0] = M->new_xy[1] = 0.0; M->ncand = 0; M->pick = 0;
    *n_dev = n;
}

double put_in_box_host(double x, double L) {       // Periodic.py:8-20 on the host (Python float modulo)
    double r = std::fmod(x, L);
    if (r != 0.0) { if (r < 0.0) r += L; } else r = 0.0;
    if (std::fabs(r) > L / 2) r -= L;
    return r;
}

constexpr size_t MAIL_EVENT_OFFSET = 1024;         // EventMail's place in the pinned mailbox (the relaxation block uses [0, 512))

// queue the placement of an event on the stream.  force_kind -1: rate table + selection decide.
int enqueue_placement(KmcCtx* c, const KmcBins* sb, double u0, double u1, int force_kind, int who_override, bool need_surf) {
    const int n = (int)c->st_n;
    double tq = kmc_now();
#define ENQ_MARK(i) { const double _t = kmc_now(); c->enq_s[i] += _t - tq; tq = _t; }
    TRY(kmc_state_reserve(c, n + 1));
    EventMail* mail;
    TRY(slot(c, S_EVMAIL, 1, &mail));
    ENQ_MARK(4)
    LAUNCH(c, "event_init", k_ev_init, 1, 1, mail, c->st_n_dev, n);
    ENQ_MARK(5)
    double* rate = nullptr;
    uint8_t* surf = nullptr;
    CellView ad;
    ad.start = nullptr; ad.sxy = nullptr; ad.sidx = nullptr; ad.n = 0;
    if (n > 0) {
        TRY(slot(c, S_EVRATE, (size_t)n, &rate));
        TRY(slot(c, S_EVSURF, (size_t)n, &surf));
        if (force_kind < 0 || need_surf) {
            // :318 the rate table of the current state (its cell list serves the placement as well)
            TRY(kmc_rates_dev(c, (const double2*)c->st_xy, n, sb, rate, surf, nullptr, nullptr, nullptr, nullptr, nullptr));
            ad.start = (int*)c->slots[S_START].p; ad.sxy = (double2*)c->slots[S_SXY].p; ad.sidx = (int*)c->slots[S_SIDX].p; ad.n = n;
        } else {
            TRY(kmc_scratch_cells(c, (const double2*)c->st_xy, n, 1, &ad));
        }
    }
    ENQ_MARK(6)
    if (force_kind < 0) TRY(kmc_select_dev(c, rate, surf, n, u0, mail->res, &mail->rtot, nullptr));      // :319-324
    ENQ_MARK(7)
    EvArgs A{};
    A.P = c->dp; A.ad = ad; A.sub = view_of(sb);
    A.xy = (double2*)c->st_xy;
    TRY(slot(c, S_EVTMP, (size_t)n + 1, &A.tmp));
    A.surf = surf; A.mail = mail; A.n_dev = c->st_n_dev; A.n = n;
    A.hop_index_mode = c->hp.hop_index_mode; A.deposit_mode = c->hp.deposit_mode;
    A.who_override = who_override; A.force_kind = force_kind;
    A.u1 = u1;
    A.new_x = put_in_box_host(u1 * c->hp.L, c->hp.L);                    // :76
    A.two_ra = 2 * c->hp.r_a; A.two_ras = 2 * c->hp.r_as; A.half_ra = c->hp.r_a / 2; A.three_ra = 3 * c->hp.r_a;
    for (int i = 0; i < 6; ++i) {                                         // :249-250
        const double t = i * M_PI / 3;
        A.hex_dx[i] = std::cos(t) * c->hp.r_a;
        A.hex_dy[i] = std::sin(t) * c->hp.r_a;
    }
    ENQ_MARK(8)
    if (force_kind != 1) LAUNCH(c, "event_deposit", k_ev_deposit, 1, EV_THREADS, A);
    ENQ_MARK(9)
    if (force_kind != 0 && n > 0) {
        LAUNCH(c, "event_hop", k_ev_hop, 1, EV_THREADS, A);
        LAUNCH(c, "event_apply", k_ev_apply, cdiv(n, 256), 256, A);
        LAUNCH(c, "event_apply", k_ev_copyback, cdiv(n, 256), 256, A);
    }
    CK(cudaGetLastError());
    ENQ_MARK(10)
    static_assert(sizeof(EventMail) % 4 == 0, "EventMail is copied word by word");
    TRY(kmc_to_mailbox_async(c, MAIL_EVENT_OFFSET, mail, sizeof(EventMail)));
    CK(cudaGetLastError());
    ENQ_MARK(11)
#undef ENQ_MARK
    return 0;
}

// after a synchronisation: take the placement's outcome from the mailbox
int collect_placement(KmcCtx* c, int32_t* kind_out, int64_t* hop_k_out, double* rtot_out) {
    const EventMail* M = (const EventMail*)((const char*)c->h_mail + MAIL_EVENT_OFFSET);
    if (kind_out) *kind_out = M->kind;
    if (hop_k_out) *hop_k_out = M->kind == 1 ? M->hop_who : -1;
    if (rtot_out) *rtot_out = M->rtot;
    const int n_before = (int)c->st_n;
    if (M->status == KMC_E_NOCANDIDATE) return kmc_fail(KMC_E_NOCANDIDATE, "HopAtom: no surface atom within 3 r_a of adatom %d", M->hop_who);
    if (M->status == KMC_E_DEPOSIT) return kmc_fail(KMC_E_DEPOSIT, "deposition probe never came within 2 r of an atom");
    if (M->status == KMC_E_ARG) return kmc_fail(KMC_E_ARG, "hop index %d out of range (n=%d)", M->hop_who, n_before);
    if (M->status != 0) return kmc_fail(M->status, "event placement failed");
    c->st_n = M->n_new;
    c->rc_valid = false;           // the placement bypassed the incremental rate cache's bookkeeping
    c->pl_valid = false;
    return 0;
}

}  // namespace

extern "C" {

int kmc_set_event_mode(KmcCtx* c, int mode) {
    if (!c || mode < 0 || mode > 1) return kmc_fail(KMC_E_ARG, "event mode must be 0 (host-stepped placement) or 1 (device-resident event)");
    c->event_mode = mode;
    return 0;
}

int kmc_event_enqueue(KmcCtx* c, const KmcBins* sb, double u0, double u1, int64_t max_ls) {
    if (!c || !sb) return kmc_fail(KMC_E_ARG, "NULL argument");
    if (c->ev_pending) return kmc_fail(KMC_E_ARG, "kmc_event_enqueue: the previous event of this context was not finished");
    DevGuard guard(c);
    const int n = (int)c->st_n;
    const double t0 = kmc_now();
    TRY(enqueue_placement(c, sb, u0, u1, -1, -1, false));
    const double t1 = kmc_now();
    EventMail* mail = (EventMail*)c->slots[S_EVMAIL].p;
    if (c->relax_mode == 2)       // :94 / :256 -- the relaxation reads the adatom count and the placement status on the device
        TRY(kmc_relax2_enqueue(c, sb, 16, 1e-4, max_ls, n + 1, c->st_n_dev, &mail->status));
    c->enq_s[0] += t1 - t0;
    c->enq_s[1] += kmc_now() - t1;
    c->ev_pending = true;
    c->ev_max_ls = max_ls;
    c->ev_sbins = sb;
    return 0;
}

int kmc_event_finish(KmcCtx* c, int32_t* kind_out, int64_t* hop_k_out, double* rtot_out, KmcRelaxStats* stats) {
    if (!c) return kmc_fail(KMC_E_ARG, "ctx is NULL");
    if (!c->ev_pending) return kmc_fail(KMC_E_ARG, "kmc_event_finish: no event was enqueued on this context");
    DevGuard guard(c);
    c->ev_pending = false;
    CK(cudaStreamSynchronize(c->stream));                  // the ONE host synchronisation of the event
    if (stats) memset(stats, 0, sizeof(*stats));
    TRY(collect_placement(c, kind_out, hop_k_out, rtot_out));
    if (c->relax_mode != 2) return kmc_relax(c, c->ev_sbins, 16, 1e-4, c->ev_max_ls, stats);      // cross-check engines: host-launched
    KmcRelaxStats st{};
    bool again = false;
    TRY(kmc_relax2_collect(c, (int)c->st_n, &st, &again));
    if (again) return kmc_relax(c, c->ev_sbins, 16, 1e-4, c->ev_max_ls, stats);                    // work list grew: relax again (x0 untouched)
    if (stats) *stats = st;
    if (st.status != 0) return kmc_fail(st.status, "relaxation stopped at the line-search cap (%lld)", (long long)c->ev_max_ls);
    return 0;
}

// rate table -> selection (u0) -> placement (u1) of one event WITHOUT the relaxation (2DGrowth.py:318-329 up to the call of
// Relaxation inside DepositAdatom / HopAtom): the driver of a domain-decomposed trajectory places the atom on the full
// state and relaxes with its own (multi-GPU) loop.
int kmc_event_place(KmcCtx* c, const KmcBins* sb, double u0, double u1, int32_t* kind_out, int64_t* hop_k_out, double* rtot_out) {
    if (!c || !sb) return kmc_fail(KMC_E_ARG, "NULL argument");
    if (c->ev_pending) return kmc_fail(KMC_E_ARG, "an enqueued event has not been finished (kmc_event_finish)");
    DevGuard guard(c);
    TRY(enqueue_placement(c, sb, u0, u1, -1, -1, false));
    CK(cudaStreamSynchronize(c->stream));
    return collect_placement(c, kind_out, hop_k_out, rtot_out);
}

// 1 when the enqueued event of this context has completed on the device (kmc_event_finish will not block), 0 when it is
// still running, negative on error.  Lets one host thread keep many replicas in flight without lockstep.
int kmc_event_ready(KmcCtx* c) {
    if (!c) return kmc_fail(KMC_E_ARG, "ctx is NULL");
    if (!c->ev_pending) return 1;
    DevGuard guard(c);
    const cudaError_t e = cudaStreamQuery(c->stream);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    return kmc_fail(KMC_E_CUDA, "cudaStreamQuery -> %s", cudaGetErrorString(e));
}

// R replicas (one context, stream and substrate each): every event is queued before the first one is waited for, so
// the replicas' kernels overlap on the device (small systems relax in single-block / single-cluster launches).
// rc_out[r] = status of replica r; returns the first non-zero status other than KMC_E_ITERCAP, else 0.
int kmc_events_batch(KmcCtx** ctxs, KmcBins** sbins, int32_t R, const double* u0, const double* u1, int64_t max_ls, int32_t* kind_out,
                     int64_t* hop_k_out, double* rtot_out, KmcRelaxStats* stats, int32_t* rc_out) {
    if (!ctxs || !sbins || !u0 || !u1 || R < 0) return kmc_fail(KMC_E_ARG, "NULL argument");
    int first = 0;
    for (int r = 0; r < R; ++r) {
        const int rc = kmc_event_enqueue(ctxs[r], sbins[r], u0[r], u1[r], max_ls);
        if (rc_out) rc_out[r] = rc;
        if (rc != 0 && first == 0) first = rc;
    }
    for (int r = 0; r < R; ++r) {
        if (!ctxs[r] || !ctxs[r]->ev_pending) continue;
        const int rc = kmc_event_finish(ctxs[r], kind_out ? kind_out + r : nullptr, hop_k_out ? hop_k_out + r : nullptr,
                                        rtot_out ? rtot_out + r : nullptr, stats ? stats + r : nullptr);
        if (rc_out) rc_out[r] = rc;
        if (rc != 0 && rc != KMC_E_ITERCAP && first == 0) first = rc;
    }
    return first;
}

}  // extern "C"

// device-resident placement for kmc_deposit_place / kmc_hop_place / kmc_event (called from kmc_b200.cu)
int kmc_deposit_place_dev(KmcCtx* c, const KmcBins* sb, double u) {
    TRY(enqueue_placement(c, sb, 0.0, u, 0, -1, false));
    CK(cudaStreamSynchronize(c->stream));
    return collect_placement(c, nullptr, nullptr, nullptr);
}

int kmc_hop_place_dev(KmcCtx* c, const KmcBins* sb, int64_t k, double u) {
    const int n = (int)c->st_n;
    if (k < 0 || k >= n) return kmc_fail(KMC_E_ARG, "hop index %lld out of range (n=%d)", (long long)k, n);
    TRY(enqueue_placement(c, sb, 0.0, u, 1, (int)k, true));
    CK(cudaStreamSynchronize(c->stream));
    return collect_placement(c, nullptr, nullptr, nullptr);
}
