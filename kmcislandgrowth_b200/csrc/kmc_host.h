// kmc_host.h -- host-side state shared by the translation units of libkmcb200.so: the context, scratch
// slots, error plumbing and the launch helpers.  (The library is split into several .cu files so that the
// persistent relaxation kernels compile in parallel with the rest; every kernel is private to its file.)
#pragma once
#include "../../include/kmc_b200.h"
#include "kmc_device.cuh"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

namespace kmc {
struct KmcRelaxStatsDev;
}

// ------------------------------------------------------------------------------------ errors
int kmc_fail(int code, const char* fmt, ...);

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t _e = (call);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return kmc_fail(KMC_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)

#define TRY(expr)                 \
    do {                          \
        int _rc = (expr);         \
        if (_rc != 0) return _rc; \
    } while (0)

// ------------------------------------------------------------------------------------ buffers
struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct KmcBins {
    int64_t n = 0;
    int ncell2 = 0;
    double2* xy = nullptr;     // original order
    int* cell_of = nullptr;    // [n]
    int* start = nullptr;      // [ncell+2]
    double2* sxy = nullptr;    // [n]
    int* sidx = nullptr;       // [n]
};

struct TimingRec {
    int family;
    cudaEvent_t a, b;
};

// what one event leaves in device memory for the host to read back ONCE (kmc_event: 2DGrowth.py:317-329)
struct EventMail {
    long long res[2];       // compacted hop index k / dense adatom index (or -1, -1: deposit)
    double rtot;            // Rd + sum(Rk)
    int kind;               // 0 deposit, 1 hop
    int status;             // 0, KMC_E_NOCANDIDATE, KMC_E_DEPOSIT
    int n_new;              // adatoms after the placement
    int hop_who;            // adatom index handed to HopAtom
    double ref_y;           // min / max y of the film (DepositAdatom :73-75)
    double new_xy[2];       // where the atom was put
    int ncand, pick;        // HopAtom :244-248
};

struct KmcCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    KmcParams hp{};
    kmc::DevParams dp{};
    std::vector<DBuf> slots;        // scratch, indexed by small integers
    int64_t launches = 0;
    bool timing = false;
    std::vector<TimingRec> recs;
    std::vector<cudaEvent_t> event_pool;      // recycled timing events
    std::vector<std::string> families;
    // pinned mailbox for small readbacks
    void* h_mail = nullptr;
    void* h_mail_dev = nullptr;     // the same pinned page as seen from the device (mapped)
    // trajectory state
    double* st_xy = nullptr;
    int64_t st_n = 0, st_cap = 0;
    int* st_n_dev = nullptr;        // device copy of the adatom count (the on-device event path changes it)
    int relax_mode = 2;             // 2: persistent kernel on cell-sorted state (default), 1: first-generation persistent kernel, 0: host-driven loop
    size_t item_cap = 0;
    bool item_cap_forced = false;   // KMC_ITEM_CAP (tests) was applied to this context
    bool relax2_smem_set = false;   // the >48 KB dynamic shared memory opt-in of k_relax2 is per device / context
    // incremental rate table (kmc_rates_update): cached per-atom results and the positions they were computed at
    int rate_mode = 0;              // 0: full sweep per event (default, parity mode); 1: kmc_event uses kmc_rates_update
    double rate_threshold = 0.0;    // displacement (Angstrom) below which a neighbourhood counts as unchanged; 0 = exact
    double2* rc_xy = nullptr;
    double* rc_rate = nullptr;
    double* rc_du = nullptr;
    uint8_t* rc_surf = nullptr;
    double2* rc_pending = nullptr;  // positions of removed atoms whose neighbourhood must be refreshed
    int rc_npending = 0;
    int64_t rc_cap = 0, rc_n = 0, rc_last = 0;
    bool rc_valid = false;
    // persistent adatom cell list of the trajectory state (kmc_cells.cu): edited in place by the incremental rate table
    bool pl_valid = false;
    int pl_cur = 0;                 // which of the two buffer sets holds the list
    int64_t pl_n = 0;               // atoms filed in it
    int pl_removed = -1;            // index HopAtom deleted since the last update (-1: none)
    int pl_appended = 0;            // atoms appended since the last update
    uintptr_t pl_sig = 0;           // signature of the buffer addresses (a grown scratch slot loses its content)
    int64_t pl_stats[4] = {0, 0, 0, 0};   // full builds, in-place updates, edits of the last update (atoms after a full build), last kmc_cells_check
    int select_mode = 0;            // 0: sequential-order sums up to 65536 atoms, block scan beyond; 1: always sequential; 2: always block scan
    int event_mode = 1;             // 1 (default): the whole event is enqueued without a host round trip; 0: host-stepped placement (cross-check)
    int sm_count = 148;
    bool debug_sync = false;        // KMC_SYNC_DEBUG
    double enq_s[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // diagnostics (KMC_ENQ_TIMING): host seconds spent in the parts of kmc_event_enqueue
    int64_t diag[4] = {0, 0, 0, 0};  // counters of the last relaxation (kmc_relax_diag)
    // an event that was enqueued and not yet finished (kmc_event_enqueue / kmc_event_finish)
    bool ev_pending = false;
    int64_t ev_max_ls = 0;
    const KmcBins* ev_sbins = nullptr;
};

// binds the calling thread to the context's device for the duration of a C-ABI call (two contexts on different
// GPUs in one process, or a torch.cuda.set_device between calls, must not redirect allocations and launches)
struct DevGuard {
    int prev = -1;
    bool switched = false;
    explicit DevGuard(const KmcCtx* c) {
        if (!c) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != c->device) switched = cudaSetDevice(c->device) == cudaSuccess;
    }
    ~DevGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

enum Slot {
    S_IN0 = 0, S_IN1, S_IN2, S_IN3, S_OUT0, S_OUT1, S_OUT2, S_OUT3, S_OUT4, S_OUT5, S_OUT6, S_OUT7,
    S_CELLOF, S_START, S_CURSOR, S_SXY, S_SIDX, S_PARTIAL, S_DONE, S_CAND, S_E, S_CTRL,
    S_F, S_FAA, S_FAS, S_X0, S_XN, S_XNP, S_SN, S_TMP0, S_TMP1, S_TMP2, S_TMP3,
    S_FLAG, S_MOVERS, S_NMOV, S_LPART, S_FIX, S_AMIN, S_RED, S_SYNC,
    S_POS0, S_POS1, S_SD0, S_SD1, S_OI0, S_OI1, S_KEY, S_PERM, S_CHUNKS, S_ITEMS, S_NBR,
    S_EVMAIL, S_EVTMP, S_EVRATE, S_EVSURF, S_HALO0, S_HALO1, S_PX, S_BX, S_COLR,
    S_PL_START0, S_PL_START1, S_PL_SIDX0, S_PL_SIDX1, S_PL_CELL0, S_PL_CELL1, S_PL_SLOT0, S_PL_SLOT1, S_PL_SXY0, S_PL_SXY1, S_PL_EDITS,
    S_COUNT
};

int kmc_ensure(KmcCtx* c, DBuf& b, size_t bytes);

template <class T>
static inline int slot(KmcCtx* c, int id, size_t count, T** out) {
    TRY(kmc_ensure(c, c->slots[id], (count > 0 ? count : 1) * sizeof(T)));
    *out = (T*)c->slots[id].p;
    return 0;
}

// read `bytes` from device memory into the pinned mailbox and wait
int kmc_mail(KmcCtx* c, const void* dev, size_t bytes);

// ------------------------------------------------------------------------------------ launches
int kmc_family_id(KmcCtx* c, const char* name);
cudaEvent_t kmc_event_from_pool(KmcCtx* c);

struct LaunchScope {
    KmcCtx* c;
    TimingRec rec{};
    bool on;
    LaunchScope(KmcCtx* ctx, const char* fam) : c(ctx), on(ctx->timing) {
        c->launches++;
        if (on) {
            rec.family = kmc_family_id(c, fam);
            rec.a = kmc_event_from_pool(c);
            rec.b = kmc_event_from_pool(c);
            cudaEventRecord(rec.a, c->stream);
        }
    }
    ~LaunchScope() {
        if (on) {
            cudaEventRecord(rec.b, c->stream);
            c->recs.push_back(rec);
        }
    }
};

// KMC_SYNC_DEBUG=1: synchronise after every launch and name the kernel family that faulted (diagnostics)
static inline void kmc_debug_sync(KmcCtx* c, const char* what) {
    if (!c->debug_sync) return;
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) fprintf(stderr, "[kmc sync debug] %s -> %s\n", what, cudaGetErrorString(e));
}

#define LAUNCH(ctx, fam, kernel, grid, block, ...)                      \
    do {                                                                \
        {                                                               \
            LaunchScope _ls((ctx), (fam));                              \
            kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__); \
        }                                                               \
        kmc_debug_sync((ctx), #kernel);                                 \
    } while (0)

static inline double kmc_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Tiny utility kernels used instead of cudaMemsetAsync / cudaMemcpyAsync on the event path: with the default 8 hardware
// connections those copy-engine operations BLOCK THE HOST while another stream's long kernel is running (measured: 1-2 ms
// per call with 4-16 replicas in flight), kernel launches do not.  Results for the host go to pinned, device-mapped memory.
#ifdef __CUDACC__
static __global__ void k_zero_words(unsigned int* __restrict__ p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}
static __global__ void k_copy_words(unsigned int* __restrict__ dst, const unsigned int* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
// zero `bytes` (a multiple of 4) of device memory on the context's stream
static inline int kmc_zero_async(KmcCtx* c, void* p, size_t bytes) {
    const size_t n = bytes / 4;
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256);
    c->launches++;
    k_zero_words<<<grid, 256, 0, c->stream>>>((unsigned int*)p, n);
    kmc_debug_sync(c, "k_zero_words");
    return 0;
}
// copy `bytes` (a multiple of 4, small) from device memory to the pinned, device-mapped mailbox
static inline int kmc_to_mailbox_async(KmcCtx* c, size_t mail_offset, const void* dev, size_t bytes) {
    c->launches++;
    k_copy_words<<<1, 128, 0, c->stream>>>((unsigned int*)((char*)c->h_mail_dev + mail_offset), (const unsigned int*)dev, (int)(bytes / 4));
    kmc_debug_sync(c, "k_copy_words (mailbox)");
    return 0;
}
#endif

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

static inline kmc::CellView view_of(const KmcBins* b) {
    kmc::CellView v;
    v.start = b->start; v.sxy = b->sxy; v.sidx = b->sidx; v.n = (int)b->n;
    return v;
}

// ------------------------------------------------------------------------------------ cross-file entry points
// kmc_b200.cu
int kmc_scratch_cells(KmcCtx* c, const double2* d_xy, int n, int B, kmc::CellView* out);
int kmc_build_cells(KmcCtx* c, const double2* d_xy, int n, int B, int* cell_of, int* start, int* cursor, int* sidx, double2* sxy);
// kmc_cells.cu
int kmc_cells_update(KmcCtx* c, kmc::CellView* out);
void kmc_cells_on_remove(KmcCtx* c, int k);
void kmc_cells_on_append(KmcCtx* c);
int kmc_rates_dev(KmcCtx* c, const double2* d_xy, int n, const KmcBins* sb, double* rate, uint8_t* surf, double* du, double* ua, double* ul,
                  int* na, int* ns);
int kmc_state_reserve(KmcCtx* c, int64_t n);
int kmc_rates_update_impl(KmcCtx* c, const KmcBins* sb, double threshold, int64_t* recomputed);
int kmc_rc_on_remove(KmcCtx* c, int k);
int kmc_rc_on_append(KmcCtx* c);
int kmc_select_dev(KmcCtx* c, const double* rate, const uint8_t* surf, int n, double u, long long* res, double* rtot, double* pj);
// kmc_event.cu
int kmc_deposit_place_dev(KmcCtx* c, const KmcBins* sb, double u);
int kmc_hop_place_dev(KmcCtx* c, const KmcBins* sb, int64_t k, double u);
// kmc_relax1.cu / kmc_relax2.cu
int kmc_relax_persistent1(KmcCtx* c, const KmcBins* sb, int32_t scale0, double threshold0, int64_t max_ls, KmcRelaxStats* stats);
// n_dev != nullptr: the adatom count is read from device memory at kernel start (<= n_upper); status_dev != nullptr: the kernel does
// nothing when *status_dev != 0.  enqueue only: the results stay in the S_SYNC mailbox until kmc_relax2_collect.
int kmc_relax2_enqueue(KmcCtx* c, const KmcBins* sb, int32_t scale0, double threshold0, int64_t max_ls, int n_upper, const int* n_dev,
                       const int* status_dev);
// returns 1 when the work list overflowed and the relaxation has to be launched again
int kmc_relax2_collect(KmcCtx* c, int n, KmcRelaxStats* stats, bool* again);
int kmc_relax2_persistent(KmcCtx* c, const KmcBins* sb, int32_t scale0, double threshold0, int64_t max_ls, KmcRelaxStats* stats);
