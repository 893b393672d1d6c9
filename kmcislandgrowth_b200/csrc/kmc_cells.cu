// kmc_cells.cu -- the PERSISTENT adatom cell list of the trajectory state, updated in place of being rebuilt.
//
// The reference re-bins every adatom for every neighbour query (Bins.py:91, PutInBins of the whole array; 2DGrowth.py:
// 201-215 does that 2S+1 times per rate table).  The full-sweep path here rebuilds one scratch list per rate table; the
// incremental rate table (kmc_rates_update / kmc_set_rate_mode 1) keeps THIS list on the context instead and edits it:
//
//   * k_cells_detect compares the cell of every atom's current position with the cell it is filed under and appends
//     the atoms whose cell changed -- plus the atom HopAtom deleted (2DGrowth.py:241) and the atoms that were appended
//     (:93, :255) -- to a short edit list (removals by slot, insertions by (cell, index));
//   * with at most CELL_EDITS edits the three apply kernels produce the new list from the old one without sorting
//     or scanning: a kept atom's slot moves by (#insertions before it - #removals before it), an inserted atom goes
//     to its rank inside its cell, a cell's start moves by the edits in the cells before it; positions are refreshed
//     from the state in the same pass;
//   * more edits than that (or a first use, an uploaded state, a large out-of-grid population) rebuild the list with
//     the counting-sort kernels of kmc_b200.cu.
//
// The result is IDENTICAL to a rebuilt list -- same starts, same slots, ascending original index inside every cell
// (Bins.py:40-41 appends in input order) -- so every sum formed over it has the reference's order of terms;
// kmc_cells_check builds the scratch list next to it and counts differing words (tests/test_gpu_parity.py).
#include "kmc_host.h"

#include <algorithm>
#include <cstdlib>
#include <vector>

using namespace kmc;

namespace {

constexpr int CELL_EDITS = 64;      // removals / insertions one update applies in place

// edit list in device memory (slot S_PL_EDITS)
struct CellEdits {
    int n_rm, n_ins;
    int force_rebuild;              // the old list cannot be edited (out-of-grid cell holds an unsorted population)
    int pad;
    int rm_slot[CELL_EDITS], rm_cell[CELL_EDITS];
    int ins_cell[CELL_EDITS], ins_idx[CELL_EDITS];
};

__device__ __forceinline__ void push_rm(CellEdits* E, int slot, int cell) {
    const int k = atomicAdd(&E->n_rm, 1);
    if (k < CELL_EDITS) { E->rm_slot[k] = slot; E->rm_cell[k] = cell; }
}
__device__ __forceinline__ void push_ins(CellEdits* E, int cell, int idx) {
    const int k = atomicAdd(&E->n_ins, 1);
    if (k < CELL_EDITS) { E->ins_cell[k] = cell; E->ins_idx[k] = idx; }
}

// new index i' -> old index: HopAtom's np.delete shifts the atoms behind the removed one down by one
__device__ __forceinline__ int old_index(int inew, int removed) { return (removed >= 0 && inew >= removed) ? inew + 1 : inew; }

__global__ void k_cells_detect(DevParams P, const double2* __restrict__ xy, int n_new, int n_old, int removed,
                               const int* __restrict__ cell_old, const int* __restrict__ slot_old, const int* __restrict__ start_old,
                               int* __restrict__ cell_new, CellEdits* E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        if (removed >= 0) push_rm(E, slot_old[removed], cell_old[removed]);
        if (start_old[P.ncell + 1] - start_old[P.ncell] > 64) E->force_rebuild = 1;       // k_bin_order leaves such a cell unsorted
    }
    if (i >= n_new) return;
    const double2 p = xy[i];
    int nx, ny;
    bin_index(P, p.x, p.y, nx, ny);
    const int c = flat_cell(P, nx, ny);
    cell_new[i] = c;
    const int io = old_index(i, removed);
    if (io >= n_old) { push_ins(E, c, i); return; }             // appended since the last update
    const int co = cell_old[io];
    if (co != c) { push_rm(E, slot_old[io], co); push_ins(E, c, i); }
}

struct EditsSmem {
    int n_rm, n_ins;
    int rm_slot[CELL_EDITS], rm_cell[CELL_EDITS], ins_cell[CELL_EDITS], ins_idx[CELL_EDITS];
};
__device__ __forceinline__ void load_edits(EditsSmem& S, const CellEdits* E) {
    if (threadIdx.x == 0) { S.n_rm = E->n_rm; S.n_ins = E->n_ins; }
    for (int k = threadIdx.x; k < CELL_EDITS; k += blockDim.x) {
        S.rm_slot[k] = E->rm_slot[k]; S.rm_cell[k] = E->rm_cell[k];
        S.ins_cell[k] = E->ins_cell[k]; S.ins_idx[k] = E->ins_idx[k];
    }
    __syncthreads();
}

// atoms that stay in their cell: one thread per OLD slot
__global__ void k_cells_apply_kept(const double2* __restrict__ xy, int n_old, int removed, const int* __restrict__ sidx_old,
                                   const int* __restrict__ cell_old, const CellEdits* __restrict__ E, double2* __restrict__ sxy_new,
                                   int* __restrict__ sidx_new, int* __restrict__ slot_new) {
    __shared__ EditsSmem S;
    load_edits(S, E);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_old) return;
    int before = 0;
    for (int k = 0; k < S.n_rm; ++k) {
        if (S.rm_slot[k] == s) return;                          // removed from this cell
        before += S.rm_slot[k] < s;
    }
    const int io = sidx_old[s];
    const int inew = io - ((removed >= 0 && io > removed) ? 1 : 0);
    const int c = cell_old[io];
    int ins = 0;
    for (int k = 0; k < S.n_ins; ++k) ins += (S.ins_cell[k] < c) || (S.ins_cell[k] == c && S.ins_idx[k] < inew);
    const int ns = s - before + ins;
    sxy_new[ns] = xy[inew];
    sidx_new[ns] = inew;
    slot_new[inew] = ns;
}

// atoms filed under a new cell: one warp per insertion; the rank inside the cell is the number of kept members and of
// other insertions with a smaller index
__global__ void k_cells_apply_inserted(const double2* __restrict__ xy, int removed, const int* __restrict__ start_old,
                                       const int* __restrict__ sidx_old, const CellEdits* __restrict__ E, double2* __restrict__ sxy_new,
                                       int* __restrict__ sidx_new, int* __restrict__ slot_new) {
    __shared__ EditsSmem S;
    load_edits(S, E);
    const int lane = threadIdx.x & 31;
    for (int e = threadIdx.x >> 5; e < S.n_ins; e += blockDim.x >> 5) {
        const int c = S.ins_cell[e], j = S.ins_idx[e];
        const int lo = start_old[c], hi = start_old[c + 1];
        int cnt = 0;
        for (int k = lane; k < S.n_rm; k += 32) cnt -= S.rm_cell[k] < c;
        for (int k = lane; k < S.n_ins; k += 32) cnt += (S.ins_cell[k] < c) || (S.ins_cell[k] == c && S.ins_idx[k] < j);
        for (int s = lo + lane; s < hi; s += 32) {
            bool gone = false;
            for (int k = 0; k < S.n_rm; ++k) gone = gone || S.rm_slot[k] == s;
            if (gone) continue;
            const int io = sidx_old[s];
            cnt += (io - ((removed >= 0 && io > removed) ? 1 : 0)) < j;
        }
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) {
            const int ns = lo + cnt;
            sxy_new[ns] = xy[j];
            sidx_new[ns] = j;
            slot_new[j] = ns;
        }
    }
}

// start[c] = atoms in the cells before c
__global__ void k_cells_apply_starts(int ncell2, const int* __restrict__ start_old, const CellEdits* __restrict__ E, int* __restrict__ start_new) {
    __shared__ EditsSmem S;
    load_edits(S, E);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell2; c += gridDim.x * blockDim.x) {
        int d = 0;
        for (int k = 0; k < S.n_rm; ++k) d -= S.rm_cell[k] < c;
        for (int k = 0; k < S.n_ins; ++k) d += S.ins_cell[k] < c;
        start_new[c] = start_old[c] + d;
    }
}

// slot_of[sidx[s]] = s after a full build
__global__ void k_cells_inverse(int n, const int* __restrict__ sidx, int* __restrict__ slot_of) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) slot_of[sidx[s]] = s;
}

// words in which two lists differ (starts, indices, position bits).  An out-of-grid cell of more than 64 atoms is left
// unsorted by k_bin_order (nobody's stencil addresses it): its slots are compared as counts only.
__global__ void k_cells_compare(int n, int ncell2, const int* __restrict__ sa, const int* __restrict__ ia, const double2* __restrict__ pa,
                                const int* __restrict__ sb, const int* __restrict__ ib, const double2* __restrict__ pb,
                                unsigned long long* __restrict__ diff) {
    unsigned long long d = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncell2; i += gridDim.x * blockDim.x) d += sa[i] != sb[i];
    const int ovf = sb[ncell2 - 2];
    const int n_cmp = (sb[ncell2 - 1] - ovf > 64) ? ovf : n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cmp; i += gridDim.x * blockDim.x) {
        d += ia[i] != ib[i];
        d += (__double_as_longlong(pa[i].x) != __double_as_longlong(pb[i].x)) || (__double_as_longlong(pa[i].y) != __double_as_longlong(pb[i].y));
    }
    if (d) atomicAdd(diff, d);
}

struct PlBuffers {
    int *start[2], *sidx[2], *cell[2], *slot[2];
    double2* sxy[2];
    CellEdits* edits;
};

int pl_buffers(KmcCtx* c, int n, PlBuffers* b) {
    const size_t ncell2 = (size_t)c->dp.ncell + 2, cap = (size_t)n + 64;
    for (int k = 0; k < 2; ++k) {
        TRY(slot(c, S_PL_START0 + k, ncell2, &b->start[k]));
        TRY(slot(c, S_PL_SIDX0 + k, cap, &b->sidx[k]));
        TRY(slot(c, S_PL_CELL0 + k, cap, &b->cell[k]));
        TRY(slot(c, S_PL_SLOT0 + k, cap, &b->slot[k]));
        TRY(slot(c, S_PL_SXY0 + k, cap, &b->sxy[k]));
    }
    TRY(slot(c, S_PL_EDITS, 1, &b->edits));
    return 0;
}

int pl_full_build(KmcCtx* c, int n, PlBuffers& b) {
    int* cursor;
    TRY(slot(c, S_CURSOR, (size_t)c->dp.ncell + 2, &cursor));
    TRY(kmc_build_cells(c, (const double2*)c->st_xy, n, 1, b.cell[0], b.start[0], cursor, b.sidx[0], b.sxy[0]));
    LAUNCH(c, "bin_edit", k_cells_inverse, cdiv(n, 256), 256, n, (const int*)b.sidx[0], b.slot[0]);
    CK(cudaGetLastError());
    c->pl_cur = 0;
    c->pl_n = n;
    c->pl_valid = true;
    c->pl_removed = -1;
    c->pl_appended = 0;
    c->pl_stats[0]++;
    c->pl_stats[2] = n;
    return 0;
}

}  // namespace

// index bookkeeping of the trajectory state (called where the rate cache's hooks are): one removal followed by
// appends is what an event does; anything else makes the next update a rebuild
void kmc_cells_on_remove(KmcCtx* c, int k) {
    if (!c->pl_valid) return;
    if (c->pl_removed >= 0 || c->pl_appended > 0) { c->pl_valid = false; return; }
    c->pl_removed = k;
}
void kmc_cells_on_append(KmcCtx* c) {
    if (c->pl_valid) c->pl_appended++;
}

// the cell list of the trajectory state, brought up to date with the positions in c->st_xy
int kmc_cells_update(KmcCtx* c, CellView* out) {
    const int n = (int)c->st_n;
    if (n <= 0) return kmc_fail(KMC_E_ARG, "kmc_cells_update: empty state");
    PlBuffers b;
    TRY(pl_buffers(c, n, &b));
    {   // a grown slot loses its content: the list is then rebuilt
        uintptr_t sig = (uintptr_t)b.edits;
        for (int k = 0; k < 2; ++k)
            sig = sig * 1315423911u + ((uintptr_t)b.start[k] ^ ((uintptr_t)b.sidx[k] << 1) ^ ((uintptr_t)b.cell[k] << 2) ^ ((uintptr_t)b.slot[k] << 3) ^
                                       ((uintptr_t)b.sxy[k] << 4));
        if (sig != c->pl_sig) { c->pl_valid = false; c->pl_sig = sig; }
    }
    const int n_old = (int)c->pl_n;
    bool full = !c->pl_valid || n != n_old - (c->pl_removed >= 0 ? 1 : 0) + c->pl_appended || c->pl_appended > CELL_EDITS / 2;
    if (!full) {
        const int cur = c->pl_cur, nxt = cur ^ 1;
        TRY(kmc_zero_async(c, b.edits, 16));
        LAUNCH(c, "bin_edit", k_cells_detect, cdiv(n, 256), 256, c->dp, (const double2*)c->st_xy, n, n_old, c->pl_removed,
               (const int*)b.cell[cur], (const int*)b.slot[cur], (const int*)b.start[cur], b.cell[nxt], b.edits);
        CK(cudaGetLastError());
        TRY(kmc_mail(c, b.edits, 16));
        const int* h = (const int*)c->h_mail;
        const int n_rm = h[0], n_ins = h[1], force = h[2];
        if (force || n_rm > CELL_EDITS || n_ins > CELL_EDITS) {
            full = true;
        } else {
            const int ncell2 = c->dp.ncell + 2;
            LAUNCH(c, "bin_edit", k_cells_apply_kept, cdiv(n_old, 256), 256, (const double2*)c->st_xy, n_old, c->pl_removed, (const int*)b.sidx[cur],
                   (const int*)b.cell[cur], (const CellEdits*)b.edits, b.sxy[nxt], b.sidx[nxt], b.slot[nxt]);
            if (n_ins > 0)
                LAUNCH(c, "bin_edit", k_cells_apply_inserted, 1, 256, (const double2*)c->st_xy, c->pl_removed, (const int*)b.start[cur],
                       (const int*)b.sidx[cur], (const CellEdits*)b.edits, b.sxy[nxt], b.sidx[nxt], b.slot[nxt]);
            LAUNCH(c, "bin_edit", k_cells_apply_starts, std::min(cdiv(ncell2, 256), 2048u), 256, ncell2, (const int*)b.start[cur],
                   (const CellEdits*)b.edits, b.start[nxt]);
            CK(cudaGetLastError());
            c->pl_cur = nxt;
            c->pl_n = n;
            c->pl_removed = -1;
            c->pl_appended = 0;
            c->pl_stats[1]++;
            c->pl_stats[2] = n_rm + n_ins;
        }
    }
    if (full) TRY(pl_full_build(c, n, b));
    const int cur = c->pl_cur;
    out->start = b.start[cur]; out->sxy = b.sxy[cur]; out->sidx = b.sidx[cur]; out->n = n;
    return 0;
}

extern "C" int kmc_cells_stats(KmcCtx* c, int64_t out[4]) {
    if (!c || !out) return kmc_fail(KMC_E_ARG, "NULL argument");
    for (int k = 0; k < 4; ++k) out[k] = c->pl_stats[k];
    return 0;
}

extern "C" int kmc_cells_check(KmcCtx* c, int64_t* differing_words) {
    if (!c || !differing_words) return kmc_fail(KMC_E_ARG, "NULL argument");
    if (c->ev_pending) return kmc_fail(KMC_E_ARG, "an enqueued event has not been finished (kmc_event_finish)");
    DevGuard guard(c);
    const int n = (int)c->st_n;
    if (n <= 0) { *differing_words = 0; return 0; }
    CellView pl, ref;
    TRY(kmc_cells_update(c, &pl));
    TRY(kmc_scratch_cells(c, (const double2*)c->st_xy, n, 1, &ref));
    unsigned long long* diff;
    TRY(slot(c, S_OUT6, 1, &diff));
    TRY(kmc_zero_async(c, diff, 8));
    LAUNCH(c, "bin_edit", k_cells_compare, 256, 256, n, c->dp.ncell + 2, pl.start, pl.sidx, pl.sxy, ref.start, ref.sidx, ref.sxy, diff);
    CK(cudaGetLastError());
    TRY(kmc_mail(c, diff, 8));
    *differing_words = (int64_t) * (const unsigned long long*)c->h_mail;
    c->pl_stats[3] = *differing_words;
    if (*differing_words != 0 && getenv("KMC_CELLS_DEBUG")) {        // diagnostics: the first differing entries
        const int ncell2 = c->dp.ncell + 2;
        std::vector<int> sa(ncell2), sb(ncell2), ia(n), ib(n);
        cudaMemcpy(sa.data(), pl.start, sizeof(int) * ncell2, cudaMemcpyDeviceToHost);
        cudaMemcpy(sb.data(), ref.start, sizeof(int) * ncell2, cudaMemcpyDeviceToHost);
        cudaMemcpy(ia.data(), pl.sidx, sizeof(int) * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(ib.data(), ref.sidx, sizeof(int) * n, cudaMemcpyDeviceToHost);
        int shown = 0;
        for (int i = 0; i < ncell2 && shown < 8; ++i)
            if (sa[i] != sb[i]) { fprintf(stderr, "[cells] start[%d] %d != %d\n", i, sa[i], sb[i]); ++shown; }
        shown = 0;
        for (int i = 0; i < n && shown < 12; ++i)
            if (ia[i] != ib[i]) { fprintf(stderr, "[cells] sidx[%d] %d != %d\n", i, ia[i], ib[i]); ++shown; }
        fprintf(stderr, "[cells] n %d stats %lld %lld %lld\n", n, (long long)c->pl_stats[0], (long long)c->pl_stats[1], (long long)c->pl_stats[2]);
    }
    return 0;
}
