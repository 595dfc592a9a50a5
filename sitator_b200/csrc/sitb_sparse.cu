// sitator_b200 -- K3s: the assign step over cached sparse landmark vectors.
//
// Landmark vectors are ~1.5 % dense (22 of 1500 components), so the fill pass that builds the Gram
// also stores every row compressed (sitb_fill.cu, sparse_ptr/sparse_k/sparse_v: ~230 B per row
// instead of 12 KB).  The clustering plugin's later passes over "all landmark vectors"
//   cluster/mcl.py:81-83          best-matching row per cluster
//   DotProdClassifier.pyx:86-118  predict, bincount, predict again
//   cluster/mcl.py:118-122        representative landmark vectors
// then stream those rows from HBM (eight lanes per row) instead of recomputing them.
//
// "Best row" reductions (max value, first row) are kept per warp in shared memory, merged per CTA and
// then over CTAs by a second tiny kernel (no atomics or locks on the hot path).
//
// Measured (100 000 LLZO frames, 5.6e6 rows, 88 clusters; scripts/time_assign_sparse.py; ms for pass B / pass C / labels
// only): round 1's lane-per-row kernel (32 rows staged through shared memory, 12 register slots of cluster sums per
// lane: 106 warp instructions per row but 121 registers and 14 resident warps) 1.64 / 2.36 / 1.41; this kernel
// 1.82 / 1.81 / 1.45.  What moved the numbers: private copies of the representative-vector sums (pass C was bound by
// same-address FP64 atomics in L2), skipping best-table updates that cannot change the table.  What did not: row-ordered
// slots instead of a pool, entry loads that do not wait for the row pointer, the centre tables in shared memory, a peel
// loop vs. a shared-memory accumulator -- every variant runs at ~250 warp instructions per row and ~65 % issue
// utilisation (profiles/r02_assign_sparse_ncu.md); 6 % of the rows have more than 32 entries and take the whole-warp
// path, which is 14 % of the instructions.
#include "../../include/sitator_b200.h"
#include "sitb_assign.cuh"

namespace sitb {

// Per-warp private (value bits, row) tables in shared memory, merged per CTA at the end: "first row among equal
// values" is kept explicitly (larger value, or equal value and lower row).
struct WarpBest {
    unsigned long long* val;     // [C] of this warp
    unsigned long long* row;
    __device__ __forceinline__ void update(int c, double v, unsigned long long r) {
        const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
        if (vb > val[c] || (vb == val[c] && vb != 0ull && r < row[c])) { val[c] = vb; row[c] = r; }
    }
};

template <int NCH>
__device__ __forceinline__ void assign_row(
    long long r, int nent, unsigned long long off, int lane, int k0, double v0, const uint16_t* __restrict__ pk,
    const double* __restrict__ pv, long long row0, int L, const int* __restrict__ cid, const double* __restrict__ cw,
    double thr, long long* __restrict__ labels, double* __restrict__ confs, unsigned long long* counts, unsigned* hist,
    unsigned long long* best_scratch, WarpBest& wb, double* __restrict__ rep, double* __restrict__ rep_w,
    unsigned long long* site_scratch, WarpBest& ws) {
        int myc[NCH];
        double mypr[NCH], myv[NCH];
        int myk[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = 32 * c + lane;
            myc[c] = -1; mypr[c] = 0.0; myv[c] = 0.0; myk[c] = 0;
            if (e < nent) {
                const int k = c == 0 ? k0 : (int)pk[off + e];          // the first 32 entries were prefetched
                const double v = c == 0 ? v0 : pv[off + e];
                myk[c] = k; myv[c] = v;
                myc[c] = __ldg(cid + k);
                mypr[c] = v * __ldg(cw + k);
            }
        }
        // peel the row's clusters in ascending id (np.argmax: first maximum; all zero -> index 0)
        double bestc = 0.0;
        int bestid = 0;
        for (;;) {
            int mine = 0x7FFFFFFF;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                if (myc[c] >= 0 && myc[c] < mine) mine = myc[c];
            const int cur = __reduce_min_sync(0xffffffffu, mine);
            if (cur == 0x7FFFFFFF) break;
            double part = 0.0;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                if (myc[c] == cur) { part += mypr[c]; myc[c] = -1; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            const double conf = fabs(part);
            if (conf > bestc) { bestc = conf; bestid = cur; }
            if (best_scratch && lane == 0) wb.update(cur, conf, (unsigned long long)(row0 + r));   // cluster/mcl.py:81-83
        }
        long long label = bestid;
        double conf = bestc;
        if (nent == 0 || !(conf >= thr)) { label = -1; conf = 0.0; }     // DotProdClassifier.pyx:168-172,184-186
        if (lane == 0) {
            if (labels) labels[r] = label;
            if (confs) confs[r] = conf;
            if (label >= 0) {
                if (counts) atomicAdd(&hist[label], 1u);
                if (rep_w) atomicAdd(&rep_w[label], conf);
                if (site_scratch) ws.update((int)label, conf, (unsigned long long)(row0 + r));
            }
        }
        if (rep && label >= 0) {
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                if (32 * c + lane < nent) atomicAdd(&rep[(size_t)label * L + myk[c]], conf * myv[c]);
        }
}

// (value, first row) table updates from many lanes at once: raise the value with atomicMax and invalidate the row
// if this lane raised it; after a __syncwarp every lane that holds the final value atomicMins its row.
__device__ __forceinline__ void table_raise(WarpBest& t, int c, unsigned long long vb) {
    const unsigned long long old = atomicMax(&t.val[c], vb);
    if (old < vb) t.row[c] = ~0ull;
}
__device__ __forceinline__ void table_claim(WarpBest& t, int c, unsigned long long vb, unsigned long long r) {
    if (vb != 0ull && vb == t.val[c]) atomicMin(&t.row[c], r);
}

// a hard row by the whole warp, out of line (its register arrays would otherwise set the register count of the caller)
__device__ __noinline__ void assign_row_hard(
    long long r, int n, unsigned long long off, int lane, const uint16_t* __restrict__ pk, const double* __restrict__ pv,
    long long row0, int L, const int* __restrict__ cid, const double* __restrict__ cw, double thr, long long* __restrict__ labels,
    double* __restrict__ confs, unsigned long long* counts, unsigned* hist, unsigned long long* best_scratch,
    unsigned long long* wb_val, double* __restrict__ rep, double* __restrict__ rep_w, unsigned long long* site_scratch,
    unsigned long long* ws_val, int C) {
    WarpBest wb = {wb_val, wb_val + C}, ws = {ws_val, ws_val + C};
    int k0 = 0;
    double v0 = 0.0;
    if (lane < n) { k0 = pk[off + lane]; v0 = pv[off + lane]; }
    if (n <= 128)
        assign_row<4>(r, n, off, lane, k0, v0, pk, pv, row0, L, cid, cw, thr, labels, confs, counts, hist, best_scratch, wb,
                      rep, rep_w, site_scratch, ws);
    else
        assign_row<ENTRY_CAP / 32>(r, n, off, lane, k0, v0, pk, pv, row0, L, cid, cw, thr, labels, confs, counts, hist,
                                   best_scratch, wb, rep, rep_w, site_scratch, ws);
}

// ---- eight lanes per row ---------------------------------------------------------------------------------------
// A warp takes four rows at a time, eight lanes each: a lane loads up to four of its row's entries straight from the
// row's slot (no shared-memory staging, 64 registers: twice the resident warps of the lane-per-row kernel, which
// ncu showed waiting on memory with 14 warps per SM and a third of the issue slots used); the per-cluster sums are
// scattered into a small shared-memory accumulator per row.  Sums are formed in an order that depends on the row alone.
// Rows with more than 32 entries are "hard" and handled by the whole warp afterwards (assign_row).
template <bool TS, bool SLOT>
__global__ void __launch_bounds__(512, 2) k_assign_sparse8(
    const unsigned long long* __restrict__ row_ptr, const uint16_t* __restrict__ pk, const double* __restrict__ pv,
    long long n_rows, long long row0, int L, const int* __restrict__ cid, const double* __restrict__ cw,
    int n_clusters, double thr, long long* __restrict__ labels, double* __restrict__ confs,
    unsigned long long* __restrict__ counts, unsigned long long* __restrict__ best_scratch, double* __restrict__ rep,
    double* __restrict__ rep_w, unsigned long long* __restrict__ site_scratch,
    const long long* __restrict__ row_list, const unsigned long long* __restrict__ n_list, int rep_copies) {
    // rep: rep_copies private copies [C][L] (CTA b adds to copy b % rep_copies; the launcher sums them): the ~22 atomics
    // per row otherwise pile up on the few hundred (site, landmark) addresses that carry nearly all the weight
    if (n_list) n_rows = (long long)*n_list;
    // shared: per warp [C] best val | [C] best row (if asked for) | [C] site val | [C] site row (if asked for); [C] hist;
    // [C] confidence sums
    extern __shared__ unsigned long long smem_u64[];
    const int C = n_clusters;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int sub = lane >> 3, sl = lane & 7;
    const int tstride = (best_scratch ? 2 * C : 0) + (site_scratch ? 2 * C : 0);
    unsigned long long* wbase = smem_u64 + (size_t)warp * tstride;
    WarpBest wb = {wbase, wbase + C};
    unsigned long long* sbase = wbase + (best_scratch ? 2 * C : 0);
    WarpBest ws = {sbase, sbase + C};
    unsigned* hist = (unsigned*)(smem_u64 + (size_t)nwarps * tstride);
    double* wsum = (double*)(smem_u64 + (size_t)nwarps * tstride + (C + 1) / 2);
    // TS: the centre tables in shared memory.  The look-ups are random gathers (32 lanes, 32 lines): through L1 they
    // cost a tag cycle per line and bounded every variant of this kernel at ~70 cycles per row.
    double* s_cw = wsum + C;                                          // [L]
    double* acc = s_cw + (TS ? L : 0) + ((size_t)warp * 4 + sub) * C; // [C] accumulator of this 8-lane group
    int* s_cid = (int*)(s_cw + (TS ? L : 0) + (size_t)nwarps * 4 * C);   // [L]
    if (TS)
        for (int i = threadIdx.x; i < L; i += blockDim.x) { s_cw[i] = cw[i]; s_cid[i] = cid[i]; }
    for (int i = threadIdx.x; i < nwarps * 4 * C; i += blockDim.x) s_cw[(TS ? L : 0) + i] = 0.0;
    double* rep_mine = rep ? rep + (size_t)(blockIdx.x % (unsigned)rep_copies) * (size_t)C * L : nullptr;
    for (int i = threadIdx.x; i < nwarps * tstride; i += blockDim.x) smem_u64[i] = 0ull;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { hist[i] = 0u; wsum[i] = 0.0; }
    __syncthreads();
    const long long warp_global = (long long)blockIdx.x * nwarps + warp;
    const long long n_warps = (long long)gridDim.x * nwarps;
    const long long n_quads = (n_rows + 3) >> 2;
    for (long long q = warp_global; q < n_quads; q += n_warps) {
        const long long ri = (q << 2) + sub;
        const bool valid = ri < n_rows;
        const long long r = (row_list && valid) ? row_list[ri] : ri;
        unsigned k[4];
        double v[4], pr[4];
        int c[4];
        // SLOT: the rows live in fixed 32-entry slots in row order (sitb_pass_stats_slotted), so the entry loads do not
        // wait for the row's pointer: lane sl takes entries 4 sl .. 4 sl + 3 with three vector loads, issued together
        // with the pointer load, and the slots of the warp's next quad are prefetched into L2.
        uint2 kk = make_uint2(0u, 0u);
        double2 va = make_double2(0.0, 0.0), vb2 = va;
        if (SLOT && valid) {
            const size_t base = (size_t)r * 32 + 4 * sl;
            kk = __ldcs((const uint2*)(pk + base));
            va = __ldcs((const double2*)(pv + base));
            vb2 = __ldcs((const double2*)(pv + base + 2));
            if (!row_list) {
                const long long rn = ((q + n_warps) << 2) + sub;
                if (rn < n_rows) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pv + (size_t)rn * 32 + 4 * sl));
                    if ((sl & 3) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(pk + (size_t)rn * 32 + 4 * sl));
                }
            }
        }
        const unsigned long long ptr = valid ? row_ptr[r] : 0ull;
        const int nent = (int)(ptr & 0xFF);
        const unsigned long long off = ptr >> 8;
        const bool hard = nent > 32;
        const int ne = hard ? 0 : nent;
        if (SLOT) {
            k[0] = kk.x & 0xFFFFu; k[1] = kk.x >> 16; k[2] = kk.y & 0xFFFFu; k[3] = kk.y >> 16;
            v[0] = va.x; v[1] = va.y; v[2] = vb2.x; v[3] = vb2.y;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (!(4 * sl + j < ne)) { k[j] = 0u; v[j] = 0.0; }   // (beyond the row: the slot holds whatever was there)
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool act = sl + 8 * j < ne;
                const unsigned long long a = off + (unsigned long long)(sl + 8 * j);
                k[j] = 0u; v[j] = 0.0;                                   // (idle lanes: landmark 0, value 0; never counted)
                if (act) { k[j] = __ldcs(pk + a); v[j] = __ldcs(pv + a); }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool act = (SLOT ? 4 * sl + j : sl + 8 * j) < ne;
            const int cc = TS ? s_cid[k[j]] : __ldg(cid + k[j]);
            pr[j] = v[j] * (TS ? s_cw[k[j]] : __ldg(cw + k[j]));
            c[j] = act ? cc : -1;
        }
        // Per-cluster sums of the row in the group's accumulator acc[C] (shared memory, zero between rows).  A lane first
        // merges its own entries of one cluster (entry order), then the eight lanes add their partial sums one lane at a
        // time: a fixed order that depends on the row alone.  (A peel loop over the row's clusters with 8-lane butterflies
        // cost 205 warp instructions per row, this ~70.)
        int cm[4];
        double pm[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { cm[j] = c[j]; pm[j] = pr[j]; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i + 1; j < 4; ++j)
                if (cm[j] >= 0 && cm[j] == cm[i]) { pm[i] += pm[j]; cm[j] = -1; }
        const int maxn = __reduce_max_sync(0xffffffffu, ne);
        const int last_lane = SLOT ? (maxn + 3) >> 2 : (maxn < 8 ? maxn : 8);   // lanes >= this hold no entries in any row of the quad
        for (int t = 0; t < last_lane; ++t) {
            if (sl == t) {
                // (distinct clusters after the merge: the four read-modify-writes are independent)
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                if (cm[0] >= 0) a0 = acc[cm[0]];
                if (cm[1] >= 0) a1 = acc[cm[1]];
                if (cm[2] >= 0) a2 = acc[cm[2]];
                if (cm[3] >= 0) a3 = acc[cm[3]];
                if (cm[0] >= 0) acc[cm[0]] = a0 + pm[0];
                if (cm[1] >= 0) acc[cm[1]] = a1 + pm[1];
                if (cm[2] >= 0) acc[cm[2]] = a2 + pm[2];
                if (cm[3] >= 0) acc[cm[3]] = a3 + pm[3];
            }
            __syncwarp();
        }
        // arg-max over the clusters of the row: largest |sum|, lowest id among equals, all zero -> index 0 (np.argmax)
        double tot[4];
        double bestc = 0.0;
        int bestid = 0x7FFFFFFF;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tot[j] = c[j] >= 0 ? fabs(acc[c[j]]) : 0.0;
            if (c[j] >= 0 && (tot[j] > bestc || (tot[j] == bestc && c[j] < bestid))) { bestc = tot[j]; bestid = c[j]; }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, bestc, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bestid, o);
            if (ob > bestc || (ob == bestc && oi < bestid)) { bestc = ob; bestid = oi; }
        }
        if (!(bestc > 0.0)) bestid = 0;
        if (best_scratch) {                                           // cluster/mcl.py:81-83: every cluster of the row
            // (the table only grows: a value below the entry changes nothing, and after the first few thousand rows
            // that is nearly every value -- the 64-bit shared-memory atomics are compare-and-swap loops.  Several
            // entries of one cluster offer the same (value, row): harmless.)
            bool need = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long vb = (unsigned long long)__double_as_longlong(tot[j]);
                need = need || (c[j] >= 0 && vb != 0ull && vb >= wb.val[c[j]]);
            }
            if (__any_sync(0xffffffffu, need)) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c[j] >= 0) table_raise(wb, c[j], (unsigned long long)__double_as_longlong(tot[j]));
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c[j] >= 0) table_claim(wb, c[j], (unsigned long long)__double_as_longlong(tot[j]), (unsigned long long)(row0 + r));
                __syncwarp();
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c[j] >= 0) acc[c[j]] = 0.0;
        __syncwarp();
        long long label = bestid;
        double conf = bestc;
        if (ne == 0 || !(conf >= thr)) { label = -1; conf = 0.0; }       // DotProdClassifier.pyx:168-172,184-186
        const bool easy = valid && !hard;
        if (easy && sl == 0) {
            if (labels) labels[r] = label;
            if (confs) confs[r] = conf;
            if (label >= 0) {
                if (counts) atomicAdd(&hist[label], 1u);
                if (rep_w) atomicAdd(&wsum[label], conf);
            }
        }
        if (site_scratch) {
            const unsigned long long vb = (unsigned long long)__double_as_longlong(conf);
            const bool on = easy && sl == 0 && label >= 0 && vb != 0ull && vb >= ws.val[label];
            if (__any_sync(0xffffffffu, on)) {
                if (on) table_raise(ws, (int)label, vb);
                __syncwarp();
                if (on) table_claim(ws, (int)label, vb, (unsigned long long)(row0 + r));
                __syncwarp();
            }
        }
        if (rep && easy && label >= 0) {                                  // mcl.py:118-122
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((SLOT ? 4 * sl + j : sl + 8 * j) < ne) atomicAdd(&rep_mine[(size_t)label * L + k[j]], conf * v[j]);
        }
        // the hard rows of the quad, one at a time by the whole warp
        unsigned hm = __ballot_sync(0xffffffffu, valid && hard && sl == 0);
        while (hm) {
            const int i = __ffs(hm) - 1;
            hm &= hm - 1;
            const unsigned long long p = __shfl_sync(0xffffffffu, ptr, i);
            const long long rh = __shfl_sync(0xffffffffu, r, i);
            assign_row_hard(rh, (int)(p & 0xFF), p >> 8, lane, pk, pv, row0, L, cid, cw, thr, labels, confs, counts, hist,
                            best_scratch, wb.val, rep_mine, rep_w, site_scratch, ws.val, C);
            __syncwarp();
        }
    }
    __syncthreads();
    // merge the warps' tables (max value, then lowest row) and hand the CTA's table to the merge kernel
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (counts && hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
        if (rep_w && wsum[i] != 0.0) atomicAdd(&rep_w[i], wsum[i]);
        for (int t = 0; t < 2; ++t) {
            unsigned long long* scratch = t == 0 ? best_scratch : site_scratch;
            if (!scratch) continue;
            const size_t toff = (t == 1 && best_scratch) ? 2 * (size_t)C : 0;
            unsigned long long bv = 0ull, br = 0ull;
            for (int w = 0; w < nwarps; ++w) {
                const unsigned long long vv = smem_u64[(size_t)w * tstride + toff + i];
                const unsigned long long rr = smem_u64[(size_t)w * tstride + toff + C + i];
                if (vv > bv || (vv == bv && rr < br)) { bv = vv; br = rr; }
            }
            scratch[((size_t)blockIdx.x * 2) * C + i] = bv;
            scratch[((size_t)blockIdx.x * 2 + 1) * C + i] = br;
        }
    }
}

// rep[i] += sum of the private copies
__global__ void k_sum_copies(double* __restrict__ out, const double* __restrict__ copies, size_t n, int n_copies) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int c = 0; c < n_copies; ++c) s += copies[(size_t)c * n + i];
    out[i] += s;
}

// merge per-CTA (value, row) tables into the caller's table: max value, then lowest row.
// Block (32 clusters x 8 slices of the CTA list): coalesced reads, 8-way parallel over the list.
__global__ void __launch_bounds__(256) k_merge_best(const unsigned long long* __restrict__ scratch, int n_cta, int C,
                                                    unsigned long long* __restrict__ tab) {
    __shared__ unsigned long long sv[8][32], sr[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    unsigned long long bv = 0ull, br = ~0ull;
    if (c < C) {
#pragma unroll 4
        for (int b = ty; b < n_cta; b += 8) {
            const unsigned long long v = scratch[((size_t)b * 2) * C + c], r = scratch[((size_t)b * 2 + 1) * C + c];
            if (v > bv || (v == bv && r < br)) { bv = v; br = r; }
        }
    }
    sv[ty][tx] = bv; sr[ty][tx] = br;
    __syncthreads();
    if (ty == 0 && c < C) {
        bv = tab[c]; br = tab[C + c];
        for (int g = 0; g < 8; ++g) {
            const unsigned long long v = sv[g][tx], r = sr[g][tx];
            if (v > bv || (v == bv && r < br)) { bv = v; br = r; }
        }
        tab[c] = bv;
        tab[C + c] = br;
    }
}

cudaError_t launch_assign_sparse(const unsigned long long* row_ptr, const uint16_t* pk, const double* pv,
                                 long long n_rows, long long row0, int L, const int* cid, const double* cw,
                                 int n_clusters, double thr, long long* labels, double* confs,
                                 unsigned long long* counts, unsigned long long* best, double* rep, double* rep_w,
                                 unsigned long long* site_best, int n_sms, cudaStream_t st, const long long* row_list,
                                 const unsigned long long* n_list, int slot) {
    // slot = 32: the rows were written by sitb_pass_stats_slotted with 32-entry slots (row-ordered)
    if (n_rows <= 0) return cudaSuccess;
    const int C = n_clusters > 0 ? n_clusters : 1;
    cudaError_t e;
    int warps, grid;
    size_t smem;
    {
        // per warp: 16 B per cluster and table asked for; per CTA: the histogram
        const int tables = (best ? 1 : 0) + (site_best ? 1 : 0);
        const bool ts = (size_t)L * 12 <= 40 * 1024;      // centre tables in shared memory (L <= ~3400)
        const size_t table_bytes = ts ? (size_t)L * 12 + 8 : 0;
        auto smem_for = [&](int w) { return (size_t)w * (16 * (size_t)C * tables + 32 * (size_t)C) + 8 * (size_t)((C + 1) / 2) + 8 * (size_t)C + table_bytes + 16; };
        // warps per CTA: the most resident warps per SM (64 registers: at most 32) within 227 KB of shared memory; the
        // centre tables are per CTA, the accumulators and best-row tables per warp
        warps = 1;
        int best_resident = 0;
        for (int w = 16; w >= 1; w = (w > 2 ? w - 2 : w - 1)) {
            const size_t b = smem_for(w) + 1024;
            if (b > 200 * 1024) continue;
            int r = (int)((227 * 1024) / b) * w;
            if (r > 32) r = 32;
            if (r > best_resident) { best_resident = r; warps = w; }
        }
        smem = smem_for(warps);
        if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
        const bool slotted = slot == 32;
        auto kern = ts ? (slotted ? k_assign_sparse8<true, true> : k_assign_sparse8<true, false>)
                       : (slotted ? k_assign_sparse8<false, true> : k_assign_sparse8<false, false>);
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int resident = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, warps * 32, smem);
        if (e != cudaSuccess) return e;
        if (resident < 1) resident = 1;
        const long long quads = (n_rows + 3) / 4;
        const long long want = (quads + warps - 1) / warps;
        grid = (int)(want < (long long)n_sms * resident ? want : (long long)n_sms * resident);
    }
    unsigned long long *sb = nullptr, *ss = nullptr;
    if (best) { e = cudaMallocAsync((void**)&sb, sizeof(unsigned long long) * 2 * (size_t)grid * C, st); if (e != cudaSuccess) return e; }
    if (site_best) {
        e = cudaMallocAsync((void**)&ss, sizeof(unsigned long long) * 2 * (size_t)grid * C, st);
        if (e != cudaSuccess) { if (sb) cudaFreeAsync(sb, st); return e; }
    }
    {
        double* copies = nullptr;
        const size_t rep_n = (size_t)C * L;
        int n_copies = 1;
        if (rep && n_rows > 65536) {                       // a short list (the selective re-predict) adds straight into rep
            n_copies = 8;
            while (n_copies > 1 && rep_n * n_copies * sizeof(double) > ((size_t)64 << 20)) n_copies >>= 1;
        }
        if (n_copies > 1) {
            e = cudaMallocAsync((void**)&copies, sizeof(double) * rep_n * n_copies, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(copies, 0, sizeof(double) * rep_n * n_copies, st);
            if (e != cudaSuccess) { if (sb) cudaFreeAsync(sb, st); if (ss) cudaFreeAsync(ss, st); if (copies) cudaFreeAsync(copies, st); return e; }
        }
        auto kern = (size_t)L * 12 <= 40 * 1024 ? (slot == 32 ? k_assign_sparse8<true, true> : k_assign_sparse8<true, false>)
                                                : (slot == 32 ? k_assign_sparse8<false, true> : k_assign_sparse8<false, false>);
        kern<<<grid, warps * 32, smem, st>>>(row_ptr, pk, pv, n_rows, row0, L, cid, cw, n_clusters, thr, labels, confs, counts, sb,
                                             copies ? copies : rep, rep_w, ss, row_list, n_list, n_copies);
        if (copies) {
            k_sum_copies<<<(unsigned)((rep_n + 255) / 256), 256, 0, st>>>(rep, copies, rep_n, n_copies);
            cudaFreeAsync(copies, st);
        }
    }
    if (best) { k_merge_best<<<(C + 31) / 32, 256, 0, st>>>(sb, grid, n_clusters, best); cudaFreeAsync(sb, st); }
    if (site_best) { k_merge_best<<<(C + 31) / 32, 256, 0, st>>>(ss, grid, n_clusters, site_best); cudaFreeAsync(ss, st); }
    return cudaGetLastError();
}

// ---- the min_samples filter without a second full predict (DotProdClassifier.pyx:105-118) ----------------------------
// The second predict uses the centres that survived the filter, unchanged.  A row whose first-predict cluster
// survived keeps its arg-max and its confidence (removing other centres cannot raise another one above it); an
// unassigned row stays unassigned.  Only rows of removed clusters have to be predicted again: this kernel renumbers
// the labels and lists those rows.
__global__ void k_relabel_select(long long* __restrict__ labels, long long n_rows, const int* __restrict__ remap,
                                 long long* __restrict__ row_list, unsigned long long* __restrict__ n_list) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        const long long l = labels[r];
        if (l < 0) continue;
        const int nl = remap[l];
        if (nl >= 0) {
            if (nl != l) labels[r] = nl;
        } else {
            row_list[atomicAdd(n_list, 1ull)] = r;
        }
    }
}

// |row|^2 of selected cached rows (cluster/mcl.py:86: np.linalg.norm of each cluster's best-matching landmark vector);
// rows[i] < 0 or >= n_rows (not resident on this rank) give 0, so the shards' results can be summed
__global__ void k_sparse_row_norm2(const unsigned long long* __restrict__ row_ptr, const double* __restrict__ pv,
                                   long long n_rows, const long long* __restrict__ rows, int n, double* __restrict__ out) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const long long r = rows[i];
    double s = 0.0;
    if (r >= 0 && r < n_rows) {
        const unsigned long long ptr = row_ptr[r];
        const int cnt = (int)(ptr & 0xFF);
        const unsigned long long off = ptr >> 8;
        for (int e = lane; e < cnt; e += 32) { const double v = pv[off + e]; s = fma(v, v, s); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[i] = s;
}

cudaError_t launch_sparse_row_norm2(const unsigned long long* row_ptr, const double* pv, long long n_rows,
                                    const long long* rows, int n, double* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_sparse_row_norm2<<<(n + 3) / 4, 128, 0, st>>>(row_ptr, pv, n_rows, rows, n, out);
    return cudaGetLastError();
}

cudaError_t launch_relabel_select(long long* labels, long long n_rows, const int* remap, long long* row_list,
                                  unsigned long long* n_list, int n_sms, cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    long long blocks = (n_rows + 255) / 256;
    if (blocks > (long long)n_sms * 8) blocks = (long long)n_sms * 8;
    k_relabel_select<<<(unsigned)blocks, 256, 0, st>>>(labels, n_rows, remap, row_list, n_list);
    return cudaGetLastError();
}

}  // namespace sitb
