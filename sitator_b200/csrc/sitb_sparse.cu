// sitator_b200 -- K3s: the assign step over cached sparse landmark vectors.
//
// Landmark vectors are ~1.5 % dense (22 of 1500 components), so the fill pass that builds the Gram
// also stores every row compressed (sitb_fill.cu, sparse_ptr/sparse_k/sparse_v: ~230 B per row
// instead of 12 KB).  The clustering plugin's later passes over "all landmark vectors"
//   cluster/mcl.py:81-83          best-matching row per cluster
//   DotProdClassifier.pyx:86-118  predict, bincount, predict again
//   cluster/mcl.py:118-122        representative landmark vectors
// then stream those rows from HBM (one warp per row) instead of recomputing them.
//
// "Best row" reductions (max value, first row) are kept per warp in shared memory, merged per CTA and
// then over CTAs by a second tiny kernel (no atomics or locks on the hot path).
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

// ---- lane-per-row assign ------------------------------------------------------------------------------------
// A warp takes 32 consecutive rows.  Staging: row by row (16 rows' loads in flight), lane j loads entry j
// (coalesced) into the row's padded line of shared memory.  Then every lane walks ITS row out of shared memory
// and keeps the row's per-cluster sums in LR_SLOTS register slots.  No shuffles or warp reductions per row:
// ~110 warp instructions per row instead of ~300 for the warp-per-row peel loop (ncu: that one issues 70 % of
// its cycles, so instructions are its cost).
// A row with more than 32 entries or more than LR_SLOTS clusters is "hard" (a property of the row alone, so a
// row takes the same path -- and sums in the same order -- however the trajectory is sharded); the warp handles
// those afterwards with assign_row.
// Measured alternatives (100 000 LLZO frames, passes B / C / D of scripts/profile_run.py, ms): warp-per-row
// 3.15 / 2.12 / 3.27; this kernel 2.81 / 1.83 / 2.72; the same with the centre-table look-ups moved into the
// staging step, rows up to 64 entries in the lane path and a row-wise sweep for the representative vectors
// 2.86 / 1.93 / 3.41.
static constexpr int LR_SLOTS = 12;
static constexpr int LR_KSTRIDE = 34;      // uint16 per staged row (17 words: conflict-free lane-per-row reads)
static constexpr int LR_VSTRIDE = 33;      // doubles per staged row
static constexpr size_t LR_STAGE_BYTES = 32 * LR_VSTRIDE * sizeof(double) + 32 * LR_KSTRIDE * sizeof(uint16_t) + 16;

// (value, first row) table updates from many lanes at once: raise the value with atomicMax and invalidate the row
// if this lane raised it; after a __syncwarp every lane that holds the final value atomicMins its row.
__device__ __forceinline__ void table_raise(WarpBest& t, int c, unsigned long long vb) {
    const unsigned long long old = atomicMax(&t.val[c], vb);
    if (old < vb) t.row[c] = ~0ull;
}
__device__ __forceinline__ void table_claim(WarpBest& t, int c, unsigned long long vb, unsigned long long r) {
    if (vb != 0ull && vb == t.val[c]) atomicMin(&t.row[c], r);
}

// add one entry's product to the lane's slot of its cluster (slots fill in order: the first free one appends)
__device__ __forceinline__ bool slot_add(int (&sid)[LR_SLOTS], double (&ssum)[LR_SLOTS], int c, double pr) {
    bool done = false;
#pragma unroll
    for (int q = 0; q < LR_SLOTS; ++q) {
        const bool hit = !done && (sid[q] == c || sid[q] < 0);
        if (hit) { sid[q] = c; ssum[q] += pr; done = true; }
    }
    return done;
}

__global__ void __launch_bounds__(128) k_assign_sparse(
    const unsigned long long* __restrict__ row_ptr, const uint16_t* __restrict__ pk, const double* __restrict__ pv,
    long long n_rows, long long row0, int L, const int* __restrict__ cid, const double* __restrict__ cw,
    int n_clusters, double thr, long long* __restrict__ labels, double* __restrict__ confs,
    unsigned long long* __restrict__ counts, unsigned long long* __restrict__ best_scratch, double* __restrict__ rep,
    double* __restrict__ rep_w, unsigned long long* __restrict__ site_scratch,
    const long long* __restrict__ row_list, const unsigned long long* __restrict__ n_list) {
    // row_list / n_list (optional): only these rows, their number read on the device (the selective re-predict after
    // the min_samples filter, DotProdClassifier.pyx:105-118)
    if (n_list) n_rows = (long long)*n_list;
    // shared: per warp [C] best val | [C] best row | [C] site val | [C] site row ; [C] hist ; per warp staging
    extern __shared__ unsigned long long smem_u64[];
    const int C = n_clusters;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned long long* wbase = smem_u64 + (size_t)warp * 4 * C;
    WarpBest wb = {wbase, wbase + C};
    WarpBest ws = {wbase + 2 * (size_t)C, wbase + 3 * (size_t)C};
    unsigned* hist = (unsigned*)(smem_u64 + (size_t)nwarps * 4 * C);
    unsigned char* stage = (unsigned char*)(smem_u64 + (size_t)nwarps * 4 * C + (C + 1) / 2) + (size_t)warp * LR_STAGE_BYTES;
    double* sv = (double*)stage;
    uint16_t* sk = (uint16_t*)(stage + 32 * LR_VSTRIDE * sizeof(double));
    const long long warp_global = (long long)blockIdx.x * nwarps + warp;
    const long long n_warps = (long long)gridDim.x * nwarps;
    for (int i = threadIdx.x; i < nwarps * 4 * C; i += blockDim.x) smem_u64[i] = 0ull;
    for (int i = threadIdx.x; i < C; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const long long n_groups = (n_rows + 31) >> 5;
    for (long long g = warp_global; g < n_groups; g += n_warps) {
        const long long ri = (g << 5) + lane;
        const bool valid = ri < n_rows;
        const long long r = (row_list && valid) ? row_list[ri] : ri;
        const unsigned long long ptr = valid ? row_ptr[r] : 0ull;
        const int nent = (int)(ptr & 0xFF);
        bool hard = nent > 32;
        const int ne = hard ? 0 : nent;
        // stage the 32 rows.  Loads of 16 rows are issued back to back, unconditionally (idle lanes read entry 0),
        // and stored afterwards: a conditional load per row would serialise the 32 rows on DRAM latency.
        __syncwarp();
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += 16) {
            uint16_t kk[16];
            double vv[16];
            unsigned on = 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const unsigned long long p = __shfl_sync(0xffffffffu, ptr, i0 + j);
                const int n = (int)(p & 0xFF);
                const bool act = n <= 32 && lane < n;
                const unsigned long long a = act ? (p >> 8) + lane : 0ull;
                kk[j] = __ldcs(pk + a);
                vv[j] = __ldcs(pv + a);
                on |= (act ? 1u : 0u) << j;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (on & (1u << j)) {
                    sk[(i0 + j) * LR_KSTRIDE + lane] = kk[j];
                    sv[(i0 + j) * LR_VSTRIDE + lane] = vv[j];
                }
        }
        __syncwarp();
        // every lane: its row's per-cluster sums, entries in ascending landmark order
        int sid[LR_SLOTS];
        double ssum[LR_SLOTS];
#pragma unroll
        for (int q = 0; q < LR_SLOTS; ++q) { sid[q] = -1; ssum[q] = 0.0; }
        const int maxn = __reduce_max_sync(0xffffffffu, ne);
        const uint16_t* myk = sk + lane * LR_KSTRIDE;
        const double* myv = sv + lane * LR_VSTRIDE;
        for (int e0 = 0; e0 < maxn; e0 += 4) {
            // four entries' table look-ups in flight, then their slot updates in entry order
            int cq[4];
            double pq[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool act = e0 + j < ne;
                const int k = act ? (int)myk[e0 + j] : 0;
                const double v = act ? myv[e0 + j] : 0.0;
                const int c = __ldg(cid + k);
                cq[j] = act ? c : -1;
                pq[j] = v * __ldg(cw + k);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (cq[j] >= 0 && !slot_add(sid, ssum, cq[j], pq[j])) hard = true;
        }
        const bool easy = valid && !hard;
        // arg-max in ascending cluster id (np.argmax: first maximum; all zero -> index 0)
        double bestc = 0.0;
        int bestid = 0;
#pragma unroll
        for (int q = 0; q < LR_SLOTS; ++q) {
            const double cf = fabs(ssum[q]);
            if (sid[q] >= 0 && (cf > bestc || (cf == bestc && cf > 0.0 && sid[q] < bestid))) { bestc = cf; bestid = sid[q]; }
        }
        long long label = bestid;
        double conf = bestc;
        if (ne == 0 || !(conf >= thr)) { label = -1; conf = 0.0; }       // DotProdClassifier.pyx:168-172,184-186
        if (easy) {
            if (labels) labels[r] = label;
            if (confs) confs[r] = conf;
            if (label >= 0) {
                if (counts) atomicAdd(&hist[label], 1u);
                if (rep_w) atomicAdd(&rep_w[label], conf);
            }
        }
        if (best_scratch) {                                               // cluster/mcl.py:81-83: every cluster of the row
#pragma unroll
            for (int q = 0; q < LR_SLOTS; ++q)
                if (easy && sid[q] >= 0) table_raise(wb, sid[q], (unsigned long long)__double_as_longlong(fabs(ssum[q])));
            __syncwarp();
#pragma unroll
            for (int q = 0; q < LR_SLOTS; ++q)
                if (easy && sid[q] >= 0)
                    table_claim(wb, sid[q], (unsigned long long)__double_as_longlong(fabs(ssum[q])), (unsigned long long)(row0 + r));
            __syncwarp();
        }
        if (site_scratch) {
            const unsigned long long vb = (unsigned long long)__double_as_longlong(conf);
            if (easy && label >= 0) table_raise(ws, (int)label, vb);
            __syncwarp();
            if (easy && label >= 0) table_claim(ws, (int)label, vb, (unsigned long long)(row0 + r));
            __syncwarp();
        }
        if (rep) {                                                        // mcl.py:118-122
            for (int e = 0; e < maxn; ++e)
                if (easy && label >= 0 && e < ne) atomicAdd(&rep[(size_t)label * L + myk[e]], conf * myv[e]);
        }
        // the hard rows of the group, one at a time by the whole warp
        unsigned hm = __ballot_sync(0xffffffffu, valid && hard);
        while (hm) {
            const int i = __ffs(hm) - 1;
            hm &= hm - 1;
            const unsigned long long p = __shfl_sync(0xffffffffu, ptr, i);
            const int n = (int)(p & 0xFF);
            const unsigned long long off = p >> 8;
            int k0 = 0;
            double v0 = 0.0;
            if (lane < n) { k0 = pk[off + lane]; v0 = pv[off + lane]; }
            const long long rh = __shfl_sync(0xffffffffu, r, i);
            if (n <= 32)
                assign_row<1>(rh, n, off, lane, k0, v0, pk, pv, row0, L, cid, cw, thr, labels, confs, counts, hist,
                              best_scratch, wb, rep, rep_w, site_scratch, ws);
            else if (n <= 128)
                assign_row<4>(rh, n, off, lane, k0, v0, pk, pv, row0, L, cid, cw, thr, labels, confs,
                              counts, hist, best_scratch, wb, rep, rep_w, site_scratch, ws);
            else
                assign_row<ENTRY_CAP / 32>(rh, n, off, lane, k0, v0, pk, pv, row0, L, cid, cw, thr, labels, confs,
                                           counts, hist, best_scratch, wb, rep, rep_w, site_scratch, ws);
            __syncwarp();
        }
    }
    __syncthreads();
    // merge the warps' tables (max value, then lowest row) and hand the CTA's table to the merge kernel
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (counts && hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
        for (int t = 0; t < 2; ++t) {
            unsigned long long* scratch = t == 0 ? best_scratch : site_scratch;
            if (!scratch) continue;
            unsigned long long bv = 0ull, br = 0ull;
            for (int w = 0; w < nwarps; ++w) {
                const unsigned long long v = smem_u64[(size_t)w * 4 * C + 2 * t * C + i];
                const unsigned long long rr = smem_u64[(size_t)w * 4 * C + (2 * t + 1) * C + i];
                if (v > bv || (v == bv && rr < br)) { bv = v; br = rr; }
            }
            scratch[((size_t)blockIdx.x * 2) * C + i] = bv;
            scratch[((size_t)blockIdx.x * 2 + 1) * C + i] = br;
        }
    }
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
                                 const unsigned long long* n_list) {
    if (n_rows <= 0) return cudaSuccess;
    const int C = n_clusters > 0 ? n_clusters : 1;
    // per warp: 32 B of tables per cluster + the staging block; per CTA: the histogram
    auto smem_for = [&](int w) { return (size_t)w * (32 * (size_t)C + LR_STAGE_BYTES) + 8 * (size_t)((C + 1) / 2) + 16; };
    int warps = 4;
    while (warps > 1 && smem_for(warps) > 56 * 1024) warps >>= 1;       // small enough for several CTAs per SM
    const size_t smem = smem_for(warps);
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k_assign_sparse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int resident = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_assign_sparse, warps * 32, smem);
    if (e != cudaSuccess) return e;
    if (resident < 1) resident = 1;
    const long long groups = (n_rows + 31) / 32;
    long long want = (groups + warps - 1) / warps;
    const int grid = (int)(want < (long long)n_sms * resident ? want : (long long)n_sms * resident);
    unsigned long long *sb = nullptr, *ss = nullptr;
    if (best) { e = cudaMallocAsync((void**)&sb, sizeof(unsigned long long) * 2 * (size_t)grid * C, st); if (e != cudaSuccess) return e; }
    if (site_best) {
        e = cudaMallocAsync((void**)&ss, sizeof(unsigned long long) * 2 * (size_t)grid * C, st);
        if (e != cudaSuccess) { if (sb) cudaFreeAsync(sb, st); return e; }
    }
    k_assign_sparse<<<grid, warps * 32, smem, st>>>(row_ptr, pk, pv, n_rows, row0, L, cid, cw, n_clusters, thr, labels, confs,
                                            counts, sb, rep, rep_w, ss, row_list, n_list);
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
