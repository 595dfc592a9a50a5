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

// Per-warp private (value bits, row) tables in shared memory: a warp walks its rows in ascending order,
// so "first row among equal values" is simply "do not replace on equality"; no atomics, no locks.
struct WarpBest {
    unsigned long long* val;     // [C] of this warp
    unsigned long long* row;
    __device__ __forceinline__ void update(int c, double v, unsigned long long r) {
        const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
        if (vb > val[c]) { val[c] = vb; row[c] = r; }
    }
};

template <int NCH>
__device__ __forceinline__ void assign_row(
    long long r, int nent, unsigned long long off, int lane, const uint16_t* __restrict__ pk,
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
                const int k = pk[off + e];
                const double v = pv[off + e];
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

__global__ void __launch_bounds__(256) k_assign_sparse(
    const unsigned long long* __restrict__ row_ptr, const uint16_t* __restrict__ pk, const double* __restrict__ pv,
    long long n_rows, long long row0, int L, const int* __restrict__ cid, const double* __restrict__ cw,
    int n_clusters, double thr, long long* __restrict__ labels, double* __restrict__ confs,
    unsigned long long* __restrict__ counts, unsigned long long* __restrict__ best_scratch, double* __restrict__ rep,
    double* __restrict__ rep_w, unsigned long long* __restrict__ site_scratch) {
    // shared: per warp [C] best val | [C] best row | [C] site val | [C] site row ; then [C] hist
    extern __shared__ unsigned long long smem_u64[];
    const int C = n_clusters;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned long long* wbase = smem_u64 + (size_t)warp * 4 * C;
    WarpBest wb = {wbase, wbase + C};
    WarpBest ws = {wbase + 2 * (size_t)C, wbase + 3 * (size_t)C};
    unsigned* hist = (unsigned*)(smem_u64 + (size_t)nwarps * 4 * C);
    const long long warp_global = (long long)blockIdx.x * nwarps + warp;
    const long long n_warps = (long long)gridDim.x * nwarps;
    for (int i = threadIdx.x; i < nwarps * 4 * C; i += blockDim.x) smem_u64[i] = 0ull;
    for (int i = threadIdx.x; i < C; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (long long r = warp_global; r < n_rows; r += n_warps) {
        const unsigned long long ptr = row_ptr[r];
        const int nent = (int)(ptr & 0xFF);
        const unsigned long long off = ptr >> 8;
        if (nent <= 32)          // the common case: one entry per lane
            assign_row<1>(r, nent, off, lane, pk, pv, row0, L, cid, cw, thr, labels, confs, counts, hist, best_scratch,
                          wb, rep, rep_w, site_scratch, ws);
        else
            assign_row<ENTRY_CAP / 32>(r, nent, off, lane, pk, pv, row0, L, cid, cw, thr, labels, confs, counts, hist,
                                       best_scratch, wb, rep, rep_w, site_scratch, ws);
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
                                 unsigned long long* site_best, int n_sms, cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    const int C = n_clusters > 0 ? n_clusters : 1;
    int warps = 8;                                            // per-warp tables: 32 B per cluster and warp
    while (warps > 1 && (32 * (size_t)warps + 4) * C > 160 * 1024) warps >>= 1;
    const size_t smem = (32 * (size_t)warps + 4) * C;
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k_assign_sparse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = n_sms * 8;
    unsigned long long *sb = nullptr, *ss = nullptr;
    if (best) { e = cudaMallocAsync((void**)&sb, sizeof(unsigned long long) * 2 * (size_t)grid * C, st); if (e != cudaSuccess) return e; }
    if (site_best) { e = cudaMallocAsync((void**)&ss, sizeof(unsigned long long) * 2 * (size_t)grid * C, st); if (e != cudaSuccess) return e; }
    k_assign_sparse<<<grid, warps * 32, smem, st>>>(row_ptr, pk, pv, n_rows, row0, L, cid, cw, n_clusters, thr, labels, confs,
                                            counts, sb, rep, rep_w, ss);
    if (best) { k_merge_best<<<(C + 31) / 32, 256, 0, st>>>(sb, grid, n_clusters, best); cudaFreeAsync(sb, st); }
    if (site_best) { k_merge_best<<<(C + 31) / 32, 256, 0, st>>>(ss, grid, n_clusters, site_best); cudaFreeAsync(ss, st); }
    return cudaGetLastError();
}

}  // namespace sitb
