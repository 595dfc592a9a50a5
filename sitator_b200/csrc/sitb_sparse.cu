// sitator_b200 -- K3s: the assign step over cached sparse landmark vectors.
//
// Landmark vectors are ~1.5 % dense (22 of 1500 components), so the fill pass that builds the Gram
// also stores every row compressed (sitb_fill.cu, sparse_ptr/sparse_k/sparse_v: ~230 B per row
// instead of 12 KB).  The clustering plugin's later passes over "all landmark vectors"
//   cluster/mcl.py:81-83          best-matching row per cluster
//   DotProdClassifier.pyx:86-118  predict, bincount, predict again
//   cluster/mcl.py:118-122        representative landmark vectors
// then stream those rows from HBM (one warp per row) instead of recomputing them: they become
// bandwidth-bound passes of ~1 ms per 5.6e6 rows.
#include "../../include/sitator_b200.h"
#include "sitb_assign.cuh"

namespace sitb {

__global__ void __launch_bounds__(256) k_assign_sparse(
    const unsigned long long* __restrict__ row_ptr, const uint16_t* __restrict__ pk, const double* __restrict__ pv,
    long long n_rows, long long row0, int L, const int* __restrict__ cid, const double* __restrict__ cw,
    int n_clusters, double thr, long long* __restrict__ labels, double* __restrict__ confs,
    unsigned long long* __restrict__ counts, unsigned long long* __restrict__ best, double* __restrict__ rep,
    double* __restrict__ rep_w, unsigned long long* __restrict__ site_best) {
    extern __shared__ unsigned hist[];
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    if (counts)
        for (int i = threadIdx.x; i < n_clusters; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (long long r = warp_global; r < n_rows; r += n_warps) {
        const unsigned long long ptr = row_ptr[r];
        const int nent = (int)(ptr & 0xFF);
        const unsigned long long off = ptr >> 8;
        int myc[ENTRY_CAP / 32];
        double mypr[ENTRY_CAP / 32], myv[ENTRY_CAP / 32];
        int myk[ENTRY_CAP / 32];
#pragma unroll
        for (int c = 0; c < ENTRY_CAP / 32; ++c) {
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
        double bestc;
        int bestid;
        peel_clusters<ENTRY_CAP / 32>(myc, mypr, lane, best, n_clusters, (unsigned long long)(row0 + r), bestc, bestid);
        long long label = bestid;
        double conf = bestc;
        if (nent == 0 || !(conf >= thr)) { label = -1; conf = 0.0; }     // DotProdClassifier.pyx:168-172,184-186
        if (lane == 0) {
            if (labels) labels[r] = label;
            if (confs) confs[r] = conf;
            if (label >= 0) {
                if (counts) atomicAdd(&hist[label], 1u);
                if (rep_w) atomicAdd(&rep_w[label], conf);
                if (site_best) best_update(site_best, n_clusters, (int)label, conf, (unsigned long long)(row0 + r));
            }
        }
        if (rep && label >= 0) {
#pragma unroll
            for (int c = 0; c < ENTRY_CAP / 32; ++c)
                if (32 * c + lane < nent) atomicAdd(&rep[(size_t)label * L + myk[c]], conf * myv[c]);
        }
    }
    if (counts) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_clusters; i += blockDim.x)
            if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
    }
}

cudaError_t launch_assign_sparse(const unsigned long long* row_ptr, const uint16_t* pk, const double* pv,
                                 long long n_rows, long long row0, int L, const int* cid, const double* cw,
                                 int n_clusters, double thr, long long* labels, double* confs,
                                 unsigned long long* counts, unsigned long long* best, double* rep, double* rep_w,
                                 unsigned long long* site_best, int n_sms, cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    const size_t smem = sizeof(unsigned) * (size_t)(n_clusters > 0 ? n_clusters : 1);
    k_assign_sparse<<<n_sms * 8, 256, smem, st>>>(row_ptr, pk, pv, n_rows, row0, L, cid, cw, n_clusters, thr, labels,
                                                 confs, counts, best, rep, rep_w, site_best);
    return cudaGetLastError();
}

}  // namespace sitb
