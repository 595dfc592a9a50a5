// sitator_b200 -- the assign step shared by the fused kernel (sitb_fill.cu, MODE_ASSIGN) and the
// sparse-row kernel (sitb_sparse.cu): DotProdClassifier.predict (DotProdClassifier.pyx:166-189,
// predict_normed=False) for cluster centres with disjoint supports (cluster/mcl.py:80).
#pragma once
#include "sitb_fill.cuh"

namespace sitb {

// lock-protected lexicographic max of (value, first row): slot layout [C] value bits | [C] row | [C] lock.
// cache (optional, shared memory, [C] value bits): a per-CTA lower bound of the global maximum; values
// only grow, so a candidate below the cache is below the global table too.
__device__ __forceinline__ void best_update(unsigned long long* tab, int C, int c, double v, unsigned long long row,
                                            unsigned long long* cache = nullptr) {
    const unsigned long long vb = (unsigned long long)__double_as_longlong(v);     // v >= 0: bits are monotone
    if (cache) {
        if (vb < *((volatile unsigned long long*)(cache + c))) return;
        atomicMax(cache + c, vb);
    }
    volatile unsigned long long* val = tab + c;
    volatile unsigned long long* rw = tab + C + c;
    if (vb < *val) return;                                  // a stale read can only let us in
    unsigned* lock = (unsigned*)(tab + 2 * (size_t)C + c);
    while (atomicCAS(lock, 0u, 1u) != 0u) {}
    __threadfence();
    const unsigned long long cv = *val, cr = *rw;
    if (vb > cv || (vb == cv && row < cr)) { *val = vb; *rw = row; }
    __threadfence();
    atomicExch(lock, 0u);
}

// A row touches few clusters: peel them off one at a time (ascending id) with warp votes.
// myc/mypr: per lane, up to NCH entries (cluster id or -1, value * centre weight).
// Returns the winning cluster (np.argmax: first maximum; all-zero -> index 0) and its |dot|.
template <int NCH>
__device__ __forceinline__ void peel_clusters(int (&myc)[NCH], double (&mypr)[NCH], int lane,
                                              unsigned long long* best_tab, int n_clusters,
                                              unsigned long long row_global, double& bestc, int& bestid,
                                              unsigned long long* best_cache = nullptr) {
    bestc = 0.0;
    bestid = 0;
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
        if (conf > bestc) { bestc = conf; bestid = cur; }     // ascending ids: ties keep the lower
        if (best_tab && lane == 0) best_update(best_tab, n_clusters, cur, conf, row_global, best_cache);
    }
}

}  // namespace sitb
