// sitator_b200 -- landmark tables (once per analysis).
//
//   LandmarkAnalysis.py:194-202   verts_np (-1 padded) and site_vert_dists = distance from each
//                                 landmark centre to each of its vertex atoms at their ideal
//                                 positions (PBCCalculator.distances, PBCCalculator.pyx:64-103)
//   helpers.pyx:197-203           the cut-off test  dist/site_vert_dist > cutoff_round_to_zero
//
// The test is turned into a compare on the squared distance: with IEEE sqrt and divide both
// monotone, there is a largest double T with T/svd <= cutoff and a largest double Q with
// sqrt(Q) <= T; then  sqrt(q)/svd > cutoff  <=>  q > Q  for every double q, bit for bit.
#include "sitb_fill.cuh"
#include <math_constants.h>

namespace sitb {

__global__ void k_tables(Cell cell, const double* __restrict__ centers, const double* __restrict__ ideal,
                         const int* __restrict__ verts_in, int L, int V, int Lpad, int S, double cutoff,
                         double steep_log2e, double* __restrict__ svd_out, uint16_t* __restrict__ verts,
                         float* __restrict__ qf, double* __restrict__ q64, double* __restrict__ acoef) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= V * Lpad) return;
    const int h = idx / Lpad, k = idx % Lpad;
    int v = -1;
    if (k < L) {
        v = verts_in[k * V + h];
        // the reference stops at the first -1 (helpers.pyx:192-193)
        for (int hh = 0; hh < h; ++hh)
            if (verts_in[k * V + hh] < 0) v = -1;
    }
    if (v < 0 || v >= S) {
        verts[idx] = VERT_END;
        qf[idx] = CUDART_INF_F;
        q64[idx] = CUDART_INF;
        acoef[idx] = 0.0;
        if (k < L) svd_out[k * V + h] = CUDART_NAN;
        return;
    }
    const double ox = __dsub_rn(cell.cen[0], centers[3 * k + 0]);
    const double oy = __dsub_rn(cell.cen[1], centers[3 * k + 1]);
    const double oz = __dsub_rn(cell.cen[2], centers[3 * k + 2]);
    const double q = cell.diag
        ? shifted_dist2<true, false>(cell, ideal[3 * v], ideal[3 * v + 1], ideal[3 * v + 2], ox, oy, oz)
        : shifted_dist2<false, false>(cell, ideal[3 * v], ideal[3 * v + 1], ideal[3 * v + 2], ox, oy, oz);
    const double svd = __dsqrt_rn(q);
    svd_out[k * V + h] = svd;
    verts[idx] = (uint16_t)v;
    double Q;
    if (!(svd > 0.0) || !(cutoff > 0.0)) {
        Q = -1.0;                       // degenerate landmark: every ratio is inf/nan -> never counted
        acoef[idx] = 0.0;
    } else {
        double t = __dmul_rn(cutoff, svd);
        for (int it = 0; it < 64 && __ddiv_rn(t, svd) > cutoff; ++it) t = nextafter(t, 0.0);
        for (int it = 0; it < 64 && __ddiv_rn(nextafter(t, CUDART_INF), svd) <= cutoff; ++it) t = nextafter(t, CUDART_INF);
        Q = __dmul_rn(t, t);
        for (int it = 0; it < 64 && __dsqrt_rn(Q) > t; ++it) Q = nextafter(Q, 0.0);
        for (int it = 0; it < 64 && __dsqrt_rn(nextafter(Q, CUDART_INF)) <= t; ++it) Q = nextafter(Q, CUDART_INF);
        acoef[idx] = steep_log2e / svd;
    }
    q64[idx] = Q;
    qf[idx] = __double2float_rn(Q);
}

cudaError_t launch_tables(const Cell& cell, const double* centers, const double* ideal, const int* verts_in, int L,
                          int V, int Lpad, int S, double cutoff, double steep_log2e, double* svd_out,
                          uint16_t* verts, float* qf, double* q64, double* acoef, cudaStream_t stream) {
    const int n = V * Lpad;
    k_tables<<<(n + 127) / 128, 128, 0, stream>>>(cell, centers, ideal, verts_in, L, V, Lpad, S, cutoff, steep_log2e,
                                                 svd_out, verts, qf, q64, acoef);
    return cudaGetLastError();
}

}  // namespace sitb
