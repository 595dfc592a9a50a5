// sitator_b200 -- principal eigenvector of every landmark cluster's covariance block, on the device.
//
//   cluster/mcl.py:73-80   per cluster: eigenvec = eigsh(cov[cluster][:, cluster], k=1)  ([1] for a singleton)
//
// The blocks are tiny (<= ~50 x 50) but there is one per cluster; gathering them to the host and calling LAPACK per
// block size took 3.7 ms per analysis (and two device synchronisations).  Here: one CTA per cluster, the block
// gathered from the device-resident covariance, cyclic Jacobi with the round-robin (tournament) ordering so that
// n/2 disjoint rotations are applied concurrently, float64 throughout.  Jacobi converges to working precision
// (the eigenvector agrees with LAPACK's to a few 1e-16; its sign is arbitrary in the reference too: ARPACK starts
// from a random vector, and every consumer takes the absolute value, mcl.py:82,85, DotProdClassifier.pyx:179).
#include "../../include/sitator_b200.h"
#include "sitb_common.cuh"

namespace sitb {
int set_error(int code, const char* fmt, ...);

static constexpr int EIG_NMAX = 64;       // larger blocks: the caller falls back to the host
static constexpr int EIG_THREADS = 128;
static constexpr int EIG_MAX_SWEEPS = 40;

__global__ void __launch_bounds__(EIG_THREADS) k_principal_vectors(const double* __restrict__ cov, int L,
                                                                   const int* __restrict__ members,
                                                                   const int* __restrict__ offsets, double* __restrict__ out_w,
                                                                   int* __restrict__ sweeps_out) {
    extern __shared__ double sm[];
    const int c = blockIdx.x;
    const int beg = offsets[c], n = offsets[c + 1] - beg;
    const int tid = threadIdx.x;
    if (n <= 0) return;
    if (n == 1) {
        if (tid == 0) { out_w[members[beg]] = 1.0; sweeps_out[c] = 0; }
        return;
    }
    if (n > EIG_NMAX) {
        if (tid == 0) sweeps_out[c] = -1;
        return;
    }
    const int ld = n + 1;                       // padded rows: column walks hit different banks
    double* A = sm;                             // [n][ld]
    double* V = A + (size_t)EIG_NMAX * (EIG_NMAX + 1);
    double* cs = V + (size_t)EIG_NMAX * (EIG_NMAX + 1);     // [m/2][2]
    __shared__ int s_idx[EIG_NMAX];
    __shared__ int s_rotated;
    __shared__ int s_pq[EIG_NMAX / 2 + 1];
    for (int i = tid; i < n; i += blockDim.x) s_idx[i] = members[beg + i];
    __syncthreads();
    for (int e = tid; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e - i * n;
        A[i * ld + j] = cov[(size_t)s_idx[i] * L + s_idx[j]];
        V[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int m = (n + 1) & ~1;                 // even number of players; index n (if any) is a bye
    const int half = m >> 1;
    int sweep = 0;
    for (; sweep < EIG_MAX_SWEEPS; ++sweep) {
        if (tid == 0) s_rotated = 0;
        __syncthreads();
        for (int r = 0; r < m - 1; ++r) {
            // phase 1: rotation angles of this round's disjoint pairs
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = m - 1; q = r; }
                else { p = (r + tid) % (m - 1); q = (r - tid + (m - 1)) % (m - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                double cc = 1.0, ss = 0.0;
                s_pq[tid] = p | (q << 8);
                if (q < n) {
                    const double apq = A[p * ld + q], app = A[p * ld + p], aqq = A[q * ld + q];
                    // rotate unless the off-diagonal element is already negligible against the diagonal (relative
                    // criterion of one-sided / two-sided Jacobi: the eigenvectors then carry a few eps of error)
                    if (fabs(apq) > 4.4e-16 * sqrt(fabs(app * aqq)) && fabs(apq) > 1e-300) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        cc = 1.0 / sqrt(t * t + 1.0);
                        ss = t * cc;
                        s_rotated = 1;
                    }
                }
                cs[2 * tid] = cc; cs[2 * tid + 1] = ss;
            }
            __syncthreads();
            // phase 2: rows  A <- J^T A
            for (int e = tid; e < half * n; e += blockDim.x) {
                const int t = e / n, k = e - t * n;
                const int p = s_pq[t] & 0xFF, q = s_pq[t] >> 8;
                const double cc = cs[2 * t], ss = cs[2 * t + 1];
                if (ss != 0.0) {
                    const double ap = A[p * ld + k], aq = A[q * ld + k];
                    A[p * ld + k] = cc * ap - ss * aq;
                    A[q * ld + k] = ss * ap + cc * aq;
                }
            }
            __syncthreads();
            // phase 3: columns  A <- A J,  V <- V J
            for (int e = tid; e < half * n; e += blockDim.x) {
                const int t = e / n, k = e - t * n;
                const int p = s_pq[t] & 0xFF, q = s_pq[t] >> 8;
                const double cc = cs[2 * t], ss = cs[2 * t + 1];
                if (ss != 0.0) {
                    const double ap = A[k * ld + p], aq = A[k * ld + q];
                    A[k * ld + p] = cc * ap - ss * aq;
                    A[k * ld + q] = ss * ap + cc * aq;
                    const double vp = V[k * ld + p], vq = V[k * ld + q];
                    V[k * ld + p] = cc * vp - ss * vq;
                    V[k * ld + q] = ss * vp + cc * vq;
                }
            }
            __syncthreads();
        }
        if (!s_rotated) break;
        __syncthreads();
    }
    // largest eigenvalue (first maximum), its eigenvector normalised
    if (tid < 32) {
        double bv = -1e300;
        int bj = 0;
        for (int j = tid; j < n; j += 32) {
            const double v = A[j * ld + j];
            if (v > bv) { bv = v; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
        }
        double nn = 0.0;
        for (int i = tid; i < n; i += 32) nn += V[i * ld + bj] * V[i * ld + bj];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        const double inv = 1.0 / sqrt(nn);
        for (int i = tid; i < n; i += 32) out_w[s_idx[i]] = V[i * ld + bj] * inv;
        if (tid == 0) sweeps_out[c] = sweep;
    }
}

}  // namespace sitb

using namespace sitb;

extern "C" int sitb_principal_vectors(int device, const double* dev_cov, int32_t n_landmarks, const int32_t* dev_members,
                                      const int32_t* dev_offsets, int32_t n_clusters, double* dev_weights,
                                      int32_t* dev_sweeps, void* cuda_stream) {
    if (!dev_cov || !dev_members || !dev_offsets || !dev_weights || !dev_sweeps || n_landmarks <= 0 || n_clusters < 0)
        return set_error(SITB_E_INVALID, "sitb_principal_vectors: bad argument");
    if (n_clusters == 0) return SITB_OK;
    cudaError_t e = cudaSetDevice(device);
    const size_t smem = sizeof(double) * (2 * (size_t)EIG_NMAX * (EIG_NMAX + 1) + EIG_NMAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_principal_vectors, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "sitb_principal_vectors: %s", cudaGetErrorString(e));
    k_principal_vectors<<<n_clusters, EIG_THREADS, smem, (cudaStream_t)cuda_stream>>>(dev_cov, n_landmarks, dev_members, dev_offsets,
                                                                                  dev_weights, dev_sweeps);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "sitb_principal_vectors: %s", cudaGetErrorString(e));
    return SITB_OK;
}
