// sitator_b200 -- K4: landmark graph + Markov clustering in FP64 on the device.
//
//   cluster/mcl.py:54-59   cov = Gram / N ; cov2corr (:34-41) ; clip at 0 ; unit self loops
//   util/mcl.py:3-60       markov_clustering: normalise columns, expand (matrix power),
//                          inflate (element power), normalise, prune, converge (np.allclose)
//
// Everything stays in double so that the discrete decisions (prune < 1e-5, allclose, attractor
// rows) follow the reference; only the summation order inside the GEMM differs from OpenBLAS.
// L is ~10^3: the GEMM is 2 L^3 = 7 GFLOP per iteration, a millisecond on the FP64 pipe, so a
// plain shared-memory tiled DFMA kernel is used (the FP64 tensor path would save microseconds).
#include "../../include/sitator_b200.h"
#include "sitb_common.cuh"
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdio>
#include <vector>

namespace sitb {

// ---- C = A * B, n x n row-major doubles -------------------------------------------------------
constexpr int GB = 64, GK = 16;

// 1 where a (th x tw) tile of the n x n matrix holds a non-zero (or non-finite) entry.  MCL iterates are sparse
// (a landmark correlates with few others, and pruning re-sparsifies every iteration), and a product
// tile with an all-zero operand tile adds exact zeros: skipping it changes no bit of the result.
__global__ void __launch_bounds__(256) k_tilemap(const double* __restrict__ m, int n, int th, int tw,
                                                 unsigned char* __restrict__ map) {
    const int r0 = blockIdx.y * th, c0 = blockIdx.x * tw;
    int any = 0;
    for (int i = threadIdx.x; i < th * tw; i += 256) {
        const int r = r0 + i / tw, c = c0 + i % tw;
        if (r < n && c < n && m[(size_t)r * n + c] != 0.0) any = 1;
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) map[blockIdx.y * gridDim.x + blockIdx.x] = (unsigned char)any;
}

__global__ void __launch_bounds__(256) k_dgemm(const double* __restrict__ A, const double* __restrict__ B,
                                               double* __restrict__ C, int n, const unsigned char* __restrict__ amap,
                                               const unsigned char* __restrict__ bmap) {
    __shared__ double As[GK][GB + 1];
    __shared__ double Bs[GK][GB];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row0 = blockIdx.y * GB, col0 = blockIdx.x * GB;
    const int kt = (n + GK - 1) / GK;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < n; k0 += GK) {
        // amap: [n/64][n/16] tiles of A, bmap: [n/16][n/64] tiles of B (block-uniform test)
        if (!amap[blockIdx.y * kt + k0 / GK] || !bmap[(k0 / GK) * gridDim.x + blockIdx.x]) continue;
        for (int i = threadIdx.x; i < GB * GK; i += 256) {
            const int r = i / GK, c = i % GK;                    // A tile: 64 rows x 16 k
            const int gr = row0 + r, gc = k0 + c;
            As[c][r] = (gr < n && gc < n) ? A[(size_t)gr * n + gc] : 0.0;
            const int r2 = i / GB, c2 = i % GB;                  // B tile: 16 k x 64 cols
            const int gr2 = k0 + r2, gc2 = col0 + c2;
            Bs[r2][c2] = (gr2 < n && gc2 < n) ? B[(size_t)gr2 * n + gc2] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = row0 + ty + 16 * i, c = col0 + tx + 16 * j;
            if (r < n && c < n) C[(size_t)r * n + c] = acc[i][j];
        }
}

// ---- column sums in the row order NumPy uses for axis=0 on a C-contiguous matrix ----------------
__global__ void k_colsum(const double* __restrict__ m, int n, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
#pragma unroll 32
    for (int i = 0; i < n; ++i) s = __dadd_rn(s, m[(size_t)i * n + j]);     // loads pipeline, adds stay in order
    out[j] = s;
}

__global__ void k_coldiv(double* __restrict__ m, int n, const double* __restrict__ colsum) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    m[idx] = m[idx] / colsum[idx % n];
}

__global__ void k_power(double* __restrict__ m, size_t count, double r) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    m[idx] = pow(m[idx], r);
}

// first row index of the column maximum (np.argmax(m, axis=0)); block = 32 columns x 8 interleaved row groups
__global__ void __launch_bounds__(256) k_colargmax(const double* __restrict__ m, int n, int* __restrict__ arg) {
    __shared__ double sv[8][32];
    __shared__ int si[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    double best = -CUDART_INF;
    int bi = 0x7FFFFFFF;
    if (j < n) {
#pragma unroll 8
        for (int i = ty; i < n; i += 8) {
            const double v = m[(size_t)i * n + j];
            if (v > best || bi == 0x7FFFFFFF) { best = v; bi = i; }     // NaN-free input; first maximum of this group
        }
    }
    sv[ty][tx] = best; si[ty][tx] = bi;
    __syncthreads();
    if (ty == 0 && j < n) {
        for (int g = 1; g < 8; ++g) {
            const double v = sv[g][tx];
            const int i = si[g][tx];
            if (i != 0x7FFFFFFF && (v > best || (v == best && i < bi))) { best = v; bi = i; }
        }
        arg[j] = bi;
    }
}

// prune (util/mcl.py:37-40) and compare with the previous iterate (np.allclose, :42) in one pass
__global__ void k_prune_compare(double* __restrict__ m2, const double* __restrict__ m1, int n,
                                const int* __restrict__ arg, double thr, int* __restrict__ not_close) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    double v = m2[idx];
    if (v < thr && i != arg[j]) { v = 0.0; m2[idx] = 0.0; }
    const double a = m1[idx];
    // np.allclose(a, b): |a - b| <= atol + rtol * |b|, atol 1e-8, rtol 1e-5; NaN never close
    if (!(fabs(a - v) <= 1e-8 + 1e-5 * fabs(v))) *not_close = 1;
}

// cluster/mcl.py:54-59 from the un-normalised upper-triangular Gram
__global__ void k_graph(const double* __restrict__ gram_upper, int n, double n_rows, double* __restrict__ cov,
                        double* __restrict__ graph) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const double c = gram_upper[(size_t)lo * n + hi] / n_rows;
    cov[idx] = c;
    double di = sqrt(gram_upper[(size_t)i * n + i] / n_rows);
    double dj = sqrt(gram_upper[(size_t)j * n + j] / n_rows);
    if (di == 0.0) di = CUDART_INF;
    if (dj == 0.0) dj = CUDART_INF;
    double g = (c / di) / dj;                 // ((A.T/d).T)/d
    g = fmax(g, 0.0);                         // np.clip(corr, 0, None)
    if (i == j && g == 0.0) g = 1.0;          // never-seen landmark: self loop (:57-59)
    graph[idx] = g;
}

}  // namespace sitb

using namespace sitb;

namespace sitb { int set_error(int code, const char* fmt, ...); }

#define CKM(call)                                                                  \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            rc = sitb::set_error(SITB_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
            goto done;                                                             \
        }                                                                          \
    } while (0)

extern "C" int sitb_landmark_graph(int device, const double* dev_gram_upper, int32_t n, double n_rows,
                                   double* dev_cov, double* dev_graph, void* cuda_stream) {
    int rc = SITB_OK;
    if (!dev_gram_upper || !dev_cov || !dev_graph || n <= 0 || !(n_rows > 0))
        return sitb::set_error(SITB_E_INVALID, "sitb_landmark_graph: bad argument");
    {
        CKM(cudaSetDevice(device));
        const size_t cnt = (size_t)n * n;
        k_graph<<<(unsigned)((cnt + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(dev_gram_upper, n, n_rows, dev_cov, dev_graph);
        CKM(cudaGetLastError());
    }
done:
    return rc;
}

extern "C" int sitb_markov_clustering(int device, const double* dev_graph, int32_t n, int32_t expansion,
                                      double inflation, double pruning_threshold, int32_t iterlimit,
                                      double* dev_result, int32_t* n_iterations, int32_t* converged,
                                      void* cuda_stream) {
    int rc = SITB_OK;
    if (!dev_graph || !dev_result || n <= 0 || expansion < 1 || iterlimit < 1)
        return sitb::set_error(SITB_E_INVALID, "sitb_markov_clustering: bad argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t cnt = (size_t)n * n, bytes = cnt * sizeof(double);
    double *m1 = nullptr, *m2 = nullptr, *z = nullptr, *tmp = nullptr, *colsum = nullptr;
    int *arg = nullptr, *flag = nullptr;
    unsigned char *amap = nullptr, *bmap = nullptr;
    const unsigned eb = (unsigned)((cnt + 255) / 256), cb = (unsigned)((n + 31) / 32);
    const dim3 gg((n + GB - 1) / GB, (n + GB - 1) / GB);
    int it = 0, conv = 0;
    {
        CKM(cudaSetDevice(device));
        CKM(sitb::pool_alloc((void**)&m1, bytes, st));
        CKM(sitb::pool_alloc((void**)&z, bytes, st));
        CKM(sitb::pool_alloc((void**)&tmp, bytes, st));
        CKM(sitb::pool_alloc((void**)&colsum, sizeof(double) * n, st));
        CKM(sitb::pool_alloc((void**)&arg, sizeof(int) * n, st));
        CKM(sitb::pool_alloc((void**)&flag, sizeof(int), st));
        const int kt = (n + GK - 1) / GK;
        CKM(sitb::pool_alloc((void**)&amap, (size_t)gg.y * kt, st));
        CKM(sitb::pool_alloc((void**)&bmap, (size_t)kt * gg.x, st));
        auto gemm = [&](const double* X, const double* Y, double* Z) {
            k_tilemap<<<dim3(kt, gg.y), 256, 0, st>>>(X, n, GB, GK, amap);
            k_tilemap<<<dim3(gg.x, kt), 256, 0, st>>>(Y, n, GK, GB, bmap);
            k_dgemm<<<gg, 256, 0, st>>>(X, Y, Z, n, amap, bmap);
        };
        m2 = dev_result;
        // m1 = graph / colsum (util/mcl.py:22-25)
        CKM(cudaMemcpyAsync(m1, dev_graph, bytes, cudaMemcpyDeviceToDevice, st));
        k_colsum<<<cb, 32, 0, st>>>(m1, n, colsum);
        k_coldiv<<<eb, 256, 0, st>>>(m1, n, colsum);
        for (it = 0; it < iterlimit; ++it) {
            // expansion: np.linalg.matrix_power(m1, expansion) with NumPy's multiplication order
            if (expansion == 1) {
                CKM(cudaMemcpyAsync(m2, m1, bytes, cudaMemcpyDeviceToDevice, st));
            } else if (expansion == 2) {
                gemm(m1, m1, m2);
            } else if (expansion == 3) {
                gemm(m1, m1, tmp);
                gemm(tmp, m1, m2);
            } else {
                // binary decomposition: z = a, a^2, a^4, ...; result *= z for set bits
                int e = expansion;
                bool have_z = false, have_r = false;
                double* zc = z;      // current power
                double* zn = tmp;    // scratch
                double* res = m2;
                std::vector<double*> spare;
                double* res_tmp = nullptr;
                CKM(sitb::pool_alloc((void**)&res_tmp, bytes, st));
                while (e > 0) {
                    if (!have_z) { CKM(cudaMemcpyAsync(zc, m1, bytes, cudaMemcpyDeviceToDevice, st)); have_z = true; }
                    else { gemm(zc, zc, zn); double* t = zc; zc = zn; zn = t; }
                    const int bit = e & 1;
                    e >>= 1;
                    if (bit) {
                        if (!have_r) { CKM(cudaMemcpyAsync(res, zc, bytes, cudaMemcpyDeviceToDevice, st)); have_r = true; }
                        else { gemm(res, zc, res_tmp); CKM(cudaMemcpyAsync(res, res_tmp, bytes, cudaMemcpyDeviceToDevice, st)); }
                    }
                }
                CKM(cudaStreamSynchronize(st));
                sitb::pool_free(res_tmp, st);
            }
            k_power<<<eb, 256, 0, st>>>(m2, cnt, inflation);                 // :34
            k_colsum<<<cb, 32, 0, st>>>(m2, n, colsum);                      // :35
            k_coldiv<<<eb, 256, 0, st>>>(m2, n, colsum);
            k_colargmax<<<cb, 256, 0, st>>>(m2, n, arg);                      // :39
            CKM(cudaMemsetAsync(flag, 0, sizeof(int), st));
            k_prune_compare<<<eb, 256, 0, st>>>(m2, m1, n, arg, pruning_threshold, flag);   // :37-42
            int h_flag = 1;
            CKM(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
            CKM(cudaStreamSynchronize(st));
            CKM(cudaGetLastError());
            if (!h_flag) { conv = 1; ++it; break; }
            CKM(cudaMemcpyAsync(m1, m2, bytes, cudaMemcpyDeviceToDevice, st));   // :46
        }
        CKM(cudaStreamSynchronize(st));
    }
done:
    sitb::pool_free(m1, st); sitb::pool_free(z, st); sitb::pool_free(tmp, st); sitb::pool_free(colsum, st); sitb::pool_free(arg, st); sitb::pool_free(flag, st);
    sitb::pool_free(amap, st); sitb::pool_free(bmap, st);
    if (n_iterations) *n_iterations = it;
    if (converged) *converged = conv;
    return rc;
}
