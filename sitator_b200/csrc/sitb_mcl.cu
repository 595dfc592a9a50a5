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
// The additions of one column must stay in row order (bit-compatible with NumPy's reduction), so a column is
// one thread's serial chain; what is parallel is the fetching: 8 warps bring 64 rows x 32 columns at a time
// into shared memory (double-buffered), warp 0 then adds them in order.
constexpr int CS_ROWS = 64;
__global__ void __launch_bounds__(256) k_colsum(const double* __restrict__ m, int n, double* __restrict__ out) {
    __shared__ double buf[2][CS_ROWS][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    const bool live = j < n;
    double s = 0.0;
    const int n_chunks = (n + CS_ROWS - 1) / CS_ROWS;
    auto fetch = [&](int chunk, int first, int step) {
        double (*b)[32] = buf[chunk & 1];
        const int r0 = chunk * CS_ROWS;
#pragma unroll 8
        for (int r = first; r < CS_ROWS; r += step) {
            const int i = r0 + r;
            b[r][tx] = (live && i < n) ? m[(size_t)i * n + j] : 0.0;
        }
    };
    fetch(0, ty, 8);
    __syncthreads();
    for (int c = 0; c < n_chunks; ++c) {
        if (ty != 0) {
            if (c + 1 < n_chunks) fetch(c + 1, ty - 1, 7);      // warps 1..7 fetch while warp 0 adds
        } else {
            const double (*b)[32] = buf[c & 1];
            const int rows = (n - c * CS_ROWS < CS_ROWS) ? (n - c * CS_ROWS) : CS_ROWS;
#pragma unroll 8
            for (int r = 0; r < rows; ++r) s = __dadd_rn(s, b[r][tx]);
        }
        __syncthreads();
    }
    if (ty == 0 && live) out[j] = s;
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
// flags[0] = "not close", flags[1] += number of non-zero entries left in m2 (picks the product kernel of the next
// iteration)
__global__ void k_prune_compare(double* __restrict__ m2, const double* __restrict__ m1, int n,
                                const int* __restrict__ arg, double thr, unsigned long long* __restrict__ flags) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool nz = false;
    if (idx < (size_t)n * n) {
        const int i = (int)(idx / n), j = (int)(idx % n);
        double v = m2[idx];
        if (v < thr && i != arg[j]) { v = 0.0; m2[idx] = 0.0; }
        const double a = m1[idx];
        // np.allclose(a, b): |a - b| <= atol + rtol * |b|, atol 1e-8, rtol 1e-5; NaN never close
        if (!(fabs(a - v) <= 1e-8 + 1e-5 * fabs(v))) flags[0] = 1ull;
        nz = v != 0.0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&flags[1], (unsigned long long)__popc(m));
}

__global__ void k_count_nonzero(const double* __restrict__ m, size_t count, unsigned long long* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned b = __ballot_sync(0xffffffffu, idx < count && m[idx] != 0.0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, (unsigned long long)__popc(b));
}

// C = A * B for a sparse A (MCL iterates after pruning: a few dozen entries per row).  One CTA per row i of A:
// the row's non-zeros are compacted in ascending k, 256 at a time, and every thread accumulates its columns
// over that list -- the same fma chain in the same k order as k_dgemm (a zero a_ik leaves an fma chain
// unchanged), so both kernels give bit-identical products.
constexpr int SP_COLS = 8;          // columns per thread and pass: 2048 columns per pass
__global__ void __launch_bounds__(256) k_spgemm_rows(const double* __restrict__ A, const double* __restrict__ B,
                                                     double* __restrict__ C, int n) {
    __shared__ unsigned short lk[256];
    __shared__ double la[256];
    __shared__ int wcount[8];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* arow = A + (size_t)i * n;
    for (int c0 = 0; c0 < n; c0 += 256 * SP_COLS) {
        double acc[SP_COLS];
#pragma unroll
        for (int c = 0; c < SP_COLS; ++c) acc[c] = 0.0;
        for (int k0 = 0; k0 < n; k0 += 256) {
            const int k = k0 + tid;
            const double a = (k < n) ? arow[k] : 0.0;
            const unsigned m = __ballot_sync(0xffffffffu, a != 0.0);
            if (lane == 0) wcount[warp] = __popc(m);
            __syncthreads();
            int base = 0, total = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { if (w < warp) base += wcount[w]; total += wcount[w]; }
            if (a != 0.0) {
                const int pos = base + __popc(m & ((1u << lane) - 1u));
                lk[pos] = (unsigned short)(k - k0); la[pos] = a;
            }
            __syncthreads();
            for (int e = 0; e < total; ++e) {
                const double av = la[e];
                const double* brow = B + (size_t)(k0 + lk[e]) * n + c0 + tid;
#pragma unroll
                for (int c = 0; c < SP_COLS; ++c)
                    if (c0 + tid + 256 * c < n) acc[c] = fma(av, brow[256 * c], acc[c]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int c = 0; c < SP_COLS; ++c)
            if (c0 + tid + 256 * c < n) C[(size_t)i * n + c0 + tid + 256 * c] = acc[c];
    }
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
    int* arg = nullptr;
    unsigned long long* flag = nullptr;      // [0] not close, [1] non-zeros of the pruned iterate
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
        CKM(sitb::pool_alloc((void**)&flag, 2 * sizeof(unsigned long long), st));
        const int kt = (n + GK - 1) / GK;
        CKM(sitb::pool_alloc((void**)&amap, (size_t)gg.y * kt, st));
        CKM(sitb::pool_alloc((void**)&bmap, (size_t)kt * gg.x, st));
        long long m1_nnz = -1;                     // non-zeros of m1 when known
        auto gemm = [&](const double* X, const double* Y, double* Z) {
            if (X == m1 && m1_nnz >= 0 && (double)m1_nnz < 0.06 * (double)cnt) {
                k_spgemm_rows<<<n, 256, 0, st>>>(X, Y, Z, n);
                return;
            }
            k_tilemap<<<dim3(kt, gg.y), 256, 0, st>>>(X, n, GB, GK, amap);
            k_tilemap<<<dim3(gg.x, kt), 256, 0, st>>>(Y, n, GK, GB, bmap);
            k_dgemm<<<gg, 256, 0, st>>>(X, Y, Z, n, amap, bmap);
        };
        m2 = dev_result;
        // m1 = graph / colsum (util/mcl.py:22-25)
        CKM(cudaMemcpyAsync(m1, dev_graph, bytes, cudaMemcpyDeviceToDevice, st));
        k_colsum<<<cb, 256, 0, st>>>(m1, n, colsum);
        k_coldiv<<<eb, 256, 0, st>>>(m1, n, colsum);
        {
            unsigned long long h_nnz = 0;
            CKM(cudaMemsetAsync(flag, 0, 2 * sizeof(unsigned long long), st));
            k_count_nonzero<<<eb, 256, 0, st>>>(m1, cnt, flag + 1);
            CKM(cudaMemcpyAsync(&h_nnz, flag + 1, sizeof(h_nnz), cudaMemcpyDeviceToHost, st));
            CKM(cudaStreamSynchronize(st));
            m1_nnz = (long long)h_nnz;
        }
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
            k_colsum<<<cb, 256, 0, st>>>(m2, n, colsum);                      // :35
            k_coldiv<<<eb, 256, 0, st>>>(m2, n, colsum);
            k_colargmax<<<cb, 256, 0, st>>>(m2, n, arg);                      // :39
            CKM(cudaMemsetAsync(flag, 0, 2 * sizeof(unsigned long long), st));
            k_prune_compare<<<eb, 256, 0, st>>>(m2, m1, n, arg, pruning_threshold, flag);   // :37-42
            unsigned long long h_flag[2] = {1ull, 0ull};
            CKM(cudaMemcpyAsync(h_flag, flag, sizeof(h_flag), cudaMemcpyDeviceToHost, st));
            CKM(cudaStreamSynchronize(st));
            CKM(cudaGetLastError());
            if (!h_flag[0]) { conv = 1; ++it; break; }
            m1_nnz = (long long)h_flag[1];
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
