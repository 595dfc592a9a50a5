// sitator_b200 -- the "dotprod" clustering plugin's passes over the cached compressed landmark vectors.
//
//   util/DotProdClassifier.pyx:199-315  fit_centers: online "leader" clustering of the rows IN ORDER: a row
//                                       joins the centre of highest cosine similarity if that reaches the
//                                       threshold (the centre becomes the running mean of its members),
//                                       otherwise it founds a new centre.
//   util/DotProdClassifier.pyx:129-197  predict (predict_normed=True): |normalised centre . x| / |x|, first
//                                       arg-max, threshold.
//
// fit: the loop is sequential by definition (row i sees the centres as rows < i left them), so one warp walks
// the rows; what is parallel is the work inside a row.  State is kept in SUM form: S_c = sum of the member
// rows, n_c = their number.  The cosine does not depend on the scale of the centre, so every decision is the
// one the reference takes with centre = S_c / n_c, and the centre's norm follows from
// |S + v|^2 = |S|^2 + 2 S.v + |v|^2 with S.v already known.  Rows are ~23 of 1500 non-zero and a centre only
// matters if it shares a landmark with the row: per landmark a short list of the centres that are non-zero
// there gives the candidates (a bitmap in shared memory, walked in ascending centre order = np.argmax's
// first maximum), and each candidate's dot product is a gather of the row's ~23 components from the dense
// S matrix (L1 / L2 resident) reduced over the warp.
#include "../../include/sitator_b200.h"
#include "sitb_common.cuh"

namespace sitb {

static constexpr int DP_MAX_WORDS = 64;               // at most 2048 centres (shared-memory tables per centre)
static constexpr int DP_MAX_CENTERS = DP_MAX_WORDS * 32;
static constexpr int DP_ROW_CHUNKS = 8;               // a row has at most 255 entries (8-bit count): 8 per lane

// the row's entries sit in registers, 32 per chunk; chunks past the row's length are skipped by a uniform branch
#define FOR_ROW_CHUNKS(j) _Pragma("unroll") for (int j = 0; j < DP_ROW_CHUNKS; ++j) if (j < nch)

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// out[0] = number of centres, out[1] = status (0 ok, 1 more than max_c centres, 2 a landmark list is full),
// out[2] = rows consumed
__global__ void __launch_bounds__(32) k_dotprod_fit(const unsigned long long* __restrict__ row_ptr,
                                                     const uint16_t* __restrict__ pk, const double* __restrict__ pv,
                                                     long long n_rows, int L, double thr, int max_c, int cap,
                                                     double* S, long long* cnt, double* nrm2, uint16_t* lists,
                                                     uint16_t* llen, long long* out) {
    __shared__ unsigned emask[DP_MAX_CENTERS];        // per centre: which entries of the current row it is non-zero at
    __shared__ double nrm2s[DP_MAX_CENTERS];          // |S_c|^2 (written back to nrm2 at the end)
    __shared__ uint16_t cands[DP_MAX_CENTERS];        // candidate centres of the current row, ascending
    __shared__ unsigned cmask[DP_MAX_CENTERS];        // and their entry masks
    __shared__ uint16_t rk[32 * DP_ROW_CHUNKS];       // the current row, readable by every lane
    __shared__ double rv[32 * DP_ROW_CHUNKS];
    const int lane = threadIdx.x;
    for (int i = lane; i < DP_MAX_CENTERS; i += 32) emask[i] = 0u;
    int C = 0;
    int status = 0;
    long long r = 0;
    long long t_load = 0, t_flag = 0, t_enum = 0, t_dot = 0, t_commit = 0, n_cand = 0;
    __syncwarp();
    unsigned long long ptr_next = n_rows > 0 ? row_ptr[0] : 0ull;
    int k_next = 0;
    double v_next = 0.0;
    if (n_rows > 0 && lane < (int)(ptr_next & 0xFF)) { k_next = pk[(ptr_next >> 8) + lane]; v_next = pv[(ptr_next >> 8) + lane]; }
    for (; r < n_rows; ++r) {
        long long tk = clock64();
#define DP_TICK(acc) { const long long now_ = clock64(); acc += now_ - tk; tk = now_; }
        const unsigned long long ptr = ptr_next;
        const int nnz = (int)(ptr & 0xFF);
        const int nch = (nnz + 31) >> 5;
        const unsigned long long off = ptr >> 8;
        // one warp, in-order issue: a load only overlaps with what is issued before its first use, so everything
        // the next rows need is requested early -- the pointer two rows ahead, the entries one row ahead
        int kk[DP_ROW_CHUNKS];
        double vv[DP_ROW_CHUNKS];
        kk[0] = k_next; vv[0] = v_next;                         // first 32 entries: loaded while the previous row ran
        double sq = (lane < nnz) ? vv[0] * vv[0] : 0.0;
        if (r + 1 < n_rows) {
            const unsigned long long np = row_ptr[r + 1];        // (its line was touched two rows ago)
            ptr_next = np;
            const unsigned long long no = np >> 8;
            if (lane < (int)(np & 0xFF)) { k_next = pk[no + lane]; v_next = pv[no + lane]; }   // used one row later
            if (r + 2 < n_rows && lane == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(row_ptr + r + 2));
        }
#pragma unroll
        for (int j = 1; j < DP_ROW_CHUNKS; ++j)
            if (j < nch) {
                const int e = 32 * j + lane;
                kk[j] = 0; vv[j] = 0.0;
                if (e < nnz) { kk[j] = pk[off + e]; vv[j] = pv[off + e]; }
                sq = fma(vv[j], vv[j], sq);
            }
        const double vn2 = warp_sum(sq);
        DP_TICK(t_load)
        if (C == 0) {                                           // the first row is always its own cluster (:231-233)
            FOR_ROW_CHUNKS(j)
                if (32 * j + lane < nnz) {
                    S[kk[j]] = vv[j];
                    lists[(size_t)kk[j] * cap] = 0;
                    llen[kk[j]] = 1;
                }
            if (lane == 0) { cnt[0] = 1; nrm2s[0] = vn2; }
            C = 1;
            __syncwarp();
            continue;
        }
        // np.argmax over similarities that are all NaN returns 0 and NaN < threshold is False: an all-zero row, or
        // any row while centre 0 is still the zero vector, joins cluster 0 (:243-248)
        int a = -1;
        double dot_a = 0.0;
        if (nnz == 0 || nrm2s[0] == 0.0) {
            a = 0;
            if (nnz > 0) {
                double part = 0.0;
                FOR_ROW_CHUNKS(j)
                    if (32 * j + lane < nnz) part = fma(vv[j], S[kk[j]], part);
                dot_a = warp_sum(part);
            }
        } else {
            // candidates: the centres listed at the row's landmarks.  emask[c] collects, per centre, the row entries
            // where it is non-zero (rows of <= 32 entries; longer rows only flag the centre)
            const bool small = nnz <= 32;
            FOR_ROW_CHUNKS(j)
                if (32 * j + lane < nnz) {
                    rk[32 * j + lane] = (uint16_t)kk[j];
                    rv[32 * j + lane] = vv[j];
                    const unsigned bit = small ? (1u << lane) : 1u;
                    const int n = llen[kk[j]];
                    const uint4* lst = (const uint4*)(lists + (size_t)kk[j] * cap);    // cap is a multiple of 32
                    for (int t0 = 0; t0 < n; t0 += 32) {
                        const uint4 q0 = lst[(t0 >> 3)], q1 = lst[(t0 >> 3) + 1], q2 = lst[(t0 >> 3) + 2], q3 = lst[(t0 >> 3) + 3];
                        const unsigned w[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w,
                                                q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (t0 + 2 * i < n) atomicOr(&emask[w[i] & 0xFFFFu], bit);
                            if (t0 + 2 * i + 1 < n) atomicOr(&emask[w[i] >> 16], bit);
                        }
                    }
                }
            __syncwarp();
            DP_TICK(t_flag)
            // -> candidate array, ascending (32 centres per round)
            int ncand = 0;
            for (int c0 = 0; c0 < C; c0 += 32) {
                const unsigned m = (c0 + lane < C) ? emask[c0 + lane] : 0u;
                const unsigned any = __ballot_sync(0xffffffffu, m != 0u);
                if (m) {
                    const int pos = ncand + __popc(any & lanemask_lt());
                    cands[pos] = (uint16_t)(c0 + lane);
                    cmask[pos] = m;
                    emask[c0 + lane] = 0u;
                }
                ncand += __popc(any);
            }
            __syncwarp();
            DP_TICK(t_enum)
            n_cand += ncand;
            // one candidate per lane: its dot product with the row is a chain of independent gathers from S
            // (all in flight at once), summed in ascending entry order
            const double vnorm = sqrt(vn2);
            double best = 0.0;                                    // similarities are >= 0; untouched centres have 0
            int besta = 0;                                        // np.argmax of all zeros
            for (int c0 = 0; c0 < ncand; c0 += 32) {
                const bool has = c0 + lane < ncand;
                const int c = has ? (int)cands[c0 + lane] : 0;
                const double* __restrict__ sc = S + (size_t)c * L;
                double dot = 0.0;
                if (has && small) {
                    unsigned m = cmask[c0 + lane];                          // only the entries where S_c is non-zero
                    while (m) {                                             // 8 gathers in flight, then the sum in order
                        double sv[8];
                        int ee[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            ee[i] = -1; sv[i] = 0.0;
                            if (m) { ee[i] = __ffs(m) - 1; m &= m - 1u; sv[i] = sc[rk[ee[i]]]; }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (ee[i] >= 0) dot = fma(rv[ee[i]], sv[i], dot);
                    }
                } else if (has) {
                    for (int e0 = 0; e0 < nnz; e0 += 16) {
                        double sv[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) sv[i] = (e0 + i < nnz) ? sc[rk[e0 + i]] : 0.0;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (e0 + i < nnz) dot = fma(rv[e0 + i], sv[i], dot);
                    }
                }
                double cosang = has ? (dot / sqrt(nrm2s[c])) / vnorm : -1.0;    // :241-243
                int cbest = c;
                double dbest = dot;
                // arg-max over the lanes: highest similarity, then lowest centre index (np.argmax: first maximum)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double oc = __shfl_xor_sync(0xffffffffu, cosang, o);
                    const int ob = __shfl_xor_sync(0xffffffffu, cbest, o);
                    const double od = __shfl_xor_sync(0xffffffffu, dbest, o);
                    if (oc > cosang || (oc == cosang && ob < cbest)) { cosang = oc; cbest = ob; dbest = od; }
                }
                if (cosang > best) { best = cosang; besta = cbest; dot_a = dbest; }   // chunks ascend: ties keep the earlier
            }
            __syncwarp();
            DP_TICK(t_dot)
            if (!(best < thr)) a = besta;                           // :248 (cos < threshold -> new cluster)
            if (a == besta && best == 0.0) dot_a = 0.0;
        }
        if (a < 0) {
            // new cluster (:252-262)
            if (C >= max_c) { status = 1; break; }
            a = C;
            bool full = false;
            FOR_ROW_CHUNKS(j)
                if (32 * j + lane < nnz) {
                    const int n = llen[kk[j]];
                    if (n >= cap) { full = true; continue; }
                    S[(size_t)a * L + kk[j]] = vv[j];
                    lists[(size_t)kk[j] * cap + n] = (uint16_t)a;
                    llen[kk[j]] = (uint16_t)(n + 1);
                }
            if (__any_sync(0xffffffffu, full)) { status = 2; break; }
            if (lane == 0) { cnt[a] = 1; nrm2s[a] = vn2; }
            ++C;
        } else {
            // join cluster a: the running mean of :283-289 in sum form
            bool full = false;
            FOR_ROW_CHUNKS(j)
                if (32 * j + lane < nnz) {
                    double* s = S + (size_t)a * L + kk[j];
                    const double old = *s;
                    if (old == 0.0) {                               // the centre gains a landmark
                        const int n = llen[kk[j]];
                        if (n >= cap) { full = true; continue; }
                        lists[(size_t)kk[j] * cap + n] = (uint16_t)a;
                        llen[kk[j]] = (uint16_t)(n + 1);
                    }
                    *s = old + vv[j];
                }
            if (__any_sync(0xffffffffu, full)) { status = 2; break; }
            if (lane == 0) {
                atomicAdd((unsigned long long*)&cnt[a], 1ull);          // fire and forget: nothing waits for the count
                if (nnz > 0) nrm2s[a] = nrm2s[a] + 2.0 * dot_a + vn2;
            }
        }
        __syncwarp();
        DP_TICK(t_commit)
    }
    __syncwarp();
    for (int i = lane; i < C; i += 32) nrm2[i] = nrm2s[i];
    if (lane == 0) {
        out[0] = C; out[1] = status; out[2] = r;
        // diagnostics: SM cycles per phase (row load, candidate flags, enumeration, dot products, commit), candidates
        out[3] = t_load; out[4] = t_flag; out[5] = t_enum; out[6] = t_dot; out[7] = t_commit; out[8] = n_cand;
    }
}

// predict: one warp per row.  Centres are given dense and already normalised (C x L), plus per landmark the
// list of centres that are non-zero there (CSR).  Candidates are walked in ascending centre order, so the first
// maximum wins as in np.argmax; a row no centre touches gets similarity 0 everywhere -> index 0 -> below any
// positive threshold -> unassigned.
__global__ void __launch_bounds__(256) k_dotprod_predict(const unsigned long long* __restrict__ row_ptr,
                                                          const uint16_t* __restrict__ pk, const double* __restrict__ pv,
                                                          long long n_rows, int L, int C, const double* __restrict__ centres,
                                                          const unsigned* __restrict__ cptr, const uint16_t* __restrict__ cc,
                                                          double thr, long long* __restrict__ labels,
                                                          double* __restrict__ confs, unsigned long long* __restrict__ counts) {
    __shared__ unsigned bitmap_all[8][DP_MAX_WORDS];
    __shared__ unsigned hist[DP_MAX_CENTERS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned* bitmap = bitmap_all[warp];
    for (int i = lane; i < DP_MAX_WORDS; i += 32) bitmap[i] = 0u;
    for (int i = threadIdx.x; i < C; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const int nwords = (C + 31) >> 5;
    for (long long r = (long long)blockIdx.x * nwarps + warp; r < n_rows; r += (long long)gridDim.x * nwarps) {
        const unsigned long long ptr = row_ptr[r];
        const int nnz = (int)(ptr & 0xFF);
        const int nch = (nnz + 31) >> 5;
        const unsigned long long off = ptr >> 8;
        int kk[DP_ROW_CHUNKS];
        double vv[DP_ROW_CHUNKS];
        double sq = 0.0;
        FOR_ROW_CHUNKS(j) {
            const int e = 32 * j + lane;
            kk[j] = 0; vv[j] = 0.0;
            if (e < nnz) { kk[j] = pk[off + e]; vv[j] = pv[off + e]; }
            sq = fma(vv[j], vv[j], sq);
        }
        const double vnorm = sqrt(warp_sum(sq));
        long long label = -1;
        double conf = 0.0;
        if (nnz > 0) {                                              // all-zero rows: label -1 (:168-172)
            FOR_ROW_CHUNKS(j)
                if (32 * j + lane < nnz) {
                    const unsigned b = cptr[kk[j]], e = cptr[kk[j] + 1];
                    for (unsigned t = b; t < e; ++t) {
                        const unsigned c = cc[t];
                        atomicOr(&bitmap[c >> 5], 1u << (c & 31));
                    }
                }
            __syncwarp();
            double best = 0.0;
            int besta = 0;
            for (int w0 = 0; w0 < nwords; w0 += 32) {
                const unsigned mine = (w0 + lane < nwords) ? bitmap[w0 + lane] : 0u;
                if (w0 + lane < nwords) bitmap[w0 + lane] = 0u;
                unsigned any = __ballot_sync(0xffffffffu, mine != 0u);
                while (any) {
                    const int wl = __ffs(any) - 1;
                    any &= any - 1u;
                    unsigned bits = __shfl_sync(0xffffffffu, mine, wl);
                    while (bits) {
                        const int c = ((w0 + wl) << 5) + __ffs(bits) - 1;
                        bits &= bits - 1u;
                        double part = 0.0;
                        FOR_ROW_CHUNKS(j)
                            if (32 * j + lane < nnz) part = fma(vv[j], __ldg(centres + (size_t)c * L + kk[j]), part);
                        const double d = fabs(warp_sum(part) / vnorm);     // :176-179
                        if (d > best) { best = d; besta = c; }
                    }
                }
            }
            __syncwarp();
            if (!(best < thr)) { label = besta; conf = best; }          // :184-186
        }
        if (lane == 0) {
            if (labels) labels[r] = label;
            if (confs) confs[r] = conf;
            if (label >= 0 && counts) atomicAdd(&hist[label], 1u);
        }
    }
    __syncthreads();
    if (counts)
        for (int i = threadIdx.x; i < C; i += blockDim.x)
            if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
}

}  // namespace sitb

using namespace sitb;

namespace sitb { int set_error(int code, const char* fmt, ...); }

extern "C" int sitb_dotprod_limits(int32_t* max_centers, int32_t* max_row_entries) {
    if (max_centers) *max_centers = DP_MAX_CENTERS;
    if (max_row_entries) *max_row_entries = 32 * DP_ROW_CHUNKS;
    return SITB_OK;
}

extern "C" int sitb_dotprod_fit(int device, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k,
                                const double* dev_pool_v, int64_t n_rows, int32_t n_landmarks, double threshold,
                                int32_t max_centers, int32_t list_cap, double* dev_sums, int64_t* dev_counts,
                                double* dev_norm2, uint16_t* dev_lists, uint16_t* dev_list_len, int64_t* dev_out3,
                                void* stream) {
    if (!dev_row_ptr || !dev_pool_k || !dev_pool_v || !dev_sums || !dev_counts || !dev_norm2 || !dev_lists ||
        !dev_list_len || !dev_out3 || n_rows < 0 || n_landmarks <= 0 || list_cap <= 0)
        return set_error(SITB_E_INVALID, "sitb_dotprod_fit: bad argument");
    if (max_centers <= 0 || max_centers > DP_MAX_CENTERS)
        return set_error(SITB_E_LIMIT, "sitb_dotprod_fit: max_centers %d outside [1, %d]", max_centers, DP_MAX_CENTERS);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    k_dotprod_fit<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)dev_row_ptr, dev_pool_k, dev_pool_v, n_rows,
                                                     n_landmarks, threshold, max_centers, list_cap, dev_sums,
                                                     (long long*)dev_counts, dev_norm2, dev_lists, dev_list_len,
                                                     (long long*)dev_out3);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "k_dotprod_fit: %s", cudaGetErrorString(e));
    return SITB_OK;
}

extern "C" int sitb_dotprod_predict(int device, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k,
                                    const double* dev_pool_v, int64_t n_rows, int32_t n_landmarks, int32_t n_centers,
                                    const double* dev_normed_centers, const uint32_t* dev_list_ptr,
                                    const uint16_t* dev_list_centers, double threshold, int64_t* dev_labels,
                                    double* dev_confs, uint64_t* dev_counts, void* stream) {
    if (!dev_row_ptr || !dev_pool_k || !dev_pool_v || !dev_normed_centers || !dev_list_ptr || !dev_list_centers ||
        n_rows < 0 || n_landmarks <= 0)
        return set_error(SITB_E_INVALID, "sitb_dotprod_predict: bad argument");
    if (n_centers <= 0 || n_centers > DP_MAX_CENTERS)
        return set_error(SITB_E_LIMIT, "sitb_dotprod_predict: %d centres outside [1, %d]", n_centers, DP_MAX_CENTERS);
    if (n_rows == 0) return SITB_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    int n_sms = 0;
    e = cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    long long grid = (long long)n_sms * 8;
    if (grid > (n_rows + 7) / 8) grid = (n_rows + 7) / 8;
    k_dotprod_predict<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long*)dev_row_ptr, dev_pool_k, dev_pool_v, n_rows, n_landmarks, n_centers, dev_normed_centers,
        dev_list_ptr, dev_list_centers, threshold, (long long*)dev_labels, dev_confs, (unsigned long long*)dev_counts);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "k_dotprod_predict: %s", cudaGetErrorString(e));
    return SITB_OK;
}
