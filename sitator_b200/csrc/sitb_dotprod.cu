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
// out[2] = rows consumed, out[3..6] = SM cycles in the candidate / dot-product / commit phases and the number
// of candidates (diagnostics).
//
// One CTA of DP_FIT_WARPS warps works on ONE row at a time (the rows are sequential by definition); what the
// warps share out is the work inside the row, so that the row's critical path is a few short dependent steps
// instead of one warp's long in-order instruction stream:
//   1. candidates: warp w takes the row's entries w, w+8, ...; its lanes walk that landmark's centre list and
//      flag the centres (first toucher appends the centre to the candidate array)
//   2. dot products: warp w takes candidates w, w+8, ...; lane e gathers S_c[k_e] (all of a warp's candidates in
//      flight before the first reduction), warp sum, one lane per candidate forms the cosine; best per warp
//   3. warp 0 merges the warp bests (highest cosine, then lowest centre index = np.argmax), decides and commits;
//      the other warps clear the flags and pull the next row's lists towards L1 meanwhile
#ifndef SITB_DP_FIT_WARPS
#define SITB_DP_FIT_WARPS 16
#endif
static constexpr int DP_FIT_WARPS = SITB_DP_FIT_WARPS;
#ifndef SITB_DP_FIT_UNROLL
#define SITB_DP_FIT_UNROLL ((SITB_DP_FIT_WARPS >= 16) ? 3 : 4)
#endif
static constexpr int DP_FIT_UNROLL = SITB_DP_FIT_UNROLL;   // candidates per warp and pass
static constexpr int DP_SVAL_CANDS = 64;              // candidates whose gathered S values are kept for the commit

__global__ void __launch_bounds__(32 * DP_FIT_WARPS) k_dotprod_fit(
    const unsigned long long* __restrict__ row_ptr, const uint16_t* __restrict__ pk, const double* __restrict__ pv,
    long long n_rows, int L, double thr, int max_c, int cap, double* S, long long* cnt, double* nrm2,
    uint16_t* lists, uint16_t* llen, long long* out) {
    __shared__ unsigned flag[DP_MAX_CENTERS];         // centre is a candidate of the current row
    __shared__ double nrm2s[DP_MAX_CENTERS];          // |S_c|^2 (written back to nrm2 at the end)
    __shared__ uint16_t cands[DP_MAX_CENTERS];        // candidate centres of the current row, in order of first touch
    __shared__ int ncand_s, sh_C, sh_status;
    __shared__ unsigned long long nx_ptr[2];
    __shared__ uint16_t nx_k[2][32];
    __shared__ double nx_v[2][32];
    __shared__ double wb_cos[DP_FIT_WARPS], wb_dot[DP_FIT_WARPS];
    __shared__ int wb_c[DP_FIT_WARPS], wb_ci[DP_FIT_WARPS];
    __shared__ double sval[DP_SVAL_CANDS][32];        // S_c[k_e] of the first 32 entries, per candidate (phase 2 -> commit)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < DP_MAX_CENTERS; i += blockDim.x) flag[i] = 0u;
    if (tid == 0) { ncand_s = 0; sh_C = 0; sh_status = 0; }
    __syncthreads();
    int C = 0;
    long long r = 0;
    long long t_cand = 0, t_dot = 0, t_commit = 0, n_cand = 0, t_arrive = 0;
    // the last warp fetches the next row (pointer + first 32 entries) while the current one is processed and leaves
    // it in shared memory (double buffered), so a row starts from shared memory instead of two dependent global loads
    int k_next = 0;
    double v_next = 0.0;
    unsigned long long ptr_n = 0ull;
    if (warp == DP_FIT_WARPS - 1 && n_rows > 0) {
        ptr_n = row_ptr[0];
        if (lane < (int)(ptr_n & 0xFF)) { k_next = pk[(ptr_n >> 8) + lane]; v_next = pv[(ptr_n >> 8) + lane]; }
        if (lane == 0) nx_ptr[0] = ptr_n;
        nx_k[0][lane] = (uint16_t)k_next; nx_v[0][lane] = v_next;
    }
    __syncthreads();
    for (; r < n_rows; ++r) {
        long long tk = clock64();
#define DP_TICK(acc) { const long long now_ = clock64(); acc += now_ - tk; tk = now_; }
        const int buf = (int)(r & 1);
        const unsigned long long ptr = nx_ptr[buf];
        const int nnz = (int)(ptr & 0xFF);
        const int nch = (nnz + 31) >> 5;
        const unsigned long long off = ptr >> 8;
        if (warp == DP_FIT_WARPS - 1 && r + 1 < n_rows) {         // start pulling the next row in (used at the end of this row)
            ptr_n = row_ptr[r + 1];
            if (lane < (int)(ptr_n & 0xFF)) { k_next = pk[(ptr_n >> 8) + lane]; v_next = pv[(ptr_n >> 8) + lane]; }
            if (r + 2 < n_rows && lane == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(row_ptr + r + 2));
        }
        // every warp holds the row: lane e <-> entries e, e + 32, ...
        int kk[DP_ROW_CHUNKS];
        double vv[DP_ROW_CHUNKS];
        kk[0] = nx_k[buf][lane]; vv[0] = nx_v[buf][lane];
        double sq = (lane < nnz) ? vv[0] * vv[0] : 0.0;
        if (lane >= nnz) { kk[0] = 0; vv[0] = 0.0; }
#pragma unroll
        for (int j = 1; j < DP_ROW_CHUNKS; ++j)
            if (j < nch) {                                       // rows longer than 32 entries: the rest from global memory
                const int e = 32 * j + lane;
                kk[j] = 0; vv[j] = 0.0;
                if (e < nnz) { kk[j] = pk[off + e]; vv[j] = pv[off + e]; }
                sq = fma(vv[j], vv[j], sq);
            }
        const double vn2 = warp_sum(sq);                         // the same value in every warp
        if (C == 0) {                                           // the first row is always its own cluster (:231-233)
            if (warp == 0) {
                FOR_ROW_CHUNKS(j)
                    if (32 * j + lane < nnz) {
                        S[kk[j]] = vv[j];
                        lists[(size_t)kk[j] * cap] = 0;
                        llen[kk[j]] = 1;
                    }
                if (lane == 0) { cnt[0] = 1; nrm2s[0] = vn2; sh_C = 1; }
            }
            if (warp == DP_FIT_WARPS - 1 && r + 1 < n_rows) {
                if (lane == 0) nx_ptr[buf ^ 1] = ptr_n;
                nx_k[buf ^ 1][lane] = (uint16_t)k_next; nx_v[buf ^ 1][lane] = v_next;
            }
            __syncthreads();
            C = 1;
            continue;
        }
        // np.argmax over similarities that are all NaN returns 0 and NaN < threshold is False: an all-zero row, or
        // any row while centre 0 is still the zero vector, joins cluster 0 (:243-248)
        const bool forced0 = (nnz == 0) || (nrm2s[0] == 0.0);
        int ncand = 0;
        if (!forced0) {
            // 1. candidates
            FOR_ROW_CHUNKS(j)
                for (int el = warp; el < 32; el += DP_FIT_WARPS) {
                    if (32 * j + el >= nnz) break;
                    const int k = __shfl_sync(0xffffffffu, kk[j], el);
                    const int n = llen[k];
                    const uint16_t* lst = lists + (size_t)k * cap;
                    for (int t = lane; t < n; t += 32) {
                        const unsigned c = lst[t];
                        if (atomicExch(&flag[c], 1u) == 0u) cands[atomicAdd(&ncand_s, 1)] = (uint16_t)c;
                    }
                }
            __syncthreads();
            if (tid == 0) DP_TICK(t_cand)
            ncand = ncand_s;
            // 2. dot products and cosines
            double bcos = -1.0, bdot = 0.0;                       // this lane's best (lanes 0 .. DP_FIT_UNROLL-1 form cosines)
            int bc = 0x7FFFFFFF, bci = 0x7FFFFFFF;
            for (int ci0 = warp; ci0 < ncand; ci0 += DP_FIT_WARPS * DP_FIT_UNROLL) {
                int cu[DP_FIT_UNROLL];
                double part[DP_FIT_UNROLL];
#pragma unroll
                for (int u = 0; u < DP_FIT_UNROLL; ++u) {
                    const int ci = ci0 + u * DP_FIT_WARPS;
                    cu[u] = (ci < ncand) ? (int)cands[ci] : -1;
                    part[u] = 0.0;
                }
                FOR_ROW_CHUNKS(j) {
                    const bool in = 32 * j + lane < nnz;
                    double sv[DP_FIT_UNROLL];
#pragma unroll
                    for (int u = 0; u < DP_FIT_UNROLL; ++u)
                        sv[u] = (in && cu[u] >= 0) ? S[(size_t)cu[u] * L + kk[j]] : 0.0;
#pragma unroll
                    for (int u = 0; u < DP_FIT_UNROLL; ++u) part[u] = fma(vv[j], sv[u], part[u]);
                    if (j == 0) {
#pragma unroll
                        for (int u = 0; u < DP_FIT_UNROLL; ++u)
                            if (cu[u] >= 0 && ci0 + u * DP_FIT_WARPS < DP_SVAL_CANDS) sval[ci0 + u * DP_FIT_WARPS][lane] = sv[u];
                    }
                }
                double mydot = 0.0;
                int myc = -1, myci = 0;
#pragma unroll
                for (int u = 0; u < DP_FIT_UNROLL; ++u) {
                    const double d = warp_sum(part[u]);
                    if (lane == u) { mydot = d; myc = cu[u]; myci = ci0 + u * DP_FIT_WARPS; }
                }
                if (myc >= 0) {
                    // :241-243 (dot / |centre|) / |row|, as one reciprocal square root: equal to ~1 ulp, and only
                    // the order of the cosines and their side of the threshold are used
                    const double cosang = mydot * rsqrt(nrm2s[myc] * vn2);
                    if (cosang > bcos || (cosang == bcos && myc < bc)) { bcos = cosang; bc = myc; bdot = mydot; bci = myci; }
                }
            }
            // best of the warp: highest similarity, then lowest centre index (np.argmax: first maximum)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double oc = __shfl_xor_sync(0xffffffffu, bcos, o);
                const int ob = __shfl_xor_sync(0xffffffffu, bc, o);
                const double od = __shfl_xor_sync(0xffffffffu, bdot, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bci, o);
                if (oc > bcos || (oc == bcos && ob < bc)) { bcos = oc; bc = ob; bdot = od; bci = oi; }
            }
            if (lane == 0) { wb_cos[warp] = bcos; wb_c[warp] = bc; wb_dot[warp] = bdot; wb_ci[warp] = bci; }
            __syncthreads();
            if (tid == 0) DP_TICK(t_dot)
        }
        // 3. decision and commit (warp 0); the other warps clear the flags / prefetch for the next row
        const long long t_phase3 = clock64();
        if (warp == 0) {
            int a = -1, a_ci = 0x7FFFFFFF;                       // a_ci: the winner's place in the candidate array
            double dot_a = 0.0;
            int status = 0;
            if (forced0) {
                a = 0;
                if (nnz > 0) {
                    double part = 0.0;
                    FOR_ROW_CHUNKS(j)
                        if (32 * j + lane < nnz) part = fma(vv[j], S[kk[j]], part);
                    dot_a = warp_sum(part);
                }
            } else {
                double bcos = (lane < DP_FIT_WARPS) ? wb_cos[lane] : -1.0;
                int bc = (lane < DP_FIT_WARPS) ? wb_c[lane] : 0x7FFFFFFF;
                double bdot = (lane < DP_FIT_WARPS) ? wb_dot[lane] : 0.0;
                int bci = (lane < DP_FIT_WARPS) ? wb_ci[lane] : 0x7FFFFFFF;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double oc = __shfl_xor_sync(0xffffffffu, bcos, o);
                    const int ob = __shfl_xor_sync(0xffffffffu, bc, o);
                    const double od = __shfl_xor_sync(0xffffffffu, bdot, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bci, o);
                    if (oc > bcos || (oc == bcos && ob < bc)) { bcos = oc; bc = ob; bdot = od; bci = oi; }
                }
                // similarities are >= 0 and centres the row does not touch have 0: np.argmax of all zeros is 0
                double best = 0.0;
                int besta = 0;
                if (bcos > 0.0) { best = bcos; besta = bc; dot_a = bdot; a_ci = bci; }
                if (!(best < thr)) a = besta;                       // :248 (cos < threshold -> new cluster)
            }
            if (a < 0) {
                // new cluster (:252-262)
                if (C >= max_c) {
                    status = 1;
                } else {
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
                    if (__any_sync(0xffffffffu, full)) status = 2;
                    if (lane == 0) { cnt[a] = 1; nrm2s[a] = vn2; }
                    ++C;
                }
            } else {
                // join cluster a: the running mean of :283-289 in sum form
                bool full = false;
                FOR_ROW_CHUNKS(j)
                    if (32 * j + lane < nnz) {
                        double* s = S + (size_t)a * L + kk[j];
                        // (the winner's components were gathered in phase 2: first 32 entries from shared memory)
                        const double old = (j == 0 && a_ci < DP_SVAL_CANDS) ? sval[a_ci][lane] : *s;
                        if (old == 0.0) {                               // the centre gains a landmark
                            const int n = llen[kk[j]];
                            if (n >= cap) { full = true; continue; }
                            lists[(size_t)kk[j] * cap + n] = (uint16_t)a;
                            llen[kk[j]] = (uint16_t)(n + 1);
                        }
                        *s = old + vv[j];
                    }
                if (__any_sync(0xffffffffu, full)) status = 2;
                if (lane == 0) {
                    atomicAdd((unsigned long long*)&cnt[a], 1ull);          // fire and forget: nothing waits for the count
                    if (nnz > 0) nrm2s[a] = nrm2s[a] + 2.0 * dot_a + vn2;
                }
            }
            if (lane == 0) { sh_C = C; sh_status = status; ncand_s = 0; }
        } else {
            for (int ci = tid - 32; ci < ncand; ci += blockDim.x - 32) flag[cands[ci]] = 0u;
            if (warp == DP_FIT_WARPS - 1 && r + 1 < n_rows) {
                if (lane == 0) nx_ptr[buf ^ 1] = ptr_n;
                nx_k[buf ^ 1][lane] = (uint16_t)k_next; nx_v[buf ^ 1][lane] = v_next;
                if (lane < (int)(ptr_n & 0xFF)) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(llen + k_next));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(lists + (size_t)k_next * cap));
                }
            }
        }
        if (lane == 0) t_arrive += clock64() - t_phase3;          // diagnostics: when this warp reaches the row's last barrier
        __syncthreads();
        if (tid == 0) { DP_TICK(t_commit) n_cand += ncand; }
        C = sh_C;
        if (sh_status != 0) break;
    }
    __syncthreads();
    for (int i = tid; i < C; i += blockDim.x) nrm2[i] = nrm2s[i];
    if (tid == 0) {
        out[0] = C; out[1] = sh_status; out[2] = r;
        out[3] = t_cand; out[4] = t_dot; out[5] = t_commit; out[6] = n_cand;
    }
    if (lane == 0) out[8 + warp] = t_arrive;
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
    k_dotprod_fit<<<1, 32 * DP_FIT_WARPS, 0, (cudaStream_t)stream>>>((const unsigned long long*)dev_row_ptr, dev_pool_k, dev_pool_v, n_rows,
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
