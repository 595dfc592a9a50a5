// sitator_b200 -- K2s: the landmark Gram  G = sum over rows of lv^T lv  (cluster/mcl.py:54) in FP64 from the
// cached compressed rows (sitb_fill.cu, sparse_ptr/k/v).
//
// A row has ~22 non-zero components: 253 pair products, and adding each straight into the L x L matrix costs
// 253 L2 atomics per row (1.4e9 per 10^5 LLZO frames; that was 9 of pass A's 25 ms).  But one mobile atom sees
// almost the same landmarks from frame to frame (union over 16 consecutive frames: ~34 landmarks), so:
//   one warp per (mobile atom, window of GW_T = 16 consecutive frames)
//     1. union of the window's landmarks as a bitmap in shared memory -> local slot of every landmark
//     2. scatter the window's rows into a dense GW_T x slots FP64 matrix X in shared memory
//     3. X^T X on the FP64 tensor cores (mma.sync m8n8k4.f64, 8x8 tiles of the upper triangle, K = the 16 frames)
//     4. add the non-zero entries of the tiles into G straight from the accumulator registers:
//        ~35 atomics per row instead of 253, and ~50 instructions per row instead of ~300
//   a window whose union exceeds GW_CAP slots (an atom in transit through many sites: 11 % of the LLZO windows) is
//   split into halves, recursively, each handled the same way; only a single row beyond GW_CAP entries adds directly.
// The sums are the same FP64 products in a different association; G is exact to rounding either way.
//
// Deterministic accumulation (EXACT = true, the default of the clustering plugin): FP64 atomics add in arrival
// order, so two runs differed in the last bits of G.  Here every addend (a window's tile entry, or a single pair
// product on the direct path) is split into THREE 64-bit integers -- units of 2^-30, 2^-62 and 2^-94 -- and added
// with integer atomics.  An FP64 addend >= 2^-42 is represented exactly (addends below that are rounded to 2^-94),
// and integer addition is associative: G is the exact sum of the addends, rounded once at the end, bit-identical
// from run to run and, with the windows aligned to GLOBAL frame numbers, for every sharding whose boundaries are
// multiples of GW_T frames (the integer words are all-reduced before they are converted).  The third word is
// touched only by addends with bits below 2^-62 (addends < 2^-10; a product of two components is rarely that small).
// Sums stay below 2^63: an entry is at most the number of rows (< 2^32, sitb_api.cu), i.e. < 2^62 units of 2^-30.
// Layout of the word matrix long long[2 (L + 1)][L]: plane 0: upper triangle = the 2^-30 words, lower triangle = the
// 2^-62 words of the mirrored entry, row L = the 2^-62 words of the diagonal; plane 1, same positions as the 2^-62
// words: the 2^-94 words.
#include "../../include/sitator_b200.h"
#include "sitb_common.cuh"

namespace sitb {

#ifndef SITB_GW_T
#define SITB_GW_T 16
#endif
#ifndef SITB_GW_WARPS
#define SITB_GW_WARPS 14
#endif
constexpr int GW_T = SITB_GW_T;  // frames per window = K of the local product (mma k-steps of 4)
#ifndef SITB_GW_CAP
#define SITB_GW_CAP 48
#endif
constexpr int GW_CAP = SITB_GW_CAP;        // landmark slots per window (6 tiles of 8)
constexpr int GW_STRIDE = GW_CAP + 4;      // row stride of X in doubles: = 4 mod 16, so the 4 x 8 fragment loads are conflict-free
constexpr int GW_WARPS = SITB_GW_WARPS;   // warps per CTA (two CTAs per SM; X is GW_T x GW_STRIDE doubles per warp)
constexpr int GW_NT = GW_CAP / 8;

// add x >= 0 to entry (r <= c): FP64 atomic, or the two integer words described above
template <bool EXACT>
__device__ __forceinline__ void gram_add(double* __restrict__ gram, int L, unsigned r, unsigned c, double x) {
    if (!EXACT) {
        atomicAdd(&gram[(size_t)r * L + c], x);
    } else {
        unsigned long long* w = (unsigned long long*)gram;
        const size_t plane = (size_t)(L + 1) * L;
        // x = m 2^(ex - 1075) as a 128-bit fixed-point number in units of 2^-94 (integer shifts: the float64
        // floor / convert sequence cost more than the atomics).  x < 2^33 (an entry is bounded by the row count).
        const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
        const int ex = (int)(bits >> 52);
        if (ex == 0) return;                                    // zero (or subnormal: below any grid)
        const unsigned long long m = (bits & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
        const int sh = ex - 981;                                // = (ex - 1075) + 94
        unsigned __int128 F;
        if (sh >= 0) F = (unsigned __int128)m << sh;            // exact
        else if (sh > -54) F = (unsigned __int128)((m + (1ull << (-sh - 1))) >> -sh);      // rounded to 2^-94
        else return;
        const unsigned long long hi = (unsigned long long)(F >> 64);
        const unsigned long long mid = ((unsigned long long)F) >> 32;
        const unsigned long long lo = ((unsigned long long)F) & 0xFFFFFFFFull;
        const size_t at = (r == c) ? ((size_t)L * L + r) : ((size_t)c * L + r);
        if (hi) atomicAdd(&w[(size_t)r * L + c], hi);
        if (mid) atomicAdd(&w[at], mid);
        if (lo) atomicAdd(&w[plane + at], lo);
    }
}

template <bool EXACT>
__device__ __forceinline__ void gram_row_direct(const uint16_t* __restrict__ pk, const double* __restrict__ pv,
                                                unsigned long long off, int n, int lane, int L,
                                                double* __restrict__ gram) {
    for (int a = 0; a < n; ++a) {
        const unsigned ka = pk[off + a];
        const double va = pv[off + a];
        for (int b = a + lane; b < n; b += 32) {
            const unsigned kb = pk[off + b];
            const unsigned lo = ka < kb ? ka : kb, hi = ka < kb ? kb : ka;
            gram_add<EXACT>(gram, L, lo, hi, va * pv[off + b]);
        }
    }
}

// D(8x8) += A(8x4, row) * B(4x8, col) in FP64 on the tensor cores
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

__host__ __device__ inline size_t gram_warp_bytes(int words) {
    return ((sizeof(double) * GW_T * GW_STRIDE + 4 * (size_t)words + 2 * (size_t)words + 2 * GW_CAP) + 15) & ~(size_t)15;
}

template <bool EXACT>
__global__ void __launch_bounds__(GW_WARPS * 32, 2)
k_gram_windows(const unsigned long long* __restrict__ row_ptr, const uint16_t* __restrict__ pk,
               const double* __restrict__ pv, long long n_frames, int M, int L, int words,
               double* __restrict__ gram, int lead) {
    // lead: frames before the first multiple of GW_T in GLOBAL frame numbers (windows are aligned to those, so that
    // a window holds the same frames however the trajectory is sharded); 0 = the shard starts on a boundary
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per warp: X [GW_T][GW_STRIDE] f64 | bitmap [words] u32 | slot offset of each word [words] u16 | ids [GW_CAP] u16
    unsigned char* base = smem_raw + (size_t)warp * gram_warp_bytes(words);
    double* X = (double*)base;
    unsigned* bits = (unsigned*)(base + sizeof(double) * GW_T * GW_STRIDE);
    uint16_t* wofs = (uint16_t*)(bits + words);
    uint16_t* ids = wofs + words;
    const int r4 = lane & 3, c8 = lane >> 2;

    const long long shift = lead ? (GW_T - lead) : 0;             // window w covers local frames [w GW_T - shift, ...)
    const long long n_windows = (n_frames + shift + GW_T - 1) / GW_T;
    const long long n_tasks = n_windows * M;
    for (long long t = (long long)blockIdx.x * GW_WARPS + warp; t < n_tasks; t += (long long)gridDim.x * GW_WARPS) {
        const int j = (int)(t % M);
        long long f0 = (t / M) * GW_T - shift;
        long long f1 = f0 + GW_T;
        if (f0 < 0) f0 = 0;
        if (f1 > n_frames) f1 = n_frames;
        const int nf = (int)(f1 - f0);
        // lane i holds row i of the window
        unsigned long long my_off = 0;
        int my_n = 0;
        if (lane < nf) {
            const unsigned long long e = row_ptr[(f0 + lane) * M + j];
            my_off = e >> 8; my_n = (int)(e & 0xFFull);
        }
        const bool long_rows = __any_sync(0xffffffffu, my_n > 32);
        // the first 32 entries of all rows are fetched together (longer rows are rare)
        unsigned kk[GW_T];
#pragma unroll
        for (int i = 0; i < GW_T; ++i) {
            const unsigned long long off = __shfl_sync(0xffffffffu, my_off, i);
            const int n = __shfl_sync(0xffffffffu, my_n, i);
            kk[i] = (lane < n) ? (unsigned)pk[off + lane] : 0xFFFFFFFFu;
        }
        // Segments of the window, depth first: the whole window if its landmarks fit GW_CAP slots, else its halves, their
        // halves, ... (an atom in transit between sites: 11 % of the 16-frame windows at the LLZO shape exceed 48 slots,
        // 1 % of the 4-frame ones).  Adding such a window's rows pair by pair instead cost 253 atomics per row and was a
        // third of the kernel's atomics.  The partition depends on the window's data alone.
        int lo = 0, len = GW_T;
        while (lo < nf) {
            const int hi = lo + len;
            // 1. union bitmap of the segment's rows
            for (int w = lane; w < words; w += 32) bits[w] = 0u;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < GW_T; ++i)
                if (i >= lo && i < hi && kk[i] != 0xFFFFFFFFu) atomicOr(&bits[kk[i] >> 5], 1u << (kk[i] & 31u));
            if (long_rows) {
                for (int i = lo; i < hi && i < nf; ++i) {
                    const unsigned long long off = __shfl_sync(0xffffffffu, my_off, i);
                    const int n = __shfl_sync(0xffffffffu, my_n, i);
                    for (int e = 32 + lane; e < n; e += 32) {
                        const unsigned k = pk[off + e];
                        atomicOr(&bits[k >> 5], 1u << (k & 31u));
                    }
                }
            }
            __syncwarp();
            // slot offsets: exclusive prefix of the words' popcounts
            int carry = 0;
            for (int w0 = 0; w0 < words; w0 += 32) {
                const int w = w0 + lane;
                const int c = (w < words) ? __popc(bits[w]) : 0;
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (w < words) wofs[w] = (uint16_t)(carry + incl - c);
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            const int u = carry;
            __syncwarp();
            if (u > GW_CAP && len > 1) { len >>= 1; continue; }            // descend into the left half
            if (u > GW_CAP) {
                // a single row with more than GW_CAP entries: pair by pair
                gram_row_direct<EXACT>(pk, pv, __shfl_sync(0xffffffffu, my_off, lo), __shfl_sync(0xffffffffu, my_n, lo), lane, L, gram);
            } else if (u > 0) {
                for (int w = lane; w < words; w += 32) {
                    unsigned b = bits[w];
                    int s = wofs[w];
                    while (b) { ids[s++] = (uint16_t)(w * 32 + __ffs(b) - 1); b &= b - 1u; }
                }
                // 2. X[row][slot] = value for the segment's rows; the rows of its k-steps are cleared first (a k-step is
                // four rows: a segment of one or two rows shares its k-step with rows of earlier segments)
                const int r_lo = lo & ~3, r_hi = (hi + 3) & ~3;
                for (int i = r_lo * GW_STRIDE + lane; i < r_hi * GW_STRIDE; i += 32) X[i] = 0.0;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < GW_T; ++i) {
                    const unsigned long long off = __shfl_sync(0xffffffffu, my_off, i);
                    if (i >= lo && i < hi && kk[i] != 0xFFFFFFFFu) {
                        const unsigned k = kk[i];
                        const int s = wofs[k >> 5] + __popc(bits[k >> 5] & ((1u << (k & 31u)) - 1u));
                        X[i * GW_STRIDE + s] = pv[off + lane];
                    }
                }
                if (long_rows) {
                    for (int i = lo; i < hi && i < nf; ++i) {
                        const unsigned long long off = __shfl_sync(0xffffffffu, my_off, i);
                        const int n = __shfl_sync(0xffffffffu, my_n, i);
                        for (int e = 32 + lane; e < n; e += 32) {
                            const unsigned k = pk[off + e];
                            const int s = wofs[k >> 5] + __popc(bits[k >> 5] & ((1u << (k & 31u)) - 1u));
                            X[i * GW_STRIDE + s] = pv[off + e];
                        }
                    }
                }
                __syncwarp();
                // 3. + 4. upper-triangular tiles of X^T X over the segment's k-steps, flushed from the accumulators
                const int nt = (u + 7) >> 3;
                const int ks_lo = r_lo >> 2, ks_hi = r_hi >> 2;
                for (int ti = 0; ti < nt; ++ti) {
                    double acc[GW_NT][2];
#pragma unroll
                    for (int d = 0; d < GW_NT; ++d) acc[d][0] = acc[d][1] = 0.0;
                    for (int ks = ks_lo; ks < ks_hi; ++ks) {
                        const double* xr = X + (ks * 4 + r4) * GW_STRIDE + c8;
                        const double a = xr[ti * 8];
#pragma unroll
                        for (int d = 0; d < GW_NT; ++d)
                            if (ti + d < nt) dmma(acc[d], a, xr[(ti + d) * 8]);
                    }
                    const int r = ti * 8 + c8;                    // accumulator row (slot)
                    if (r < u) {
#pragma unroll
                        for (int d = 0; d < GW_NT; ++d) {
                            if (ti + d >= nt) break;
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const int c = (ti + d) * 8 + r4 * 2 + q;
                                if (c < u && c >= r && acc[d][q] != 0.0) gram_add<EXACT>(gram, L, ids[r], ids[c], acc[d][q]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            lo = hi;
            len = lo & -lo;                                       // the largest aligned segment that starts here
        }
    }
}

// exact = 0: gram is double[L][L] (upper triangle, FP64 atomics).  exact = 1: gram is the word matrix long long[(L+1)][L]
// (see the top of this file), windows aligned to global frame numbers (frame0 = global index of the first cached frame).
cudaError_t launch_gram_sparse(const unsigned long long* row_ptr, const uint16_t* pk, const double* pv, long long n_frames,
                               int M, int L, double* gram, int n_sms, cudaStream_t st, int exact, long long frame0) {
    if (n_frames <= 0) return cudaSuccess;
    const int words = (L + 31) / 32;
    const size_t smem = gram_warp_bytes(words) * GW_WARPS;
    if (smem > 110 * 1024) return cudaErrorInvalidConfiguration;      // L > ~9000: the caller keeps the in-kernel Gram
    const int lead = exact ? (int)((GW_T - (frame0 % GW_T)) % GW_T) : 0;
    cudaError_t e;
    if (exact) {
        e = cudaFuncSetAttribute(k_gram_windows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_gram_windows<true><<<n_sms * 2, GW_WARPS * 32, smem, st>>>(row_ptr, pk, pv, n_frames, M, L, words, gram, lead);
    } else {
        e = cudaFuncSetAttribute(k_gram_windows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_gram_windows<false><<<n_sms * 2, GW_WARPS * 32, smem, st>>>(row_ptr, pk, pv, n_frames, M, L, words, gram, 0);
    }
    return cudaGetLastError();
}

// word matrix -> double[L][L] upper triangle (lower triangle zero): value = hi 2^-24 + lo 2^-56, one rounding
__global__ void k_gram_words_finish(const long long* __restrict__ w, int L, double* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * L) return;
    const int r = (int)(idx / L), c = (int)(idx % L);
    double v = 0.0;
    if (c >= r) {
        const size_t plane = (size_t)(L + 1) * L;
        const size_t at = (r == c) ? ((size_t)L * L + r) : ((size_t)c * L + r);
        const unsigned long long hi = (unsigned long long)w[(size_t)r * L + c];
        const unsigned long long mid = (unsigned long long)w[at];
        const unsigned long long lo = (unsigned long long)w[plane + at];
        // the exact sum is (hi 2^64 + mid 2^32 + lo) 2^-94: the two low words as a 128-bit integer, its carry into hi,
        // then two conversions and one addition (a fixed function of the words; error below one ulp of the result)
        const unsigned __int128 low = ((unsigned __int128)mid << 32) + (unsigned __int128)lo;     // < 2^97
        const unsigned long long carry = (unsigned long long)(low >> 64);                          // units of 2^-30
        const unsigned long long rest = (unsigned long long)low;                                   // units of 2^-94
        v = ldexp((double)rest, -94) + ldexp((double)(hi + carry), -30);
    }
    out[idx] = v;
}

cudaError_t launch_gram_words_finish(const long long* words, int L, double* out, cudaStream_t st) {
    const size_t cnt = (size_t)L * L;
    k_gram_words_finish<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(words, L, out);
    return cudaGetLastError();
}

}  // namespace sitb
