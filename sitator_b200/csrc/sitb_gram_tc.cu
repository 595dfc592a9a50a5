// sitator_b200 -- K2: the landmark Gram  G = sum over rows of lv^T lv  (cluster/mcl.py:54) as a
// SYRK on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands by TMA).
//
// Input: the staged landmark vectors written by K1 in MODE_STAGE -- transposed (landmark-major) fp16 matrices
// Xhi, Xlo with value = hi + lo * 2^-12 (22 significant bits; the scale keeps lo in fp16's normal range), stored
// tile by tile: tile (rt, kt) = landmarks [128 rt, +128) x rows [64 kt, +64) is one contiguous 16 KB block at
// ((rt * ld/64) + kt) * 16 KB, already in the tensor core's K-major 128-byte-swizzled shared-memory layout
//     element (r, k) at byte  r*128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7)*2        (sitb_stage_offset)
// so one TMA bulk copy per tile brings an MMA operand in with fully sequential global reads, and both MMA
// operands are K-major tiles of the same matrix (K = landmark-vector index).
// Output: the upper triangle of G in float64 (+=), tile by tile.
//
//   G_ij = Hi_i Hi_j^T + 2^-12 (Hi_i Lo_j^T + Lo_i Hi_j^T)        (lo*lo ~ 2^-24 relative, dropped)
// with one TMEM accumulator for the hi.hi term and one for the cross terms.
//
// One CTA per (upper-triangular 128x128 tile, slice of K); the K split is chosen so the grid fills the SMs:
//   warp 0      TMA producer   (cp.async.bulk, 3-stage mbarrier ring, 4 tiles = 64 KB per stage)
//   warp 1      MMA issuer     (one elected lane; 3 x 4 tcgen05.mma per stage; tcgen05.commit frees the stage)
//   warps 2..9  epilogue       (tcgen05.ld 32x32b, FP32 -> FP64 accumulate in registers, one atomic add at the end)
// The tensor core's FP32 accumulation truncates, so its error grows linearly with the run length: the
// accumulator (double-buffered in TMEM, so the MMAs never wait for the drain) is drained into FP64
// registers every DRAIN_K rows, which bounds the error independently of the trajectory length.
#include "../../include/sitator_b200.h"
#include "sitb_common.cuh"
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sitb {
int set_error(int code, const char* fmt, ...);

constexpr int TM = 128;            // tile rows (UMMA M)
constexpr int TN = 128;            // tile cols (UMMA N)
constexpr int TK = 64;             // K per stage: 64 fp16 = 128 B = one swizzle atom row
constexpr int STAGES = 3;
constexpr int DRAIN_K = 512;       // rows between TMEM -> FP64 drains (the FP32 accumulator truncates: keep runs short)
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int TMEM_COLS = 4 * TN;   // double-buffered pair of accumulators (hi.hi and the scaled cross terms): all of TMEM
constexpr float LO_UNSCALE = 1.0f / 4096.0f;   // K1 stages lo * 2^12 so it stays in fp16's normal range
constexpr uint32_t TILE_BYTES = TM * TK * 2;            // 16 KB
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;        // Hi_i, Lo_i, Hi_j, Lo_j
constexpr long long SPIN_LIMIT = 400000000LL;           // watchdog: ~ seconds, then abort instead of hanging

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: returns false (and raises *abort) if the barrier never flips
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
    for (long long spin = 0; spin < SPIN_LIMIT; ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((spin & 0xFFFF) == 0xFFFF && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}

// one pre-swizzled 16 KB operand tile, contiguous in global memory -> shared memory, by the TMA engine's bulk copy
__device__ __forceinline__ void tma_load_tile(void* dst, const void* src, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(TILE_BYTES), "r"(smem_u32(bar)) : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset
    d |= (uint64_t)1 << 46;                               // descriptor version
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}

// kind::f16, A/B = fp16 (format 0), D = fp32 (format 1), both K-major, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct __align__(1024) GramSmem {
    unsigned char tiles[STAGES][4][TILE_BYTES];
    uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(THREADS, 1)
k_gram_syrk(const unsigned char* __restrict__ stage_hi, const unsigned char* __restrict__ stage_lo, long long n_ktiles,
            int n_row_tiles,
            int n_tiles, long long n_ksteps_total, long long ksteps_per_part, int L, double* __restrict__ gram_upper,
            int* __restrict__ abort_flag) {
    extern __shared__ unsigned char smem_dyn[];
    GramSmem& sm = *reinterpret_cast<GramSmem*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // this CTA's tile (ti <= tj), enumerated row by row over the upper triangle, and its slice of K.
    // CTAs with consecutive ids work on the same K slice, so the row blocks they share meet in L2.
    int t = blockIdx.x % n_tiles, ti = 0;
    const long long kpart = blockIdx.x / n_tiles;
    while (t >= n_row_tiles - ti) { t -= n_row_tiles - ti; ++ti; }
    const int tj = ti + t;
    const long long ks_begin = kpart * ksteps_per_part;
    const long long ks_end = ks_begin + ksteps_per_part < n_ksteps_total ? ks_begin + ksteps_per_part : n_ksteps_total;
    const long long n_ksteps = ks_end > ks_begin ? ks_end - ks_begin : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: two FP32 accumulators of 128 columns each
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = sm.tmem_base;
    constexpr long long steps_per_drain = DRAIN_K / TK;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (long long ks = 0; ks < n_ksteps; ++ks) {
                const int s = (int)(ks % STAGES);
                const uint32_t phase = (uint32_t)((ks / STAGES) & 1);
                if (!mbar_wait(&sm.empty[s], phase ^ 1, abort_flag)) break;
                mbar_expect_tx(&sm.full[s], STAGE_BYTES);
                const size_t oi = ((size_t)ti * n_ktiles + (size_t)(ks_begin + ks)) * TILE_BYTES;
                const size_t oj = ((size_t)tj * n_ktiles + (size_t)(ks_begin + ks)) * TILE_BYTES;
                tma_load_tile(sm.tiles[s][0], stage_hi + oi, &sm.full[s]);
                tma_load_tile(sm.tiles[s][1], stage_lo + oi, &sm.full[s]);
                tma_load_tile(sm.tiles[s][2], stage_hi + oj, &sm.full[s]);
                tma_load_tile(sm.tiles[s][3], stage_lo + oj, &sm.full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (long long ks = 0; ks < n_ksteps; ++ks) {
                const int s = (int)(ks % STAGES);
                const uint32_t phase = (uint32_t)((ks / STAGES) & 1);
                const long long in_drain = ks % steps_per_drain;
                const long long dr = ks / steps_per_drain;
                const uint32_t acc = tmem + (uint32_t)(dr & 1) * (2 * TN), acc_x = acc + TN;
                if (in_drain == 0) {     // this accumulator must have been drained by the epilogue
                    if (!mbar_wait(&sm.acc_empty[dr & 1], (uint32_t)(((dr >> 1) & 1) ^ 1), abort_flag)) break;
                    asm volatile("tcgen05.fence::after_thread_sync;");
                }
                if (!mbar_wait(&sm.full[s], phase, abort_flag)) break;
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint64_t d_hi_i = make_smem_desc(smem_u32(sm.tiles[s][0]));
                const uint64_t d_lo_i = make_smem_desc(smem_u32(sm.tiles[s][1]));
                const uint64_t d_hi_j = make_smem_desc(smem_u32(sm.tiles[s][2]));
                const uint64_t d_lo_j = make_smem_desc(smem_u32(sm.tiles[s][3]));
#pragma unroll
                for (int kk = 0; kk < TK / 16; ++kk) {
                    const uint64_t adv = (uint64_t)((kk * 32) >> 4);       // 16 fp16 = 32 B along K inside the swizzle atom
                    umma_f16(acc, d_hi_i + adv, d_hi_j + adv, (in_drain | kk) != 0);
                    umma_f16(acc_x, d_hi_i + adv, d_lo_j + adv, (in_drain | kk) != 0);
                    umma_f16(acc_x, d_lo_i + adv, d_hi_j + adv, 1u);
                }
                umma_commit(&sm.empty[s]);                                  // frees the smem stage when the MMAs retire
                if (in_drain == steps_per_drain - 1 || ks == n_ksteps - 1) umma_commit(&sm.acc_full[dr & 1]);
            }
        }
    } else {
        // ===== epilogue: TMEM -> FP64 registers (every DRAIN_K rows), then one atomic add into G =====
        const int ew = warp - 2;
        const int quad = warp & 3;                        // TMEM lane quadrant this warp may read
        const int chalf = ew >> 2;                        // which 64 of the tile's 128 columns
        const int row = quad * 32 + lane;                 // tile row
        const long long n_drains = (n_ksteps + steps_per_drain - 1) / steps_per_drain;
        double sum[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) sum[c] = 0.0;
        bool ok = true;
        for (long long dr = 0; dr < n_drains; ++dr) {
            if (!mbar_wait(&sm.acc_full[dr & 1], (uint32_t)((dr >> 1) & 1), abort_flag)) { ok = false; break; }
            asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32], x[32];
                const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)((dr & 1) * (2 * TN) + chalf * 64 + h * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]),
                      "=r"(x[8]), "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15]),
                      "=r"(x[16]), "=r"(x[17]), "=r"(x[18]), "=r"(x[19]), "=r"(x[20]), "=r"(x[21]), "=r"(x[22]), "=r"(x[23]),
                      "=r"(x[24]), "=r"(x[25]), "=r"(x[26]), "=r"(x[27]), "=r"(x[28]), "=r"(x[29]), "=r"(x[30]), "=r"(x[31])
                    : "r"(taddr + (uint32_t)TN));
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    sum[h * 32 + c] += (double)fmaf(__uint_as_float(x[c]), LO_UNSCALE, __uint_as_float(v[c]));
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            mbar_arrive(&sm.acc_empty[dr & 1]);
        }
        const int gi = ti * TM + row;
        if (ok && n_drains > 0 && gi < L) {
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                const int gj = tj * TN + chalf * 64 + c;
                if (gj < L && gj >= gi) atomicAdd(&gram_upper[(size_t)gi * L + gj], sum[c]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

}  // namespace sitb

using namespace sitb;

extern "C" int sitb_gram_syrk_tc(int device, const void* dev_stage_hi, const void* dev_stage_lo, int32_t n_landmarks,
                                 int32_t lpad, int64_t ld, int64_t k_rows, double* dev_gram_upper, void* cuda_stream) {
    if (!dev_stage_hi || !dev_stage_lo || !dev_gram_upper || n_landmarks <= 0 || lpad % TM != 0 || lpad < n_landmarks ||
        ld <= 0 || ld % TK != 0 || k_rows <= 0 || k_rows > ld)
        return set_error(SITB_E_INVALID, "sitb_gram_syrk_tc: bad argument (lpad %% 128 and ld %% 64 must be 0, k_rows <= ld)");
    k_rows = (k_rows + TK - 1) / TK * TK;        // the columns up to ld are zero-filled by contract
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    const int n_row_tiles = lpad / TM;
    const int n_tiles = n_row_tiles * (n_row_tiles + 1) / 2;
    const size_t smem = sizeof(GramSmem) + 1024;
    e = cudaFuncSetAttribute(k_gram_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int* abort_flag = nullptr;
    e = pool_alloc((void**)&abort_flag, sizeof(int), st);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "device allocation: %s", cudaGetErrorString(e));
    cudaMemsetAsync(abort_flag, 0, sizeof(int), st);
    // K split: fill whole waves of SMs (one CTA per SM), slices a multiple of the drain length and not too short
    int n_sms = 148;
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device);
    const long long n_ksteps = k_rows / TK, spd = DRAIN_K / TK;
    long long best_split = 1;
    double best_eff = 0.0;
    for (long long sp = 1; sp <= 64; ++sp) {
        long long per = ((n_ksteps + sp - 1) / sp + spd - 1) / spd * spd;
        if (sp > 1 && per < 4 * spd) break;
        const long long parts = (n_ksteps + per - 1) / per, ctas = parts * n_tiles;
        const long long waves = (ctas + n_sms - 1) / n_sms;
        const double eff = (double)n_ksteps * n_tiles / ((double)waves * n_sms * per);
        if (eff > best_eff * 1.02) { best_eff = eff; best_split = sp; }
    }
    const long long per_part = ((n_ksteps + best_split - 1) / best_split + spd - 1) / spd * spd;
    const long long n_parts = (n_ksteps + per_part - 1) / per_part;
    k_gram_syrk<<<(unsigned)(n_tiles * n_parts), THREADS, smem, st>>>((const unsigned char*)dev_stage_hi, (const unsigned char*)dev_stage_lo, ld / TK,
                                                                     n_row_tiles, n_tiles, n_ksteps, per_part,
                                                                     n_landmarks, dev_gram_upper, abort_flag);
    e = cudaGetLastError();
    int h_abort = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_abort, abort_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    pool_free(abort_flag, st);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "sitb_gram_syrk_tc: %s", cudaGetErrorString(e));
    if (h_abort) return set_error(SITB_E_CUDA, "sitb_gram_syrk_tc: pipeline watchdog fired (an mbarrier never completed)");
    return SITB_OK;
}
