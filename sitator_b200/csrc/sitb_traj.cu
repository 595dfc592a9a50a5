// sitator_b200 -- K5/K6: integer passes over the (frames x mobile atoms) assignment stream and
// the periodic site-centre average.
//
//   SiteTrajectory.py:205-232     check_multiple_occupancy                      (k_occupancy)
//   SiteTrajectory.py:307-373     jumps() / _jumped_generator                   (k_chunk_last, k_jump_from, compaction)
//   dynamics/JumpAnalysis.py:27-135  run(): n_ij, jump lag, residence sums     (k_ja_*)
//   LandmarkAnalysis.py:276-287 + PBCCalculator.pyx:106-139  site centres       (k_site_*)
//
// All of these are HBM-bound streams over int64 labels (8 B per mobile atom and frame).  The
// frame axis is cut into chunks; anything sequential in time ("last known site", "frames since
// the last jump") is a carry resolved over the chunk summaries by one thread per atom.
#include "../../include/sitator_b200.h"
#include "sitb_common.cuh"
#include <cuda_runtime.h>
#include <vector>

namespace sitb {

int set_error(int code, const char* fmt, ...);

// ------------------------------------------------------------------------------------------------
// occupancy check
// ------------------------------------------------------------------------------------------------
// out[0] += #(frame, site) with more than one atom ; out[1] += #assigned atoms ; out[2] += #(frame, site)
// occupied ; first_bad = min over offending frames of (frame << 32 | lowest offending site)
__global__ void k_occupancy(const long long* __restrict__ traj, long long F, int M, long long frame0,
                            int max_per_site, unsigned long long* __restrict__ out,
                            unsigned long long* __restrict__ first_bad) {
    extern __shared__ int row[];               // site ids fit 32 bits (sitb_set_centers: < 32768 sites)
    unsigned long long n_more = 0, n_assigned = 0, n_distinct = 0;
    for (long long f = blockIdx.x; f < F; f += gridDim.x) {
        for (int a = threadIdx.x; a < M; a += blockDim.x) {
            const long long v = traj[f * M + a];
            row[a] = v < 0 ? -1 : (int)v;
        }
        __syncthreads();
        for (int a = threadIdx.x; a < M; a += blockDim.x) {
            const int s = row[a];
            if (s < 0) continue;
            int cnt = 0;
            bool first = true;
            for (int b = 0; b < M; ++b) {
                if (row[b] == s) { ++cnt; if (b < a) first = false; }
            }
            ++n_assigned;
            if (first) {
                ++n_distinct;
                if (cnt > 1) ++n_more;
                if (cnt > max_per_site)
                    atomicMin(first_bad, ((unsigned long long)(frame0 + f) << 32) | (unsigned long long)(unsigned)s);
            }
        }
        __syncthreads();
    }
    // block reduce
    __shared__ unsigned long long red[3];
    if (threadIdx.x == 0) red[0] = red[1] = red[2] = 0;
    __syncthreads();
    atomicAdd(&red[0], n_more); atomicAdd(&red[1], n_assigned); atomicAdd(&red[2], n_distinct);
    __syncthreads();
    if (threadIdx.x == 0) { atomicAdd(&out[0], red[0]); atomicAdd(&out[1], red[1]); atomicAdd(&out[2], red[2]); }
}

// ------------------------------------------------------------------------------------------------
// jump list
// ------------------------------------------------------------------------------------------------
constexpr int CHUNK = 256;   // frames per chunk

// last known (non -1) site of every atom inside each chunk, -1 if none
__global__ void k_chunk_last(const long long* __restrict__ traj, long long F, int M, int* __restrict__ chunk_last) {
    const long long c = blockIdx.x;
    const long long f0 = c * CHUNK, f1 = (f0 + CHUNK < F) ? (f0 + CHUNK) : F;
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
        int last = -1;
        for (long long f = f0; f < f1; ++f) {
            const long long s = traj[f * M + a];
            if (s >= 0) last = (int)s;
        }
        chunk_last[c * M + a] = last;
    }
}

// exclusive "last known" carry over chunks: carry_in[c][a] = last known site before chunk c
__global__ void k_chunk_carry(const int* __restrict__ chunk_last, long long n_chunks, int M,
                              const long long* __restrict__ shard_carry, int* __restrict__ carry_in) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= M) return;
    int cur = shard_carry ? (int)shard_carry[a] : -1;
    for (long long c = 0; c < n_chunks; ++c) {
        carry_in[c * M + a] = cur;
        const int l = chunk_last[c * M + a];
        if (l >= 0) cur = l;
    }
}

// from[f][a] = site the atom jumped from at frame f, or -2 if it did not jump there
// (SiteTrajectory.py:353-373).  first_frame_is_start: global frame 0 is in this shard (no jump there).
__global__ void k_jump_from(const long long* __restrict__ traj, long long F, int M, const int* __restrict__ carry_in,
                            int unknown_as_jump, int first_frame_is_start, const long long* __restrict__ prev_row,
                            int* __restrict__ from, unsigned long long* __restrict__ total) {
    const long long c = blockIdx.x;
    const long long f0 = c * CHUNK, f1 = (f0 + CHUNK < F) ? (f0 + CHUNK) : F;
    unsigned long long cnt = 0;
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
        // unknown_as_jump: last_known is simply the previous frame's entry
        long long last = unknown_as_jump
            ? (f0 > 0 ? traj[(f0 - 1) * M + a] : (prev_row ? prev_row[a] : -1))
            : (long long)carry_in[c * M + a];
        for (long long f = f0; f < f1; ++f) {
            const long long s = traj[f * M + a];
            int out = -2;
            if (!(f == 0 && first_frame_is_start)) {
                const bool known = unknown_as_jump ? true : (s != -1);
                if (known && s != last) { out = (int)last; ++cnt; }
            }
            from[f * M + a] = out;
            if (unknown_as_jump || s != -1) last = s;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(total, cnt);
}

// ordered compaction of the flat (frame-major, atom-minor) stream: (frame, atom, from, to) rows
constexpr int CBLOCK = 1024;

__global__ void k_jump_count(const int* __restrict__ from, long long n, unsigned* __restrict__ block_count) {
    __shared__ unsigned cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * CBLOCK + threadIdx.x;
    const bool j = (i < n) && from[i] != -2;
    const unsigned m = __ballot_sync(0xffffffffu, j);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0) block_count[blockIdx.x] = cnt;
}

__global__ void k_block_scan(const unsigned* __restrict__ block_count, long long n_blocks,
                             unsigned long long* __restrict__ block_offset) {
    // single CTA: serial over tiles of 1024 block counts, parallel scan inside a tile
    __shared__ unsigned long long tile[1024];
    __shared__ unsigned long long base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (long long t0 = 0; t0 < n_blocks; t0 += 1024) {
        const long long i = t0 + threadIdx.x;
        const unsigned long long v = (i < n_blocks) ? block_count[i] : 0ull;
        tile[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            unsigned long long add = 0;
            if ((int)threadIdx.x >= o) add = tile[threadIdx.x - o];
            __syncthreads();
            tile[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < n_blocks) block_offset[i] = base + tile[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) base += tile[1023];
        __syncthreads();
    }
}

__global__ void k_jump_write(const long long* __restrict__ traj, const int* __restrict__ from, long long n, int M,
                             long long frame0, const unsigned long long* __restrict__ block_offset,
                             long long* __restrict__ out, unsigned long long capacity) {
    __shared__ unsigned warp_cnt[CBLOCK / 32];
    const long long i = (long long)blockIdx.x * CBLOCK + threadIdx.x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fr = (i < n) ? from[i] : -2;
    const bool j = fr != -2;
    const unsigned m = __ballot_sync(0xffffffffu, j);
    if (lane == 0) warp_cnt[w] = __popc(m);
    __syncthreads();
    unsigned before = 0;
    for (int k = 0; k < w; ++k) before += warp_cnt[k];
    if (j) {
        const unsigned long long pos = block_offset[blockIdx.x] + before + __popc(m & lanemask_lt());
        if (pos < capacity) {
            out[4 * pos + 0] = frame0 + i / M;
            out[4 * pos + 1] = i % M;
            out[4 * pos + 2] = fr;
            out[4 * pos + 3] = traj[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// JumpAnalysis.run
// ------------------------------------------------------------------------------------------------
struct JaSummary {          // per (chunk, atom)
    int first_known_frame;  // local frame index, -1 if the atom is unknown throughout the chunk
    int first_label, last_label;
    int last_internal_jump; // last frame (local index in shard) with a jump decided inside the chunk, -1 none
};

__global__ void k_ja_summary(const long long* __restrict__ traj, long long F, int M, JaSummary* __restrict__ sum) {
    const long long c = blockIdx.x;
    const long long f0 = c * CHUNK, f1 = (f0 + CHUNK < F) ? (f0 + CHUNK) : F;
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
        JaSummary s = {-1, -1, -1, -1};
        int last = -1;
        for (long long f = f0; f < f1; ++f) {
            const long long v = traj[f * M + a];
            if (v < 0) continue;
            if (s.first_known_frame < 0) { s.first_known_frame = (int)f; s.first_label = (int)v; }
            else if ((int)v != last) s.last_internal_jump = (int)f;
            last = (int)v;
        }
        s.last_label = last;
        sum[c * M + a] = s;
    }
}

// carries into each chunk: last known label and the frame of the atom's last jump (-1: none yet)
__global__ void k_ja_carry(const JaSummary* __restrict__ sum, long long n_chunks, int M,
                           const long long* __restrict__ carry_label_in, const long long* __restrict__ carry_jump_in,
                           int2* __restrict__ carry) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= M) return;
    int lab = carry_label_in ? (int)carry_label_in[a] : -1;
    int jmp = carry_jump_in ? (int)carry_jump_in[a] : -1;     // local frame index (may be negative: previous shard)
    for (long long c = 0; c < n_chunks; ++c) {
        carry[c * M + a] = make_int2(lab, jmp);
        const JaSummary s = sum[c * M + a];
        if (s.first_known_frame >= 0) {
            if (lab >= 0 && s.first_label != lab) jmp = s.first_known_frame;
            if (s.last_internal_jump >= 0) jmp = s.last_internal_jump;
            lab = s.last_label;
        }
    }
}

// one shard's summary per atom, for the carry over frame shards: (first known frame, first label, last label, frame of
// the last jump decided inside the shard), local frame indices, -1 = none
__global__ void k_ja_shard_summary(const JaSummary* __restrict__ sum, long long n_chunks, int M, long long* __restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= M) return;
    long long first_frame = -1, first_label = -1, last_label = -1, last_jump = -1;
    int lab = -1;
    for (long long c = 0; c < n_chunks; ++c) {
        const JaSummary s = sum[c * M + a];
        if (s.first_known_frame < 0) continue;
        if (first_frame < 0) { first_frame = s.first_known_frame; first_label = s.first_label; }
        else if (s.first_label != lab) last_jump = s.first_known_frame;
        if (s.last_internal_jump >= 0) last_jump = s.last_internal_jump;
        lab = s.last_label;
        last_label = lab;
    }
    out[4 * a + 0] = first_frame; out[4 * a + 1] = first_label; out[4 * a + 2] = last_label; out[4 * a + 3] = last_jump;
}

// per element: last known site, filled current site, frames at current site (JumpAnalysis.py:63-93)
__global__ void k_ja_expand(const long long* __restrict__ traj, long long F, int M, const int2* __restrict__ carry,
                            int first_frame_is_start, int* __restrict__ last_o, int* __restrict__ cur_o,
                            int* __restrict__ time_o) {
    const long long c = blockIdx.x;
    const long long f0 = c * CHUNK, f1 = (f0 + CHUNK < F) ? (f0 + CHUNK) : F;
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
        const int2 cr = carry[c * M + a];
        int last = cr.x;
        long long jmp = cr.y;        // local frame index of the atom's last jump; global -1 = "never"
        for (long long f = f0; f < f1; ++f) {
            const long long v = traj[f * M + a];
            // at global frame 0 the reference starts with last_known = traj[0] (JumpAnalysis.py:46-47)
            int lk = last;
            if (f == 0 && first_frame_is_start) lk = (int)v;
            const int cur = (v < 0) ? lk : (int)v;
            // time_at_current read at frame f = f - (frame of the last jump), "never" being frame -1:
            // it starts at 1, grows by one per frame without a jump and restarts at 1 after one (:49,:91-93)
            const long long t = f - jmp;
            last_o[f * M + a] = lk;
            cur_o[f * M + a] = cur;
            time_o[f * M + a] = (int)t;
            if (cur >= 0 && lk >= 0 && cur != lk) jmp = f;
            if (v >= 0) last = (int)v;
        }
    }
}

// per frame accumulate with NumPy's buffered fancy-index semantics: a duplicated (i, j) pair inside
// one frame counts once, and for the lag sum the last duplicate's value wins (JumpAnalysis.py:75-88)
__global__ void k_ja_accumulate(const int* __restrict__ last_a, const int* __restrict__ cur_a,
                                const int* __restrict__ time_a, long long F, int M, int C,
                                double* __restrict__ n_ij, unsigned long long* __restrict__ total_time,
                                double* __restrict__ lag_sum, unsigned long long* __restrict__ lag_n,
                                unsigned long long* __restrict__ n_problems) {
    extern __shared__ int sh[];
    int* sl = sh;
    int* sc = sh + M;
    unsigned long long prob = 0;
    for (long long f = blockIdx.x; f < F; f += gridDim.x) {
        for (int a = threadIdx.x; a < M; a += blockDim.x) { sl[a] = last_a[f * M + a]; sc[a] = cur_a[f * M + a]; }
        __syncthreads();
        for (int a = threadIdx.x; a < M; a += blockDim.x) {
            const int l = sl[a], c = sc[a];
            if (!(l >= 0 && c >= 0)) { ++prob; continue; }
            bool last_pair = true, last_site = true;
            for (int b = a + 1; b < M; ++b) {
                if (sl[b] >= 0 && sc[b] >= 0) {
                    if (sc[b] == c) { last_site = false; if (sl[b] == l) last_pair = false; }
                }
            }
            if (last_site) atomicAdd(&total_time[c], 1ull);
            if (last_pair) {
                atomicAdd(&n_ij[(size_t)l * C + c], 1.0);
                if (l != c) {
                    atomicAdd(&lag_sum[(size_t)l * C + c], (double)time_a[f * M + a]);
                    atomicAdd(&lag_n[(size_t)l * C + c], 1ull);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) prob += __shfl_xor_sync(0xffffffffu, prob, o);
    if ((threadIdx.x & 31) == 0 && prob) atomicAdd(n_problems, prob);
}

// ------------------------------------------------------------------------------------------------
// site centres
// ------------------------------------------------------------------------------------------------
// wrapped (LandmarkAnalysis.py:182-189) positions of selected (frame, mobile) rows
__global__ void k_wrapped_rows(Cell cell, const double* __restrict__ frames, int A, int M,
                               const int* __restrict__ mobile_idx, long long frame0, long long n_frames,
                               const long long* __restrict__ rows, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long r = rows[i];
    const long long f = r / M - frame0;
    const int j = (int)(r % M);
    if (r < 0 || f < 0 || f >= n_frames) { out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = 0.0; return; }
    const double* p = frames + ((size_t)f * A + mobile_idx[j]) * 3;
    double x = p[0], y = p[1], z = p[2];
    if (cell.diag) wrap_point<true, false>(cell, x, y, z); else wrap_point<false, false>(cell, x, y, z);
    out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
}

// sums[s][0..2] += w * wrap(p + offset[s]),  sums[s][3] += w   (PBCCalculator.pyx:124-132)
// A thread follows one mobile atom through SITE_CHUNK consecutive frames and keeps the running sums of the
// site it currently sits at in registers: an atom stays at a site for many frames, so the global atomics drop
// from four per row to four per residence (a warp = 32 neighbouring atoms of one frame: coalesced loads).
constexpr int SITE_CHUNK = 64;
__global__ void k_site_accumulate(Cell cell, const double* __restrict__ frames, int A, int M,
                                  const int* __restrict__ mobile_idx, long long n_frames,
                                  const long long* __restrict__ labels, const double* __restrict__ confs,
                                  const double* __restrict__ offset, int C, int weighted, double* __restrict__ sums) {
    const long long n_chunks = (n_frames + SITE_CHUNK - 1) / SITE_CHUNK;
    const long long total = n_chunks * M;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % M);
        const long long f0 = (t / M) * SITE_CHUNK;
        const long long f1 = f0 + SITE_CHUNK < n_frames ? f0 + SITE_CHUNK : n_frames;
        const int atom = mobile_idx[j];
        long long cur = -1;
        double sx = 0.0, sy = 0.0, sz = 0.0, sw = 0.0, ox = 0.0, oy = 0.0, oz = 0.0;
        for (long long f = f0; f < f1; ++f) {
            const long long r = f * M + j;
            const long long s = labels[r];
            if (s < 0 || s >= C) continue;
            if (s != cur) {
                if (cur >= 0) {
                    atomicAdd(&sums[4 * cur + 0], sx); atomicAdd(&sums[4 * cur + 1], sy);
                    atomicAdd(&sums[4 * cur + 2], sz); atomicAdd(&sums[4 * cur + 3], sw);
                }
                cur = s; sx = sy = sz = sw = 0.0;
                ox = offset[3 * s]; oy = offset[3 * s + 1]; oz = offset[3 * s + 2];
            }
            const double* p = frames + ((size_t)f * A + atom) * 3;
            double x = p[0], y = p[1], z = p[2];
            if (cell.diag) wrap_point<true, false>(cell, x, y, z); else wrap_point<false, false>(cell, x, y, z);
            x = __dadd_rn(x, ox); y = __dadd_rn(y, oy); z = __dadd_rn(z, oz);
            if (cell.diag) wrap_point<true, false>(cell, x, y, z); else wrap_point<false, false>(cell, x, y, z);
            const double w = weighted ? confs[r] : 1.0;
            sx += w * x; sy += w * y; sz += w * z; sw += w;
        }
        if (cur >= 0) {
            atomicAdd(&sums[4 * cur + 0], sx); atomicAdd(&sums[4 * cur + 1], sy);
            atomicAdd(&sums[4 * cur + 2], sz); atomicAdd(&sums[4 * cur + 3], sw);
        }
    }
}

// centre = wrap(sum / weight - offset)   (PBCCalculator.pyx:132-135)
__global__ void k_site_finish(Cell cell, const double* __restrict__ sums, const double* __restrict__ offset, int C,
                              double* __restrict__ centers) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= C) return;
    const double w = sums[4 * s + 3];
    double x = sums[4 * s] / w - offset[3 * s];
    double y = sums[4 * s + 1] / w - offset[3 * s + 1];
    double z = sums[4 * s + 2] / w - offset[3 * s + 2];
    if (cell.diag) wrap_point<true, false>(cell, x, y, z); else wrap_point<false, false>(cell, x, y, z);
    centers[3 * s] = x; centers[3 * s + 1] = y; centers[3 * s + 2] = z;
}

// first (lowest) row assigned to each site: the centring point of the unweighted average
__global__ void k_first_row(const long long* __restrict__ labels, long long n, long long row0, int C,
                            unsigned long long* __restrict__ first) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const long long s = labels[r];
        if (s >= 0 && s < C) atomicMin(&first[s], (unsigned long long)(row0 + r));
    }
}

}  // namespace sitb

using namespace sitb;

#define CKT(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return sitb::set_error(SITB_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

extern "C" int sitb_check_multiple_occupancy(int device, const int64_t* dev_traj, int64_t n_frames, int32_t n_mobile,
                                             int64_t frame0, int32_t max_mobile_per_site, uint64_t* dev_out3,
                                             uint64_t* dev_first_bad, void* cuda_stream) {
    if (!dev_traj || !dev_out3 || !dev_first_bad || n_frames <= 0 || n_mobile <= 0)
        return set_error(SITB_E_INVALID, "sitb_check_multiple_occupancy: bad argument");
    CKT(cudaSetDevice(device));
    long long grid = n_frames < 148 * 8 ? n_frames : 148 * 8;
    const size_t occ_smem = sizeof(int) * (size_t)n_mobile;
    if (occ_smem > 200 * 1024) return set_error(SITB_E_LIMIT, "sitb_check_multiple_occupancy: %d mobile atoms (limit 51200)", n_mobile);
    if (occ_smem > 48 * 1024)
        CKT(cudaFuncSetAttribute(k_occupancy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)occ_smem));
    k_occupancy<<<(unsigned)grid, 128, occ_smem, (cudaStream_t)cuda_stream>>>(
        (const long long*)dev_traj, n_frames, n_mobile, frame0, max_mobile_per_site,
        (unsigned long long*)dev_out3, (unsigned long long*)dev_first_bad);
    CKT(cudaGetLastError());
    return SITB_OK;
}

extern "C" int sitb_jump_scan(int device, const int64_t* dev_traj, int64_t n_frames, int32_t n_mobile,
                              int32_t unknown_as_jump, int32_t first_frame_is_start, const int64_t* dev_carry_in,
                              int32_t* dev_from, uint64_t* dev_total, void* cuda_stream) {
    if (!dev_traj || !dev_from || !dev_total || n_frames <= 0 || n_mobile <= 0)
        return set_error(SITB_E_INVALID, "sitb_jump_scan: bad argument");
    CKT(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long n_chunks = (n_frames + CHUNK - 1) / CHUNK;
    int *chunk_last = nullptr, *carry = nullptr;
    CKT(pool_alloc((void**)&chunk_last, sizeof(int) * n_chunks * n_mobile, st));
    CKT(pool_alloc((void**)&carry, sizeof(int) * n_chunks * n_mobile, st));
    k_chunk_last<<<(unsigned)n_chunks, 128, 0, st>>>((const long long*)dev_traj, n_frames, n_mobile, chunk_last);
    k_chunk_carry<<<(n_mobile + 127) / 128, 128, 0, st>>>(chunk_last, n_chunks, n_mobile, (const long long*)dev_carry_in, carry);
    k_jump_from<<<(unsigned)n_chunks, 128, 0, st>>>((const long long*)dev_traj, n_frames, n_mobile, carry, unknown_as_jump,
                                                    first_frame_is_start, (const long long*)dev_carry_in, dev_from,
                                                    (unsigned long long*)dev_total);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    pool_free(chunk_last, st); pool_free(carry, st);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "sitb_jump_scan: %s", cudaGetErrorString(e));
    return SITB_OK;
}

extern "C" int sitb_jump_compact(int device, const int64_t* dev_traj, const int32_t* dev_from, int64_t n_frames,
                                 int32_t n_mobile, int64_t frame0, int64_t* dev_out, uint64_t capacity,
                                 void* cuda_stream) {
    if (!dev_traj || !dev_from || n_frames <= 0 || n_mobile <= 0 || (!dev_out && capacity))
        return set_error(SITB_E_INVALID, "sitb_jump_compact: bad argument");
    CKT(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long n = n_frames * n_mobile, n_blocks = (n + CBLOCK - 1) / CBLOCK;
    unsigned* bc = nullptr;
    unsigned long long* bo = nullptr;
    CKT(pool_alloc((void**)&bc, sizeof(unsigned) * n_blocks, st));
    CKT(pool_alloc((void**)&bo, sizeof(unsigned long long) * n_blocks, st));
    k_jump_count<<<(unsigned)n_blocks, CBLOCK, 0, st>>>(dev_from, n, bc);
    k_block_scan<<<1, 1024, 0, st>>>(bc, n_blocks, bo);
    k_jump_write<<<(unsigned)n_blocks, CBLOCK, 0, st>>>((const long long*)dev_traj, dev_from, n, n_mobile, frame0, bo,
                                                       (long long*)dev_out, capacity);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    pool_free(bc, st); pool_free(bo, st);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "sitb_jump_compact: %s", cudaGetErrorString(e));
    return SITB_OK;
}

extern "C" int sitb_jump_analysis(int device, const int64_t* dev_traj, int64_t n_frames, int32_t n_mobile,
                                  int32_t n_sites, int32_t first_frame_is_start, const int64_t* dev_carry_label,
                                  const int64_t* dev_carry_jump, double* dev_n_ij, uint64_t* dev_total_time,
                                  double* dev_lag_sum, uint64_t* dev_lag_n, uint64_t* dev_n_problems,
                                  void* cuda_stream) {
    if (!dev_traj || !dev_n_ij || !dev_total_time || !dev_lag_sum || !dev_lag_n || !dev_n_problems ||
        n_frames <= 0 || n_mobile <= 0 || n_sites <= 0)
        return set_error(SITB_E_INVALID, "sitb_jump_analysis: bad argument");
    CKT(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long n_chunks = (n_frames + CHUNK - 1) / CHUNK, n = n_frames * n_mobile;
    JaSummary* sum = nullptr;
    int2* carry = nullptr;
    int *la = nullptr, *ca = nullptr, *ta = nullptr;
    CKT(pool_alloc((void**)&sum, sizeof(JaSummary) * n_chunks * n_mobile, st));
    CKT(pool_alloc((void**)&carry, sizeof(int2) * n_chunks * n_mobile, st));
    CKT(pool_alloc((void**)&la, sizeof(int) * n, st));
    CKT(pool_alloc((void**)&ca, sizeof(int) * n, st));
    CKT(pool_alloc((void**)&ta, sizeof(int) * n, st));
    k_ja_summary<<<(unsigned)n_chunks, 128, 0, st>>>((const long long*)dev_traj, n_frames, n_mobile, sum);
    k_ja_carry<<<(n_mobile + 127) / 128, 128, 0, st>>>(sum, n_chunks, n_mobile, (const long long*)dev_carry_label,
                                                      (const long long*)dev_carry_jump, carry);
    k_ja_expand<<<(unsigned)n_chunks, 128, 0, st>>>((const long long*)dev_traj, n_frames, n_mobile, carry,
                                                    first_frame_is_start, la, ca, ta);
    long long grid = n_frames < 148 * 8 ? n_frames : 148 * 8;
    k_ja_accumulate<<<(unsigned)grid, 128, sizeof(int) * 2 * n_mobile, st>>>(
        la, ca, ta, n_frames, n_mobile, n_sites, dev_n_ij, (unsigned long long*)dev_total_time, dev_lag_sum,
        (unsigned long long*)dev_lag_n, (unsigned long long*)dev_n_problems);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    pool_free(sum, st); pool_free(carry, st); pool_free(la, st); pool_free(ca, st); pool_free(ta, st);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "sitb_jump_analysis: %s", cudaGetErrorString(e));
    return SITB_OK;
}

extern "C" int sitb_jump_analysis_summary(int device, const int64_t* dev_traj, int64_t n_frames, int32_t n_mobile,
                                          int64_t* dev_summary, void* cuda_stream) {
    if (!dev_traj || !dev_summary || n_frames <= 0 || n_mobile <= 0)
        return set_error(SITB_E_INVALID, "sitb_jump_analysis_summary: bad argument");
    CKT(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long n_chunks = (n_frames + CHUNK - 1) / CHUNK;
    JaSummary* sum = nullptr;
    CKT(pool_alloc((void**)&sum, sizeof(JaSummary) * n_chunks * n_mobile, st));
    k_ja_summary<<<(unsigned)n_chunks, 128, 0, st>>>((const long long*)dev_traj, n_frames, n_mobile, sum);
    k_ja_shard_summary<<<(n_mobile + 127) / 128, 128, 0, st>>>(sum, n_chunks, n_mobile, (long long*)dev_summary);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    pool_free(sum, st);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "sitb_jump_analysis_summary: %s", cudaGetErrorString(e));
    return SITB_OK;
}

// ---- launchers used by the context-bound entry points in sitb_api.cu ----------------------------
namespace sitb {

cudaError_t launch_wrapped_rows(const Cell& cell, const double* frames, int A, int M, const int* mobile_idx,
                                long long frame0, long long n_frames, const long long* rows, int n, double* out,
                                cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_wrapped_rows<<<(n + 127) / 128, 128, 0, st>>>(cell, frames, A, M, mobile_idx, frame0, n_frames, rows, n, out);
    return cudaGetLastError();
}

cudaError_t launch_site_accumulate(const Cell& cell, const double* frames, int A, int M, const int* mobile_idx,
                                   long long n_frames, const long long* labels, const double* confs,
                                   const double* offset, int C, int weighted, double* sums, int n_sms, cudaStream_t st) {
    if (n_frames <= 0) return cudaSuccess;
    k_site_accumulate<<<n_sms * 8, 256, 0, st>>>(cell, frames, A, M, mobile_idx, n_frames, labels, confs, offset, C,
                                                weighted, sums);
    return cudaGetLastError();
}

cudaError_t launch_site_finish(const Cell& cell, const double* sums, const double* offset, int C, double* centers,
                               cudaStream_t st) {
    if (C <= 0) return cudaSuccess;
    k_site_finish<<<(C + 127) / 128, 128, 0, st>>>(cell, sums, offset, C, centers);
    return cudaGetLastError();
}

cudaError_t launch_first_row(const long long* labels, long long n, long long row0, int C, unsigned long long* first,
                             int n_sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_first_row<<<n_sms * 8, 256, 0, st>>>(labels, n, row0, C, first);
    return cudaGetLastError();
}

}  // namespace sitb

// ---- PBCCalculator.average over a fixed point set with per-site weights ---------------------------
// LandmarkAnalysis.py:288-296 (SITE_CENTERS_REPRESENTATIVE_LANDMARK): per site, the periodic weighted
// average of the landmark centres whose weight is > 0, centred on the max-weight one (first maximum).
namespace sitb {

__global__ void k_weighted_point_average(Cell cell, const double* __restrict__ pts, const double* __restrict__ w,
                                         int C, int P, double* __restrict__ out) {
    const int s = blockIdx.x;
    if (s >= C) return;
    const double* ws = w + (size_t)s * P;
    __shared__ double red[4][128];
    __shared__ double bw[128];
    __shared__ int bi[128];
    // arg max weight, first maximum
    double mw = -1.0;
    int mi = 0x7FFFFFFF;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const double v = ws[i];
        if (v > 0.0 && (v > mw)) { mw = v; mi = i; }
    }
    bw[threadIdx.x] = mw; bi[threadIdx.x] = mi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double ow = bw[threadIdx.x + o];
            const int oi = bi[threadIdx.x + o];
            if (ow > bw[threadIdx.x] || (ow == bw[threadIdx.x] && oi < bi[threadIdx.x])) { bw[threadIdx.x] = ow; bi[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    const int anchor = bi[0];
    if (anchor == 0x7FFFFFFF) {   // no positive weight: np.average would raise; report NaN
        if (threadIdx.x == 0) out[3 * s] = out[3 * s + 1] = out[3 * s + 2] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    const double ox = __dsub_rn(cell.cen[0], pts[3 * anchor]);
    const double oy = __dsub_rn(cell.cen[1], pts[3 * anchor + 1]);
    const double oz = __dsub_rn(cell.cen[2], pts[3 * anchor + 2]);
    double sx = 0, sy = 0, sz = 0, sw = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const double v = ws[i];
        if (!(v > 0.0)) continue;
        double x = __dadd_rn(pts[3 * i], ox), y = __dadd_rn(pts[3 * i + 1], oy), z = __dadd_rn(pts[3 * i + 2], oz);
        if (cell.diag) wrap_point<true, false>(cell, x, y, z); else wrap_point<false, false>(cell, x, y, z);
        sx += v * x; sy += v * y; sz += v * z; sw += v;
    }
    red[0][threadIdx.x] = sx; red[1][threadIdx.x] = sy; red[2][threadIdx.x] = sz; red[3][threadIdx.x] = sw;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int c = 0; c < 4; ++c) red[c][threadIdx.x] += red[c][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double x = red[0][0] / red[3][0] - ox, y = red[1][0] / red[3][0] - oy, z = red[2][0] / red[3][0] - oz;
        if (cell.diag) wrap_point<true, false>(cell, x, y, z); else wrap_point<false, false>(cell, x, y, z);
        out[3 * s] = x; out[3 * s + 1] = y; out[3 * s + 2] = z;
    }
}

cudaError_t launch_weighted_point_average(const Cell& cell, const double* pts, const double* w, int C, int P,
                                          double* out, cudaStream_t st) {
    if (C <= 0) return cudaSuccess;
    k_weighted_point_average<<<C, 128, 0, st>>>(cell, pts, w, C, P, out);
    return cudaGetLastError();
}

// PBCCalculator.distances (PBCCalculator.pyx:64-103) for every (a_i, b_j): shift so that a_i sits on the cell centroid,
// wrap, distance to the centroid
__global__ void k_pbc_distances(Cell cell, const double* __restrict__ a, const double* __restrict__ b, int na, int nb,
                                double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)na * nb) return;
    const int i = (int)(idx / nb), j = (int)(idx % nb);
    const double ox = __dsub_rn(cell.cen[0], a[3 * i]), oy = __dsub_rn(cell.cen[1], a[3 * i + 1]), oz = __dsub_rn(cell.cen[2], a[3 * i + 2]);
    const double q = cell.diag ? shifted_dist2<true, false>(cell, b[3 * j], b[3 * j + 1], b[3 * j + 2], ox, oy, oz)
                               : shifted_dist2<false, false>(cell, b[3 * j], b[3 * j + 1], b[3 * j + 2], ox, oy, oz);
    out[idx] = __dsqrt_rn(q);
}

// the cell block of a context-free call: cellmat = cell^T and its inverse, row major, as PBCCalculator.__init__ builds them
static Cell cell_from_host(const double* cellmat, const double* cellmat_inv) {
    Cell c;
    for (int i = 0; i < 9; ++i) { c.c[i] = cellmat[i]; c.ci[i] = cellmat_inv[i]; }
    for (int k = 0; k < 3; ++k) c.cen[k] = (0.5 * c.c[3 * k + 0] + 0.5 * c.c[3 * k + 1]) + 0.5 * c.c[3 * k + 2];
    bool diag = true;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (i != j && (c.c[3 * i + j] != 0.0 || c.ci[3 * i + j] != 0.0)) diag = false;
    c.diag = diag ? 1 : 0;
    return c;
}

}  // namespace sitb

extern "C" int sitb_pbc_distances(int device, const double* host_cellmat, const double* host_cellmat_inv, const double* dev_a,
                                  const double* dev_b, int32_t na, int32_t nb, double* dev_out, void* cuda_stream) {
    if (!host_cellmat || !host_cellmat_inv || !dev_a || !dev_b || !dev_out || na <= 0 || nb <= 0)
        return sitb::set_error(SITB_E_INVALID, "sitb_pbc_distances: bad argument");
    CKT(cudaSetDevice(device));
    const long long n = (long long)na * nb;
    sitb::k_pbc_distances<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(
        sitb::cell_from_host(host_cellmat, host_cellmat_inv), dev_a, dev_b, na, nb, dev_out);
    CKT(cudaGetLastError());
    return SITB_OK;
}

extern "C" int sitb_pbc_weighted_average(int device, const double* host_cellmat, const double* host_cellmat_inv,
                                         const double* dev_points, const double* dev_weights, int32_t n_sets, int32_t n_points,
                                         double* dev_out, void* cuda_stream) {
    if (!host_cellmat || !host_cellmat_inv || !dev_points || !dev_weights || !dev_out || n_sets <= 0 || n_points <= 0)
        return sitb::set_error(SITB_E_INVALID, "sitb_pbc_weighted_average: bad argument");
    CKT(cudaSetDevice(device));
    CKT(sitb::launch_weighted_point_average(sitb::cell_from_host(host_cellmat, host_cellmat_inv), dev_points, dev_weights, n_sets,
                                            n_points, dev_out, (cudaStream_t)cuda_stream));
    return SITB_OK;
}

