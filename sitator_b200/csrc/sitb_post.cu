// sitator_b200 -- integer post-processing of the assignment stream traj [n_frames][n_mobile] int64 (-1 unknown):
//
//   SiteTrajectory.assign_to_last_known_site     SiteTrajectory.py:235-304
//   SmoothSiteTrajectory: running_windowed_mode  dynamics/SmoothSiteTrajectory.pyx:79-111
//   RemoveUnoccupiedSites: seen mask / relabel   dynamics/RemoveUnoccupiedSites.py:30-57
//
// All three are per mobile atom.  The first is sequential in time ("last known site", "frames since it was last
// known"): frames are cut into chunks, a thread per (chunk, atom) summarises its chunk, a thread per atom chains
// the summaries, then a thread per (chunk, atom) replays its chunk with the carried state.  Threads of a warp are
// neighbouring atoms of the same chunk, so every frame row is read coalesced.
#include "../../include/sitator_b200.h"
#include "sitb_common.cuh"

namespace sitb {
int set_error(int code, const char* fmt, ...);

static constexpr int LK_CHUNK = 128;

// chunk summary: last known label inside the chunk (-1: none) and the unknown frames after it (or the whole chunk)
__global__ void k_lk_summary(const long long* __restrict__ traj, long long F, int M, long long n_chunks,
                             long long* __restrict__ sum_label, long long* __restrict__ sum_tail) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_chunks * M) return;
    const long long c = t / M;
    const int m = (int)(t - c * M);
    const long long f0 = c * LK_CHUNK, f1 = (f0 + LK_CHUNK < F) ? f0 + LK_CHUNK : F;
    long long lk = -1, tail = 0;
    for (long long f = f0; f < f1; ++f) {
        const long long s = traj[f * M + m];
        if (s == -1) ++tail; else { lk = s; tail = 0; }
    }
    sum_label[t] = lk;
    sum_tail[t] = tail;
}

// per atom: state before every chunk (in place over the summaries) and after the last one
__global__ void k_lk_chain(long long* __restrict__ sum_label, long long* __restrict__ sum_tail, int M, long long n_chunks,
                           const long long* __restrict__ carry_label, const long long* __restrict__ carry_time,
                           long long* __restrict__ end_label, long long* __restrict__ end_time) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    long long lk = carry_label ? carry_label[m] : -1;
    long long tu = carry_time ? carry_time[m] : 0;
    for (long long c = 0; c < n_chunks; ++c) {
        const long long l = sum_label[c * M + m], tail = sum_tail[c * M + m];
        sum_label[c * M + m] = lk;
        sum_tail[c * M + m] = tu;
        if (l != -1) { lk = l; tu = tail; } else { tu += tail; }
    }
    if (end_label) end_label[m] = lk;
    if (end_time) end_time[m] = tu;
}

// replay a chunk with its carried state (SiteTrajectory.py:262-283)
__global__ void k_lk_apply(long long* __restrict__ traj, long long F, int M, long long n_chunks, long long frame0,
                           long long threshold, const long long* __restrict__ in_label,
                           const long long* __restrict__ in_time, unsigned long long* __restrict__ stats) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long reassigned = 0, sum_times = 0, n_times = 0, maxkey = 0;
    if (t < n_chunks * M) {
        const long long c = t / M;
        const int m = (int)(t - c * M);
        const long long f0 = c * LK_CHUNK, f1 = (f0 + LK_CHUNK < F) ? f0 + LK_CHUNK : F;
        long long lk = in_label[t], tu = in_time[t];
        for (long long f = f0; f < f1; ++f) {
            const long long s = traj[f * M + m];
            if (s != -1) {
                lk = s;                                                   // :266
                if (tu != 0) {                                            // :268-276 an unknown stretch ends here
                    sum_times += (unsigned long long)tu;
                    ++n_times;
                    if (tu > threshold) {
                        const unsigned long long key = ((unsigned long long)(frame0 + f) << 24) | (unsigned long long)(tu < 0xFFFFFF ? tu : 0xFFFFFF);
                        if (key > maxkey) maxkey = key;
                    }
                }
                tu = 0;                                                   // :278
            } else {
                if (tu < threshold) { traj[f * M + m] = lk; ++reassigned; }   // :280-283 (lk may still be -1)
                ++tu;                                                     // :284
            }
        }
    }
    // warp-level then global accumulation
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        reassigned += __shfl_xor_sync(0xffffffffu, reassigned, o);
        sum_times += __shfl_xor_sync(0xffffffffu, sum_times, o);
        n_times += __shfl_xor_sync(0xffffffffu, n_times, o);
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, maxkey, o);
        if (ok > maxkey) maxkey = ok;
    }
    if ((threadIdx.x & 31) == 0) {
        if (reassigned) atomicAdd(&stats[0], reassigned);
        if (sum_times) atomicAdd(&stats[1], sum_times);
        if (n_times) atomicAdd(&stats[2], n_times);
        if (maxkey) atomicMax(&stats[3], maxkey);
    }
}

// running_windowed_mode (SmoothSiteTrajectory.pyx:79-111): per (frame, atom) the most frequent label in frames
// [max(f - wleft, 0), min(f + wright, F)) -- unknown counts as a label; ties go to the lowest label, unknown first
// (the reference scans its count buffer from "unknown" upwards with a strict >).  A CTA takes 256 frames of one
// atom; the window lives in shared memory.
__global__ void __launch_bounds__(256) k_windowed_mode(const long long* __restrict__ traj, long long* __restrict__ out,
                                                       long long F, int M, int wleft, int wright, long long threshold,
                                                       int replace_unknown, long long halo_before,
                                                       const long long* __restrict__ before, long long halo_after,
                                                       const long long* __restrict__ after) {
    extern __shared__ int tile[];                      // [256 + wleft + wright] labels of frames f0 - wleft ...
    const int m = blockIdx.y;
    const long long f0 = (long long)blockIdx.x * blockDim.x;
    const int span = (int)blockDim.x + wleft + wright;
    // frames outside [0, F) come from the neighbouring shards' halos when given (before: the last halo_before frames
    // of the previous shards, after: the first halo_after frames of the next), else they do not exist (-2)
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const long long f = f0 - wleft + i;
        int v = -2;
        if (f >= 0 && f < F) v = (int)traj[f * M + m];
        else if (f < 0 && before && -f <= halo_before) v = (int)before[(halo_before + f) * M + m];
        else if (f >= F && after && f - F < halo_after) v = (int)after[(f - F) * M + m];
        tile[i] = v;
    }
    __syncthreads();
    const long long f = f0 + threadIdx.x;
    if (f >= F) return;
    const int lo = threadIdx.x, hi = threadIdx.x + wleft + wright;      // window = tile[lo, hi)
    int winner = -1, best = 0;
    for (int i = lo; i < hi; ++i) {
        const int s = tile[i];
        if (s == -2) continue;
        int n = 0;
        for (int j = lo; j < hi; ++j) n += (tile[j] == s);
        if (n > best || (n == best && s < winner)) { best = n; winner = s; }
    }
    long long r;
    if (best >= threshold) r = winner;
    else r = replace_unknown ? -1 : (long long)tile[threadIdx.x + wleft];
    out[f * M + m] = r;
}

// RemoveUnoccupiedSites: which sites occur at all
__global__ void k_seen_sites(const long long* __restrict__ traj, long long n, int n_sites, unsigned* __restrict__ seen) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long s = traj[i];
        if (s >= 0 && s < n_sites && seen[s] == 0u) seen[s] = 1u;
    }
}

// traj[i] = translation[traj[i]] (unknown stays unknown)
__global__ void k_relabel(long long* __restrict__ traj, long long n, int n_sites, const long long* __restrict__ translation) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long s = traj[i];
        if (s >= 0 && s < n_sites) traj[i] = translation[s];
    }
}

// RecenterTrajectory (util/RecenterTrajectory.pyx:61-100): per frame, subtract the mass-weighted centre of the
// atoms with factor != 0 (accumulated over the atoms IN ORDER with the reference's operation order, so the centre
// is bit-identical), then add `shift` (the cell centroid; zero for velocities).  One warp per frame: lanes 0..2
// chain the sum of one coordinate each, then all lanes update the frame.
__global__ void __launch_bounds__(256) k_recenter(double* __restrict__ arr, long long F, int A, const double* __restrict__ w,
                                                  double shift_x, double shift_y, double shift_z, int add_shift) {
    const int lane = threadIdx.x & 31;
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= F) return;
    double* fr = arr + (size_t)f * A * 3;
    double com = 0.0;
    if (lane < 3)
        for (int j = 0; j < A; ++j) com = __dadd_rn(com, __dmul_rn(w[j], fr[3 * j + lane]));   // :92-94
    const double cx = __shfl_sync(0xffffffffu, com, 0), cy = __shfl_sync(0xffffffffu, com, 1), cz = __shfl_sync(0xffffffffu, com, 2);
    for (int i = lane; i < 3 * A; i += 32) {
        const int d = i % 3;
        double v = __dsub_rn(fr[i], d == 0 ? cx : (d == 1 ? cy : cz));                          // :97-99
        if (add_shift) v = __dadd_rn(v, d == 0 ? shift_x : (d == 1 ? shift_y : shift_z));       // :56-57
        fr[i] = v;
    }
}

}  // namespace sitb

using namespace sitb;

#define CKP(call, what)                                                                                  \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return set_error(SITB_E_CUDA, "%s: %s", what, cudaGetErrorString(e_));     \
    } while (0)

extern "C" int sitb_assign_last_known(int device, int64_t* dev_traj, int64_t n_frames, int32_t n_mobile, int64_t frame0,
                                      int64_t frame_threshold, const int64_t* dev_carry_label,
                                      const int64_t* dev_carry_time, int64_t* dev_end_label, int64_t* dev_end_time,
                                      uint64_t* dev_stats4, int32_t apply, void* stream) {
    if (!dev_traj || n_frames < 0 || n_mobile <= 0 || (apply && !dev_stats4))
        return set_error(SITB_E_INVALID, "sitb_assign_last_known: bad argument");
    if (frame0 + n_frames >= (1ll << 40)) return set_error(SITB_E_LIMIT, "sitb_assign_last_known: more than 2^40 frames");
    CKP(cudaSetDevice(device), "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_chunks = (n_frames + LK_CHUNK - 1) / LK_CHUNK;
    const long long n_items = n_chunks * n_mobile;
    long long *d_label = nullptr, *d_tail = nullptr;
    CKP(pool_alloc((void**)&d_label, sizeof(long long) * (size_t)(n_items ? n_items : 1), st), "scratch");
    cudaError_t e = pool_alloc((void**)&d_tail, sizeof(long long) * (size_t)(n_items ? n_items : 1), st);
    if (e != cudaSuccess) { pool_free(d_label, st); return set_error(SITB_E_CUDA, "scratch: %s", cudaGetErrorString(e)); }
    if (n_items > 0)
        k_lk_summary<<<(unsigned)((n_items + 255) / 256), 256, 0, st>>>((const long long*)dev_traj, n_frames, n_mobile, n_chunks,
                                                                      d_label, d_tail);
    k_lk_chain<<<(n_mobile + 127) / 128, 128, 0, st>>>(d_label, d_tail, n_mobile, n_chunks, (const long long*)dev_carry_label,
                                                      (const long long*)dev_carry_time, (long long*)dev_end_label,
                                                      (long long*)dev_end_time);
    if (apply && n_items > 0)
        k_lk_apply<<<(unsigned)((n_items + 255) / 256), 256, 0, st>>>((long long*)dev_traj, n_frames, n_mobile, n_chunks, frame0,
                                                                    frame_threshold, d_label, d_tail,
                                                                    (unsigned long long*)dev_stats4);
    e = cudaGetLastError();
    pool_free(d_label, st); pool_free(d_tail, st);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "sitb_assign_last_known: %s", cudaGetErrorString(e));
    return SITB_OK;
}

extern "C" int sitb_windowed_mode(int device, const int64_t* dev_traj, int64_t* dev_out, int64_t n_frames, int32_t n_mobile,
                                  int32_t wleft, int32_t wright, int64_t threshold, int32_t replace_no_winner_unknown,
                                  int64_t halo_before, const int64_t* dev_before, int64_t halo_after,
                                  const int64_t* dev_after, void* stream) {
    if (!dev_traj || !dev_out || n_frames < 0 || n_mobile <= 0 || wleft < 0 || wright < 0)
        return set_error(SITB_E_INVALID, "sitb_windowed_mode: bad argument");
    if (wleft + wright > 8192) return set_error(SITB_E_LIMIT, "sitb_windowed_mode: window of %d frames (limit 8192)", wleft + wright);
    if (n_frames == 0) return SITB_OK;
    CKP(cudaSetDevice(device), "cudaSetDevice");
    const size_t smem = sizeof(int) * (size_t)(256 + wleft + wright);
    dim3 grid((unsigned)((n_frames + 255) / 256), (unsigned)n_mobile);
    k_windowed_mode<<<grid, 256, smem, (cudaStream_t)stream>>>((const long long*)dev_traj, (long long*)dev_out, n_frames, n_mobile,
                                                             wleft, wright, threshold, replace_no_winner_unknown, halo_before,
                                                             (const long long*)dev_before, halo_after,
                                                             (const long long*)dev_after);
    CKP(cudaGetLastError(), "k_windowed_mode");
    return SITB_OK;
}

extern "C" int sitb_seen_sites(int device, const int64_t* dev_traj, int64_t n_entries, int32_t n_sites, uint32_t* dev_seen,
                               void* stream) {
    if (!dev_traj || !dev_seen || n_entries < 0 || n_sites <= 0) return set_error(SITB_E_INVALID, "sitb_seen_sites: bad argument");
    if (n_entries == 0) return SITB_OK;
    CKP(cudaSetDevice(device), "cudaSetDevice");
    long long blocks = (n_entries + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_seen_sites<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)dev_traj, n_entries, n_sites, dev_seen);
    CKP(cudaGetLastError(), "k_seen_sites");
    return SITB_OK;
}

extern "C" int sitb_relabel_sites(int device, int64_t* dev_traj, int64_t n_entries, int32_t n_sites,
                                  const int64_t* dev_translation, void* stream) {
    if (!dev_traj || !dev_translation || n_entries < 0 || n_sites <= 0)
        return set_error(SITB_E_INVALID, "sitb_relabel_sites: bad argument");
    if (n_entries == 0) return SITB_OK;
    CKP(cudaSetDevice(device), "cudaSetDevice");
    long long blocks = (n_entries + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_relabel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((long long*)dev_traj, n_entries, n_sites,
                                                                (const long long*)dev_translation);
    CKP(cudaGetLastError(), "k_relabel");
    return SITB_OK;
}

extern "C" int sitb_recenter(int device, double* dev_array, int64_t n_frames, int32_t n_atoms, const double* dev_weights,
                             const double* host_shift3, void* stream) {
    if (!dev_array || !dev_weights || n_frames < 0 || n_atoms <= 0) return set_error(SITB_E_INVALID, "sitb_recenter: bad argument");
    if (n_frames == 0) return SITB_OK;
    CKP(cudaSetDevice(device), "cudaSetDevice");
    const double sx = host_shift3 ? host_shift3[0] : 0.0, sy = host_shift3 ? host_shift3[1] : 0.0, sz = host_shift3 ? host_shift3[2] : 0.0;
    k_recenter<<<(unsigned)((n_frames + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dev_array, n_frames, n_atoms, dev_weights, sx, sy, sz,
                                                                             host_shift3 ? 1 : 0);
    CKP(cudaGetLastError(), "k_recenter");
    return SITB_OK;
}
