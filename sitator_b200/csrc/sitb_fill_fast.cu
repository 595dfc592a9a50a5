// sitator_b200 -- K1f: first tier of the two-tier fused fill + assign pass (orthorhombic cells, candidate grid).
//
// The same path as k_fill<DIAG, MODE_ASSIGN> (sitb_fill.cu):
//   LandmarkAnalysis.py:182-189   wrap                                  -> fractional coordinates, float
//   helpers.pyx:55-92             static-lattice check / dynamic map    -> FP32 screen; anything unusual leaves the frame
//                                                                          to the exact kernel
//   helpers.pyx:95-114, :134-212  shift, wrap, logistic-cutoff product  -> FP32 (MUFU sqrt / ex2 / lg2)
//   DotProdClassifier.pyx:166-189 |centre . x| argmax + threshold       -> FP32 with an error bound
//
// Two tiers: every landmark-vector component is evaluated in FP32 with a proven bound on its error
// (sitb_api.cu: fast_tables).  A row's DECISIONS are taken from the FP32 values only when they hold for every value
// inside the bound:
//   support    q * ib > 1       =>  d^2 > Q in exact arithmetic (component is zero, helpers.pyx:199-203)
//              q * ib <= kappa  =>  d^2 <= Q in exact arithmetic; in between the row is undecided
//   arg-max    best - second > 2 B,   B = tau * sum |value * weight|   (DotProdClassifier.pyx:181)
//   threshold  best - B > thr  or  best + B < thr                      (DotProdClassifier.pyx:184-186)
// Undecided rows (and whole frames with a static atom beyond the candidate-grid margin, an ambiguous dynamic
// lattice map, or more than 64 non-zero components) are flagged; the caller then runs the exact float64 kernel on
// exactly those rows (FillParams::row_filter).  Labels therefore equal the exact kernel's by construction;
// confidences of the rows decided here carry the FP32 error (<= tau * sum |value * weight|, ~1e-5).
//
// Work decomposition as in k_fill: a CTA takes a batch of frames, its warps claim (frame, mobile atom) tasks.  Per
// task: FP32 squared distances to the static sites of the atom's grid box; every candidate landmark of the box
// (sorted by cluster, landmarks of no cluster left out: sitb_api.cu) is tested against its cut-off with one
// compare on max_h q_h * ib_h; survivors are compacted, their values evaluated by one lane each, and the per-cluster
// sums fall out of one segmented warp scan because the list is sorted by cluster.
#include "sitb_fill.cuh"
#include <math_constants.h>

namespace sitb {

struct FastSmem {
    int Spad, Mpad;
    size_t off_va, off_ib, off_ac, off_cw, off_if, off_fs, off_fm, off_lmap, off_seen, off_warp, warp_bytes, off_hist,
        off_misc, total;
};

static constexpr int SURV_CAP = 64;

__host__ __device__ inline FastSmem fast_layout(int S, int M, int Lpad, int NB, int warps, int fb, int n_clusters,
                                                int dynamic, int with_hist) {
    FastSmem l;
    l.Spad = (S + 4) & ~3;          // room for the dummy site S
    l.Mpad = (M + 3) & ~3;
    size_t o = 0;
    l.off_ib = o;   o += sizeof(float4) * (size_t)NB * Lpad;
    l.off_ac = o;   o += sizeof(float4) * (size_t)NB * Lpad;
    l.off_va = o;   o += sizeof(ushort4) * (size_t)NB * Lpad;
    l.off_cw = o;   o += sizeof(float2) * (size_t)Lpad;
    l.off_if = o;   o += sizeof(float) * 3 * (size_t)l.Spad;
    l.off_fs = o;   o += sizeof(float) * 3 * (size_t)l.Spad * fb;
    l.off_fm = o;   o += sizeof(float) * 3 * (size_t)l.Mpad * fb;
    l.off_lmap = o; o += dynamic ? sizeof(unsigned) * (size_t)l.Spad * fb : 0;
    l.off_seen = o; o += dynamic ? sizeof(unsigned) * (size_t)l.Spad * fb : 0;
    l.warp_bytes = (sizeof(float) * (size_t)l.Spad + sizeof(unsigned) * SURV_CAP + 15) & ~(size_t)15;
    o = (o + 15) & ~(size_t)15;
    l.off_warp = o; o += l.warp_bytes * (size_t)warps;
    l.off_hist = o; o += with_hist ? sizeof(unsigned) * (size_t)(n_clusters > 0 ? n_clusters : 1) : 0;
    l.off_misc = o; o += sizeof(int) * (size_t)(fb + 1);
    l.total = (o + 15) & ~(size_t)15;
    return l;
}

__device__ __forceinline__ float f_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// u - round(u) for |u| < 2^22
__device__ __forceinline__ float cfrac(float u) {
    const float magic = 12582912.0f;   // 1.5 * 2^23
    return __fsub_rn(u, __fsub_rn(__fadd_rn(u, magic), magic));
}

__device__ __forceinline__ float dist2f(float ax, float ay, float az, float bx, float by, float bz, float Lx, float Ly,
                                        float Lz) {
    const float cx = cfrac(ax - bx) * Lx, cy = cfrac(ay - by) * Ly, cz = cfrac(az - bz) * Lz;
    return fmaf(cz, cz, fmaf(cy, cy, cx * cx));
}

__device__ __forceinline__ void flag_row(const FastParams& p, long long row, long long frame, int reason) {
    p.recheck[row] = 1;
    if (atomicExch(&p.frame_flag[frame], 1) == 0) {
        const unsigned long long at = atomicAdd(p.n_list, 1ull);
        p.frame_list[at] = frame;
    }
    atomicAdd(&p.counters[reason], 1ull);
    atomicAdd(&p.counters[RECHECK_ROWS], 1ull);
}

__global__ void __launch_bounds__(1024, 1) k_assign_fast(const __grid_constant__ FastParams p, const int FB,
                                                         const __grid_constant__ FastSmem lay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int S = p.S, M = p.M, Lpad = p.Lpad, NB = p.NB;
    const int Spad = lay.Spad, Mpad = lay.Mpad;
    float4* tib = (float4*)(smem_raw + lay.off_ib);
    float4* tac = (float4*)(smem_raw + lay.off_ac);
    ushort4* tva = (ushort4*)(smem_raw + lay.off_va);
    float2* tcw = (float2*)(smem_raw + lay.off_cw);
    float* idf = (float*)(smem_raw + lay.off_if);          // [3][Spad]
    float* fs = (float*)(smem_raw + lay.off_fs);           // [FB][3][Spad]
    float* fm = (float*)(smem_raw + lay.off_fm);           // [FB][3][Mpad]
    unsigned* lmap_all = (unsigned*)(smem_raw + lay.off_lmap);
    unsigned* seen_all = (unsigned*)(smem_raw + lay.off_seen);
    unsigned char* wscratch = smem_raw + lay.off_warp + (size_t)warp * lay.warp_bytes;
    float* qfw = (float*)wscratch;                         // [Spad] squared screen distances, by static site
    unsigned* surv = (unsigned*)(wscratch + sizeof(float) * (size_t)Spad);   // [SURV_CAP]
    unsigned* hist = (unsigned*)(smem_raw + lay.off_hist);
    int* task_counter = (int*)(smem_raw + lay.off_misc);
    int* glevel = task_counter + 1;                        // [FB]

    for (int i = threadIdx.x; i < NB * Lpad; i += blockDim.x) {
        tib[i] = p.tab.ib[i];
        tac[i] = p.tab.ac[i];
        tva[i] = p.tab.va[i];
    }
    for (int i = threadIdx.x; i < Lpad; i += blockDim.x) tcw[i] = p.tab.cw[i];
    for (int i = threadIdx.x; i < 3 * Spad; i += blockDim.x) idf[i] = p.ideal_frac[i];
    if (p.counts)
        for (int i = threadIdx.x; i < p.n_clusters; i += blockDim.x) hist[i] = 0u;
    if (lane == 0) qfw[S] = 0.f;                           // dummy vertex: distance 0
    __syncthreads();

    const float Lx = p.Lx, Ly = p.Ly, Lz = p.Lz;
    const int n_levels = p.n_levels;
    const float m0 = p.grid[0].margin_sq, m1 = p.grid[1].margin_sq, lim = p.static_lim_sq;
    const float bc = p.bc, kappa = p.kappa;

    for (long long w0 = (long long)blockIdx.x * FB; w0 < p.n_work; w0 += (long long)gridDim.x * FB) {
        const int nb = (int)((p.n_work - w0 < FB) ? (p.n_work - w0) : FB);

        // ---- 1. fractional coordinates of the batch's atoms (LandmarkAnalysis.py:182-189; float64, then rounded) ----
        for (int t = threadIdx.x; t < nb * (S + M); t += blockDim.x) {
            const int b = t / (S + M), r = t - b * (S + M);
            const double* __restrict__ fr = p.frames + (size_t)(w0 + b) * (size_t)p.A * 3;
            const int a = (r < S) ? p.static_idx[r] : p.mobile_idx[r - S];
            double f0 = fr[3 * a + 0] * p.ci0, f1 = fr[3 * a + 1] * p.ci1, f2 = fr[3 * a + 2] * p.ci2;
            f0 -= floor(f0); f1 -= floor(f1); f2 -= floor(f2);
            if (r < S) {
                float* d = fs + (size_t)b * 3 * Spad;
                d[r] = (float)f0; d[Spad + r] = (float)f1; d[2 * Spad + r] = (float)f2;
            } else {
                float* d = fm + (size_t)b * 3 * Mpad;
                d[r - S] = (float)f0; d[Mpad + r - S] = (float)f1; d[2 * Mpad + r - S] = (float)f2;
            }
        }
        if (p.dynamic)
            for (int t = threadIdx.x; t < nb * Spad; t += blockDim.x) seen_all[t] = 0u;
        if (threadIdx.x == 0) *task_counter = 0;
        if (threadIdx.x < FB) glevel[threadIdx.x] = 0;
        __syncthreads();

        // ---- 2. static lattice (helpers.pyx:55-92): FP32 screen; frames it cannot clear go to the exact kernel ----
        if (!p.dynamic) {
            for (int t = threadIdx.x; t < nb * S; t += blockDim.x) {
                const int b = t / S, s = t - b * S;
                const float* fsb = fs + (size_t)b * 3 * Spad;
                const float q = dist2f(fsb[s], fsb[Spad + s], fsb[2 * Spad + s], idf[s], idf[Spad + s], idf[2 * Spad + s], Lx, Ly, Lz);
                const int lvl = (q > lim) ? n_levels : ((q > m0) ? ((q > m1) ? n_levels : 1) : 0);
                if (lvl) atomicMax(&glevel[b], lvl);
            }
        } else {
            for (int t = warp; t < nb * S; t += nwarps) {
                const int b = t / S, li = t - b * S;
                const float* fsb = fs + (size_t)b * 3 * Spad;
                const float ix = idf[li], iy = idf[Spad + li], iz = idf[2 * Spad + li];
                float qmin = CUDART_INF_F;
                for (int j = lane; j < S; j += 32)
                    qmin = fminf(qmin, dist2f(fsb[j], fsb[Spad + j], fsb[2 * Spad + j], ix, iy, iz, Lx, Ly, Lz));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) qmin = fminf(qmin, __shfl_xor_sync(0xffffffffu, qmin, o));
                // an atom whose float distance is within the float error of the smallest could be the exact arg-min
                const float dc = p.dyn_dc;
                const float bound = qmin + 2.0f * (2.0f * sqrtf(3.0f * qmin) * dc + 3.0f * dc * dc + 1e-6f * qmin) + 1e-12f;
                int cnt = 0, bj = 0;
                for (int j = lane; j < S; j += 32) {
                    const float q = dist2f(fsb[j], fsb[Spad + j], fsb[2 * Spad + j], ix, iy, iz, Lx, Ly, Lz);
                    if (!(q > bound)) { ++cnt; bj = j; }
                }
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                bj = __reduce_max_sync(0xffffffffu, bj);
                if (lane == 0) {
                    lmap_all[(size_t)b * Spad + li] = (unsigned)bj;
                    atomicAdd(&seen_all[(size_t)b * Spad + bj], 1u);
                    int lvl = (qmin > lim) ? n_levels : ((qmin > m0) ? ((qmin > m1) ? n_levels : 1) : 0);
                    if (cnt != 1) lvl = n_levels;                 // ambiguous nearest atom
                    if (lvl) atomicMax(&glevel[b], lvl);
                }
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nb * S; t += blockDim.x) {
                const int b = t / S, s = t - b * S;
                if (seen_all[(size_t)b * Spad + s] != 1u) atomicMax(&glevel[b], n_levels);   // unassigned / doubly assigned
            }
        }
        __syncthreads();

        // ---- 3. one warp per (frame, mobile atom) ------------------------------------------------------------
        for (;;) {
            int jj = 0;
            if (lane == 0) jj = atomicAdd(task_counter, 1);
            jj = __shfl_sync(0xffffffffu, jj, 0);
            if (jj >= nb * M) break;
            const int b = p.m_magic ? (int)__umulhi((unsigned)jj, p.m_magic) : jj / M;
            const int j = jj - b * M;
            const long long row = w0 * M + jj;
            const int lv = glevel[b];
            if (lv >= n_levels) {
                if (lane == 0) flag_row(p, row, w0 + b, RECHECK_FRAME);
                continue;
            }
            const float* fsb = fs + (size_t)b * 3 * Spad;
            const float* fmb = fm + (size_t)b * 3 * Mpad;
            const unsigned* lmap = lmap_all + (size_t)b * Spad;
            const float mx = fmb[j], my = fmb[Mpad + j], mz = fmb[2 * Mpad + j];
            int ix = (int)(mx * (float)p.gx), iy = (int)(my * (float)p.gy), iz = (int)(mz * (float)p.gz);
            ix = ix < 0 ? 0 : (ix >= p.gx ? p.gx - 1 : ix);
            iy = iy < 0 ? 0 : (iy >= p.gy ? p.gy - 1 : iy);
            iz = iz < 0 ? 0 : (iz >= p.gz ? p.gz - 1 : iz);
            const int box = (ix * p.gy + iy) * p.gz + iz;
            const FastGrid& g = p.grid[lv];
            const uint2 sp = __ldg(g.sbox + box);
            const uint2 cp = __ldg(g.cbox + box);

            // 3a. squared distances to the box's static sites (helpers.pyx:99-103, :174-178)
            for (unsigned i = lane; i < sp.y; i += 32) {
                const int s = (int)__ldg(g.slist + sp.x + i);
                const int src = p.dynamic ? (int)lmap[s] : s;
                qfw[s] = dist2f(fsb[src], fsb[Spad + src], fsb[2 * Spad + src], mx, my, mz, Lx, Ly, Lz);
            }
            __syncwarp();

            // 3b. cut-off test of every candidate (helpers.pyx:197-203) on max_h q_h / Q_h; survivors compacted
            int nsurv = 0;
            unsigned amb_any = 0u;
            for (unsigned i0 = 0; i0 < cp.y; i0 += 32) {
                const unsigned i = i0 + lane;
                const bool in = i < cp.y;
                const unsigned rec = in ? __ldg(g.clist + cp.x + i) : 0u;
                const int k = (int)(rec & 0xFFFFu);
                float m = 0.f;
                for (int blk = 0; blk < NB; ++blk) {
                    const ushort4 vv = tva[(size_t)blk * Lpad + k];
                    const float4 bb = tib[(size_t)blk * Lpad + k];
                    m = fmaxf(fmaxf(fmaxf(m, qfw[vv.x] * bb.x), fmaxf(qfw[vv.y] * bb.y, qfw[vv.z] * bb.z)), qfw[vv.w] * bb.w);
                }
                const bool keep = in && !(m > 1.0f);
                const unsigned km = __ballot_sync(0xffffffffu, keep);
                amb_any |= __ballot_sync(0xffffffffu, keep && (m > kappa));
                if (keep) {
                    const int q = nsurv + __popc(km & lanemask_lt());
                    if (q < SURV_CAP) surv[q] = rec;
                }
                nsurv += __popc(km);
            }
            if (amb_any != 0u || nsurv > SURV_CAP) {
                if (lane == 0) flag_row(p, row, w0 + b, amb_any ? RECHECK_SUPPORT : RECHECK_LONG);
                continue;
            }
            __syncwarp();

            // 3c. values of the survivors (helpers.pyx:205-212), one lane each, times the centre weight
            float best = 0.f, second = 0.f, sumabs = 0.f;
            int bestid = 0;
            int cid[2];
            float pr[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                cid[c] = 0x7FFF;
                pr[c] = 0.f;
                const int e = 32 * c + lane;
                if (c * 32 < nsurv && e < nsurv) {
                    const unsigned rec = surv[e];
                    const int k = (int)(rec & 0xFFFFu);
                    cid[c] = (int)(rec >> 16);
                    float P = 1.0f;
                    for (int blk = 0; blk < NB; ++blk) {
                        const ushort4 vv = tva[(size_t)blk * Lpad + k];
                        const float4 aa = tac[(size_t)blk * Lpad + k];
                        float e0 = f_ex2(fmaf(f_sqrt(qfw[vv.x]), aa.x, -bc));
                        float e1 = f_ex2(fmaf(f_sqrt(qfw[vv.y]), aa.y, -bc));
                        float e2 = f_ex2(fmaf(f_sqrt(qfw[vv.z]), aa.z, -bc));
                        float e3 = f_ex2(fmaf(f_sqrt(qfw[vv.w]), aa.w, -bc));
                        P = fmaf(P, e0, P); P = fmaf(P, e1, P); P = fmaf(P, e2, P); P = fmaf(P, e3, P);
                    }
                    const float2 cw = tcw[k];
                    pr[c] = f_ex2(f_lg2(P) * cw.y) * cw.x;
                }
            }
            if (nsurv <= 32) {
                // 3d. per-cluster sums by a segmented scan (the list is sorted by cluster), then the two largest |sums|
                const int prev = __shfl_up_sync(0xffffffffu, cid[0], 1);
                const bool head = (lane == 0) || (cid[0] != prev);
                const unsigned heads = __ballot_sync(0xffffffffu, head);
                const int dist = lane - (31 - __clz((int)(heads & (lanemask_lt() | (1u << lane)))));
                float v = pr[0];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float t = __shfl_up_sync(0xffffffffu, v, o);
                    if (dist >= o) v += t;
                }
                const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
                const unsigned ab = (tail && lane < nsurv) ? __float_as_uint(fabsf(v)) : 0u;
                const unsigned mxb = __reduce_max_sync(0xffffffffu, ab);
                const unsigned wm = __ballot_sync(0xffffffffu, ab == mxb);
                const int wl = __ffs((int)wm) - 1;                 // lowest lane = lowest cluster id (np.argmax: first maximum)
                bestid = __shfl_sync(0xffffffffu, cid[0], wl);
                const unsigned sb = __reduce_max_sync(0xffffffffu, (lane == wl) ? 0u : ab);
                best = __uint_as_float(mxb);
                second = __uint_as_float(sb);
                sumabs = (float)__reduce_add_sync(0xffffffffu, __float2uint_ru(fabsf(pr[0]) * 1048576.0f)) * (1.0f / 1048576.0f);
            } else {
                // rare (more than 32 non-zero components): peel the clusters one at a time
                sumabs = (float)__reduce_add_sync(0xffffffffu, __float2uint_ru((fabsf(pr[0]) + fabsf(pr[1])) * 1048576.0f)) *
                         (1.0f / 1048576.0f);
                for (;;) {
                    const int mine = min(cid[0], cid[1]);
                    const int cur = __reduce_min_sync(0xffffffffu, mine);
                    if (cur == 0x7FFF) break;
                    float part = 0.f;
                    if (cid[0] == cur) { part += pr[0]; cid[0] = 0x7FFF; }
                    if (cid[1] == cur) { part += pr[1]; cid[1] = 0x7FFF; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    const float a = fabsf(part);
                    if (a > best) { second = best; best = a; bestid = cur; }      // ascending ids: ties keep the lower
                    else if (a > second) second = a;
                }
            }

            // 3e. decide (DotProdClassifier.pyx:181-186) if the decision holds for every value inside the error bound
            const float B = p.tau * sumabs + 2.4e-7f * p.thr;
            int decision;                   // 1 assigned, 0 unassigned, -1 undecided (top-2), -2 undecided (threshold)
            if (best + B < p.thr) decision = 0;
            else if (best - B > p.thr) decision = (best - second > 2.0f * B) ? 1 : -1;
            else decision = -2;
            if (lane == 0) {
                if (decision < 0) {
                    flag_row(p, row, w0 + b, decision == -1 ? RECHECK_MARGIN : RECHECK_THRESHOLD);
                } else {
                    p.labels[row] = decision ? (long long)bestid : -1ll;
                    p.confs[row] = decision ? (double)best : 0.0;
                    if (decision && p.counts) atomicAdd(&hist[bestid], 1u);
                }
            }
            __syncwarp();
        }
        __syncthreads();
    }
    if (p.counts) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.n_clusters; i += blockDim.x)
            if (hist[i]) atomicAdd(&p.counts[i], (unsigned long long)hist[i]);
    }
}

cudaError_t launch_assign_fast(const FastParams& p, int n_sms, cudaStream_t stream) {
    if (p.n_work <= 0) return cudaSuccess;
    if (p.n_levels < 1) return cudaErrorInvalidConfiguration;
    const size_t budget = (size_t)227 * 1024 - 1024;
    int best_w = 0, best_fb = 0;
    size_t best_bytes = 0;
    for (int w = 32; w >= 4 && !best_w; w -= 4) {
        for (int fb = 16; fb >= 1; --fb) {
            if (fb > 1 && (long long)fb * p.M > 16LL * w && (long long)(fb - 1) * p.M >= 8LL * w) continue;   // long enough
            const size_t bytes = fast_layout(p.S, p.M, p.Lpad, p.NB, w, fb, p.n_clusters, p.dynamic, p.counts != nullptr).total;
            if (bytes <= budget) { best_w = w; best_fb = fb; best_bytes = bytes; break; }
        }
    }
    if (!best_w) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k_assign_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)best_bytes);
    if (e != cudaSuccess) return e;
    const long long batches = (p.n_work + best_fb - 1) / best_fb;
    long long grid = n_sms;
    if (grid > batches) grid = batches;
    const FastSmem lay = fast_layout(p.S, p.M, p.Lpad, p.NB, best_w, best_fb, p.n_clusters, p.dynamic, p.counts != nullptr);
    k_assign_fast<<<(unsigned)grid, best_w * 32, best_bytes, stream>>>(p, best_fb, lay);
    return cudaGetLastError();
}

// ---- candidate lists of the first tier: the grid's landmark lists sorted by cluster, unclustered landmarks left out ----
// one thread per box: filter, then insertion sort on (cluster << 16 | landmark) in the output segment
__global__ void k_sort_box_lists(const unsigned* __restrict__ ptr, const uint16_t* __restrict__ list, const int* __restrict__ cid,
                                 long long cells, uint2* __restrict__ cbox, unsigned* __restrict__ clist) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= cells) return;
    const unsigned beg = ptr[id], end = ptr[id + 1];
    unsigned* out = clist + beg;
    unsigned n = 0;
    for (unsigned i = beg; i < end; ++i) {
        const unsigned k = list[i];
        const int c = cid[k];
        if (c < 0) continue;
        const unsigned key = ((unsigned)c << 16) | k;
        unsigned pos = n;
        while (pos > 0 && out[pos - 1] > key) { out[pos] = out[pos - 1]; --pos; }
        out[pos] = key;
        ++n;
    }
    cbox[id] = make_uint2(beg, n);
}

cudaError_t launch_sort_box_lists(const unsigned* ptr, const uint16_t* list, const int* cid, long long cells, uint2* cbox,
                                  unsigned* clist, cudaStream_t stream) {
    if (cells <= 0) return cudaSuccess;
    k_sort_box_lists<<<(unsigned)((cells + 127) / 128), 128, 0, stream>>>(ptr, list, cid, cells, cbox, clist);
    return cudaGetLastError();
}

}  // namespace sitb
