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
// (sitb_api.cu: sitb_create, "error model").  A row's DECISIONS are taken from the FP32 values only when they hold
// for every value inside the bound:
//   support    q * ib > 1       =>  d^2 > Q in exact arithmetic (component is zero, helpers.pyx:199-203)
//              q * ib <= kappa  =>  d^2 <= Q in exact arithmetic; in between the row is undecided
//   arg-max    best - second > 2 B,   B = tau * sum |value * weight|   (DotProdClassifier.pyx:181)
//   threshold  best - B > thr  or  best + B < thr                      (DotProdClassifier.pyx:184-186)
// Undecided rows (and whole frames with a static atom beyond the candidate-grid margin, an ambiguous dynamic
// lattice map, or more than 64 non-zero components) are flagged; the caller then runs the exact float64 kernel on
// exactly those rows (FillParams::row_filter).  Labels therefore equal the exact kernel's by construction;
// confidences of the rows decided here carry the FP32 error (<= tau * sum |value * weight|, measured ~1e-6).
//
// Work decomposition as in k_fill: a CTA takes a batch of frames, its warps claim (frame, mobile atom) tasks.  Per
// task: FP32 squared distances to the static sites of the atom's grid box; every candidate landmark of the box
// (sorted by cluster, landmarks of no cluster left out: k_sort_box_lists) is tested against its cut-off with one
// compare on max_h q_h * ib_h; survivors are compacted, their values evaluated by one lane each, and the per-cluster
// sums fall out of one segmented warp scan because the list is sorted by cluster.
//
// Shared-memory layout is chosen for few address computations: one record per landmark (reciprocal bounds, slopes,
// vertex byte offsets, centre weight) read with immediate offsets from one base; positions as float4 per atom.
#include "sitb_fill.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace sitb {

static constexpr int SURV_CAP = 64;

// record of a landmark in shared memory, in 16-byte units: [0, NB) reciprocal bounds, [NB, 2 NB) slopes,
// then NB x 8 bytes of vertex byte offsets (4 * site, uint16) and (centre weight, -1 / n_vertices)
__host__ __device__ constexpr int rec_units(int NB) { return 2 * NB + (NB * 8 + 8 + 15) / 16; }

struct FastSmem {
    int Spad;                  // floats per warp for the squared distances (S + 1 rounded up)
    unsigned off_rec, off_if, off_fs, off_fm, off_lmap, off_seen, off_warp, warp_bytes, off_hist, off_misc, total;
};

__host__ __device__ inline FastSmem fast_layout(int S, int M, int Lpad, int NB, int warps, int fb, int n_clusters,
                                                int dynamic, int with_hist) {
    FastSmem l;
    l.Spad = (S + 4) & ~3;          // room for the dummy site S
    size_t o = 0;
    l.off_rec = (unsigned)o;  o += 16 * (size_t)rec_units(NB) * Lpad;
    l.off_if = (unsigned)o;   o += sizeof(float4) * (size_t)S;
    l.off_fs = (unsigned)o;   o += sizeof(float4) * (size_t)S * fb;
    l.off_fm = (unsigned)o;   o += sizeof(float4) * (size_t)M * fb;
    l.off_lmap = (unsigned)o; o += dynamic ? sizeof(unsigned) * (size_t)l.Spad * fb : 0;
    l.off_seen = (unsigned)o; o += dynamic ? sizeof(unsigned) * (size_t)l.Spad * fb : 0;
    l.warp_bytes = (unsigned)((sizeof(float) * (size_t)l.Spad + sizeof(unsigned) * SURV_CAP + 15) & ~(size_t)15);
    o = (o + 15) & ~(size_t)15;
    l.off_warp = (unsigned)o; o += (size_t)l.warp_bytes * warps;
    l.off_hist = (unsigned)o; o += with_hist ? sizeof(unsigned) * (size_t)(n_clusters > 0 ? n_clusters : 1) : 0;
    l.off_misc = (unsigned)o; o += sizeof(int) * (size_t)(fb + 1);
    l.total = (unsigned)((o + 15) & ~(size_t)15);
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

__device__ __forceinline__ float dist2f(const float4 a, const float4 b, float Lx, float Ly, float Lz) {
    const float cx = cfrac(a.x - b.x) * Lx, cy = cfrac(a.y - b.y) * Ly, cz = cfrac(a.z - b.z) * Lz;
    return fmaf(cz, cz, fmaf(cy, cy, cx * cx));
}

// shared-memory accesses by byte address (one base register + immediate offsets)
__device__ __forceinline__ float lds_f(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_f4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_u2(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds_f2(unsigned a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __noinline__ void flag_row(const FastParams& p, unsigned row, long long frame, int reason) {
    p.recheck[row] = 1;
    if (atomicExch(&p.frame_flag[frame], 1) == 0) {
        const unsigned long long at = atomicAdd(p.n_list, 1ull);
        p.frame_list[at] = frame;
    }
    atomicAdd(&p.counters[reason], 1ull);
    atomicAdd(&p.counters[RECHECK_ROWS], 1ull);
}

template <int NB, bool DYN>
__global__ void __launch_bounds__(1024, 1) k_assign_fast(const __grid_constant__ FastParams p, const int FB,
                                                         const __grid_constant__ FastSmem lay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RU = rec_units(NB);
    constexpr unsigned REC = 16u * RU;            // bytes per landmark record
    constexpr unsigned OFF_AC = 16u * NB, OFF_VA = 32u * NB, OFF_CW = 32u * NB + 8u * NB;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int S = p.S, M = p.M;
    const int Spad = lay.Spad;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned a_rec = sbase + lay.off_rec;
    float4* idf = (float4*)(smem_raw + lay.off_if);          // [S]
    float4* fs = (float4*)(smem_raw + lay.off_fs);           // [FB][S]
    float4* fm = (float4*)(smem_raw + lay.off_fm);           // [FB][M]
    unsigned* lmap_all = (unsigned*)(smem_raw + lay.off_lmap);
    unsigned* seen_all = (unsigned*)(smem_raw + lay.off_seen);
    const unsigned a_q = sbase + lay.off_warp + (unsigned)warp * lay.warp_bytes;      // float[Spad], by 4 * static site
    const unsigned a_surv = a_q + 4u * (unsigned)Spad;                                // unsigned[SURV_CAP]
    unsigned* hist = (unsigned*)(smem_raw + lay.off_hist);
    int* task_counter = (int*)(smem_raw + lay.off_misc);
    const unsigned a_task = sbase + lay.off_misc;
    int* glevel = task_counter + 1;                          // [FB]

    // ---- stage the landmark records once per CTA ----
    for (int i = threadIdx.x; i < p.Lpad; i += blockDim.x) {
        uint4* r = (uint4*)(smem_raw + lay.off_rec) + (size_t)i * RU;
        unsigned tail[4 * (RU - 2 * NB)];
#pragma unroll
        for (int blk = 0; blk < NB; ++blk) {
            const float4 b = p.tab.ib[(size_t)blk * p.Lpad + i], a = p.tab.ac[(size_t)blk * p.Lpad + i];
            r[blk] = make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
            r[NB + blk] = make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w));
            const ushort4 v = p.tab.va[(size_t)blk * p.Lpad + i];
            tail[2 * blk] = (4u * v.x) | ((4u * v.y) << 16);
            tail[2 * blk + 1] = (4u * v.z) | ((4u * v.w) << 16);
        }
        const float2 cw = p.tab.cw[i];
        tail[2 * NB] = __float_as_uint(cw.x);
        tail[2 * NB + 1] = __float_as_uint(cw.y);
#pragma unroll
        for (int t = 2 * NB + 2; t < 4 * (RU - 2 * NB); ++t) tail[t] = 0u;
#pragma unroll
        for (int t = 0; t < RU - 2 * NB; ++t) r[2 * NB + t] = make_uint4(tail[4 * t], tail[4 * t + 1], tail[4 * t + 2], tail[4 * t + 3]);
    }
    for (int i = threadIdx.x; i < S; i += blockDim.x) idf[i] = p.ideal_frac[i];
    if (p.counts)
        for (int i = threadIdx.x; i < p.n_clusters; i += blockDim.x) hist[i] = 0u;
    if (lane == 0) sts_f(a_q + 4u * (unsigned)S, 0.f);       // dummy vertex: distance 0
    __syncthreads();

    const float Lx = p.Lx, Ly = p.Ly, Lz = p.Lz;
    const float m0 = p.margin_sq[0], m1 = p.margin_sq[1], lim = p.static_lim_sq;
    const float bc = p.bc, kappa = p.kappa;
    const int SM = S + M;

    for (long long w0 = (long long)blockIdx.x * FB; w0 < p.n_work; w0 += (long long)gridDim.x * FB) {
        const int nb = (int)((p.n_work - w0 < FB) ? (p.n_work - w0) : FB);

        // ---- 1. fractional coordinates of the batch's atoms (LandmarkAnalysis.py:182-189; float64, then rounded) ----
        for (int t = threadIdx.x; t < nb * SM; t += blockDim.x) {
            const int b = (int)__umulhi((unsigned)t, p.sm_magic), r = t - b * SM;
            const double* __restrict__ fr = p.frames + (size_t)(w0 + b) * (size_t)p.A * 3;
            const int a = (r < S) ? p.static_idx[r] : p.mobile_idx[r - S];
            double f0 = fr[3 * a + 0] * p.ci0, f1 = fr[3 * a + 1] * p.ci1, f2 = fr[3 * a + 2] * p.ci2;
            f0 -= floor(f0); f1 -= floor(f1); f2 -= floor(f2);
            const float4 v = make_float4((float)f0, (float)f1, (float)f2, 0.f);
            if (r < S) fs[b * S + r] = v;
            else fm[b * M + (r - S)] = v;
        }
        if (DYN)
            for (int t = threadIdx.x; t < nb * Spad; t += blockDim.x) seen_all[t] = 0u;
        if (threadIdx.x == 0) *task_counter = 0;
        if (threadIdx.x < FB) glevel[threadIdx.x] = 0;
        __syncthreads();

        // ---- 2. static lattice (helpers.pyx:55-92): FP32 screen; frames it cannot clear go to the exact kernel ----
        if (!DYN) {
            for (int b = 0; b < nb; ++b)
                for (int s = threadIdx.x; s < S; s += blockDim.x) {
                    const float q = dist2f(fs[b * S + s], idf[s], Lx, Ly, Lz);
                    const int lvl = (q > lim) ? 2 : ((q > m0) ? ((q > m1) ? 2 : 1) : 0);
                    if (lvl) atomicMax(&glevel[b], lvl);
                }
        } else {
            for (int t = warp; t < nb * S; t += nwarps) {
                const int b = t / S, li = t - b * S;
                const float4* fsb = fs + b * S;
                const float4 id = idf[li];
                float qmin = CUDART_INF_F;
                for (int j = lane; j < S; j += 32) qmin = fminf(qmin, dist2f(fsb[j], id, Lx, Ly, Lz));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) qmin = fminf(qmin, __shfl_xor_sync(0xffffffffu, qmin, o));
                // an atom whose float distance is within the float error of the smallest could be the exact arg-min
                const float dc = p.dyn_dc;
                const float bound = qmin + 2.0f * (2.0f * sqrtf(3.0f * qmin) * dc + 3.0f * dc * dc + 1e-6f * qmin) + 1e-12f;
                int cnt = 0, bj = 0;
                for (int j = lane; j < S; j += 32)
                    if (!(dist2f(fsb[j], id, Lx, Ly, Lz) > bound)) { ++cnt; bj = j; }
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                bj = __reduce_max_sync(0xffffffffu, bj);
                if (lane == 0) {
                    lmap_all[b * Spad + li] = (unsigned)bj;
                    atomicAdd(&seen_all[b * Spad + bj], 1u);
                    int lvl = (qmin > lim) ? 2 : ((qmin > m0) ? ((qmin > m1) ? 2 : 1) : 0);
                    if (cnt != 1) lvl = 2;                        // ambiguous nearest atom
                    if (lvl) atomicMax(&glevel[b], lvl);
                }
            }
            __syncthreads();
            for (int b = 0; b < nb; ++b)
                for (int s = threadIdx.x; s < S; s += blockDim.x)
                    if (seen_all[b * Spad + s] != 1u) atomicMax(&glevel[b], 2);      // unassigned / doubly assigned
        }
        __syncthreads();

        // ---- 3. one warp per (frame, mobile atom) ------------------------------------------------------------
        const int ntask = nb * M;
        const unsigned row0 = (unsigned)(w0 * M);
        for (;;) {
            int jj = 0;
            if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(jj) : "r"(a_task) : "memory");
            jj = __shfl_sync(0xffffffffu, jj, 0);
            if (jj >= ntask) break;
            const int b = (int)__umulhi((unsigned)jj, p.m_magic);
            const unsigned row = row0 + (unsigned)jj;          // < 2^32 landmark vectors per launch (sitb_api.cu: base_params)
            const int lv = glevel[b];
            if (lv >= 2) {
                if (lane == 0) flag_row(p, row, w0 + b, RECHECK_FRAME);
                continue;
            }
            const unsigned a_fs = sbase + lay.off_fs + 16u * (unsigned)(b * S);
            const float4 mp = fm[jj];                          // fm is [b][M]: index b * M + j = jj
            const int ix = min((int)(mp.x * p.gxf), p.gx - 1), iy = min((int)(mp.y * p.gyf), p.gy - 1),
                      iz = min((int)(mp.z * p.gzf), p.gz - 1);
            const int box = lv * p.cells + (ix * p.gy + iy) * p.gz + iz;
            const uint2 sp = __ldg(p.sbox + box);
            const uint2 cp = __ldg(p.cbox + box);

            // 3a. squared distances to the box's static sites (helpers.pyx:99-103, :174-178)
            {
                const uint16_t* sl = p.slist + sp.x;
#pragma unroll 1
                for (unsigned i0 = 0; i0 < sp.y; i0 += 32) {
                    const unsigned i = i0 + lane;
                    if (i < sp.y) {
                        const unsigned s4 = __ldg(sl + i);         // 4 * site
                        unsigned src4 = s4;
                        if (DYN) src4 = 4u * lmap_all[b * Spad + (s4 >> 2)];
                        sts_f(a_q + s4, dist2f(lds_f4(a_fs + 4u * src4), mp, Lx, Ly, Lz));
                    }
                }
            }
            __syncwarp();

            // 3b. cut-off test of every candidate (helpers.pyx:197-203) on max_h q_h / Q_h; survivors compacted
            int nsurv = 0;
            unsigned amb_any = 0u;
            {
                const unsigned* cl = p.clist + cp.x;
#pragma unroll 1
                for (unsigned i0 = 0; i0 < cp.y; i0 += 32) {
                    const unsigned i = i0 + lane;
                    const bool in = i < cp.y;
                    unsigned rec = 0u;
                    if (in) rec = __ldg(cl + i);
                    const unsigned ar = a_rec + (rec & 0xFFFFu) * REC;
                    float m = 0.f;
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
                        const uint2 vv = lds_u2(ar + OFF_VA + 8u * blk);
                        const float4 bb = lds_f4(ar + 16u * blk);
                        const float q0 = lds_f(a_q + (vv.x & 0xFFFFu)), q1 = lds_f(a_q + (vv.x >> 16));
                        const float q2 = lds_f(a_q + (vv.y & 0xFFFFu)), q3 = lds_f(a_q + (vv.y >> 16));
                        m = fmaxf(fmaxf(fmaxf(m, q0 * bb.x), fmaxf(q1 * bb.y, q2 * bb.z)), q3 * bb.w);
                    }
                    const bool keep = in && !(m > 1.0f);
                    const unsigned km = __ballot_sync(0xffffffffu, keep);
                    amb_any |= __ballot_sync(0xffffffffu, keep && (m > kappa));
                    const int q = nsurv + __popc(km & lanemask_lt());
                    if (keep && q < SURV_CAP) sts_u(a_surv + 4u * (unsigned)q, rec);
                    nsurv += __popc(km);
                }
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
                if ((c == 0 || nsurv > 32) && e < nsurv) {
                    const unsigned rec = lds_u(a_surv + 4u * (unsigned)e);
                    const unsigned ar = a_rec + (rec & 0xFFFFu) * REC;
                    cid[c] = (int)(rec >> 16);
                    float P = 1.0f;
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
                        const uint2 vv = lds_u2(ar + OFF_VA + 8u * blk);
                        const float4 aa = lds_f4(ar + OFF_AC + 16u * blk);
                        const float e0 = f_ex2(fmaf(f_sqrt(lds_f(a_q + (vv.x & 0xFFFFu))), aa.x, -bc));
                        const float e1 = f_ex2(fmaf(f_sqrt(lds_f(a_q + (vv.x >> 16))), aa.y, -bc));
                        const float e2 = f_ex2(fmaf(f_sqrt(lds_f(a_q + (vv.y & 0xFFFFu))), aa.z, -bc));
                        const float e3 = f_ex2(fmaf(f_sqrt(lds_f(a_q + (vv.y >> 16))), aa.w, -bc));
                        P = fmaf(P, e0, P); P = fmaf(P, e1, P); P = fmaf(P, e2, P); P = fmaf(P, e3, P);
                    }
                    const float2 cw = lds_f2(ar + OFF_CW);
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
            if (lane == 0) {
                const float B = p.tau * sumabs + 2.4e-7f * p.thr;
                const bool assigned = best - B > p.thr;
                if (best + B < p.thr || (assigned && best - second > 2.0f * B)) {
                    p.labels[row] = assigned ? (long long)bestid : -1ll;
                    p.confs[row] = assigned ? (double)best : 0.0;
                    if (assigned && p.counts) atomicAdd(&hist[bestid], 1u);
                } else {
                    flag_row(p, row, w0 + b, assigned ? RECHECK_MARGIN : RECHECK_THRESHOLD);
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

template <int NB, bool DYN>
static cudaError_t launch_fast_one(const FastParams& p, int n_sms, cudaStream_t stream) {
    // CTAs per SM x warps x frames per batch: as many resident warps as registers (32 per thread) and the 227 KB of
    // shared memory allow (the landmark records are per CTA), batches of ~12 tasks per warp
    auto kern = k_assign_fast<NB, DYN>;
    int force_ctas = 0;
    if (const char* env = getenv("SITB_FAST_CTAS")) force_ctas = atoi(env);       // developer knob
    int tasks_per_warp = 12;
    if (const char* env = getenv("SITB_FAST_TPW")) tasks_per_warp = atoi(env);   // developer knob
    int best_w = 0, best_fb = 0, best_ctas = 0;
    size_t best_bytes = 0;
    double best_score = -1.0;
    for (int ctas = 2; ctas >= 1; --ctas) {
        if (force_ctas && ctas != force_ctas) continue;
        const size_t budget = (size_t)(227 * 1024) / ctas - 1024;
        for (int w = 32; w >= 4; w -= 4) {
            int fb_fit = 0;
            size_t bytes_fit = 0;
            for (int fb = 1; fb <= 16; ++fb) {
                const size_t bytes = fast_layout(p.S, p.M, p.Lpad, NB, w, fb, p.n_clusters, DYN, p.counts != nullptr).total;
                if (bytes > budget) break;
                fb_fit = fb; bytes_fit = bytes;
                if ((long long)fb * p.M >= (long long)tasks_per_warp * w) break;
            }
            if (!fb_fit) continue;
            const double fill = (double)(fb_fit * p.M) / (6.0 * w);      // short batches: the CTA barrier costs
            const double score = (double)(ctas * w) * (fill < 1.0 ? fill : 1.0);
            if (score > best_score) { best_score = score; best_w = w; best_fb = fb_fit; best_ctas = ctas; best_bytes = bytes_fit; }
        }
    }
    if (!best_w) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)best_bytes);
    if (e != cudaSuccess) return e;
    const long long batches = (p.n_work + best_fb - 1) / best_fb;
    long long grid = (long long)n_sms * best_ctas;
    if (grid > batches) grid = batches;
    const FastSmem lay = fast_layout(p.S, p.M, p.Lpad, NB, best_w, best_fb, p.n_clusters, DYN, p.counts != nullptr);
    kern<<<(unsigned)grid, best_w * 32, best_bytes, stream>>>(p, best_fb, lay);
    return cudaGetLastError();
}

cudaError_t launch_assign_fast(const FastParams& p, int n_sms, cudaStream_t stream) {
    if (p.n_work <= 0) return cudaSuccess;
    if (p.S > 16000 || p.M < 2 || p.M >= 16384 || (long long)p.S + p.M >= 65536) return cudaErrorInvalidConfiguration;
    if (p.NB == 1) return p.dynamic ? launch_fast_one<1, true>(p, n_sms, stream) : launch_fast_one<1, false>(p, n_sms, stream);
    if (p.NB == 2) return p.dynamic ? launch_fast_one<2, true>(p, n_sms, stream) : launch_fast_one<2, false>(p, n_sms, stream);
    return cudaErrorInvalidConfiguration;
}

// ---- candidate lists of the first tier: the grid's landmark lists sorted by cluster, unclustered landmarks left out ----
// one thread per box: filter, then insertion sort on (cluster << 16 | landmark) in the output segment
__global__ void k_sort_box_lists(const unsigned* __restrict__ ptr, const uint16_t* __restrict__ list, const int* __restrict__ cid,
                                 long long cells, unsigned base, uint2* __restrict__ cbox, unsigned* __restrict__ clist) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= cells) return;
    const unsigned beg = ptr[id], end = ptr[id + 1];
    unsigned* out = clist + base + beg;
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
    cbox[id] = make_uint2(base + beg, n);
}

cudaError_t launch_sort_box_lists(const unsigned* ptr, const uint16_t* list, const int* cid, long long cells, unsigned base,
                                  uint2* cbox, unsigned* clist, cudaStream_t stream) {
    if (cells <= 0) return cudaSuccess;
    k_sort_box_lists<<<(unsigned)((cells + 127) / 128), 128, 0, stream>>>(ptr, list, cid, cells, base, cbox, clist);
    return cudaGetLastError();
}

}  // namespace sitb
