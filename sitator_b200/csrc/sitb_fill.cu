// sitator_b200 -- K1: fused wrap + static-lattice check + landmark-vector fill (+ assign).
//
// Replaces, per frame and fused in one kernel (reference file:line):
//   LandmarkAnalysis.py:182-189   wrap a copy of the frame into the cell       (step 1 below)
//   helpers.pyx:55-92             static-lattice check / dynamic lattice map   (step 2)
//   helpers.pyx:95-114            per mobile atom: shift statics, wrap         (step 3a)
//   helpers.pyx:134-212           fill_landmark_vec: logistic-cutoff product   (step 3b,3c)
//   helpers.pyx:116-122           all-zero landmark vector check
//   cluster/mcl.py:53             seen_ntimes                                  (MODE_STATS/STAGE)
//   cluster/mcl.py:54             Gram (staged for the tcgen05 SYRK, or sparse outer products)
//   DotProdClassifier.pyx:166-189 predict: |centre . x| argmax + threshold     (MODE_ASSIGN)
//   cluster/mcl.py:81-83          best-matching landmark vector per cluster    (MODE_ASSIGN, best)
//   cluster/mcl.py:118-122        representative landmark vector sums          (MODE_ASSIGN, rep)
//
// Work decomposition: one CTA per frame (grid-stride over frames), one warp per mobile atom.
// Squared distances are computed in IEEE double exactly as the reference does (sitb_common.cuh);
// the cut-off test "distance/site_vert_dist > 1.807" is applied as "d^2 > Q" with Q the exact
// double boundary (sitb_tables.cu), pre-screened on float(d^2) vs float(Q) (rounding is monotone,
// so only exact float ties need the double compare).  Only surviving landmarks (~1-2 %) evaluate
// the logistic, in FP32 with an FP64-prepared argument.
#include "sitb_fill.cuh"
#include <math_constants.h>

namespace sitb {

struct SmemLayout {
    int Spad;
    size_t off_ss, off_sm, off_wq64, off_tq, off_wqf, off_wev, off_wepr, off_hist, off_tv, off_wek,
        off_wec, off_lmap, total;
};

__host__ __device__ inline SmemLayout make_layout(int S, int M, int L, int V, int Lpad, int warps, int mode,
                                                  int n_clusters) {
    SmemLayout l;
    l.Spad = (S + 3) & ~3;
    size_t o = 0;
    l.off_ss = o;   o += sizeof(double) * 3 * (size_t)S;
    l.off_sm = o;   o += sizeof(double) * 3 * (size_t)M;
    l.off_wq64 = o; o += sizeof(double) * (size_t)warps * l.Spad;
    l.off_tq = o;   o += sizeof(float) * (size_t)V * Lpad;
    l.off_wqf = o;  o += sizeof(float) * (size_t)warps * l.Spad;
    l.off_wev = o;  o += sizeof(float) * (size_t)warps * ENTRY_CAP;
    l.off_wepr = o; o += sizeof(float) * (size_t)warps * ENTRY_CAP;
    l.off_hist = o;
    if (mode == MODE_STATS || mode == MODE_STAGE) o += sizeof(unsigned) * (size_t)L;
    if (mode == MODE_ASSIGN) o += sizeof(unsigned) * (size_t)(n_clusters > 0 ? n_clusters : 1);
    l.off_tv = o;   o += sizeof(uint16_t) * (size_t)V * Lpad;
    l.off_wek = o;  o += sizeof(uint16_t) * (size_t)warps * ENTRY_CAP;
    l.off_wec = o;  o += sizeof(int16_t) * (size_t)warps * ENTRY_CAP;
    l.off_lmap = o; o += sizeof(unsigned) * (size_t)l.Spad * 2;   // lattice map + seen counts
    l.total = (o + 15) & ~(size_t)15;
    return l;
}

__device__ __forceinline__ unsigned long long pack_key(float v, unsigned long long row) {
    return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)row);
}

__device__ __forceinline__ void atomic_max_checked(unsigned long long* addr, unsigned long long key) {
    // monotone non-decreasing cell: a stale read can only under-estimate, so the check is safe
    if (key > *((volatile unsigned long long*)addr)) atomicMax(addr, key);
}

// logistic cut-off 1/(1+exp(steep*(d/svd-mid))) from the squared distance (helpers.pyx:197,205)
__device__ __forceinline__ float cutoff_factor(double q, double acoef, double bcoef) {
    // d = sqrt(q): float estimate + one Newton step with the residual taken in double
    const float qf = __double2float_rn(q);
    const float s0 = __fsqrt_rn(qf);
    const double s0d = (double)s0;
    const double res = fma(-s0d, s0d, q);
    const double d = (s0 > 0.f) ? s0d + (double)(__double2float_rn(res) * __frcp_rn(s0 + s0)) : 0.0;
    // exponent in base 2:  x2 = steep*log2e*(d/svd - mid)
    double x2 = fma(d, acoef, -bcoef);
    x2 = fmax(x2, -100.0);
    const int n = __double2int_rn(x2);
    const float fr = __double2float_rn(x2 - (double)n);
    float e = exp2f(fr);
    e = __int_as_float(__float_as_int(e) + (n << 23));   // e * 2^n, result stays normal (n >= -100)
    return __frcp_rn(1.0f + e);
}

// ci^(1/n_verts)  (helpers.pyx:212)
__device__ __forceinline__ float nth_root(float prod, int nv) {
    switch (nv) {
        case 1: return prod;
        case 2: return __fsqrt_rn(prod);
        case 3: return cbrtf(prod);
        case 4: return __fsqrt_rn(__fsqrt_rn(prod));
        default: return powf(prod, 1.0f / (float)nv);
    }
}

template <bool DIAG, int MODE>
__global__ void __launch_bounds__(256) k_fill(const __grid_constant__ FillParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int S = p.S, M = p.M, L = p.L, V = p.V, Lpad = p.Lpad;
    const SmemLayout lay = make_layout(S, M, L, V, Lpad, nwarps, MODE, p.n_clusters);
    double* ss = (double*)(smem_raw + lay.off_ss);
    double* sm = (double*)(smem_raw + lay.off_sm);
    double* q64w = (double*)(smem_raw + lay.off_wq64) + (size_t)warp * lay.Spad;
    float* tq = (float*)(smem_raw + lay.off_tq);
    float* qfw = (float*)(smem_raw + lay.off_wqf) + (size_t)warp * lay.Spad;
    float* ev = (float*)(smem_raw + lay.off_wev) + (size_t)warp * ENTRY_CAP;
    float* epr = (float*)(smem_raw + lay.off_wepr) + (size_t)warp * ENTRY_CAP;
    unsigned* hist = (unsigned*)(smem_raw + lay.off_hist);
    uint16_t* tv = (uint16_t*)(smem_raw + lay.off_tv);
    uint16_t* ek = (uint16_t*)(smem_raw + lay.off_wek) + (size_t)warp * ENTRY_CAP;
    int16_t* ec = (int16_t*)(smem_raw + lay.off_wec) + (size_t)warp * ENTRY_CAP;
    unsigned* lmap = (unsigned*)(smem_raw + lay.off_lmap);
    unsigned* seen_cnt = lmap + lay.Spad;

    // ---- stage the landmark tables once per CTA ------------------------------------------
    for (int i = threadIdx.x; i < V * Lpad; i += blockDim.x) {
        tq[i] = p.qf[i];
        tv[i] = p.verts[i];
    }
    const bool use_hist = (MODE == MODE_STATS || MODE == MODE_STAGE) ||
                          (MODE == MODE_ASSIGN && p.counts != nullptr);
    const int hist_n = (MODE == MODE_ASSIGN) ? p.n_clusters : L;
    if (use_hist)
        for (int i = threadIdx.x; i < hist_n; i += blockDim.x) hist[i] = 0u;
    __syncthreads();

    unsigned long long loc_zero = 0, loc_nnz = 0, loc_tie = 0, loc_over = 0, loc_dup = 0;
    const Cell& cell = p.cell;

    for (long long wi = blockIdx.x; wi < p.n_work; wi += gridDim.x) {
        const long long f = p.frame_list ? p.frame_list[wi] : wi;
        const long long gframe = p.frame0 + f;
        const double* __restrict__ fr = p.frames + (size_t)f * (size_t)p.A * 3;

        // ---- 1. wrap this frame's static and mobile atoms (LandmarkAnalysis.py:182-189) --
        for (int t = threadIdx.x; t < S + M; t += blockDim.x) {
            const int a = (t < S) ? p.static_idx[t] : p.mobile_idx[t - S];
            double x = fr[3 * a + 0], y = fr[3 * a + 1], z = fr[3 * a + 2];
            wrap_point<DIAG, false>(cell, x, y, z);
            double* dst = (t < S) ? (ss + 3 * t) : (sm + 3 * (t - S));
            dst[0] = x; dst[1] = y; dst[2] = z;
        }
        if (p.dynamic)
            for (int t = threadIdx.x; t < S; t += blockDim.x) seen_cnt[t] = 0u;
        __syncthreads();

        // ---- 2. static lattice (helpers.pyx:55-92) ---------------------------------------
        if (!p.dynamic) {
            for (int s = threadIdx.x; s < S; s += blockDim.x) {
                const double ox = __dsub_rn(cell.cen[0], p.ideal[3 * s + 0]);
                const double oy = __dsub_rn(cell.cen[1], p.ideal[3 * s + 1]);
                const double oz = __dsub_rn(cell.cen[2], p.ideal[3 * s + 2]);
                const double q = shifted_dist2<DIAG, false>(cell, ss[3 * s], ss[3 * s + 1], ss[3 * s + 2], ox, oy, oz);
                if (__dsqrt_rn(q) > p.static_thr)
                    atomicMin(p.errkey, make_error_key(gframe, PHASE_STATIC_MOVED, (unsigned)s));
            }
        } else {
            for (int li = warp; li < S; li += nwarps) {
                const double ox = __dsub_rn(cell.cen[0], p.ideal[3 * li + 0]);
                const double oy = __dsub_rn(cell.cen[1], p.ideal[3 * li + 1]);
                const double oz = __dsub_rn(cell.cen[2], p.ideal[3 * li + 2]);
                double bd = CUDART_INF;
                int bj = 0x7FFFFFFF;
                for (int j = lane; j < S; j += 32) {
                    const double d = __dsqrt_rn(shifted_dist2<DIAG, false>(cell, ss[3 * j], ss[3 * j + 1], ss[3 * j + 2], ox, oy, oz));
                    if (d < bd) { bd = d; bj = j; }           // first minimum within the lane
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {            // argmin, ties -> lower index (np.argmin)
                    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                    if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
                }
                if (lane == 0) {
                    atomicAdd(&seen_cnt[bj], 1u);
                    lmap[li] = (unsigned)bj;
                    if (bd > p.static_thr)
                        atomicMin(p.errkey, make_error_key(gframe, PHASE_STATIC_MOVED, (unsigned)li));
                }
            }
            __syncthreads();
            for (int t = threadIdx.x; t < S; t += blockDim.x) {
                const unsigned c = seen_cnt[t];
                if (c == 0u && !p.relaxed)
                    atomicMin(p.errkey, make_error_key(gframe, PHASE_STATIC_UNASSIGNED, 0u));
                if (c > 1u) loc_dup += c - 1u;
            }
        }
        __syncthreads();

        // ---- 3. one warp per mobile atom --------------------------------------------------
        for (int j = warp; j < M; j += nwarps) {
            const long long row_local = wi * M + j;                 // row in this launch's outputs
            const unsigned long long row_global = (unsigned long long)(gframe * M + j);
            // 3a. squared distances static -> mobile (helpers.pyx:99-103, :174-178)
            const double ox = __dsub_rn(cell.cen[0], sm[3 * j + 0]);
            const double oy = __dsub_rn(cell.cen[1], sm[3 * j + 1]);
            const double oz = __dsub_rn(cell.cen[2], sm[3 * j + 2]);
            for (int s = lane; s < S; s += 32) {
                const int src = p.dynamic ? (int)lmap[s] : s;
                const double q = shifted_dist2<DIAG, true>(cell, ss[3 * src], ss[3 * src + 1], ss[3 * src + 2], ox, oy, oz);
                q64w[s] = q;
                qfw[s] = __double2float_rn(q);
            }
            __syncwarp();

            // 3b. landmark walk (helpers.pyx:186-212)
            int nent = 0;
            for (int k0 = 0; k0 < L; k0 += 32) {
                const int k = k0 + lane;
                bool alive = k < L;
                int nv = 0;
                if (alive) {
                    for (int h = 0; h < V; ++h) {
                        const unsigned v = tv[h * Lpad + k];
                        if (v == VERT_END) break;
                        ++nv;
                        const float a = qfw[v];
                        const float b = tq[h * Lpad + k];
                        if (a > b) { alive = false; break; }
                        if (a == b) {                          // float tie: decide in double
                            ++loc_tie;
                            if (q64w[v] > p.q64[h * Lpad + k]) { alive = false; break; }
                        }
                    }
                    if (nv == 0) alive = false;
                }
                float val = 0.f;
                if (alive) {
                    float prod = 1.f;
                    for (int h = 0; h < nv; ++h) {
                        const unsigned v = tv[h * Lpad + k];
                        prod *= cutoff_factor(q64w[v], p.acoef[h * Lpad + k], p.bcoef);
                    }
                    val = nth_root(prod, nv);
                }
                if (MODE == MODE_DENSE) {
                    if (k < L) {
                        if (p.dense_f64) ((double*)p.dense_out)[(size_t)row_local * L + k] = (double)val;
                        else ((float*)p.dense_out)[(size_t)row_local * L + k] = val;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, alive);
                if (MODE != MODE_DENSE) {
                    if (alive) {
                        const int pos = nent + __popc(m & lanemask_lt());
                        if (pos < ENTRY_CAP) { ek[pos] = (uint16_t)k; ev[pos] = val; }
                    }
                }
                nent += __popc(m);
            }
            if (lane == 0) {
                loc_nnz += (unsigned long long)nent;
                if (nent == 0) {
                    ++loc_zero;
                    if (p.errkey)
                        atomicMin(p.errkey + 1, make_error_key(gframe, PHASE_ZERO_LVEC, (unsigned)j));
                }
                if (nent > ENTRY_CAP) ++loc_over;
            }
            if (nent > ENTRY_CAP) nent = ENTRY_CAP;
            __syncwarp();

            // 3c. sinks
            if (MODE == MODE_STATS || MODE == MODE_STAGE) {
                for (int e = lane; e < nent; e += 32) atomicAdd(&hist[ek[e]], 1u);
            }
            if (MODE == MODE_STATS) {
                // upper triangle of the outer product; entries are sorted by landmark index
                for (int a = 0; a < nent; ++a) {
                    const unsigned ka = ek[a];
                    const double va = (double)ev[a];
                    for (int b = a + lane; b < nent; b += 32)
                        atomicAdd(&p.gram[(size_t)ka * L + ek[b]], va * (double)ev[b]);
                }
            }
            if (MODE == MODE_STAGE) {
                for (int e = lane; e < nent; e += 32) {
                    const float v = ev[e];
                    const __half hi = __float2half_rn(v);
                    const __half lo = __float2half_rn(v - __half2float(hi));
                    const size_t o = (size_t)ek[e] * (size_t)p.stage_ld + (size_t)row_local;
                    p.stage_hi[o] = hi;
                    p.stage_lo[o] = lo;
                }
            }
            if (MODE == MODE_ASSIGN) {
                // centres have disjoint supports (cluster/mcl.py:80): dot = sum over the row's
                // non-zeros of weight[landmark], grouped by cluster[landmark]
                for (int e = lane; e < nent; e += 32) {
                    const int k = ek[e];
                    const int c = p.cid[k];
                    ec[e] = (int16_t)c;
                    epr[e] = (c >= 0) ? ev[e] * p.cw[k] : 0.f;
                }
                __syncwarp();
                float bestc = 0.f;      // untouched clusters have |dot| = 0; np.argmax -> index 0
                int bestid = 0;
                for (int base = 0; base < nent; base += 32) {
                    const int e = base + lane;
                    const int my = (e < nent) ? (int)ec[e] : -1;
                    float tot = 0.f;
                    bool first = true;
                    if (my >= 0) {
                        for (int e2 = 0; e2 < nent; ++e2) {
                            if ((int)ec[e2] == my) {
                                tot += epr[e2];
                                if (e2 < e) first = false;
                            }
                        }
                        const float conf = fabsf(tot);
                        if (conf > bestc || (conf == bestc && my < bestid)) { bestc = conf; bestid = my; }
                        if (p.best && first) atomic_max_checked(p.best + my, pack_key(conf, row_global));
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float oc = __shfl_xor_sync(0xffffffffu, bestc, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bestid, o);
                    if (oc > bestc || (oc == bestc && oi < bestid)) { bestc = oc; bestid = oi; }
                }
                long long label = bestid;
                float conf = bestc;
                if (nent == 0 || !(conf >= p.assign_thr)) { label = -1; conf = 0.f; }   // DotProdClassifier.pyx:168-172,184-186
                if (lane == 0) {
                    if (p.labels) p.labels[row_local] = label;
                    if (p.confs) p.confs[row_local] = (double)conf;
                    if (label >= 0) {
                        if (p.counts) atomicAdd(&hist[label], 1u);
                        if (p.rep_w) atomicAdd(&p.rep_w[label], (double)conf);
                        if (p.site_best) atomic_max_checked(p.site_best + label, pack_key(conf, row_global));
                    }
                }
                if (p.rep && label >= 0) {
                    for (int e = lane; e < nent; e += 32)
                        atomicAdd(&p.rep[(size_t)label * L + ek[e]], (double)conf * (double)ev[e]);
                }
            }
            __syncwarp();
        }
        __syncthreads();
    }

    // ---- flush per-CTA accumulators -----------------------------------------------------------
    if (use_hist) {
        __syncthreads();
        unsigned long long* dst = (MODE == MODE_ASSIGN) ? p.counts : p.seen;
        if (dst)
            for (int i = threadIdx.x; i < hist_n; i += blockDim.x)
                if (hist[i]) atomicAdd(&dst[i], (unsigned long long)hist[i]);
    }
    if (p.counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            loc_tie += __shfl_xor_sync(0xffffffffu, loc_tie, o);
            loc_dup += __shfl_xor_sync(0xffffffffu, loc_dup, o);
        }
        if (lane == 0) {
            if (loc_zero) atomicAdd(&p.counters[CNT_ZERO_ROWS], loc_zero);
            if (loc_nnz) atomicAdd(&p.counters[CNT_NNZ], loc_nnz);
            if (loc_over) atomicAdd(&p.counters[CNT_LIST_OVERFLOW], loc_over);
            if (loc_tie) atomicAdd(&p.counters[CNT_TIE_EXACT], loc_tie);
            if (loc_dup) atomicAdd(&p.counters[CNT_DUP_NEAREST], loc_dup);
        }
    }
}

template <bool DIAG, int MODE>
static cudaError_t launch_one(const FillParams& p, int n_sms, cudaStream_t stream) {
    // as many warps per CTA as fit the shared-memory budget, grid = a multiple of the SM count
    int warps = 8;
    size_t bytes = 0;
    const int ncl = p.n_clusters;
    for (; warps >= 1; warps >>= 1) {
        bytes = make_layout(p.S, p.M, p.L, p.V, p.Lpad, warps, MODE, ncl).total;
        if (bytes <= 200 * 1024) break;
    }
    if (warps < 1) return cudaErrorInvalidConfiguration;
    auto kern = k_fill<DIAG, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)n_sms * per_sm;
    if (grid > p.n_work) grid = p.n_work;
    if (grid < 1) return cudaSuccess;
    kern<<<(unsigned)grid, warps * 32, bytes, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fill(const FillParams& p, int mode, int n_sms, cudaStream_t stream) {
    const bool diag = p.cell.diag != 0;
#define SITB_DISPATCH(MODE_)                                                     \
    case MODE_:                                                                  \
        return diag ? launch_one<true, MODE_>(p, n_sms, stream) : launch_one<false, MODE_>(p, n_sms, stream);
    switch (mode) {
        SITB_DISPATCH(MODE_DENSE)
        SITB_DISPATCH(MODE_STATS)
        SITB_DISPATCH(MODE_STAGE)
        SITB_DISPATCH(MODE_ASSIGN)
        default: return cudaErrorInvalidValue;
    }
#undef SITB_DISPATCH
}

}  // namespace sitb
