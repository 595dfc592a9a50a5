// sitator_b200 -- K1: fused wrap + static-lattice check + landmark-vector fill (+ assign).
//
// Replaces, per frame and fused in one kernel (reference file:line):
//   LandmarkAnalysis.py:182-189   wrap a copy of the frame into the cell       (step 1 below)
//   helpers.pyx:55-92             static-lattice check / dynamic lattice map   (step 2)
//   helpers.pyx:95-114            per mobile atom: shift statics, wrap         (step 3a)
//   helpers.pyx:134-212           fill_landmark_vec: logistic-cutoff product   (steps 3b-3d)
//   helpers.pyx:116-122           all-zero landmark vector check
//   cluster/mcl.py:53             seen_ntimes                                  (MODE_STATS/STAGE)
//   cluster/mcl.py:54             Gram (staged for the tcgen05 SYRK, or sparse outer products)
//   DotProdClassifier.pyx:166-189 predict: |centre . x| argmax + threshold     (MODE_ASSIGN)
//   cluster/mcl.py:81-83          best-matching landmark vector per cluster    (MODE_ASSIGN, best)
//   cluster/mcl.py:118-122        representative landmark vector sums          (MODE_ASSIGN, rep)
//
// Work decomposition: a CTA takes a small batch of frames (grid-stride); its warps claim the
// batch's mobile atoms one at a time from a shared counter.  Per mobile atom the landmark walk is
// split into phases so that no lane sits in a long serial dependency chain and the rare expensive
// work is done on compacted lists:
//   3a  squared distance to every static atom in FP32 (orthorhombic cells) -- a *screen* only
//   3b  screen the first vertex of every landmark against a float bound that includes the FP32
//       error margin (sitb_tables.cu): no landmark of the true support is ever dropped.  Passes
//       are kept as one bit per lane and sub-iteration and compacted once per 1024 landmarks.
//       Landmarks are numbered so that neighbours share their first vertex (few bank conflicts).
//   3c  screen the remaining vertices of the ~13 % that passed (vector loads of the vertex block)
//   3d  for the ~2 % left: recompute the squared distances in IEEE double exactly as the
//       reference does (sitb_common.cuh), apply the exact cut-off  d^2 > Q  (bit-identical
//       support), evaluate the logistic product in FP32 (log2 domain) from an FP64-prepared
//       argument
// For triclinic cells 3a is done in exact double (the reference's shift-and-wrap is not a
// continuous function of position there, so a float screen can not be made conservative).
#include "sitb_assign.cuh"
#include <math_constants.h>

namespace sitb {

// inner offsets of a warp's scratch block
static constexpr size_t WARP_EV = 0;                                          // double[ENTRY_CAP]
static constexpr size_t WARP_CAND = WARP_EV + sizeof(double) * ENTRY_CAP;     // uint16[CAND_CAP]
static constexpr size_t WARP_EK = WARP_CAND + sizeof(uint16_t) * CAND_CAP;    // uint16[ENTRY_CAP]
static constexpr size_t WARP_POOL = WARP_EK + sizeof(uint16_t) * ENTRY_CAP;   // u64 offset, u32 entries left: the warp's
                                                                              // current slice of the compressed-row pool
static constexpr size_t WARP_QFW = WARP_POOL + 16;                            // float[qstride]
static constexpr unsigned POOL_SLICE = 256;   // entries a warp reserves at a time (>= ENTRY_CAP): one global atomic per
                                              // ~11 rows instead of one per row (5.6e6 atomics on one address per pass)

struct SmemLayout {
    int Spad, Mpad, qstride;
    // per-warp scratch is one block per warp (off_warp + warp * warp_bytes) with compile-time inner offsets, so one
    // base address serves the four arrays: values, candidate list, survivor / entry list, screen distances
    size_t off_warp, warp_bytes;
    size_t off_ss, off_sm, off_ba, off_b0, off_fs, off_fm, off_hist, off_lmap, off_seen,
        off_cw, off_va, off_v0, off_cid, off_task, off_flag, off_ca, off_cb, total;
};

__host__ __device__ inline SmemLayout make_layout(int S, int M, int L, int Lpad, int NB, int warps, int fb, int mode,
                                                  int n_clusters, int dynamic) {
    SmemLayout l;
    l.Spad = (S + 3) & ~3;
    l.Mpad = (M + 3) & ~3;
    l.qstride = l.Spad + 4;     // per-warp screen distances + the dummy vertex slot [S]
    size_t o = 0;
    l.off_ss = o;    o += sizeof(double) * 3 * (size_t)S * fb;
    l.warp_bytes = (WARP_QFW + sizeof(float) * (size_t)l.qstride + 15) & ~(size_t)15;
    l.off_warp = o;  o += l.warp_bytes * (size_t)warps;
    l.off_cw = o;    o += (mode == MODE_ASSIGN) ? sizeof(double) * (size_t)Lpad : 0;
    l.off_sm = o;    o += sizeof(double) * 3 * (size_t)M * fb;
    o = (o + 15) & ~(size_t)15;
    l.off_ba = o;    o += sizeof(float4) * (size_t)NB * Lpad;
    l.off_cb = o;    o += sizeof(float4) * (size_t)(Lpad / 32);
    l.off_b0 = o;    o += sizeof(float) * (size_t)Lpad;
    l.off_fs = o;    o += sizeof(float) * 3 * (size_t)l.Spad * fb;
    l.off_fm = o;    o += sizeof(float) * 3 * (size_t)l.Mpad * fb;
    l.off_hist = o;
    if (mode == MODE_STATS || mode == MODE_STAGE) o += sizeof(unsigned) * (size_t)L;
    if (mode == MODE_ASSIGN) o += sizeof(unsigned) * (size_t)(n_clusters > 0 ? n_clusters : 1);
    l.off_lmap = o;  o += dynamic ? sizeof(unsigned) * (size_t)l.Spad * fb : 0;
    l.off_seen = o;  o += dynamic ? sizeof(unsigned) * (size_t)l.Spad * fb : 0;
    o = (o + 7) & ~(size_t)7;
    l.off_va = o;    o += sizeof(ushort4) * (size_t)NB * Lpad;
    l.off_ca = o;    o += sizeof(ushort4) * (size_t)(Lpad / 32);
    l.off_v0 = o;    o += sizeof(uint16_t) * (size_t)Lpad;
    l.off_cid = o;   o += (mode == MODE_ASSIGN) ? sizeof(int16_t) * (size_t)Lpad : 0;
    o = (o + 3) & ~(size_t)3;
    l.off_task = o;  o += sizeof(int);
    l.off_flag = o;  o += sizeof(int) * (size_t)fb;
    l.total = (o + 15) & ~(size_t)15;
    return l;
}

// ---- float64 kernels of step 3d ------------------------------------------------------------------
// The landmark component is (prod_h (1 + exp(steep*(d_h/svd_h - mid))))^(-1/n) (helpers.pyx:197-212).  It is
// evaluated in double: the reference's own argmax-over-rows decisions (cluster/mcl.py:83) separate rows whose
// components differ by ~1e-8 relative, and the parity tolerance is 1e-12.  The library exp/sqrt/cbrt carry
// range checks, slow paths and 64-bit immediates (two moves per polynomial coefficient); the arguments here
// have a known range, so the three functions are written out with their constants in the constant bank.
// Each agrees with the correctly rounded result to ~1 ulp.
__constant__ double c_exp[14] = {
    1.4426950408889634074,        // [0] log2(e)
    6755399441055744.0,           // [1] 1.5 * 2^52: adding it leaves rint(t) in the low word
    -6.93147180369123816490e-01,  // [2] -ln2, high part (32 significant bits)
    -1.90821492927058770002e-10,  // [3] -ln2, low part
    // exp(r) = 1 + r + r^2 Q(r) on |r| <= ln2/2: Q = Chebyshev interpolant of degree 9 (max rel. error 1.6e-17)
    2.510038549551032e-08, 2.7620088445409746e-07, 2.7557268459997064e-06, 2.4801521295954376e-05,
    0.00019841269863053618, 0.0013888888917213717, 0.0083333333333300615, 0.041666666666624129,
    0.16666666666666669, 0.50000000000000011};

// MUFU seeds without the library's denormal handling (the arguments are normal floats here)
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// exp(x) for |x| < 700 (no overflow / underflow handling)
__device__ __forceinline__ double exp_bounded(double x) {
    const double t = fma(x, c_exp[0], c_exp[1]);
    const int n = __double2loint(t);
    const double nf = t - c_exp[1];
    double r = fma(nf, c_exp[2], x);
    r = fma(nf, c_exp[3], r);
    double q = c_exp[4];
#pragma unroll
    for (int i = 5; i < 14; ++i) q = fma(q, r, c_exp[i]);
    q = fma(q * r, r, r);                 // r + r^2 Q(r)
    q = q + 1.0;
    return __hiloint2double(__double2hiint(q) + n * 1048576, __double2loint(q));
}

// sqrt(q) for 1e-30 < q < 1e30: float reciprocal-square-root seed, two coupled Newton steps, one correction
__device__ __forceinline__ double sqrt_bounded(double q) {
    const double y = (double)rsqrt_approx((float)q);
    double g = q * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    return fma(fma(-g, g, q), h, g);
}

// P^(-1/n) for 1 <= P < 1e300, 1 <= n <= 16: float seed from the exponent and the mantissa's log2 (P itself can exceed
// the float range with more than nine vertices), two Newton steps on y^-n = P (same code for every n)
__device__ __forceinline__ double inv_root(double P, int n) {
    const float rn = __fdividef(1.0f, (float)n);     // approximate is enough: it only scales the Newton step
    const int hi = __double2hiint(P);
    const float l2 = (float)((hi >> 20) - 1023) +
                     lg2_approx((float)__hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(P)));
    double y = (double)ex2_approx(-l2 * rn);
    const double rnd = (double)rn;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double y2 = y * y, y4 = y2 * y2, y8 = y4 * y4;
        double t = (n & 1) ? y : 1.0;
        if (n & 2) t *= y2;
        if (n & 4) t *= y4;
        if (n & 8) t *= y8;
        if (n & 16) t *= y8 * y8;
        y = fma(y * fma(-P, t, 1.0), rnd, y);
    }
    return y;
}

// u - round(u) for |u| < 2^22, two adds
__device__ __forceinline__ float centre_frac(float u) {
    const float magic = 12582912.0f;   // 1.5 * 2^23
    return __fsub_rn(u, __fsub_rn(__fadd_rn(u, magic), magic));
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

template <bool DIAG, int MODE>
#ifndef SITB_K1_WARPS
#define SITB_K1_WARPS 32
#endif
__global__ void __launch_bounds__(DIAG ? SITB_K1_WARPS * 32 : 512) k_fill(const __grid_constant__ FillParams p, const int FB,
                                                            const __grid_constant__ SmemLayout lay) {
    // (the layout is computed by the host and read from the constant bank: under the 64-register cap
    // the compiler otherwise rebuilds these offsets inside the hot loops)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int S = p.S, M = p.M, L = p.L, Lpad = p.Lpad, NB = p.NB;
    const int Spad = lay.Spad, Mpad = lay.Mpad;
    double* ss = (double*)(smem_raw + lay.off_ss);          // [FB][S][3] wrapped statics (double)
    double* sm = (double*)(smem_raw + lay.off_sm);          // [FB][M][3] wrapped mobiles (double)
    float4* tba = (float4*)(smem_raw + lay.off_ba);         // [NB][Lpad] screen bounds, 4 vertices
    float* tb0 = (float*)(smem_raw + lay.off_b0);           // [Lpad] screen bound of vertex 0
    float* fs = (float*)(smem_raw + lay.off_fs);            // [FB][3][Spad] fractional statics (float, SoA)
    float* fm = (float*)(smem_raw + lay.off_fm);            // [FB][3][Mpad]
    unsigned char* wscratch = smem_raw + lay.off_warp + (size_t)warp * lay.warp_bytes;
    float* qfw = (float*)(wscratch + WARP_QFW);
    double* ev = (double*)(wscratch + WARP_EV);
    unsigned* hist = (unsigned*)(smem_raw + lay.off_hist);
    unsigned* lmap_all = (unsigned*)(smem_raw + lay.off_lmap);
    unsigned* seen_all = (unsigned*)(smem_raw + lay.off_seen);
    double* tcw = (double*)(smem_raw + lay.off_cw);         // [Lpad] centre weight (MODE_ASSIGN)
    ushort4* tva = (ushort4*)(smem_raw + lay.off_va);       // [NB][Lpad] vertex ids, 4 per block
    uint16_t* tv0 = (uint16_t*)(smem_raw + lay.off_v0);     // [Lpad] vertex 0
    int16_t* tcid = (int16_t*)(smem_raw + lay.off_cid);     // [Lpad] cluster of landmark (MODE_ASSIGN)
    uint16_t* cand = (uint16_t*)(wscratch + WARP_CAND);
    uint16_t* ek = (uint16_t*)(wscratch + WARP_EK);
    unsigned long long* pool_off = (unsigned long long*)(wscratch + WARP_POOL);
    unsigned* pool_left = (unsigned*)(wscratch + WARP_POOL + 8);
    int* task_counter = (int*)(smem_raw + lay.off_task);
    int* glevel = (int*)(smem_raw + lay.off_flag);         // [FB] grid level of the frame (n_levels: walk all landmarks)
    ushort4* tca = (ushort4*)(smem_raw + lay.off_ca);       // [Lpad/32] chunk skip table: atoms
    float4* tcb = (float4*)(smem_raw + lay.off_cb);         // [Lpad/32] chunk skip table: bounds

    // ---- stage the landmark tables once per CTA ------------------------------------------
    for (int i = threadIdx.x; i < (Lpad >> 5); i += blockDim.x) {
        tca[i] = p.tab.chunk_atoms[i];
        tcb[i] = p.tab.chunk_bound[i];
    }
    for (int i = threadIdx.x; i < NB * Lpad; i += blockDim.x) {
        tba[i] = p.tab.ba[i];
        tva[i] = p.tab.va[i];
    }
    for (int i = threadIdx.x; i < Lpad; i += blockDim.x) {
        tb0[i] = p.tab.b0[i];
        tv0[i] = p.tab.v0[i];
        if (MODE == MODE_ASSIGN) {
            tcid[i] = (i < L) ? (int16_t)p.cid[i] : (int16_t)-1;
            tcw[i] = (i < L) ? p.cw[i] : 0.0;
        }
    }
    const bool use_hist = (MODE == MODE_STATS || MODE == MODE_STAGE) ||
                          (MODE == MODE_ASSIGN && p.counts != nullptr);
    const int hist_n = (MODE == MODE_ASSIGN) ? p.n_clusters : L;
    if (use_hist)
        for (int i = threadIdx.x; i < hist_n; i += blockDim.x) hist[i] = 0u;
    if (lane == 0) { qfw[S] = 0.f; *pool_left = 0u; *pool_off = 0ull; }   // dummy vertex: passes every screen
    __syncthreads();

    unsigned long long loc_zero = 0, loc_nnz = 0, loc_rej = 0, loc_over = 0, loc_dup = 0, loc_full = 0;
    const int n_levels = DIAG ? p.n_grid_levels : 0;
    unsigned long long loc_loose = 0;
    const Cell& cell = p.cell;
    const float Lx = (float)cell.c[0], Ly = (float)cell.c[4], Lz = (float)cell.c[8];
    const int W = 4 * NB;

    // (second tier of the two-tier assign pass: the list length is only known on the device)
    const long long n_work = p.n_work_dev ? (long long)*p.n_work_dev : p.n_work;
    for (long long w0 = (long long)blockIdx.x * FB; w0 < n_work; w0 += (long long)gridDim.x * FB) {
        const int nb = (int)((n_work - w0 < FB) ? (n_work - w0) : FB);

        // ---- 1. wrap the batch's static and mobile atoms (LandmarkAnalysis.py:182-189) ----
        for (int t = threadIdx.x; t < nb * (S + M); t += blockDim.x) {
            const int b = t / (S + M), r = t - b * (S + M);
            const long long f = p.frame_list ? p.frame_list[w0 + b] : (w0 + b);
            const double* __restrict__ fr = p.frames + (size_t)f * (size_t)p.A * 3;
            const int a = (r < S) ? p.static_idx[r] : p.mobile_idx[r - S];
            double x = fr[3 * a + 0], y = fr[3 * a + 1], z = fr[3 * a + 2];
            if (DIAG) {
                double f0, f1, f2;
                wrap_diag_frac(cell, x, y, z, f0, f1, f2);
                if (r < S) {
                    float* d = fs + (size_t)b * 3 * Spad;
                    d[r] = (float)f0; d[Spad + r] = (float)f1; d[2 * Spad + r] = (float)f2;
                } else {
                    float* d = fm + (size_t)b * 3 * Mpad;
                    d[r - S] = (float)f0; d[Mpad + r - S] = (float)f1; d[2 * Mpad + r - S] = (float)f2;
                }
            } else {
                wrap_point<false, false>(cell, x, y, z);
            }
            double* dst = (r < S) ? (ss + ((size_t)b * S + r) * 3) : (sm + ((size_t)b * M + (r - S)) * 3);
            dst[0] = x; dst[1] = y; dst[2] = z;
        }
        if (p.dynamic)
            for (int t = threadIdx.x; t < nb * Spad; t += blockDim.x) seen_all[t] = 0u;
        if (threadIdx.x == 0) *task_counter = 0;
        if (threadIdx.x < FB) glevel[threadIdx.x] = 0;
        __syncthreads();

        // ---- 2. static lattice (helpers.pyx:55-92) ---------------------------------------
        if (!p.dynamic) {
            for (int t = threadIdx.x; t < nb * S; t += blockDim.x) {
                const int b = t / S, s = t - b * S;
                const long long gframe = p.frame0 + (p.frame_list ? p.frame_list[w0 + b] : (w0 + b));
                const double* pt = ss + ((size_t)b * S + s) * 3;
                const double ox = __dsub_rn(cell.cen[0], p.ideal[3 * s + 0]);
                const double oy = __dsub_rn(cell.cen[1], p.ideal[3 * s + 1]);
                const double oz = __dsub_rn(cell.cen[2], p.ideal[3 * s + 2]);
                const double q = shifted_dist2<DIAG, false>(cell, pt[0], pt[1], pt[2], ox, oy, oz);
                if (__dsqrt_rn(q) > p.static_thr)
                    atomicMin(p.errkey, make_error_key(gframe, PHASE_STATIC_MOVED, (unsigned)s));
                if (n_levels > 0 && q > p.grid[0].margin_sq)
                    atomicMax(&glevel[b], (n_levels > 1 && !(q > p.grid[1].margin_sq)) ? 1 : n_levels);
            }
        } else {
            for (int t = warp; t < nb * S; t += nwarps) {
                const int b = t / S, li = t - b * S;
                const long long gframe = p.frame0 + (p.frame_list ? p.frame_list[w0 + b] : (w0 + b));
                const double* sb = ss + (size_t)b * S * 3;
                const double ox = __dsub_rn(cell.cen[0], p.ideal[3 * li + 0]);
                const double oy = __dsub_rn(cell.cen[1], p.ideal[3 * li + 1]);
                const double oz = __dsub_rn(cell.cen[2], p.ideal[3 * li + 2]);
                double bd = CUDART_INF;
                int bj = 0x7FFFFFFF;
                if (DIAG) {
                    // FP32 screen first: only atoms whose float distance is within the float error of the smallest one can
                    // be the exact arg-min (normally one atom), and only those get the exact double evaluation
                    const float* fsb = fs + (size_t)b * 3 * Spad;
                    double i0 = __dmul_rn(cell.ci[0], p.ideal[3 * li + 0]), i1 = __dmul_rn(cell.ci[4], p.ideal[3 * li + 1]),
                           i2 = __dmul_rn(cell.ci[8], p.ideal[3 * li + 2]);
                    const float fx = (float)(i0 - floor(i0)), fy = (float)(i1 - floor(i1)), fz = (float)(i2 - floor(i2));
                    float q32[8];                                   // up to 256 static atoms take the screened route
                    float qmin = CUDART_INF_F;
                    const int n_it = (S + 31) >> 5;
                    if (n_it <= 8) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int j = lane + 32 * it;
                            q32[it] = CUDART_INF_F;
                            if (it < n_it && j < S) {
                                const float cx = centre_frac(fsb[j] - fx) * Lx;
                                const float cy = centre_frac(fsb[Spad + j] - fy) * Ly;
                                const float cz = centre_frac(fsb[2 * Spad + j] - fz) * Lz;
                                q32[it] = fmaf(cz, cz, fmaf(cy, cy, cx * cx));
                                qmin = fminf(qmin, q32[it]);
                            }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) qmin = fminf(qmin, __shfl_xor_sync(0xffffffffu, qmin, o));
                        // float error of a squared distance q (sitb_tables.cu: screen_bound), twice, for both sides
                        const float lmax = fmaxf(Lx, fmaxf(Ly, Lz));
                        const float dc = 24.0f * 5.9604644775390625e-08f * lmax;
                        const float bound = qmin + 2.0f * (2.0f * sqrtf(3.0f * qmin) * dc + 3.0f * dc * dc + 1e-6f * qmin) + 1e-12f;
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int j = lane + 32 * it;
                            if (it < n_it && j < S && !(q32[it] > bound)) {
                                const double d = __dsqrt_rn(shifted_dist2<DIAG, false>(cell, sb[3 * j], sb[3 * j + 1], sb[3 * j + 2], ox, oy, oz));
                                if (d < bd) { bd = d; bj = j; }   // first minimum within the lane (j ascends)
                            }
                        }
                    } else {
                        for (int j = lane; j < S; j += 32) {
                            const double d = __dsqrt_rn(shifted_dist2<DIAG, false>(cell, sb[3 * j], sb[3 * j + 1], sb[3 * j + 2], ox, oy, oz));
                            if (d < bd) { bd = d; bj = j; }
                        }
                    }
                } else {
                    for (int j = lane; j < S; j += 32) {
                        const double d = __dsqrt_rn(shifted_dist2<DIAG, false>(cell, sb[3 * j], sb[3 * j + 1], sb[3 * j + 2], ox, oy, oz));
                        if (d < bd) { bd = d; bj = j; }           // first minimum within the lane
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {            // argmin, ties -> lower index (np.argmin)
                    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                    if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
                }
                if (lane == 0) {
                    atomicAdd(&seen_all[(size_t)b * Spad + bj], 1u);
                    lmap_all[(size_t)b * Spad + li] = (unsigned)bj;
                    const double bq = __dmul_rn(bd, bd);
                    if (n_levels > 0 && bq > p.grid[0].margin_sq)
                        atomicMax(&glevel[b], (n_levels > 1 && !(bq > p.grid[1].margin_sq)) ? 1 : n_levels);
                    if (bd > p.static_thr)
                        atomicMin(p.errkey, make_error_key(gframe, PHASE_STATIC_MOVED, (unsigned)li));
                }
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nb * S; t += blockDim.x) {
                const int b = t / S, s = t - b * S;
                const unsigned c = seen_all[(size_t)b * Spad + s];
                if (c == 0u && !p.relaxed) {
                    const long long gframe = p.frame0 + (p.frame_list ? p.frame_list[w0 + b] : (w0 + b));
                    atomicMin(p.errkey, make_error_key(gframe, PHASE_STATIC_UNASSIGNED, 0u));
                }
                if (c > 1u) loc_dup += c - 1u;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int b = 0; b < nb; ++b) {
                loc_full += (unsigned long long)(glevel[b] >= n_levels);
                loc_loose += (unsigned long long)(glevel[b] > 0 && glevel[b] < n_levels);
            }

        // ---- 3. one warp per mobile atom, claimed from a per-batch counter ------------------
        for (;;) {
            int jj = 0;
            if (lane == 0) jj = atomicAdd(task_counter, 1);
            jj = __shfl_sync(0xffffffffu, jj, 0);
            if (jj >= nb * M) break;
            // frame of the batch and mobile atom: jj / M by a multiply with ceil(2^32 / M) (exact for jj < 2^17, M < 2^14;
            // a signed 32-bit division is ~35 instructions on the uniform datapath)
            const int b = p.m_magic ? (int)__umulhi((unsigned)jj, p.m_magic) : jj / M;
            const int j = jj - b * M;
            const long long fl = p.frame_list ? p.frame_list[w0 + b] : (w0 + b);
            const long long gframe = p.frame0 + fl;
            // row in this launch's outputs: by list position, = (w0 + b) * M + j, or by frame
            const long long row_local = p.rows_by_frame ? fl * M + j : w0 * M + jj;
            if (p.row_filter && !p.row_filter[row_local]) continue;
            const unsigned long long row_global = (unsigned long long)(gframe * M + j);
            const double* sb = ss + (size_t)b * S * 3;
            const unsigned* lmap = lmap_all + (size_t)b * Spad;
            const double ox = __dsub_rn(cell.cen[0], sm[((size_t)b * M + j) * 3 + 0]);
            const double oy = __dsub_rn(cell.cen[1], sm[((size_t)b * M + j) * 3 + 1]);
            const double oz = __dsub_rn(cell.cen[2], sm[((size_t)b * M + j) * 3 + 2]);

            // 3a. screen distances static -> mobile (helpers.pyx:99-103, :174-178)
            float mx = 0.f, my = 0.f, mz = 0.f;
            int box = 0;
            const int lv = glevel[b];
            if (DIAG) {
                const float* fsb = fs + (size_t)b * 3 * Spad;
                const float* fmb = fm + (size_t)b * 3 * Mpad;
                mx = fmb[j]; my = fmb[Mpad + j]; mz = fmb[2 * Mpad + j];
                if (lv < n_levels) {
                    // the grid box of the mobile atom: only the static-lattice sites its candidate landmarks use
                    int ix = (int)(mx * (float)p.gx), iy = (int)(my * (float)p.gy), iz = (int)(mz * (float)p.gz);
                    ix = ix < 0 ? 0 : (ix >= p.gx ? p.gx - 1 : ix);
                    iy = iy < 0 ? 0 : (iy >= p.gy ? p.gy - 1 : iy);
                    iz = iz < 0 ? 0 : (iz >= p.gz ? p.gz - 1 : iz);
                    box = (ix * p.gy + iy) * p.gz + iz;
                    const unsigned sbeg = __ldg(p.grid[lv].sptr + box), send = __ldg(p.grid[lv].sptr + box + 1);
                    for (unsigned i = sbeg + lane; i < send; i += 32) {
                        const int s = (int)__ldg(p.grid[lv].slist + i);
                        const int src = p.dynamic ? (int)lmap[s] : s;
                        const float cx = centre_frac(fsb[src] - mx) * Lx;
                        const float cy = centre_frac(fsb[Spad + src] - my) * Ly;
                        const float cz = centre_frac(fsb[2 * Spad + src] - mz) * Lz;
                        qfw[s] = fmaf(cz, cz, fmaf(cy, cy, cx * cx));
                    }
                } else
                for (int s = lane; s < S; s += 32) {
                    const int src = p.dynamic ? (int)lmap[s] : s;
                    const float cx = centre_frac(fsb[src] - mx) * Lx;
                    const float cy = centre_frac(fsb[Spad + src] - my) * Ly;
                    const float cz = centre_frac(fsb[2 * Spad + src] - mz) * Lz;
                    qfw[s] = fmaf(cz, cz, fmaf(cy, cy, cx * cx));
                }
            } else {
                for (int s = lane; s < S; s += 32) {
                    const int src = p.dynamic ? (int)lmap[s] : s;
                    qfw[s] = __double2float_rn(
                        shifted_dist2<false, true>(cell, sb[3 * src], sb[3 * src + 1], sb[3 * src + 2], ox, oy, oz));
                }
            }
            if (MODE == MODE_DENSE) {     // rows are written as zeros + scattered non-zeros
                if (p.dense_f64) {
                    double* o = (double*)p.dense_out + (size_t)row_local * L;
                    for (int k = lane; k < L; k += 32) o[k] = 0.0;
                } else {
                    float* o = (float*)p.dense_out + (size_t)row_local * L;
                    for (int k = lane; k < L; k += 32) o[k] = 0.f;
                }
            }
            __syncwarp();

            int nsurv = 0, ncand = 0;
            const int n_chunks = Lpad >> 5;
            if (DIAG && lv < n_levels) {
                // 3b'. candidates = the list of the grid box the mobile atom is in (sitb_tables.cu: k_grid_lists);
                // every vertex of a candidate is screened at once (3c)
                const unsigned beg = __ldg(p.grid[lv].ptr + box), end = __ldg(p.grid[lv].ptr + box + 1);
                for (unsigned i0 = beg; i0 < end; i0 += 32) {
                    const unsigned i = i0 + lane;
                    const bool in = i < end;
                    const int k = in ? (int)__ldg(p.grid[lv].list + i) : 0;
                    const ushort4 vv = tva[k];
                    const float4 bb = tba[k];
                    // (no short-circuit: four independent gathers and compares, no branches)
                    bool ok = in & !(qfw[vv.x] > bb.x) & !(qfw[vv.y] > bb.y) & !(qfw[vv.z] > bb.z) & !(qfw[vv.w] > bb.w);
                    for (int blk = 1; blk < NB; ++blk) {
                        const ushort4 v2 = tva[(size_t)blk * Lpad + k];
                        const float4 b2 = tba[(size_t)blk * Lpad + k];
                        ok = ok & !(qfw[v2.x] > b2.x) & !(qfw[v2.y] > b2.y) & !(qfw[v2.z] > b2.z) & !(qfw[v2.w] > b2.w);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const int q = nsurv + __popc(m & lanemask_lt());
                        if (q < ENTRY_CAP) ek[q] = (uint16_t)k;
                    }
                    nsurv += __popc(m);
                }
                __syncwarp();
            } else
            for (int c0 = 0; c0 < n_chunks; c0 += 32) {
                // 3b. chunks of 32 landmarks (renumbered along a Morton curve of their first vertex): a lane
                // tests one chunk -- is any of its (<= 4) first-vertex atoms within its loosest bound? --
                // then only the active chunks (~30 %) test their 32 landmarks' first vertex.
                bool act = false;
                if (c0 + lane < n_chunks) {
                    const ushort4 ca = tca[c0 + lane];
                    const float4 cb = tcb[c0 + lane];
                    act = !(qfw[ca.x] > cb.x) || !(qfw[ca.y] > cb.y) || !(qfw[ca.z] > cb.z) || !(qfw[ca.w] > cb.w);
                }
                unsigned active = __ballot_sync(0xffffffffu, act);
                for (;;) {
                while (active) {
                    const int c = c0 + __ffs(active) - 1;
                    active &= active - 1u;
                    const int k = 32 * c + lane;
                    const bool ok = !(qfw[tv0[k]] > tb0[k]);
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (ok) cand[ncand + __popc(m & lanemask_lt())] = (uint16_t)k;
                    ncand += __popc(m);
                    if (ncand + 32 > CAND_CAP) break;      // list full: resume this round after 3c
                }
                __syncwarp();
                // run 3c when the list is (nearly) full or all chunks are done; otherwise keep collecting
                if (active == 0u && c0 + 32 < n_chunks && ncand + 32 <= CAND_CAP) break;
                // 3c. remaining vertices of the candidates -> survivors appended to ek[]
                for (int i0 = 0; i0 < ncand; i0 += 32) {
                    const int i = i0 + lane;
                    bool ok = i < ncand;
                    const int k = ok ? (int)cand[i] : 0;
                    if (ok) {
                        const ushort4 vv = tva[k];
                        const float4 bb = tba[k];
                        ok = !(qfw[vv.y] > bb.y) && !(qfw[vv.z] > bb.z) && !(qfw[vv.w] > bb.w);
                        for (int blk = 1; blk < NB && ok; ++blk) {
                            const ushort4 v2 = tva[(size_t)blk * Lpad + k];
                            const float4 b2 = tba[(size_t)blk * Lpad + k];
                            ok = !(qfw[v2.x] > b2.x) && !(qfw[v2.y] > b2.y) && !(qfw[v2.z] > b2.z) && !(qfw[v2.w] > b2.w);
                        }
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const int q = nsurv + __popc(m & lanemask_lt());
                        if (q < ENTRY_CAP) ek[q] = (uint16_t)k;
                    }
                    nsurv += __popc(m);
                }
                ncand = 0;
                __syncwarp();
                if (active == 0u) break;
                }
            }
            bool over = nsurv > ENTRY_CAP;
            if (over) nsurv = ENTRY_CAP;

            // 3d. exact test + value for the survivors (in-place: ek[i] -> ek[pos], ev[pos], pos <= i)
            int nent = 0;
            for (int i0 = 0; i0 < nsurv; i0 += 32) {
                const int i = i0 + lane;
                const bool in = i < nsurv;
                const int k = in ? (int)ek[i] : 0;
                bool alive = in;
                double P = 1.0;
                int nv = 0;
                if (in) {
                    for (int blk = 0; blk < NB; ++blk) {
                        const ushort4 vv = tva[(size_t)blk * Lpad + k];
                        const double2* q2 = (const double2*)(p.tab.q64 + (size_t)k * W + 4 * blk);
                        const double2* a2 = (const double2*)(p.tab.acoef + (size_t)k * W + 4 * blk);
                        const double2 q01 = __ldg(q2), q23 = __ldg(q2 + 1), a01 = __ldg(a2), a23 = __ldg(a2 + 1);
                        const unsigned vs[4] = {vv.x, vv.y, vv.z, vv.w};
                        const double qs[4] = {q01.x, q01.y, q23.x, q23.y};
                        const double as[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            if (vs[h] != (unsigned)S) {
                                const int src = p.dynamic ? (int)lmap[vs[h]] : (int)vs[h];
                                const double q = shifted_dist2<DIAG, true>(cell, sb[3 * src], sb[3 * src + 1], sb[3 * src + 2], ox, oy, oz);
                                if (q > qs[h]) alive = false;                 // helpers.pyx:199-203, exact
                                // 1 + exp(steep*(d/svd - mid))                 helpers.pyx:197,205
                                P *= 1.0 + exp_bounded(fmax(fma((q > 1e-30) ? sqrt_bounded(q) : 0.0, as[h], -p.bcoef), -700.0));
                                ++nv;
                            }
                        }
                    }
                    if (!alive) ++loc_rej;
                }
                const double val = inv_root(P, nv > 0 ? nv : 1);                // helpers.pyx:212
                const unsigned m = __ballot_sync(0xffffffffu, alive);
                if (alive) {
                    const int q = nent + __popc(m & lanemask_lt());
                    ek[q] = (uint16_t)k;
                    ev[q] = val;
                }
                nent += __popc(m);
            }
            __syncwarp();
            if (nent > 255) { over = true; nent = 255; }          // 8-bit entry count of a compressed row
            if (over && lane == 0) ++loc_over;
            if (lane == 0) {
                loc_nnz += (unsigned long long)nent;
                if (nent == 0) {
                    ++loc_zero;
                    if (p.errkey) atomicMin(p.errkey + 1, make_error_key(gframe, PHASE_ZERO_LVEC, (unsigned)j));
                }
            }

            // 3e. sinks
            if (MODE == MODE_ASSIGN) {
                // centres have disjoint supports (cluster/mcl.py:80): dot = sum over the row's
                // non-zeros of weight[landmark], grouped by cluster[landmark].  A row touches few
                // clusters: peel them off one at a time with warp votes.
                double bestc;           // untouched clusters have |dot| = 0; np.argmax -> index 0
                int bestid;
                if (nent <= 32) {
                    int myc[1] = {-1};
                    double mypr[1] = {0.0};
                    if (lane < nent) { const int k = ek[lane]; myc[0] = tcid[k]; mypr[0] = ev[lane] * tcw[k]; }
                    peel_clusters<1>(myc, mypr, lane, p.best, p.n_clusters, row_global, bestc, bestid);
                } else if (nent <= 128) {
                    int myc[4];
                    double mypr[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int e = 32 * c + lane;
                        myc[c] = -1; mypr[c] = 0.0;
                        if (e < nent) { const int k = ek[e]; myc[c] = tcid[k]; mypr[c] = ev[e] * tcw[k]; }
                    }
                    peel_clusters<4>(myc, mypr, lane, p.best, p.n_clusters, row_global, bestc, bestid);
                } else {
                    int myc[ENTRY_CAP / 32];
                    double mypr[ENTRY_CAP / 32];
#pragma unroll
                    for (int c = 0; c < ENTRY_CAP / 32; ++c) {
                        const int e = 32 * c + lane;
                        myc[c] = -1; mypr[c] = 0.0;
                        if (e < nent) { const int k = ek[e]; myc[c] = tcid[k]; mypr[c] = ev[e] * tcw[k]; }
                    }
                    peel_clusters<ENTRY_CAP / 32>(myc, mypr, lane, p.best, p.n_clusters, row_global, bestc, bestid);
                }
                long long label = bestid;
                double conf = bestc;
                if (nent == 0 || !(conf >= p.assign_thr)) { label = -1; conf = 0.0; }   // DotProdClassifier.pyx:168-172,184-186
                if (lane == 0) {
                    if (p.labels) p.labels[row_local] = label;
                    if (p.confs) p.confs[row_local] = conf;
                    if (label >= 0) {
                        if (p.counts) atomicAdd(&hist[label], 1u);
                        if (p.rep_w) atomicAdd(&p.rep_w[label], conf);
                        if (p.site_best) best_update(p.site_best, p.n_clusters, (int)label, conf, row_global);
                    }
                }
                if (p.rep && label >= 0) {
                    for (int e = lane; e < nent; e += 32)
                        atomicAdd(&p.rep[(size_t)label * L + p.tab.orig_of[ek[e]]], conf * ev[e]);
                }
            } else {
                // the other sinks address landmarks by the caller's numbering
                if (MODE == MODE_STATS || MODE == MODE_STAGE)
                    for (int e = lane; e < nent; e += 32) atomicAdd(&hist[ek[e]], 1u);
                for (int e = lane; e < nent; e += 32) ek[e] = p.tab.orig_of[ek[e]];
                __syncwarp();
                if ((MODE == MODE_STATS || MODE == MODE_STAGE) && p.sparse_ptr) {
                    // keep the row in compressed form: later passes read ~250 B instead of recomputing it
                    unsigned long long off = 0;
                    const bool slotted = p.sparse_slot && (unsigned)nent <= p.sparse_slot;
                    if (slotted) {
                        // row-ordered fixed slots: the later passes stream them in order (rows scattered over the pool cost
                        // a DRAM page per row: those passes ran at a fifth of the HBM rate)
                        off = (unsigned long long)(p.sparse_row_base + row_local) * p.sparse_slot;
                    } else if (lane == 0 && nent > 0) {            // (an all-zero row needs no space: offset 0, count 0)
                        if (*pool_left < (unsigned)nent) {         // next slice (the rest of the old one stays unused)
                            *pool_off = atomicAdd(p.sparse_cursor, (unsigned long long)POOL_SLICE);
                            *pool_left = POOL_SLICE;
                        }
                        off = *pool_off;
                        *pool_off = off + (unsigned long long)nent;
                        *pool_left -= (unsigned)nent;
                    }
                    if (!slotted) off = __shfl_sync(0xffffffffu, off, 0);
                    if (off + (unsigned long long)nent <= p.sparse_capacity) {
                        for (int e = lane; e < nent; e += 32) { p.sparse_k[off + e] = ek[e]; p.sparse_v[off + e] = ev[e]; }
                        if (lane == 0) p.sparse_ptr[row_local] = (off << 8) | (unsigned long long)nent;
                    } else if (lane == 0) {
                        p.sparse_ptr[row_local] = ~0ull;          // pool exhausted: the host grows it and reruns
                    }
                }
                if (MODE == MODE_DENSE) {
                    if (p.dense_f64) {
                        double* o = (double*)p.dense_out + (size_t)row_local * L;
                        for (int e = lane; e < nent; e += 32) o[ek[e]] = ev[e];
                    } else {
                        float* o = (float*)p.dense_out + (size_t)row_local * L;
                        for (int e = lane; e < nent; e += 32) o[ek[e]] = (float)ev[e];
                    }
                }
                if (MODE == MODE_STATS && p.gram) {
                    // upper triangle of the outer product (when the rows are cached, sitb_gram_sparse.cu
                    // builds the Gram from them with far fewer atomics and p.gram is null here)
                    for (int a = 0; a < nent; ++a) {
                        const unsigned ka = ek[a];
                        const double va = ev[a];
                        for (int bq = a + lane; bq < nent; bq += 32) {
                            const unsigned kq = ek[bq];
                            const unsigned lo = ka < kq ? ka : kq, hi = ka < kq ? kq : ka;
                            atomicAdd(&p.gram[(size_t)lo * L + hi], va * ev[bq]);
                        }
                    }
                }
                if (MODE == MODE_STAGE) {
                    for (int e = lane; e < nent; e += 32) {
                        const double v = ev[e];
                        const __half hi = __double2half(v);
                        const __half lo = __double2half((v - (double)__half2float(hi)) * 4096.0);   // scaled: stays normal
                        // tile (landmark / 128, row / 64), element (r, k) at the 128B-swizzled K-major position
                        const unsigned l = ek[e], r = l & 127u, k = (unsigned)(row_local & 63);
                        const size_t o = (((size_t)(l >> 7) * (size_t)(p.stage_ld >> 6) + (size_t)(row_local >> 6)) << 13) +
                                         (size_t)(r * 64u + ((((k >> 3) ^ (r & 7u)) << 3) | (k & 7u)));
                        p.stage_hi[o] = hi;
                        p.stage_lo[o] = lo;
                    }
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
                if (hist[i]) atomicAdd(&dst[(MODE == MODE_ASSIGN) ? i : (int)p.tab.orig_of[i]], (unsigned long long)hist[i]);
    }
    if (p.counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            loc_rej += __shfl_xor_sync(0xffffffffu, loc_rej, o);
            loc_dup += __shfl_xor_sync(0xffffffffu, loc_dup, o);
        }
        if (lane == 0) {
            if (loc_zero) atomicAdd(&p.counters[CNT_ZERO_ROWS], loc_zero);
            if (loc_nnz) atomicAdd(&p.counters[CNT_NNZ], loc_nnz);
            if (loc_over) atomicAdd(&p.counters[CNT_LIST_OVERFLOW], loc_over);
            if (loc_rej) atomicAdd(&p.counters[CNT_SCREEN_REJECT], loc_rej);
            if (loc_dup) atomicAdd(&p.counters[CNT_DUP_NEAREST], loc_dup);
            if (loc_full) atomicAdd(&p.counters[CNT_FULL_WALK_FRAMES], loc_full);
            if (loc_loose) atomicAdd(&p.counters[CNT_LOOSE_GRID_FRAMES], loc_loose);
        }
    }
}

template <bool DIAG, int MODE>
static cudaError_t launch_one(const FillParams& p, int n_sms, cudaStream_t stream) {
    // warps per CTA and frames per batch: as many warps as fit the shared-memory budget with two
    // CTAs per SM if possible; mobile atoms are claimed dynamically inside a batch, so the longest
    // batch that fits is best (fewer CTA barriers, better balance)
    auto kern = k_fill<DIAG, MODE>;
    // CTAs per SM x warps per CTA x frames per batch: maximise resident warps per SM within the
    // 227 KB of shared memory (tables are per CTA, lists per warp, frame buffers per batch); mobile
    // atoms are claimed dynamically inside a batch, so want >= 3 tasks per warp and batch
    const int max_w = DIAG ? SITB_K1_WARPS : 16;
    // resident warps are also bounded by the register file (64 K registers per SM, allocated per warp in units of
    // 8 registers per thread)
    cudaFuncAttributes fa;
    cudaError_t e0 = cudaFuncGetAttributes(&fa, kern);
    if (e0 != cudaSuccess) return e0;
    const int regs = ((fa.numRegs > 0 ? fa.numRegs : 64) + 7) & ~7;
    const int reg_warps = 65536 / (32 * regs);
    int best_w = 0, best_fb = 1;
    size_t best_bytes = 0;
    double best_score = -1.0;
#ifdef SITB_K1_FORCE_CTAS
    for (int ctas = SITB_K1_FORCE_CTAS; ctas >= SITB_K1_FORCE_CTAS; --ctas) {
#else
    for (int ctas = 2; ctas >= 1; --ctas) {
#endif
        const size_t budget = (size_t)(227 * 1024) / ctas - 1024;
        for (int w = max_w; w >= 1; --w) {
            if (w * ctas > 64 || w * ctas > reg_warps) continue;
            int fb_fit = 0;
            size_t bytes_fit = 0;
            for (int fb = 1; fb <= 8; ++fb) {
                const size_t bytes = make_layout(p.S, p.M, p.L, p.Lpad, p.NB, w, fb, MODE, p.n_clusters, p.dynamic).total;
                if (bytes > budget) break;
                fb_fit = fb; bytes_fit = bytes;
            }
            if (!fb_fit) continue;
            const double fill = (double)(fb_fit * p.M) / (3.0 * w);
            const double score = (double)(ctas * w) * (fill < 1.0 ? fill : 1.0) - 0.01 * ctas;   // equal warps: one CTA with longer batches measured faster
            if (score > best_score) { best_score = score; best_w = w; best_fb = fb_fit; best_bytes = bytes_fit; }
        }
    }
    if (best_w == 0) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)best_bytes);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, best_w * 32, best_bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long batches = (p.n_work + best_fb - 1) / best_fb;
    long long grid = (long long)n_sms * per_sm;
    if (grid > batches) grid = batches;
    if (grid < 1) return cudaSuccess;
    const SmemLayout lay = make_layout(p.S, p.M, p.L, p.Lpad, p.NB, best_w, best_fb, MODE, p.n_clusters, p.dynamic);
    kern<<<(unsigned)grid, best_w * 32, best_bytes, stream>>>(p, best_fb, lay);
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
