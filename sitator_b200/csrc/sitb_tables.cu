// sitator_b200 -- landmark tables (once per analysis).
//
//   LandmarkAnalysis.py:194-202   verts_np (-1 padded) and site_vert_dists = distance from each
//                                 landmark centre to each of its vertex atoms at their ideal
//                                 positions (PBCCalculator.distances, PBCCalculator.pyx:64-103)
//   helpers.pyx:197-203           the cut-off test  dist/site_vert_dist > cutoff_round_to_zero
//
// The test is turned into a compare on the squared distance: with IEEE sqrt and divide both
// monotone, there is a largest double T with T/svd <= cutoff and a largest double Q with
// sqrt(Q) <= T; then  sqrt(q)/svd > cutoff  <=>  q > Q  for every double q, bit for bit.
//
// The device computes svd and Q per (landmark, vertex) in the reference's order; the host then
// lays the kernel tables out (build_landmark_tables): landmarks are renumbered so that those
// sharing a first vertex are adjacent (the first-vertex gather in K1 then hits one or two
// shared-memory words per warp instead of 32 random ones), each landmark's tightest vertex is
// tested first, and missing vertices become the always-passing dummy vertex S.
#include "sitb_fill.cuh"
#include <math_constants.h>
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <vector>

namespace sitb {

__device__ double exact_q_cutoff(double svd, double cutoff) {
    double t = __dmul_rn(cutoff, svd);
    for (int it = 0; it < 64 && __ddiv_rn(t, svd) > cutoff; ++it) t = nextafter(t, 0.0);
    for (int it = 0; it < 64 && __ddiv_rn(nextafter(t, CUDART_INF), svd) <= cutoff; ++it) t = nextafter(t, CUDART_INF);
    double Q = __dmul_rn(t, t);
    for (int it = 0; it < 64 && __dsqrt_rn(Q) > t; ++it) Q = nextafter(Q, 0.0);
    for (int it = 0; it < 64 && __dsqrt_rn(nextafter(Q, CUDART_INF)) <= t; ++it) Q = nextafter(Q, CUDART_INF);
    return Q;
}

// svd_out[L][V] (NaN padded) and q_out[L][V] (+inf padded), both in the reference's layout
__global__ void k_tables(Cell cell, const double* __restrict__ centers, const double* __restrict__ ideal,
                         const int* __restrict__ verts_in, int L, int V, int S, double cutoff,
                         double* __restrict__ svd_out, double* __restrict__ q_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L * V) return;
    const int k = idx / V, h = idx - k * V;
    int v = verts_in[idx];
    for (int hh = 0; hh < h; ++hh)
        if (verts_in[k * V + hh] < 0) v = -1;       // the reference stops at the first -1 (helpers.pyx:192-193)
    if (v < 0 || v >= S) {
        svd_out[idx] = CUDART_NAN;
        q_out[idx] = CUDART_INF;
        return;
    }
    const double ox = __dsub_rn(cell.cen[0], centers[3 * k + 0]);
    const double oy = __dsub_rn(cell.cen[1], centers[3 * k + 1]);
    const double oz = __dsub_rn(cell.cen[2], centers[3 * k + 2]);
    const double q = cell.diag
        ? shifted_dist2<true, false>(cell, ideal[3 * v], ideal[3 * v + 1], ideal[3 * v + 2], ox, oy, oz)
        : shifted_dist2<false, false>(cell, ideal[3 * v], ideal[3 * v + 1], ideal[3 * v + 2], ox, oy, oz);
    const double svd = __dsqrt_rn(q);
    svd_out[idx] = svd;
    // degenerate landmark (centre on a vertex): every ratio is inf/nan -> never counted
    q_out[idx] = (svd > 0.0 && cutoff > 0.0) ? exact_q_cutoff(svd, cutoff) : -1.0;
}

cudaError_t launch_tables(const Cell& cell, const double* centers, const double* ideal, const int* verts_in, int L,
                          int V, int S, double cutoff, double* svd_out, double* q_out, cudaStream_t stream) {
    const int n = L * V;
    k_tables<<<(n + 127) / 128, 128, 0, stream>>>(cell, centers, ideal, verts_in, L, V, S, cutoff, svd_out, q_out);
    return cudaGetLastError();
}

// ---- cell grid of candidate landmarks (orthorhombic cells) ---------------------------------------
// A landmark's component is non-zero only if the mobile atom lies within R_h = sqrt(Q_h) of every vertex
// atom (helpers.pyx:197-203).  For a frame whose static atoms all sit within `margin` of their ideal
// positions (measured anyway by the static-lattice check, helpers.pyx:57-67), that region is contained in
// the intersection of the balls of radius R_h + margin about the IDEAL vertex positions (the per-axis
// minimum-image metric of an orthorhombic cell obeys the triangle inequality).  The cell is cut into
// gx*gy*gz boxes; a box lists every landmark whose balls all reach it.  K1 then tests only the list of the
// box the mobile atom is in (~6 % of the landmarks at the LLZO shape) instead of walking all of them; a
// frame with a static atom beyond the margin walks all landmarks as before.  One warp per box; lists
// are in ascending internal landmark order, the same order the full walk produces.
// (ideal = the static-lattice positions wrapped into the cell, Cartesian; q64 holds the cut-off RADII sqrt(Q) here)
// Both margins are tested in one sweep: masks[box * n_chunks + chunk] = (ballot for margin0, ballot for margin1) over the
// chunk's 32 landmarks, count[box] / count[cells + box] = list lengths.  margin0 <= margin1 (level 0 is a subset of level 1).
__global__ void k_grid_masks(Cell cell, const double* __restrict__ ideal, const ushort4* __restrict__ va,
                             const double* __restrict__ q64, int L, int Lpad, int NB, int S, int gx, int gy, int gz,
                             double margin0, double margin1, uint2* __restrict__ masks, unsigned* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const long long cells = (long long)gx * gy * gz;
    const long long id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id >= cells) return;
    const int iz = (int)(id % gz), iy = (int)((id / gz) % gy), ix = (int)(id / ((long long)gz * gy));
    const double len[3] = {cell.c[0], cell.c[4], cell.c[8]};
    const double inv_len[3] = {1.0 / len[0], 1.0 / len[1], 1.0 / len[2]};
    const double half[3] = {0.5 * len[0] / gx, 0.5 * len[1] / gy, 0.5 * len[2] / gz};
    const double mid[3] = {(2 * ix + 1) * half[0], (2 * iy + 1) * half[1], (2 * iz + 1) * half[2]};
    const int W = 4 * NB;
    const int n_chunks = (L + 31) >> 5;
    unsigned n0 = 0, n1 = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int k = 32 * ch + lane;
        bool in1 = k < L, in0 = in1;
        for (int blk = 0; blk < NB && in1; ++blk) {
            const ushort4 vv = va[(size_t)blk * Lpad + k];
            const unsigned vs[4] = {vv.x, vv.y, vv.z, vv.w};
            for (int h = 0; h < 4 && in1; ++h) {
                if (vs[h] == (unsigned)S) continue;
                const double R = q64[(size_t)k * W + 4 * blk + h];
                if (!(R >= 0.0)) { in1 = false; break; }          // degenerate landmark: never non-zero
                double d2 = 0.0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    double x = ideal[3 * vs[h] + d] - mid[d];
                    x -= len[d] * rint(x * inv_len[d]);           // |x| <= len/2 (+ one rounding: covered by the margin's slack)
                    const double a = fmax(fabs(x) - half[d], 0.0);
                    d2 += a * a;
                }
                if (d2 > (R + margin1) * (R + margin1)) in1 = false;
                if (d2 > (R + margin0) * (R + margin0)) in0 = false;
            }
        }
        in0 = in0 && in1;
        const unsigned m0 = __ballot_sync(0xffffffffu, in0), m1 = __ballot_sync(0xffffffffu, in1);
        if (lane == 0) masks[id * n_chunks + ch] = make_uint2(m0, m1);
        n0 += __popc(m0); n1 += __popc(m1);
    }
    if (lane == 0) { count[id] = n0; count[cells + id] = n1; }
}

// The same grid for the static atoms: a box lists every static-lattice site that is a vertex of one of the
// box's candidate landmarks, i.e. lies within rmax[s] + margin of the box (rmax[s] = the largest cut-off
// radius of any landmark vertex on s).  K1 then computes screen distances for those sites only.
__global__ void k_grid_static_masks(Cell cell, const double* __restrict__ ideal, const double* __restrict__ rmax, int S,
                                    int gx, int gy, int gz, double margin0, double margin1, uint2* __restrict__ masks,
                                    unsigned* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const long long cells = (long long)gx * gy * gz;
    const long long id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id >= cells) return;
    const int iz = (int)(id % gz), iy = (int)((id / gz) % gy), ix = (int)(id / ((long long)gz * gy));
    const double len[3] = {cell.c[0], cell.c[4], cell.c[8]};
    const double inv_len[3] = {1.0 / len[0], 1.0 / len[1], 1.0 / len[2]};
    const double half[3] = {0.5 * len[0] / gx, 0.5 * len[1] / gy, 0.5 * len[2] / gz};
    const double mid[3] = {(2 * ix + 1) * half[0], (2 * iy + 1) * half[1], (2 * iz + 1) * half[2]};
    const int n_chunks = (S + 31) >> 5;
    unsigned n0 = 0, n1 = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int s = 32 * ch + lane;
        bool in0 = false, in1 = false;
        if (s < S) {
            const double r = rmax[s];
            if (r >= 0.0) {                                       // (< 0: not a vertex of any landmark)
                double d2 = 0.0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    double x = ideal[3 * s + d] - mid[d];
                    x -= len[d] * rint(x * inv_len[d]);
                    const double a = fmax(fabs(x) - half[d], 0.0);
                    d2 += a * a;
                }
                in1 = !(d2 > (r + margin1) * (r + margin1));
                in0 = in1 && !(d2 > (r + margin0) * (r + margin0));
            }
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, in0), m1 = __ballot_sync(0xffffffffu, in1);
        if (lane == 0) masks[id * n_chunks + ch] = make_uint2(m0, m1);
        n0 += __popc(m0); n1 += __popc(m1);
    }
    if (lane == 0) { count[id] = n0; count[cells + id] = n1; }
}

// masks -> lists (ascending index, as the full walk produces them), one warp per box.  cat_list / cat_box (optional): both
// levels in one array with entries 4 * index and (offset, count) per box -- the form the two-tier first tier reads.
__global__ void k_grid_lists_from_masks(const uint2* __restrict__ masks, int n_chunks, long long cells,
                                        const unsigned* __restrict__ ptr0, const unsigned* __restrict__ ptr1,
                                        uint16_t* __restrict__ list0, uint16_t* __restrict__ list1,
                                        uint16_t* __restrict__ cat_list, uint2* __restrict__ cat_box) {
    const int lane = threadIdx.x & 31;
    const long long id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id >= cells) return;
    const unsigned b0 = ptr0[id], b1 = ptr1[id];
    const unsigned cat1 = ptr0[cells];                            // level 1 follows all of level 0
    unsigned n0 = 0, n1 = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
        const uint2 m = masks[id * n_chunks + ch];
        const unsigned k = 32u * ch + lane;
        if ((m.x >> lane) & 1u) {
            const unsigned at = b0 + n0 + __popc(m.x & lanemask_lt());
            list0[at] = (uint16_t)k;
            if (cat_list) cat_list[at] = (uint16_t)(4u * k);
        }
        if ((m.y >> lane) & 1u) {
            const unsigned at = b1 + n1 + __popc(m.y & lanemask_lt());
            list1[at] = (uint16_t)k;
            if (cat_list) cat_list[cat1 + at] = (uint16_t)(4u * k);
        }
        n0 += __popc(m.x); n1 += __popc(m.y);
    }
    if (cat_box && lane == 0) {
        cat_box[id] = make_uint2(b0, n0);
        cat_box[cells + id] = make_uint2(cat1 + b1, n1);
    }
}

cudaError_t launch_grid_static_masks(const Cell& cell, const double* ideal, const double* rmax, int S, int gx, int gy,
                                     int gz, double margin0, double margin1, uint2* masks, unsigned* count,
                                     cudaStream_t stream) {
    const long long cells = (long long)gx * gy * gz;
    const int wpb = 8;
    k_grid_static_masks<<<(unsigned)((cells + wpb - 1) / wpb), wpb * 32, 0, stream>>>(cell, ideal, rmax, S, gx, gy, gz, margin0,
                                                                                 margin1, masks, count);
    return cudaGetLastError();
}

cudaError_t launch_grid_masks(const Cell& cell, const double* ideal, const ushort4* va, const double* q64, int L,
                              int Lpad, int NB, int S, int gx, int gy, int gz, double margin0, double margin1, uint2* masks,
                              unsigned* count, cudaStream_t stream) {
    const long long cells = (long long)gx * gy * gz;
    const int wpb = 8;
    k_grid_masks<<<(unsigned)((cells + wpb - 1) / wpb), wpb * 32, 0, stream>>>(cell, ideal, va, q64, L, Lpad, NB, S, gx, gy,
                                                                          gz, margin0, margin1, masks, count);
    return cudaGetLastError();
}

cudaError_t launch_grid_lists_from_masks(const uint2* masks, int n_chunks, long long cells, const unsigned* ptr0,
                                         const unsigned* ptr1, uint16_t* list0, uint16_t* list1, uint16_t* cat_list,
                                         uint2* cat_box, cudaStream_t stream) {
    const int wpb = 8;
    k_grid_lists_from_masks<<<(unsigned)((cells + wpb - 1) / wpb), wpb * 32, 0, stream>>>(masks, n_chunks, cells, ptr0, ptr1,
                                                                                     list0, list1, cat_list, cat_box);
    return cudaGetLastError();
}

// Float screen bound (sitb_fill.cu steps 3a-3c).  Orthorhombic cells: the screen distance is
// computed in FP32 from float fractional coordinates as |(u - round(u)) * L|^2; each Cartesian
// component is then off by at most ~1.8e-7 * L (two float conversions, one subtract, one
// multiply); we allow 24 * 2^-24 * Lmax = 1.4e-6 * Lmax, i.e. 8x that, and bound the effect on
// the squared distance at the cut-off.  Triclinic cells: the screen value is the exact double
// rounded to float, and rounding is monotone, so float(Q) itself is the bound.
static float screen_bound(const Cell& cell, double Q) {
    if (!(Q > 0.0)) return -1.0f;
    if (std::isinf(Q)) return std::numeric_limits<float>::infinity();
    if (cell.diag) {
        const double lmax = std::max(std::fabs(cell.c[0]), std::max(std::fabs(cell.c[4]), std::fabs(cell.c[8])));
        const double dc = 24.0 * 5.9604644775390625e-08 * lmax;
        const double marg = 2.0 * std::sqrt(3.0 * Q) * dc + 3.0 * dc * dc + 8.0 * 5.9604644775390625e-08 * Q;
        return std::nextafter((float)(Q + marg), std::numeric_limits<float>::infinity());
    }
    return (float)Q;    // round to nearest
}

void build_landmark_tables(const Cell& cell, int L, int V, int Lpad, int NB, int S, double steep_log2e,
                           const int* verts_in, const double* ideal, const double* svd, const double* q,
                           HostTables& out) {
    const int W = 4 * NB;
    // vertex order per landmark: tightest cut-off first, the rest in the reference's order
    std::vector<int> nv(L), first(L);
    for (int k = 0; k < L; ++k) {
        int n = 0;
        while (n < V && verts_in[(size_t)k * V + n] >= 0 && verts_in[(size_t)k * V + n] < S) ++n;
        nv[k] = n;
        int best = 0;
        for (int h = 1; h < n; ++h)
            if (q[(size_t)k * V + h] < q[(size_t)k * V + best]) best = h;
        first[k] = best;
    }
    // renumber: by the position of the first-vertex atom along a Morton curve (so that the landmarks a
    // mobile atom can see sit in few 32-landmark chunks), then by atom, then by original index
    std::vector<unsigned> morton(S, 0u);
    for (int s = 0; s < S; ++s) {
        unsigned code = 0;
        unsigned g[3];
        for (int d = 0; d < 3; ++d) {
            double f = cell.ci[3 * d] * ideal[3 * s] + cell.ci[3 * d + 1] * ideal[3 * s + 1] + cell.ci[3 * d + 2] * ideal[3 * s + 2];
            f -= std::floor(f);
            int v = (int)(f * 1024.0);
            g[d] = (unsigned)(v < 0 ? 0 : (v > 1023 ? 1023 : v));
        }
        for (int b = 9; b >= 0; --b)
            for (int d = 0; d < 3; ++d) code = (code << 1) | ((g[d] >> b) & 1u);
        morton[s] = code;
    }
    std::vector<int> order(L);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const int va = verts_in[(size_t)a * V + first[a]], vb = verts_in[(size_t)b * V + first[b]];
        if (morton[va] != morton[vb]) return morton[va] < morton[vb];
        return va < vb;
    });
    out.v0.assign(Lpad, (uint16_t)S);
    out.b0.assign(Lpad, -1.0f);
    out.va.assign((size_t)NB * Lpad, make_ushort4((uint16_t)S, (uint16_t)S, (uint16_t)S, (uint16_t)S));
    const float inf = std::numeric_limits<float>::infinity();
    out.ba.assign((size_t)NB * Lpad, make_float4(inf, inf, inf, inf));
    out.q64.assign((size_t)Lpad * W, std::numeric_limits<double>::infinity());
    out.acoef.assign((size_t)Lpad * W, 0.0);
    out.nverts.assign(Lpad, 0);
    out.orig_of.assign(Lpad, 0);
    out.internal_of.assign(L, 0);
    for (int ki = 0; ki < L; ++ki) {
        const int k = order[ki];
        out.orig_of[ki] = (uint16_t)k;
        out.internal_of[k] = ki;
        out.nverts[ki] = (uint8_t)nv[k];
        uint16_t vs[MAX_VERTS];
        float bs[MAX_VERTS];
        for (int h = 0; h < W; ++h) { vs[h] = (uint16_t)S; bs[h] = inf; }
        int slot = 0;
        for (int pass = 0; pass < 2; ++pass) {
            for (int h = 0; h < nv[k]; ++h) {
                if ((pass == 0) != (h == first[k])) continue;
                const double Q = q[(size_t)k * V + h], sv = svd[(size_t)k * V + h];
                vs[slot] = (uint16_t)verts_in[(size_t)k * V + h];
                bs[slot] = screen_bound(cell, Q);
                out.q64[(size_t)ki * W + slot] = Q;
                out.acoef[(size_t)ki * W + slot] = (sv > 0.0) ? steep_log2e / sv : 0.0;
                ++slot;
            }
        }
        out.v0[ki] = vs[0];
        out.b0[ki] = (nv[k] > 0) ? bs[0] : -1.0f;
        for (int blk = 0; blk < NB; ++blk) {
            out.va[(size_t)blk * Lpad + ki] = make_ushort4(vs[4 * blk], vs[4 * blk + 1], vs[4 * blk + 2], vs[4 * blk + 3]);
            out.ba[(size_t)blk * Lpad + ki] = make_float4(bs[4 * blk], bs[4 * blk + 1], bs[4 * blk + 2], bs[4 * blk + 3]);
        }
    }
    // per static-lattice site: the largest cut-off radius of any landmark vertex on it (-1: not a vertex)
    out.rmax.assign((size_t)S, -1.0);
    for (int k = 0; k < L; ++k)
        for (int h = 0; h < nv[k]; ++h) {
            const int v = verts_in[(size_t)k * V + h];
            const double Q = q[(size_t)k * V + h];
            if (Q >= 0.0 && std::sqrt(Q) > out.rmax[v]) out.rmax[v] = std::sqrt(Q);
        }
    // chunk skip table: per 32 consecutive landmarks, the (up to 4) distinct first-vertex atoms and, per
    // atom, the loosest first-vertex bound among the chunk's landmarks that use it.  A chunk whose atoms
    // all lie beyond their bound holds no candidate.  More than 4 distinct atoms: never skipped.
    const int n_chunks = Lpad / 32;
    out.chunk_atoms.assign(n_chunks, make_ushort4((uint16_t)S, (uint16_t)S, (uint16_t)S, (uint16_t)S));
    out.chunk_bound.assign(n_chunks, make_float4(-1.f, -1.f, -1.f, -1.f));
    for (int c = 0; c < n_chunks; ++c) {
        uint16_t at[4];
        float bd[4];
        int n = 0;
        bool overflow = false;
        for (int i = 0; i < 32; ++i) {
            const int ki = 32 * c + i;
            if (ki >= L) break;
            const uint16_t a = out.v0[ki];
            const float b = out.b0[ki];
            int j = 0;
            while (j < n && at[j] != a) ++j;
            if (j == n) {
                if (n == 4) { overflow = true; break; }
                at[n] = a; bd[n] = b; ++n;
            } else if (b > bd[j]) {
                bd[j] = b;
            }
        }
        if (overflow) {   // dummy atom S has screen distance 0: 0 <= +inf always passes
            out.chunk_atoms[c] = make_ushort4((uint16_t)S, (uint16_t)S, (uint16_t)S, (uint16_t)S);
            out.chunk_bound[c] = make_float4(inf, inf, inf, inf);
        } else {
            for (int j = n; j < 4; ++j) { at[j] = (uint16_t)S; bd[j] = -1.f; }
            out.chunk_atoms[c] = make_ushort4(at[0], at[1], at[2], at[3]);
            out.chunk_bound[c] = make_float4(bd[0], bd[1], bd[2], bd[3]);
        }
    }
}

}  // namespace sitb
