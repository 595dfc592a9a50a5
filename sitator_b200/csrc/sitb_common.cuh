// sitator_b200 -- shared device helpers (sm_100a).
//
// The periodic-boundary maths here restates the reference's PBCCalculator
// (sitator/util/PBCCalculator.pyx:341-366 wrap_points, :64-103 distances) operation by
// operation in IEEE double with explicit round-to-nearest intrinsics, so that nvcc can not
// contract a*b+c into an FMA: the C the reference is compiled to (gcc -O2, x86-64) has no
// FMA either, hence squared distances come out bit-identical and the landmark support
// (helpers.pyx:199-203, a discontinuity at ratio 1.807) is reproduced exactly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sitb {

// Device scratch and context buffers come from the device's stream-ordered memory pool, kept warm (release
// threshold = everything): after the first analysis no call goes back to the driver's cudaMalloc/cudaFree,
// which synchronise the device and take milliseconds for trajectory-sized blocks.
inline cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t st) {
    static bool warmed[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !warmed[dev]) {
        cudaMemPool_t pool;
        e = cudaDeviceGetDefaultMemPool(&pool, dev);
        if (e != cudaSuccess) return e;
        unsigned long long keep = ~0ull;
        e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return e;
        warmed[dev] = true;
    }
    return cudaMallocAsync(p, bytes ? bytes : 1, st);
}
inline void pool_free(void* p, cudaStream_t st) {
    if (p) cudaFreeAsync(p, st);
}


struct Cell {
    double c[9];    // cellmat  = cell^T, row major   (PBCCalculator.pyx:33)
    double ci[9];   // cellmat^-1, row major          (PBCCalculator.pyx:34)
    double cen[3];  // centroid = sum(0.5 * cell)     (PBCCalculator.pyx:35)
    int diag;       // 1 when both matrices are exactly diagonal (orthorhombic cell)
};

// error key: [frame:40][phase:4][index:20]; smaller key = earlier in the reference's loop order
// (helpers.pyx:50 frame loop; :57-92 lattice checks before :95-122 mobile loop).
enum : unsigned { PHASE_STATIC_MOVED = 1, PHASE_STATIC_UNASSIGNED = 2, PHASE_ZERO_LVEC = 3 };
static constexpr unsigned long long NO_ERROR_KEY = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ inline unsigned long long make_error_key(long long frame, unsigned phase, unsigned index) {
    return ((unsigned long long)frame << 24) | ((unsigned long long)(phase & 0xF) << 20) |
           (unsigned long long)(index & 0xFFFFF);
}

// m[d,0]*p0 + m[d,1]*p1 + m[d,2]*p2, evaluated left to right without contraction
__device__ __forceinline__ double rowdot(const double* __restrict__ m, int d, double p0, double p1, double p2) {
    return __dadd_rn(__dadd_rn(__dmul_rn(m[3 * d + 0], p0), __dmul_rn(m[3 * d + 1], p1)),
                     __dmul_rn(m[3 * d + 2], p2));
}

// x - floor(x), as the reference computes it (PBCCalculator.pyx:358-360).  Callers use it for the fractional
// coordinate of (static + centroid - mobile) with both atoms already wrapped into the cell (step 1 of K1).
__device__ __forceinline__ double frac_near(double f) {
    return __dsub_rn(f, floor(f));      // FRND.F64.FLOOR + DADD: two instructions (two compares and selects were six)
}

// PBCCalculator.wrap_points for one point (general triclinic cell). NEAR: |frac| is small.
template <bool NEAR>
__device__ __forceinline__ void wrap_general(const Cell& cell, double& x, double& y, double& z) {
    double f0 = rowdot(cell.ci, 0, x, y, z);
    double f1 = rowdot(cell.ci, 1, x, y, z);
    double f2 = rowdot(cell.ci, 2, x, y, z);
    if (NEAR) {
        f0 = frac_near(f0); f1 = frac_near(f1); f2 = frac_near(f2);
    } else {
        f0 = __dsub_rn(f0, floor(f0)); f1 = __dsub_rn(f1, floor(f1)); f2 = __dsub_rn(f2, floor(f2));
    }
    x = rowdot(cell.c, 0, f0, f1, f2);
    y = rowdot(cell.c, 1, f0, f1, f2);
    z = rowdot(cell.c, 2, f0, f1, f2);
}

// Same for an exactly diagonal cell: the off-diagonal products are +-0 and adding them
// changes nothing, so one multiply per axis gives the identical bits.
template <bool NEAR>
__device__ __forceinline__ void wrap_diag(const Cell& cell, double& x, double& y, double& z) {
    double f0 = __dmul_rn(cell.ci[0], x);
    double f1 = __dmul_rn(cell.ci[4], y);
    double f2 = __dmul_rn(cell.ci[8], z);
    if (NEAR) {
        f0 = frac_near(f0); f1 = frac_near(f1); f2 = frac_near(f2);
    } else {
        f0 = __dsub_rn(f0, floor(f0)); f1 = __dsub_rn(f1, floor(f1)); f2 = __dsub_rn(f2, floor(f2));
    }
    x = __dmul_rn(cell.c[0], f0);
    y = __dmul_rn(cell.c[4], f1);
    z = __dmul_rn(cell.c[8], f2);
}

// diagonal cell, also returning the wrapped fractional coordinates (in [0,1]) for the FP32 pre-screen
__device__ __forceinline__ void wrap_diag_frac(const Cell& cell, double& x, double& y, double& z,
                                               double& f0, double& f1, double& f2) {
    f0 = __dmul_rn(cell.ci[0], x);
    f1 = __dmul_rn(cell.ci[4], y);
    f2 = __dmul_rn(cell.ci[8], z);
    f0 = __dsub_rn(f0, floor(f0)); f1 = __dsub_rn(f1, floor(f1)); f2 = __dsub_rn(f2, floor(f2));
    x = __dmul_rn(cell.c[0], f0);
    y = __dmul_rn(cell.c[4], f1);
    z = __dmul_rn(cell.c[8], f2);
}

template <bool DIAG, bool NEAR>
__device__ __forceinline__ void wrap_point(const Cell& cell, double& x, double& y, double& z) {
    if (DIAG) wrap_diag<NEAR>(cell, x, y, z);
    else wrap_general<NEAR>(cell, x, y, z);
}

// Squared shift-and-wrap distance between a point p and a reference point whose offset
// (centroid - ref) is (ox, oy, oz):  | wrap(p + off) - centroid |^2
// (helpers.pyx:99-103 + :174-178 ; PBCCalculator.pyx:84-100).
template <bool DIAG, bool NEAR>
__device__ __forceinline__ double shifted_dist2(const Cell& cell, double px, double py, double pz,
                                                double ox, double oy, double oz) {
    double x = __dadd_rn(px, ox), y = __dadd_rn(py, oy), z = __dadd_rn(pz, oz);
    wrap_point<DIAG, NEAR>(cell, x, y, z);
    const double dx = __dsub_rn(x, cell.cen[0]);
    const double dy = __dsub_rn(y, cell.cen[1]);
    const double dz = __dsub_rn(z, cell.cen[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace sitb
