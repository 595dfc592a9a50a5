// sitator_b200 -- K1: fused wrap + static-lattice check + landmark-vector fill (+ assign).
// Parameter block shared by the kernel (sitb_fill.cu) and the C-ABI (sitb_api.cu).
#pragma once
#include <cuda_fp16.h>
#include "sitb_common.cuh"

namespace sitb {

enum FillMode : int {
    MODE_DENSE = 0,   // materialise landmark vectors (helpers._fill_landmark_vectors drop-in)
    MODE_STATS = 1,   // seen counts, zero rows, Gram by sparse outer products (FP64 atomics)
    MODE_STAGE = 2,   // seen counts, zero rows, fp16 hi/lo transposed staging chunk for the tcgen05 SYRK
    MODE_ASSIGN = 3,  // centre similarity, threshold, argmax (+ optional reductions)
};

static constexpr int ENTRY_CAP = 256;        // non-zero components kept per landmark vector
static constexpr uint16_t VERT_END = 0xFFFF; // end of a landmark's vertex list (the reference's -1)

// counters[] slots
enum : int { CNT_ZERO_ROWS = 0, CNT_DUP_NEAREST = 1, CNT_LIST_OVERFLOW = 2, CNT_NNZ = 3,
             CNT_TIE_EXACT = 4, CNT_ROWS = 5, CNT_SLOTS = 8 };

struct FillParams {
    Cell cell;
    const double* frames;        // [n_frames][A][3] float64, unwrapped allowed
    const long long* frame_list; // optional: process these frame indices only
    long long n_work;            // frames (or list entries) in this launch
    long long frame0;            // global index of frames[0]: row ids / error keys are global
    int A, S, M, L, V, Lpad;
    const int* static_idx;       // [S] atom index of static lattice atom s
    const int* mobile_idx;       // [M]
    const double* ideal;         // [S][3] ideal static positions
    const uint16_t* verts;       // [V][Lpad] vertex table, SoA, VERT_END terminated
    const float* qf;             // [V][Lpad] float(Q): squared-distance cut-off, rounded
    const double* q64;           // [V][Lpad] Q exact: ratio > cutoff  <=>  d^2 > Q
    const double* acoef;         // [V][Lpad] steepness*log2(e)/site_vert_dist
    double bcoef;                // steepness*log2(e)*midpoint
    double static_thr;           // static_movement_threshold
    int dynamic, relaxed;
    unsigned long long* errkey;  // [2] atomicMin of make_error_key: [0] lattice errors, [1] zero landmark vectors
    unsigned long long* counters;
    // MODE_DENSE
    void* dense_out;             // [n_work*M][L]
    int dense_f64;
    // MODE_STATS / MODE_STAGE
    unsigned long long* seen;    // [L]
    double* gram;                // [L][L] upper+lower filled
    __half* stage_hi;            // [Lpad][stage_ld]  (landmark-major: K-major operand for UMMA)
    __half* stage_lo;
    long long stage_ld;
    // MODE_ASSIGN
    const int* cid;              // [L] cluster of landmark, -1 none
    const float* cw;             // [L] centre weight of landmark
    int n_clusters;
    float assign_thr;
    long long* labels;           // [n_work*M] int64, -1 unknown
    double* confs;               // [n_work*M]
    unsigned long long* counts;  // [C] optional bincount of labels
    unsigned long long* best;    // [C] optional max over rows of (|dot| bits << 32 | ~row)
    double* rep;                 // [C][L] optional sum conf*lvec
    double* rep_w;               // [C]
    unsigned long long* site_best; // [C] optional max over rows of (conf bits << 32 | ~row)
};

cudaError_t launch_fill(const FillParams& p, int mode, int n_sms, cudaStream_t stream);

}  // namespace sitb
