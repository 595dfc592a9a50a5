// sitator_b200 -- K1: fused wrap + static-lattice check + landmark-vector fill (+ assign).
// Parameter block shared by the kernel (sitb_fill.cu) and the C-ABI (sitb_api.cu).
#pragma once
#include <cuda_fp16.h>
#include <vector>
#include "sitb_common.cuh"

namespace sitb {

enum FillMode : int {
    MODE_DENSE = 0,   // materialise landmark vectors (helpers._fill_landmark_vectors drop-in)
    MODE_STATS = 1,   // seen counts, zero rows, Gram by sparse outer products (FP64 atomics)
    MODE_STAGE = 2,   // seen counts, zero rows, fp16 hi/lo transposed staging chunk for the tcgen05 SYRK
    MODE_ASSIGN = 3,  // centre similarity, threshold, argmax (+ optional reductions)
};

static constexpr int ENTRY_CAP = 256;        // survivors of the float screen kept per landmark vector; a compressed row holds
                                             // at most 255 entries (8-bit count), more is reported as n_list_overflow
static constexpr int CAND_CAP = 512;         // landmarks screened per block (= candidate list capacity)
static constexpr int MAX_VERTS = 16;

// counters[] slots
enum : int { CNT_ZERO_ROWS = 0, CNT_DUP_NEAREST = 1, CNT_LIST_OVERFLOW = 2, CNT_NNZ = 3,
             CNT_SCREEN_REJECT = 4, CNT_ROWS = 5, CNT_FULL_WALK_FRAMES = 6, CNT_LOOSE_GRID_FRAMES = 7, CNT_SLOTS = 8 };

// Landmark tables (built by sitb_tables.cu).  Vertex ids index the static lattice; a missing
// vertex (the reference's -1 padding) is the dummy id S, whose screen distance is 0 and whose
// bound is +inf, so it passes every screen test.  NB = ceil(V/4) blocks of 4 vertices.
struct LandmarkTables {
    const uint16_t* v0;     // [Lpad] first vertex (SoA copy for the first screen); pad rows: S
    const float* b0;        // [Lpad] its screen bound; pad rows: -1 (never passes)
    const ushort4* va;      // [NB][Lpad] vertices 4*blk .. 4*blk+3
    const float4* ba;       // [NB][Lpad] screen bounds (float, including the FP32 error margin)
    const double* q64;      // [Lpad][4*NB] exact squared cut-off: ratio > cutoff <=> d^2 > q64
    const double* acoef;    // [Lpad][4*NB] steepness/site_vert_dist
    const uint8_t* nverts;  // [Lpad]
    const uint16_t* orig_of;// [Lpad] internal landmark number -> the caller's landmark index
    const ushort4* chunk_atoms;  // [Lpad/32] distinct first-vertex atoms of each 32-landmark chunk (S = none)
    const float4* chunk_bound;   // [Lpad/32] loosest first-vertex bound per such atom (-1 = none)
};

// Host-side image of the tables (sitb_tables.cu: build_landmark_tables).  Landmarks are renumbered
// internally (sorted by first vertex); every output is mapped back through orig_of.
struct HostTables {
    std::vector<uint16_t> v0;
    std::vector<float> b0;
    std::vector<ushort4> va;
    std::vector<float4> ba;
    std::vector<double> q64, acoef;
    std::vector<uint8_t> nverts;
    std::vector<uint16_t> orig_of;
    std::vector<ushort4> chunk_atoms;
    std::vector<float4> chunk_bound;
    std::vector<double> rmax;       // [S] largest cut-off radius of any landmark vertex on the site (-1: none)
    std::vector<int> internal_of;   // [L] caller's index -> internal
};

struct GridLevel {
    const unsigned* ptr;         // [gx*gy*gz + 1] landmark lists
    const uint16_t* list;        // internal landmark ids, ascending within a box
    const unsigned* sptr;        // [gx*gy*gz + 1] the static-lattice sites those landmarks have as vertices
    const uint16_t* slist;
    double margin_sq;
};

struct FillParams {
    Cell cell;
    const double* frames;        // [n_frames][A][3] float64, unwrapped allowed
    const long long* frame_list; // optional: process these frame indices only
    long long n_work;            // frames (or list entries) in this launch
    long long frame0;            // global index of frames[0]: row ids / error keys are global
    int A, S, M, L, V, Lpad, NB;
    unsigned m_magic;            // ceil(2^32 / M) when M < 2^14 (task index -> frame of the batch), else 0
    const int* static_idx;       // [S] atom index of static lattice atom s
    const int* mobile_idx;       // [M]
    const double* ideal;         // [S][3] ideal static positions
    LandmarkTables tab;
    double bcoef;                // steepness*midpoint
    double static_thr;           // static_movement_threshold
    int dynamic, relaxed;
    // candidate lists per box of a gx*gy*gz grid over the (orthorhombic) cell (sitb_tables.cu: k_grid_lists), for up to two
    // margins of static-atom displacement (ascending).  A frame uses the first level whose margin covers its largest
    // static displacement; a frame beyond the last margin walks all landmarks.  n_grid_levels = 0: no grid.
    GridLevel grid[2];
    int n_grid_levels;
    int gx, gy, gz;
    unsigned long long* errkey;  // [2] atomicMin of make_error_key: [0] lattice errors, [1] zero landmark vectors
    unsigned long long* counters;
    // MODE_DENSE
    void* dense_out;             // [n_work*M][L]
    int dense_f64;
    // MODE_STATS / MODE_STAGE
    unsigned long long* seen;    // [L]
    double* gram;                // [L][L] upper triangle (+=)
    __half* stage_hi;            // [Lpad/128][stage_ld/64] pre-swizzled 128x64 tiles (sitb_gram_tc.cu)
    __half* stage_lo;
    long long stage_ld;          // rows the staging buffers hold (multiple of 64)
    // compressed rows (MODE_STATS / MODE_STAGE, optional): row r -> sparse_ptr[r] = offset << 8 | count,
    // entries (caller's landmark index, value) at sparse_k/v[offset ...]; ~0 marks "pool exhausted"
    unsigned long long* sparse_ptr;
    uint16_t* sparse_k;
    double* sparse_v;
    unsigned long long* sparse_cursor;
    unsigned long long sparse_capacity;
    unsigned sparse_slot;            // > 0: rows of at most this many entries live in fixed slots, row-ordered:
    long long sparse_row_base;       //   offset (sparse_row_base + row) * sparse_slot; the cursor only serves longer rows
    // MODE_ASSIGN
    const int* cid;              // [L] cluster of landmark (internal numbering), -1 none
    const double* cw;            // [L] centre weight of landmark (internal numbering)
    int n_clusters;
    double assign_thr;
    long long* labels;           // [n_work*M] int64, -1 unknown
    double* confs;               // [n_work*M]
    unsigned long long* counts;  // [C] optional bincount of labels
    unsigned long long* best;    // [3C] optional max over rows of (|dot|, first row): value bits | row | lock
    double* rep;                 // [C][L] optional sum conf*lvec
    double* rep_w;               // [C]
    unsigned long long* site_best; // [3C] optional max over rows of (conf, first row), same layout
    // second tier of the two-tier assign pass (sitb_fill_fast.cu): redo only the rows the FP32 kernel could not decide
    const unsigned long long* n_work_dev; // optional: number of frame_list entries, read on the device (overrides n_work)
    const uint8_t* row_filter;            // optional [frames][M]: rows with a zero byte are skipped
    int rows_by_frame;                    // with frame_list: outputs are indexed by frame, not by list position
};

cudaError_t launch_fill(const FillParams& p, int mode, int n_sms, cudaStream_t stream);

// ---- two-tier assign pass: FP32 first tier (sitb_fill_fast.cu) -----------------------------------------
// Per landmark (internal numbering): vertices, reciprocal upper cut-off bounds, logistic slopes, centre weight.
struct FastTables {
    const ushort4* va;   // [NB][Lpad] vertex ids (S = the always-passing dummy)
    const float4* ib;    // [NB][Lpad] 1 / (Q * (1 + eps)): q * ib > 1  =>  the exact test d^2 > Q holds too
    const float4* ac;    // [NB][Lpad] steepness * log2(e) / site_vert_dist
    const float2* cw;    // [Lpad] (centre weight, -1 / n_vertices)
};
enum : int { RECHECK_FRAME = 0, RECHECK_SUPPORT = 1, RECHECK_MARGIN = 2, RECHECK_THRESHOLD = 3, RECHECK_LONG = 4, RECHECK_ROWS = 5,
             RECHECK_SLOTS = 8 };
struct FastParams {
    double ci0, ci1, ci2;        // diagonal of cellmat^-1
    float Lx, Ly, Lz;
    const double* frames;        // first frame of the launch
    long long n_work;
    int A, S, M, L, Lpad, NB;
    unsigned m_magic;            // ceil(2^32 / M)
    unsigned sm_magic;           // ceil(2^32 / (S + M))
    const int* static_idx;
    const int* mobile_idx;
    const float4* ideal_frac;    // [S] wrapped fractional static-lattice positions
    FastTables tab;
    float bc;                    // steepness * midpoint * log2(e)
    float kappa;                 // q * ib <= kappa  =>  d^2 <= Q certainly
    float tau;                   // bound on the relative error of an FP32 component value (and of the sums built on it)
    float thr;                   // assignment threshold
    float static_lim_sq;         // frames with a static displacement^2 above this are left to the exact kernel
    float margin_sq[2];          // frames whose static displacements^2 stay below margin_sq[l] may use grid level l
    float dyn_dc;                // float error of a Cartesian component (dynamic lattice map screen)
    int dynamic;
    // candidate grid, both levels in one set of arrays (level l, box b -> entry l * cells + b)
    int cells;
    const uint2* cbox;           // [2 cells] (offset, count) into clist
    const unsigned* clist;       // landmark | cluster << 16, sorted by cluster; landmarks of no cluster left out
    const uint2* sbox;           // [2 cells] (offset, count) into slist
    const uint16_t* slist;       // 4 * static-lattice site, for the sites the box's candidate landmarks use
    int gx, gy, gz;
    float gxf, gyf, gzf;
    long long* labels;           // [n_work*M]
    double* confs;
    unsigned long long* counts;  // [C] optional
    int n_clusters;
    uint8_t* recheck;            // [n_work*M] zeroed by the caller; 1 = row left to the exact kernel
    int* frame_flag;             // [n_work] zeroed by the caller
    long long* frame_list;       // [n_work] frames with at least one such row
    unsigned long long* n_list;  // [1] zeroed by the caller
    unsigned long long* counters;// [RECHECK_SLOTS] reasons (+=)
};
// cudaErrorInvalidConfiguration: shape does not fit the first tier (the caller uses the exact kernel alone)
cudaError_t launch_assign_fast(const FastParams& p, int n_sms, cudaStream_t stream);
// candidate lists of level `level`: the grid's landmark lists sorted by cluster, written at clist + base
cudaError_t launch_sort_box_lists(const unsigned* ptr, const uint16_t* list, const int* cid, long long cells, unsigned base,
                                  uint2* cbox, unsigned* clist, cudaStream_t stream);

}  // namespace sitb
