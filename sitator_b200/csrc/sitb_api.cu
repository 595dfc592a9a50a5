// sitator_b200 -- C ABI (include/sitator_b200.h): context, frame residency, pass launchers.
#include "../../include/sitator_b200.h"
#include "sitb_fill.cuh"

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <new>
#include <vector>

namespace sitb {
cudaError_t launch_tables(const Cell& cell, const double* centers, const double* ideal, const int* verts_in, int L,
                          int V, int S, double cutoff, double* svd_out, double* q_out, cudaStream_t stream);
void build_landmark_tables(const Cell& cell, int L, int V, int Lpad, int NB, int S, double steep_log2e,
                           const int* verts_in, const double* ideal, const double* svd, const double* q,
                           HostTables& out);
cudaError_t launch_grid_static_masks(const Cell& cell, const double* ideal, const double* rmax, int S, int gx, int gy,
                                     int gz, double margin0, double margin1, uint2* masks, unsigned* count,
                                     cudaStream_t stream);
cudaError_t launch_grid_masks(const Cell& cell, const double* ideal, const ushort4* va, const double* q64, int L,
                              int Lpad, int NB, int S, int gx, int gy, int gz, double margin0, double margin1, uint2* masks,
                              unsigned* count, cudaStream_t stream);
cudaError_t launch_grid_lists_from_masks(const uint2* masks, int n_chunks, long long cells, const unsigned* ptr0,
                                         const unsigned* ptr1, uint16_t* list0, uint16_t* list1, uint16_t* cat_list,
                                         uint2* cat_box, cudaStream_t stream);
}

using namespace sitb;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace sitb {
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace sitb

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(SITB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct GridLevelDev {
    unsigned* ptr = nullptr;          // [cells + 1] landmark lists
    uint16_t* list = nullptr;
    unsigned* sptr = nullptr;         // [cells + 1] static-site lists
    uint16_t* slist = nullptr;
    double margin = 0.0;
    unsigned long long entries = 0, static_entries = 0;
};                                    // (the four arrays of both levels live in one block: sitb_ctx::d_grid_block)

struct sitb_ctx {
    int device = 0;
    int n_sms = 0, cc_major = 0, cc_minor = 0;
    int A = 0, S = 0, M = 0, L = 0, V = 0, Lpad = 0, NB = 1;
    Cell cell;
    double midpoint = 0, steepness = 0, cutoff = 0, static_thr = 0, bcoef = 0;
    int dynamic = 0, relaxed = 0;
    cudaStream_t stream = 0;
    // device tables
    int* d_static_idx = nullptr;
    int* d_mobile_idx = nullptr;
    double* d_ideal = nullptr;
    double* d_centers = nullptr;
    int* d_verts_in = nullptr;
    double* d_svd = nullptr;
    double* d_qorig = nullptr;
    uint16_t* d_orig_of = nullptr;
    ushort4* d_chunk_atoms = nullptr;
    float4* d_chunk_bound = nullptr;
    std::vector<int> internal_of;     // caller's landmark index -> internal
    std::vector<double> h_svd, h_qorig;
    uint16_t* d_v0 = nullptr;
    float* d_b0 = nullptr;
    ushort4* d_va = nullptr;
    float4* d_ba = nullptr;
    double* d_q64 = nullptr;
    double* d_acoef = nullptr;
    uint8_t* d_nverts = nullptr;
    // candidate grid (orthorhombic cells)
    GridLevelDev grid[2];             // [0] half the margin, [1] the margin
    unsigned char* d_grid_block = nullptr;   // every list of both levels + the first tier's site lists, one allocation
    int n_grid_levels = 0;
    double* d_rmax = nullptr;
    double* d_ideal_wrapped = nullptr;   // static-lattice positions wrapped into the cell (grid builder)
    double* d_radius = nullptr;          // [Lpad][4NB] cut-off radii sqrt(Q) (grid builder; < 0: degenerate)
    int gx = 0, gy = 0, gz = 0;
    // centres
    int* d_cid = nullptr;
    double* d_cw = nullptr;
    int* d_cid_orig = nullptr;        // the same tables in the caller's landmark numbering (sparse-row passes)
    double* d_cw_orig = nullptr;
    int n_clusters = 0;
    // frames
    const double* d_frames = nullptr;
    double* d_frames_owned = nullptr;
    size_t frames_capacity = 0;      // bytes
    float* d_frames_f32 = nullptr;   // staging for float32 sources (converted to float64 chunk by chunk)
    size_t f32_capacity = 0;
    long long n_frames = 0, frame0 = 0;
    // the upload runs chunk by chunk on its own stream; a pass waits only for the chunks it reads, so the
    // first pass over a fresh trajectory overlaps the host -> device copy
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t order_event = nullptr;
    std::vector<cudaEvent_t> up_events;   // one per chunk of up_chunk frames
    long long up_chunk = 0;
    size_t up_waited = 0;                 // chunks the compute stream already waits for
    // status
    unsigned long long* d_status = nullptr;   // [2] error keys + [CNT_SLOTS] counters
    // two-tier assign pass (sitb_fill_fast.cu): float tables, error model, scratch
    float4* d_fast_ib = nullptr;
    float4* d_fast_ac = nullptr;
    float2* d_fast_cw = nullptr;
    float4* d_ideal_frac = nullptr;
    uint2* d_fast_sbox = nullptr;             // [2 cells] both grid levels: (offset, count) into d_fast_slist
    uint16_t* d_fast_slist = nullptr;         // 4 * site
    uint2* d_fast_cbox = nullptr;             // [2 cells] (offset, count) into d_fast_clist (rebuilt when the centres change)
    unsigned* d_fast_clist = nullptr;         // landmark | cluster << 16, sorted by cluster
    std::vector<uint8_t> h_nverts;            // internal numbering
    double fast_kappa = 0.0, fast_tau = 0.0, fast_dc = 0.0, cw_absmax = 0.0;
    bool fast_tables_ok = false, fast_lists_dirty = true;
    uint8_t* d_recheck = nullptr;             // [rows]
    int* d_frame_flag = nullptr;              // [frames]
    long long* d_frame_list = nullptr;        // [frames]
    unsigned long long* d_two_tier = nullptr; // [0] list length, [1 .. RECHECK_SLOTS] reason counters
    size_t recheck_rows_cap = 0, recheck_frames_cap = 0;
    int assign_mode = 0;                      // 0 exact, 1 two-tier where the shape allows it
    // the last sitb_pass_stats_slotted: rows in these buffers live in row-ordered slots of row_slot entries (the passes
    // over cached rows then load the entries without waiting for the row pointers)
    const void* slot_pool_k = nullptr;
    const void* slot_row_ptr = nullptr;       // row 0 of the shard
    int row_slot = 0;
};

static void free_ctx(sitb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->order_event) cudaEventDestroy(c->order_event);
    for (cudaEvent_t ev : c->up_events) cudaEventDestroy(ev);
    pool_free(c->d_static_idx, c->stream); pool_free(c->d_mobile_idx, c->stream); pool_free(c->d_ideal, c->stream); pool_free(c->d_centers, c->stream);
    pool_free(c->d_chunk_atoms, c->stream); pool_free(c->d_chunk_bound, c->stream);
    pool_free(c->d_grid_block, c->stream);
    pool_free(c->d_fast_cbox, c->stream); pool_free(c->d_fast_clist, c->stream);
    pool_free(c->d_fast_ib, c->stream); pool_free(c->d_fast_ac, c->stream); pool_free(c->d_fast_cw, c->stream);
    pool_free(c->d_ideal_frac, c->stream); pool_free(c->d_recheck, c->stream); pool_free(c->d_frame_flag, c->stream);
    pool_free(c->d_frame_list, c->stream); pool_free(c->d_two_tier, c->stream);
    pool_free(c->d_rmax, c->stream); pool_free(c->d_ideal_wrapped, c->stream); pool_free(c->d_radius, c->stream);
    pool_free(c->d_verts_in, c->stream); pool_free(c->d_svd, c->stream); pool_free(c->d_qorig, c->stream); pool_free(c->d_orig_of, c->stream); pool_free(c->d_v0, c->stream); pool_free(c->d_b0, c->stream); pool_free(c->d_va, c->stream); pool_free(c->d_ba, c->stream);
    pool_free(c->d_q64, c->stream); pool_free(c->d_acoef, c->stream); pool_free(c->d_nverts, c->stream); pool_free(c->d_cid, c->stream); pool_free(c->d_cw, c->stream); pool_free(c->d_cid_orig, c->stream); pool_free(c->d_cw_orig, c->stream); pool_free(c->d_frames_owned, c->stream); pool_free(c->d_frames_f32, c->stream); pool_free(c->d_status, c->stream);
    delete c;
}

template <typename T>
static cudaError_t upload(T** dst, const T* src, size_t n, cudaStream_t st) {
    cudaError_t e = pool_alloc((void**)dst, sizeof(T) * n, st);
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // src may be a temporary
    return e;
}

// the same without the wait: the caller keeps src alive until it has synchronised the stream
template <typename T>
static cudaError_t upload_async(T** dst, const T* src, size_t n, cudaStream_t st) {
    cudaError_t e = pool_alloc((void**)dst, sizeof(T) * n, st);
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, st);
    return e;
}

static int build_grid(sitb_ctx* c, double margin);

// developer aid: SITB_TIMING=1 prints the host wall time of the stages of sitb_create / the grid build to stderr
struct StageClock {
    bool on;
    std::chrono::steady_clock::time_point t;
    StageClock() : on(getenv("SITB_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[sitb timing] %-44s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};
extern "C" int sitb_reset_status(sitb_ctx* c);

extern "C" const char* sitb_last_error(void) { return g_err; }
extern "C" int sitb_version(void) { return 200; }
extern "C" int sitb_abi_sizes(uint64_t* out) {
    if (!out) return fail(SITB_E_INVALID, "sitb_abi_sizes: null argument");
    out[0] = sizeof(sitb_network_desc);
    out[1] = sizeof(sitb_status);
    return SITB_OK;
}

extern "C" int sitb_create(const sitb_network_desc* d, int device, sitb_ctx** out) {
    if (!d || !out) return fail(SITB_E_INVALID, "sitb_create: null argument");
    *out = nullptr;
    if (d->n_atoms <= 0 || d->n_static <= 0 || d->n_mobile <= 0 || d->n_landmarks <= 0 || d->max_verts <= 0)
        return fail(SITB_E_INVALID, "sitb_create: sizes must be positive");
    if (d->max_verts > MAX_VERTS) return fail(SITB_E_LIMIT, "sitb_create: max_verts %d > %d", d->max_verts, MAX_VERTS);
    if (d->n_static >= 65535) return fail(SITB_E_LIMIT, "sitb_create: n_static %d >= 65535", d->n_static);
    if (d->n_landmarks >= 65535) return fail(SITB_E_LIMIT, "sitb_create: n_landmarks %d >= 65535", d->n_landmarks);
    if (!d->host_cellmat || !d->host_static_idx || !d->host_mobile_idx || !d->host_ideal_static ||
        !d->host_centers || !d->host_verts)
        return fail(SITB_E_INVALID, "sitb_create: null table pointer");
    for (int i = 0; i < d->n_static; ++i)
        if (d->host_static_idx[i] < 0 || d->host_static_idx[i] >= d->n_atoms)
            return fail(SITB_E_INVALID, "sitb_create: static_idx[%d] out of range", i);
    for (int i = 0; i < d->n_mobile; ++i)
        if (d->host_mobile_idx[i] < 0 || d->host_mobile_idx[i] >= d->n_atoms)
            return fail(SITB_E_INVALID, "sitb_create: mobile_idx[%d] out of range", i);
    for (int k = 0; k < d->n_landmarks; ++k) {
        if (d->host_verts[(size_t)k * d->max_verts] < 0)
            return fail(SITB_E_INVALID, "sitb_create: landmark %d has no vertices", k);
        for (int h = 0; h < d->max_verts; ++h)
            if (d->host_verts[(size_t)k * d->max_verts + h] >= d->n_static)
                return fail(SITB_E_INVALID, "sitb_create: verts[%d][%d] out of range", k, h);
    }
    StageClock clk;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SITB_E_CUDA, "sitb_create: no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    if (device < 0 || device >= ndev) return fail(SITB_E_INVALID, "sitb_create: device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    sitb_ctx* c = new (std::nothrow) sitb_ctx();
    if (!c) return fail(SITB_E_INVALID, "out of host memory");
    c->device = device;
    // three attribute reads, not cudaGetDeviceProperties (which also queries clocks and PCI state and can take
    // tens of milliseconds on a virtualised GPU)
    e = cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->cc_major, cudaDevAttrComputeCapabilityMajor, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->cc_minor, cudaDevAttrComputeCapabilityMinor, device);
    if (e != cudaSuccess) { free_ctx(c); return fail(SITB_E_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); }
    c->A = d->n_atoms; c->S = d->n_static; c->M = d->n_mobile; c->L = d->n_landmarks; c->V = d->max_verts;
    c->Lpad = (c->L + 255) & ~255;   // K1 screens landmarks in unrolled groups of 8 x 32
    c->NB = (c->V + 3) / 4;
    // cell
    for (int i = 0; i < 9; ++i) c->cell.c[i] = d->host_cellmat[i];
    if (d->host_cellmat_inv) {
        for (int i = 0; i < 9; ++i) c->cell.ci[i] = d->host_cellmat_inv[i];
    } else {
        const double* m = c->cell.c;
        const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                           m[2] * (m[3] * m[7] - m[4] * m[6]);
        if (det == 0.0) { free_ctx(c); return fail(SITB_E_INVALID, "sitb_create: singular cell"); }
        double* o = c->cell.ci;
        o[0] = (m[4] * m[8] - m[5] * m[7]) / det; o[1] = (m[2] * m[7] - m[1] * m[8]) / det; o[2] = (m[1] * m[5] - m[2] * m[4]) / det;
        o[3] = (m[5] * m[6] - m[3] * m[8]) / det; o[4] = (m[0] * m[8] - m[2] * m[6]) / det; o[5] = (m[2] * m[3] - m[0] * m[5]) / det;
        o[6] = (m[3] * m[7] - m[4] * m[6]) / det; o[7] = (m[1] * m[6] - m[0] * m[7]) / det; o[8] = (m[0] * m[4] - m[1] * m[3]) / det;
    }
    // centroid = sum over cell vectors of 0.5*cell (PBCCalculator.pyx:35): cell rows = cellmat columns
    for (int k = 0; k < 3; ++k) {
        // np.sum(0.5*cell, axis=0)[k] = ((0.5*cell[0][k] + 0.5*cell[1][k]) + 0.5*cell[2][k]); cell[i][k] = cellmat[k][i]
        c->cell.cen[k] = (0.5 * c->cell.c[3 * k + 0] + 0.5 * c->cell.c[3 * k + 1]) + 0.5 * c->cell.c[3 * k + 2];
    }
    bool diag = true;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (i != j && (c->cell.c[3 * i + j] != 0.0 || c->cell.ci[3 * i + j] != 0.0)) diag = false;
    c->cell.diag = diag ? 1 : 0;
    c->midpoint = d->cutoff_midpoint; c->steepness = d->cutoff_steepness;
    c->cutoff = d->cutoff_round_to_zero > 0.0 ? d->cutoff_round_to_zero
                                              : d->cutoff_midpoint + std::log((1.0 / 0.0001) - 1.0) / d->cutoff_steepness;
    c->static_thr = d->static_movement_threshold;
    c->dynamic = d->dynamic_lattice_mapping; c->relaxed = d->relaxed_lattice_checks;
    const double steep_log2e = c->steepness;          // natural-exponent units (values are evaluated in double)
    c->bcoef = c->steepness * c->midpoint;

    const size_t LV = (size_t)c->L * c->V;
    clk.lap("create: validation, device, cell");
#define CKC(call)                                                                          \
    do {                                                                                   \
        cudaError_t e2_ = (call);                                                          \
        if (e2_ != cudaSuccess) {                                                          \
            free_ctx(c);                                                                   \
            return fail(SITB_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2_));     \
        }                                                                                  \
    } while (0)
    CKC(upload_async(&c->d_static_idx, d->host_static_idx, (size_t)c->S, c->stream));
    CKC(upload_async(&c->d_mobile_idx, d->host_mobile_idx, (size_t)c->M, c->stream));
    CKC(upload_async(&c->d_ideal, d->host_ideal_static, (size_t)c->S * 3, c->stream));
    CKC(upload_async(&c->d_centers, d->host_centers, (size_t)c->L * 3, c->stream));
    CKC(upload_async(&c->d_verts_in, d->host_verts, LV, c->stream));
    CKC(pool_alloc((void**)&c->d_svd, sizeof(double) * LV, c->stream));
    CKC(pool_alloc((void**)&c->d_qorig, sizeof(double) * LV, c->stream));
    CKC(pool_alloc((void**)&c->d_cid, sizeof(int) * (size_t)c->L, c->stream));
    CKC(pool_alloc((void**)&c->d_cw, sizeof(double) * (size_t)c->L, c->stream));
    CKC(cudaMemset(c->d_cid, 0xFF, sizeof(int) * (size_t)c->L));
    CKC(cudaMemset(c->d_cw, 0, sizeof(double) * (size_t)c->L));
    CKC(pool_alloc((void**)&c->d_status, sizeof(unsigned long long) * (2 + CNT_SLOTS), c->stream));
    CKC(launch_tables(c->cell, c->d_centers, c->d_ideal, c->d_verts_in, c->L, c->V, c->S, c->cutoff, c->d_svd,
                      c->d_qorig, 0));
    c->h_svd.resize(LV); c->h_qorig.resize(LV);
    CKC(cudaMemcpy(c->h_svd.data(), c->d_svd, sizeof(double) * LV, cudaMemcpyDeviceToHost));
    CKC(cudaMemcpy(c->h_qorig.data(), c->d_qorig, sizeof(double) * LV, cudaMemcpyDeviceToHost));
    clk.lap("create: uploads, site-vertex distances, D2H");
    {
        HostTables ht;
        build_landmark_tables(c->cell, c->L, c->V, c->Lpad, c->NB, c->S, steep_log2e, d->host_verts,
                              d->host_ideal_static, c->h_svd.data(), c->h_qorig.data(), ht);
        clk.lap("create: host landmark tables");
        CKC(upload_async(&c->d_chunk_atoms, ht.chunk_atoms.data(), ht.chunk_atoms.size(), c->stream));
        CKC(upload_async(&c->d_chunk_bound, ht.chunk_bound.data(), ht.chunk_bound.size(), c->stream));
        CKC(upload_async(&c->d_rmax, ht.rmax.data(), ht.rmax.size(), c->stream));
        {
            // what the grid builder reads per (box, landmark, vertex): wrapped Cartesian positions and radii, once
            std::vector<double> iw((size_t)c->S * 3, 0.0), rad(ht.q64.size());
            if (c->cell.diag)
                for (int s = 0; s < c->S; ++s)
                    for (int k = 0; k < 3; ++k) {
                        double f = c->cell.ci[4 * k] * d->host_ideal_static[3 * s + k];
                        f -= std::floor(f);
                        iw[(size_t)s * 3 + k] = f * c->cell.c[4 * k];
                    }
            for (size_t i = 0; i < rad.size(); ++i) rad[i] = (ht.q64[i] >= 0.0) ? std::sqrt(ht.q64[i]) : -1.0;
            CKC(upload_async(&c->d_ideal_wrapped, iw.data(), iw.size(), c->stream));
            CKC(upload_async(&c->d_radius, rad.data(), rad.size(), c->stream));
            CKC(cudaStreamSynchronize(c->stream));                 // iw, rad go out of scope
        }
        c->internal_of = ht.internal_of;
        CKC(upload_async(&c->d_v0, ht.v0.data(), ht.v0.size(), c->stream));
        CKC(upload_async(&c->d_b0, ht.b0.data(), ht.b0.size(), c->stream));
        CKC(upload_async(&c->d_va, ht.va.data(), ht.va.size(), c->stream));
        CKC(upload_async(&c->d_ba, ht.ba.data(), ht.ba.size(), c->stream));
        CKC(upload_async(&c->d_q64, ht.q64.data(), ht.q64.size(), c->stream));
        CKC(upload_async(&c->d_acoef, ht.acoef.data(), ht.acoef.size(), c->stream));
        CKC(upload_async(&c->d_nverts, ht.nverts.data(), ht.nverts.size(), c->stream));
        CKC(upload_async(&c->d_orig_of, ht.orig_of.data(), ht.orig_of.size(), c->stream));
        c->h_nverts = ht.nverts;
        clk.lap("create: table uploads");
        if (c->cell.diag) {
            // ---- float tables and error model of the two-tier assign pass (sitb_fill_fast.cu) ----
            // Squared distances are formed in FP32 from float fractional coordinates as |(u - round(u)) L|^2.  With
            // u24 = 2^-24: a fractional coordinate carries 2^-25 from its rounding to float, their difference one more
            // rounding, the product with float(L) two; each Cartesian component is off by at most 2 u24 L (3 allowed).
            const double u24 = 5.9604644775390625e-08;
            const double lmax = std::max(std::fabs(c->cell.c[0]), std::max(std::fabs(c->cell.c[4]), std::fabs(c->cell.c[8])));
            const double dc = 3.0 * u24 * lmax;
            const double log2e = 1.4426950408889634074, ln2 = 0.69314718055994530942;
            const double bcl = c->steepness * c->midpoint * log2e;
            const int W = 4 * c->NB;
            std::vector<float4> ib((size_t)c->NB * c->Lpad, make_float4(0.f, 0.f, 0.f, 0.f)), ac(ib);
            double eps_max = 0.0, tau_max = 0.0;
            for (int ki = 0; ki < c->L; ++ki)
                for (int h = 0; h < W; ++h) {
                    const double Q = ht.q64[(size_t)ki * W + h];
                    const double a = ht.acoef[(size_t)ki * W + h] * log2e;     // steepness * log2(e) / site_vert_dist
                    if (!(Q > 0.0) || std::isinf(Q) || !(a > 0.0)) continue;    // dummy vertex / degenerate landmark: (0, 0)
                    const double err = 2.0 * std::sqrt(3.0 * Q) * dc + 3.0 * dc * dc + 4.0 * u24 * Q;
                    const double eps = 1.5 * err / Q + 8.0 * u24;
                    eps_max = std::max(eps_max, eps);
                    const double R = std::sqrt(Q);
                    const double dd = std::sqrt(3.0) * dc + 6.0 * u24 * R;       // |d_float - d|
                    tau_max = std::max(tau_max, ln2 * (a * dd + (a * R + bcl) * 3.0 * u24));
                    float* pi = &ib[(size_t)(h / 4) * c->Lpad + ki].x;
                    float* pa = &ac[(size_t)(h / 4) * c->Lpad + ki].x;
                    pi[h & 3] = (float)(1.0 / (Q * (1.0 + eps)));
                    pa[h & 3] = (float)a;
                }
            c->fast_kappa = (1.0 - eps_max) / (1.0 + eps_max) * (1.0 - 4.0 * u24);
            // + lg2.approx (2^-22 relative on |log2 P| <= 54), ex2.approx, product and summation roundings
            c->fast_tau = 1.25 * (tau_max + ln2 * 54.0 * 4.0 * u24 + 16.0 * u24) + 16.0 * u24;
            c->fast_dc = dc;
            std::vector<float4> idf((size_t)c->S);
            for (int s2 = 0; s2 < c->S; ++s2) {
                float fr[3];
                for (int k = 0; k < 3; ++k) {
                    double f = c->cell.ci[4 * k] * d->host_ideal_static[3 * s2 + k];
                    f -= std::floor(f);
                    fr[k] = (float)f;
                }
                idf[s2] = make_float4(fr[0], fr[1], fr[2], 0.f);
            }
            CKC(upload_async(&c->d_fast_ib, ib.data(), ib.size(), c->stream));
            CKC(upload_async(&c->d_fast_ac, ac.data(), ac.size(), c->stream));
            CKC(upload_async(&c->d_ideal_frac, idf.data(), idf.size(), c->stream));
            CKC(pool_alloc((void**)&c->d_fast_cw, sizeof(float2) * (size_t)c->Lpad, c->stream));
            CKC(pool_alloc((void**)&c->d_two_tier, sizeof(unsigned long long) * (1 + RECHECK_SLOTS), c->stream));
            CKC(cudaMemsetAsync(c->d_two_tier, 0, sizeof(unsigned long long) * (1 + RECHECK_SLOTS), c->stream));
            c->fast_tables_ok = c->fast_kappa > 0.9 && c->fast_tau < 1e-2;
        }
        CKC(cudaStreamSynchronize(c->stream));                     // ht and the float tables go out of scope
    }
    clk.lap("create: float tables of the first tier");
#undef CKC
    *out = c;
    int rc = sitb_reset_status(c);
    if (rc == SITB_OK) {
        // default margin 0.5 A (thermal displacements of a static lattice are a few tenths of an Angstrom);
        // SITB_GRID_MARGIN overrides it, <= 0 disables the grid
        double margin = 0.5;
        if (const char* env = getenv("SITB_GRID_MARGIN")) margin = atof(env);
        rc = build_grid(c, margin);
        clk.lap("create: candidate grid (total)");
    }
    if (rc != SITB_OK) { free_ctx(c); *out = nullptr; return rc; }
    return SITB_OK;
}

static void free_grid(sitb_ctx* c) {
    pool_free(c->d_grid_block, c->stream);
    c->d_grid_block = nullptr;
    for (int l = 0; l < 2; ++l) c->grid[l] = GridLevelDev();
    pool_free(c->d_fast_cbox, c->stream); pool_free(c->d_fast_clist, c->stream);
    c->d_fast_sbox = nullptr; c->d_fast_slist = nullptr; c->d_fast_cbox = nullptr; c->d_fast_clist = nullptr;
    c->fast_lists_dirty = true;
    c->n_grid_levels = 0;
    c->gx = c->gy = c->gz = 0;
}

// (Re)build the candidate grid for a static-atom margin (Angstrom); margin <= 0 or a triclinic cell: no grid.
// Two levels: lists for half the margin (what a quiet lattice needs: fewer candidates) and for the margin itself;
// K1 picks per frame the tightest level that covers the frame's largest static displacement.
static int build_grid(sitb_ctx* c, double margin) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_grid(c);
    if (!(margin > 0.0) || !c->cell.diag) return SITB_OK;
    const double len[3] = {std::fabs(c->cell.c[0]), std::fabs(c->cell.c[4]), std::fabs(c->cell.c[8])};
    if (c->cell.c[0] <= 0.0 || c->cell.c[4] <= 0.0 || c->cell.c[8] <= 0.0) return SITB_OK;
    // boxes of ~0.5 A (a seventh of a cut-off radius), at most 64 per axis and ~2e8 (box, landmark) tests
    double side = 0.5;
    if (const char* env = getenv("SITB_GRID_BOX")) { const double v = atof(env); if (v > 0.05) side = v; }   // developer knob
    int g[3];
    for (;;) {
        double cells = 1.0;
        for (int d = 0; d < 3; ++d) {
            int n = (int)std::floor(len[d] / side + 0.5);
            g[d] = n < 1 ? 1 : (n > 64 ? 64 : n);
            cells *= g[d];
        }
        if (cells * (double)c->L <= 2.0e8 || (g[0] == 1 && g[1] == 1 && g[2] == 1)) break;
        side *= 1.26;
    }
    // list margin = the bound on static displacements + the float rounding of the box lookup in K1
    const double lmax = std::max(len[0], std::max(len[1], len[2]));
    const double eps = 1e-5 * lmax + 1e-9;
    const double margins[2] = {0.5 * margin, margin};
    // Both levels in one sweep (sitb_tables.cu): a ballot mask per (box, 32 landmarks) and margin, the list lengths to the
    // host for the prefix sums (the allocation needs the totals), then the lists are written from the masks -- no second
    // round of distance tests, one device block, one round trip.
    StageClock clk;
    const size_t cells = (size_t)g[0] * g[1] * g[2];
    const int chL = (c->L + 31) / 32, chS = (c->S + 31) / 32;
    unsigned char* scratch = nullptr;
    const size_t mask_bytes = sizeof(uint2) * cells * (size_t)(chL + chS);
    CK(pool_alloc((void**)&scratch, mask_bytes + sizeof(unsigned) * 4 * cells, c->stream));
    uint2* masksL = (uint2*)scratch;
    uint2* masksS = masksL + cells * (size_t)chL;
    unsigned* d_count = (unsigned*)(scratch + mask_bytes);         // [0, 2 cells): landmarks, levels 0 / 1; then sites
    cudaError_t e = launch_grid_masks(c->cell, c->d_ideal_wrapped, c->d_va, c->d_radius, c->L, c->Lpad, c->NB, c->S, g[0], g[1], g[2],
                                      margins[0] + eps, margins[1] + eps, masksL, d_count, c->stream);
    if (e == cudaSuccess)
        e = launch_grid_static_masks(c->cell, c->d_ideal_wrapped, c->d_rmax, c->S, g[0], g[1], g[2], margins[0] + eps,
                                     margins[1] + eps, masksS, d_count + 2 * cells, c->stream);
    std::vector<unsigned> ptr(4 * (cells + 1), 0u);                // four prefix arrays: ptr0, ptr1, sptr0, sptr1
    {
        std::vector<unsigned> cnt(4 * cells);
        if (e == cudaSuccess) e = cudaMemcpyAsync(cnt.data(), d_count, sizeof(unsigned) * 4 * cells, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { pool_free(scratch, c->stream); return fail(SITB_E_CUDA, "candidate grid (count): %s", cudaGetErrorString(e)); }
        clk.lap("  grid: masks of both levels + counts D2H");
        unsigned long long totals[4];
        for (int a = 0; a < 4; ++a) {
            unsigned long long total = 0;
            unsigned* pa = ptr.data() + (size_t)a * (cells + 1);
            for (size_t i = 0; i < cells; ++i) { pa[i] = (unsigned)total; total += cnt[(size_t)a * cells + i]; }
            pa[cells] = (unsigned)total;
            totals[a] = total;
        }
        if (totals[0] + totals[1] >= 0xFFFFFFFFull || totals[2] + totals[3] >= 0xFFFFFFFFull) {
            pool_free(scratch, c->stream);
            return SITB_OK;                                        // absurdly large: keep the full walk
        }
        c->grid[0].entries = totals[0]; c->grid[1].entries = totals[1];
        c->grid[0].static_entries = totals[2]; c->grid[1].static_entries = totals[3];
    }
    const size_t nL0 = (size_t)c->grid[0].entries, nL1 = (size_t)c->grid[1].entries;
    const size_t nS0 = (size_t)c->grid[0].static_entries, nS1 = (size_t)c->grid[1].static_entries;
    const bool first_tier = c->fast_tables_ok && c->S <= 16000;
    // block layout: 4 prefix arrays | (offset, count) boxes of the first tier | the lists (uint16)
    size_t o = 0;
    const size_t off_ptr = o;   o += sizeof(unsigned) * 4 * (cells + 1);
    o = (o + 7) & ~(size_t)7;
    const size_t off_sbox = o;  o += first_tier ? sizeof(uint2) * 2 * cells : 0;
    const size_t off_list = o;  o += sizeof(uint16_t) * (nL0 + nL1 + nS0 + nS1 + (first_tier ? nS0 + nS1 : 0) + 8);
    CK(pool_alloc((void**)&c->d_grid_block, o, c->stream));
    unsigned* d_ptr = (unsigned*)(c->d_grid_block + off_ptr);
    uint16_t* d_lists = (uint16_t*)(c->d_grid_block + off_list);
    CK(cudaMemcpyAsync(d_ptr, ptr.data(), sizeof(unsigned) * ptr.size(), cudaMemcpyHostToDevice, c->stream));
    c->grid[0].ptr = d_ptr;                       c->grid[1].ptr = d_ptr + (cells + 1);
    c->grid[0].sptr = d_ptr + 2 * (cells + 1);    c->grid[1].sptr = d_ptr + 3 * (cells + 1);
    c->grid[0].list = d_lists;                    c->grid[1].list = d_lists + nL0;
    c->grid[0].slist = d_lists + nL0 + nL1;       c->grid[1].slist = d_lists + nL0 + nL1 + nS0;
    if (first_tier) {
        // first tier of the two-tier assign pass: both levels' static-site lists in one array (entries = 4 * site)
        c->d_fast_slist = d_lists + nL0 + nL1 + nS0 + nS1;
        c->d_fast_sbox = (uint2*)(c->d_grid_block + off_sbox);
    }
    CK(launch_grid_lists_from_masks(masksL, chL, (long long)cells, c->grid[0].ptr, c->grid[1].ptr, c->grid[0].list, c->grid[1].list,
                                    nullptr, nullptr, c->stream));
    CK(launch_grid_lists_from_masks(masksS, chS, (long long)cells, c->grid[0].sptr, c->grid[1].sptr, c->grid[0].slist,
                                    c->grid[1].slist, c->d_fast_slist, c->d_fast_sbox, c->stream));
    pool_free(scratch, c->stream);
    CK(cudaStreamSynchronize(c->stream));          // (ptr is a temporary)
    for (int l = 0; l < 2; ++l) c->grid[l].margin = margins[l];
    c->gx = g[0]; c->gy = g[1]; c->gz = g[2];
    c->n_grid_levels = 2;
    if (first_tier) {
        CK(pool_alloc((void**)&c->d_fast_cbox, sizeof(uint2) * 2 * cells, c->stream));
        CK(pool_alloc((void**)&c->d_fast_clist, sizeof(unsigned) * (nL0 + nL1 + 1), c->stream));
        c->fast_lists_dirty = true;
    }
    clk.lap("  grid: prefix sums, lists from the masks");
    return SITB_OK;
}

extern "C" int sitb_set_candidate_grid(sitb_ctx* c, double static_margin) {
    if (!c) return fail(SITB_E_INVALID, "null context");
    return build_grid(c, static_margin);
}

extern "C" int sitb_candidate_grid_info(sitb_ctx* c, int32_t* dims, double* static_margin, uint64_t* n_entries) {
    if (!c) return fail(SITB_E_INVALID, "null context");
    if (dims) { dims[0] = c->gx; dims[1] = c->gy; dims[2] = c->gz; }
    if (static_margin) *static_margin = c->n_grid_levels ? c->grid[c->n_grid_levels - 1].margin : 0.0;
    if (n_entries) *n_entries = c->grid[0].entries + c->grid[1].entries;
    return SITB_OK;
}

extern "C" void sitb_destroy(sitb_ctx* ctx) { free_ctx(ctx); }

extern "C" int sitb_set_stream(sitb_ctx* c, void* s) {
    if (!c) return fail(SITB_E_INVALID, "null context");
    c->stream = (cudaStream_t)s;
    c->up_waited = 0;                 // a new compute stream has to wait for the upload again
    return SITB_OK;
}

extern "C" int sitb_device_info(sitb_ctx* c, int32_t* n_sms, int32_t* major, int32_t* minor) {
    if (!c) return fail(SITB_E_INVALID, "null context");
    if (n_sms) *n_sms = c->n_sms;
    if (major) *major = c->cc_major;
    if (minor) *minor = c->cc_minor;
    return SITB_OK;
}

extern "C" int sitb_get_tables(sitb_ctx* c, double* svd, double* q) {
    if (!c) return fail(SITB_E_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    const size_t LV = (size_t)c->L * c->V;
    if (svd) memcpy(svd, c->h_svd.data(), sizeof(double) * LV);
    if (q) memcpy(q, c->h_qorig.data(), sizeof(double) * LV);
    return SITB_OK;
}

extern "C" int sitb_upload_frames(sitb_ctx* c, const double* host, int64_t n, int64_t frame0) {
    if (!c || !host || n <= 0) return fail(SITB_E_INVALID, "sitb_upload_frames: bad argument");
    CK(cudaSetDevice(c->device));
    const size_t frame_bytes = sizeof(double) * (size_t)c->A * 3;
    const size_t bytes = frame_bytes * (size_t)n;
    if (bytes > c->frames_capacity) {
        pool_free(c->d_frames_owned, c->stream);
        c->d_frames_owned = nullptr; c->frames_capacity = 0;
        CK(pool_alloc((void**)&c->d_frames_owned, bytes, c->stream));
        c->frames_capacity = bytes;
    }
    if (!c->copy_stream) {
        CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->order_event, cudaEventDisableTiming));
    }
    // the copy may start once everything queued on the compute stream (allocation, older passes) is done
    CK(cudaEventRecord(c->order_event, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->order_event, 0));
    // host memory should be page-locked (then the copies are asynchronous; the caller keeps it alive and
    // unchanged until the passes that read it have run); pageable memory works but copies synchronously
    c->up_chunk = (long long)((32ull << 20) / frame_bytes);
    if (c->up_chunk < 1) c->up_chunk = 1;
    const size_t n_chunks = (size_t)((n + c->up_chunk - 1) / c->up_chunk);
    while (c->up_events.size() < n_chunks) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->up_events.push_back(ev);
    }
    for (size_t k = 0; k < n_chunks; ++k) {
        const long long f0 = (long long)k * c->up_chunk;
        const long long nf = (n - f0 < c->up_chunk) ? (n - f0) : c->up_chunk;
        CK(cudaMemcpyAsync(c->d_frames_owned + (size_t)f0 * c->A * 3, host + (size_t)f0 * c->A * 3, frame_bytes * (size_t)nf,
                           cudaMemcpyHostToDevice, c->copy_stream));
        CK(cudaEventRecord(c->up_events[k], c->copy_stream));
    }
    c->up_waited = 0;
    c->d_frames = c->d_frames_owned; c->n_frames = n; c->frame0 = frame0;
    return SITB_OK;
}

__global__ void k_widen_frames(const float* __restrict__ in, double* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

// float32 trajectories (what most MD codes write): half the PCIe traffic; widened to float64 on the device, chunk by
// chunk on the copy stream, so every result equals a run on frames.astype(float64).
extern "C" int sitb_upload_frames_f32(sitb_ctx* c, const float* host, int64_t n, int64_t frame0) {
    if (!c || !host || n <= 0) return fail(SITB_E_INVALID, "sitb_upload_frames_f32: bad argument");
    CK(cudaSetDevice(c->device));
    const size_t frame_elems = (size_t)c->A * 3;
    const size_t bytes64 = sizeof(double) * frame_elems * (size_t)n, bytes32 = sizeof(float) * frame_elems * (size_t)n;
    if (bytes64 > c->frames_capacity) {
        pool_free(c->d_frames_owned, c->stream);
        c->d_frames_owned = nullptr; c->frames_capacity = 0;
        CK(pool_alloc((void**)&c->d_frames_owned, bytes64, c->stream));
        c->frames_capacity = bytes64;
    }
    if (bytes32 > c->f32_capacity) {
        pool_free(c->d_frames_f32, c->stream);
        c->d_frames_f32 = nullptr; c->f32_capacity = 0;
        CK(pool_alloc((void**)&c->d_frames_f32, bytes32, c->stream));
        c->f32_capacity = bytes32;
    }
    if (!c->copy_stream) {
        CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->order_event, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(c->order_event, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->order_event, 0));
    c->up_chunk = (long long)((32ull << 20) / (sizeof(float) * frame_elems));
    if (c->up_chunk < 1) c->up_chunk = 1;
    const size_t n_chunks = (size_t)((n + c->up_chunk - 1) / c->up_chunk);
    while (c->up_events.size() < n_chunks) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->up_events.push_back(ev);
    }
    for (size_t k = 0; k < n_chunks; ++k) {
        const long long f0 = (long long)k * c->up_chunk;
        const long long nf = (n - f0 < c->up_chunk) ? (n - f0) : c->up_chunk;
        const size_t o = (size_t)f0 * frame_elems, cnt = (size_t)nf * frame_elems;
        CK(cudaMemcpyAsync(c->d_frames_f32 + o, host + o, sizeof(float) * cnt, cudaMemcpyHostToDevice, c->copy_stream));
        size_t blocks = (cnt + 255) / 256;
        if (blocks > (size_t)c->n_sms * 8) blocks = (size_t)c->n_sms * 8;
        k_widen_frames<<<(unsigned)blocks, 256, 0, c->copy_stream>>>(c->d_frames_f32 + o, c->d_frames_owned + o, cnt);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->up_events[k], c->copy_stream));
    }
    c->up_waited = 0;
    c->d_frames = c->d_frames_owned; c->n_frames = n; c->frame0 = frame0;
    return SITB_OK;
}

// Make the compute stream wait for the uploaded chunks that cover frames [0, end).
static int wait_for_frames(sitb_ctx* c, long long end) {
    if (c->d_frames != c->d_frames_owned || c->up_chunk <= 0) return SITB_OK;
    const size_t n_chunks = (size_t)((c->n_frames + c->up_chunk - 1) / c->up_chunk);
    size_t need = (size_t)((end + c->up_chunk - 1) / c->up_chunk);
    if (need > n_chunks) need = n_chunks;
    if (need > c->up_waited) {
        CK(cudaStreamWaitEvent(c->stream, c->up_events[need - 1], 0));   // copies complete in order
        c->up_waited = need;
    }
    return SITB_OK;
}

extern "C" int sitb_upload_chunk_frames(sitb_ctx* c, int64_t* frames_per_chunk) {
    if (!c || !frames_per_chunk) return fail(SITB_E_INVALID, "sitb_upload_chunk_frames: null argument");
    *frames_per_chunk = (c->d_frames && c->d_frames == c->d_frames_owned) ? c->up_chunk : 0;
    return SITB_OK;
}

extern "C" int sitb_borrow_frames(sitb_ctx* c, const double* dev, int64_t n, int64_t frame0) {
    if (!c || !dev || n <= 0) return fail(SITB_E_INVALID, "sitb_borrow_frames: bad argument");
    c->d_frames = dev; c->n_frames = n; c->frame0 = frame0;
    return SITB_OK;
}

extern "C" int sitb_reset_status(sitb_ctx* c) {
    if (!c) return fail(SITB_E_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    unsigned long long init[2 + CNT_SLOTS];
    init[0] = init[1] = NO_ERROR_KEY;
    for (int i = 0; i < CNT_SLOTS; ++i) init[2 + i] = 0ull;
    CK(cudaMemcpyAsync(c->d_status, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SITB_OK;
}

extern "C" int sitb_get_status(sitb_ctx* c, sitb_status* out) {
    if (!c || !out) return fail(SITB_E_INVALID, "sitb_get_status: null argument");
    CK(cudaSetDevice(c->device));
    unsigned long long h[2 + CNT_SLOTS];
    CK(cudaMemcpyAsync(h, c->d_status, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    memset(out, 0, sizeof(*out));
    if (h[0] != NO_ERROR_KEY) {
        out->error_code = (int32_t)((h[0] >> 20) & 0xF);
        out->index = (int32_t)(h[0] & 0xFFFFF);
        out->frame = (int64_t)(h[0] >> 24);
    }
    if (h[1] != NO_ERROR_KEY) {
        out->zero_error = 1;
        out->zero_index = (int32_t)(h[1] & 0xFFFFF);
        out->zero_frame = (int64_t)(h[1] >> 24);
    }
    out->n_zero_rows = h[2 + CNT_ZERO_ROWS];
    out->n_duplicate_nearest = h[2 + CNT_DUP_NEAREST];
    out->n_list_overflow = h[2 + CNT_LIST_OVERFLOW];
    out->nnz = h[2 + CNT_NNZ];
    out->n_screen_rejects = h[2 + CNT_SCREEN_REJECT];
    out->n_full_walk_frames = h[2 + CNT_FULL_WALK_FRAMES];
    out->n_loose_grid_frames = h[2 + CNT_LOOSE_GRID_FRAMES];
    return SITB_OK;
}

static int base_params(sitb_ctx* c, int64_t begin, int64_t n, FillParams& p, const char* who) {
    if (!c) return fail(SITB_E_INVALID, "%s: null context", who);
    if (!c->d_frames) return fail(SITB_E_STATE, "%s: no frames resident (sitb_upload_frames / sitb_borrow_frames)", who);
    if (begin < 0 || n < 0 || begin + n > c->n_frames)
        return fail(SITB_E_INVALID, "%s: frame range [%lld, %lld) outside [0, %lld)", who, (long long)begin,
                    (long long)(begin + n), (long long)c->n_frames);
    if ((unsigned long long)(c->frame0 + c->n_frames) * (unsigned long long)c->M >= 0xFFFFFFFFull)
        return fail(SITB_E_LIMIT, "%s: more than 2^32 landmark vectors", who);
    {
        const int wrc = wait_for_frames(c, n > 0 ? begin + n : c->n_frames);   // n == 0: a frame-list pass
        if (wrc) return wrc;
    }
    memset(&p, 0, sizeof(p));
    p.cell = c->cell;
    p.frames = c->d_frames + (size_t)begin * c->A * 3;
    p.frame_list = nullptr;
    p.n_work = n;
    p.frame0 = c->frame0 + begin;
    p.A = c->A; p.S = c->S; p.M = c->M; p.L = c->L; p.V = c->V; p.Lpad = c->Lpad; p.NB = c->NB;
    p.m_magic = (c->M > 1 && c->M < 16384) ? (unsigned)((0x100000000ull + (unsigned long long)c->M - 1ull) / (unsigned long long)c->M) : 0u;
    p.static_idx = c->d_static_idx; p.mobile_idx = c->d_mobile_idx; p.ideal = c->d_ideal;
    p.tab.v0 = c->d_v0; p.tab.b0 = c->d_b0; p.tab.va = c->d_va; p.tab.ba = c->d_ba;
    p.tab.q64 = c->d_q64; p.tab.acoef = c->d_acoef; p.tab.nverts = c->d_nverts; p.tab.orig_of = c->d_orig_of;
    p.tab.chunk_atoms = c->d_chunk_atoms; p.tab.chunk_bound = c->d_chunk_bound;
    p.bcoef = c->bcoef; p.static_thr = c->static_thr; p.dynamic = c->dynamic; p.relaxed = c->relaxed;
    p.n_grid_levels = c->n_grid_levels; p.gx = c->gx; p.gy = c->gy; p.gz = c->gz;
    for (int l = 0; l < 2; ++l) {
        p.grid[l].ptr = c->grid[l].ptr; p.grid[l].list = c->grid[l].list;
        p.grid[l].sptr = c->grid[l].sptr; p.grid[l].slist = c->grid[l].slist;
        p.grid[l].margin_sq = c->grid[l].margin * c->grid[l].margin;
    }
    p.errkey = c->d_status; p.counters = c->d_status + 2;
    p.cid = c->d_cid; p.cw = c->d_cw; p.n_clusters = c->n_clusters;
    return SITB_OK;
}

extern "C" int sitb_fill_dense(sitb_ctx* c, int64_t begin, int64_t n, void* dev_out, int32_t is_f64) {
    FillParams p;
    int rc = base_params(c, begin, n, p, "sitb_fill_dense");
    if (rc) return rc;
    if (!dev_out) return fail(SITB_E_INVALID, "sitb_fill_dense: null output");
    CK(cudaSetDevice(c->device));
    p.dense_out = dev_out; p.dense_f64 = is_f64;
    CK(launch_fill(p, MODE_DENSE, c->n_sms, c->stream));
    return SITB_OK;
}

extern "C" int sitb_fill_dense_frames(sitb_ctx* c, const int64_t* dev_list, int64_t n, void* dev_out, int32_t is_f64) {
    FillParams p;
    int rc = base_params(c, 0, 0, p, "sitb_fill_dense_frames");
    if (rc) return rc;
    if (!dev_out || !dev_list) return fail(SITB_E_INVALID, "sitb_fill_dense_frames: null argument");
    CK(cudaSetDevice(c->device));
    p.frame_list = (const long long*)dev_list; p.n_work = n;
    p.dense_out = dev_out; p.dense_f64 = is_f64;
    p.errkey = nullptr; p.counters = nullptr;   // a re-evaluation of selected rows must not disturb the run's status
    CK(launch_fill(p, MODE_DENSE, c->n_sms, c->stream));
    return SITB_OK;
}

extern "C" int sitb_pass_stats(sitb_ctx* c, int64_t begin, int64_t n, uint64_t* dev_seen, double* dev_gram) {
    FillParams p;
    int rc = base_params(c, begin, n, p, "sitb_pass_stats");
    if (rc) return rc;
    if (!dev_seen || !dev_gram) return fail(SITB_E_INVALID, "sitb_pass_stats: null output");
    CK(cudaSetDevice(c->device));
    p.seen = (unsigned long long*)dev_seen; p.gram = dev_gram;
    CK(launch_fill(p, MODE_STATS, c->n_sms, c->stream));
    return SITB_OK;
}

extern "C" int sitb_set_centers(sitb_ctx* c, const int32_t* cid, const double* w, int32_t n_clusters) {
    if (!c || !cid || !w) return fail(SITB_E_INVALID, "sitb_set_centers: null argument");
    if (n_clusters < 0 || n_clusters > 32767) return fail(SITB_E_LIMIT, "sitb_set_centers: %d clusters (limit 32767)", n_clusters);
    for (int k = 0; k < c->L; ++k)
        if (cid[k] < -1 || cid[k] >= n_clusters) return fail(SITB_E_INVALID, "sitb_set_centers: cluster id %d of landmark %d out of range", cid[k], k);
    CK(cudaSetDevice(c->device));
    std::vector<int> cid_i((size_t)c->L);
    std::vector<double> w_i((size_t)c->L);
    for (int k = 0; k < c->L; ++k) { cid_i[c->internal_of[k]] = cid[k]; w_i[c->internal_of[k]] = w[k]; }
    CK(cudaMemcpyAsync(c->d_cid, cid_i.data(), sizeof(int) * (size_t)c->L, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_cw, w_i.data(), sizeof(double) * (size_t)c->L, cudaMemcpyHostToDevice, c->stream));
    if (!c->d_cid_orig) {
        CK(pool_alloc((void**)&c->d_cid_orig, sizeof(int) * (size_t)c->L, c->stream));
        CK(pool_alloc((void**)&c->d_cw_orig, sizeof(double) * (size_t)c->L, c->stream));
    }
    CK(cudaMemcpyAsync(c->d_cid_orig, cid, sizeof(int) * (size_t)c->L, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_cw_orig, w, sizeof(double) * (size_t)c->L, cudaMemcpyHostToDevice, c->stream));
    std::vector<float2> cwf;
    if (c->d_fast_cw) {
        cwf.assign((size_t)c->Lpad, make_float2(0.f, -1.f));
        c->cw_absmax = 0.0;
        for (int k = 0; k < c->L; ++k) {
            const int nv = c->h_nverts[k] ? c->h_nverts[k] : 1;
            cwf[k] = make_float2((float)w_i[k], -1.0f / (float)nv);
            if (cid_i[k] >= 0) c->cw_absmax = std::max(c->cw_absmax, std::fabs(w_i[k]));
        }
        CK(cudaMemcpyAsync(c->d_fast_cw, cwf.data(), sizeof(float2) * cwf.size(), cudaMemcpyHostToDevice, c->stream));
        c->fast_lists_dirty = true;
    }
    CK(cudaStreamSynchronize(c->stream));
    c->n_clusters = n_clusters;
    return SITB_OK;
}

extern "C" int sitb_set_assign_mode(sitb_ctx* c, int32_t mode) {
    if (!c) return fail(SITB_E_INVALID, "null context");
    if (mode != SITB_ASSIGN_EXACT && mode != SITB_ASSIGN_TWO_TIER) return fail(SITB_E_INVALID, "sitb_set_assign_mode: mode %d", mode);
    c->assign_mode = mode;
    return SITB_OK;
}

extern "C" int sitb_two_tier_info(sitb_ctx* c, int32_t* available, double* tau, double* kappa, uint64_t* counts, int32_t reset) {
    if (!c) return fail(SITB_E_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    if (available) *available = (c->fast_tables_ok && c->n_grid_levels == 2 && c->d_fast_cbox) ? 1 : 0;
    if (tau) *tau = c->fast_tau;
    if (kappa) *kappa = c->fast_kappa;
    if (counts) {
        for (int i = 0; i < SITB_TWO_TIER_SLOTS; ++i) counts[i] = 0;
        if (c->d_two_tier) {
            unsigned long long h[1 + RECHECK_SLOTS];
            CK(cudaMemcpyAsync(h, c->d_two_tier, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            for (int i = 0; i < SITB_TWO_TIER_SLOTS && i < RECHECK_SLOTS; ++i) counts[i] = h[1 + i];
        }
    }
    if (reset && c->d_two_tier) CK(cudaMemsetAsync(c->d_two_tier, 0, sizeof(unsigned long long) * (1 + RECHECK_SLOTS), c->stream));
    return SITB_OK;
}

// First tier (FP32, sitb_fill_fast.cu) over all frames, then the exact kernel over the rows it left undecided.
// Returns 1 if the shape does not fit the first tier (the caller runs the exact kernel alone).
static int two_tier_assign(sitb_ctx* c, const FillParams& base, int64_t n, double thr, int64_t* labels, double* confs,
                           uint64_t* counts) {
    if (!c->fast_tables_ok || c->n_grid_levels != 2 || !c->d_fast_cbox || c->n_clusters <= 0 || !(thr == thr) || std::isinf(thr) ||
        !(c->cw_absmax <= 64.0) || !labels || !confs || n <= 0)
        return 1;
    const size_t rows = (size_t)n * c->M;
    if (rows > c->recheck_rows_cap) {
        pool_free(c->d_recheck, c->stream); c->d_recheck = nullptr; c->recheck_rows_cap = 0;
        CK(pool_alloc((void**)&c->d_recheck, rows, c->stream));
        c->recheck_rows_cap = rows;
    }
    if ((size_t)n > c->recheck_frames_cap) {
        pool_free(c->d_frame_flag, c->stream); pool_free(c->d_frame_list, c->stream);
        c->d_frame_flag = nullptr; c->d_frame_list = nullptr; c->recheck_frames_cap = 0;
        CK(pool_alloc((void**)&c->d_frame_flag, sizeof(int) * (size_t)n, c->stream));
        CK(pool_alloc((void**)&c->d_frame_list, sizeof(long long) * (size_t)n, c->stream));
        c->recheck_frames_cap = (size_t)n;
    }
    if (c->fast_lists_dirty) {
        const long long cells = (long long)c->gx * c->gy * c->gz;
        for (int l = 0; l < 2; ++l)
            CK(launch_sort_box_lists(c->grid[l].ptr, c->grid[l].list, c->d_cid, cells, l ? (unsigned)c->grid[0].entries : 0u,
                                     c->d_fast_cbox + (size_t)l * cells, c->d_fast_clist, c->stream));
        c->fast_lists_dirty = false;
    }
    CK(cudaMemsetAsync(c->d_recheck, 0, rows, c->stream));
    CK(cudaMemsetAsync(c->d_frame_flag, 0, sizeof(int) * (size_t)n, c->stream));
    CK(cudaMemsetAsync(c->d_two_tier, 0, sizeof(unsigned long long), c->stream));
    FastParams f;
    memset(&f, 0, sizeof(f));
    f.ci0 = c->cell.ci[0]; f.ci1 = c->cell.ci[4]; f.ci2 = c->cell.ci[8];
    f.Lx = (float)c->cell.c[0]; f.Ly = (float)c->cell.c[4]; f.Lz = (float)c->cell.c[8];
    f.frames = base.frames; f.n_work = n;
    f.A = c->A; f.S = c->S; f.M = c->M; f.L = c->L; f.Lpad = c->Lpad; f.NB = c->NB; f.m_magic = base.m_magic;
    f.sm_magic = (unsigned)((0x100000000ull + (unsigned long long)(c->S + c->M) - 1ull) / (unsigned long long)(c->S + c->M));
    f.static_idx = c->d_static_idx; f.mobile_idx = c->d_mobile_idx; f.ideal_frac = c->d_ideal_frac;
    f.tab.va = c->d_va; f.tab.ib = c->d_fast_ib; f.tab.ac = c->d_fast_ac; f.tab.cw = c->d_fast_cw;
    f.bc = (float)(c->steepness * c->midpoint * 1.4426950408889634074);
    f.kappa = std::nextafter((float)c->fast_kappa, 0.0f);
    f.tau = std::nextafter((float)c->fast_tau, 1.0f);
    f.thr = (float)thr;
    f.dyn_dc = (float)(8.0 * c->fast_dc);
    f.dynamic = c->dynamic;
    const double shrink = 0.999;          // float rounding of the screen distances: stay inside the margins
    for (int l = 0; l < 2; ++l) f.margin_sq[l] = (float)(c->grid[l].margin * c->grid[l].margin * shrink);
    f.cells = c->gx * c->gy * c->gz;
    f.cbox = c->d_fast_cbox; f.clist = c->d_fast_clist; f.sbox = c->d_fast_sbox; f.slist = c->d_fast_slist;
    f.static_lim_sq = (float)(std::min(c->grid[1].margin * c->grid[1].margin, c->static_thr * c->static_thr) * shrink);
    f.gx = c->gx; f.gy = c->gy; f.gz = c->gz;
    f.gxf = (float)c->gx; f.gyf = (float)c->gy; f.gzf = (float)c->gz;
    f.labels = (long long*)labels; f.confs = confs; f.counts = (unsigned long long*)counts; f.n_clusters = c->n_clusters;
    f.recheck = c->d_recheck; f.frame_flag = c->d_frame_flag; f.frame_list = c->d_frame_list;
    f.n_list = c->d_two_tier; f.counters = c->d_two_tier + 1;
    cudaError_t e = launch_assign_fast(f, c->n_sms, c->stream);
    if (e == cudaErrorInvalidConfiguration) { cudaGetLastError(); return 1; }
    if (e != cudaSuccess) return fail(SITB_E_CUDA, "first tier of the assign pass: %s", cudaGetErrorString(e));
    // second tier: the exact kernel over the flagged frames, restricted to the flagged rows
    FillParams p = base;
    p.assign_thr = thr;
    p.labels = (long long*)labels; p.confs = confs; p.counts = (unsigned long long*)counts;
    p.frame_list = c->d_frame_list; p.n_work_dev = c->d_two_tier; p.row_filter = c->d_recheck; p.rows_by_frame = 1;
    p.counters = nullptr;
    CK(launch_fill(p, MODE_ASSIGN, c->n_sms, c->stream));
    return SITB_OK;
}

extern "C" int sitb_pass_assign(sitb_ctx* c, int64_t begin, int64_t n, double thr, int64_t* labels, double* confs,
                                uint64_t* counts, uint64_t* best, double* rep, double* rep_w, uint64_t* site_best) {
    FillParams p;
    int rc = base_params(c, begin, n, p, "sitb_pass_assign");
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    if (c->assign_mode == SITB_ASSIGN_TWO_TIER && !best && !rep && !rep_w && !site_best) {
        rc = two_tier_assign(c, p, n, thr, labels, confs, counts);
        if (rc <= 0) return rc;            // done (or failed); 1: shape not covered, fall through to the exact kernel
    }
    p.assign_thr = thr;
    p.labels = (long long*)labels; p.confs = confs; p.counts = (unsigned long long*)counts;
    p.best = (unsigned long long*)best; p.rep = rep; p.rep_w = rep_w; p.site_best = (unsigned long long*)site_best;
    CK(launch_fill(p, MODE_ASSIGN, c->n_sms, c->stream));
    return SITB_OK;
}

extern "C" int sitb_fill_landmark_vectors_host(sitb_ctx* c, const double* host_frames, int64_t n_frames,
                                               double* host_lv, sitb_status* status) {
    if (!c || !host_frames || !host_lv || n_frames <= 0)
        return fail(SITB_E_INVALID, "sitb_fill_landmark_vectors_host: bad argument");
    CK(cudaSetDevice(c->device));
    int rc = sitb_reset_status(c);
    if (rc) return rc;
    // stream the trajectory through the device in chunks: bounded device footprint for any n_frames
    const size_t row_bytes = sizeof(double) * (size_t)c->M * c->L;           // per frame, output
    const size_t in_bytes = sizeof(double) * (size_t)c->A * 3;               // per frame, input
    long long chunk = (long long)((512ull << 20) / (row_bytes + in_bytes));
    if (chunk < 1) chunk = 1;
    if (chunk > n_frames) chunk = n_frames;
    double* d_in = nullptr;
    double* d_out = nullptr;
    CK(pool_alloc((void**)&d_in, in_bytes * (size_t)chunk, c->stream));
    cudaError_t e = pool_alloc((void**)&d_out, row_bytes * (size_t)chunk, c->stream);
    if (e != cudaSuccess) { pool_free(d_in, c->stream); return fail(SITB_E_CUDA, "device allocation: %s", cudaGetErrorString(e)); }
    const double* saved_frames = c->d_frames; const long long saved_n = c->n_frames, saved_f0 = c->frame0;
    rc = SITB_OK;
    for (long long f0 = 0; f0 < n_frames && rc == SITB_OK; f0 += chunk) {
        const long long n = (n_frames - f0 < chunk) ? (n_frames - f0) : chunk;
        e = cudaMemcpyAsync(d_in, host_frames + (size_t)f0 * c->A * 3, in_bytes * (size_t)n, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { rc = fail(SITB_E_CUDA, "H2D: %s", cudaGetErrorString(e)); break; }
        c->d_frames = d_in; c->n_frames = n; c->frame0 = f0;
        rc = sitb_fill_dense(c, 0, n, d_out, 1);
        if (rc) break;
        e = cudaMemcpyAsync(host_lv + (size_t)f0 * c->M * c->L, d_out, row_bytes * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { rc = fail(SITB_E_CUDA, "D2H: %s", cudaGetErrorString(e)); break; }
    }
    c->d_frames = saved_frames; c->n_frames = saved_n; c->frame0 = saved_f0;
    pool_free(d_in, c->stream); pool_free(d_out, c->stream);
    if (rc) return rc;
    if (status) return sitb_get_status(c, status);
    return SITB_OK;
}

// ---- cached sparse landmark vectors ---------------------------------------------------------------
namespace sitb {
cudaError_t launch_assign_sparse(const unsigned long long* row_ptr, const uint16_t* pk, const double* pv,
                                 long long n_rows, long long row0, int L, const int* cid, const double* cw,
                                 int n_clusters, double thr, long long* labels, double* confs,
                                 unsigned long long* counts, unsigned long long* best, double* rep, double* rep_w,
                                 unsigned long long* site_best, int n_sms, cudaStream_t st, const long long* row_list,
                                 const unsigned long long* n_list, int slot);
cudaError_t launch_relabel_select(long long* labels, long long n_rows, const int* remap, long long* row_list,
                                  unsigned long long* n_list, int n_sms, cudaStream_t st);
cudaError_t launch_sparse_row_norm2(const unsigned long long* row_ptr, const double* pv, long long n_rows,
                                    const long long* rows, int n, double* out, cudaStream_t st);
}

static int pass_stats_cached(sitb_ctx* c, int64_t begin, int64_t n, uint64_t* dev_seen, double* dev_gram,
                             uint64_t* dev_row_ptr, uint16_t* dev_pool_k, double* dev_pool_v,
                             uint64_t* dev_cursor, uint64_t capacity, int32_t slot, const char* who) {
    FillParams p;
    int rc = base_params(c, begin, n, p, who);
    if (rc) return rc;
    if (!dev_seen || !dev_row_ptr || !dev_pool_k || !dev_pool_v || !dev_cursor)      // dev_gram may be null
        return fail(SITB_E_INVALID, "%s: null output", who);
    if (slot < 0 || slot > ENTRY_CAP) return fail(SITB_E_INVALID, "%s: slot_entries %d outside [0, %d]", who, slot, ENTRY_CAP);
    CK(cudaSetDevice(c->device));
    p.seen = (unsigned long long*)dev_seen; p.gram = dev_gram;
    p.sparse_ptr = (unsigned long long*)dev_row_ptr; p.sparse_k = dev_pool_k; p.sparse_v = dev_pool_v;
    p.sparse_cursor = (unsigned long long*)dev_cursor; p.sparse_capacity = capacity;
    p.sparse_slot = (unsigned)slot; p.sparse_row_base = (long long)begin * c->M;
    c->row_slot = slot;
    c->slot_pool_k = slot ? dev_pool_k : nullptr;
    c->slot_row_ptr = slot ? dev_row_ptr - (size_t)begin * c->M : nullptr;
    CK(launch_fill(p, MODE_STATS, c->n_sms, c->stream));
    return SITB_OK;
}

extern "C" int sitb_pass_stats_cached(sitb_ctx* c, int64_t begin, int64_t n, uint64_t* dev_seen, double* dev_gram,
                                      uint64_t* dev_row_ptr, uint16_t* dev_pool_k, double* dev_pool_v,
                                      uint64_t* dev_cursor, uint64_t capacity) {
    return pass_stats_cached(c, begin, n, dev_seen, dev_gram, dev_row_ptr, dev_pool_k, dev_pool_v, dev_cursor, capacity, 0,
                             "sitb_pass_stats_cached");
}

extern "C" int sitb_pass_stats_slotted(sitb_ctx* c, int64_t begin, int64_t n, uint64_t* dev_seen, double* dev_gram,
                                       uint64_t* dev_row_ptr, uint16_t* dev_pool_k, double* dev_pool_v,
                                       uint64_t* dev_cursor, uint64_t capacity, int32_t slot_entries) {
    return pass_stats_cached(c, begin, n, dev_seen, dev_gram, dev_row_ptr, dev_pool_k, dev_pool_v, dev_cursor, capacity,
                             slot_entries, "sitb_pass_stats_slotted");
}

extern "C" int sitb_assign_sparse(sitb_ctx* c, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k,
                                  const double* dev_pool_v, int64_t n_rows, int64_t row0, double thr,
                                  int64_t* labels, double* confs, uint64_t* counts, uint64_t* best, double* rep,
                                  double* rep_w, uint64_t* site_best) {
    if (!c || !dev_row_ptr || !dev_pool_k || !dev_pool_v || n_rows < 0)
        return fail(SITB_E_INVALID, "sitb_assign_sparse: bad argument");
    if (!c->d_cid_orig) return fail(SITB_E_STATE, "sitb_assign_sparse: no centres set (sitb_set_centers)");
    CK(cudaSetDevice(c->device));
    CK(launch_assign_sparse((const unsigned long long*)dev_row_ptr, dev_pool_k, dev_pool_v, n_rows, row0, c->L,
                            c->d_cid_orig, c->d_cw_orig, c->n_clusters, thr, (long long*)labels, confs,
                            (unsigned long long*)counts, (unsigned long long*)best, rep, rep_w,
                            (unsigned long long*)site_best, c->n_sms, c->stream, nullptr, nullptr,
                            (dev_pool_k == c->slot_pool_k && dev_row_ptr == c->slot_row_ptr) ? c->row_slot : 0));
    return SITB_OK;
}

extern "C" int sitb_sparse_row_norm2(sitb_ctx* c, const uint64_t* dev_row_ptr, const double* dev_pool_v, int64_t n_rows,
                                     const int64_t* dev_rows, int32_t n, double* dev_out) {
    if (!c || !dev_row_ptr || !dev_pool_v || !dev_rows || !dev_out || n < 0 || n_rows < 0)
        return fail(SITB_E_INVALID, "sitb_sparse_row_norm2: bad argument");
    CK(cudaSetDevice(c->device));
    CK(launch_sparse_row_norm2((const unsigned long long*)dev_row_ptr, dev_pool_v, n_rows, (const long long*)dev_rows, n, dev_out,
                               c->stream));
    return SITB_OK;
}

extern "C" int sitb_relabel_select(sitb_ctx* c, int64_t* dev_labels, int64_t n_rows, const int32_t* dev_remap,
                                   int64_t* dev_row_list, uint64_t* dev_n_list) {
    if (!c || !dev_labels || !dev_remap || !dev_row_list || !dev_n_list || n_rows < 0)
        return fail(SITB_E_INVALID, "sitb_relabel_select: bad argument");
    CK(cudaSetDevice(c->device));
    CK(launch_relabel_select((long long*)dev_labels, n_rows, dev_remap, (long long*)dev_row_list,
                             (unsigned long long*)dev_n_list, c->n_sms, c->stream));
    return SITB_OK;
}

extern "C" int sitb_assign_sparse_rows(sitb_ctx* c, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k,
                                       const double* dev_pool_v, const int64_t* dev_row_list, const uint64_t* dev_n_list,
                                       int64_t max_rows, int64_t row0, double thr, int64_t* labels, double* confs,
                                       uint64_t* counts, uint64_t* best, double* rep, double* rep_w, uint64_t* site_best) {
    if (!c || !dev_row_ptr || !dev_pool_k || !dev_pool_v || !dev_row_list || !dev_n_list || max_rows < 0)
        return fail(SITB_E_INVALID, "sitb_assign_sparse_rows: bad argument");
    if (!c->d_cid_orig) return fail(SITB_E_STATE, "sitb_assign_sparse_rows: no centres set (sitb_set_centers)");
    CK(cudaSetDevice(c->device));
    // the list is short (rows of the clusters the min_samples filter removed): a small grid is enough
    const int64_t bound = max_rows < 262144 ? max_rows : 262144;
    CK(launch_assign_sparse((const unsigned long long*)dev_row_ptr, dev_pool_k, dev_pool_v, bound > 0 ? bound : 1, row0, c->L,
                            c->d_cid_orig, c->d_cw_orig, c->n_clusters, thr, (long long*)labels, confs,
                            (unsigned long long*)counts, (unsigned long long*)best, rep, rep_w,
                            (unsigned long long*)site_best, c->n_sms, c->stream, (const long long*)dev_row_list,
                            (const unsigned long long*)dev_n_list,
                            (dev_pool_k == c->slot_pool_k && dev_row_ptr == c->slot_row_ptr) ? c->row_slot : 0));
    return SITB_OK;
}

// ---- Gram from the cached rows (sitb_gram_sparse.cu) -------------------------------------------------
namespace sitb {
cudaError_t launch_gram_sparse(const unsigned long long* row_ptr, const uint16_t* pk, const double* pv, long long n_frames,
                               int M, int L, double* gram, int n_sms, cudaStream_t st, int exact, long long frame0);
cudaError_t launch_gram_words_finish(const long long* words, int L, double* out, cudaStream_t st);
}
extern "C" int sitb_gram_from_cached(sitb_ctx* c, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k,
                                     const double* dev_pool_v, int64_t n_frames, double* dev_gram) {
    if (!c || !dev_row_ptr || !dev_pool_k || !dev_pool_v || !dev_gram || n_frames < 0)
        return fail(SITB_E_INVALID, "sitb_gram_from_cached: bad argument");
    CK(cudaSetDevice(c->device));
    CK(launch_gram_sparse((const unsigned long long*)dev_row_ptr, dev_pool_k, dev_pool_v, n_frames, c->M, c->L, dev_gram,
                          c->n_sms, c->stream, 0, 0));
    return SITB_OK;
}

extern "C" int sitb_gram_words_from_cached(sitb_ctx* c, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k,
                                           const double* dev_pool_v, int64_t n_frames, int64_t* dev_gram_words) {
    if (!c || !dev_row_ptr || !dev_pool_k || !dev_pool_v || !dev_gram_words || n_frames < 0)
        return fail(SITB_E_INVALID, "sitb_gram_words_from_cached: bad argument");
    CK(cudaSetDevice(c->device));
    CK(launch_gram_sparse((const unsigned long long*)dev_row_ptr, dev_pool_k, dev_pool_v, n_frames, c->M, c->L,
                          (double*)dev_gram_words, c->n_sms, c->stream, 1, c->frame0));
    return SITB_OK;
}

extern "C" int sitb_gram_words_finish(int device, const int64_t* dev_gram_words, int32_t n_landmarks, double* dev_gram_upper,
                                      void* cuda_stream) {
    if (!dev_gram_words || !dev_gram_upper || n_landmarks <= 0) return fail(SITB_E_INVALID, "sitb_gram_words_finish: bad argument");
    CK(cudaSetDevice(device));
    CK(launch_gram_words_finish((const long long*)dev_gram_words, n_landmarks, dev_gram_upper, (cudaStream_t)cuda_stream));
    return SITB_OK;
}

// ---- staging for the tensor-core Gram (sitb_gram_tc.cu) ---------------------------------------------
extern "C" int sitb_pass_stage(sitb_ctx* c, int64_t begin, int64_t n, uint64_t* dev_seen, void* dev_stage_hi,
                               void* dev_stage_lo, int64_t ld) {
    FillParams p;
    int rc = base_params(c, begin, n, p, "sitb_pass_stage");
    if (rc) return rc;
    if (!dev_seen || !dev_stage_hi || !dev_stage_lo || ld < n * c->M || ld % 64 != 0)
        return fail(SITB_E_INVALID, "sitb_pass_stage: bad argument (ld must hold n * n_mobile rows and be a multiple of 64)");
    CK(cudaSetDevice(c->device));
    p.seen = (unsigned long long*)dev_seen;
    p.stage_hi = (__half*)dev_stage_hi; p.stage_lo = (__half*)dev_stage_lo; p.stage_ld = ld;
    CK(launch_fill(p, MODE_STAGE, c->n_sms, c->stream));
    return SITB_OK;
}

// ---- site centres (LandmarkAnalysis.py:276-287, PBCCalculator.pyx:106-139) -----------------------
namespace sitb {
cudaError_t launch_wrapped_rows(const Cell& cell, const double* frames, int A, int M, const int* mobile_idx,
                                long long frame0, long long n_frames, const long long* rows, int n, double* out,
                                cudaStream_t st);
cudaError_t launch_site_accumulate(const Cell& cell, const double* frames, int A, int M, const int* mobile_idx,
                                   long long n_frames, const long long* labels, const double* confs,
                                   const double* offset, int C, int weighted, double* sums, int n_sms, cudaStream_t st);
cudaError_t launch_site_finish(const Cell& cell, const double* sums, const double* offset, int C, double* centers,
                               cudaStream_t st);
cudaError_t launch_first_row(const long long* labels, long long n, long long row0, int C, unsigned long long* first,
                             int n_sms, cudaStream_t st);
}

extern "C" int sitb_wrapped_mobile_rows(sitb_ctx* c, const int64_t* dev_rows, int32_t n, double* dev_out) {
    if (!c || !dev_rows || !dev_out || n < 0) return fail(SITB_E_INVALID, "sitb_wrapped_mobile_rows: bad argument");
    if (!c->d_frames) return fail(SITB_E_STATE, "sitb_wrapped_mobile_rows: no frames resident");
    { const int wrc = wait_for_frames(c, c->n_frames); if (wrc) return wrc; }
    CK(cudaSetDevice(c->device));
    CK(launch_wrapped_rows(c->cell, c->d_frames, c->A, c->M, c->d_mobile_idx, c->frame0, c->n_frames,
                           (const long long*)dev_rows, n, dev_out, c->stream));
    return SITB_OK;
}

extern "C" int sitb_site_first_rows(sitb_ctx* c, const int64_t* dev_labels, int32_t n_sites, uint64_t* dev_first) {
    if (!c || !dev_labels || !dev_first || n_sites <= 0) return fail(SITB_E_INVALID, "sitb_site_first_rows: bad argument");
    CK(cudaSetDevice(c->device));
    CK(launch_first_row((const long long*)dev_labels, c->n_frames * c->M, c->frame0 * c->M, n_sites,
                        (unsigned long long*)dev_first, c->n_sms, c->stream));
    return SITB_OK;
}

extern "C" int sitb_site_accumulate(sitb_ctx* c, const int64_t* dev_labels, const double* dev_confs,
                                    const double* dev_offsets, int32_t n_sites, int32_t weighted, double* dev_sums) {
    if (!c || !dev_labels || !dev_offsets || !dev_sums || n_sites <= 0 || (weighted && !dev_confs))
        return fail(SITB_E_INVALID, "sitb_site_accumulate: bad argument");
    if (!c->d_frames) return fail(SITB_E_STATE, "sitb_site_accumulate: no frames resident");
    { const int wrc = wait_for_frames(c, c->n_frames); if (wrc) return wrc; }
    CK(cudaSetDevice(c->device));
    CK(launch_site_accumulate(c->cell, c->d_frames, c->A, c->M, c->d_mobile_idx, c->n_frames,
                              (const long long*)dev_labels, dev_confs, dev_offsets, n_sites, weighted, dev_sums,
                              c->n_sms, c->stream));
    return SITB_OK;
}

extern "C" int sitb_site_finish(sitb_ctx* c, const double* dev_sums, const double* dev_offsets, int32_t n_sites,
                                double* dev_centers) {
    if (!c || !dev_sums || !dev_offsets || !dev_centers || n_sites <= 0)
        return fail(SITB_E_INVALID, "sitb_site_finish: bad argument");
    CK(cudaSetDevice(c->device));
    CK(launch_site_finish(c->cell, dev_sums, dev_offsets, n_sites, dev_centers, c->stream));
    return SITB_OK;
}

namespace sitb {
cudaError_t launch_weighted_point_average(const Cell& cell, const double* pts, const double* w, int C, int P,
                                          double* out, cudaStream_t st);
}

extern "C" int sitb_weighted_point_average(sitb_ctx* c, const double* dev_points, const double* dev_weights,
                                           int32_t n_sites, int32_t n_points, double* dev_out) {
    if (!c || !dev_points || !dev_weights || !dev_out || n_sites <= 0 || n_points <= 0)
        return fail(SITB_E_INVALID, "sitb_weighted_point_average: bad argument");
    CK(cudaSetDevice(c->device));
    CK(launch_weighted_point_average(c->cell, dev_points, dev_weights, n_sites, n_points, dev_out, c->stream));
    return SITB_OK;
}
