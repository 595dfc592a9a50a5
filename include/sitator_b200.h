/* sitator_b200 -- C ABI of the B200-native landmark-analysis path.
 *
 * Drop-in boundary for the native part of sitator's landmark analysis.  The reference's only
 * native entry on this path is the Cython function
 *     helpers._fill_landmark_vectors(self, sn, verts_np, site_vert_dists, frames, ...)
 *         (/root/reference/sitator/landmark/helpers.pyx:12-124)
 * plus the library calls its MCL clustering plugin makes on the landmark-vector matrix
 *     (sitator/landmark/cluster/mcl.py:53-54, :81-89, :98-122; sitator/util/mcl.py:3-60;
 *      sitator/util/DotProdClassifier.pyx:68-197)
 * and the per-frame scans of SiteTrajectory / JumpAnalysis
 *     (sitator/SiteTrajectory.py:205-232, :307-373; sitator/dynamics/JumpAnalysis.py:27-135).
 * Each entry point below names the reference code it replaces.
 *
 * Conventions: plain C, no exceptions.  Every function returns 0 on success or a negative
 * SITB_E_* code; sitb_last_error() gives the message for the calling thread.  Pointers named
 * host_* are host memory, dev_* are device memory on the context's GPU (e.g. a
 * torch tensor's data_ptr()).  Work is enqueued on the context's CUDA stream
 * (sitb_set_stream); functions that return results to the host synchronise that stream.
 * There is no CPU fallback: without a CUDA device sitb_create fails with SITB_E_CUDA.
 */
#ifndef SITATOR_B200_H
#define SITATOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SITB_OK 0
#define SITB_E_INVALID (-1)   /* bad argument */
#define SITB_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define SITB_E_STATE (-3)     /* call sequence error (e.g. no frames uploaded) */
#define SITB_E_LIMIT (-4)     /* a documented size limit was exceeded */

/* error_code values of sitb_status, in the reference's raise order inside one frame */
#define SITB_ERR_NONE 0
#define SITB_ERR_STATIC_MOVED 1       /* helpers.pyx:76-80  StaticLatticeError(lattice_atoms=[index], frame) */
#define SITB_ERR_STATIC_UNASSIGNED 2  /* helpers.pyx:87-92  StaticLatticeError(lattice_atoms=unassigned, frame) */
#define SITB_ERR_ZERO_LANDMARK 3      /* helpers.pyx:116-118 ZeroLandmarkError(mobile_index=index, frame) */

typedef struct sitb_ctx sitb_ctx;

/* The landmark basis and the analysis parameters (LandmarkAnalysis.__init__, LandmarkAnalysis.py:95-130;
 * SiteNetwork inputs, LandmarkAnalysis.py:179-202). */
typedef struct sitb_network_desc {
    int32_t n_atoms;       /* atoms per frame (sn.n_total) */
    int32_t n_static;      /* sn.n_static  */
    int32_t n_mobile;      /* sn.n_mobile  */
    int32_t n_landmarks;   /* sn.n_sites = landmark dimension */
    int32_t max_verts;     /* columns of verts (<= 16) */
    const double* host_cellmat;      /* [3][3] = cell^T           (PBCCalculator.pyx:33) */
    const double* host_cellmat_inv;  /* [3][3] inverse of cellmat (PBCCalculator.pyx:34); NULL: computed by adjugate */
    const int32_t* host_static_idx;  /* [n_static] frame index of each static-lattice atom, ascending */
    const int32_t* host_mobile_idx;  /* [n_mobile] */
    const double* host_ideal_static; /* [n_static][3] sn.static_structure.positions */
    const double* host_centers;      /* [n_landmarks][3] sn.centers */
    const int32_t* host_verts;       /* [n_landmarks][max_verts] indexes into the static lattice, -1 padded (LandmarkAnalysis.py:195) */
    double cutoff_midpoint;          /* LandmarkAnalysis.py:98 */
    double cutoff_steepness;         /* LandmarkAnalysis.py:99 */
    double cutoff_round_to_zero;     /* helpers.pyx:127-131 evaluated by the caller (libm log), or <= 0 to compute here */
    double static_movement_threshold;/* LandmarkAnalysis.py:103 */
    int32_t dynamic_lattice_mapping; /* LandmarkAnalysis.py:104 */
    int32_t relaxed_lattice_checks;  /* LandmarkAnalysis.py:105 */
} sitb_network_desc;

typedef struct sitb_status {
    int32_t error_code;   /* SITB_ERR_*: the first error in the reference's iteration order */
    int32_t index;        /* lattice index (STATIC_MOVED) or mobile index (ZERO_LANDMARK) */
    int64_t frame;        /* global frame index of that error */
    int32_t zero_error;   /* 1 if a zero landmark vector occurred; zero_frame/zero_index locate the first */
    int32_t zero_index;
    int64_t zero_frame;
    uint64_t n_zero_rows;        /* self.n_all_zero_lvecs (helpers.pyx:124) */
    uint64_t n_duplicate_nearest;/* count of the warning at helpers.pyx:69-71 */
    uint64_t n_list_overflow;    /* rows with more than 128 non-zero components (must be 0) */
    uint64_t nnz;                /* non-zero landmark-vector components produced */
    uint64_t n_screen_rejects;       /* landmarks that passed the float pre-screen but failed the exact double test */
    uint64_t n_full_walk_frames;     /* frames with a static atom beyond the candidate-grid margin (all landmarks walked) */
    uint64_t n_loose_grid_frames;    /* frames with a static atom beyond half the margin (the looser candidate lists) */
} sitb_status;

const char* sitb_last_error(void);
int sitb_version(void);
/* out[0] = sizeof(sitb_network_desc), out[1] = sizeof(sitb_status): a binding compares them with its own struct
 * definitions before the first call that writes through such a pointer (sitb_get_status fills the whole struct). */
int sitb_abi_sizes(uint64_t* out2);

/* Context: device copies of the basis + precomputed tables (replaces LandmarkAnalysis.py:179, :191-202). */
int sitb_create(const sitb_network_desc* desc, int device, sitb_ctx** out);
void sitb_destroy(sitb_ctx* ctx);
int sitb_set_stream(sitb_ctx* ctx, void* cuda_stream);
/* Candidate grid of the fill kernel (orthorhombic cells; no reference counterpart -- the reference walks every
 * landmark, helpers.pyx:188).  Frames whose static atoms all lie within static_margin (Angstrom) of their ideal
 * positions test only the landmarks that can be non-zero in the grid box of the mobile atom; other frames walk
 * all landmarks.  Lists are kept for static_margin and for half of it; a frame uses the tightest that covers its
 * largest static displacement.  Results do not depend on it.  sitb_create builds it with 0.5 A (env
 * SITB_GRID_MARGIN overrides); static_margin <= 0 removes it. */
int sitb_set_candidate_grid(sitb_ctx* ctx, double static_margin);
int sitb_candidate_grid_info(sitb_ctx* ctx, int32_t* dims3, double* static_margin, uint64_t* n_entries);
int sitb_device_info(sitb_ctx* ctx, int32_t* n_sms, int32_t* cc_major, int32_t* cc_minor);

/* site_vert_dists [L][V] (NaN padded, LandmarkAnalysis.py:196-202) and the squared cut-off table. */
int sitb_get_tables(sitb_ctx* ctx, double* host_site_vert_dists, double* host_q_cutoff);

/* Frames: copied once and kept resident for all passes, or borrowed from the caller's device buffer.
 * frame0 = global index of the first frame (frame-sharded runs). */
int sitb_upload_frames(sitb_ctx* ctx, const double* host_frames, int64_t n_frames, int64_t frame0);
/* float32 trajectories (not accepted by the reference, whose Cython fill is typed double): copied at half the
 * PCIe cost and widened to float64 on the device, so results equal a run on frames.astype(float64). */
int sitb_upload_frames_f32(sitb_ctx* ctx, const float* host_frames, int64_t n_frames, int64_t frame0);
int sitb_borrow_frames(sitb_ctx* ctx, const double* dev_frames, int64_t n_frames, int64_t frame0);
/* sitb_upload_frames copies in chunks on its own stream and returns at once when host_frames is page-locked (the
 * caller keeps it alive and unchanged until the passes reading it have run); a pass waits only for the chunks
 * it reads, so passes launched chunk by chunk overlap the copy.  *frames_per_chunk = 0 for borrowed frames. */
int sitb_upload_chunk_frames(sitb_ctx* ctx, int64_t* frames_per_chunk);

int sitb_reset_status(sitb_ctx* ctx);
int sitb_get_status(sitb_ctx* ctx, sitb_status* out);

/* helpers._fill_landmark_vectors (helpers.pyx:12-124) into a device matrix [(n*M)][L], float32 or float64. */
int sitb_fill_dense(sitb_ctx* ctx, int64_t frame_begin, int64_t n, void* dev_out, int32_t out_is_f64);
/* same for a list of (local) frame indices: rows i*M..i*M+M-1 of dev_out hold frame dev_frame_list[i] */
int sitb_fill_dense_frames(sitb_ctx* ctx, const int64_t* dev_frame_list, int64_t n, void* dev_out, int32_t out_is_f64);

/* cluster/mcl.py:53-54: seen_ntimes [L] (+=) and the un-normalised Gram sum_rows lv^T lv [L][L] (+=, upper
 * triangle) by sparse outer products in FP64 -- the exact cross-check of the tensor-core SYRK. */
int sitb_pass_stats(sitb_ctx* ctx, int64_t frame_begin, int64_t n, uint64_t* dev_seen, double* dev_gram_upper);

/* sitb_pass_stats that also keeps every landmark vector in compressed form (they are ~1.5 % dense):
 * dev_gram_upper may be NULL (then build the Gram with sitb_gram_from_cached);
 * dev_row_ptr[n*M] = offset << 8 | count (all ones: pool exhausted, grow and rerun), entries
 * (landmark index uint16, value float64) at dev_pool_k/v[offset ...], *dev_cursor = entries used. */
int sitb_pass_stats_cached(sitb_ctx* ctx, int64_t frame_begin, int64_t n, uint64_t* dev_seen, double* dev_gram_upper,
                           uint64_t* dev_row_ptr, uint16_t* dev_pool_k, double* dev_pool_v, uint64_t* dev_cursor,
                           uint64_t capacity);
/* The same with the rows in row order: a row of at most slot_entries entries is stored at offset
 * (frame_begin * n_mobile + row) * slot_entries -- fixed slots, so the passes over the cached rows stream them
 * sequentially (rows scattered over the pool cost a DRAM page each); longer rows take their space from *dev_cursor,
 * which the caller initialises to (resident frames * n_mobile) * slot_entries.  dev_row_ptr as above. */
int sitb_pass_stats_slotted(sitb_ctx* ctx, int64_t frame_begin, int64_t n, uint64_t* dev_seen, double* dev_gram_upper,
                            uint64_t* dev_row_ptr, uint16_t* dev_pool_k, double* dev_pool_v, uint64_t* dev_cursor,
                            uint64_t capacity, int32_t slot_entries);
/* The Gram of rows cached by sitb_pass_stats_cached (pass dev_gram_upper = NULL there): dev_gram_upper [L][L] += the
 * upper triangle of sum_rows lv^T lv over the n_frames * n_mobile cached rows, accumulated per (mobile atom, window
 * of consecutive frames) in shared memory before it touches the matrix.  Fails with SITB_E_CUDA
 * (invalid configuration) when n_landmarks is too large for the shared-memory tables (> ~9000). */
int sitb_gram_from_cached(sitb_ctx* ctx, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k,
                          const double* dev_pool_v, int64_t n_frames, double* dev_gram_upper);
/* The same Gram, accumulated deterministically: every addend is split into three 64-bit integers (units of 2^-30,
 * 2^-62, 2^-94; exact for addends >= 2^-42) and added with integer atomics, so the result is the exact sum rounded
 * once -- independent of the order of addition, bit-identical from run to run, and for every frame sharding whose
 * boundaries are multiples of the window length (16 frames; windows are aligned to global frame numbers, and the
 * integer words of the shards are summed before sitb_gram_words_finish converts them).
 * dev_gram_words: int64 [2 (n_landmarks + 1)][n_landmarks], zeroed by the caller (+=).  Plane 0: upper triangle =
 * units of 2^-30, lower triangle = units of 2^-62 of the mirrored entry, row n_landmarks = units of 2^-62 of the
 * diagonal; plane 1: units of 2^-94 at the positions of the 2^-62 words.
 * sitb_gram_words_finish writes the float64 upper triangle (lower triangle zero) that sitb_landmark_graph reads. */
int sitb_gram_words_from_cached(sitb_ctx* ctx, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k,
                                const double* dev_pool_v, int64_t n_frames, int64_t* dev_gram_words);
int sitb_gram_words_finish(int device, const int64_t* dev_gram_words, int32_t n_landmarks, double* dev_gram_upper,
                           void* cuda_stream);
/* sitb_pass_assign over rows cached by sitb_pass_stats_cached / sitb_pass_stats_slotted (same outputs, same
 * semantics; given the buffers of the context's last sitb_pass_stats_slotted call with dev_row_ptr = the shard's row 0,
 * the entry loads use the slot layout and do not wait for the row pointers); row0 = global
 * index of the first row.  The later passes of the clustering plugin (cluster/mcl.py:81-83, :98-122) stream the
 * compressed rows instead of recomputing them. */
int sitb_assign_sparse(sitb_ctx* ctx, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k, const double* dev_pool_v,
                       int64_t n_rows, int64_t row0, double threshold, int64_t* dev_labels, double* dev_confs,
                       uint64_t* dev_counts, uint64_t* dev_best, double* dev_rep, double* dev_rep_w,
                       uint64_t* dev_site_best);

/* |lvec|^2 of selected cached rows (dev_rows[n]: indices local to the cache; out of range -> 0, so shards can be summed):
 * the norm of each cluster's best-matching landmark vector (cluster/mcl.py:84-88) without materialising the rows. */
int sitb_sparse_row_norm2(sitb_ctx* ctx, const uint64_t* dev_row_ptr, const double* dev_pool_v, int64_t n_rows,
                          const int64_t* dev_rows, int32_t n, double* dev_out);

/* The min_samples filter of DotProdClassifier.fit_predict (util/DotProdClassifier.pyx:105-118) without a second
 * full predict.  dev_remap[C_old]: new id of each first-predict cluster, -1 = removed.  sitb_relabel_select renumbers
 * dev_labels in place and lists the rows of removed clusters (dev_row_list[<= n_rows], *dev_n_list += their number);
 * sitb_assign_sparse_rows then predicts exactly those rows again with the current centres (sitb_set_centers with the
 * surviving ones) and adds them to the reductions.  Rows of surviving clusters keep arg-max and confidence: the
 * surviving centres are unchanged, and removing centres cannot lift another one above the old maximum. */
int sitb_relabel_select(sitb_ctx* ctx, int64_t* dev_labels, int64_t n_rows, const int32_t* dev_remap,
                        int64_t* dev_row_list, uint64_t* dev_n_list);
int sitb_assign_sparse_rows(sitb_ctx* ctx, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k,
                            const double* dev_pool_v, const int64_t* dev_row_list, const uint64_t* dev_n_list,
                            int64_t max_rows, int64_t row0, double threshold, int64_t* dev_labels, double* dev_confs,
                            uint64_t* dev_counts, uint64_t* dev_best, double* dev_rep, double* dev_rep_w,
                            uint64_t* dev_site_best);

/* Cluster centres (cluster/mcl.py:70-96): centres have disjoint supports, so they are given as a
 * landmark -> cluster map (-1 none) and a landmark weight. */
int sitb_set_centers(sitb_ctx* ctx, const int32_t* host_cluster_of_landmark, const double* host_weight,
                     int32_t n_clusters);

/* DotProdClassifier.predict (DotProdClassifier.pyx:129-197, predict_normed=False) fused with the fill.
 * Any output pointer may be NULL.
 *   dev_labels[n*M] int64 (-1 unassigned), dev_confs[n*M] float64,
 *   dev_counts[C] += bincount(labels)                              (DotProdClassifier.pyx:92)
 *   dev_best[3C]  = max over rows of (|centre.x|, first row)       (cluster/mcl.py:81-83)
 *   dev_rep[C][L] += conf * lvec, dev_rep_w[C] += conf             (cluster/mcl.py:118-122)
 *   dev_site_best[3C] = max over rows of (conf, first row)         (PBCCalculator.pyx:120-122 via LandmarkAnalysis.py:285)
 * A "best" table is 3*C uint64: [0,C) the value as float64 bits, [C,2C) the global row, [2C,3C) lock words;
 * initialise it to zeros (value 0.0 at row 0 is what np.argmax gives when nothing matches). */
int sitb_pass_assign(sitb_ctx* ctx, int64_t frame_begin, int64_t n, double threshold, int64_t* dev_labels,
                     double* dev_confs, uint64_t* dev_counts, uint64_t* dev_best, double* dev_rep,
                     double* dev_rep_w, uint64_t* dev_site_best);

/* Two-tier evaluation of sitb_pass_assign (labels / confs / counts outputs only; orthorhombic cells with a candidate
 * grid; no reference counterpart -- the reference evaluates everything in float64, helpers.pyx:10).
 * SITB_ASSIGN_TWO_TIER: a first kernel evaluates every landmark-vector component in FP32 with a proven error bound
 * (tau, relative, per component) and takes a row's decisions -- which components are non-zero (helpers.pyx:199-203),
 * the arg-max cluster (DotProdClassifier.pyx:181), the threshold test (:184-186) -- only if they hold for every value
 * inside the bound; all other rows are then redone by the exact float64 kernel.  Labels are therefore identical to
 * SITB_ASSIGN_EXACT; confidences of first-tier rows differ by at most tau * sum |component * centre weight|.
 * Shapes the first tier does not cover (triclinic cells, no grid, shared memory) silently use the exact kernel. */
#define SITB_ASSIGN_EXACT 0
#define SITB_ASSIGN_TWO_TIER 1
int sitb_set_assign_mode(sitb_ctx* ctx, int32_t mode);
/* counts[SITB_TWO_TIER_SLOTS] (accumulated over passes since the last reset): rows left to the exact kernel because
 * [0] their frame has a static atom beyond the grid margin / an ambiguous lattice map, [1] a component's cut-off test
 * was inside the FP32 error band, [2] the two largest similarities were closer than the bound, [3] the largest
 * similarity was within the bound of the assignment threshold, [4] more than 64 non-zero components; [5] all of them. */
#define SITB_TWO_TIER_SLOTS 6
int sitb_two_tier_info(sitb_ctx* ctx, int32_t* available, double* tau, double* kappa, uint64_t* counts, int32_t reset);

/* ---- the "dotprod" clustering plugin (landmark/cluster/dotprod.py:11-33), over rows cached by
 * sitb_pass_stats_cached ----
 * sitb_dotprod_fit: the first (and only long) iteration of DotProdClassifier.fit_centers
 *   (util/DotProdClassifier.pyx:228-289): the rows, IN ORDER, join the centre of highest cosine similarity when it
 *   reaches threshold, else found a new centre.  State is returned in sum form: dev_sums [max_centers][L] (zeroed by
 *   the caller) = sum of the member rows, dev_counts [max_centers] (zeroed) = members (all-zero rows count for
 *   centre 0, as NumPy's arg-max of NaNs makes them), dev_norm2 [max_centers] = |sum|^2; centre = sum / count.
 *   Scratch: dev_lists [L][list_cap] uint16, dev_list_len [L] uint16 (zeroed) = the centres non-zero at a landmark.
 *   dev_out16 = {centres found, status (0 ok, 1 more than max_centers, 2 a landmark list is full: grow and rerun),
 *   rows consumed, then diagnostics: SM cycles spent finding candidates / in the dot products / committing, and
 *   the total number of candidates}.  The later iterations of fit_centers (:228 loop) act on the few hundred centres; the
 *   caller runs them.
 * sitb_dotprod_predict: DotProdClassifier.predict with predict_normed=True (:129-197) for dense, already
 *   normalised centres dev_normed_centers [n_centers][L] plus, per landmark, the centres non-zero there (CSR:
 *   dev_list_ptr [L+1], dev_list_centers ascending); dev_labels / dev_confs [n_rows], dev_counts [n_centers] +=. */
int sitb_dotprod_limits(int32_t* max_centers, int32_t* max_row_entries);
int sitb_dotprod_fit(int device, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k, const double* dev_pool_v,
                     int64_t n_rows, int32_t n_landmarks, double threshold, int32_t max_centers, int32_t list_cap,
                     double* dev_sums, int64_t* dev_counts, double* dev_norm2, uint16_t* dev_lists,
                     uint16_t* dev_list_len, int64_t* dev_out16, void* cuda_stream);
int sitb_dotprod_predict(int device, const uint64_t* dev_row_ptr, const uint16_t* dev_pool_k, const double* dev_pool_v,
                         int64_t n_rows, int32_t n_landmarks, int32_t n_centers, const double* dev_normed_centers,
                         const uint32_t* dev_list_ptr, const uint16_t* dev_list_centers, double threshold,
                         int64_t* dev_labels, double* dev_confs, uint64_t* dev_counts, void* cuda_stream);

/* cluster/mcl.py:54-59: cov = gram/n_rows, correlation graph clipped at 0 with unit self loops for
 * never-seen landmarks; dev_gram_upper is the [n][n] upper triangle written by sitb_pass_stats
 * (or the SYRK), dev_cov and dev_graph are full [n][n] float64 outputs. */
int sitb_landmark_graph(int device, const double* dev_gram_upper, int32_t n, double n_rows, double* dev_cov,
                        double* dev_graph, void* cuda_stream);

/* util/mcl.py:3-60 markov_clustering on the device in float64: dev_graph [n][n] (non-zero diagonal) ->
 * dev_result [n][n] = the converged matrix m2 (attractor rows are read off it by the caller,
 * util/mcl.py:52-60).  *converged = 0 when iterlimit was reached (the reference raises ValueError). */
int sitb_markov_clustering(int device, const double* dev_graph, int32_t n, int32_t expansion, double inflation,
                           double pruning_threshold, int32_t iterlimit, double* dev_result,
                           int32_t* n_iterations, int32_t* converged, void* cuda_stream);

/* cluster/mcl.py:73-80: the principal eigenvector of cov[cluster][:, cluster] for every cluster (what
 * scipy.sparse.linalg.eigsh(block, k=1) returns there; [1] for a singleton), float64 cyclic Jacobi, one CTA per cluster.
 * dev_members[offsets[c] .. offsets[c+1]) = the landmarks of cluster c; the unit eigenvector is scattered into
 * dev_weights[n_landmarks] (zero it first).  dev_sweeps[c] = Jacobi sweeps used, or -1 for a block larger than 64
 * (left to the caller).  The sign of an eigenvector is arbitrary here as in the reference. */
int sitb_principal_vectors(int device, const double* dev_cov, int32_t n_landmarks, const int32_t* dev_members,
                           const int32_t* dev_offsets, int32_t n_clusters, double* dev_weights, int32_t* dev_sweeps,
                           void* cuda_stream);

/* ---- site centres: LandmarkAnalysis.py:276-287 via PBCCalculator.average (PBCCalculator.pyx:106-139) ----
 * The average is centred on one point per site (the max-confidence row, or the first row when
 * unweighted); the caller turns those points into offsets = cell centroid - point.
 *   sitb_wrapped_mobile_rows: wrapped (LandmarkAnalysis.py:182-189) position of global rows -> [n][3]
 *       (rows outside the resident shard give 0 0 0, so shards can be summed)
 *   sitb_site_first_rows:     dev_first[s] = min(dev_first[s], first global row with label s)
 *   sitb_site_accumulate:     dev_sums[s] += (w*x, w*y, w*z, w) of wrap(position + offset[s])
 *   sitb_site_finish:         centre = wrap(sum/w - offset) */
int sitb_wrapped_mobile_rows(sitb_ctx* ctx, const int64_t* dev_rows, int32_t n, double* dev_out);
int sitb_site_first_rows(sitb_ctx* ctx, const int64_t* dev_labels, int32_t n_sites, uint64_t* dev_first);
int sitb_site_accumulate(sitb_ctx* ctx, const int64_t* dev_labels, const double* dev_confs,
                         const double* dev_offsets, int32_t n_sites, int32_t weighted, double* dev_sums);
int sitb_site_finish(sitb_ctx* ctx, const double* dev_sums, const double* dev_offsets, int32_t n_sites,
                     double* dev_centers);

/* LandmarkAnalysis.py:288-296 (representative-landmark site centres): per site s the periodic average of the
 * points with dev_weights[s][p] > 0, weighted by them, centred on the max-weight point -> dev_out [n_sites][3] */
int sitb_weighted_point_average(sitb_ctx* ctx, const double* dev_points, const double* dev_weights,
                                int32_t n_sites, int32_t n_points, double* dev_out);

/* ---- integer passes over the assignment stream dev_traj [n_frames][n_mobile] int64 (-1 unknown) ----
 * SiteTrajectory.check_multiple_occupancy (SiteTrajectory.py:205-232):
 *   dev_out3 += {#(frame,site) holding >1 atom, #assigned atoms, #(frame,site) occupied};
 *   dev_first_bad = min(frame << 32 | lowest site above max_mobile_per_site) (init all ones). */
int sitb_check_multiple_occupancy(int device, const int64_t* dev_traj, int64_t n_frames, int32_t n_mobile,
                                  int64_t frame0, int32_t max_mobile_per_site, uint64_t* dev_out3,
                                  uint64_t* dev_first_bad, void* cuda_stream);
/* SiteTrajectory._jumped_generator (SiteTrajectory.py:353-373): dev_from[f][a] = site left at frame f,
 * -2 where the atom did not jump; *dev_total += number of jumps.  dev_carry_in [n_mobile] (may be NULL)
 * is the last known site of each atom before this shard (unknown_as_jump: the previous frame's row). */
int sitb_jump_scan(int device, const int64_t* dev_traj, int64_t n_frames, int32_t n_mobile,
                   int32_t unknown_as_jump, int32_t first_frame_is_start, const int64_t* dev_carry_in,
                   int32_t* dev_from, uint64_t* dev_total, void* cuda_stream);
/* SiteTrajectory.jumps (SiteTrajectory.py:307-329): ordered list of (frame, atom, from, to) int64 rows */
int sitb_jump_compact(int device, const int64_t* dev_traj, const int32_t* dev_from, int64_t n_frames,
                      int32_t n_mobile, int64_t frame0, int64_t* dev_out, uint64_t capacity, void* cuda_stream);
/* JumpAnalysis.run (dynamics/JumpAnalysis.py:27-135) accumulators, with NumPy's duplicate-index
 * semantics: dev_n_ij [C][C] f64, dev_total_time [C], dev_lag_sum [C][C] f64, dev_lag_n [C][C], all +=.
 * Carries (may be NULL): last known site and local frame index of the last jump before this shard. */
int sitb_jump_analysis(int device, const int64_t* dev_traj, int64_t n_frames, int32_t n_mobile, int32_t n_sites,
                       int32_t first_frame_is_start, const int64_t* dev_carry_label, const int64_t* dev_carry_jump,
                       double* dev_n_ij, uint64_t* dev_total_time, double* dev_lag_sum, uint64_t* dev_lag_n,
                       uint64_t* dev_n_problems, void* cuda_stream);

/* Frame-sharded JumpAnalysis: one shard's per-atom summary dev_summary[n_mobile][4] = (first frame with a known site,
 * that site, last known site, frame of the last jump decided inside the shard); local frame indices, -1 = none.
 * The caller chains the shards' summaries in frame order into dev_carry_label / dev_carry_jump of sitb_jump_analysis
 * (the sequential state of JumpAnalysis.py:46-49,91-96: last_known and time_at_current). */
int sitb_jump_analysis_summary(int device, const int64_t* dev_traj, int64_t n_frames, int32_t n_mobile,
                               int64_t* dev_summary, void* cuda_stream);

/* ---- post-processing of the assignment stream ----
 * sitb_assign_last_known: SiteTrajectory.assign_to_last_known_site (SiteTrajectory.py:235-304), in place on dev_traj:
 *   an unknown entry takes the atom's last known site while fewer than frame_threshold frames have passed since
 *   it was last known.  dev_carry_label / dev_carry_time [n_mobile] (may be NULL: -1 / 0) = state before this
 *   shard; dev_end_label / dev_end_time [n_mobile] (may be NULL) = state after it.  apply = 0 only chains the
 *   state (dev_traj untouched).  dev_stats4 += {entries reassigned, sum and number of the unknown stretches that
 *   ended}; [3] = max over ended stretches longer than the threshold of (global frame << 24 | length): the
 *   reference reports the length seen at the LAST such frame (:271-273).
 * sitb_windowed_mode: running_windowed_mode (dynamics/SmoothSiteTrajectory.pyx:79-111): dev_out[f][m] = the most
 *   frequent label of frames [f - wleft, f + wright) if it occurs >= threshold times, else unknown
 *   (replace_no_winner_unknown) or the input label.  dev_before / dev_after (may be NULL) = the last halo_before /
 *   first halo_after frames of the neighbouring shards.
 * sitb_seen_sites / sitb_relabel_sites: RemoveUnoccupiedSites (dynamics/RemoveUnoccupiedSites.py:30-57): dev_seen
 *   [n_sites] uint32 (zeroed) = 1 where a site occurs; labels mapped through dev_translation [n_sites]. */
/* RecenterTrajectory (util/RecenterTrajectory.pyx:14-100), in place on dev_array [n_frames][n_atoms][3]: every frame
 * minus sum_j dev_weights[j] * x_j (dev_weights[j] = (1 / total static mass) * factor_j * mass_j, built by the caller in
 * the reference's operation order; the sum runs over the atoms in order, so the centre is the reference's bit for
 * bit), plus host_shift3 (the cell centroid, :56-57; NULL for velocities). */
int sitb_recenter(int device, double* dev_array, int64_t n_frames, int32_t n_atoms, const double* dev_weights,
                  const double* host_shift3, void* cuda_stream);
int sitb_assign_last_known(int device, int64_t* dev_traj, int64_t n_frames, int32_t n_mobile, int64_t frame0,
                           int64_t frame_threshold, const int64_t* dev_carry_label, const int64_t* dev_carry_time,
                           int64_t* dev_end_label, int64_t* dev_end_time, uint64_t* dev_stats4, int32_t apply,
                           void* cuda_stream);
int sitb_windowed_mode(int device, const int64_t* dev_traj, int64_t* dev_out, int64_t n_frames, int32_t n_mobile,
                       int32_t wleft, int32_t wright, int64_t threshold, int32_t replace_no_winner_unknown,
                       int64_t halo_before, const int64_t* dev_before, int64_t halo_after, const int64_t* dev_after,
                       void* cuda_stream);
int sitb_seen_sites(int device, const int64_t* dev_traj, int64_t n_entries, int32_t n_sites, uint32_t* dev_seen,
                    void* cuda_stream);
int sitb_relabel_sites(int device, int64_t* dev_traj, int64_t n_entries, int32_t n_sites, const int64_t* dev_translation,
                       void* cuda_stream);

/* Context-free periodic-boundary helpers for the steps after the path (site merging);
 * host_cellmat = cell^T and its inverse, row major (PBCCalculator.pyx:33-34).
 * sitb_pbc_distances: PBCCalculator.distances (PBCCalculator.pyx:64-103) of every a_i to every b_j -> dev_out [na][nb].
 * sitb_pbc_weighted_average: PBCCalculator.average (PBCCalculator.pyx:106-139) of dev_points [n_points][3] under each
 *   row of dev_weights [n_sets][n_points] (points with weight > 0 take part; centred on the first maximum weight). */
int sitb_pbc_distances(int device, const double* host_cellmat, const double* host_cellmat_inv, const double* dev_a,
                       const double* dev_b, int32_t na, int32_t nb, double* dev_out, void* cuda_stream);
int sitb_pbc_weighted_average(int device, const double* host_cellmat, const double* host_cellmat_inv,
                              const double* dev_points, const double* dev_weights, int32_t n_sets, int32_t n_points,
                              double* dev_out, void* cuda_stream);

/* Tensor-core alternative for the landmark Gram (the covariance input of cluster/mcl.py:53).
 * Staging buffers: two zero-filled fp16 arrays of lpad * ld elements (lpad % 128 == 0, ld % 64 == 0) holding the
 *   transposed landmark vectors as value = hi + lo * 2^-12, stored as contiguous 16 KB tiles: tile (rt, kt) =
 *   landmarks [128 rt, +128) x rows [64 kt, +64) starts at element ((rt * ld/64) + kt) * 8192 and element (r, k)
 *   of a tile sits at r*64 + (((k >> 3) ^ (r & 7)) << 3) + (k & 7)   (the tensor core's 128-byte swizzle).
 * sitb_pass_stage: the fused kernel writes every landmark vector of frames [begin, begin+n) into them (row =
 *   (frame - begin) * n_mobile + mobile index; ld >= n * n_mobile); dev_seen [L] uint64 +=.
 * sitb_gram_syrk_tc: dev_gram_upper [L][L] f64 (upper triangle) += Hi.Hi^T + 2^-12 (Hi.Lo^T + Lo.Hi^T) over the
 *   first k_rows rows, on tcgen05 tensor cores with TMA-fed shared-memory tiles and TMEM accumulators. */
int sitb_pass_stage(sitb_ctx* ctx, int64_t begin, int64_t n, uint64_t* dev_seen, void* dev_stage_hi,
                    void* dev_stage_lo, int64_t ld);
int sitb_gram_syrk_tc(int device, const void* dev_stage_hi, const void* dev_stage_lo, int32_t n_landmarks,
                      int32_t lpad, int64_t ld, int64_t k_rows, double* dev_gram_upper, void* cuda_stream);

/* Pipe micro-benchmarks for bench.py's roofline denominators: device-wide FP32 FMA, FP64 FMA and
 * SFU (ex2) operations per second (one instruction lane = one operation). No reference counterpart. */
int sitb_microbench(int device, double* fp32_ops, double* fp64_ops, double* sfu_ops);

/* Reference-facing, host buffers in and out: what sitator/landmark/helpers.pyx:12 computes.
 * frames [n_frames][n_atoms][3] float64 -> landmark vectors [n_frames*n_mobile][n_landmarks] float64. */
int sitb_fill_landmark_vectors_host(sitb_ctx* ctx, const double* host_frames, int64_t n_frames,
                                    double* host_landmark_vectors, sitb_status* status);

#ifdef __cplusplus
}
#endif
#endif /* SITATOR_B200_H */
