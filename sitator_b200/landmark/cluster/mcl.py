"""Cluster landmarks into sites with Markov clustering, then assign every landmark vector.

GPU restatement of the reference plugin ``sitator/landmark/cluster/mcl.py:43-131`` behind the same
plugin contract (``do_landmark_clustering(landmark_vectors, clustering_params, min_samples, verbose)``
returning the same dict keys, ``LandmarkAnalysis.py:89-93,234-256``).  ``landmark_vectors`` is a
:class:`~sitator_b200.landmark.source.LandmarkVectorSource`: the (N, L) matrix is never stored.

Valid clustering params (as in the reference):
 - ``"assignment_threshold"`` (float in [0, 1]): similarity below which a landmark vector is unassigned.
 - ``"good_site_normed_threshold"``: minimum cosine similarity between a site's unit vector and its
   best matching landmark vector.
 - ``"good_site_projected_threshold"``: minimum inner product of the same pair.
 - ``"gram_method"`` (not in the reference): ``'sparse'`` (exact FP64, default) or ``'tcgen05'`` (tensor cores).
 - everything else goes to :func:`sitator_b200.util.mcl.markov_clustering_device`.

Device passes (A is a launch of the fused kernel K1 over the resident frames and also caches every
landmark vector in compressed form; B-D then stream those ~230-byte rows, or rerun K1 if they were not cached):
  A  seen counts + Gram                      -> cov, correlation graph, MCL, per-cluster eigenvectors
  B  best matching landmark vector per cluster (max |centre . x|, first row)       (mcl.py:81-89)
  C  predict + bincount                      -> min_samples filter                  (DotProdClassifier.pyx:86-108)
  D  predict with the kept centres + representative landmark vectors + per-site max-confidence row
     (C already produces D's outputs; D only reruns when the filter dropped a cluster)
"""
import ctypes as C
import logging

import numpy as np

from ... import _native
from ...engine import new_best_table, read_best_table
from ...util.mcl import markov_clustering_device, clusters_from_matrix

logger = logging.getLogger(__name__)

DEFAULT_PARAMS = {
    'inflation': 4,
    'assignment_threshold': 0.7,
}

CLUSTERING_CLUSTER_SIZE = 'cluster-size'
CLUSTERING_LABELS = 'cluster-labels'
CLUSTERING_CONFIDENCES = 'cluster-confs'
CLUSTERING_LANDMARK_GROUPINGS = 'cluster-landmark-groupings'
CLUSTERING_REPRESENTATIVE_LANDMARKS = 'cluster-representative-lvecs'


def principal_vector(block):
    """Eigenvector of the largest eigenvalue of a (small, PSD) covariance block -- what
    ``eigsh(block, k=1)`` returns in the reference (mcl.py:74-79); its sign is arbitrary there and
    every consumer takes ``abs`` (mcl.py:82,85; DotProdClassifier.pyx:179)."""
    if block.shape[0] == 1:
        return np.array([1.0])
    w, v = np.linalg.eigh(block)
    return v[:, -1]


def principal_vectors(blocks):
    """:func:`principal_vector` of every block, with one LAPACK gufunc call per block size instead of one Python
    call per block (same routine per matrix, so the same vectors bit for bit)."""
    out = [None] * len(blocks)
    by_size = {}
    for i, b in enumerate(blocks):
        by_size.setdefault(b.shape[0], []).append(i)
    for n, idx in by_size.items():
        if n == 1:
            for i in idx:
                out[i] = np.array([1.0])
            continue
        w, v = np.linalg.eigh(np.stack([blocks[i] for i in idx]))
        for pos, i in enumerate(idx):
            out[i] = v[pos, :, -1]
    return out


def landmark_graph(source, gram_method='sparse'):
    """Pass A + mcl.py:53-59.  Returns (seen (L,) int64 numpy, cov (L,L) numpy, graph torch tensor).

    ``gram_method``: ``'sparse'`` (default) accumulates the Gram in FP64 products from the ~1.5 % dense rows, added
    as integers so that the result is independent of the order of addition (``'sparse_atomic'``: FP64 atomics), and
    caches the rows; ``'tcgen05'`` runs it as a dense SYRK on the tensor cores (fp16 hi/lo operands, FP32 TMEM
    accumulators drained to FP64; agrees to ~1e-6, see tests/test_gram_tc_gpu.py) and leaves passes B-D to the
    fused fill+assign kernel."""
    import torch
    eng = source.engine
    lib = _native.load()
    if gram_method not in ('sparse', 'sparse_atomic', 'tcgen05'):
        raise ValueError("gram_method must be 'sparse', 'sparse_atomic' or 'tcgen05', not %r" % (gram_method,))
    if source.gram_upper is None:
        # keep the rows compressed for the later passes when they fit comfortably (~400 B per row)
        # (torch's cached totals, not cudaMemGetInfo: that driver query can take tens of milliseconds)
        free_bytes = (torch.cuda.get_device_properties(eng.device).total_memory
                      - torch.cuda.memory_reserved(eng.device) - eng.frames_bytes)
        if gram_method == 'tcgen05':
            seen, gram = eng.pass_stats_tc()
        elif source.cache_rows and source.n_local * 420 < 0.5 * free_bytes:
            # 'sparse': deterministic integer accumulation (order independent: equal for every run and sharding);
            # 'sparse_atomic': FP64 atomics (round 1)
            seen, gram, source.sparse = eng.pass_stats_cached(gram_words=(gram_method == 'sparse'))
        else:
            seen, gram = eng.pass_stats()
        if source.comm is not None:
            source.comm.allreduce_sum_(seen)
            source.comm.allreduce_sum_(gram)               # integer words when deterministic: the sum is exact
        if gram.dtype == torch.int64:
            gram = eng.gram_words_finish(gram)
        source.seen, source.gram_upper = seen, gram
        source.gram_method = gram_method
    L = eng.L
    cov = torch.empty((L, L), dtype=torch.float64, device=eng.device)
    graph = torch.empty((L, L), dtype=torch.float64, device=eng.device)
    stream = torch.cuda.current_stream(eng.device).cuda_stream
    _native.check(lib.sitb_landmark_graph(eng.device.index, C.c_void_p(source.gram_upper.data_ptr()), L,
                                          float(source.n_total), C.c_void_p(cov.data_ptr()),
                                          C.c_void_p(graph.data_ptr()), C.c_void_p(stream)))
    return source.seen.cpu().numpy(), cov, graph


def _clusters_on_device(m2):
    """util/mcl.py:52-60 with only the attractor rows leaving the device.  Same construct as the reference (a ``set`` of
    tuples of column indices, in attractor order), so the clusters come out in the reference's order under the same
    interpreter; the tuples are cut from one ``nonzero`` of the whole block instead of one call per row."""
    import torch
    attractors = torch.nonzero(torch.diagonal(m2) != 0).reshape(-1)
    # (row, column) of every non-zero of the attractor rows, row-major: ascending columns within each row.  Found on the
    # device: a few KB cross PCIe instead of the rows, and NumPy's 2-D nonzero alone took 0.8 ms here.
    nz = torch.nonzero(m2.index_select(0, attractors)).cpu().numpy()
    ri, ci = nz[:, 0], nz[:, 1]
    ends = np.cumsum(np.bincount(ri, minlength=int(attractors.shape[0])))
    cols = ci.tolist()
    clusters = set()
    beg = 0
    for end in ends.tolist():
        clusters.add(tuple(cols[beg:end]))
        beg = end
    return list(clusters)


def _covariance_blocks(cov, clusters):
    """cov[cl][:, cl] for every cluster (mcl.py:78), gathered on the device in one indexing call."""
    import torch
    ri, ci, sizes = [], [], []
    for cl in clusters:
        cl = np.asarray(cl, dtype=np.int64)
        ri.append(np.repeat(cl, len(cl)))
        ci.append(np.tile(cl, len(cl)))
        sizes.append(len(cl))
    ri = torch.as_tensor(np.concatenate(ri), device=cov.device)
    ci = torch.as_tensor(np.concatenate(ci), device=cov.device)
    flat = cov[ri, ci].cpu().numpy()
    out, o = [], 0
    for n in sizes:
        out.append(flat[o:o + n * n].reshape(n, n))
        o += n * n
    return out


class _FlatClusters(object):
    """The landmark clusters as flat arrays (members, sizes, one weight per member): the per-landmark centre tables the
    kernels read, the good-site and min_samples filters and the rescaling are then a few vector operations instead of
    Python loops over the clusters (which took longer than the device passes between them)."""

    def __init__(self, clusters, L):
        import itertools
        self.clusters = clusters
        self.L = L
        self.sizes = np.fromiter((len(c) for c in clusters), dtype=np.int64, count=len(clusters))
        self.members = np.fromiter(itertools.chain.from_iterable(clusters), dtype=np.int64, count=int(self.sizes.sum()))
        if len(self.members) and np.bincount(self.members, minlength=L).max() > 1:
            raise ValueError("landmark clusters overlap; the sparse centre representation needs disjoint clusters")
        self.weights = np.zeros(len(self.members), dtype=np.float64)

    def __len__(self):
        return len(self.clusters)

    def offsets(self):
        out = np.zeros(len(self.sizes) + 1, dtype=np.int64)
        np.cumsum(self.sizes, out=out[1:])
        return out

    def tables(self):
        """(cluster of landmark (L,) int32 with -1 = none, centre weight (L,) float64)."""
        cid = np.full(self.L, -1, dtype=np.int32)
        w = np.zeros(self.L, dtype=np.float64)
        cid[self.members] = np.repeat(np.arange(len(self.sizes), dtype=np.int32), self.sizes)
        w[self.members] = self.weights
        return cid, w

    def keep(self, mask, scale=None):
        """The clusters where ``mask``; weights divided by ``scale`` (per cluster) first if given."""
        mask = np.asarray(mask, dtype=bool)
        out = _FlatClusters.__new__(_FlatClusters)
        out.L = self.L
        out.clusters = [c for c, m in zip(self.clusters, mask.tolist()) if m]
        member_mask = np.repeat(mask, self.sizes)
        weights = self.weights if scale is None else self.weights / np.repeat(np.asarray(scale, dtype=np.float64), self.sizes)
        out.sizes = self.sizes[mask]
        out.members = self.members[member_mask]
        out.weights = weights[member_mask]
        return out


def principal_vectors_device(cov, clusters, L, flat=None):
    """mcl.py:73-80 on the device (csrc/sitb_eig.cu): per-landmark centre weights (L,) float64 numpy -- each cluster's
    unit principal eigenvector scattered onto its landmarks (``flat.weights`` is set too if ``flat`` is given).  Blocks
    larger than the kernel's limit use LAPACK."""
    import torch
    lib = _native.load()
    if flat is None:
        flat = _FlatClusters([list(c) for c in clusters], L)
    n_c = len(flat)
    offsets = flat.offsets().astype(np.int32)
    dev = cov.device
    packed = torch.as_tensor(np.concatenate([offsets, flat.members.astype(np.int32)]), device=dev)
    w = torch.zeros((L,), dtype=torch.float64, device=dev)
    sweeps = torch.zeros((n_c,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _native.check(lib.sitb_principal_vectors(dev.index, C.c_void_p(cov.data_ptr()), L,
                                             C.c_void_p(packed.data_ptr() + 4 * len(offsets)), C.c_void_p(packed.data_ptr()),
                                             n_c, C.c_void_p(w.data_ptr()), C.c_void_p(sweeps.data_ptr()),
                                             C.c_void_p(stream)))
    out = torch.empty((L + n_c,), dtype=torch.float64, device=dev)
    out[:L] = w
    out[L:] = sweeps.to(torch.float64)
    h = out.cpu().numpy()                       # one small device -> host copy
    w_h, sw = h[:L].copy(), h[L:]
    for i in np.nonzero(sw < 0)[0]:             # blocks beyond the kernel's size limit
        cl = np.asarray(flat.clusters[i], dtype=np.int64)
        blk = cov[torch.as_tensor(cl, device=dev)][:, torch.as_tensor(cl, device=dev)].cpu().numpy()
        w_h[cl] = principal_vector(blk)
    flat.weights = w_h[flat.members]
    return w_h


def _to_host(t):
    """Device -> host into page-locked memory (a pageable .cpu() runs at a fraction of PCIe speed).  The
    returned array owns the pinned block; torch's caching host allocator recycles it once the caller
    drops the result, so repeated analyses do not pay for page-locking again."""
    import torch
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


_COPY_STREAMS = {}


def _to_host_async(tensors):
    """Start device -> page-locked host copies on a side stream (after everything queued on the current stream).
    Returns (numpy arrays, event): the arrays are valid once the event has been synchronised."""
    import torch
    dev = tensors[0].device
    side = _COPY_STREAMS.get(dev.index)
    if side is None:
        side = _COPY_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(dev))
    outs = []
    with torch.cuda.stream(side):
        side.wait_event(ready)
        for t in tensors:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            t.record_stream(side)
            outs.append(h)
        done = torch.cuda.Event()
        done.record(side)
    return [h.numpy() for h in outs], done


def do_landmark_clustering(landmark_vectors, clustering_params, min_samples, verbose):
    import torch
    from ...util.phases import PhaseTimer
    source = landmark_vectors
    eng = source.engine
    comm = source.comm
    L = eng.L
    params = DEFAULT_PARAMS.copy()
    params.update(clustering_params)
    timer = getattr(source, "timer", None) or PhaseTimer()

    with timer.phase("  graph: cov, corr, clip"):
        seen_ntimes, cov, graph = landmark_graph(source, params.pop('gram_method', 'sparse'))

    predict_threshold = params.pop('assignment_threshold')
    good_site_normed_threshold = params.pop('good_site_normed_threshold', predict_threshold)
    good_site_project_thresh = params.pop('good_site_projected_threshold', predict_threshold)
    weighted_reps = params.pop('weighted_representative_landmarks', True)
    eig_where = params.pop('eigenvectors', 'device')      # not in the reference: 'host' = LAPACK on gathered blocks

    # -- cluster landmarks (mcl.py:66-68)
    with timer.phase("  Markov clustering"):
        m2, n_iter = markov_clustering_device(graph, **params)
    with timer.phase("  cluster read-out (D2H)"):
        clusters = _clusters_on_device(m2)
    logger.debug("Markov clustering converged in %i iterations: %i clusters" % (n_iter, len(clusters)))
    clusters = [list(c) for c in clusters if seen_ntimes[c[0]] > 0]
    n_clusters = len(clusters)
    if n_clusters == 0:
        raise ValueError("Markov clustering found no landmark cluster that was ever seen")

    # -- centres: principal eigenvector of each cluster's covariance block (mcl.py:73-80)
    with timer.phase("  centres: eigenvectors of the covariance blocks"):
        flat = _FlatClusters(clusters, L)
        if eig_where == 'host':
            vectors = principal_vectors(_covariance_blocks(cov, clusters))
            flat.weights = np.concatenate(vectors) if len(vectors) else np.zeros(0)
        else:
            principal_vectors_device(cov, clusters, L, flat=flat)
        cid, w = flat.tables()

    # -- pass B: best matching landmark vector per cluster (mcl.py:81-89)
    with timer.phase("  pass B: best row per cluster (collective)"):
        eng.set_centers(cid, w, n_clusters)
        best = new_best_table(n_clusters, eng.device)
        source.assign(float('nan'), best=best)
        best_vals, best_rows = read_best_table(best, comm)
    with timer.phase("  best rows: norms (collective)"):
        # best_vals[i] = |best lvec . centre| (mcl.py:84-85) came out of the pass; only the row's norm (:86) is missing
        best_norms = np.sqrt(source.row_norms2(best_rows))
    best_match_dot = np.asarray(best_vals, dtype=np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        best_match_dot_norm = best_match_dot / best_norms
    good = (best_match_dot_norm >= good_site_normed_threshold) & (best_match_dot >= good_site_project_thresh)
    logger.debug("Kept %i/%i landmark clusters as good sites" % (int(np.sum(good)), len(good)))

    # -- keep the good sites, centres scaled by their best match (mcl.py:94-96)
    with np.errstate(divide='ignore', invalid='ignore'):
        flat = flat.keep(good, scale=best_match_dot)
    clusters = flat.clusters
    n_clusters = len(clusters)
    if n_clusters == 0:
        raise ValueError("`min_samples` too large; all 0 clusters under threshold.")

    # -- pass C: predict + bincount, then the min_samples filter (DotProdClassifier.pyx:86-115).  The pass also
    # writes everything the final predict needs: when the filter removes no cluster (the usual case) the
    # centres of pass D are the same and its outputs would be bit-identical, so pass D is skipped.
    cid, w = flat.tables()
    eng.set_centers(cid, w, n_clusters)
    N = source.n_local

    def final_predict(n_c, with_counts):
        out = dict(labels=torch.empty((N,), dtype=torch.int64, device=eng.device),
                   confs=torch.empty((N,), dtype=torch.float64, device=eng.device),
                   rep=torch.zeros((n_c, L), dtype=torch.float64, device=eng.device),
                   rep_w=torch.zeros((n_c,), dtype=torch.float64, device=eng.device),
                   site_best=new_best_table(n_c, eng.device))
        if with_counts:
            out['counts'] = torch.zeros((n_c,), dtype=torch.int64, device=eng.device)
        source.assign(predict_threshold, **out)
        return out

    with timer.phase("  pass C: predict + counts (collective)"):
        res = final_predict(n_clusters, True)
        counts = res.pop('counts')
        if comm is not None:
            comm.allreduce_sum_(counts)
        cluster_counts = counts.cpu().numpy()
    total_n_assigned = int(cluster_counts.sum())
    if isinstance(min_samples, (int, np.integer)):
        ms = int(min_samples)
    else:
        ms = int(np.floor(min_samples * total_n_assigned))
    ms = max(ms, 1)
    count_mask = cluster_counts >= ms
    if not np.any(count_mask):
        raise ValueError("`min_samples` too large; all %i clusters under threshold." % len(count_mask))
    logger.info("DotProdClassifier: %i/%i assignment counts below threshold %s (%s); %i clusters remain." %
                (int(np.sum(~count_mask)), len(count_mask), min_samples, ms, int(np.sum(count_mask))))
    flat = flat.keep(count_mask)
    clusters = flat.clusters
    kept_counts = cluster_counts[count_mask]
    n_sites = len(clusters)

    # -- pass D: final predict + representative landmark vectors (mcl.py:114-122) + per-site best row
    if not np.all(count_mask):
        with timer.phase("  pass D: predict with the kept centres"):
            cid, w = flat.tables()
            eng.set_centers(cid, w, n_sites)
            if source.sparse is not None:
                # the surviving centres are unchanged: rows of surviving clusters keep arg-max and confidence, only
                # the rows of removed clusters are predicted again (sitb_relabel_select / sitb_assign_sparse_rows)
                keep = torch.as_tensor(np.nonzero(count_mask)[0], device=eng.device)
                n_old = len(count_mask)
                sb_old = res['site_best']
                res = dict(labels=res['labels'], confs=res['confs'], rep=res['rep'].index_select(0, keep).contiguous(),
                           rep_w=res['rep_w'].index_select(0, keep).contiguous(),
                           site_best=torch.cat([sb_old[:n_old].index_select(0, keep), sb_old[n_old:2 * n_old].index_select(0, keep),
                                                torch.zeros((n_sites,), dtype=torch.int64, device=eng.device)]))
                remap = np.full(n_old, -1, dtype=np.int32)
                remap[count_mask] = np.arange(n_sites, dtype=np.int32)
                eng.repredict_removed(source.sparse, predict_threshold, remap, res['labels'], res['confs'],
                                      rep=res['rep'], rep_w=res['rep_w'], site_best=res['site_best'])
            else:
                res = final_predict(n_sites, False)
    labels, confs, rep, rep_w, site_best = res['labels'], res['confs'], res['rep'], res['rep_w'], res['site_best']
    with timer.phase("  representative vectors (collective)"):
        if comm is not None:
            comm.allreduce_sum_(rep)
            comm.allreduce_sum_(rep_w)
        if weighted_reps:
            reps = (rep / rep_w[:, None]).cpu().numpy()
        else:
            raise NotImplementedError("weighted_representative_landmarks=False (the reference cannot reach it either: "
                                      "the key is forwarded to markov_clustering, mcl.py:66,115)")
    d2h_done = None
    n_unassigned = None
    with timer.phase("  labels + confidences D2H"):
        if getattr(source, "defer_d2h", False):
            # (counted before the big copy starts: a scalar read queued behind it would wait for all of it)
            n_unassigned = int((labels < 0).sum().item())
            if comm is not None:
                n_unassigned = comm.allreduce_sum_scalar(n_unassigned)
            # LandmarkAnalysis.run goes on with device work (site centres, occupancy check) while the 16 bytes per
            # landmark vector cross PCIe on a side stream; it waits for `_d2h_done` before it returns
            (host_labels, host_confs), d2h_done = _to_host_async([labels, confs])
        else:
            host_labels, host_confs = _to_host(labels), _to_host(confs)

    return {
        CLUSTERING_CLUSTER_SIZE: kept_counts,
        CLUSTERING_LABELS: host_labels,
        CLUSTERING_CONFIDENCES: host_confs,
        CLUSTERING_LANDMARK_GROUPINGS: clusters,
        CLUSTERING_REPRESENTATIVE_LANDMARKS: reps,
        # device-side copies for the rest of LandmarkAnalysis.run (not part of the reference contract)
        '_dev_labels': labels, '_dev_confs': confs, '_dev_site_best': site_best,
        '_centers': (cid, w), '_mcl_iterations': n_iter, '_d2h_done': d2h_done,
        '_n_unassigned': n_unassigned,
    }
