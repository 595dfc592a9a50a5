"""Cluster landmark vectors with the online algorithm of the original landmark-analysis paper.

GPU restatement of the reference plugin ``sitator/landmark/cluster/dotprod.py:11-33`` and the classifier it
drives (``sitator/util/DotProdClassifier.pyx``) behind the same plugin contract.  ``landmark_vectors`` is a
:class:`~sitator_b200.landmark.source.LandmarkVectorSource`; its rows were cached compressed by the first pass.

  fit_centers, first iteration (DotProdClassifier.pyx:228-289): every landmark vector, in row order, joins the
      centre of highest cosine similarity or founds a new one -- sequential by definition; one kernel
      (``sitb_dotprod_fit``) walks the cached rows with the centres in sum form.
  fit_centers, later iterations (:228 loop, :291-313): the same procedure on the few hundred centres the first
      iteration left, until their number stops changing.  A (C, L) problem; done here with NumPy exactly as
      the reference does it, quirks included (an old centre is re-added with weight 1 in the sum but its old
      count in the divisor, :283-286).
  predict (:129-197, predict_normed=True), the min_samples filter (:88-115) and the second predict (:118):
      ``sitb_dotprod_predict`` over the cached rows.

Valid clustering params: ``clustering_threshold`` (0.45) and ``assignment_threshold`` (0.8), as in the reference.
"""
import ctypes as C
import logging

import numpy as np

from ... import _native
from .mcl import (CLUSTERING_CLUSTER_SIZE, CLUSTERING_LABELS, CLUSTERING_CONFIDENCES,
                  CLUSTERING_REPRESENTATIVE_LANDMARKS, _to_host)

logger = logging.getLogger(__name__)

DEFAULT_PARAMS = {
    'clustering_threshold': 0.45,
    'assignment_threshold': 0.8,
}
MAX_CONVERGE_ITERS = 10        # DotProdClassifier.__init__ default, not overridden by the plugin (dotprod.py:21-23)


def first_pass(source):
    """Landmark vectors of all resident frames, cached compressed; runs the lattice / zero-vector checks."""
    eng = source.engine
    if source.sparse is None:
        source.seen, _, source.sparse = eng.pass_stats_cached(want_gram=False)
    return source.sparse


def _refine_centers(centers, n_assigned, threshold):
    """Iterations 2.. of DotProdClassifier.fit_centers (:228-313) on the centres of the first one."""
    old_centers = centers
    old_n = n_assigned
    last_n_sites = len(centers)                    # what the first iteration left (:309)
    for iteration in range(1, MAX_CONVERGE_ITERS):
        cur = [old_centers[0].copy()]
        norms = [np.linalg.norm(cur[0])]
        n_cur = [int(old_n[0])]
        for i in range(1, len(old_centers)):
            vec = old_centers[i]
            with np.errstate(divide='ignore', invalid='ignore'):
                diffs = np.dot(np.asarray(cur), vec)
                diffs = diffs / np.asarray(norms)
                diffs = diffs / np.linalg.norm(vec)
            a = int(np.argmax(diffs))
            if diffs[a] < threshold:
                cur.append(vec.copy())
                n_cur.append(int(old_n[i]))
                norms.append(np.linalg.norm(vec))
            else:
                c = cur[a]
                c *= n_cur[a]
                c += vec
                n_cur[a] += int(old_n[i])
                c /= n_cur[a]
                norms[a] = np.linalg.norm(c)
        old_centers = np.asarray(cur)
        old_n = np.asarray(n_cur, dtype=np.int64)
        if last_n_sites == len(old_centers):
            return old_centers, old_n
        last_n_sites = len(old_centers)
    raise ValueError("Clustering did not converge after %i iterations" % MAX_CONVERGE_ITERS)


def compact_rows(ptr, k, v):
    """Cached rows without the pool's gaps: (entries per row, keys, values), rows in order."""
    import torch
    cnt = ptr & 0xFF
    off = ptr >> 8                                  # offsets stay below 2^55: the sign bit is clear
    new_off = torch.cumsum(cnt, 0) - cnt
    nnz = int(cnt.sum().item())
    idx = torch.repeat_interleave(off - new_off, cnt) + torch.arange(nnz, dtype=torch.int64, device=ptr.device)
    return cnt, k.index_select(0, idx), v.index_select(0, idx)


def gather_rows(rows, comm):
    """All ranks' cached rows on every rank, in global row order (shards are contiguous blocks in rank order).

    fit_centers is sequential over ALL landmark vectors (DotProdClassifier.pyx:236), so a frame-sharded run has no
    parallel form of it: every rank runs the same fit over the same gathered rows and arrives at the same centres
    (the fit is deterministic), which also saves the broadcast.  The compressed rows are ~250 B each."""
    from ...engine import SparseRows
    import torch
    cnt, k, v = compact_rows(rows.ptr[:rows.n_rows], rows.k, rows.v)
    sizes = comm.allgather_numpy(np.array([cnt.numel(), k.numel()], dtype=np.int64))
    cnt = comm.allgather_varlen(cnt, sizes[:, 0])
    k = comm.allgather_varlen(k, sizes[:, 1])
    v = comm.allgather_varlen(v, sizes[:, 1])
    ptr = ((torch.cumsum(cnt, 0) - cnt) << 8) | cnt
    if k.numel() == 0:                              # keep the pointers valid for the kernels
        k = torch.zeros((1,), dtype=k.dtype, device=k.device)
        v = torch.zeros((1,), dtype=v.dtype, device=v.device)
    return SparseRows(ptr, k, v, None, int(sizes[:, 1].sum()), int(sizes[:, 0].sum()), 0)


def fit_centers(source, threshold, rows=None):
    """DotProdClassifier.fit_centers over the cached rows: (centres (C, L) float64, members (C,) int64)."""
    import torch
    eng = source.engine
    rows = source.sparse if rows is None else rows
    lib = _native.load()
    L = eng.L
    stream = torch.cuda.current_stream(eng.device).cuda_stream
    lim_c, lim_e = C.c_int32(), C.c_int32()
    _native.check(lib.sitb_dotprod_limits(C.byref(lim_c), C.byref(lim_e)))
    max_c, cap = min(1024, lim_c.value), 32
    while True:
        sums = torch.zeros((max_c, L), dtype=torch.float64, device=eng.device)
        counts = torch.zeros((max_c,), dtype=torch.int64, device=eng.device)
        norm2 = torch.zeros((max_c,), dtype=torch.float64, device=eng.device)
        lists = torch.empty((L, cap), dtype=torch.int16, device=eng.device)
        llen = torch.zeros((L,), dtype=torch.int16, device=eng.device)
        out3 = torch.zeros((16,), dtype=torch.int64, device=eng.device)
        _native.check(lib.sitb_dotprod_fit(
            eng.device.index, eng._ptr(rows.ptr), eng._ptr(rows.k), eng._ptr(rows.v), rows.n_rows, L, float(threshold),
            max_c, cap, eng._ptr(sums), eng._ptr(counts), eng._ptr(norm2), eng._ptr(lists), eng._ptr(llen),
            eng._ptr(out3), C.c_void_p(stream)))
        diag = out3.cpu().numpy()
        n_c, status, consumed = (int(x) for x in diag[:3])
        if status == 0:
            break
        if status == 1:
            if max_c >= lim_c.value:
                raise ValueError("DotProdClassifier: more than %d cluster centres after %d of %d landmark vectors; "
                                 "raise clustering_threshold's selectivity" % (max_c, consumed, rows.n_rows))
            max_c = min(2 * max_c, lim_c.value)
        else:
            cap *= 2
        logger.debug("dotprod fit: rerun with max_centers=%d, list_cap=%d" % (max_c, cap))
    n_assigned = counts[:n_c].cpu().numpy()
    centers = (sums[:n_c] / counts[:n_c, None].to(torch.float64)).cpu().numpy()
    logger.debug("dotprod fit: %d centres after the pass over %d landmark vectors (longest landmark list %d of %d)"
                 % (n_c, rows.n_rows, int(llen.max().item()), cap))
    logger.debug("dotprod fit: SM cycles per row: candidates %.0f, dot products %.0f, commit %.0f; %.1f candidates per row"
                 % tuple(float(x) / max(rows.n_rows, 1) for x in diag[3:7]))
    logger.debug("dotprod fit: cycles per row from the start of phase 3 to the last barrier, per warp: %s"
                 % " ".join("%.0f" % (float(x) / max(rows.n_rows, 1)) for x in diag[8:16]))
    return _refine_centers(centers, n_assigned, threshold)


def _predict(source, centers, threshold, want_counts):
    """DotProdClassifier.predict(predict_normed=True) over the cached rows (device tensors)."""
    import torch
    eng = source.engine
    rows = source.sparse
    lib = _native.load()
    n_c, L = centers.shape
    normed = centers / np.linalg.norm(centers, axis=1)[:, np.newaxis]                     # :155-157
    # per landmark, the centres that are non-zero there (ascending)
    kk, cc = np.nonzero(normed.T)
    ptr = np.zeros(L + 1, dtype=np.uint32)
    np.cumsum(np.bincount(kk, minlength=L), out=ptr[1:])
    d_normed = torch.as_tensor(np.ascontiguousarray(normed), device=eng.device)
    d_ptr = torch.as_tensor(ptr.view(np.int32), device=eng.device)
    d_cc = torch.as_tensor(np.ascontiguousarray(cc.astype(np.uint16)).view(np.int16), device=eng.device)
    if d_cc.numel() == 0:
        d_cc = torch.zeros((1,), dtype=torch.int16, device=eng.device)
    labels = torch.empty((rows.n_rows,), dtype=torch.int64, device=eng.device)
    confs = torch.empty((rows.n_rows,), dtype=torch.float64, device=eng.device)
    counts = torch.zeros((n_c,), dtype=torch.int64, device=eng.device) if want_counts else None
    stream = torch.cuda.current_stream(eng.device).cuda_stream
    _native.check(lib.sitb_dotprod_predict(
        eng.device.index, eng._ptr(rows.ptr), eng._ptr(rows.k), eng._ptr(rows.v), rows.n_rows, L, n_c,
        eng._ptr(d_normed), eng._ptr(d_ptr), eng._ptr(d_cc), float(threshold), eng._ptr(labels), eng._ptr(confs),
        eng._ptr(counts), C.c_void_p(stream)))
    return labels, confs, counts


def _site_best_table(labels, confs, n_sites, row0):
    """(max confidence, first row) per site in the layout of engine.new_best_table -- the centring point of the
    weighted site-centre average (PBCCalculator.pyx:120-122 via LandmarkAnalysis.py:285)."""
    import torch
    tab = torch.zeros((3 * n_sites,), dtype=torch.int64, device=labels.device)
    valid = labels >= 0
    if bool(valid.any()):
        lab = labels[valid]
        cf = confs[valid]
        rows = torch.nonzero(valid).reshape(-1) + int(row0)
        best = torch.zeros((n_sites,), dtype=torch.float64, device=labels.device).scatter_reduce(0, lab, cf, 'amax')
        at_max = cf == best[lab]
        first = torch.full((n_sites,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=labels.device)
        first = first.scatter_reduce(0, lab[at_max], rows[at_max], 'amin')
        seen = first != torch.iinfo(torch.int64).max
        tab[:n_sites] = best.view(torch.int64)
        tab[n_sites:2 * n_sites] = torch.where(seen, first, torch.zeros_like(first))
    return tab


def do_landmark_clustering(landmark_vectors, clustering_params, min_samples, verbose):
    source = landmark_vectors
    comm = source.comm
    params = DEFAULT_PARAMS.copy()
    params.update(clustering_params)
    first_pass(source)

    # DotProdClassifier.fit_predict (DotProdClassifier.pyx:68-127)
    # frame-sharded: the order-dependent fit runs redundantly over the gathered rows, predict over the local ones
    fit_rows = None if comm is None else gather_rows(source.sparse, comm)
    centers, _ = fit_centers(source, params['clustering_threshold'], fit_rows)             # :83-84
    del fit_rows
    predict_threshold = params['assignment_threshold']
    labels, confs, counts = _predict(source, centers, predict_threshold, True)              # :86
    if comm is not None:
        comm.allreduce_sum_(counts)
    cluster_counts = counts.cpu().numpy()                                                   # :92
    total_n_assigned = int(cluster_counts.sum())                                            # :88
    if isinstance(min_samples, (int, np.integer)):
        ms = int(min_samples)
    else:
        ms = int(np.floor(min_samples * total_n_assigned))                                  # :100
    ms = max(ms, 1)                                                                         # :103
    count_mask = cluster_counts >= ms                                                       # :105
    if not np.any(count_mask):
        raise ValueError("`min_samples` too large; all %i clusters under threshold." % len(count_mask))
    logger.info("DotProdClassifier: %i/%i assignment counts below threshold %s (%s); %i clusters remain." %
                (int(np.sum(~count_mask)), len(count_mask), min_samples, ms, int(np.sum(count_mask))))
    if not np.all(count_mask):
        centers = centers[count_mask]
        labels, confs, _ = _predict(source, centers, predict_threshold, False)              # :118
    kept_counts = cluster_counts[count_mask]            # the first predict's counts are what the reference reports (:108)

    return {
        CLUSTERING_CLUSTER_SIZE: kept_counts,
        CLUSTERING_LABELS: _to_host(labels),
        CLUSTERING_CONFIDENCES: _to_host(confs),
        CLUSTERING_REPRESENTATIVE_LANDMARKS: centers,
        '_dev_labels': labels, '_dev_confs': confs,
        '_dev_site_best': _site_best_table(labels, confs, len(centers), source.row0),
    }
