"""Markov clustering on the GPU (mirrors reference ``sitator/util/mcl.py:3-60``).

The iteration (normalise, expand, inflate, prune, converge) runs in float64 on the device
(``csrc/sitb_mcl.cu``); reading the clusters off the converged matrix is done here exactly as the
reference does it, including the ``set`` of tuples it returns them through (``util/mcl.py:52-60``),
so the cluster order is the reference's order under the same interpreter.
"""
import ctypes as C

import numpy as np

from .. import _native


def markov_clustering_device(graph, expansion=2, inflation=2, pruning_threshold=0.00001, iterlimit=100):
    """``graph``: (n, n) float64 torch CUDA tensor.  Returns (m2 tensor, n_iterations)."""
    import torch
    lib = _native.load()
    assert graph.is_cuda and graph.dtype == torch.float64 and graph.is_contiguous()
    n = graph.shape[0]
    assert graph.shape == (n, n)
    if int(expansion) != expansion or expansion < 1:
        raise ValueError("expansion must be a positive integer (np.linalg.matrix_power)")
    out = torch.empty_like(graph)
    n_it, conv = C.c_int32(0), C.c_int32(0)
    stream = torch.cuda.current_stream(graph.device).cuda_stream
    _native.check(lib.sitb_markov_clustering(
        graph.device.index, C.c_void_p(graph.data_ptr()), n, int(expansion), float(inflation),
        float(pruning_threshold), int(iterlimit), C.c_void_p(out.data_ptr()), C.byref(n_it), C.byref(conv),
        C.c_void_p(stream)))
    if not conv.value:
        raise ValueError("Markov Clustering couldn't converge in %i iterations" % iterlimit)
    return out, n_it.value


def clusters_from_matrix(m2):
    """util/mcl.py:52-60: attractors = non-zero diagonal; a cluster = non-zero columns of an attractor row."""
    m2 = np.asarray(m2)
    attractors = m2.diagonal().nonzero()[0]
    clusters = set()
    for a in attractors:
        clusters.add(tuple(m2[a].nonzero()[0]))
    return list(clusters)


def markov_clustering(transition_matrix, expansion=2, inflation=2, pruning_threshold=0.00001, iterlimit=100):
    """Drop-in for ``sitator.util.mcl.markov_clustering``: (n, n) array -> list of tuples of indices."""
    import torch
    tm = np.asarray(transition_matrix, dtype=np.float64)
    assert tm.shape[0] == tm.shape[1]
    # self loops are needed to avoid division by zero (util/mcl.py:19-20)
    assert np.count_nonzero(tm.diagonal()) == len(tm)
    if not torch.cuda.is_available():
        raise RuntimeError("sitator_b200 needs a CUDA device; there is no CPU path")
    g = torch.as_tensor(np.ascontiguousarray(tm), device="cuda")
    m2, _ = markov_clustering_device(g, expansion, inflation, pruning_threshold, iterlimit)
    return clusters_from_matrix(m2.cpu().numpy())
