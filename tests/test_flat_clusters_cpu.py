"""Host logic of the mcl plugin (no GPU): the flat cluster tables behind the centre filters, and the cluster read-out,
which must hand back the clusters in the order the reference's ``set`` of tuples gives (util/mcl.py:52-60)."""
import numpy as np
import pytest

from sitator_b200.landmark.cluster import mcl as gm
from sitator_b200.util.mcl import clusters_from_matrix


def test_flat_clusters_tables_and_filters():
    L = 50
    clusters = [[3, 7, 9], [1], [20, 21, 22, 23], [40, 45]]
    f = gm._FlatClusters(clusters, L)
    f.weights = np.arange(1, 11, dtype=float)
    cid, w = f.tables()
    assert cid.dtype == np.int32 and cid[7] == 0 and cid[1] == 1 and cid[22] == 2 and cid[45] == 3 and cid[0] == -1
    assert w[3] == 1 and w[9] == 3 and w[1] == 4 and w[23] == 8 and w[45] == 10 and w[0] == 0
    assert list(f.offsets()) == [0, 3, 4, 8, 10]
    with np.errstate(divide='ignore'):
        g = f.keep([True, False, True, True], scale=[2.0, 0.0, 4.0, 1.0])      # mcl.py:94-96: vectors[i] / scale[i], good only
    cid, w = g.tables()
    assert g.clusters == [[3, 7, 9], [20, 21, 22, 23], [40, 45]]
    assert cid[7] == 0 and cid[1] == -1 and cid[22] == 1 and cid[45] == 2
    assert w[3] == 0.5 and w[23] == 2.0 and w[45] == 10 and w[1] == 0
    h = g.keep(np.array([False, True, True]))                                  # DotProdClassifier.pyx:105-118
    assert h.clusters == [[20, 21, 22, 23], [40, 45]] and h.tables()[0][40] == 1 and h.tables()[1][20] == 1.25
    with pytest.raises(ValueError):
        gm._FlatClusters([[1, 2], [2, 3]], L)


def test_cluster_read_out_matches_the_reference_construct():
    import torch
    rng = np.random.default_rng(3)
    n = 120
    m2 = np.zeros((n, n))
    # attractor rows with duplicate clusters (several attractors of one cluster give the same tuple)
    perm = rng.permutation(n)
    groups = np.split(perm, np.sort(rng.choice(np.arange(1, n), 17, replace=False)))
    for g in groups:
        for a in rng.choice(g, size=min(len(g), 2), replace=False):
            m2[a, g] = rng.random(len(g)) + 0.1
    got = gm._clusters_on_device(torch.as_tensor(m2))
    want = clusters_from_matrix(m2)
    assert [tuple(int(x) for x in c) for c in got] == [tuple(int(x) for x in c) for c in want]
