"""Device principal eigenvectors of the landmark clusters' covariance blocks (csrc/sitb_eig.cu) against LAPACK, the
routine the host path uses for cluster/mcl.py:73-80 (the reference calls ARPACK eigsh(k=1); sign arbitrary)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_principal_vectors_match_lapack():
    import torch
    from sitator_b200.landmark.cluster import mcl as gm
    rng = np.random.default_rng(11)
    L = 400
    # non-negative "second moment" matrix with block structure, like cov = LV^T LV / N
    X = np.abs(rng.normal(size=(900, L))) * (rng.random((900, L)) < 0.05)
    cov = X.T @ X / 900.0
    perm = rng.permutation(L)
    sizes = [1, 2, 3, 5, 8, 13, 21, 34, 50, 64, 70, 1, 17]
    clusters, o = [], 0
    for n in sizes:
        clusters.append([int(x) for x in perm[o:o + n]])
        o += n
    covd = torch.as_tensor(cov, device="cuda")
    w = gm.principal_vectors_device(covd, clusters, L)
    for cl in clusters:
        cl = np.asarray(cl)
        want = gm.principal_vector(cov[np.ix_(cl, cl)])
        got = w[cl]
        if np.dot(got, want) < 0:
            got = -got
        assert abs(np.linalg.norm(got) - 1.0) < 1e-14
        assert np.max(np.abs(got - want)) < 1e-12, (len(cl), np.max(np.abs(got - want)))
    untouched = np.setdiff1d(np.arange(L), np.concatenate(clusters))
    assert np.all(w[untouched] == 0.0)


def test_zero_block_and_repeated_eigenvalue():
    import torch
    from sitator_b200.landmark.cluster import mcl as gm
    L = 12
    cov = np.zeros((L, L))
    cov[:4, :4] = np.diag([2.0, 2.0, 1.0, 0.5])            # repeated top eigenvalue: any unit vector of its space
    cov[4:8, 4:8] = 0.0                                   # never-seen landmarks
    cov[8:, 8:] = np.full((4, 4), 0.25) + np.eye(4)
    w = gm.principal_vectors_device(torch.as_tensor(cov, device="cuda"), [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]], L)
    assert abs(np.linalg.norm(w[:4]) - 1) < 1e-14 and np.allclose(w[2:4], 0)
    assert abs(np.linalg.norm(w[4:8]) - 1) < 1e-14 and np.all(np.isfinite(w))
    assert np.allclose(np.abs(w[8:]), 0.5, atol=1e-14)
