"""Deterministic Gram (csrc/sitb_gram_sparse.cu, EXACT): cluster/mcl.py:54 accumulated as integer words, so that cov is
bit-identical from run to run and for every frame sharding on 16-frame boundaries ("GPU-count
invariant"), and agrees with an FP64 matmul of the dense landmark vectors."""
import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from . import _util as U

pytestmark = pytest.mark.gpu


def _words(system, frames, frame0=0, **kw):
    import torch
    eng = U.engine_for(system, **kw)
    eng.set_frames(frames, frame0=frame0)
    eng.reset_status()
    seen, words, rows = eng.pass_stats_cached(gram_words=True)
    torch.cuda.synchronize()
    return eng, seen, words


@pytest.mark.parametrize("name,n_frames", [("toy_bcc", 400), ("llzo", 96)])
def test_gram_is_bit_identical_across_runs_and_shardings(name, n_frames):
    import torch
    system, cfg = syn.make_config(name)
    frames = system.trajectory(n_frames)
    kw = dict(dynamic_lattice_mapping=cfg["dynamic"])
    eng, seen, w1 = _words(system, frames, **kw)
    for _ in range(2):
        _, seen2, w2 = _words(system, frames, **kw)
        assert torch.equal(w1, w2) and torch.equal(seen, seen2)
    # shards on multiples of 16 frames (the window length): the integer words add up to the same matrix, bit for bit
    for bounds in ([0, 48, n_frames], [0, 16, 64, 80, n_frames]):
        tot = torch.zeros_like(w1)
        for a, b in zip(bounds[:-1], bounds[1:]):
            tot += _words(system, np.ascontiguousarray(frames[a:b]), frame0=a, **kw)[2]
        assert torch.equal(tot, w1), bounds
    g = eng.gram_words_finish(w1).cpu().numpy()
    # shards off the 16-frame grid: windows differ, every addend is still rounded to 2^-56 only
    tot = torch.zeros_like(w1)
    for a, b in ((0, 37), (37, n_frames)):
        tot += _words(system, np.ascontiguousarray(frames[a:b]), frame0=a, **kw)[2]
    g_off = eng.gram_words_finish(tot).cpu().numpy()
    assert np.max(np.abs(g_off - g)) <= 1e-12 * max(1.0, np.max(g))
    # against the dense FP64 product (cluster/mcl.py:54)
    import torch as _t
    lv = eng.fill_dense(dtype=_t.float64).cpu().numpy()
    want = np.triu(lv.T @ lv)
    assert np.array_equal(g != 0, want != 0)
    nz = want != 0
    assert np.max(np.abs(g[nz] - want[nz]) / want[nz]) < 1e-13
    assert np.all(np.tril(g, -1) == 0)


def test_run_reports_the_gram_path_and_tensor_core_option_is_reachable():
    """round-1 advice: clustering_params={'gram_method': ...} must reach the pass that builds the Gram."""
    from sitator_b200.landmark import LandmarkAnalysis
    g, system, cfg, frames = U.load_golden("toy_bcc_300")
    want = None
    for method in ("sparse", "sparse_atomic", "tcgen05"):
        la = LandmarkAnalysis(clustering_algorithm='mcl', clustering_params={'gram_method': method}, verbose=False,
                              **U.analysis_kwargs(cfg))
        st = la.run(syn.site_network_for(system), frames)
        assert la.stats["gram_method"] == method
        assert np.array_equal(st.traj, g["labels"])
