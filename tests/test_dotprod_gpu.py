"""GPU parity of the constructor-default clustering, clustering_algorithm='dotprod' (landmark/cluster/dotprod.py,
util/DotProdClassifier.pyx): whole LandmarkAnalysis.run against the golden outputs of the compiled reference."""
import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from tests import _util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", U.DOTPROD_GOLDEN_CASES)
def test_default_clustering_matches_reference_golden(name):
    from oracle import landmark_oracle as orc
    from sitator_b200.landmark import LandmarkAnalysis
    g, system, cfg, frames = U.load_dotprod_golden(name)
    la = LandmarkAnalysis(verbose=False, **U.analysis_kwargs(cfg))            # default algorithm, as in the reference
    assert la._cluster_algo == 'dotprod'
    st = la.run(syn.site_network_for(system), frames)
    out_sn = st.site_network
    assert out_sn.n_sites == len(g["site_centers"])
    # decision margins of the reference's own final predict: a label may differ only inside TIE_TOL
    kw = U.analysis_kwargs(cfg)
    lv, _, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                         system.lm_centers, system.lm_vertices, frames, check_for_zeros=False,
                                         dynamic_lattice_mapping=kw["dynamic_lattice_mapping"])
    res = orc.do_landmark_clustering_dotprod(lv, {}, 0.01 / system.n_mobile)
    centers = res["cluster-representative-lvecs"]
    normed = centers / np.linalg.norm(centers, axis=1)[:, None]
    with np.errstate(divide='ignore', invalid='ignore'):
        dots = np.abs(lv @ normed.T) / np.linalg.norm(lv, axis=1)[:, None]
    dots[~lv.any(axis=1)] = 0.0
    srt = np.sort(dots, axis=1)
    margin = np.minimum(srt[:, -1] - srt[:, -2], np.abs(srt[:, -1] - 0.8)) if dots.shape[1] > 1 else np.abs(srt[:, -1] - 0.8)
    margin[~lv.any(axis=1)] = 1.0
    n_diff = U.compare_labels(st.traj, g["labels"], margin)
    same = (st.traj == g["labels"]).reshape(-1)
    assert np.max(np.abs(st.confidences.reshape(-1)[same] - g["confs"].reshape(-1)[same])) < U.CONF_ATOL
    assert np.max(np.abs(np.asarray(out_sn.centers) - g["site_centers"])) < U.CENTER_ATOL
    if n_diff == 0:
        assert la.n_multiple_assignments == int(g["n_multiple_assignments"])
        assert np.array_equal(st.jump_array(), g["jumps"])


def test_fit_centers_matches_oracle_on_a_longer_trajectory():
    """The sequential fit over more rows (toy cell, 2000 frames = 32000 rows incl. all-zero rows): the same
    centres, member counts and labels as the NumPy restatement."""
    import torch
    from oracle import landmark_oracle as orc
    from sitator_b200.landmark.source import LandmarkVectorSource
    from sitator_b200.landmark.cluster import dotprod
    system, cfg = syn.make_config("toy_bcc")
    frames = system.trajectory(2000)
    lv, _, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                         system.lm_centers, system.lm_vertices, frames, check_for_zeros=False)
    assert (~lv.any(axis=1)).sum() > 0
    want_c, want_n = orc.dotprod_fit_centers(lv, 0.45)
    eng = U.engine_for(system)
    eng.set_frames(frames)
    src = LandmarkVectorSource(eng)
    dotprod.first_pass(src)
    got_c, got_n = dotprod.fit_centers(src, 0.45)
    assert got_c.shape == want_c.shape
    assert np.array_equal(got_n, want_n)
    assert np.array_equal(got_c != 0, want_c != 0)
    assert np.max(np.abs(got_c - want_c)) < 1e-12


@pytest.mark.parametrize("threshold", [0.9, 0.999])
def test_fit_centers_many_centres_and_growing_tables(threshold):
    """A high clustering threshold makes hundreds to thousands of centres: the per-landmark centre lists and the
    centre table outgrow their first sizes and the fit reruns with larger ones; the result is still the oracle's."""
    from oracle import landmark_oracle as orc
    from sitator_b200.landmark.source import LandmarkVectorSource
    from sitator_b200.landmark.cluster import dotprod
    system, cfg = syn.make_config("llzo")
    frames = system.trajectory(40)
    lv, _, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                         system.lm_centers, system.lm_vertices, frames, check_for_zeros=False)
    want_c, want_n = orc.dotprod_fit_centers(lv, threshold)
    eng = U.engine_for(system)
    eng.set_frames(frames)
    src = LandmarkVectorSource(eng)
    dotprod.first_pass(src)
    got_c, got_n = dotprod.fit_centers(src, threshold)
    assert got_c.shape == want_c.shape, (got_c.shape, want_c.shape)
    assert np.array_equal(got_n, want_n)
    assert np.array_equal(got_c != 0, want_c != 0)
    assert np.max(np.abs(got_c - want_c)) < 1e-12


def test_fit_centers_when_the_first_row_is_all_zero():
    """NumPy's arg-max over NaN similarities sends every row to cluster 0 while centre 0 is the zero vector
    (DotProdClassifier.pyx:241-248): a trajectory whose very first landmark vector is all zero exercises that."""
    from oracle import landmark_oracle as orc
    from sitator_b200.landmark.source import LandmarkVectorSource
    from sitator_b200.landmark.cluster import dotprod
    system, cfg = syn.make_config("toy_bcc")
    full = system.trajectory(300)
    lv_full, _, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                              system.lm_centers, system.lm_vertices, full, check_for_zeros=False)
    zero_rows = np.where(~lv_full.any(axis=1))[0]
    f, j = int(zero_rows[0]) // system.n_mobile, int(zero_rows[0]) % system.n_mobile
    frames = full[:120].copy()
    frames[0, system.mobile_idx[0]] = full[f, system.mobile_idx[j]]           # mobile 0 of frame 0 <- that stray atom
    frames[0, system.static_idx] = full[f, system.static_idx]
    lv, _, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                         system.lm_centers, system.lm_vertices, frames, check_for_zeros=False)
    assert not lv[0].any()
    want_c, want_n = orc.dotprod_fit_centers(lv, 0.45)
    eng = U.engine_for(system)
    eng.set_frames(frames)
    src = LandmarkVectorSource(eng)
    dotprod.first_pass(src)
    got_c, got_n = dotprod.fit_centers(src, 0.45)
    assert got_c.shape == want_c.shape
    assert np.array_equal(got_n, want_n)
    assert np.max(np.abs(got_c - want_c)) < 1e-12
