"""CPU: the NumPy oracle against the golden outputs of the compiled reference (tests/golden/*.npz),
plus known-answer tests derived from the reference code."""
import math

import numpy as np
import pytest

from oracle import landmark_oracle as orc
from tests import _util as U


@pytest.mark.parametrize("name", U.GOLDEN_CASES)
def test_oracle_reproduces_reference_golden(name):
    g, system, cfg, frames = U.load_golden(name)
    kw = U.analysis_kwargs(cfg)
    res = orc.run_landmark_analysis(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                    system.lm_centers, system.lm_vertices, frames,
                                    dynamic_lattice_mapping=kw["dynamic_lattice_mapping"],
                                    check_for_zero_landmarks=kw["check_for_zero_landmarks"],
                                    max_mobile_per_site=kw["max_mobile_per_site"])
    lv = res["landmark_vectors"]
    assert np.array_equal(lv != 0, g["landmark_vectors"] != 0)
    assert np.max(np.abs(lv - g["landmark_vectors"])) < 1e-14
    assert res["n_all_zero_lvecs"] == int(g["n_all_zero_lvecs"])
    assert np.array_equal(res["labels"], g["labels"])              # same site order: same set() logic
    assert np.max(np.abs(res["confs"] - g["confs"])) < 1e-12
    assert np.max(np.abs(res["site_centers"] - g["site_centers"])) < 1e-11
    assert res["site_vertices"] == g["site_vertex_sets"]
    assert res["n_multiple_assignments"] == int(g["n_multiple_assignments"])
    assert abs(res["avg_mobile_per_site"] - float(g["avg_mobile_per_site"])) < 1e-12
    assert np.array_equal(orc.jumps(res["labels"]), g["jumps"])
    assert np.array_equal(orc.jumps(res["labels"], unknown_as_jump=True), g["jumps_unknown_as_jump"])
    ja = orc.jump_analysis(res["labels"], len(res["cluster-size"]))
    for key in ("n_ij", "jump_lag", "residence_times", "occupancy_freqs", "total_corrected_residences"):
        assert np.array_equal(ja[key], g[key]), key
    assert np.array_equal(np.nan_to_num(ja["p_ij"], nan=-1.0), np.nan_to_num(g["p_ij"], nan=-1.0))


@pytest.mark.parametrize("name", U.DOTPROD_GOLDEN_CASES)
def test_oracle_dotprod_clustering_reproduces_reference_golden(name):
    """The constructor-default clustering (DotProdClassifier.fit_centers / predict, landmark/cluster/dotprod.py)."""
    g, system, cfg, frames = U.load_dotprod_golden(name)
    kw = U.analysis_kwargs(cfg)
    lv, n_zero, wrapped = orc.fill_landmark_vectors(
        system.cell, system.static_pos, system.static_idx, system.mobile_idx, system.lm_centers, system.lm_vertices,
        frames, check_for_zeros=kw["check_for_zero_landmarks"], dynamic_lattice_mapping=kw["dynamic_lattice_mapping"])
    res = orc.do_landmark_clustering_dotprod(lv, {}, 0.01 / system.n_mobile)
    labels = res["cluster-labels"].reshape(g["labels"].shape)
    assert len(res["cluster-size"]) == len(g["site_centers"])
    assert np.array_equal(labels, g["labels"])
    assert np.max(np.abs(res["cluster-confs"].reshape(g["confs"].shape) - g["confs"])) < 1e-12
    sc = orc.site_centers_real(orc.PBC(system.cell), wrapped[:, system.mobile_idx], labels,
                               res["cluster-confs"].reshape(labels.shape), len(res["cluster-size"]), weighted=True)
    assert np.max(np.abs(sc - g["site_centers"])) < 1e-11
    assert np.array_equal(orc.jumps(labels), g["jumps"])


def test_oracle_postprocessing_reproduces_reference_golden():
    """assign_to_last_known_site and SmoothSiteTrajectory (+ RemoveUnoccupiedSites) against tests/golden/postprocess.npz."""
    import os
    g = dict(np.load(os.path.join(U.GOLDEN_DIR, "postprocess.npz"), allow_pickle=False))
    for name in ("toy", "synth"):
        traj, n_sites = g[name + "_traj"], int(g[name + "_n_sites"])
        assert np.sum(traj == -1) > 0
        for thr in (1, 2, 5):
            out, info = orc.assign_to_last_known_site(traj, thr)
            assert np.array_equal(out, g["%s_lk%d" % (name, thr)])
            want = g["%s_lk%d_info" % (name, thr)]
            assert info['max_time_unknown'] == int(want[0]) and info['total_reassigned'] == int(want[2])
            assert abs(info['avg_time_unknown'] - want[1]) < 1e-12
        for thr, flag in ((3, True), (4, False), (10, True)):
            out, kept = orc.smooth_site_trajectory(traj, n_sites, thr, set_unassigned_under_threshold=flag)
            assert np.array_equal(out, g["%s_smooth%d_%d" % (name, thr, int(flag))])
            assert int(kept.sum()) == int(g["%s_smooth%d_%d_n_sites" % (name, thr, int(flag))])


def test_oracle_recenter_reproduces_reference_golden():
    """RecenterTrajectory (util/RecenterTrajectory.pyx) bit for bit."""
    import os
    from tests.golden.make_golden import recenter_inputs
    g = dict(np.load(os.path.join(U.GOLDEN_DIR, "recenter.npz"), allow_pickle=False))
    system, pos, vel, masses = recenter_inputs()
    factors = system.static_mask.astype(np.float64)
    centroid = np.sum(0.5 * system.cell, axis=0)
    assert np.array_equal(orc.recenter_trajectory(pos, masses, factors, centroid), g["positions"])
    assert np.array_equal(orc.recenter_trajectory(vel, masses, factors, None), g["velocities"])


def test_oracle_against_live_reference_if_built():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built in this environment")
    ref = ref_loader.load()
    from sitator_b200 import synthetic as syn
    system, cfg = syn.make_config("toy_bcc")
    frames = system.trajectory(80, seed=11)
    pb = ref.PBCCalculator(system.cell)
    op = orc.PBC(system.cell)
    rng = np.random.default_rng(0)
    pts = rng.normal(0, 20, (50, 3))
    a = pts.copy(); pb.wrap_points(a)
    b = op.wrap_points(pts.copy())
    assert np.array_equal(a, b)
    assert np.array_equal(pb.distances(pts[0], pts[1:]), op.distances(pts[0], pts[1:]))
    w = rng.random(49)
    near = system.static_pos[:10] + rng.normal(0, 0.1, (10, 3))
    assert np.array_equal(pb.average(near, weights=w[:10]), op.average(near, weights=w[:10]))


def test_known_answers_from_the_reference_formulas():
    # helpers.pyx:127-131 with midpoint 1.5, steepness 30, threshold 1e-4
    assert orc.cutoff_round_to_zero_point(1.5, 30.0) == pytest.approx(1.8070080122325283, abs=1e-15)
    # a mobile atom exactly on a landmark centre of an ideal lattice: every ratio is 1
    cell = np.eye(3) * 10.0
    static = np.array([[4.0, 5.0, 5.0], [6.0, 5.0, 5.0], [5.0, 4.0, 5.0], [5.0, 6.0, 5.0]])
    centre = np.array([[5.0, 5.0, 5.0]])
    frames = np.concatenate([static, centre])[None]
    lv, nz, _ = orc.fill_landmark_vectors(cell, static, np.arange(4), np.array([4]), centre, [[0, 1, 2, 3]], frames)
    assert lv[0, 0] == pytest.approx(1.0 / (1.0 + math.exp(30 * (1 - 1.5))), rel=1e-15)
    # just inside / outside the short-circuit (helpers.pyx:199-203): the component jumps from ~0.1 to 0
    thr = orc.cutoff_round_to_zero_point(1.5, 30.0)
    for scale, expect_zero in ((thr * (1 - 1e-9), False), (thr * (1 + 1e-9), True)):
        fr = frames.copy()
        fr[0, 4] = [5.0 + (scale - 1.0), 5.0, 5.0]      # distance to atom 0 becomes `scale`
        lv, _, _ = orc.fill_landmark_vectors(cell, static, np.arange(4), np.array([4]), centre, [[0, 1, 2, 3]], fr,
                                             check_for_zeros=False)
        assert (lv[0, 0] == 0.0) == expect_zero
        if not expect_zero:
            assert 0.05 < lv[0, 0] < 0.2
    # wrap of a point on a cell face and in a triclinic cell (PBCCalculator.pyx:341-366)
    p = orc.PBC(cell).wrap_points(np.array([[10.0, -0.0, 25.0]]))
    assert np.allclose(p, [[0.0, 0.0, 5.0]])
    tri = np.array([[9.0, 0.5, 0.0], [1.0, 8.0, 0.2], [0.0, 1.5, 10.0]])
    q = orc.PBC(tri).wrap_points(np.array([[30.0, -7.0, 12.0]]))
    f = q @ np.linalg.inv(tri)
    assert np.all(f >= -1e-12) and np.all(f < 1 + 1e-12)


def test_markov_clustering_returns_blocks():
    g = np.zeros((7, 7))
    g[:3, :3] = 1.0
    g[3:5, 3:5] = 1.0
    g[5:, 5:] = 1.0
    g += 1e-7
    np.fill_diagonal(g, 1.0)
    cl = sorted(tuple(int(x) for x in c) for c in orc.markov_clustering(g, inflation=4))
    assert cl == [(0, 1, 2), (3, 4), (5, 6)]


def test_jump_scan_hand_written_table():
    t = np.array([[-1, 2, 3],
                  [4, 2, -1],
                  [4, 5, -1],
                  [-1, 5, 3],
                  [6, -1, 7]])
    j = orc.jumps(t)
    # a first appearance after an unknown start is a jump from -1 (SiteTrajectory.py:361-373)
    assert j.tolist() == [[1, 0, -1, 4], [2, 1, 2, 5], [4, 0, 4, 6], [4, 2, 3, 7]]
    ju = orc.jumps(t, unknown_as_jump=True)
    assert [1, 2, 3, -1] in ju.tolist() and [3, 0, 4, -1] in ju.tolist()


def test_oracle_general_cell_path_reproduces_reference_golden():
    """Triclinic cell, ragged vertex lists, frames displaced by whole lattice vectors: the compiled reference's
    landmark vectors and PBCCalculator outputs (tests/golden/make_triclinic_golden.py).  The run() goldens are all
    orthorhombic; this pins the oracle's general wrap, which the GPU's triclinic test is checked against."""
    import os
    g = np.load(os.path.join(U.GOLDEN_DIR, "triclinic_fill.npz"))
    t = U.triclinic_system()
    lv, n_zero, _ = orc.fill_landmark_vectors(t["cell"], t["static"], t["static_idx"], t["mobile_idx"], t["centers"],
                                              t["verts"], t["frames"], check_for_zeros=False)
    want = np.zeros(tuple(int(x) for x in g["lv_shape"]))
    want[g["lv_rows"], g["lv_cols"]] = g["lv_vals"]
    assert lv.shape == want.shape
    assert np.array_equal(lv != 0, want != 0)
    nz = want != 0
    assert np.max(np.abs(lv[nz] - want[nz]) / want[nz]) < 1e-14
    assert int(n_zero) == int(g["n_all_zero_lvecs"])
    pb = orc.PBC(t["cell"])
    assert np.array_equal(pb.wrap_points(g["points"].copy()), g["wrapped"])
    assert np.array_equal(pb.distances(g["points"][0], g["points"][1:].copy()), g["distances"])
    assert np.array_equal(pb.average(g["avg_points"].copy(), weights=g["avg_weights"]), g["average"])


def test_oracle_failures_reproduce_reference_golden():
    """Which error, at which frame, naming which atoms: the oracle's failure classes against what the compiled
    reference raised on the same seeded inputs (tests/golden/errors.json, make_error_golden.py)."""
    import json
    import os
    with open(os.path.join(U.GOLDEN_DIR, "errors.json")) as f:
        want = json.load(f)
    for name, (system, frames, kw) in U.error_cases().items():
        w = want[name]
        fill_kw = dict(check_for_zeros=kw.get("check_for_zero_landmarks", True),
                       dynamic_lattice_mapping=kw.get("dynamic_lattice_mapping", False),
                       static_movement_threshold=kw.get("static_movement_threshold", 1.0))
        with pytest.raises((orc.StaticLatticeFailure, orc.ZeroLandmarkFailure)) as e:
            orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                      system.lm_centers, system.lm_vertices, frames, **fill_kw)
        if w["error"] == "StaticLatticeError":
            assert isinstance(e.value, orc.StaticLatticeFailure), name
            assert e.value.frame == w["frame"], name
            assert [int(x) for x in e.value.lattice_atoms] == w["lattice_atoms"], name
        else:
            assert w["error"] == "ZeroLandmarkError" and isinstance(e.value, orc.ZeroLandmarkFailure), name
            assert (e.value.frame, e.value.mobile_index) == (w["frame"], w["mobile_index"]), name
    traj, n_sites = U.occupancy_error_table()
    w = want["multiple_occupancy"]
    with pytest.raises(orc.MultipleOccupancyFailure) as e:
        orc.check_multiple_occupancy(traj, max_mobile_per_site=1)
    assert (e.value.frame, e.value.site) == (w["frame"], w["site"])
    assert [int(x) for x in e.value.mobile] == w["mobile_particles"]
    n_more, avg = orc.check_multiple_occupancy(traj, max_mobile_per_site=traj.shape[1])
    assert n_more == want["multiple_occupancy_stats"]["n_more_than_one"]
    assert avg == want["multiple_occupancy_stats"]["avg_mobile_per_site"]


@pytest.mark.parametrize("name", ["toy_bcc_300", "llzo_60"])
@pytest.mark.parametrize("method", ["real-unweighted", "representative-landmark"])
def test_oracle_site_center_methods_reproduce_reference_golden(name, method):
    """The non-default site_centers_method values (LandmarkAnalysis.py:286-296) against the compiled reference's
    centres (tests/golden/site_center_methods.npz, make_site_center_golden.py)."""
    import os
    want = np.load(os.path.join(U.GOLDEN_DIR, "site_center_methods.npz"))["%s/%s" % (name, method)]
    g, system, cfg, frames = U.load_golden(name)
    kw = U.analysis_kwargs(cfg)
    res = orc.run_landmark_analysis(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                    system.lm_centers, system.lm_vertices, frames, site_centers_method=method,
                                    dynamic_lattice_mapping=kw["dynamic_lattice_mapping"],
                                    check_for_zero_landmarks=kw["check_for_zero_landmarks"],
                                    max_mobile_per_site=kw["max_mobile_per_site"])
    assert res["site_centers"].shape == want.shape
    # compared modulo lattice vectors: the toy landmarks lie ON cell faces, where a last-bit difference in the weights
    # turns the final wrap's 0.0 into 12.0 (the same point; seen once, site 16 of the toy case)
    diff = (res["site_centers"] - want) @ np.linalg.inv(system.cell)
    diff -= np.round(diff)
    assert np.max(np.abs(diff @ system.cell)) < 1e-11
    assert np.sum(np.abs(res["site_centers"] - want) > 1e-11) <= 1


def test_oracle_relaxed_lattice_checks_reproduce_reference_golden():
    """relaxed_lattice_checks=True with dynamic lattice mapping and a static atom that no lattice position picks
    (helpers.pyx:84-92): the run goes on with the duplicated map; landmark vectors of the compiled reference
    (tests/golden/relaxed_dynamic_fill.npz, make_relaxed_golden.py)."""
    import os
    g = np.load(os.path.join(U.GOLDEN_DIR, "relaxed_dynamic_fill.npz"))
    system, frames, kw = U.error_cases()["dynamic_unassigned"]
    lv, n_zero, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                              system.lm_centers, system.lm_vertices, frames, check_for_zeros=False,
                                              dynamic_lattice_mapping=True, relaxed_lattice_checks=True,
                                              static_movement_threshold=kw["static_movement_threshold"])
    want = np.zeros(tuple(int(x) for x in g["lv_shape"]))
    want[g["lv_rows"], g["lv_cols"]] = g["lv_vals"]
    assert np.array_equal(lv != 0, want != 0)
    nz = want != 0
    assert np.max(np.abs(lv[nz] - want[nz]) / want[nz]) < 1e-14
    assert int(n_zero) == int(g["n_all_zero_lvecs"])


@pytest.mark.parametrize("name", ["toy_soft", "llzo_sharp", "llzo_steep"])
def test_oracle_non_default_cutoff_reproduces_reference_golden(name):
    """cutoff_midpoint / cutoff_steepness away from 1.5 / 30 (helpers.pyx:41-43,127-131,197-209): the compiled
    reference's landmark vectors (tests/golden/cutoff_params_fill.npz, make_cutoff_golden.py)."""
    system, frames, midpoint, steepness = U.cutoff_cases()[name]
    want, want_zero = U.load_cutoff_golden(name)
    lv, n_zero, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                              system.lm_centers, system.lm_vertices, frames, midpoint, steepness,
                                              check_for_zeros=False)
    assert np.array_equal(lv != 0, want != 0)
    nz = want != 0
    assert np.max(np.abs(lv[nz] - want[nz]) / want[nz]) < 1e-13
    assert int(n_zero) == want_zero


@pytest.mark.parametrize("case", ["toy_inflation3", "toy_thresholds", "llzo_inflation2"])
def test_oracle_non_default_clustering_params_reproduce_reference_golden(case):
    """'mcl' clustering parameters and minimum_site_occupancy away from their defaults (cluster/mcl.py:28-31,60-66,
    87-88,98-109): labels, confidences and site centres of the compiled reference (clustering_params.npz)."""
    import os
    name, params, min_occ = U.clustering_param_cases()[case]
    g = np.load(os.path.join(U.GOLDEN_DIR, "clustering_params.npz"))
    _, system, cfg, frames = U.load_golden(name)
    kw = U.analysis_kwargs(cfg)
    res = orc.run_landmark_analysis(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                    system.lm_centers, system.lm_vertices, frames, clustering_params=dict(params),
                                    minimum_site_occupancy=min_occ,
                                    dynamic_lattice_mapping=kw["dynamic_lattice_mapping"],
                                    check_for_zero_landmarks=kw["check_for_zero_landmarks"],
                                    max_mobile_per_site=kw["max_mobile_per_site"])
    assert np.array_equal(res["labels"], g[case + "/labels"])
    assert np.max(np.abs(res["confs"] - g[case + "/confs"])) < 1e-12
    assert res["site_centers"].shape == g[case + "/site_centers"].shape
    assert np.max(np.abs(res["site_centers"] - g[case + "/site_centers"])) < 1e-11


@pytest.mark.parametrize("case", ["toy_loose", "toy_tight", "llzo_tight"])
def test_oracle_non_default_dotprod_params_reproduce_reference_golden(case):
    """clustering_threshold / assignment_threshold of the default 'dotprod' clustering and minimum_site_occupancy away
    from their defaults (cluster/dotprod.py:6-9): the compiled reference's labels, confidences, site centres."""
    import os
    name, params, min_occ = U.dotprod_param_cases()[case]
    g = np.load(os.path.join(U.GOLDEN_DIR, "dotprod_params.npz"))
    _, system, cfg, frames = U.load_dotprod_golden(name)
    kw = U.analysis_kwargs(cfg)
    lv, n_zero, wrapped = orc.fill_landmark_vectors(
        system.cell, system.static_pos, system.static_idx, system.mobile_idx, system.lm_centers, system.lm_vertices,
        frames, check_for_zeros=kw["check_for_zero_landmarks"], dynamic_lattice_mapping=kw["dynamic_lattice_mapping"])
    res = orc.do_landmark_clustering_dotprod(lv, dict(params), min_occ / system.n_mobile)
    want = g[case + "/labels"]
    labels = res["cluster-labels"].reshape(want.shape)
    assert len(res["cluster-size"]) == len(g[case + "/site_centers"])
    assert np.array_equal(labels, want)
    confs = res["cluster-confs"].reshape(want.shape)
    assert np.max(np.abs(confs - g[case + "/confs"])) < 1e-12
    sc = orc.site_centers_real(orc.PBC(system.cell), wrapped[:, system.mobile_idx], labels, confs,
                               len(res["cluster-size"]), weighted=True)
    assert np.max(np.abs(sc - g[case + "/site_centers"])) < 1e-11
