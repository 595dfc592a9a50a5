"""GPU, the less-travelled parameters against goldens of the compiled reference.
relaxed_lattice_checks=True (helpers.pyx:84-92) under dynamic lattice mapping with a static atom that no
lattice position picks: no error, and the landmark vectors are the reference's (relaxed_dynamic_fill.npz; without the
flag the same input raises, tests/test_fill_gpu.py).  Non-default cutoff_midpoint / cutoff_steepness (cutoff_params_fill.npz)."""
import os

import numpy as np
import pytest

from tests import _util as U

pytestmark = pytest.mark.gpu


def test_relaxed_lattice_checks_match_reference_golden():
    import torch
    g = np.load(os.path.join(U.GOLDEN_DIR, "relaxed_dynamic_fill.npz"))
    want = np.zeros(tuple(int(x) for x in g["lv_shape"]))
    want[g["lv_rows"], g["lv_cols"]] = g["lv_vals"]
    system, frames, kw = U.error_cases()["dynamic_unassigned"]
    eng = U.engine_for(system, dynamic_lattice_mapping=True, relaxed_lattice_checks=True,
                       static_movement_threshold=kw["static_movement_threshold"])
    eng.set_frames(frames)
    eng.reset_status()
    got = eng.fill_dense(dtype=torch.float64).cpu().numpy()
    st = eng.status()
    assert st.error_code == 0
    assert st.n_duplicate_nearest >= 1
    assert got.shape == want.shape
    assert np.array_equal(got != 0, want != 0)
    nz = want != 0
    assert np.max(np.abs(got[nz] - want[nz]) / want[nz]) < U.LV_RTOL
    assert st.n_zero_rows == int(g["n_all_zero_lvecs"])


@pytest.mark.parametrize("name", ["toy_soft", "llzo_sharp", "llzo_steep"])
def test_non_default_cutoff_matches_reference_golden(name):
    """cutoff_midpoint / cutoff_steepness away from the defaults: tables, candidate grid, screens and the float64
    values all follow the parameters; landmark vectors of the compiled reference (cutoff_params_fill.npz)."""
    import torch
    system, frames, midpoint, steepness = U.cutoff_cases()[name]
    want, want_zero = U.load_cutoff_golden(name)
    eng = U.engine_for(system, cutoff_midpoint=midpoint, cutoff_steepness=steepness)
    eng.set_frames(frames)
    eng.reset_status()
    got = eng.fill_dense(dtype=torch.float64).cpu().numpy()
    st = eng.status()
    assert st.error_code == 0 and st.n_list_overflow == 0
    assert np.array_equal(got != 0, want != 0)
    nz = want != 0
    assert np.max(np.abs(got[nz] - want[nz]) / want[nz]) < U.LV_RTOL
    assert st.n_zero_rows == want_zero


@pytest.mark.parametrize("case", ["toy_inflation3", "toy_thresholds", "llzo_inflation2"])
def test_non_default_clustering_params_match_reference_golden(case):
    """'mcl' clustering parameters (inflation, assignment / good-site thresholds) and minimum_site_occupancy away
    from their defaults, whole run() against the compiled reference (clustering_params.npz)."""
    from sitator_b200 import synthetic as syn
    from sitator_b200.landmark import LandmarkAnalysis
    name, params, min_occ = U.clustering_param_cases()[case]
    g = np.load(os.path.join(U.GOLDEN_DIR, "clustering_params.npz"))
    _, system, cfg, frames = U.load_golden(name)
    la = LandmarkAnalysis(clustering_algorithm='mcl', clustering_params=dict(params), verbose=False,
                          minimum_site_occupancy=min_occ, **U.analysis_kwargs(cfg))
    st = la.run(syn.site_network_for(system), frames)
    want = g[case + "/labels"]
    assert st.site_network.n_sites == len(g[case + "/site_centers"])
    assert np.array_equal(st.traj, want)
    assert np.max(np.abs(st.confidences - g[case + "/confs"])) < U.CONF_ATOL
    assert np.max(np.abs(np.asarray(st.site_network.centers) - g[case + "/site_centers"])) < U.CENTER_ATOL


@pytest.mark.parametrize("case", ["toy_loose", "toy_tight", "llzo_tight"])
def test_non_default_dotprod_params_match_reference_golden(case):
    """clustering_threshold / assignment_threshold of the default 'dotprod' clustering and minimum_site_occupancy
    away from their defaults, whole run() against the compiled reference (dotprod_params.npz)."""
    from sitator_b200 import synthetic as syn
    from sitator_b200.landmark import LandmarkAnalysis
    name, params, min_occ = U.dotprod_param_cases()[case]
    g = np.load(os.path.join(U.GOLDEN_DIR, "dotprod_params.npz"))
    _, system, cfg, frames = U.load_dotprod_golden(name)
    la = LandmarkAnalysis(clustering_params=dict(params), verbose=False, minimum_site_occupancy=min_occ,
                          **U.analysis_kwargs(cfg))
    assert la._cluster_algo == 'dotprod'
    st = la.run(syn.site_network_for(system), frames)
    assert st.site_network.n_sites == len(g[case + "/site_centers"])
    assert np.array_equal(st.traj, g[case + "/labels"])
    assert np.max(np.abs(st.confidences - g[case + "/confs"])) < U.CONF_ATOL
    assert np.max(np.abs(np.asarray(st.site_network.centers) - g[case + "/site_centers"])) < U.CENTER_ATOL
