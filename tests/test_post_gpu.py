"""GPU parity of the assignment-stream post-processing against the compiled reference's outputs
(tests/golden/postprocess.npz): SiteTrajectory.assign_to_last_known_site, SmoothSiteTrajectory, RemoveUnoccupiedSites."""
import os

import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from tests import _util as U

pytestmark = pytest.mark.gpu


def _site_network(n_sites):
    system, cfg = syn.make_config("toy_bcc")
    sn = syn.site_network_for(system)
    sn.centers = np.zeros((n_sites, 3))
    return sn


@pytest.mark.parametrize("name", ["toy", "synth"])
def test_assign_to_last_known_site_matches_reference(name):
    from sitator_b200 import SiteTrajectory
    g = dict(np.load(os.path.join(U.GOLDEN_DIR, "postprocess.npz"), allow_pickle=False))
    traj, n_sites = g[name + "_traj"], int(g[name + "_n_sites"])
    sn = _site_network(n_sites)
    for thr in (1, 2, 5):
        st = SiteTrajectory(sn, traj.copy())
        info = st.assign_to_last_known_site(frame_threshold=thr)
        assert np.array_equal(st.traj, g["%s_lk%d" % (name, thr)])
        want = g["%s_lk%d_info" % (name, thr)]
        assert info['max_time_unknown'] == int(want[0]) and info['total_reassigned'] == int(want[2])
        assert abs(info['avg_time_unknown'] - want[1]) < 1e-12
    # nothing unknown: the reference's "None to correct." result
    st = SiteTrajectory(sn, np.zeros((10, sn.n_mobile), dtype=np.int64))
    assert st.assign_to_last_known_site() == {'max_time_unknown': 0, 'avg_time_unknown': 0, 'total_reassigned': 0}


@pytest.mark.parametrize("name", ["toy", "synth"])
def test_smooth_site_trajectory_matches_reference(name):
    from sitator_b200 import SiteTrajectory
    from sitator_b200.dynamics import SmoothSiteTrajectory, RemoveUnoccupiedSites
    g = dict(np.load(os.path.join(U.GOLDEN_DIR, "postprocess.npz"), allow_pickle=False))
    traj, n_sites = g[name + "_traj"], int(g[name + "_n_sites"])
    sn = _site_network(n_sites)
    for thr, flag in ((3, True), (4, False), (10, True)):
        st = SiteTrajectory(sn, traj.copy())
        sm = SmoothSiteTrajectory(set_unassigned_under_threshold=flag).run(st, thr)
        assert np.array_equal(sm.traj, g["%s_smooth%d_%d" % (name, thr, int(flag))])
        assert sm.site_network.n_sites == int(g["%s_smooth%d_%d_n_sites" % (name, thr, int(flag))])
        assert np.array_equal(st.traj, traj)                       # the input trajectory is not modified
    # RemoveUnoccupiedSites on its own: a stream that never visits sites 2 and 5
    t2 = np.where(np.isin(traj, (2, 5)), -1, traj)
    st = SiteTrajectory(sn, t2)
    if n_sites - 2 >= sn.n_mobile:
        new_st, kept = RemoveUnoccupiedSites().run(st, return_kept_sites=True)
        assert new_st.site_network.n_sites == len(kept[0]) and 2 not in kept[0] and 5 not in kept[0]
        remap = np.full(n_sites + 1, -1)
        remap[kept[0]] = np.arange(len(kept[0]))
        assert np.array_equal(new_st.traj, remap[t2])
    else:
        from sitator_b200.errors import InsufficientSitesError
        with pytest.raises(InsufficientSitesError):
            RemoveUnoccupiedSites().run(st)


def test_long_stream_against_the_oracle():
    """Chunk boundaries of the last-known-site scan and the window tiles: 3000 frames, unknown stretches across
    chunk edges."""
    from oracle import landmark_oracle as orc
    from sitator_b200 import SiteTrajectory
    from sitator_b200.dynamics import SmoothSiteTrajectory
    rng = np.random.default_rng(3)
    sn = _site_network(20)
    F, M = 3000, sn.n_mobile
    traj = rng.integers(0, 20, (F, M))
    traj = np.where(rng.random((F, M)) < 0.4, -1, traj).astype(np.int64)
    traj[100:420, 0] = -1
    traj[:300, 1] = -1
    for thr in (1, 7, 200):
        st = SiteTrajectory(sn, traj.copy())
        info = st.assign_to_last_known_site(frame_threshold=thr)
        want, winfo = orc.assign_to_last_known_site(traj, thr)
        assert np.array_equal(st.traj, want)
        assert info['max_time_unknown'] == winfo['max_time_unknown'] and info['total_reassigned'] == winfo['total_reassigned']
        assert abs(info['avg_time_unknown'] - winfo['avg_time_unknown']) < 1e-12
    st = SiteTrajectory(sn, traj.copy())
    sm = SmoothSiteTrajectory(remove_unoccupied_sites=False).run(st, 6)
    want = orc.running_windowed_mode(traj, 6, 7, 6, 20, True)
    assert np.array_equal(sm.traj, want)


def test_recenter_trajectory_matches_reference_bit_for_bit():
    from tests.golden.make_golden import recenter_inputs
    from sitator_b200.util.RecenterTrajectory import RecenterTrajectory
    from sitator_b200.structure import Atoms
    g = dict(np.load(os.path.join(U.GOLDEN_DIR, "recenter.npz"), allow_pickle=False))
    system, pos, vel, masses = recenter_inputs()
    structure = Atoms(system.initial_structure_positions(), system.cell, np.where(system.static_mask, 8, 3))
    p, v = pos.copy(), vel.copy()
    assert RecenterTrajectory().run(structure, system.static_mask.copy(), p, velocities=v, masses=masses) is None
    assert np.array_equal(p, g["positions"])
    assert np.array_equal(v, g["velocities"])
