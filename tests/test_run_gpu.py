"""GPU parity: the whole LandmarkAnalysis.run + jump extraction + JumpAnalysis, through the public API,
against the golden outputs of the compiled reference (tests/golden/*.npz)."""
import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from tests import _util as U

pytestmark = pytest.mark.gpu


def _run(system, cfg, frames, **over):
    from sitator_b200.landmark import LandmarkAnalysis
    sn = syn.site_network_for(system)
    kw = U.analysis_kwargs(cfg)
    kw.update(over)
    la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **kw)
    st = la.run(sn, frames)
    return la, st


@pytest.mark.parametrize("name", U.GOLDEN_CASES)
def test_run_matches_reference_golden(name):
    from oracle import landmark_oracle as orc
    g, system, cfg, frames = U.load_golden(name)
    frames_before = frames.copy()
    la, st = _run(system, cfg, frames)
    assert np.array_equal(frames, frames_before), "run() must not modify the caller's frames"
    assert st.real_trajectory is frames

    # landmark vectors (lazy property): support bit-exact, values to LV_RTOL
    lv = np.asarray(la.landmark_vectors)
    want_lv = g["landmark_vectors"]
    assert np.array_equal(lv != 0, want_lv != 0)
    nz = want_lv != 0
    assert np.max(np.abs(lv[nz] - want_lv[nz]) / want_lv[nz]) < U.LV_RTOL
    assert la.n_all_zero_lvecs == int(g["n_all_zero_lvecs"])
    assert la.landmark_dimension == system.n_landmarks

    # sites: same number, same landmark->vertex unions, same order (the set() logic is shared)
    out_sn = st.site_network
    assert out_sn.n_sites == len(g["site_centers"])
    assert [set(v) for v in out_sn.vertices] == g["site_vertex_sets"]

    # assignments: exact except decisions whose reference margin is below TIE_TOL (counted)
    kw = U.analysis_kwargs(cfg)
    res = orc.do_landmark_clustering_mcl(want_lv, {}, 0.01 / system.n_mobile)
    dots = np.abs(want_lv @ res["_centers"].T)
    srt = np.sort(dots, axis=1)
    margin = np.minimum(srt[:, -1] - srt[:, -2], np.abs(srt[:, -1] - 0.7))
    n_diff = U.compare_labels(st.traj, g["labels"], margin)
    same = (st.traj == g["labels"]).reshape(-1)
    assert np.max(np.abs(st.confidences.reshape(-1)[same] - g["confs"].reshape(-1)[same])) < U.CONF_ATOL
    assert np.max(np.abs(np.asarray(out_sn.centers) - g["site_centers"])) < U.CENTER_ATOL
    if n_diff == 0:
        assert la.n_multiple_assignments == int(g["n_multiple_assignments"])
        assert abs(la.avg_mobile_per_site - float(g["avg_mobile_per_site"])) < 1e-12
        # jump list: exact, both modes, same order
        assert np.array_equal(st.jump_array(), g["jumps"])
        assert np.array_equal(st.jump_array(unknown_as_jump=True), g["jumps_unknown_as_jump"])
        assert [tuple(j) for j in st.jumps()] == [tuple(int(x) for x in r) for r in g["jumps"]]
        from sitator_b200.dynamics import JumpAnalysis
        JumpAnalysis().run(st)
        for key in ("n_ij", "jump_lag", "residence_times", "occupancy_freqs", "total_corrected_residences"):
            assert np.array_equal(getattr(out_sn, key), g[key]), key
        assert np.array_equal(np.nan_to_num(out_sn.p_ij, nan=-1.0), np.nan_to_num(g["p_ij"], nan=-1.0))


def test_integer_kernels_against_oracle_on_adversarial_tables():
    """Occupancy check, jump scan and JumpAnalysis on random tables with many unknowns and shared sites."""
    from oracle import landmark_oracle as orc
    from sitator_b200 import SiteNetwork, SiteTrajectory, Atoms
    from sitator_b200.dynamics import JumpAnalysis
    from sitator_b200.errors import MultipleOccupancyError
    rng = np.random.default_rng(3)
    M, C, F = 13, 9, 1500
    atoms = Atoms(rng.random((M + 4, 3)) * 5, np.eye(3) * 5)
    mob = np.zeros(M + 4, dtype=bool); mob[:M] = True
    sn = SiteNetwork(atoms, ~mob, mob)
    sn.centers = rng.random((C, 3)) * 5
    traj = rng.integers(0, C, (F, M))
    # long dwell times with occasional changes and unknown stretches (incl. unknown at frame 0)
    hold = rng.random((F, M)) < 0.9
    for f in range(1, F):
        traj[f, hold[f]] = traj[f - 1, hold[f]]
    traj[rng.random((F, M)) < 0.25] = -1
    traj[0, :4] = -1
    st = SiteTrajectory(sn, traj)
    want = orc.check_multiple_occupancy(traj, max_mobile_per_site=M)
    got = st.check_multiple_occupancy(max_mobile_per_site=M)
    assert got[0] == want[0] and abs(got[1] - want[1]) < 1e-12
    with pytest.raises(MultipleOccupancyError) as ei:
        st.check_multiple_occupancy(max_mobile_per_site=1)
    with pytest.raises(orc.MultipleOccupancyFailure) as eo:
        orc.check_multiple_occupancy(traj, max_mobile_per_site=1)
    assert (ei.value.frame, ei.value.site) == (eo.value.frame, eo.value.site)
    assert np.array_equal(ei.value.mobile_particles, eo.value.mobile)
    for uaj in (False, True):
        assert np.array_equal(st.jump_array(unknown_as_jump=uaj), orc.jumps(traj, unknown_as_jump=uaj))
    frames_seen = [f for f, *_ in st.jumps_by_frame()]
    assert frames_seen == list(range(1, F))
    JumpAnalysis().run(st)
    ja = orc.jump_analysis(traj, C)
    for key in ("n_ij", "jump_lag", "residence_times", "occupancy_freqs", "total_corrected_residences"):
        assert np.array_equal(getattr(sn, key), ja[key]), key


@pytest.mark.parametrize("n", [150, 640])        # 640: sparse enough for the sparse-row product from iteration 1
def test_mcl_kernels_against_oracle(n):
    from oracle import landmark_oracle as orc
    from sitator_b200.util.mcl import markov_clustering
    rng = np.random.default_rng(1)
    # noisy block structure with weak cross links
    blocks = np.repeat(np.arange(n // 10), 10)
    g = (blocks[:, None] == blocks[None, :]).astype(float) * rng.uniform(0.3, 1.0, (n, n))
    g += rng.uniform(0, 0.02, (n, n)) * (rng.random((n, n)) < 0.05)
    g = (g + g.T) / 2
    np.fill_diagonal(g, 1.0)
    for infl, exp in ((4, 2), (2, 2), (2.5, 3), (2, 5)):
        want = orc.markov_clustering(g, expansion=exp, inflation=infl)
        got = markov_clustering(g, expansion=exp, inflation=infl)
        assert got == want, (infl, exp)     # same tuples in the same (set) order


def test_site_center_methods_and_errors():
    from oracle import landmark_oracle as orc
    from sitator_b200.landmark import LandmarkAnalysis, ZeroLandmarkError, StaticLatticeError
    g, system, cfg, frames = U.load_golden("toy_bcc_300")
    want_lv = g["landmark_vectors"]
    pbc = orc.PBC(system.cell)
    wrapped = orc.wrap_frames(pbc, frames)
    res = orc.do_landmark_clustering_mcl(want_lv, {}, 0.01 / system.n_mobile)
    labels = res["cluster-labels"].reshape(len(frames), -1)
    confs = res["cluster-confs"].reshape(len(frames), -1)
    n_sites = len(res["cluster-size"])
    la, st = _run(system, cfg, frames, site_centers_method='real-unweighted')
    if np.array_equal(st.traj, labels):
        want = orc.site_centers_real(pbc, wrapped[:, system.mobile_idx], labels, confs, n_sites, weighted=False)
        assert np.max(np.abs(np.asarray(st.site_network.centers) - want)) < U.CENTER_ATOL
    la, st = _run(system, cfg, frames, site_centers_method='representative-landmark')
    want = orc.site_centers_representative(pbc, system.lm_centers, res["cluster-representative-lvecs"])
    # this toy's landmarks are symmetric, so centres land exactly on a cell face: 0 and L are the same point
    d = np.asarray(st.site_network.centers) - want
    d -= system.lengths * np.round(d / system.lengths)
    assert np.max(np.abs(d)) < U.CENTER_ATOL
    # the reference raises ZeroLandmarkError(mobile 5, frame 58) on this trajectory when checking (helpers.pyx:116-118)
    with pytest.raises(ZeroLandmarkError) as e:
        _run(system, cfg, frames, check_for_zero_landmarks=True)
    assert (e.value.mobile_index, e.value.frame) == (5, 58)
    bad = frames.copy()
    bad[41, system.static_idx[9]] += 1.2
    with pytest.raises(StaticLatticeError) as e:
        _run(system, cfg, bad)
    assert e.value.frame == 41 and e.value.lattice_atoms == [9]
    # one-shot semantics, wrong shapes, missing pieces (LandmarkAnalysis.py:165-172)
    with pytest.raises(ValueError, match="Cannot rerun"):
        la.run(syn.site_network_for(system), frames)
    with pytest.raises(ValueError, match="Wrong shape"):
        LandmarkAnalysis(clustering_algorithm='mcl').run(syn.site_network_for(system), frames[:, :-1])
    with pytest.raises(ZeroLandmarkError):        # the default algorithm ('dotprod') runs the same checks first
        LandmarkAnalysis().run(syn.site_network_for(system), frames)


def test_float32_and_memmap_frame_sources(tmp_path):
    """float32 trajectories are widened on the device: the result equals a run on frames.astype(float64) (up to the
    run-to-run rounding of the atomically accumulated Gram, ~1e-16); an np.memmap (float32 MD dump on disk) is a
    valid source."""
    system, cfg = syn.make_config("toy_bcc")
    frames32 = system.trajectory(200).astype(np.float32)
    la64, st64 = _run(system, cfg, frames32.astype(np.float64))
    la32, st32 = _run(system, cfg, frames32)
    assert np.array_equal(st32.traj, st64.traj)
    assert np.max(np.abs(st32.confidences - st64.confidences)) < 1e-13
    assert np.max(np.abs(np.asarray(st32.site_network.centers) - np.asarray(st64.site_network.centers))) < 1e-12
    lv32, lv64 = np.asarray(la32.landmark_vectors), np.asarray(la64.landmark_vectors)
    assert np.array_equal(lv32, lv64)                              # the landmark vectors themselves are bit-equal
    path = str(tmp_path / "traj.f32")
    frames32.tofile(path)
    mm = np.memmap(path, dtype=np.float32, mode="r", shape=frames32.shape)
    lam, stm = _run(system, cfg, mm)
    assert np.array_equal(stm.traj, st64.traj) and np.max(np.abs(stm.confidences - st64.confidences)) < 1e-13
    with pytest.raises(ValueError, match="float64"):
        _run(system, cfg, frames32.astype(np.float16))
