"""Two-tier fused fill + assign pass (csrc/sitb_fill_fast.cu): an FP32 first tier with a proven error bound, the exact
float64 kernel only for rows whose decisions lie inside the bound.  Stated parity: labels and cluster
counts IDENTICAL to the exact pass (which reproduces the reference, tests/test_run_gpu.py); confidences of first-tier
rows within CONF_TOL_TWO_TIER; the rows left to the exact kernel are counted by reason.
Reference path: helpers.pyx:134-212 (values), DotProdClassifier.pyx:166-189 (decisions)."""
import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from . import _util as U

pytestmark = pytest.mark.gpu

CONF_TOL_TWO_TIER = 5e-4      # absolute; bound = tau * sum |component * weight| (tau ~ 1.5e-4 at the LLZO shape), measured ~1e-5


def _run_and_engine(name, n_frames, **over):
    from sitator_b200.landmark import LandmarkAnalysis
    system, cfg = syn.make_config(name)
    frames = system.trajectory(n_frames)
    kw = U.analysis_kwargs(cfg)
    kw["max_mobile_per_site"] = max(4, kw["max_mobile_per_site"])
    kw.update(over)
    la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **kw)
    st = la.run(syn.site_network_for(system), frames)
    return system, frames, la, st


def _assign(eng, mode, thr, n_rows):
    import torch
    eng.set_assign_mode(mode)
    eng.two_tier_info(reset=True)
    labels = torch.full((n_rows,), -7, dtype=torch.int64, device=eng.device)
    confs = torch.full((n_rows,), -7.0, dtype=torch.float64, device=eng.device)
    counts = torch.zeros(eng.n_clusters, dtype=torch.int64, device=eng.device)
    eng.pass_assign(thr, labels=labels, confs=confs, counts=counts)
    torch.cuda.synchronize()
    return labels.cpu().numpy(), confs.cpu().numpy(), counts.cpu().numpy(), eng.two_tier_info(reset=True)


@pytest.mark.parametrize("name,n_frames", [("toy_bcc", 400), ("llzo", 300), ("lgps_dynamic", 120), ("laso", 12)])
def test_two_tier_labels_identical_to_exact(name, n_frames):
    system, frames, la, st = _run_and_engine(name, n_frames)
    eng = la._engine
    N = n_frames * system.n_mobile
    le, ce, ne, _ = _assign(eng, "exact", 0.7, N)
    lf, cf, nf, info = _assign(eng, "two_tier", 0.7, N)
    assert info["available"], "first tier not available for an orthorhombic cell with a candidate grid"
    assert np.array_equal(le.reshape(n_frames, -1), st.traj)
    assert np.array_equal(lf, le), "%d labels differ between the two-tier and the exact pass" % int((lf != le).sum())
    assert np.array_equal(nf, ne)
    assert float(np.max(np.abs(cf - ce))) < CONF_TOL_TWO_TIER
    assert info["recheck_rows"] < 0.05 * N + 60, info     # the first tier must decide nearly everything itself
    assert info["recheck_rows"] == sum(info["recheck_" + k] for k in ("frame", "support", "margin", "threshold", "long"))


def test_two_tier_threshold_on_a_confidence_goes_to_the_exact_kernel():
    """An assignment threshold sitting exactly on rows' confidences: those rows are inside the bound, must be
    counted as such and come out as the exact kernel decides them."""
    system, frames, la, st = _run_and_engine("llzo", 60)
    eng = la._engine
    N = 60 * system.n_mobile
    le, ce, _, _ = _assign(eng, "exact", 0.7, N)
    thr = float(np.sort(ce[ce > 0])[len(ce[ce > 0]) // 2])       # the median confidence: `conf >= thr` holds with equality there
    le2, ce2, ne2, _ = _assign(eng, "exact", thr, N)
    lf2, cf2, nf2, info = _assign(eng, "two_tier", thr, N)
    assert info["recheck_threshold"] >= 1
    assert np.array_equal(lf2, le2) and np.array_equal(nf2, ne2)
    hit = ce == thr
    assert np.all(lf2[hit] >= 0)                                   # conf >= thr: assigned (DotProdClassifier.pyx:184)


def test_two_tier_moved_static_atom_leaves_the_frame_to_the_exact_kernel():
    """Frames with a static atom beyond the candidate-grid margin are not handled by the first tier."""
    import torch
    system, frames, la, st = _run_and_engine("toy_bcc", 100)
    eng = la._engine
    moved = frames.copy()
    moved[17, system.static_idx[3]] += 0.7        # beyond the 0.5 A grid margin, inside static_movement_threshold
    moved[40, system.static_idx[9]] -= 0.3        # beyond half the margin: the looser candidate lists
    eng.set_frames(moved)
    N = 100 * system.n_mobile
    le, ce, ne, _ = _assign(eng, "exact", 0.7, N)
    lf, cf, nf, info = _assign(eng, "two_tier", 0.7, N)
    assert info["recheck_frame"] == system.n_mobile
    assert np.array_equal(lf, le) and np.array_equal(nf, ne)
    assert float(np.max(np.abs(cf - ce))) < CONF_TOL_TWO_TIER


def test_two_tier_falls_back_for_triclinic_cells():
    import torch
    from sitator_b200.engine import LandmarkEngine
    t = U.triclinic_system()
    eng = LandmarkEngine(t["cell"], t["static_idx"], t["mobile_idx"], t["n_atoms"], t["static"], t["centers"], t["verts"])
    eng.set_frames(t["frames"])
    L = len(t["centers"])
    eng.set_centers(np.arange(L, dtype=np.int32) % 7, np.ones(L), 7)
    N = len(t["frames"]) * len(t["mobile_idx"])
    le, ce, ne, _ = _assign(eng, "exact", 0.3, N)
    lf, cf, nf, info = _assign(eng, "two_tier", 0.3, N)
    assert not info["available"]
    assert np.array_equal(lf, le) and np.array_equal(cf, ce) and np.array_equal(nf, ne)


@pytest.mark.parametrize("dynamic", [False, True])
def test_two_tier_with_two_vertex_blocks_and_soft_cutoff(dynamic):
    """Landmarks with 3..8 vertices (two vertex blocks: the <2, *> kernels), a softer and wider cut-off (larger tau),
    against the float64 pass and against DotProdClassifier.predict restated on the oracle's landmark vectors."""
    import torch
    from oracle import landmark_oracle as orc
    from sitator_b200.engine import LandmarkEngine
    system, cfg = syn.make_config("toy_bcc")
    pbc = orc.PBC(system.cell)
    rng = np.random.default_rng(5)
    verts = []
    for c in system.lm_centers:
        d = pbc.distances(c, system.static_pos)
        verts.append([int(x) for x in np.argsort(d, kind="stable")[:int(rng.integers(3, 9))]])
    frames = system.trajectory(80)
    mid, steep = 1.4, 18.0
    eng = LandmarkEngine(system.cell, system.static_idx, system.mobile_idx, system.n_total, system.static_pos,
                         system.lm_centers, verts, cutoff_midpoint=mid, cutoff_steepness=steep, dynamic_lattice_mapping=dynamic)
    eng.set_frames(frames)
    d = system.lm_centers[:, None, :] - system.site_pos[None, :, :]
    d -= system.lengths * np.round(d / system.lengths)
    dist = np.sqrt((d ** 2).sum(-1))
    cid = np.where(dist.min(1) < 2.0, dist.argmin(1), -1).astype(np.int32)
    n_sites = len(system.site_pos)
    w = 0.5 + rng.random(system.n_landmarks)
    eng.set_centers(cid, w, n_sites)
    N = len(frames) * system.n_mobile
    thr = 0.9
    le, ce, ne, _ = _assign(eng, "exact", thr, N)
    lf, cf, nf, info = _assign(eng, "two_tier", thr, N)
    assert info["available"]
    assert np.array_equal(lf, le) and np.array_equal(nf, ne)
    assert float(np.max(np.abs(cf - ce))) < 20 * info["tau"] + 1e-6
    assert info["recheck_rows"] < 0.1 * N
    want, _, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                           system.lm_centers, verts, frames, midpoint=mid, steepness=steep,
                                           check_for_zeros=False, dynamic_lattice_mapping=dynamic)
    centers = np.zeros((n_sites, system.n_landmarks))
    sel = cid >= 0
    centers[cid[sel], np.nonzero(sel)[0]] = w[sel]
    dots = np.abs(want @ centers.T)
    srt = np.sort(dots, axis=1)
    decided = np.minimum(srt[:, -1] - srt[:, -2], np.abs(srt[:, -1] - thr)) > 1e-9
    want_labels = np.where((srt[:, -1] >= thr) & want.any(1), dots.argmax(1), -1)
    assert np.array_equal(lf[decided], want_labels[decided])


def test_two_tier_falls_back_for_large_centre_weights():
    """Centre weights beyond the fixed-point range of the error bound's sum: the float64 kernel runs alone."""
    system, frames, la, st = _run_and_engine("toy_bcc", 60)
    eng = la._engine
    cid, w = la.cluster_centers_
    eng.set_centers(cid, np.asarray(w) * 1000.0, st.site_network.n_sites)
    N = 60 * system.n_mobile
    le, ce, ne, _ = _assign(eng, "exact", 700.0, N)
    lf, cf, nf, info = _assign(eng, "two_tier", 700.0, N)
    assert info["recheck_rows"] == 0                       # the first tier did not run
    assert np.array_equal(lf, le) and np.array_equal(cf, ce) and np.array_equal(nf, ne)
