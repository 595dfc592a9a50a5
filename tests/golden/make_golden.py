#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference (compiled into
oracle/_ref from /root/reference by oracle/build_ref.py) on seeded synthetic inputs.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The inputs are regenerated from (config name, n_frames) by sitator_b200.synthetic, so only the
reference's outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref, ref_loader          # noqa: E402
from sitator_b200 import synthetic as syn          # noqa: E402

CASES = [
    # (fixture name, config, n_frames, trajectory kwargs)
    ("toy_bcc_300", "toy_bcc", 300, {}),
    ("llzo_60", "llzo", 60, {}),
    ("lgps_dynamic_40", "lgps_dynamic", 40, {"swap_statics_at": 13}),
    # BASELINE configs[0] at its stated length, and the configs[3] shape (1400 + 200 atoms, 3000 landmarks)
    ("toy_bcc_2000", "toy_bcc", 2000, {}),
    ("laso_16", "laso", 16, {}),
]


DOTPROD_CASES = [
    # the constructor-default clustering (landmark/cluster/dotprod.py), same inputs as the cases above
    ("toy_bcc_300_dotprod", "toy_bcc", 300, {}),
    ("llzo_60_dotprod", "llzo", 60, {}),
    ("lgps_dynamic_40_dotprod", "lgps_dynamic", 40, {"swap_statics_at": 13}),
]


def run_dotprod_case(ref, name, config, n_frames, traj_kw):
    system, cfg = syn.make_config(config)
    frames = system.trajectory(n_frames, **traj_kw)
    sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
    la = ref.LandmarkAnalysis(verbose=False, force_no_memmap=True,          # clustering_algorithm defaults to 'dotprod'
                              dynamic_lattice_mapping=cfg["dynamic"],
                              check_for_zero_landmarks=cfg.get("check_for_zero_landmarks", True),
                              max_mobile_per_site=cfg.get("max_mobile_per_site", 1))
    st = la.run(sn, frames)
    lv = np.asarray(la.landmark_vectors)
    zero_rows = ~lv.any(axis=1)
    confs = st.confidences.copy()
    confs.reshape(-1)[zero_rows] = 0.0
    out_sn = st.site_network
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        config=config, n_frames=n_frames, traj_kw=repr(traj_kw),
        labels=st.traj, confs=confs, site_centers=np.asarray(out_sn.centers),
        n_multiple_assignments=la.n_multiple_assignments, avg_mobile_per_site=la.avg_mobile_per_site,
        jumps=np.asarray(list(st.jumps()), dtype=np.int64).reshape(-1, 4),
    )
    print("%s: %d frames, %d sites, %d unassigned" % (name, n_frames, out_sn.n_sites, int(np.sum(st.traj < 0))))


def run_postprocess(ref):
    """assign_to_last_known_site / SmoothSiteTrajectory of the compiled reference on assignment streams with
    plenty of unknowns: the toy golden's labels and a synthetic stream (random walks over 12 sites, 25 % unknown)."""
    import ast
    out = {}
    g = dict(np.load(os.path.join(HERE, "toy_bcc_300.npz"), allow_pickle=False))
    system, cfg = syn.make_config("toy_bcc")
    rng = np.random.default_rng(7)
    F, M, C = 700, 16, 12
    steps = rng.random((F, M)) < 0.05
    synth = (np.cumsum(steps, axis=0) + rng.integers(0, C, M)[None, :]) % C
    flick = rng.random((F, M)) < 0.04                           # one-frame excursions for the smoother to remove
    synth = np.where(flick, (synth + 1) % C, synth)
    holes = rng.random((F, M)) < 0.25
    holes[200:260, 3] = True                                    # a long unknown stretch
    holes[:5, 7] = True                                         # unknown from the start
    synth = np.where(holes, -1, synth).astype(np.int64)
    streams = {"toy": (g["labels"].astype(np.int64), int(g["site_centers"].shape[0])), "synth": (synth, C)}
    for name, (traj, n_sites) in streams.items():
        sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
        sn.centers = np.zeros((n_sites, 3))
        out[name + "_traj"] = traj
        out[name + "_n_sites"] = n_sites
        for thr in (1, 2, 5):
            st = ref.SiteTrajectory(sn, traj.copy())
            info = st.assign_to_last_known_site(frame_threshold=thr)
            out["%s_lk%d" % (name, thr)] = st.traj.copy()
            out["%s_lk%d_info" % (name, thr)] = np.array([info['max_time_unknown'], info['avg_time_unknown'], info['total_reassigned']], dtype=np.float64)
        for thr, flag in ((3, True), (4, False), (10, True)):
            st = ref.SiteTrajectory(sn, traj.copy())
            sm = ref.SmoothSiteTrajectory(set_unassigned_under_threshold=flag).run(st, thr)
            out["%s_smooth%d_%d" % (name, thr, int(flag))] = sm.traj.copy()
            out["%s_smooth%d_%d_n_sites" % (name, thr, int(flag))] = sm.site_network.n_sites
    np.savez_compressed(os.path.join(HERE, "postprocess.npz"), **out)
    print("postprocess: %d arrays" % len(out))


def recenter_inputs():
    """Drifting toy trajectory + masses for the RecenterTrajectory fixture (regenerated by the tests)."""
    system, cfg = syn.make_config("toy_bcc")
    rng = np.random.default_rng(21)
    pos = system.trajectory(25) + np.cumsum(rng.normal(0, 0.05, (25, 1, 3)), axis=0)
    vel = rng.normal(0, 1.0, pos.shape)
    masses = np.where(system.static_mask, 15.999, 6.94) * (1.0 + 0.01 * rng.random(system.n_total))
    return system, pos, vel, masses


def run_recenter(ref):
    from sitator.util.RecenterTrajectory import RecenterTrajectory
    system, pos, vel, masses = recenter_inputs()
    structure = ref.Atoms(positions=system.initial_structure_positions(), cell=system.cell, numbers=np.where(system.static_mask, 8, 3))
    p, v = pos.copy(), vel.copy()
    RecenterTrajectory().run(structure, system.static_mask.copy(), p, velocities=v, masses=masses)
    np.savez_compressed(os.path.join(HERE, "recenter.npz"), positions=p, velocities=v)
    print("recenter: max shift %.3f" % np.max(np.abs(p - pos)))


def run_case(ref, name, config, n_frames, traj_kw):
    system, cfg = syn.make_config(config)
    frames = system.trajectory(n_frames, **traj_kw)
    sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
    la = ref.LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, force_no_memmap=True,
                              dynamic_lattice_mapping=cfg["dynamic"],
                              check_for_zero_landmarks=cfg.get("check_for_zero_landmarks", True),
                              max_mobile_per_site=cfg.get("max_mobile_per_site", 1))
    st = la.run(sn, frames)
    lv = np.asarray(la.landmark_vectors)
    nz = np.nonzero(lv)
    jumps = np.asarray(list(st.jumps()), dtype=np.int64).reshape(-1, 4)
    jumps_u = np.asarray(list(st.jumps(unknown_as_jump=True)), dtype=np.int64).reshape(-1, 4)
    ref.JumpAnalysis().run(st)
    out_sn = st.site_network
    zero_rows = ~lv.any(axis=1)
    confs = st.confidences.copy()
    confs.reshape(-1)[zero_rows] = 0.0          # uninitialised in the reference (DotProdClassifier.pyx:168-172)
    vert_len = np.array([len(v) for v in out_sn.vertices])
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        config=config, n_frames=n_frames, traj_kw=repr(traj_kw),
        lv_rows=nz[0].astype(np.int32), lv_cols=nz[1].astype(np.int16), lv_vals=lv[nz], lv_shape=np.array(lv.shape),
        n_all_zero_lvecs=la.n_all_zero_lvecs,
        labels=st.traj, confs=confs, site_centers=np.asarray(out_sn.centers),
        site_vertices=np.concatenate([sorted(v) for v in out_sn.vertices]), site_vertices_len=vert_len,
        n_multiple_assignments=la.n_multiple_assignments, avg_mobile_per_site=la.avg_mobile_per_site,
        jumps=jumps, jumps_unknown_as_jump=jumps_u,
        n_ij=out_sn.n_ij, p_ij=out_sn.p_ij, jump_lag=out_sn.jump_lag, residence_times=out_sn.residence_times,
        occupancy_freqs=out_sn.occupancy_freqs, total_corrected_residences=out_sn.total_corrected_residences,
    )
    print("%s: %d frames, %d sites, %d jumps, %d non-zero lvec components, %d unassigned"
          % (name, n_frames, out_sn.n_sites, len(jumps), len(nz[0]), int(np.sum(st.traj < 0))))


if __name__ == "__main__":
    if not build_ref.build(verbose=False):
        sys.exit("needs /root/reference to build oracle/_ref")
    ref = ref_loader.load()
    which = sys.argv[1:] or ["mcl", "dotprod", "post"]
    only = [w[5:] for w in which if w.startswith("only:")]       # e.g. only:laso_16 (leaves the other fixtures untouched)
    if only:
        for case in CASES:
            if case[0] in only:
                run_case(ref, *case)
        sys.exit(0)
    if "mcl" in which:
        for case in CASES:
            run_case(ref, *case)
    if "post" in which:
        run_postprocess(ref)
        run_recenter(ref)
    if "dotprod" in which:
        for case in DOTPROD_CASES:
            run_dotprod_case(ref, *case)
