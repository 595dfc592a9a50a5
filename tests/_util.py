"""Shared helpers for the parity tests (tests only; may import oracle/)."""
import numpy as np

from sitator_b200 import synthetic as syn

# stated tolerances ("Parity rules")
LV_RTOL = 1e-12     # landmark-vector components, relative, on the common support (measured: 6e-15)
CONF_ATOL = 1e-12   # confidences (measured: 8e-16)
CENTER_ATOL = 1e-10 # site centres, Angstrom (measured: 9e-15)
TIE_TOL = 1e-12     # top-2 similarity margin / distance to the assignment threshold below which a label may differ


def engine_for(system, **kw):
    from sitator_b200.engine import LandmarkEngine
    return LandmarkEngine(system.cell, system.static_idx, system.mobile_idx, system.n_total, system.static_pos,
                          system.lm_centers, system.lm_vertices, **kw)


def canonical_relabel(labels, landmark_clusters):
    """Relabel sites by the lexicographic order of their sorted landmark tuples."""
    keys = [tuple(sorted(int(x) for x in c)) for c in landmark_clusters]
    order = sorted(range(len(keys)), key=lambda i: keys[i])
    remap = np.full(len(keys) + 1, -1, dtype=np.int64)
    for new, old in enumerate(order):
        remap[old] = new
    out = remap[np.asarray(labels)]          # -1 indexes the trailing -1 slot
    return out, [keys[i] for i in order], order


def triclinic_system(seed=5, n_static=40, n_mobile=6, n_landmarks=90, n_frames=12):
    """A small triclinic cell with random landmarks, for the general (non-diagonal) wrap path."""
    from oracle import landmark_oracle as orc
    rng = np.random.default_rng(seed)
    cell = np.array([[9.1, 0.3, -0.4], [1.7, 8.2, 0.6], [-0.8, 1.1, 10.3]])
    pbc = orc.PBC(cell)
    static = rng.random((n_static, 3)) @ cell
    centers = rng.random((n_landmarks, 3)) @ cell
    verts = []
    for c in centers:
        d = pbc.distances(c, static)
        nv = int(rng.integers(2, 6))
        verts.append(sorted(int(x) for x in np.argsort(d, kind="stable")[:nv]))
    A = n_static + n_mobile
    order = rng.permutation(A)
    static_idx = np.sort(order[:n_static])
    mobile_idx = np.sort(order[n_static:])
    frames = np.empty((n_frames, A, 3))
    frames[:, static_idx] = static[None] + rng.normal(0, 0.05, (n_frames, n_static, 3))
    # mobiles sit near landmark centres; add whole lattice vectors so wrapping is exercised
    pick = rng.integers(0, n_landmarks, (n_frames, n_mobile))
    frames[:, mobile_idx] = centers[pick] + rng.normal(0, 0.15, (n_frames, n_mobile, 3))
    frames += (rng.integers(-2, 3, (n_frames, A, 3)).astype(float)) @ cell
    return dict(cell=cell, static=static, static_idx=static_idx, mobile_idx=mobile_idx, n_atoms=A,
                centers=centers, verts=verts, frames=frames)


GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")
GOLDEN_CASES = ["toy_bcc_300", "llzo_60", "lgps_dynamic_40", "toy_bcc_2000", "laso_16"]


DOTPROD_GOLDEN_CASES = ["toy_bcc_300_dotprod", "llzo_60_dotprod", "lgps_dynamic_40_dotprod"]


def load_dotprod_golden(name):
    """Outputs of the compiled reference with its default clustering_algorithm='dotprod' + the inputs."""
    import ast
    import os
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    system, cfg = syn.make_config(str(g["config"]))
    frames = system.trajectory(int(g["n_frames"]), **ast.literal_eval(str(g["traj_kw"])))
    return g, system, cfg, frames


def load_golden(name):
    """Fixture written by tests/golden/make_golden.py (outputs of the compiled reference) + its inputs."""
    import ast
    import os
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    config = str(g["config"])
    n_frames = int(g["n_frames"])
    traj_kw = ast.literal_eval(str(g["traj_kw"]))
    system, cfg = syn.make_config(config)
    frames = system.trajectory(n_frames, **traj_kw)
    shape = tuple(int(x) for x in g["lv_shape"])
    lv = np.zeros(shape)
    lv[g["lv_rows"].astype(np.int64), g["lv_cols"].astype(np.int64)] = g["lv_vals"]
    g["landmark_vectors"] = lv
    ends = np.cumsum(g["site_vertices_len"])
    g["site_vertex_sets"] = [set(int(x) for x in g["site_vertices"][e - n:e]) for e, n in zip(ends, g["site_vertices_len"])]
    return g, system, cfg, frames


def analysis_kwargs(cfg):
    return dict(dynamic_lattice_mapping=cfg["dynamic"],
                check_for_zero_landmarks=cfg.get("check_for_zero_landmarks", True),
                max_mobile_per_site=cfg.get("max_mobile_per_site", 1))


def compare_labels(got, want, decision_margin, what="labels"):
    """Labels must agree except where the reference's own decision margin is below TIE_TOL."""
    got = np.asarray(got).reshape(-1)
    want = np.asarray(want).reshape(-1)
    diff = got != want
    bad = diff & ~(decision_margin.reshape(-1) < TIE_TOL)
    assert not np.any(bad), "%d %s differ outside the tie tolerance (of %d differing)" % (int(bad.sum()), what, int(diff.sum()))
    return int(diff.sum())


def error_cases():
    """Seeded inputs on which the reference raises from inside its fill: name -> (system, frames, LandmarkAnalysis kwargs).
    Shared by tests/golden/make_error_golden.py (compiled reference) and the oracle / GPU tests."""
    from oracle import landmark_oracle as orc
    cases = {}
    # static atoms pushed beyond static_movement_threshold in two frames: the first frame in iteration order wins,
    # and within it the first lattice position (helpers.pyx:66-80)
    system, cfg = syn.make_config("toy_bcc")
    frames = system.trajectory(60)
    frames[20, system.static_idx[7]] += 1.5
    frames[9, system.static_idx[5]] += 1.3
    frames[9, system.static_idx[2]] -= 1.3
    cases["static_moved"] = (system, frames, dict(check_for_zero_landmarks=False))
    # dynamic lattice mapping, a static atom sitting on another one: no lattice position picks it (helpers.pyx:87-92)
    system, cfg = syn.make_config("lgps_dynamic")
    frames = system.trajectory(12)
    d = orc.PBC(system.cell).distances(system.static_pos[4], system.static_pos)
    second = int(np.argsort(d, kind="stable")[2])
    frames[7, system.static_idx[4]] = frames[7, system.static_idx[second]] + 0.01
    cases["dynamic_unassigned"] = (system, frames, dict(dynamic_lattice_mapping=True, static_movement_threshold=50.0,
                                                        max_mobile_per_site=2))
    # a mobile atom far from every landmark: all-zero landmark vector with check_for_zero_landmarks=True (helpers.pyx:116-122)
    system, cfg = syn.make_config("toy_bcc")
    cases["zero_vector"] = (system, system.trajectory(300), dict(check_for_zero_landmarks=True))
    return cases


def occupancy_error_table():
    """A random assignment table (F, M) with shared sites and unknowns, and its number of sites."""
    rng = np.random.default_rng(3)
    M, n_sites, F = 16, 24, 400
    traj = rng.integers(0, n_sites, (F, M))
    for f in range(37):                                  # no shared site before frame 37
        traj[f] = rng.permutation(n_sites)[:M]
    hold = rng.random((F, M)) < 0.9
    for f in range(38, F):
        traj[f, hold[f]] = traj[f - 1, hold[f]]
    traj[rng.random((F, M)) < 0.3] = -1
    return traj.astype(np.int64), n_sites


def cutoff_cases():
    """Non-default cut-off parameters: name -> (system, frames, cutoff_midpoint, cutoff_steepness)."""
    cases = {}
    system, cfg = syn.make_config("toy_bcc")
    cases["toy_soft"] = (system, system.trajectory(40), 1.3, 12.0)          # wide, soft cut-off: many more non-zeros
    system, cfg = syn.make_config("llzo")
    cases["llzo_sharp"] = (system, system.trajectory(8), 1.7, 45.0)         # sharp cut-off further out, ragged vertex lists
    cases["llzo_steep"] = (system, system.trajectory(8), 1.2, 120.0)        # nearly a step function
    return cases


def load_cutoff_golden(name):
    import os
    g = np.load(os.path.join(GOLDEN_DIR, "cutoff_params_fill.npz"))
    want = np.zeros(tuple(int(x) for x in g[name + "/shape"]))
    want[g[name + "/rows"], g[name + "/cols"]] = g[name + "/vals"]
    return want, int(g[name + "/n_zero"])


def clustering_param_cases():
    """Non-default 'mcl' clustering parameters: case -> (golden input name, clustering_params, minimum_site_occupancy)."""
    return {
        "toy_inflation3": ("toy_bcc_300", {"inflation": 3, "assignment_threshold": 0.8}, 0.01),
        "toy_thresholds": ("toy_bcc_300", {"assignment_threshold": 0.6, "good_site_normed_threshold": 0.9,
                                           "good_site_projected_threshold": 0.5}, 0.2),
        "llzo_inflation2": ("llzo_60", {"inflation": 2.5, "assignment_threshold": 0.75}, 0.05),
    }


def dotprod_param_cases():
    """Non-default 'dotprod' clustering parameters: case -> (golden input name, clustering_params, minimum_site_occupancy)."""
    return {
        "toy_loose": ("toy_bcc_300_dotprod", {"clustering_threshold": 0.3, "assignment_threshold": 0.6}, 0.01),
        "toy_tight": ("toy_bcc_300_dotprod", {"clustering_threshold": 0.7, "assignment_threshold": 0.9}, 0.1),
        "llzo_tight": ("llzo_60_dotprod", {"clustering_threshold": 0.6, "assignment_threshold": 0.85}, 0.05),
    }


MERGE_CASES = {
    # case: (golden input, distance_threshold, post_check_thresh_factor, markov_parameters, weighted_spatial_average)
    "toy_default": ("toy_bcc_2000", 2.0, 1.5, {}, True),
    "toy_inflation": ("toy_bcc_2000", 3.4, 3.0, {"inflation": 1.25, "expansion": 2}, False),
    "llzo_loose": ("llzo_60", 3.0, 3.0, {"inflation": 1.4}, True),
    # atoms flicker between the members of eight close site pairs every few frames: strong i <-> j flux, pairs merge
    "toy_flicker": ("toy_bcc_2000+flicker", 3.0, 2.0, {}, True),
}


def merge_case_inputs(gname):
    """Inputs of the site-merging fixtures: a golden's SiteTrajectory without the sites that never hold an atom for
    two consecutive frames (their n_ii = 0, which the reference's markov_clustering refuses, util/mcl.py:20).
    Returns (system, frames, labels (F, M), confidences, site centres, site vertex lists)."""
    flicker = gname.endswith("+flicker")
    gname = gname.split("+")[0]
    g, system, cfg, frames = load_golden(gname)
    keep = np.diag(g["n_ij"]) != 0
    remap = np.full(len(keep) + 1, -1, dtype=np.int64)
    remap[:-1][keep] = np.arange(int(keep.sum()))
    labels = remap[g["labels"].astype(np.int64)]
    confs = np.where(labels >= 0, g["confs"], 0.0)
    verts = [sorted(v) for v, k in zip(g["site_vertex_sets"], keep) if k]
    centers = g["site_centers"][keep].copy()
    if flicker:
        from oracle import landmark_oracle as orc
        pbc = orc.PBC(system.cell)
        rng = np.random.default_rng(17)
        partner = np.full(len(centers), -1, dtype=np.int64)
        for s_ in range(len(centers)):
            if partner[s_] >= 0:
                continue
            d = pbc.distances(centers[s_], centers)
            d[s_] = np.inf
            d[partner >= 0] = np.inf
            j = int(np.argmin(d))
            if d[j] < 2.9 and (partner >= 0).sum() < 16:
                partner[s_], partner[j] = j, s_
        swap = (rng.random(labels.shape) < 0.35) & (labels >= 0)
        swap &= partner[np.clip(labels, 0, None)] >= 0
        labels = np.where(swap, partner[np.clip(labels, 0, None)], labels)
    return system, frames, labels, confs, centers, verts
