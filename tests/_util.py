"""Shared helpers for the parity tests (tests only; may import oracle/)."""
import numpy as np

from sitator_b200 import synthetic as syn

# stated tolerances ("Parity rules")
LV_RTOL = 2e-5      # landmark-vector components, relative, on the common support
CONF_ATOL = 1e-4    # confidences
CENTER_ATOL = 1e-4  # site centres, Angstrom
TIE_TOL = 1e-5      # top-2 similarity margin / distance to the assignment threshold below which a label may differ


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
