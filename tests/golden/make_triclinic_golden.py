#!/usr/bin/env python
"""Golden fixture for the GENERAL (triclinic) cell path: landmark vectors and PBCCalculator outputs of the
UNMODIFIED compiled reference (oracle/_ref) on tests/_util.triclinic_system() -- a triclinic cell, ragged vertex
lists (2-5 vertices), frames displaced by whole lattice vectors.  The three run() goldens are all orthorhombic, so
this is what pins the oracle's non-diagonal wrap (PBCCalculator.pyx:341-366, helpers.pyx:95-103).

The reference's fill is reached through its own run(): a capture plugin registered as
sitator.landmark.cluster.capture (a file put on the plugin package's path) takes a copy of the matrix run() hands to the clustering step and stops there.

Run in the build container (needs oracle/_ref):  python tests/golden/make_triclinic_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader          # noqa: E402
from tests import _util as U           # noqa: E402


class Captured(Exception):
    pass


PLUGIN = """
import sys
import numpy as np
def do_landmark_clustering(landmark_vectors, clustering_params, min_samples, verbose):
    sys._sitb_captured_lv = np.array(landmark_vectors, dtype=np.float64, copy=True)
    raise sys._sitb_captured_exc()
"""


def reference_landmark_vectors(ref, t, **kw):
    # run() imports and then RELOADS its plugin (LandmarkAnalysis.py:234-235), so it has to be a real file on the
    # plugin package's path
    import tempfile
    import sitator.landmark.cluster as cluster_pkg
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "capture.py"), "w") as f:
        f.write(PLUGIN)
    cluster_pkg.__path__.append(d)
    sys._sitb_captured_exc = Captured
    A = t["n_atoms"]
    static_mask = np.zeros(A, dtype=bool)
    static_mask[t["static_idx"]] = True
    positions = np.zeros((A, 3))
    positions[t["static_idx"]] = t["static"]
    positions[t["mobile_idx"]] = t["frames"][0][t["mobile_idx"]]
    atoms = ref.Atoms(positions=positions, cell=t["cell"], numbers=np.where(static_mask, 8, 3))
    sn = ref.SiteNetwork(atoms, static_mask, ~static_mask)
    sn.centers = t["centers"].copy()
    verts = np.empty(len(t["verts"]), dtype=object)
    for i, v in enumerate(t["verts"]):
        verts[i] = list(v)
    sn.vertices = verts
    la = ref.LandmarkAnalysis(clustering_algorithm='capture', verbose=False, force_no_memmap=True,
                              check_for_zero_landmarks=False, **kw)
    try:
        la.run(sn, t["frames"])
    except Captured:
        pass
    return sys._sitb_captured_lv, int(la.n_all_zero_lvecs)


def main():
    if not ref_loader.available():
        sys.exit("needs oracle/_ref (python oracle/build_ref.py in a container with /root/reference)")
    ref = ref_loader.load()
    t = U.triclinic_system()
    lv, n_zero = reference_landmark_vectors(ref, t)
    rows, cols = np.nonzero(lv)
    pb = ref.PBCCalculator(t["cell"])
    rng = np.random.default_rng(21)
    pts = rng.normal(0, 25, (64, 3))
    wrapped = pts.copy()
    pb.wrap_points(wrapped)
    dists = pb.distances(pts[0], pts[1:].copy())
    near = t["static"][:12] + rng.normal(0, 0.2, (12, 3)) + t["cell"][1]
    w = rng.random(12)
    np.savez_compressed(
        os.path.join(HERE, "triclinic_fill.npz"),
        lv_shape=np.array(lv.shape), lv_rows=rows.astype(np.int32), lv_cols=cols.astype(np.int32), lv_vals=lv[rows, cols],
        n_all_zero_lvecs=n_zero, points=pts, wrapped=wrapped, distances=dists,
        avg_points=near, avg_weights=w, average=np.asarray(pb.average(near.copy(), weights=w)),
    )
    print("triclinic_fill: %s landmark vectors, %d non-zeros, %d all-zero rows" % (lv.shape, len(rows), n_zero))


if __name__ == "__main__":
    main()
