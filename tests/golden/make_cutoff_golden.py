#!/usr/bin/env python
"""Golden fixture for NON-DEFAULT cut-off parameters (`cutoff_midpoint`, `cutoff_steepness`, helpers.pyx:41-43,
127-131, 197-209): landmark vectors of the UNMODIFIED compiled reference (oracle/_ref).  Every other golden uses the
defaults 1.5 / 30.

Run in the build container (needs oracle/_ref):  python tests/golden/make_cutoff_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_loader                                  # noqa: E402
from tests import _util as U                                   # noqa: E402
from make_triclinic_golden import reference_landmark_vectors   # noqa: E402


def main():
    if not ref_loader.available():
        sys.exit("needs oracle/_ref (python oracle/build_ref.py in a container with /root/reference)")
    ref = ref_loader.load()
    out = {}
    for name, (system, frames, midpoint, steepness) in U.cutoff_cases().items():
        t = dict(n_atoms=system.n_total, static_idx=system.static_idx, mobile_idx=system.mobile_idx,
                 static=system.static_pos, frames=frames, cell=system.cell, centers=system.lm_centers,
                 verts=system.lm_vertices)
        lv, n_zero = reference_landmark_vectors(ref, t, cutoff_midpoint=midpoint, cutoff_steepness=steepness)
        rows, cols = np.nonzero(lv)
        out[name + "/shape"] = np.array(lv.shape)
        out[name + "/rows"] = rows.astype(np.int32)
        out[name + "/cols"] = cols.astype(np.int32)
        out[name + "/vals"] = lv[rows, cols]
        out[name + "/n_zero"] = n_zero
        print("%s: midpoint %g steepness %g, %s, %d non-zeros (min %.3g), %d all-zero rows"
              % (name, midpoint, steepness, lv.shape, len(rows), lv[rows, cols].min(), n_zero))
    np.savez_compressed(os.path.join(HERE, "cutoff_params_fill.npz"), **out)


if __name__ == "__main__":
    main()
