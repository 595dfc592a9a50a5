#!/usr/bin/env python
"""Golden fixture for `relaxed_lattice_checks=True` (helpers.pyx:84-92): the UNMODIFIED compiled reference's landmark
vectors on the 'dynamic_unassigned' error case of tests/_util.error_cases() -- dynamic lattice mapping with a static
atom sitting on another one, so that one lattice position maps to its neighbour's atom and one atom is picked by
nobody.  Without the flag the reference raises; with it the run goes on with that lattice map.

Run in the build container (needs oracle/_ref):  python tests/golden/make_relaxed_golden.py
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
    system, frames, kw = U.error_cases()["dynamic_unassigned"]
    t = dict(n_atoms=system.n_total, static_idx=system.static_idx, mobile_idx=system.mobile_idx,
             static=system.static_pos, frames=frames, cell=system.cell, centers=system.lm_centers,
             verts=system.lm_vertices)
    lv, n_zero = reference_landmark_vectors(ref, t, dynamic_lattice_mapping=True, relaxed_lattice_checks=True,
                                            static_movement_threshold=kw["static_movement_threshold"])
    rows, cols = np.nonzero(lv)
    np.savez_compressed(os.path.join(HERE, "relaxed_dynamic_fill.npz"), lv_shape=np.array(lv.shape),
                        lv_rows=rows.astype(np.int32), lv_cols=cols.astype(np.int32), lv_vals=lv[rows, cols],
                        n_all_zero_lvecs=n_zero)
    print("relaxed_dynamic_fill: %s landmark vectors, %d non-zeros, %d all-zero rows" % (lv.shape, len(rows), n_zero))


if __name__ == "__main__":
    main()
