"""GPU: relaxed_lattice_checks=True (helpers.pyx:84-92) under dynamic lattice mapping with a static atom that no
lattice position picks: no error, and the landmark vectors are those of the compiled reference
(tests/golden/relaxed_dynamic_fill.npz).  Without the flag the same input raises (tests/test_fill_gpu.py)."""
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
