"""The workload property behind the Gram kernel's window splitting (csrc/sitb_gram_sparse.cu), from the oracle's landmark
vectors (no GPU): how many distinct landmarks one mobile atom sees over a window of consecutive frames.  The kernel
holds 48 landmark slots per (atom, 16-frame window); an atom in transit between sites exceeds that, and before the
split such windows were added pair by pair (253 atomics per row).  This pins the numbers DESIGN.md quotes."""
import numpy as np

from sitator_b200 import synthetic as syn
from oracle import landmark_oracle as orc


def test_landmark_union_per_window_at_the_llzo_shape():
    system, cfg = syn.make_config("llzo")
    F = 128
    frames = system.trajectory(F, seed=system.seed)
    lv, n_zero, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                              system.lm_centers, system.lm_vertices, frames, check_for_zeros=False)
    M, L = system.n_mobile, system.n_landmarks
    sup = (lv != 0).reshape(F, M, L)
    nnz = sup.sum(axis=2)
    assert 18 < nnz.mean() < 28 and nnz.max() <= 64
    frac_long_rows = float((nnz > 32).mean())          # rows beyond a 32-entry slot (engine.ROW_SLOT)
    assert 0.01 < frac_long_rows < 0.15

    def over(T, cap):
        u = sup.reshape(F // T, T, M, L).any(axis=1).sum(axis=2)
        return float((u > cap).mean())

    assert 0.03 < over(16, 48) < 0.25       # the windows the kernel now splits (11 % over 512 frames)
    assert over(4, 48) < 0.04               # two levels down almost everything fits
    assert over(1, 48) == 0.0               # a single row never needs the pair-by-pair path here
    assert over(128, 64) > 0.25             # why much longer windows are not an option: the union keeps growing
