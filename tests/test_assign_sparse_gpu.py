"""The assign pass over cached compressed rows (sitb_sparse.cu) against a direct NumPy evaluation, on hand-made
rows that exercise both paths of the kernel: lane-per-row (at most 32 entries and 8 clusters in the row) and the
warp-per-row fallback (longer rows, rows touching many clusters), empty rows, landmarks outside every cluster,
and exact ties between rows (the "first row among equal values" rule of np.argmax, cluster/mcl.py:81-83)."""
import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from . import _util as U

pytestmark = pytest.mark.gpu


def _make_rows(rng, n_rows, L, max_len, dup_every=7):
    ks, vs = [], []
    for r in range(n_rows):
        if dup_every and r >= dup_every and r % dup_every == 0:
            src = int(rng.integers(0, r))                     # an exact copy of an earlier row: ties in every cluster
            ks.append(ks[src].copy()); vs.append(vs[src].copy())
            continue
        n = int(rng.integers(0, max_len + 1))
        if rng.random() < 0.05:
            n = 0
        k = np.sort(rng.choice(L, size=n, replace=False)).astype(np.uint16)
        v = rng.random(n) * 0.9 + 0.05
        ks.append(k); vs.append(v)
    return ks, vs


@pytest.mark.parametrize("case", ["few_big_clusters", "many_small_clusters", "long_rows"])
def test_assign_sparse_against_numpy(case):
    import torch
    from sitator_b200.engine import SparseRows, new_best_table, read_best_table
    rng = np.random.default_rng({"few_big_clusters": 1, "many_small_clusters": 2, "long_rows": 3}[case])
    system, cfg = syn.make_config("toy_bcc")
    eng = U.engine_for(system)
    L = eng.L
    if case == "few_big_clusters":
        n_c, max_len = 6, 30
    elif case == "many_small_clusters":
        n_c, max_len = 90, 32              # most rows touch more than 8 clusters: the fallback path
    else:
        n_c, max_len = 25, 100             # rows beyond 32 entries
    cid = rng.integers(-1, n_c, size=L).astype(np.int32)     # -1: in no cluster
    w = rng.standard_normal(L)
    w[cid < 0] = 0.0
    n_rows, row0 = 1500 + 13, 4242
    ks, vs = _make_rows(rng, n_rows, L, max_len)
    cnt = np.array([len(k) for k in ks], dtype=np.uint64)
    off = np.zeros(n_rows, dtype=np.uint64)
    # rows scattered through the pool in a shuffled order, like the per-warp slices of the fill pass
    order = rng.permutation(n_rows)
    pos = 5
    for r in order:
        off[r] = pos
        pos += int(cnt[r]) + int(rng.integers(0, 3))
    pool_k = np.zeros(pos + 64, dtype=np.uint16)
    pool_v = np.zeros(pos + 64, dtype=np.float64)
    for r in range(n_rows):
        pool_k[int(off[r]):int(off[r]) + len(ks[r])] = ks[r]
        pool_v[int(off[r]):int(off[r]) + len(ks[r])] = vs[r]
    ptr = (off << np.uint64(8)) | cnt
    dev = eng.device
    rows = SparseRows(torch.as_tensor(ptr.view(np.int64), device=dev), torch.as_tensor(pool_k.view(np.int16), device=dev),
                      torch.as_tensor(pool_v, device=dev), None, len(pool_k), n_rows, row0)
    eng.set_centers(cid, w, n_c)

    # NumPy: dense rows x dense centres
    dense = np.zeros((n_rows, L))
    for r in range(n_rows):
        dense[r, ks[r].astype(np.int64)] = vs[r]
    centers = np.zeros((n_c, L))
    inside = cid >= 0
    centers[cid[inside], np.nonzero(inside)[0]] = w[inside]
    dots = np.abs(dense @ centers.T)

    for thr in (0.35, float("nan")):
        labels = torch.full((n_rows,), -7, dtype=torch.int64, device=dev)
        confs = torch.full((n_rows,), -7.0, dtype=torch.float64, device=dev)
        counts = torch.zeros((n_c,), dtype=torch.int64, device=dev)
        best = new_best_table(n_c, dev)
        site_best = new_best_table(n_c, dev)
        rep = torch.zeros((n_c, L), dtype=torch.float64, device=dev)
        rep_w = torch.zeros((n_c,), dtype=torch.float64, device=dev)
        eng.assign_sparse(rows, thr, labels=labels, confs=confs, counts=counts, best=best, rep=rep, rep_w=rep_w,
                          site_best=site_best)
        labels, confs, counts = labels.cpu().numpy(), confs.cpu().numpy(), counts.cpu().numpy()
        want_l = np.argmax(dots, axis=1)
        want_c = dots[np.arange(n_rows), want_l]
        srt = np.sort(dots, axis=1)
        near = ((srt[:, -1] - srt[:, -2]) < U.TIE_TOL) | (np.abs(srt[:, -1] - thr) < U.TIE_TOL)
        off_thr = ~(want_c >= thr)
        off_thr |= cnt == 0
        want_l = np.where(off_thr, -1, want_l)
        want_c = np.where(off_thr, 0.0, want_c)
        bad = (labels != want_l) & ~near
        assert not np.any(bad), "%d labels differ outside the tie tolerance" % int(bad.sum())
        same = labels == want_l
        assert np.max(np.abs(confs[same] - want_c[same])) < U.CONF_ATOL
        assert np.array_equal(counts, np.bincount(labels[labels >= 0], minlength=n_c))
        # best row per cluster: the maximum over ALL rows, first row among (near-)equal values
        vals, brow = read_best_table(best)
        assert np.max(np.abs(vals - dots.max(axis=0))) < U.CONF_ATOL
        for c in range(n_c):
            if vals[c] == 0.0:
                continue
            cand = np.nonzero(dots[:, c] >= vals[c] - U.TIE_TOL)[0]
            assert brow[c] - row0 in cand
            # identical rows give bit-identical sums: among them the first one must have been kept
            r = int(brow[c] - row0)
            twins = [q for q in cand if len(ks[q]) == len(ks[r]) and np.array_equal(ks[q], ks[r]) and np.array_equal(vs[q], vs[r])]
            assert r == min(twins)
        svals, srow = read_best_table(site_best)
        for c in range(n_c):
            mine = np.nonzero(labels == c)[0]
            if len(mine) == 0:
                assert svals[c] == 0.0
                continue
            assert svals[c] == confs[mine].max()
            assert srow[c] - row0 == mine[confs[mine] == svals[c]].min()
        # representative landmark vectors: sum of conf * row over the rows of each site
        want_rep = np.zeros((n_c, L)); want_w = np.zeros(n_c)
        for c in range(n_c):
            mine = labels == c
            want_rep[c] = (confs[mine, None] * dense[mine]).sum(axis=0)
            want_w[c] = confs[mine].sum()
        assert np.max(np.abs(rep.cpu().numpy() - want_rep)) < 1e-9
        assert np.max(np.abs(rep_w.cpu().numpy() - want_w)) < 1e-9


def test_slot_layout_path_equals_pool_path():
    """Rows written by the slotted fill pass are read through the slot path (vector loads at 32 * row, entries beyond
    the row's count masked); the same buffers under other addresses go through the generic path (offset from the row
    pointer).  The two paths deal the entries to the lanes differently, so sums may differ in the last bits: labels, counts
    and the rows of the tables must agree, values to rounding."""
    import torch
    from sitator_b200.engine import SparseRows, new_best_table
    system, cfg = syn.make_config("llzo")
    eng = U.engine_for(system)
    frames = system.trajectory(300, seed=11)
    eng.set_frames(frames)
    eng.reset_status()
    seen, gram, rows = eng.pass_stats_cached(want_gram=False)
    L = eng.L
    rng = np.random.default_rng(5)
    n_c = 40
    cid = rng.integers(-1, n_c, size=L).astype(np.int32)
    w = rng.standard_normal(L)
    w[cid < 0] = 0.0
    eng.set_centers(cid, w, n_c)
    ptr = rows.ptr.cpu().numpy().view(np.uint64)
    cnt = (ptr & np.uint64(0xFF)).astype(np.int64)
    off = (ptr >> np.uint64(8)).astype(np.int64)
    short = cnt <= eng.ROW_SLOT
    assert np.array_equal(off[short], np.arange(rows.n_rows)[short] * eng.ROW_SLOT)       # the slot layout itself
    assert np.all(off[~short] >= rows.n_rows * eng.ROW_SLOT)
    # poison what lies beyond every short row's count: the slot path must not let it through
    k_h = rows.k.cpu().numpy().copy(); v_h = rows.v.cpu().numpy().copy()
    for r in np.nonzero(short)[0][:2000]:
        k_h[off[r] + cnt[r]: off[r] + eng.ROW_SLOT] = -1            # landmark 65535
        v_h[off[r] + cnt[r]: off[r] + eng.ROW_SLOT] = np.nan
    rows.k.copy_(torch.as_tensor(k_h, device=eng.device)); rows.v.copy_(torch.as_tensor(v_h, device=eng.device))
    clone = SparseRows(rows.ptr.clone(), rows.k.clone(), rows.v.clone(), None, rows.capacity, rows.n_rows, rows.row0)

    def run(rr):
        out = dict(labels=torch.full((rr.n_rows,), -7, dtype=torch.int64, device=eng.device),
                   confs=torch.full((rr.n_rows,), -7.0, dtype=torch.float64, device=eng.device),
                   counts=torch.zeros((n_c,), dtype=torch.int64, device=eng.device), best=new_best_table(n_c, eng.device),
                   site_best=new_best_table(n_c, eng.device), rep=torch.zeros((n_c, L), dtype=torch.float64, device=eng.device),
                   rep_w=torch.zeros((n_c,), dtype=torch.float64, device=eng.device))
        eng.assign_sparse(rr, 0.3, **out)
        return {k: v.cpu().numpy() for k, v in out.items()}

    a, b = run(rows), run(clone)
    assert (a["labels"] >= 0).sum() > 100 and (a["labels"] < 0).sum() > 0
    assert np.array_equal(a["labels"], b["labels"]) and np.array_equal(a["counts"], b["counts"])
    assert np.allclose(a["confs"], b["confs"], rtol=1e-13, atol=0.0)
    for key in ("best", "site_best"):
        va, vb = a[key][:n_c].view(np.float64), b[key][:n_c].view(np.float64)
        assert np.allclose(va, vb, rtol=1e-13, atol=0.0), key
        assert np.array_equal(a[key][n_c:2 * n_c], b[key][n_c:2 * n_c]), key
    assert np.allclose(a["rep"], b["rep"], rtol=1e-12, atol=1e-12) and np.allclose(a["rep_w"], b["rep_w"], rtol=1e-12)
