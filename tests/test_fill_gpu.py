"""GPU parity: K1 (wrap + lattice check + landmark fill + assign) against the NumPy oracle."""
import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from tests import _util as U

pytestmark = pytest.mark.gpu


def _oracle_lv(system, frames, **kw):
    from oracle import landmark_oracle as orc
    cnt = {}
    lv, nzero, wrapped = orc.fill_landmark_vectors(
        system.cell, system.static_pos, system.static_idx, system.mobile_idx, system.lm_centers,
        system.lm_vertices, frames, check_for_zeros=False, counters=cnt, **kw)
    return lv, nzero, cnt


def _compare_lv(got, want, rtol=None):
    rtol = U.LV_RTOL if rtol is None else rtol
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == want.shape
    assert np.array_equal(got != 0, want != 0), "support differs in %d components" % int(np.sum((got != 0) != (want != 0)))
    nz = want != 0
    rel = np.abs(got[nz] - want[nz]) / want[nz]
    assert rel.max() < rtol, "max rel err %.3g" % rel.max()
    return float(rel.max())


@pytest.mark.parametrize("name,n_frames", [("toy_bcc", 300), ("llzo", 40), ("llzo_v4", 40)])
def test_tables_and_dense_fill_match_oracle(name, n_frames):
    import torch
    from oracle import landmark_oracle as orc
    system, cfg = syn.make_config(name)
    frames = system.trajectory(n_frames)
    eng = U.engine_for(system)
    # Step 1 tables: bit-exact (LandmarkAnalysis.py:194-202)
    pbc = orc.PBC(system.cell)
    verts_np, svd = orc.vertex_tables(pbc, system.lm_centers, system.lm_vertices, system.static_pos)
    svd_gpu, q_gpu = eng.tables()
    assert np.array_equal(np.isnan(svd), np.isnan(svd_gpu))
    assert np.array_equal(svd[~np.isnan(svd)], svd_gpu[~np.isnan(svd)])
    # the squared cut-off is the exact boundary of  sqrt(q)/svd > cutoff
    thr = orc.cutoff_round_to_zero_point(1.5, 30.0)
    ok = ~np.isnan(svd)
    q = q_gpu[ok]
    assert np.all(np.sqrt(q) / svd[ok] <= thr)
    assert np.all(np.sqrt(np.nextafter(q, np.inf)) / svd[ok] > thr)

    want, nzero, cnt = _oracle_lv(system, frames)
    eng.set_frames(frames)
    eng.reset_status()
    got32 = eng.fill_dense(dtype=torch.float32).cpu().numpy()
    got64 = eng.fill_dense(dtype=torch.float64).cpu().numpy()
    st = eng.status()
    assert st.error_code == 0
    assert st.n_list_overflow == 0
    assert st.n_zero_rows == 2 * nzero          # two passes
    assert st.nnz == 2 * cnt["nnz"]
    _compare_lv(got32, want, rtol=1e-7)      # float32 output: one rounding of the float64 value
    _compare_lv(got64, want)
    # host-buffer drop-in entry
    got_host, st2 = eng.fill_landmark_vectors_host(frames)
    assert np.array_equal(got_host, got64)
    assert st2.n_zero_rows == nzero
    # selected frames
    sel = [n_frames - 1, 0, n_frames // 2]
    rows = eng.fill_frames(sel).cpu().numpy()
    M = system.n_mobile
    for i, f in enumerate(sel):
        assert np.array_equal(rows[i * M:(i + 1) * M], got32[f * M:(f + 1) * M])
    # cached compressed rows == the dense rows, and an assign pass over them == the fused pass
    seen, gram, sp = eng.pass_stats_cached()
    ptr = sp.ptr.cpu().numpy().view(np.uint64)
    k = sp.k.cpu().numpy().view(np.uint16)
    v = sp.v.cpu().numpy()
    rebuilt = np.zeros_like(got64)
    for r in range(0, len(ptr), max(1, len(ptr) // 500)):
        off, n = int(ptr[r] >> np.uint64(8)), int(ptr[r] & np.uint64(0xFF))
        rebuilt[r, k[off:off + n]] = v[off:off + n]
        assert np.array_equal(rebuilt[r], got64[r])
    # every row's entry count is its number of non-zero components (all-zero rows: count 0, no pool space)
    assert np.array_equal((ptr & np.uint64(0xFF)).astype(np.int64), np.count_nonzero(want, axis=1))
    assert np.array_equal(seen.cpu().numpy(), np.count_nonzero(want, axis=0))
    G = gram.cpu().numpy()
    Gw = np.triu(want.T @ want)
    assert np.max(np.abs(G - Gw)) < 1e-9 * max(1.0, np.abs(Gw).max())


def test_triclinic_cell_general_wrap_path():
    import torch
    from oracle import landmark_oracle as orc
    from sitator_b200.engine import LandmarkEngine
    t = U.triclinic_system()
    eng = LandmarkEngine(t["cell"], t["static_idx"], t["mobile_idx"], t["n_atoms"], t["static"], t["centers"], t["verts"])
    want, nzero, _ = orc.fill_landmark_vectors(t["cell"], t["static"], t["static_idx"], t["mobile_idx"], t["centers"],
                                               t["verts"], t["frames"], check_for_zeros=False)
    eng.set_frames(t["frames"])
    got = eng.fill_dense(dtype=torch.float64).cpu().numpy()
    assert np.count_nonzero(want) > 50
    _compare_lv(got, want)


def test_dynamic_lattice_mapping_with_swapped_statics():
    import torch
    system, cfg = syn.make_config("lgps_dynamic")
    frames = system.trajectory(30, swap_statics_at=11)
    want, nzero, cnt = _oracle_lv(system, frames, dynamic_lattice_mapping=True)
    eng = U.engine_for(system, dynamic_lattice_mapping=True)
    eng.set_frames(frames)
    eng.reset_status()
    got = eng.fill_dense(dtype=torch.float64).cpu().numpy()
    st = eng.status()
    assert st.error_code == 0
    assert st.n_duplicate_nearest == cnt["n_duplicate_nearest"]
    _compare_lv(got, want)
    # without the dynamic map the swapped pair breaks the 1 A movement limit at frame 11 (helpers.pyx:76-80)
    eng2 = U.engine_for(system, dynamic_lattice_mapping=False)
    eng2.set_frames(frames)
    eng2.reset_status()
    eng2.fill_dense()
    st2 = eng2.status()
    assert st2.error_code == 1 and st2.frame == 11


def test_errors_first_in_reference_order():
    import torch
    system, cfg = syn.make_config("toy_bcc")
    frames = system.trajectory(60)
    # push static atom 7 away in frames 20 and 9; mobile 3 far from everything? (zero vectors occur naturally)
    frames[20, system.static_idx[7]] += 1.5
    frames[9, system.static_idx[5]] += 1.3
    frames[9, system.static_idx[2]] -= 1.3
    eng = U.engine_for(system)
    eng.set_frames(frames, frame0=1000)
    eng.reset_status()
    eng.fill_dense()
    st = eng.status()
    assert (st.error_code, st.frame, st.index) == (1, 1009, 2)
    assert st.first_error(check_for_zeros=False) == (1, 1009, 2)


def test_assign_matches_oracle_predict():
    import torch
    from oracle import landmark_oracle as orc
    system, cfg = syn.make_config("toy_bcc")
    frames = system.trajectory(300)
    res = orc.run_landmark_analysis(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                    system.lm_centers, system.lm_vertices, frames, check_for_zero_landmarks=False)
    centers = res["_centers"]                      # (C, L) rows with disjoint supports
    C, L = centers.shape
    cid = np.full(L, -1, dtype=np.int32)
    w = np.zeros(L, dtype=np.float64)
    for c in range(C):
        nz = np.nonzero(centers[c])[0]
        assert np.all(cid[nz] == -1)
        cid[nz] = c
        w[nz] = centers[c, nz]
    eng = U.engine_for(system)
    eng.set_frames(frames)
    eng.set_centers(cid, w, C)
    N = frames.shape[0] * system.n_mobile
    labels = torch.empty(N, dtype=torch.int64, device="cuda")
    confs = torch.empty(N, dtype=torch.float64, device="cuda")
    counts = torch.zeros(C, dtype=torch.int64, device="cuda")
    eng.pass_assign(0.7, labels=labels, confs=confs, counts=counts)
    _, _, sp = eng.pass_stats_cached()
    labels2 = torch.empty_like(labels); confs2 = torch.empty_like(confs); counts2 = torch.zeros_like(counts)
    eng.assign_sparse(sp, 0.7, labels=labels2, confs=confs2, counts=counts2)
    assert torch.equal(labels, labels2) and torch.equal(counts, counts2)
    assert float((confs - confs2).abs().max()) < 1e-14
    labels, confs, counts = labels.cpu().numpy(), confs.cpu().numpy(), counts.cpu().numpy()
    want_l, want_c = res["cluster-labels"], res["cluster-confs"]
    diff = labels != want_l
    # a label may differ only where the decision is within TIE_TOL (reported, not hidden)
    lv = res["landmark_vectors"]
    dots = np.abs(lv @ centers.T)
    srt = np.sort(dots, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    near = (margin < U.TIE_TOL) | (np.abs(srt[:, -1] - 0.7) < U.TIE_TOL)
    assert not np.any(diff & ~near), "%d labels differ outside the tie tolerance" % int(np.sum(diff & ~near))
    same = ~diff
    assert np.max(np.abs(confs[same] - want_c[same])) < U.CONF_ATOL
    assert np.array_equal(counts, np.bincount(labels[labels >= 0], minlength=C))


@pytest.mark.parametrize("name,n_frames,dynamic", [("llzo", 200, False), ("lgps_dynamic", 60, True), ("toy_bcc", 200, False)])
def test_candidate_grid_does_not_change_results(name, n_frames, dynamic):
    """The grid of candidate lists (sitb_tables.cu: k_grid_lists) only prunes the landmark walk: rows, their
    entry order and every statistic are identical with the grid, without it, and with a margin so small that
    most frames fall back to the full walk."""
    import torch
    system, cfg = syn.make_config(name)
    frames = system.trajectory(n_frames)
    results = []
    for margin in (0.5, 0.0, 0.08, 0.22):
        eng = U.engine_for(system, dynamic_lattice_mapping=dynamic, candidate_grid_margin=margin)
        info = eng.candidate_grid_info()
        assert (info["entries"] > 0) == (margin > 0)
        eng.set_frames(frames)
        eng.reset_status()
        seen, gram, sp = eng.pass_stats_cached()
        st = eng.status()
        assert st.error_code == 0 and st.n_list_overflow == 0
        ptr = sp.ptr.cpu().numpy().view(np.uint64)
        off = (ptr >> np.uint64(8)).astype(np.int64)
        cnt = (ptr & np.uint64(0xFF)).astype(np.int64)
        k = sp.k.cpu().numpy().view(np.uint16)
        v = sp.v.cpu().numpy()
        rows = [(k[o:o + n].copy(), v[o:o + n].copy()) for o, n in zip(off, cnt)]
        results.append((st, seen.cpu().numpy(), rows))
    st_grid, st_none, st_tight = results[0][0], results[1][0], results[2][0]
    assert st_grid.n_full_walk_frames == 0
    assert st_none.n_full_walk_frames == n_frames
    assert 0 < st_tight.n_full_walk_frames <= n_frames        # sigma_static = 0.05 A: most frames exceed 0.08 A
    # lists are kept for the margin and for half of it: at 0.22 A frames spread over both levels (and the full walk)
    assert results[3][0].n_loose_grid_frames > 0 or name == "toy_bcc"
    for st, seen, rows in results[1:]:
        assert st.nnz == st_grid.nnz and st.n_zero_rows == st_grid.n_zero_rows
        assert np.array_equal(seen, results[0][1])
        for (k0, v0), (k1, v1) in zip(results[0][2], rows):
            assert np.array_equal(k0, k1) and np.array_equal(v0, v1)


@pytest.mark.parametrize("max_verts", [4, 7])
def test_dense_lattice_with_far_reaching_landmarks(max_verts):
    """A dense lattice with far-reaching landmarks: dozens of static atoms inside every cut-off radius, long
    candidate lists per grid box (the opposite regime of the LLZO-shaped cases)."""
    import torch
    from oracle import landmark_oracle as orc
    from sitator_b200.engine import LandmarkEngine
    rng = np.random.default_rng(11)
    cell = np.diag([9.0, 10.0, 11.0])
    n_static, n_mobile, n_landmarks, n_frames = 260, 5, 120, 6
    static = rng.random((n_static, 3)) * np.diag(cell)
    centers = rng.random((n_landmarks, 3)) * np.diag(cell)
    pbc = orc.PBC(cell)
    verts = []
    for c in centers:
        d = pbc.distances(c, static)
        order = np.argsort(d, kind="stable")
        nv = int(rng.integers(2, max_verts + 1))                   # 7: two blocks of four vertices per landmark
        verts.append(sorted(int(x) for x in order[8:8 + nv]))      # far vertices: cut-off radius ~ 4-5 A
    A = n_static + n_mobile
    static_idx = np.arange(n_static)
    mobile_idx = np.arange(n_static, A)
    frames = np.empty((n_frames, A, 3))
    frames[:, static_idx] = static[None] + rng.normal(0, 0.03, (n_frames, n_static, 3))
    pick = rng.integers(0, n_landmarks, (n_frames, n_mobile))
    frames[:, mobile_idx] = centers[pick] + rng.normal(0, 0.2, (n_frames, n_mobile, 3))
    want, nzero, _ = orc.fill_landmark_vectors(cell, static, static_idx, mobile_idx, centers, verts, frames,
                                               check_for_zeros=False)
    # the premise: some mobile atom has more than 48 static atoms inside the loosest cut-off radius
    svd_max = max(np.max(pbc.distances(c, static[v])) for c, v in zip(centers, verts))
    reach = 1.8070080122325283 * svd_max
    n_near = max(int(np.sum(pbc.distances(m, frames[f, static_idx]) < 0.8 * reach))
                 for f in range(n_frames) for m in frames[f, mobile_idx])
    assert n_near > 48, n_near
    eng = LandmarkEngine(cell, static_idx, mobile_idx, A, static, centers, verts)
    eng.set_frames(frames)
    got = eng.fill_dense(dtype=torch.float64).cpu().numpy()
    assert np.count_nonzero(want) > (100 if max_verts == 4 else 20)
    _compare_lv(got, want)
    # and without the candidate grid (every landmark walked)
    eng2 = LandmarkEngine(cell, static_idx, mobile_idx, A, static, centers, verts, candidate_grid_margin=0.0)
    eng2.set_frames(frames)
    assert np.array_equal(eng2.fill_dense(dtype=torch.float64).cpu().numpy(), got)


def test_unassigned_static_atoms_are_reported():
    """Dynamic lattice mapping with a static atom displaced onto another one: no lattice position picks it, and the
    reference raises StaticLatticeError(lattice_atoms = the atoms no lattice position picked, frame) (helpers.pyx:87-92)."""
    from oracle import landmark_oracle as orc
    from sitator_b200.landmark import LandmarkAnalysis, StaticLatticeError
    system, cfg = syn.make_config("lgps_dynamic")
    frames = system.trajectory(12)
    pbc = orc.PBC(system.cell)
    d = pbc.distances(system.static_pos[4], system.static_pos)
    second = int(np.argsort(d, kind="stable")[2])                      # [0] is position 4 itself, [1] its nearest neighbour
    # in frame 7 atom 4 sits on the second-nearest neighbour of its lattice position: position 4 then picks its nearest
    # neighbour's atom, and nobody picks atom 4
    frames[7, system.static_idx[4]] = frames[7, system.static_idx[second]] + 0.01
    with pytest.raises(orc.StaticLatticeFailure) as o:
        orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx, system.lm_centers,
                                  system.lm_vertices, frames, dynamic_lattice_mapping=True, static_movement_threshold=50.0)
    assert o.value.frame == 7 and o.value.kind == "unassigned"
    la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, dynamic_lattice_mapping=True, max_mobile_per_site=2,
                          static_movement_threshold=50.0)
    with pytest.raises(StaticLatticeError) as e:
        la.run(syn.site_network_for(system), frames)
    assert e.value.frame == 7
    assert list(e.value.lattice_atoms) == list(o.value.lattice_atoms) == [4]


def test_landmarks_with_up_to_twelve_vertices():
    """More than 8 vertices per landmark (the reference has no limit, helpers.pyx:188-209): vertex blocks 3 of 4."""
    import torch
    from oracle import landmark_oracle as orc
    system, cfg = syn.make_config("toy_bcc")
    pbc = orc.PBC(system.cell)
    rng = np.random.default_rng(3)
    verts = []
    for c in system.lm_centers:
        d = pbc.distances(c, system.static_pos)
        verts.append([int(x) for x in np.argsort(d, kind="stable")[:int(rng.integers(3, 13))]])
    frames = system.trajectory(30)
    from sitator_b200.engine import LandmarkEngine
    eng = LandmarkEngine(system.cell, system.static_idx, system.mobile_idx, system.n_total, system.static_pos,
                         system.lm_centers, verts, cutoff_midpoint=1.6, cutoff_steepness=20.0)
    eng.set_frames(frames)
    got = eng.fill_dense(dtype=torch.float64).cpu().numpy()
    want, _, _ = orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                                           system.lm_centers, verts, frames, midpoint=1.6, steepness=20.0,
                                           check_for_zeros=False)
    assert np.array_equal(got != 0, want != 0)
    nz = want != 0
    assert nz.sum() > 100
    assert np.max(np.abs(got[nz] - want[nz]) / want[nz]) < U.LV_RTOL
