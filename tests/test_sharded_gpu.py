"""GPU: a frame-sharded run (2 ranks, each holding half of the frames) must give exactly the single-rank
result: same sites, labels, confidences, site centres, occupancy statistics and jump list.

Both ranks use the one visible GPU and talk over gloo (their kernels never wait on one another; the
collectives are host-side), so this runs on a single-GPU box.  On NVLink boxes the same code path
runs over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_shard(rank, world, port, bounds, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SITATOR_PROGRESSBAR"] = "false"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sitator_b200 import synthetic as syn
        from sitator_b200.landmark import LandmarkAnalysis
        system, cfg = syn.make_config("toy_bcc")
        frames = system.trajectory(bounds[-1])
        mine = np.ascontiguousarray(frames[bounds[rank]:bounds[rank + 1]])
        la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, check_for_zero_landmarks=False)
        st = la.run(syn.site_network_for(system), mine)
        # post-processing across the shard boundary: the last-known-site scan carries its state over, the
        # smoothing window reads the neighbour's frames
        from sitator_b200.dynamics import SmoothSiteTrajectory
        smooth = SmoothSiteTrajectory(remove_unoccupied_sites=False).run(st, 3).traj
        lk = st.copy()
        lk_info = lk.assign_to_last_known_site(frame_threshold=2)
        # JumpAnalysis over the shard boundary: last known site and time-at-current-site are carried in
        from sitator_b200.dynamics import JumpAnalysis
        ja_st = JumpAnalysis().run(st.copy())
        ja = {k: np.asarray(getattr(ja_st.site_network, k))
              for k in ("n_ij", "p_ij", "jump_lag", "residence_times", "occupancy_freqs", "total_corrected_residences")}
        q.put((rank, dict(traj=st.traj, confs=st.confidences, centers=np.asarray(st.site_network.centers), ja=ja,
                          smooth=smooth, lk=lk.traj, lk_info=lk_info,
                          verts=[sorted(v) for v in st.site_network.vertices], frame0=st.frame0,
                          n_multi=la.n_multiple_assignments, avg=la.avg_mobile_per_site, nzero=la.n_all_zero_lvecs,
                          jumps=st.jump_array(), jumps_u=st.jump_array(unknown_as_jump=True))))
    finally:
        dist.destroy_process_group()


def test_two_shards_equal_one():
    from sitator_b200 import synthetic as syn
    from sitator_b200.landmark import LandmarkAnalysis
    F = 500
    bounds = [0, 230, F]                       # uneven shards
    system, cfg = syn.make_config("toy_bcc")
    frames = system.trajectory(F)
    la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, check_for_zero_landmarks=False)
    st = la.run(syn.site_network_for(system), frames)
    want_j, want_ju = st.jump_array(), st.jump_array(unknown_as_jump=True)
    from sitator_b200.dynamics import SmoothSiteTrajectory
    want_smooth = SmoothSiteTrajectory(remove_unoccupied_sites=False).run(st, 3).traj
    want_lk = st.copy()
    want_lk_info = want_lk.assign_to_last_known_site(frame_threshold=2)

    from sitator_b200.dynamics import JumpAnalysis
    ja_st = JumpAnalysis().run(st.copy())
    want_ja = {k: np.asarray(getattr(ja_st.site_network, k))
               for k in ("n_ij", "p_ij", "jump_lag", "residence_times", "occupancy_freqs", "total_corrected_residences")}

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_shard, args=(r, 2, port, bounds, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["frame0"] == 0 and res[1]["frame0"] == 230
    traj = np.concatenate([res[0]["traj"], res[1]["traj"]])
    confs = np.concatenate([res[0]["confs"], res[1]["confs"]])
    assert np.array_equal(traj, st.traj)
    assert np.max(np.abs(confs - st.confidences)) < 1e-13
    for r in range(2):
        assert res[r]["verts"] == [sorted(v) for v in st.site_network.vertices]
        assert np.max(np.abs(res[r]["centers"] - np.asarray(st.site_network.centers))) < 1e-11
        assert res[r]["n_multi"] == la.n_multiple_assignments
        assert abs(res[r]["avg"] - la.avg_mobile_per_site) < 1e-12
        assert res[r]["nzero"] == la.n_all_zero_lvecs
    assert np.array_equal(np.concatenate([res[0]["smooth"], res[1]["smooth"]]), want_smooth)
    assert np.array_equal(np.concatenate([res[0]["lk"], res[1]["lk"]]), want_lk.traj)
    for r in range(2):
        assert res[r]["lk_info"] == want_lk_info
    # every JumpAnalysis attribute of the sharded run equals the single-rank run, on every rank (JumpAnalysis.py:68-129)
    for r in range(2):
        for k, want in want_ja.items():
            assert np.array_equal(res[r]["ja"][k], want, equal_nan=True), (r, k)
    assert np.array_equal(np.concatenate([res[0]["jumps"], res[1]["jumps"]]), want_j)
    assert np.array_equal(np.concatenate([res[0]["jumps_u"], res[1]["jumps_u"]]), want_ju)
