"""CPU, world size 2, gloo: the host-side logic of the frame-sharded run (landmark/parallel.py):
shard offsets, key reductions, best-row merge and the jump-scan carry across shard boundaries."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import landmark_oracle as orc
        from sitator_b200.landmark.parallel import default_comm
        from sitator_b200.engine import read_best_table
        comm = default_comm()
        assert comm is not None and comm.world == world
        out = {}
        # contiguous frame blocks: rank r holds n_r frames, frame0 = sum of the lower ranks'
        n_local = 7 + 3 * rank
        out["frame0"] = comm.exclusive_scan_int(n_local)
        out["sum"] = comm.allreduce_sum_scalar(n_local)
        out["min"] = comm.allreduce_min_int(100 - rank)
        # uint64 keys kept in int64 tensors (all-ones = "none")
        t = torch.tensor([-1, 5 + rank, (1 << 62) + rank], dtype=torch.int64)
        out["min_u64"] = comm.allreduce_min_u64_(t.clone()).tolist()
        out["max_u64"] = comm.allreduce_max_u64_(t.clone()).tolist()
        # best-row tables: value bits | row | lock; merge = max value, then lowest row
        vals = np.array([0.5 + 0.1 * rank, 0.9, 0.0])
        rows = np.array([10 + rank, 40 - 7 * rank, 0], dtype=np.int64)
        tab = torch.as_tensor(np.concatenate([vals.view(np.int64), rows, np.zeros(3, dtype=np.int64)]))
        v, r = read_best_table(tab, comm)
        out["best"] = (v.tolist(), r.tolist())
        # jump-scan carry: the reference's last_known across a shard boundary (SiteTrajectory.py:361-373)
        rng = np.random.default_rng(5)
        full = rng.integers(-1, 4, (n_local_total(world), 6))
        full[rng.random(full.shape) < 0.3] = -1
        f0 = out["frame0"]
        shard = torch.as_tensor(full[f0:f0 + n_local])
        for uaj in (False, True):
            carry, first = comm.jump_carry(shard, uaj)
            want = np.full(6, -1, dtype=np.int64)
            if f0 > 0:
                if uaj:
                    want = full[f0 - 1]
                else:
                    for f in range(f0):
                        k = full[f] != -1
                        want[k] = full[f][k]
            assert first == int(rank == 0)
            assert np.array_equal(carry.numpy(), want), (rank, uaj)
        # variable-length all-gather (any dtype, as bytes) and the gather of cached compressed rows that the
        # frame-sharded dotprod fit runs over: pool gaps squeezed out, rows in global order
        from sitator_b200.landmark.cluster.dotprod import gather_rows
        from sitator_b200.engine import SparseRows
        t16 = torch.arange(3 + 2 * rank, dtype=torch.int16) + 100 * rank
        out["varlen"] = comm.allgather_varlen(t16).tolist()
        r2 = np.random.default_rng(40 + rank)
        n_rows = 5 + 4 * rank
        cnt = r2.integers(0, 6, n_rows)
        cnt[1] = 0                                       # an all-zero row takes no pool entries
        gaps = r2.integers(0, 4, n_rows)
        off = np.cumsum(cnt + gaps) - cnt
        pool_k = np.full(int(off[-1] + cnt[-1] + 3), -7, dtype=np.int16)
        pool_v = np.full(len(pool_k), np.nan)
        dense = np.zeros((n_rows, 40))
        for i in range(n_rows):
            kk = np.sort(r2.choice(40, cnt[i], replace=False))
            vv = r2.random(cnt[i]) + 0.1
            pool_k[off[i]:off[i] + cnt[i]] = kk
            pool_v[off[i]:off[i] + cnt[i]] = vv
            dense[i, kk] = vv
        rows = SparseRows(torch.as_tensor((off << 8) | cnt), torch.as_tensor(pool_k), torch.as_tensor(pool_v),
                          None, len(pool_k), n_rows, 0)
        g = gather_rows(rows, comm)
        gp = g.ptr.numpy()
        got = np.zeros((g.n_rows, 40))
        for i in range(g.n_rows):
            o, c = gp[i] >> 8, gp[i] & 0xFF
            got[i, g.k.numpy()[o:o + c]] = g.v.numpy()[o:o + c]
        out["gathered"] = got
        out["dense"] = dense
        # site occupancies of a frame-sharded SiteTrajectory are those of the whole trajectory (SiteTrajectory.py:187-202)
        from sitator_b200 import SiteNetwork, SiteTrajectory, Atoms
        r3 = np.random.default_rng(9)
        whole = r3.integers(-1, 5, (n_local_total(world), 4))
        atoms = Atoms(r3.random((6, 3)) * 5, np.eye(3) * 5)
        mob = np.zeros(6, dtype=bool); mob[:4] = True
        sn = SiteNetwork(atoms, ~mob, mob)
        sn.centers = r3.random((5, 3)) * 5
        st = SiteTrajectory(sn, whole[f0:f0 + n_local])
        st._comm, st.frame0 = comm, f0
        out["occ"] = st.compute_site_occupancies()
        out["occ_want"] = np.bincount(whole[whole >= 0], minlength=5) / float(len(whole))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def n_local_total(world):
    return sum(7 + 3 * r for r in range(world))


def test_comm_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["frame0"] == 0 and res[1]["frame0"] == 7
    for r in range(world):
        assert res[r]["sum"] == 17 and res[r]["min"] == 99
        assert res[r]["min_u64"] == [-1, 5, 1 << 62]                   # both ranks hold "none" (all ones) in slot 0
        assert res[r]["max_u64"] == [-1, 6, (1 << 62) + 1]          # -1 = all ones = the largest key
        assert res[r]["best"] == ([0.6, 0.9, 0.0], [11, 33, 0])
        assert res[r]["varlen"] == [0, 1, 2, 100, 101, 102, 103, 104]
        assert np.array_equal(res[r]["gathered"], np.concatenate([res[0]["dense"], res[1]["dense"]]))
        assert np.array_equal(res[r]["occ"], res[r]["occ_want"])
