"""GPU: the constructor-default clustering ('dotprod') in a frame-sharded run (2 ranks on the one GPU, gloo).

The fit is order dependent over all landmark vectors, so every rank runs it over the gathered cached rows; predict,
site centres and the jump scan stay sharded.  Each rank's shard of the result must be the compiled reference's
golden output for those frames (labels, confidences, site centres, site count, jump list)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAME = "toy_bcc_300_dotprod"
CUT = 130                                      # uneven shards


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_shard(rank, world, port):
    """One rank: returns None when its shard equals the golden, else raises."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SITATOR_PROGRESSBAR"] = "false"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sitator_b200 import synthetic as syn
        from sitator_b200.landmark import LandmarkAnalysis
        from tests import _util as U
        g, system, cfg, frames = U.load_dotprod_golden(NAME)
        bounds = [0, CUT, len(frames)]
        lo, hi = bounds[rank], bounds[rank + 1]
        la = LandmarkAnalysis(verbose=False, **U.analysis_kwargs(cfg))
        assert la._cluster_algo == 'dotprod'
        st = la.run(syn.site_network_for(system), np.ascontiguousarray(frames[lo:hi]))
        assert st.frame0 == lo
        assert st.site_network.n_sites == len(g["site_centers"])
        assert np.array_equal(st.traj, g["labels"][lo:hi])
        assert np.max(np.abs(st.confidences - g["confs"][lo:hi])) < U.CONF_ATOL
        assert np.max(np.abs(np.asarray(st.site_network.centers) - g["site_centers"])) < U.CENTER_ATOL
        assert la.n_multiple_assignments == int(g["n_multiple_assignments"])
        jumps = g["jumps"]
        mine = jumps[(jumps[:, 0] >= lo) & (jumps[:, 0] < hi)] if len(jumps) else jumps
        assert np.array_equal(st.jump_array(), mine)
    finally:
        dist.destroy_process_group()


def _entry(rank, world, port, q):
    try:
        run_shard(rank, world, port)
        q.put((rank, "ok"))
    except BaseException as e:                 # reported to the parent, which fails the test
        import traceback
        q.put((rank, traceback.format_exc()))
        raise


def test_two_shards_default_clustering_match_reference_golden():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0] == "ok", res[0]
    assert res[1] == "ok", res[1]


if __name__ == "__main__":                     # python -m tests.test_zz_dotprod_sharded_gpu RANK PORT
    run_shard(int(sys.argv[1]), 2, int(sys.argv[2]))
    print("shard %s ok" % sys.argv[1])
