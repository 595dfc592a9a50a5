"""Full-size property check of the frame-sharded path (BASELINE north star: 10^6 synthetic LLZO frames on 8 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        scripts/million_frames.py [total_frames]

Every rank analyses its contiguous block of the trajectory (NCCL collectives as in sitator_b200/landmark/parallel.py);
rank 0 then repeats the analysis of the WHOLE trajectory alone on its GPU and compares: number of sites, site
vertex sets, every label, the confidences, the site centres and the jump list (count + checksum of the set of rows;
the rows are unique and sorted by (frame, atom), so equal sets mean equal lists).
The reference itself cannot serve as the checker at this size (about 7 h of CPU); its parity is pinned at small
sizes by tests/golden.  Prints one JSON line.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis

CHUNK = 20000


def block(system, rank, n_frames, out):
    for f0 in range(0, n_frames, CHUNK):
        n = min(CHUNK, n_frames - f0)
        out[f0:f0 + n] = system.trajectory(n, seed=1000 * rank + f0 // CHUNK + system.seed)


def jump_checksum(j):
    """64-bit checksum of the set of rows of an (n, 4) int64 jump array (rows carry global frame numbers)."""
    if len(j) == 0:
        return 0
    j = j.astype(np.uint64)
    h = (j[:, 0] * np.uint64(0x9E3779B97F4A7C15)) ^ (j[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F)) ^ \
        (j[:, 2] * np.uint64(0x165667B19E3779F9)) ^ (j[:, 3] * np.uint64(0x27D4EB2F165667C5))
    return int(np.bitwise_xor.reduce(h * (j[:, 0] + np.uint64(1))) >> np.uint64(1))      # 63 bits: fits an int64 tensor


def analysis():
    return LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, max_mobile_per_site=4)


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    system, cfg = syn.make_config("llzo")
    per = total // world
    sn = syn.site_network_for(system)
    pinned = torch.empty((per, system.n_total, 3), dtype=torch.float64, pin_memory=True)
    frames = pinned.numpy()
    block(system, rank, per, frames)

    def timed_run(fr):
        if world > 1 and dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        la = analysis()
        st = la.run(sn, fr)
        torch.cuda.synchronize()
        return la, st, (time.perf_counter() - t) * 1e3

    timed_run(frames)                                   # warm-up (allocator pools, module load)
    la, st, ms = timed_run(frames)
    jumps = st.jump_array()
    stats = torch.tensor([ms, float(len(jumps))], dtype=torch.float64, device="cuda")
    csum = torch.tensor([jump_checksum(jumps)], dtype=torch.int64, device="cuda")
    labels = torch.as_tensor(st.traj, device="cuda").contiguous()
    confs = torch.as_tensor(st.confidences, device="cuda").contiguous()
    if world > 1:
        ms_max = stats[:1].clone()
        dist.all_reduce(ms_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats[1:], op=dist.ReduceOp.SUM)
        cs = [torch.zeros_like(csum) for _ in range(world)]
        dist.all_gather(cs, csum)
        all_labels = torch.empty((world,) + tuple(labels.shape), dtype=labels.dtype, device="cuda") if rank == 0 else None
        all_confs = torch.empty((world,) + tuple(confs.shape), dtype=confs.dtype, device="cuda") if rank == 0 else None
        dist.gather(labels, list(all_labels.unbind(0)) if rank == 0 else None, dst=0)
        dist.gather(confs, list(all_confs.unbind(0)) if rank == 0 else None, dst=0)
        sharded_ms = float(ms_max.item())
        checksum = 0
        for c in cs:
            checksum ^= int(c.item())
        dist.barrier()
        dist.destroy_process_group()
    else:
        all_labels, all_confs, sharded_ms, checksum = labels[None], confs[None], ms, int(csum.item())
    if rank != 0:
        return
    n_jumps = int(stats[1].item())
    sharded_labels = all_labels.reshape(-1, system.n_mobile).cpu().numpy()
    sharded_confs = all_confs.reshape(-1, system.n_mobile).cpu().numpy()
    sharded_centers = np.asarray(st.site_network.centers)
    sharded_verts = [sorted(int(x) for x in v) for v in st.site_network.vertices]
    del all_labels, all_confs, labels, confs, la, st
    torch.cuda.empty_cache()

    # the same trajectory, whole, on one GPU
    whole_pinned = torch.empty((per * world, system.n_total, 3), dtype=torch.float64, pin_memory=True)
    whole = whole_pinned.numpy()
    whole[:per] = frames
    for r in range(1, world):
        block(system, r, per, whole[r * per:(r + 1) * per])
    la1, st1, single_ms = timed_run(whole)
    j1 = st1.jump_array()
    diff = sharded_labels != st1.traj
    same = ~diff
    print(json.dumps({
        "total_frames": per * world, "n_gpus": world, "frames_per_gpu": per, "n_mobile": system.n_mobile,
        "sharded_run_ms": sharded_ms, "single_gpu_run_ms_cold": single_ms,
        "sharded_frame_atoms_per_s": per * world * system.n_total / (sharded_ms * 1e-3),
        "n_sites_sharded": len(sharded_verts), "n_sites_single": int(st1.site_network.n_sites),
        "site_vertex_sets_equal": sharded_verts == [sorted(int(x) for x in v) for v in st1.site_network.vertices],
        "labels_differ": int(diff.sum()), "labels_total": int(diff.size),
        "confs_max_abs_diff_where_equal": float(np.max(np.abs(sharded_confs[same] - st1.confidences[same]))),
        "centers_max_abs_diff": float(np.max(np.abs(sharded_centers - np.asarray(st1.site_network.centers))))
        if len(sharded_verts) == st1.site_network.n_sites else None,
        "n_jumps_sharded": n_jumps, "n_jumps_single": int(len(j1)),
        "jump_checksum_equal": checksum == jump_checksum(j1),
    }))


if __name__ == "__main__":
    main()
