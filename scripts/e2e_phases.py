"""Per-phase, per-rank timeline of LandmarkAnalysis.run (mcl) on the bench workload (developer tool).
    python scripts/e2e_phases.py [frames per GPU] [out.json]            (1 GPU)
    torchrun --nproc-per-node N scripts/e2e_phases.py ...                (N GPUs, weak scaling)
Runs the analysis warm (3 untimed runs), then once unsynchronised (true wall-clock) and once with SITB_PHASE_SYNC=1
(each phase's full cost), and writes both timelines for every rank."""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
out_path = sys.argv[2] if len(sys.argv) > 2 else None
system, cfg = syn.make_config("llzo")
pinned = torch.empty((F, system.n_total, 3), dtype=torch.float64, pin_memory=True)
frames = pinned.numpy()
for f0 in range(0, F, 20000):
    n = min(20000, F - f0)
    frames[f0:f0 + n] = system.trajectory(n, seed=1000 * rank + f0 // 20000 + system.seed)
kw = dict(max_mobile_per_site=4, check_for_zero_landmarks=True)


def one(sync):
    os.environ["SITB_PHASE_SYNC"] = "1" if sync else "0"
    la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **kw)
    sn = syn.site_network_for(system)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    la.run(sn, frames)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, la.stats


for _ in range(3):
    one(False)
res = {"rank": rank, "world": world, "frames_per_gpu": F}
ms, stats = [], None
for _ in range(3):
    t, stats = one(False)
    ms.append(t)
res["wall_ms_unsynchronised"] = ms
res["phases_unsynchronised"] = stats["phases_ms"]
t, stats = one(True)
res["wall_ms_synchronised"] = t
res["phases_synchronised"] = stats["phases_ms"]
allres = [res]
if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    allres = gathered
if rank == 0:
    txt = json.dumps({"n_gpus": world, "frames_per_gpu": F, "ranks": allres}, indent=1)
    if out_path:
        open(out_path, "w").write(txt)
    r0 = allres[0]
    print("wall (unsync) ms per rank:", [round(float(np.mean(r["wall_ms_unsynchronised"])), 1) for r in allres])
    print("phase: synchronised ms, max over ranks / rank 0 unsynchronised")
    for k in r0["phases_synchronised"]:
        print("  %-55s %8.2f %8.2f" % (k, max(r["phases_synchronised"].get(k, 0) for r in allres), r0["phases_unsynchronised"].get(k, 0)))
if world > 1:
    dist.destroy_process_group()
