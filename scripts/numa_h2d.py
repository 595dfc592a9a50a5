"""Concurrent host<->device copy bandwidth of all ranks, with and without binding each rank to its GPU's NUMA node
(developer diagnostic for the weak-scaling e2e curve).   torchrun --nproc-per-node N scripts/numa_h2d.py"""
import json, os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from sitator_b200.util import numa

def bench(tag):
    n = 461 * (1 << 20) // 8
    host = torch.empty(n, dtype=torch.float64, pin_memory=True); host.fill_(1.0)
    dev = torch.empty(n, dtype=torch.float64, device="cuda")
    back = torch.empty(n // 5, dtype=torch.float64, pin_memory=True); back.fill_(0.0)
    res = {}
    for name, fn, nbytes in (("h2d", lambda: dev.copy_(host, non_blocking=True), n * 8),
                             ("d2h", lambda: back.copy_(dev[:n // 5], non_blocking=True), n // 5 * 8)):
        for rep in range(3):
            if world > 1: dist.barrier()
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res[name] = nbytes / dt / 1e9
    out = [None] * world
    if world > 1: dist.all_gather_object(out, res)
    else: out = [res]
    if rank == 0:
        print(tag, "h2d GB/s per rank", [round(r["h2d"], 1) for r in out], "sum %.0f" % sum(r["h2d"] for r in out))
        print(tag, "d2h GB/s per rank", [round(r["d2h"], 1) for r in out], "sum %.0f" % sum(r["d2h"] for r in out))

info = numa.gpu_numa_info(local)
allinfo = [None] * world
if world > 1: dist.all_gather_object(allinfo, info)
else: allinfo = [info]
if rank == 0:
    print("gpu numa info:", json.dumps(allinfo))
    print("affinity before:", len(os.sched_getaffinity(0)), "cpus")
bench("unbound")
ok = numa.bind_to_gpu_node(local)
if rank == 0: print("bound:", ok, "affinity now", len(os.sched_getaffinity(0)))
bench("bound  ")
if world > 1: dist.destroy_process_group()
