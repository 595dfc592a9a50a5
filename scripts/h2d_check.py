"""Developer tool: host->device copy rate of the frame upload (plain torch copy vs the engine's chunked upload)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from sitator_b200 import synthetic as syn
from tests import _util as U
F = 100000
system, cfg = syn.make_config("llzo")
pinned = torch.empty((F, system.n_total, 3), dtype=torch.float64, pin_memory=True)
pinned.numpy()[:] = 1.0
dev = torch.empty_like(pinned, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    dev.copy_(pinned, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("plain copy: %.2f ms, %.1f GB/s" % (dt * 1e3, pinned.numel() * 8 / dt / 1e9))
del dev
eng = U.engine_for(system)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    eng.set_frames(pinned.numpy()); torch.cuda.synchronize(); 
    import ctypes
    dt = time.perf_counter() - t
    print("engine.set_frames (returns before the copy ends): %.2f ms" % (dt * 1e3))
    t = time.perf_counter(); eng.status(); torch.cuda.synchronize()
    # force completion of the copy stream: a tiny pass over the last frame waits for all chunks
    out = eng.fill_dense(begin=F - 1, n=1); torch.cuda.synchronize()
    dt2 = time.perf_counter() - t
    print("  + wait for all chunks: %.2f ms -> %.1f GB/s" % (dt2 * 1e3, pinned.numel() * 8 / (dt + dt2) / 1e9))
