"""Developer tool: time the Gram-from-cached-rows kernels (FP64 atomics vs deterministic integer words)."""
import sys
sys.path.insert(0, ".")
import ctypes as C
import numpy as np, torch
from sitator_b200 import synthetic as syn, _native
from tests import _util as U
F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
system, cfg = syn.make_config("llzo")
frames = np.concatenate([system.trajectory(20000, seed=s) for s in range(F // 20000)])
eng = U.engine_for(system); eng.set_frames(frames); eng.reset_status()
seen, gram, rows = eng.pass_stats_cached(want_gram=False)
lib = _native.load()
L = eng.L
for name, fn, shape, dt in (("atomic f64", lib.sitb_gram_from_cached, (L, L), torch.float64),
                            ("integer words", lib.sitb_gram_words_from_cached, (2 * (L + 1), L), torch.int64)):
    g = torch.zeros(shape, dtype=dt, device="cuda")
    ts = []
    for rep in range(4):
        g.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        _native.check(fn(eng._ctx, eng._ptr(rows.ptr), eng._ptr(rows.k), eng._ptr(rows.v), eng.n_frames, eng._ptr(g)))
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("%-14s %.3f ms (best of %s)" % (name, min(ts), [round(t, 2) for t in ts]))
