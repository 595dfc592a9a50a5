"""Time the tensor-core Gram path (K1 stage + tcgen05 SYRK) against the sparse FP64 Gram of pass A.

    python scripts/time_gram_tc.py [n_frames] [block_frames]
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from sitator_b200 import _native, synthetic as syn
from sitator_b200.engine import LandmarkEngine


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    block = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    system, cfg = syn.make_config("llzo")
    eng = LandmarkEngine(system.cell, system.static_idx, system.mobile_idx, system.n_total, system.static_pos,
                         system.lm_centers, system.lm_vertices)
    base = system.trajectory(2000)
    reps = -(-n_frames // len(base))
    frames = torch.as_tensor(np.concatenate([base] * reps)[:n_frames]).cuda()
    eng.set_frames(frames)
    lib = _native.load()
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(n):
            a, b = ev(), ev()
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best, out

    t_sparse, (seen, gram) = timed(lambda: eng.pass_stats())
    t_tc, (seen_tc, gram_tc) = timed(lambda: eng.pass_stats_tc(block_frames=block))
    # SYRK alone on one staged block
    L, M = eng.L, eng.M
    lpad = -(-L // 128) * 128
    nb = min(block, n_frames)
    ld = -(-(nb * M) // 64) * 64
    hi = torch.zeros((lpad, ld), dtype=torch.float16, device="cuda")
    lo = torch.zeros((lpad, ld), dtype=torch.float16, device="cuda")
    s2 = torch.zeros((L,), dtype=torch.int64, device="cuda")
    _native.check(lib.sitb_pass_stage(eng._ctx, 0, nb, C.c_void_p(s2.data_ptr()), C.c_void_p(hi.data_ptr()),
                                      C.c_void_p(lo.data_ptr()), ld))
    g2 = torch.zeros((L, L), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    t_syrk, _ = timed(lambda: _native.check(lib.sitb_gram_syrk_tc(0, C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()),
                                                                   L, lpad, ld, nb * M, C.c_void_p(g2.data_ptr()),
                                                                   C.c_void_p(stream))), n=5)
    t_stage, _ = timed(lambda: _native.check(lib.sitb_pass_stage(eng._ctx, 0, nb, C.c_void_p(s2.data_ptr()),
                                                                  C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()), ld)))
    n_tiles = (lpad // 128) * (lpad // 128 + 1) // 2
    flops = 3 * 2.0 * 128 * 128 * n_tiles * ld                        # three fp16 MMAs per tile pair
    g, gt = np.triu(gram.cpu().numpy()), np.triu(gram_tc.cpu().numpy())
    d = np.sqrt(np.diag(g))
    sc = np.outer(d, d) + 1e-300
    print(json.dumps({
        "n_frames": n_frames, "block_frames": block, "L": L, "k_rows_per_block": nb * M,
        "pass_stats_sparse_fp64_ms": round(t_sparse, 3), "pass_stats_tc_ms": round(t_tc, 3),
        "syrk_block_ms": round(t_syrk, 4), "stage_block_ms": round(t_stage, 4),
        "syrk_tflops_fp16": round(flops / (t_syrk * 1e-3) / 1e12, 1),
        "gram_max_scaled_err": float(np.max(np.abs(g - gt) / sc)),
        "seen_equal": bool(torch.equal(seen, seen_tc)),
    }))


if __name__ == "__main__":
    main()
