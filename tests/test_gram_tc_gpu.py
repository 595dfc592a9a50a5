"""tcgen05 tensor-core Gram (sitb_gram_syrk_tc) against FP64 references.

The tensor-core path is a floating-point kernel: operands are fp16 hi + 2^-12 lo splits of the FP64 landmark
vectors (relative operand error 2^-22, the lo.lo term is dropped) and TMEM accumulates in FP32 (truncating) over
256 rows before draining into FP64 registers.  Stated tolerance: |G_tc - G| <= 2e-6 * sqrt(G_ii * G_jj).
"""
import ctypes as C

import numpy as np
import pytest

from tests import _util as U
from sitator_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

GRAM_TC_RTOL = 2e-6


def _stage(xt_f16, lpad, ld):
    """(L, K) fp16 -> the tiled, 128B-swizzled staging layout of include/sitator_b200.h (sitb_gram_syrk_tc)."""
    import torch
    L, K = xt_f16.shape
    full = torch.zeros((lpad, ld), dtype=torch.float16, device=xt_f16.device)
    full[:L, :K] = xt_f16
    t = full.view(lpad // 128, 128, ld // 64, 8, 8).permute(0, 2, 1, 3, 4).contiguous()    # [rt][kt][r][chunk][8]
    r = torch.arange(128, device=full.device)[:, None]
    pos = torch.arange(8, device=full.device)[None, :]
    src_chunk = (pos ^ (r & 7))                                                          # stored[r][p] = src[r][p ^ (r & 7)]
    idx = src_chunk[None, None, :, :, None].expand(t.shape[0], t.shape[1], 128, 8, 8)
    return torch.gather(t, 3, idx).contiguous()


def _check(g_tc, g_ref):
    d = np.sqrt(np.clip(np.diag(g_ref), 0, None))
    scale = np.outer(d, d)
    iu = np.triu_indices(g_ref.shape[0])
    err = np.abs(g_tc - g_ref)[iu]
    assert np.all(err <= GRAM_TC_RTOL * scale[iu] + 1e-12), "max scaled err %g" % np.max(err / (scale[iu] + 1e-30))


@pytest.mark.parametrize("L,K", [(128, 64), (300, 5000), (1500, 40000)])
def test_syrk_against_fp64_matmul(L, K):
    import torch
    from sitator_b200 import _native
    lib = _native.load()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(1234 + L)
    x = torch.rand((K, L), generator=g, dtype=torch.float64)
    x = x * (torch.rand((K, L), generator=g) < 0.05)                     # sparse like landmark vectors
    x = x.to(dev)
    lpad = -(-L // 128) * 128
    ld = -(-K // 64) * 64 + 64                                           # garbage-free padding columns stay zero
    xt = x.t().contiguous()
    hi_flat = xt.to(torch.float16)
    lo_flat = ((xt - hi_flat.to(torch.float64)) * 4096.0).to(torch.float16)
    hi, lo = _stage(hi_flat, lpad, ld), _stage(lo_flat, lpad, ld)
    gram = torch.zeros((L, L), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(2):                                                   # += semantics: two calls double it
        _native.check(lib.sitb_gram_syrk_tc(0, C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()), L, lpad, ld, K,
                                            C.c_void_p(gram.data_ptr()), C.c_void_p(stream)))
    torch.cuda.synchronize()
    ref = (2.0 * (xt @ x)).cpu().numpy()
    _check(np.triu(gram.cpu().numpy()), np.triu(ref))


@pytest.mark.parametrize("name,frames", [("toy_bcc", 300), ("llzo", 60)])
def test_pass_stats_tc_matches_sparse_gram(name, frames):
    import torch
    system, cfg = syn.make_config(name)
    eng = U.engine_for(system)
    eng.set_frames(system.trajectory(frames))
    seen, gram = eng.pass_stats()
    seen_tc, gram_tc = eng.pass_stats_tc(block_frames=37)                # ragged blocks on purpose
    torch.cuda.synchronize()
    assert torch.equal(seen, seen_tc)
    _check(np.triu(gram_tc.cpu().numpy()), np.triu(gram.cpu().numpy()))


@pytest.mark.parametrize("name", ["toy_bcc_300", "llzo_60"])
def test_run_with_tensor_core_gram_reproduces_reference_sites(name):
    """The whole analysis with the Gram on tensor cores: same sites, labels and jumps as the reference golden."""
    from sitator_b200.landmark import LandmarkAnalysis
    g, system, cfg, frames = U.load_golden(name)
    sn = syn.site_network_for(system)
    la = LandmarkAnalysis(clustering_algorithm='mcl', clustering_params={'gram_method': 'tcgen05'}, verbose=False,
                          **U.analysis_kwargs(cfg))
    st = la.run(sn, frames)
    assert [set(v) for v in st.site_network.vertices] == g["site_vertex_sets"]
    assert np.array_equal(st.traj, g["labels"])
    assert np.max(np.abs(st.confidences - g["confs"])) < 1e-6
    assert np.array_equal(st.jump_array(), g["jumps"])


@pytest.mark.parametrize("name,frames", [("toy_bcc", 300), ("llzo", 100), ("lgps_dynamic", 70)])
def test_windowed_gram_from_cached_rows_matches_direct(name, frames):
    """sitb_gram_from_cached (per atom-and-window shared-memory tables) against the one-atomic-per-product Gram
    of sitb_pass_stats: the same FP64 products in another association, so equal to rounding (1e-13)."""
    import torch
    system, cfg = syn.make_config(name)
    eng = U.engine_for(system, dynamic_lattice_mapping=cfg["dynamic"])
    eng.set_frames(system.trajectory(frames))
    seen, gram = eng.pass_stats()
    for from_rows in (True, False):
        seen_c, gram_c, rows = eng.pass_stats_cached(gram_from_rows=from_rows)
        torch.cuda.synchronize()
        assert torch.equal(seen, seen_c)
        g, gc = np.triu(gram.cpu().numpy()), np.triu(gram_c.cpu().numpy())
        assert np.array_equal(g != 0, gc != 0)
        assert np.max(np.abs(g - gc) / np.maximum(np.abs(g), 1e-300)) < 1e-13
