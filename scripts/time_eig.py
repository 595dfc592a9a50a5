"""Developer tool: time the device eigenvector kernel on the bench workload's real clusters."""
import sys, time
sys.path.insert(0, ".")
import ctypes as C
import numpy as np, torch
from sitator_b200 import synthetic as syn, _native
from sitator_b200.landmark import LandmarkAnalysis
from sitator_b200.landmark.source import LandmarkVectorSource
from sitator_b200.landmark.cluster import mcl as gm
from sitator_b200.util.mcl import markov_clustering_device
from tests import _util as U
system, cfg = syn.make_config("llzo")
frames = system.trajectory(20000)
eng = U.engine_for(system); eng.set_frames(frames); eng.reset_status()
src = LandmarkVectorSource(eng)
seen, cov, graph = gm.landmark_graph(src)
m2, nit = markov_clustering_device(graph, inflation=4)
clusters = [list(c) for c in gm._clusters_on_device(m2) if seen[c[0]] > 0]
print("clusters", len(clusters), "sizes max", max(len(c) for c in clusters), "mean %.1f" % np.mean([len(c) for c in clusters]))
L = eng.L
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w = gm.principal_vectors_device(cov, clusters, L)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    v = gm.principal_vectors(gm._covariance_blocks(cov, clusters))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("device path %.3f ms, host LAPACK path %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
lib = _native.load()
members = np.concatenate([np.asarray(c, dtype=np.int32) for c in clusters]); offsets = np.zeros(len(clusters) + 1, dtype=np.int32); offsets[1:] = np.cumsum([len(c) for c in clusters])
md = torch.as_tensor(members, device="cuda"); od = torch.as_tensor(offsets, device="cuda")
wd = torch.zeros(L, dtype=torch.float64, device="cuda"); sw = torch.zeros(len(clusters), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    _native.check(lib.sitb_principal_vectors(0, C.c_void_p(cov.data_ptr()), L, C.c_void_p(md.data_ptr()), C.c_void_p(od.data_ptr()), len(clusters), C.c_void_p(wd.data_ptr()), C.c_void_p(sw.data_ptr()), C.c_void_p(st)))
    b.record(); torch.cuda.synchronize()
    print("kernel %.3f ms, sweeps max %d mean %.1f" % (a.elapsed_time(b), int(sw.max()), float(sw.float().mean())))
err = 0
for cl, vv in zip(clusters, v):
    g = w[np.asarray(cl)]
    if np.dot(g, vv) < 0: g = -g
    err = max(err, float(np.max(np.abs(g - vv))))
print("max |device - LAPACK| = %.2e" % err)
