"""Developer tool: wall-clock breakdown of LandmarkAnalysis.run phases (synchronising after each)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from sitator_b200 import synthetic as syn
from sitator_b200.engine import LandmarkEngine, new_best_table, read_best_table
from sitator_b200.landmark.source import LandmarkVectorSource
from sitator_b200.landmark.cluster import mcl as gm
from sitator_b200.util.mcl import markov_clustering_device, clusters_from_matrix

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
system, cfg = syn.make_config("llzo")
pinned = torch.empty((F, system.n_total, 3), dtype=torch.float64, pin_memory=True)
frames = pinned.numpy()
for f0 in range(0, F, 20000):
    n = min(20000, F - f0); frames[f0:f0 + n] = system.trajectory(n, seed=f0 // 20000 + 1)
sn = syn.site_network_for(system)
T = {}
G = {}
ev_last = [None]
def tick(name, t0):
    e = torch.cuda.Event(enable_timing=True); e.record()
    torch.cuda.synchronize(); T[name] = T.get(name, 0) + (time.perf_counter() - t0) * 1e3
    G[name] = G.get(name, 0) + ev_last[0].elapsed_time(e)
    e2 = torch.cuda.Event(enable_timing=True); e2.record(); ev_last[0] = e2
    return time.perf_counter()
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for rep in range(REPS):
    if rep: print("rep", rep - 1, " ".join("%s=%.1f/%.1f" % (k.split()[0], T[k], G[k]) for k in T))
    T.clear(); G.clear()
    torch.cuda.synchronize()
    ev_last[0] = torch.cuda.Event(enable_timing=True); ev_last[0].record()
    t = time.perf_counter()
    eng = LandmarkEngine.from_site_network(sn); t = tick("engine+tables", t)
    eng.set_frames(frames); t = tick("H2D frames", t)
    eng.reset_status(); src = LandmarkVectorSource(eng)
    seen, gram, src.sparse = eng.pass_stats_cached(); src.seen, src.gram_upper = seen, gram; t = tick("pass A (stats+gram+cache)", t)
    seen_n, cov, graph = gm.landmark_graph(src); t = tick("graph + cov D2H", t)
    m2, nit = markov_clustering_device(graph, inflation=4); t = tick("MCL (%d it)" % nit, t)
    clusters = gm._clusters_on_device(m2); t = tick("clusters from m2", t)
    clusters = [list(c) for c in clusters if seen_n[c[0]] > 0]
    vectors = gm.principal_vectors(gm._covariance_blocks(cov, clusters)); t = tick("eigh host", t)
    flat = gm._FlatClusters(clusters, eng.L); flat.weights = np.concatenate(vectors)
    cid, w = flat.tables(); eng.set_centers(cid, w, len(clusters))
    best = new_best_table(len(clusters), eng.device); src.assign(float('nan'), best=best); t = tick("pass B", t)
    _, rows = read_best_table(best); lv = src.rows(rows); t = tick("best rows refill", t)
    N = F * system.n_mobile
    labels = torch.empty(N, dtype=torch.int64, device="cuda"); confs = torch.empty(N, dtype=torch.float64, device="cuda")
    counts = torch.zeros(len(clusters), dtype=torch.int64, device="cuda")
    src.assign(0.7, counts=counts); t = tick("pass C", t)
    rep_ = torch.zeros((len(clusters), eng.L), dtype=torch.float64, device="cuda"); rw = torch.zeros(len(clusters), dtype=torch.float64, device="cuda")
    sb = new_best_table(len(clusters), eng.device)
    src.assign(0.7, labels=labels, confs=confs, rep=rep_, rep_w=rw, site_best=sb); t = tick("pass D", t)
    l = gm._to_host(labels); c = gm._to_host(confs); t = tick("D2H labels+confs", t)
    eng.site_centers(labels, confs, len(clusters), True, sb); t = tick("site centres", t)
print("F =", F); tot = 0
for k, v in T.items():
    print("  %-32s %8.2f ms wall %8.2f ms gpu-span" % (k, v, G[k])); tot += v
print("  %-32s %8.2f ms" % ("total", tot))
