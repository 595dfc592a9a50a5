"""Developer diagnostic: compare the GPU clustering stages with the oracle on a golden case."""
import sys, logging
sys.path.insert(0, ".")
import numpy as np, torch
from oracle import landmark_oracle as orc
from tests import _util as U
from sitator_b200 import synthetic as syn
from sitator_b200.engine import LandmarkEngine, unpack_key, initial_key
from sitator_b200.landmark.source import LandmarkVectorSource
from sitator_b200.landmark.cluster import mcl as gm
from sitator_b200.util.mcl import markov_clustering_device, clusters_from_matrix

name = sys.argv[1] if len(sys.argv) > 1 else "lgps_dynamic_40"
g, system, cfg, frames = U.load_golden(name)
lv = g["landmark_vectors"]
seen_o, cov_o, graph_o = orc.mcl_graph(lv)
cl_o = orc.markov_clustering(graph_o, inflation=4)
eng = U.engine_for(system, dynamic_lattice_mapping=cfg["dynamic"])
eng.set_frames(frames)
src = LandmarkVectorSource(eng)
seen, cov, graph = gm.landmark_graph(src)
cov = cov.cpu().numpy()
print("seen equal", np.array_equal(seen, seen_o), "cov max abs diff", np.abs(cov - cov_o).max(), "rel", (np.abs(cov - cov_o) / (np.abs(cov_o) + 1e-300)).max())
gg = graph.cpu().numpy()
print("graph max abs diff", np.abs(gg - graph_o).max())
m2, nit = markov_clustering_device(graph, inflation=4)
cl_g = clusters_from_matrix(m2.cpu().numpy())
print("mcl iterations", nit, "clusters gpu", len(cl_g), "oracle", len(cl_o), "same sets", set(cl_g) == set(cl_o), "same order", cl_g == cl_o)
# MCL on the oracle's graph, on the GPU
m2b, nitb = markov_clustering_device(torch.as_tensor(graph_o, device="cuda"), inflation=4)
cl_b = clusters_from_matrix(m2b.cpu().numpy())
print("gpu mcl on oracle graph: same sets", set(cl_b) == set(cl_o), nitb)
if set(cl_g) != set(cl_o):
    a, b = set(cl_g), set(cl_o)
    print("only gpu:", sorted(a - b)[:10]); print("only oracle:", sorted(b - a)[:10])

logging.basicConfig(level=logging.DEBUG)
eng2 = U.engine_for(system, dynamic_lattice_mapping=cfg["dynamic"])
eng2.set_frames(frames)
src2 = LandmarkVectorSource(eng2)
res = gm.do_landmark_clustering(src2, {}, 0.01 / system.n_mobile, False)
res_o = orc.do_landmark_clustering_mcl(lv, {}, 0.01 / system.n_mobile)
print("sites gpu", len(res['cluster-size']), "oracle", len(res_o['cluster-size']))
go = [tuple(c) for c in res_o['cluster-landmark-groupings']]; gg_ = [tuple(c) for c in res['cluster-landmark-groupings']]
missing = [c for c in go if c not in gg_]
print("missing in gpu:", missing)
# oracle-side numbers for the missing clusters
clusters = [list(c) for c in cl_o if seen_o[c[0]] > 0]
for mc in missing:
    cl = list(mc)
    vec = orc.principal_vector(cov_o[cl][:, cl]); centre = np.zeros(lv.shape[1]); centre[cl] = vec
    mb = np.abs(lv @ centre); best = int(np.argmax(mb)); srt = np.sort(mb)
    print("  oracle: best row", best, "best_dot", mb[best], "norm'd", mb[best] / np.linalg.norm(lv[best]), "runner-up", srt[-2], "count rows>0:", (mb > 0).sum())
    vecg = gm.principal_vector(cov[np.ix_(cl, cl)])
    print("  eigvec oracle", vec, "gpu-cov eigvec", vecg)
