import sys
sys.path.insert(0, ".")
import numpy as np, torch
from oracle import landmark_oracle as orc
from tests import _util as U
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis
g, system, cfg, frames = U.load_golden("toy_bcc_300")
res = orc.do_landmark_clustering_mcl(g["landmark_vectors"], {}, 0.01 / system.n_mobile)
pbc = orc.PBC(system.cell)
want = orc.site_centers_representative(pbc, system.lm_centers, res["cluster-representative-lvecs"])
kw = U.analysis_kwargs(cfg); kw["site_centers_method"] = 'representative-landmark'
la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **kw)
st = la.run(syn.site_network_for(system), frames)
got = np.asarray(st.site_network.centers)
d = got - want
print("max abs", np.abs(d).max()); bad = np.where(np.abs(d).max(1) > 1e-6)[0]
print("bad sites", bad, "\n got", got[bad], "\n want", want[bad])
eng = la._engine
reps = res["cluster-representative-lvecs"]
got2 = eng.weighted_point_averages(system.lm_centers, reps)
print("kernel on oracle reps: max abs", np.abs(got2 - want).max())
