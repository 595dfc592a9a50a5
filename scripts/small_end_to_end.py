"""Developer tool: a tiny end-to-end run through every kernel family (both clustering plugins, the dynamic lattice map, post-processing)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis
from sitator_b200.dynamics import JumpAnalysis, SmoothSiteTrajectory
for name, F, algo in (("toy_bcc", 60, "mcl"), ("toy_bcc", 60, "dotprod"), ("lgps_dynamic", 6, "mcl")):
    system, cfg = syn.make_config(name)
    frames = system.trajectory(F)
    la = LandmarkAnalysis(clustering_algorithm=algo, verbose=False, dynamic_lattice_mapping=cfg["dynamic"],
                          check_for_zero_landmarks=False, max_mobile_per_site=3)
    st = la.run(syn.site_network_for(system), frames)
    JumpAnalysis().run(st)
    st.jump_array()
    st2 = st.copy(); st2.assign_to_last_known_site(2)
    SmoothSiteTrajectory(remove_unoccupied_sites=False).run(st, 3)
    print(name, algo, st.site_network.n_sites, "sites")
