"""The case compute-sanitizer runs: BASELINE configs[0] (toy BCC) through the whole LandmarkAnalysis.run
(mcl), the jump scan, JumpAnalysis, and the two-tier fused fill + assign pass.  A few hundred frames: sanitizer tools
slow kernels down 10-100x."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis
from sitator_b200.dynamics import JumpAnalysis
F = int(sys.argv[1]) if len(sys.argv) > 1 else 200
system, cfg = syn.make_config("toy_bcc")
frames = system.trajectory(F)
la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, check_for_zero_landmarks=False)
st = la.run(syn.site_network_for(system), frames)
jumps = st.jump_array()
JumpAnalysis().run(st)
eng = la._engine
N = F * system.n_mobile
for mode in ("exact", "two_tier"):
    eng.set_assign_mode(mode)
    labels = torch.empty(N, dtype=torch.int64, device="cuda"); confs = torch.empty(N, dtype=torch.float64, device="cuda")
    counts = torch.zeros(eng.n_clusters, dtype=torch.int64, device="cuda")
    eng.pass_assign(0.7, labels=labels, confs=confs, counts=counts)
    torch.cuda.synchronize()
    assert np.array_equal(labels.cpu().numpy().reshape(F, -1), st.traj)
print("sanitizer case ok: %d frames, %d sites, %d jumps" % (F, st.site_network.n_sites, len(jumps)))
