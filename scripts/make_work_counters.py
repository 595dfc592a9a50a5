"""Count the reference's algorithmic work per landmark vector with the oracle:
E = vertex-ratio evaluations under its short-circuit (helpers.pyx:190-203), X = logistic evaluations
(:205-209), nnz = non-zero components.  Writes profiles/work_counters.json, which bench.py reads for
roofline.achieved (the bench's timed arm itself never imports oracle/)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import landmark_oracle as orc
from sitator_b200 import synthetic as syn

out = {}
for name, nf in (("llzo", 24), ("toy_bcc", 100), ("llzo_v4", 24), ("lgps_dynamic", 24)):
    system, cfg = syn.make_config(name)
    frames = system.trajectory(nf)
    cnt = {}
    orc.fill_landmark_vectors(system.cell, system.static_pos, system.static_idx, system.mobile_idx, system.lm_centers,
                              system.lm_vertices, frames, check_for_zeros=False, counters=cnt,
                              dynamic_lattice_mapping=cfg["dynamic"])
    rows = float(cnt["rows"])
    S = system.n_static
    E, X, nnz = cnt["E"] / rows, cnt["X"] / rows, cnt["nnz"] / rows
    out[name] = {"S": S, "M": system.n_mobile, "A": system.n_total, "L": system.n_landmarks, "frames_sampled": nf,
                 "E_per_lvec": E, "X_per_lvec": X, "nnz_per_lvec": nnz,
                 "fp32_ops_per_lvec": 47.0 * S + 2.0 * E + 4.0 * X + 2.0 * nnz,
                 "sfu_ops_per_lvec": S + 2.0 * X + 2.0 * nnz,
                 "bytes_per_frame": 24.0 * system.n_total + 16.0 * system.n_mobile}
    print(name, out[name])
json.dump(out, open(os.path.join(ROOT, "profiles", "work_counters.json"), "w"), indent=1)
