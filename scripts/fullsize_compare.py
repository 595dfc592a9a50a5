"""Compare two scripts/fullsize_run.py outputs of the same (config, total frames) at different GPU counts."""
import json, sys
a, b = (json.load(open(p)) for p in sys.argv[1:3])
assert (a["config"], a["total_frames"]) == (b["config"], b["total_frames"])
import numpy as np
ca, cb = np.asarray(a["site_centers"]), np.asarray(b["site_centers"])
out = {
    "config": a["config"], "total_frames": a["total_frames"], "n_gpus": [a["n_gpus"], b["n_gpus"]],
    "run_ms_warm": [a["run_ms_warm"], b["run_ms_warm"]],
    "sites_equal": a["n_sites"] == b["n_sites"] and a["site_vertex_crc"] == b["site_vertex_crc"],
    "label_blocks_equal": a["label_block_crc32"] == b["label_block_crc32"], "label_blocks": len(a["label_block_crc32"]),
    "n_unassigned": [a["n_unassigned"], b["n_unassigned"]],
    "jump_lists_equal": (a["n_jumps"], a["jump_checksum"]) == (b["n_jumps"], b["jump_checksum"]), "n_jumps": a["n_jumps"],
    "site_centers_max_abs_diff": float(np.max(np.abs(ca - cb))) if ca.shape == cb.shape else None,
    "conf_sum_rel_diff": abs(a["conf_sum"] - b["conf_sum"]) / max(abs(a["conf_sum"]), 1e-300),
    "occupancy_stats_equal": (a["n_multiple_assignments"], a["avg_mobile_per_site"]) == (b["n_multiple_assignments"], b["avg_mobile_per_site"]),
}
print(json.dumps(out))
