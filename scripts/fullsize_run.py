"""Full-length runs of the BASELINE configs, frame-sharded over the launched ranks, with evidence that survives the
size:

    [torchrun --nproc-per-node N] scripts/fullsize_run.py --config llzo|laso --total FRAMES --out x.json [--ref-frames K]

* the trajectory is a function of the GLOBAL frame index only (5000-frame chunks seeded by chunk number), so runs at
  different N analyse the same frames: their JSONs carry CRC32s of every 5000-frame block of labels, the jump-list
  checksum and the site vertex sets, and `scripts/fullsize_compare.py a.json b.json` decides "N-GPU run == 1-GPU run";
* `run()` is timed warm (third run) as the max over ranks;
* parity on the CPU-sized prefix: rank 0 runs the COMPILED REFERENCE's fill (helpers._fill_landmark_vectors) and its
  assign step (DotProdClassifier.predict with the site centres this run found) on the first K frames and compares
  landmark vectors, labels, confidences and the jump list of those frames (the reference cannot cluster 10^6 frames:
  its dense matrix would be 672 GB).
"""
import argparse
import json
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis

CHUNK = 5000


def jump_checksum(j):
    if len(j) == 0:
        return 0
    j = j.astype(np.uint64)
    h = (j[:, 0] * np.uint64(0x9E3779B97F4A7C15)) ^ (j[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F)) ^ \
        (j[:, 2] * np.uint64(0x165667B19E3779F9)) ^ (j[:, 3] * np.uint64(0x27D4EB2F165667C5))
    return int(np.bitwise_xor.reduce(h * (j[:, 0] + np.uint64(1))) >> np.uint64(1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="llzo")
    ap.add_argument("--total", type=int, default=1000000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--ref-frames", type=int, default=0)
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=5000, help="frames per seeded trajectory chunk (part of the trajectory's definition)")
    ap.add_argument("--count-zero-vectors", action="store_true",
                    help="check_for_zero_landmarks=False: count all-zero landmark vectors instead of raising (the synthetic "
                         "LASO walk leaves an atom far from every landmark now and then)")
    args = ap.parse_args()
    global CHUNK
    CHUNK = args.chunk
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    system, cfg = syn.make_config(args.config)
    per = args.total // world
    assert per % CHUNK == 0 and per * world == args.total, "frames per rank must be a multiple of %d" % CHUNK
    M, A = system.n_mobile, system.n_total
    t0 = time.perf_counter()
    pinned = torch.empty((per, A, 3), dtype=torch.float64, pin_memory=True)
    frames = pinned.numpy()
    c0 = rank * per // CHUNK
    for c in range(per // CHUNK):
        frames[c * CHUNK:(c + 1) * CHUNK] = system.trajectory(CHUNK, seed=7919 * (c0 + c) + system.seed)
    gen_s = time.perf_counter() - t0
    kw = dict(max_mobile_per_site=max(4, cfg.get("max_mobile_per_site", 1)), dynamic_lattice_mapping=cfg["dynamic"],
              check_for_zero_landmarks=cfg.get("check_for_zero_landmarks", True) and not args.count_zero_vectors)
    sn = syn.site_network_for(system)
    ms = []
    for i in range(args.runs):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **kw)
        st = la.run(sn, frames)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t) * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ms.append(float(dt.item()))
    jumps = st.jump_array()
    crcs = [zlib.crc32(np.ascontiguousarray(st.traj[b:b + CHUNK]).tobytes()) for b in range(0, per, CHUNK)]
    mine = dict(crcs=crcs, n_jumps=int(len(jumps)), jump_checksum=jump_checksum(jumps), conf_sum=float(st.confidences.sum()),
                n_unassigned=int((st.traj < 0).sum()), gen_s=gen_s)
    allr = [None] * world
    if world > 1:
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    out = None
    if rank == 0:
        cs = 0
        for r in allr:
            cs ^= r["jump_checksum"]
        verts = [sorted(int(x) for x in v) for v in st.site_network.vertices]
        out = {
            "config": args.config, "total_frames": args.total, "chunk": CHUNK, "n_gpus": world, "frames_per_gpu": per, "n_atoms": A, "n_mobile": M,
            "n_landmarks": system.n_landmarks, "run_ms_all": ms, "run_ms_warm": ms[-1],
            "frame_atoms_per_s": args.total * A / (ms[-1] * 1e-3), "trajectory_generation_s_per_rank": [r["gen_s"] for r in allr],
            "n_sites": int(st.site_network.n_sites), "site_vertex_crc": zlib.crc32(json.dumps(verts).encode()),
            "site_centers": np.asarray(st.site_network.centers).round(9).tolist(),
            "label_block_crc32": [c for r in allr for c in r["crcs"]],
            "n_unassigned": sum(r["n_unassigned"] for r in allr), "conf_sum": sum(r["conf_sum"] for r in allr),
            "n_jumps": sum(r["n_jumps"] for r in allr), "jump_checksum": cs,
            "n_all_zero_lvecs": int(la.n_all_zero_lvecs), "n_multiple_assignments": int(la.n_multiple_assignments), "avg_mobile_per_site": float(la.avg_mobile_per_site),
            "phases_ms_rank0": la.stats.get("phases_ms"),
        }
        if args.ref_frames:
            out["reference_prefix"] = reference_prefix(system, cfg, frames[:args.ref_frames], la, st, jumps)
        txt = json.dumps(out)
        if args.out:
            open(args.out, "w").write(txt)
        brief = {k: v for k, v in out.items() if k not in ("label_block_crc32", "site_centers", "phases_ms_rank0")}
        print(json.dumps(brief))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def reference_prefix(system, cfg, frames, la, st, jumps):
    """The compiled reference's own fill + predict on the first frames, with the centres of this run."""
    from oracle import ref_loader
    if not ref_loader.available():
        return {"skipped": "oracle/_ref not built"}
    ref = ref_loader.load()
    K, M, L = len(frames), system.n_mobile, system.n_landmarks
    sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
    # the reference's fill, called as LandmarkAnalysis.run calls it (LandmarkAnalysis.py:179-220)
    ra = ref.LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, force_no_memmap=True,
                              dynamic_lattice_mapping=cfg["dynamic"], check_for_zero_landmarks=False)
    captured = {}
    mod = ref.cluster_mcl

    class Stop(Exception):
        pass

    def capture(lv, *a, **k):
        captured["lv"] = np.array(lv)
        raise Stop()
    orig = mod.do_landmark_clustering
    mod.do_landmark_clustering = capture
    t = time.perf_counter()
    try:
        try:
            ra.run(sn, np.ascontiguousarray(frames))
        except Stop:
            pass
    finally:
        mod.do_landmark_clustering = orig
    fill_s = time.perf_counter() - t
    lv_ref = captured["lv"]
    eng = la._engine
    lv = eng.fill_dense(0, K, dtype=torch.float64).cpu().numpy()
    nz = lv_ref != 0
    res = {"frames": K, "reference_fill_s": fill_s, "lv_support_equal": bool(np.array_equal(lv != 0, nz)),
           "lv_max_rel_err": float(np.max(np.abs(lv[nz] - lv_ref[nz]) / lv_ref[nz])) if nz.any() else 0.0}
    cid, w = la.cluster_centers_
    n_sites = int(st.site_network.n_sites)
    centers = np.zeros((n_sites, L))
    sel = cid >= 0
    centers[cid[sel], np.nonzero(sel)[0]] = w[sel]
    clf = ref.DotProdClassifier(threshold=np.nan, min_samples=1)
    clf.set_cluster_centers(centers)
    clf._featuredim = L
    t = time.perf_counter()
    lab_ref, conf_ref = clf.predict(lv_ref, return_confidences=True, threshold=0.7, predict_normed=False, verbose=False)
    res["reference_predict_s"] = time.perf_counter() - t
    lab = st.traj[:K].reshape(-1)
    conf = st.confidences[:K].reshape(-1)
    zero = ~lv_ref.any(axis=1)                 # the reference leaves these confidences uninitialised
    res["labels_differ"] = int((lab != lab_ref).sum())
    res["labels_total"] = int(lab.size)
    ok = (lab == lab_ref) & ~zero
    res["conf_max_abs_diff"] = float(np.max(np.abs(conf[ok] - conf_ref[ok])))
    rst = ref.SiteTrajectory(sn_with_sites(ref, system, n_sites), lab_ref.reshape(K, M).astype(np.int64))
    rj = np.asarray(list(rst.jumps()), dtype=np.int64).reshape(-1, 4)
    mj = jumps[jumps[:, 0] < K]
    res["jump_list_equal"] = bool(np.array_equal(rj, mj))
    res["n_jumps_prefix"] = int(len(rj))
    return res


def sn_with_sites(ref, system, n_sites):
    sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
    sn.centers = np.zeros((n_sites, 3))
    return sn


if __name__ == "__main__":
    main()
