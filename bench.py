#!/usr/bin/env python
"""Benchmark of the landmark-analysis hot path (BASELINE.json: frame*atoms/s, landmark + assign).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload: BASELINE.json configs[1] -- synthetic cubic LLZO cell (136 static + 56 mobile = 192 atoms,
1500 landmarks), 10^5 frames per GPU, static lattice (sitator_b200.synthetic 'llzo').

One *step* = one pass of the fused landmark-fill + site-assign kernel (K1, MODE_ASSIGN) over the
rank's 10^5 resident frames with the cluster centres of a previous full analysis.  ``value`` =
frames * atoms * n_gpus / (max over ranks of the CUDA-event time per step).  The resident frames
(461 MB) are larger than L2 (126 MB), so every step streams them from HBM.

``e2e`` = the same metric for the whole ``LandmarkAnalysis(clustering_algorithm='mcl').run(sn, frames)``
through the public API with the frames in pinned HOST memory: host->device copy, all device passes
(Gram, Markov clustering, best-match, two assign passes, site centres, occupancy check) and the
device->host read of labels and confidences are inside the timed region.

``--impl reference`` times the unmodified reference (compiled into oracle/_ref) on the host cores on
bounded samples of the same workload and prints the same line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

# The reference arm pins its BLAS thread count to the same value at every --gpus N (torchrun exports OMP_NUM_THREADS=1
# to its workers; the reference's np.dot / matrix_power calls would then run on one thread at N >= 2 and on all cores at
# N = 1).  OpenBLAS sizes its pool when numpy is imported, so this has to happen first.
REFERENCE_BLAS_THREADS = min(16, os.cpu_count() or 1)
if "reference" in sys.argv:
    os.environ["OMP_NUM_THREADS"] = str(REFERENCE_BLAS_THREADS)
    os.environ["OPENBLAS_NUM_THREADS"] = str(REFERENCE_BLAS_THREADS)
    os.environ["MKL_NUM_THREADS"] = str(REFERENCE_BLAS_THREADS)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "llzo"
METRIC = "frame*atoms/s (landmark fill + site assign)"
UNIT = "frame*atoms/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU (default: the config's 100000)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=400)
    ap.add_argument("--strong-frames", type=int, default=1000000,
                    help="total frames of the strong-scaling leg (BASELINE configs[4]: 10^6 LLZO frames over the N GPUs); 0 = skip")
    ap.add_argument("--assign-mode", default="two_tier", choices=["two_tier", "exact"],
                    help="the timed fill+assign pass: FP32 first tier + float64 for undecided rows (default), or all float64")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: an in-process NVML thread (a sample every
    ~2 ms), or `nvidia-smi -lms 20` when the NVML binding is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []
        self.mask = 0
        self.max_mhz = None
        self._stop = False

    def _nvml_loop(self, nv, handle):
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            import threading
            nv.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES if it remaps them
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.gpu
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.gpu < len(ids) and ids[self.gpu].isdigit():
                    phys = int(ids[self.gpu])
            handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            if self.samples:
                out["sm_mhz"] = float(np.median(self.samples))
                out["sm_max_mhz"] = self.max_mhz
                out["samples"] = len(self.samples)
                out["source"] = "nvml"
            out["reasons"] = sorted(n for n, bit in self.REASONS if self.mask & bit)
            return out
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            mx = [float(r[2]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(max(mx))
                out["samples"] = len(sm)
                out["source"] = "nvidia-smi"
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            seen = set()
            for r in rows:
                if len(r) >= 9:
                    for n, v in zip(names, r[5:9]):
                        if v.strip().lower().startswith("active"):
                            seen.add(n)
            out["reasons"] = sorted(seen)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ----------------------------------------------------------------------------------------------
def reference_run_once(ref, system, cfg, frames):
    from sitator_b200 import synthetic as syn
    sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
    la = ref.LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, force_no_memmap=True,
                              max_mobile_per_site=cfg.get("max_mobile_per_site", 1))
    # time the reference's own fill (helpers.pyx:12, the Cython function this repository replaces) and its assign
    # step (DotProdClassifier.predict, DotProdClassifier.pyx:129-197, called twice by fit_predict) inside its run:
    # the unmodified functions are called through wrappers that only read the clock
    fill = ref.helpers._fill_landmark_vectors
    predict = ref.DotProdClassifier.predict
    spent = [0.0]
    spent_predict = []

    def timed_fill(*a, **k):
        t0 = time.perf_counter()
        try:
            return fill(*a, **k)
        finally:
            spent[0] += time.perf_counter() - t0

    def timed_predict(self, *a, **k):
        t0 = time.perf_counter()
        try:
            return predict(self, *a, **k)
        finally:
            spent_predict.append(time.perf_counter() - t0)
    ref.helpers._fill_landmark_vectors = timed_fill
    ref.DotProdClassifier.predict = timed_predict
    try:
        t = time.perf_counter()
        la.run(sn, frames)
        dt = time.perf_counter() - t
    finally:
        ref.helpers._fill_landmark_vectors = fill
        ref.DotProdClassifier.predict = predict
    fa = spent[0] + sum(spent_predict)
    LAST_REFERENCE_SPLIT["fill_seconds"] = spent[0]
    LAST_REFERENCE_SPLIT["predict_seconds"] = sum(spent_predict)
    LAST_REFERENCE_SPLIT["predict_calls"] = len(spent_predict)
    LAST_REFERENCE_SPLIT["clustering_and_rest_seconds"] = dt - fa
    LAST_REFERENCE_SPLIT["fill_only_value"] = len(frames) * system.n_total / max(spent[0], 1e-9)   # same unit as value
    LAST_REFERENCE_SPLIT["fill_assign_seconds"] = fa
    LAST_REFERENCE_SPLIT["fill_assign_value"] = len(frames) * system.n_total / max(fa, 1e-9)
    return dt


LAST_REFERENCE_SPLIT = {}      # fill / everything else of the last reference_run_once (time them separately)


def load_reference():
    """The compiled reference if oracle/_ref is present (kind 'reference'), else None."""
    try:
        from oracle import ref_loader
        if ref_loader.available():
            import logging
            logging.disable(logging.WARNING)
            try:
                import sklearn.covariance  # noqa: F401  (cluster/mcl.py:22 imports it; do not bill the import)
            except Exception:
                pass
            return ref_loader.load()
    except Exception as e:  # pragma: no cover
        sys.stderr.write("reference unavailable: %r\n" % (e,))
    return None


def port_run_once(system, cfg, frames):
    from oracle import landmark_oracle as orc
    t = time.perf_counter()
    orc.run_landmark_analysis(system.cell, system.static_pos, system.static_idx, system.mobile_idx,
                              system.lm_centers, system.lm_vertices, frames,
                              max_mobile_per_site=cfg.get("max_mobile_per_site", 1))
    return time.perf_counter() - t


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return None


def cpu_baseline(system, cfg, frames, n_frames):
    ref = load_reference()
    sample = np.ascontiguousarray(frames[:n_frames])
    if ref is not None:
        dt = reference_run_once(ref, system, cfg, sample)
        kind = "reference"
    else:
        dt = port_run_once(system, cfg, sample)
        kind = "port"
    return {
        "value": n_frames * system.n_total / dt, "unit": UNIT, "cores": 1, "kind": kind,
        "blas_threads": blas_threads(), "host_cpus": os.cpu_count(), "seconds": dt,
        **({k: round(v, 4) for k, v in LAST_REFERENCE_SPLIT.items()} if kind == "reference" else {}),
        "sample": "whole LandmarkAnalysis.run (mcl) on the first %d frames of the same trajectory; fill and assign "
                  "are single-threaded in the reference, only np.dot/matrix_power use BLAS threads" % n_frames,
    }


def run_reference_arm(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    from sitator_b200 import synthetic as syn
    system, cfg = syn.make_config(WORKLOAD)
    n_steps = args.steps + args.warmup
    # bounded sample per step: the reference's whole run() takes ~26 ms per frame on this shape (most of it the
    # clustering plugin's per-cluster passes over the dense matrix); 400 frames per step unless the step count
    # would push the arm beyond ~4 minutes
    per_step = max(60, min(400, int(240.0 / max(n_steps, 1) / 0.026)))
    frames = system.trajectory(per_step)
    ref = load_reference()
    kind = "reference" if ref is not None else "port"
    times, fa_times = [], []
    for i in range(n_steps):
        dt = reference_run_once(ref, system, cfg, frames) if ref is not None else port_run_once(system, cfg, frames)
        if i >= args.warmup:
            times.append(dt)
            fa_times.append(LAST_REFERENCE_SPLIT.get("fill_assign_seconds", dt))
    t_run = float(np.mean(times))
    t = float(np.mean(fa_times))
    # like for like: `value` = the reference's fill (helpers._fill_landmark_vectors) + its assign step (the two
    # DotProdClassifier.predict calls of fit_predict) -- what our fused fill + assign pass replaces; `e2e` = its whole run()
    value = per_step * system.n_total / t
    e2e_value = per_step * system.n_total / t_run
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(system, per_step), "frames_per_step": per_step,
                   "step": "the reference's fill + assign (helpers._fill_landmark_vectors + both DotProdClassifier.predict calls) "
                           "clocked inside its own LandmarkAnalysis.run (mcl); e2e = that whole run()"
                           if kind == "reference" else "whole run of the NumPy port"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "blas_threads": blas_threads(),
                         "blas_threads_pinned": REFERENCE_BLAS_THREADS, "host_cpus": os.cpu_count(),
                         "whole_run_value": e2e_value, "whole_run_ms_per_step": t_run * 1e3,
                         **({k: round(v, 4) for k, v in LAST_REFERENCE_SPLIT.items()} if kind == "reference" else {}),
                         "sample": "%d frames per step; fill and assign are single-threaded in the reference, only "
                                   "np.dot / matrix_power use the BLAS threads" % per_step},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "ms": t_run * 1e3, "what": "whole LandmarkAnalysis.run (mcl) of the compiled reference"},
    }
    emit(line)


def workload_name(system, n_frames):
    return ("synthetic LLZO-shaped cell (BASELINE configs[1]): %d static + %d mobile = %d atoms, %d landmarks, "
            "%d frames per GPU, static lattice" % (system.n_static, system.n_mobile, system.n_total,
                                                   system.n_landmarks, n_frames))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def work_counters(system, cfg, frames):
    """E (vertex-ratio evaluations under the reference's short-circuit), X (logistic evaluations) and nnz
    per landmark vector, counted once with the oracle by
    scripts/make_work_counters.py and committed as profiles/work_counters.json."""
    w = json.load(open(os.path.join(ROOT, "profiles", "work_counters.json")))[WORKLOAD]
    assert (w["S"], w["M"], w["L"]) == (system.n_static, system.n_mobile, system.n_landmarks)
    return w["E_per_lvec"], w["X_per_lvec"], w["nnz_per_lvec"]


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sitator_b200 import synthetic as syn, _native
    from sitator_b200.landmark import LandmarkAnalysis
    from sitator_b200.landmark.cluster import mcl as cluster_mcl
    from sitator_b200.landmark.source import LandmarkVectorSource
    from sitator_b200.landmark import parallel

    system, cfg = syn.make_config(WORKLOAD)
    F = args.frames or cfg["n_frames"]
    A, M, L = system.n_total, system.n_mobile, system.n_landmarks
    # every rank: its own block of the trajectory, in pinned host memory
    pinned = torch.empty((F, A, 3), dtype=torch.float64, pin_memory=True)
    frames = pinned.numpy()
    chunk = 20000
    for f0 in range(0, F, chunk):
        n = min(chunk, F - f0)
        frames[f0:f0 + n] = system.trajectory(n, seed=1000 * rank + f0 // chunk + system.seed)
    # MCL merges neighbouring true sites of the synthetic hop model into one site, so a site can hold more than
    # one atom; over 10^5..10^6 random frames three on one merged site do occur (seen at 8 ranks), hence 4 here
    # (the parameter only sets where MultipleOccupancyError is raised; it changes no result)
    kw = dict(max_mobile_per_site=max(4, cfg.get("max_mobile_per_site", 1)),
              check_for_zero_landmarks=cfg.get("check_for_zero_landmarks", True))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- e2e: whole run() from pinned host frames (also yields the centres for the kernel steps) -------
    e2e_ms = []
    la = None
    E2E_WARM = 2      # context creation paths, host-pinned result blocks and the memory pool warm up in two runs
    for i in range(E2E_WARM + max(1, args.e2e_steps)):
        la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **kw)
        sn = syn.site_network_for(system)
        barrier()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        st = la.run(sn, frames)
        b.record()
        barrier()
        if i >= E2E_WARM:
            e2e_ms.append(max_over_ranks(a.elapsed_time(b)))
    n_sites = st.site_network.n_sites
    e2e_t = float(np.mean(e2e_ms)) * 1e-3
    # what the e2e number is made of on the host side: all ranks upload their block at once (as inside run()); with N > 1
    # the ranks share the host's memory / PCIe root bandwidth, which is what separates the N-GPU e2e from the 1-GPU one
    # (profiles/r02_h2d_concurrency.md)
    barrier()
    ua = torch.cuda.Event(enable_timing=True); ub = torch.cuda.Event(enable_timing=True)
    ua.record()
    la._engine.set_frames(frames, frame0=la._engine.frame0)
    la._engine.fill_dense(begin=F - 1, n=1)          # a pass over the last frame waits for every upload chunk
    ub.record()
    barrier()
    up_ms = ua.elapsed_time(ub)
    up_max = max_over_ranks(up_ms)
    up_min = -max_over_ranks(-up_ms)
    e2e = {"value": F * A * world / e2e_t, "unit": UNIT, "h2d_bytes_per_step": int(frames.nbytes),
           "d2h_bytes_per_step": int(F * M * 16), "ms": e2e_t * 1e3, "steps": len(e2e_ms),
           "ms_each_step": [round(x, 3) for x in e2e_ms], "n_sites": int(n_sites),
           "upload_alone_ms_slowest_rank": up_max, "upload_alone_ms_fastest_rank": up_min,
           "upload_gbs_slowest_rank": frames.nbytes / (up_max * 1e-3) / 1e9,
           "what": "LandmarkAnalysis(clustering_algorithm='mcl').run(sn, frames) with frames in pinned host memory"}

    # ---- e2e_strong: BASELINE configs[4], a fixed 10^6-frame LLZO trajectory split over the N ranks ------------
    e2e_strong = None
    if args.strong_frames and args.frames is None:
        Fs = args.strong_frames // world
        big = torch.empty((Fs, A, 3), dtype=torch.float64, pin_memory=True)
        bf = big.numpy()
        for f0 in range(0, Fs, F):          # the rank's share, tiled from its weak-scaling block (synthetic data either way)
            n = min(F, Fs - f0)
            bf[f0:f0 + n] = frames[:n]
        sms = []
        STRONG_WARM = 2         # the memory pools and the page-locked result blocks grow to this size in the first two runs
        for i in range(STRONG_WARM + 2):
            las = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **kw)
            sns = syn.site_network_for(system)
            barrier()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            sts = las.run(sns, bf)
            b.record()
            barrier()
            if i >= STRONG_WARM:
                sms.append(max_over_ranks(a.elapsed_time(b)))
        ts = float(np.mean(sms)) * 1e-3
        e2e_strong = {"value": Fs * world * A / ts, "unit": UNIT, "ms": ts * 1e3, "scaling": "strong", "steps": len(sms),
                      "total_frames": Fs * world, "frames_per_gpu": Fs, "h2d_bytes_per_step": int(bf.nbytes),
                      "d2h_bytes_per_step": int(Fs * M * 16), "n_sites": int(sts.site_network.n_sites),
                      "what": "the same run() on BASELINE configs[4]: a fixed 10^6-frame trajectory split over the ranks"}
        del las, sts, big, bf

    # ---- value: the fused fill + assign pass over the resident frames --------------------------------
    eng = la._engine                       # frames of the last run are still resident
    labels = torch.empty(F * M, dtype=torch.int64, device="cuda")
    confs = torch.empty(F * M, dtype=torch.float64, device="cuda")
    counts = torch.zeros(eng.n_clusters, dtype=torch.int64, device="cuda")

    def step():
        eng.pass_assign(0.7, labels=labels, confs=confs, counts=counts)

    # the exact (all float64) pass first: its labels are the reference point for the two-tier pass timed below
    eng.set_assign_mode("exact")
    step()
    torch.cuda.synchronize()
    exact_labels = labels.clone()
    exact_confs = confs.clone()
    ex = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
    for a, b in ex:
        a.record(); step(); b.record()
    torch.cuda.synchronize()
    exact_ms = float(np.mean([a.elapsed_time(b) for a, b in ex]))
    eng.set_assign_mode(args.assign_mode)
    eng.two_tier_info(reset=True)
    for _ in range(max(args.warmup, 3)):
        step()
    eng.two_tier_info(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for a, b in evs:
        a.record(); step(); b.record()
    t_all1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(t_all0.elapsed_time(t_all1))
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    ms_per_step = total_ms / args.steps
    value = F * A * world / (ms_per_step * 1e-3)
    # labels of the timed pass must equal the run's own (same centres, same frames) and the exact pass's
    same = bool(np.array_equal(labels.view(F, M).cpu().numpy(), st.traj))
    same_exact = bool(torch.equal(labels, exact_labels))
    conf_err = float((confs - exact_confs).abs().max().item())
    tt = eng.two_tier_info(reset=True)
    n_passes = args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1) -------------------------------------------------------
    lib = _native.load()
    fp32, fp64, sfu = C.c_double(), C.c_double(), C.c_double()
    _native.check(lib.sitb_microbench(local, C.byref(fp32), C.byref(fp64), C.byref(sfu)))
    E, X, nnz = work_counters(system, cfg, frames)
    S = system.n_static
    ops_per_lvec = 47.0 * S + 2.0 * E + 4.0 * X + 2.0 * nnz
    lvec_per_launch = F * M
    achieved = ops_per_lvec * lvec_per_launch / (kernel_ms * 1e-3)
    peaks, peaks_src = measured_peaks()
    bytes_per_launch = F * (24.0 * A + 16.0 * M)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.isfile(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_frame", 0.0) * F
        except Exception:
            traffic = None
    roofline = {
        "kernel": ("k_assign_fast (fused wrap + lattice check + landmark fill + assign, FP32 first tier) + k_fill<DIAG, MODE_ASSIGN> "
                   "over the rows it leaves undecided" if args.assign_mode == "two_tier" else
                   "k_fill<DIAG, MODE_ASSIGN> (fused wrap + lattice check + landmark fill + assign, float64)"),
        "bound": "fp32", "achieved": achieved / 1e12, "peak": fp32.value / 1e12, "unit": "TFLOP/s",
        "frac": achieved / fp32.value, "traffic": traffic,
        "note": "FLOP = the FP32-pipe lane operations of SURVEY.md 8d (47*S + 2*E + 4*X + 2*nnz per landmark vector, "
                "FMA = 1), peak = FFMA issue rate measured on this GPU by sitb_microbench; K1 is issue-bound, not HBM-bound",
        "ops_per_landmark_vector": ops_per_lvec, "E": E, "X": X, "nnz": nnz,
        "kernel_ms": kernel_ms, "measured_fp64_tflops": fp64.value / 1e12, "measured_sfu_tops": sfu.value / 1e12,
        "hbm": {"bound": "hbm", "achieved": bytes_per_launch / (kernel_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": bytes_per_launch / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "peak_source": peaks_src, "algorithmic_bytes_per_frame": 24.0 * A + 16.0 * M},
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(system, F), "frames_per_gpu": F, "n_sites": int(n_sites),
                   "l2": "resident frames (%d MB) larger than L2 (126 MB); every step streams them from HBM" % (frames.nbytes >> 20),
                   "step": "one K1 pass (fill + assign) over all resident frames, centres from a previous run",
                   "assign_mode": args.assign_mode},
        "e2e": e2e, "e2e_strong": e2e_strong, "gpu_launches": args.steps * (2 if args.assign_mode == "two_tier" else 1), "clocks": clocks, "roofline": roofline,
        "labels_match_run": same,
        "assign_mode": args.assign_mode,
        "two_tier": {
            "labels_equal_exact_pass": same_exact, "conf_max_abs_err_vs_exact_pass": conf_err,
            "exact_pass_ms": exact_ms, "tau": tt["tau"],
            "rows_per_pass": F * M,
            "rows_left_to_exact_kernel_per_pass": {k[len("recheck_"):]: v / float(n_passes) for k, v in tt.items() if k.startswith("recheck_")},
            "what": "first tier: every component in FP32 with a proven error bound; rows whose support, arg-max or threshold "
                    "decision lies inside the bound are redone by the float64 kernel (counted by reason); labels identical by construction",
        },
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(system, cfg, frames, min(args.cpu_frames, F))
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = None


def emit(line):
    """The one JSON line, on the process's real stdout."""
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    a = parse_args()
    # stdout carries exactly one JSON line: whatever libraries write to file descriptor 1 (e.g. NCCL's version banner
    # under torchrun) is sent to stderr instead
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
