#!/usr/bin/env python
"""Headline benchmark: keyframes/s of the sliding-window BA solve (10 KF, 2 k landmarks) on B200.

    python bench.py --gpus N --steps K --warmup W            # CUDA engine (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # CPU path on the box's host cores

A "step" is one pass of the hot path over one batch of synthetic windows (BASELINE.json configs[1]:
EuRoC-shape stereo, 10-keyframe window + 3 IMU frames, 2 000 landmarks, ~10 k observations, Cauchy(1),
dense marginalisation prior; Ceres defaults as set by Estimator::optimize, max 10 iterations).
`value` times the solve with inputs resident in HBM; `e2e` times svin_ba_optimize with HOST buffers
(host pack + H2D + solve + D2H inside the timed region).  One process per GPU; windows are independent
so ranks share nothing on the data path (weak scaling, no collective) — torch.distributed is only the
barrier / max-over-ranks plumbing.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from svin_b200.synthetic import make_window  # noqa: E402
from svin_b200.window import default_options  # noqa: E402

METRIC = "keyframes/sec sliding-window BA solve (10 KF, 2k landmarks)"
UNIT = "keyframes/s"
WORKLOAD = "EuRoC-shape synthetic stereo+IMU, 10-KF window (+3 IMU frames), 2k landmarks (BASELINE configs[1])"
# Algorithmic bytes per observation.  SURVEY.md §8(d) counts 43 B read per observation
# (z 16 + info 8 + idx 12 + landmark ~6.4) and 160 B of materialised r / J_pose / J_lm.  With the compact linearisation
# (r + J_lm planes only, J_pose rebuilt in registers; the default) the per-observation record is 64 B + a 4-byte index.
FUSED = os.environ.get("SVIN_BA_FUSED", "1") != "0"
BYTES_PER_OBS = ({"linearize": 43.0 + 64.0, "schur": 64.0 + 4.0, "backsub": 64.0 + 8.0} if FUSED else
                 {"linearize": 43.0 + 160.0, "schur": 160.0, "backsub": 160.0})


# stdout carries exactly ONE line - the JSON result.  Libraries chat on fd 1 (NCCL prints its version banner there at
# NCCL_DEBUG=VERSION and WARN alike), so fd 1 is pointed at stderr for the whole run and the result goes to a saved copy.
_RESULT_OUT = None


def claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def make_batch(n_windows: int, n_distinct: int, seed0: int = 20260925, prior: str = "random", backend=None):
    """prior = "chain": `n_distinct` consecutive windows of a closed-loop run (svin_b200.synthetic_sequence.
    make_chain_windows: real marginalisation prior from running the window forward through `backend`);
    "random": round 1's generator (random PSD prior, truth (+) noise initial values)."""
    if prior == "chain":
        from svin_b200.synthetic_sequence import make_chain_windows
        base = make_chain_windows(backend, seed=seed0, n_windows=n_distinct)
    else:
        base = [make_window(seed=seed0 + i, num_keyframes=10, num_imu_frames=3, num_landmarks=2000, mode="steady")[0]
                for i in range(n_distinct)]
    return [base[i % n_distinct].copy() for i in range(n_windows)]


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        # NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION (this image's default): keep stdout to the
        # one JSON line the driver parses
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
        return world, rank, local, dist
    return 1, 0, 0, None


def max_over_ranks(dist, value: float, local: int) -> float:
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=f"cuda:{local}" if torch.cuda.is_available() else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}


def ncu_traffic(kernel: str, windows: int, obs_per_window: float):
    """DRAM bytes per launch of a kernel family from the committed `ncu --set full` capture (profiles/), scaled to
    this run's batch and observation count (the per-observation kernels' traffic is proportional to both);
    (None, None) when no capture is committed.  Not measured in this run: ncu replays kernels."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        k = t["kernels"][kernel]
        scale = windows / t["windows"]
        src = t.get("source")
        if t.get("observations_per_window"):
            scale *= obs_per_window / t["observations_per_window"]
            src = "%s (captured at %d windows x %.0f observations, scaled to this run's %d x %.0f)" % (
                src, t["windows"], t["observations_per_window"], windows, obs_per_window)
        return k["dram_bytes_per_launch"] * scale, src
    except (OSError, KeyError, ValueError):
        return None, None


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ front-end
FE_BYTES_DETECT = 372e3     # read 752x480 + write <= 400 keypoints
FE_BYTES_DESCRIBE = 380e3   # patches (<= image) + 400 x 48 B
FE_BYTES_MATCH = 51e3       # one 400 x 400 Hamming top-4 call


def _inv(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def frame_match_problems(seq, feats, fa, fb):
    """The ~17 DenseMatcher::match calls of one dataAssociationAndInitialization between
    frame `fb` (current) and frame `fa` (keyframe / last frame): 3 KF x 2 cams 3D-2D, 2 KF x 2 cams 2D-2D,
    last frame 2 x (3D-2D + 2D-2D), stereo 2D-2D + 3D-2D + 2D-3D."""
    from svin_b200 import capi
    from svin_b200.frontend import MatchProblem
    from svin_b200.synthetic import T_to_pose
    from svin_b200.synthetic_images import wall_point
    W, H = 752, 480
    intr = seq["intrinsics"]
    probs = []

    def p3d2d(f0, c0, f1, c1):
        kA, dA = feats[(f0, c0)]
        kB, dB = feats[(f1, c1)]
        pw = wall_point(seq["T_WC"][f0][c0], c0, np.stack([kA["x"], kA["y"]], axis=1))
        lm = np.concatenate([pw, np.ones((len(pw), 1))], axis=1)
        return MatchProblem(capi.SVIN_MATCH_3D2D, dA, dB, kA, kB, intr[c0], intr[c1], W, H, landmarksA=lm,
                            T_CbW=T_to_pose(_inv(seq["T_WC"][f1][c1])), pose_uncertainty=1e-2)

    def p2d2d(f0, c0, f1, c1):
        kA, dA = feats[(f0, c0)]
        kB, dB = feats[(f1, c1)]
        T = _inv(seq["T_WC"][f0][c0]) @ seq["T_WC"][f1][c1]
        return MatchProblem(capi.SVIN_MATCH_2D2D, dA, dB, kA, kB, intr[c0], intr[c1], W, H, T_CaCb=T_to_pose(T))

    for _ in range(3):
        for c in range(2):
            probs.append(p3d2d(fa, c, fb, c))
    for _ in range(2):
        for c in range(2):
            probs.append(p2d2d(fa, c, fb, c))
    for c in range(2):
        probs.append(p3d2d(fa, c, fb, c))
        probs.append(p2d2d(fa, c, fb, c))
    probs.append(p2d2d(fb, 0, fb, 1))
    probs.append(p3d2d(fb, 0, fb, 1))
    probs.append(p3d2d(fb, 1, fb, 0))
    return probs


def run_frontend(args, local, rank, world, dist, barrier):
    """BRISK detect/describe + the ~17 match calls per stereo frame, batched over `--frames` frames."""
    from svin_b200.frontend import FeEngine
    from svin_b200.synthetic_images import make_stereo_sequence
    F = args.frames
    nd = 4
    seq = make_stereo_sequence(seed=20260925 + rank, n_frames=nd)
    imgs, intr, edir = [], [], []
    for k in range(F):
        f = k % nd
        for c in range(2):
            imgs.append(seq["images"][f][c])
            intr.append(seq["intrinsics"][c])
            edir.append(seq["extraction_dir"][f][c])
    fe = FeEngine(752, 480, max_images=2 * F, device=local)
    fe.upload(imgs, intr, edir)
    for _ in range(3):
        fe.run()
    out = fe.download()
    feats = {(f, c): out[2 * f + c] for f in range(nd) for c in range(2)}
    nkp = float(np.mean([len(k) for k, _ in out]))
    probs = []
    for k in range(F):
        probs += frame_match_problems(seq, feats, k % nd, (k + 1) % nd)
    for _ in range(2):
        fe.match(probs)
    barrier()
    t0 = time.perf_counter()
    dev_detect = dev_match = 0.0
    kms = {}
    for _ in range(args.steps):
        fe.run()
        t = fe.timings()
        dev_detect += t["run_ms"]
        for n, v in t["kernel_ms"].items():
            if n not in ("match", "assign"):      # those belong to svin_match (timed below); here they are stale
                kms[n] = kms.get(n, 0.0) + v
    barrier()
    t_detect = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fe.match(probs)
        t = fe.timings()
        dev_match += t["run_ms"]
        kms["match"] = kms.get("match", 0.0) + t["kernel_ms"]["match"]
        kms["assign"] = kms.get("assign", 0.0) + t["kernel_ms"]["assign"]
    barrier()
    t_match_e2e = time.perf_counter() - t0
    mt = fe.timings()
    # end to end detect: host images in, host keypoints out
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fe.detect_describe(imgs, intr, edir)
    barrier()
    t_detect_e2e = time.perf_counter() - t0
    dt = fe.timings()
    # pipelined end to end: detection of step k+1 (context 1) overlaps the matching of step k (context 2), one host
    # thread each, as ThreadedKFVio runs its per-camera detection threads beside the matching thread
    # (okvis_multisensor_processing/src/ThreadedKFVio.cpp:528-640 frameConsumerLoop / :693-780 matchingLoop)
    import queue
    import threading
    fe2 = FeEngine(752, 480, max_images=2, device=local)
    fe2.match(probs)
    n_pipe = max(args.steps, 8)
    ready: "queue.Queue[int]" = queue.Queue()
    errs = []

    def detect_loop():
        try:
            for k in range(n_pipe):
                fe.detect_describe(imgs, intr, edir)
                ready.put(k)
        except Exception as e:  # noqa: BLE001
            errs.append(e)
            ready.put(-1)

    def match_loop():
        try:
            for _ in range(n_pipe):
                if ready.get() < 0:
                    return
                fe2.match(probs)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    gc.collect()
    gc.disable()        # no collector pause inside the timed region (the match path allocates ~10 k ctypes objects a step)
    barrier()
    t0 = time.perf_counter()
    if args.skip_e2e:   # single-threaded under ncu
        n_pipe = 1
        detect_loop()
        match_loop()
    else:
        th = [threading.Thread(target=detect_loop), threading.Thread(target=match_loop)]
        for x in th:
            x.start()
        for x in th:
            x.join()
    barrier()
    t_pipe = time.perf_counter() - t0
    gc.enable()
    if errs:
        raise errs[0]
    fe2.close()
    e2e_pipe = world * F * n_pipe / max_over_ranks(dist, t_pipe, local)
    dev_s = (dev_detect + dev_match) * 1e-3
    dev_s = max_over_ranks(dist, dev_s, local)
    e2e_s = max_over_ranks(dist, t_detect_e2e + t_match_e2e, local)
    value = world * F * args.steps / dev_s
    e2e = world * F * args.steps / e2e_s
    peak, peak_src = measured_peak_hbm()
    n_img = 2 * F
    harris_gbs = FE_BYTES_DETECT * n_img / (kms["harris"] / args.steps * 1e-3) / 1e9
    res = {
        "metric": "BRISK detect+describe+match stereo frames/s (752x480, <=400 kp/image, 17 match calls/frame)",
        "value": value, "unit": "frames/s", "frames_per_step": F, "keypoints_per_image": nkp,
        "match_calls_per_frame": len(probs) // F,
        "device_ms_per_step": {"detect_describe": dev_detect / args.steps, "match": dev_match / args.steps},
        "wall_ms_per_step_resident_detect": 1e3 * t_detect / args.steps,
        "e2e": {"value": e2e_pipe, "unit": "frames/s",
                "h2d_bytes_per_step": int(dt["h2d_bytes"] + mt["h2d_bytes"]),
                "d2h_bytes_per_step": int(dt["d2h_bytes"] + mt["d2h_bytes"]),
                "ms_per_step": 1e3 * t_pipe / n_pipe, "steps": n_pipe,
                "pipeline": "2 contexts x 1 host thread: detect+describe of step k+1 overlaps the match calls of step k",
                "serial": {"value": e2e, "ms_per_step": {"detect_describe": 1e3 * t_detect_e2e / args.steps,
                                                         "match": 1e3 * t_match_e2e / args.steps}}},
        "kernels_ms_per_step": {k: v / args.steps for k, v in kms.items() if v > 0},
        "roofline": {"bound": "hbm", "kernel": "harris_nms", "achieved": harris_gbs, "peak": peak, "unit": "GB/s",
                     "frac": harris_gbs / peak, "traffic": None, "peak_source": peak_src,
                     "note": "images fit the 126 MB L2 only up to ~340 frames; the score image (4 B/px) is extra "
                             "traffic the algorithmic figure does not count"},
    }
    if rank == 0 and world == 1:
        orc = oracle_solver()
        t0 = time.perf_counter()
        nf = 0
        while time.perf_counter() - t0 < 5.0:
            f = nf % nd
            for c in range(2):
                orc.fe_detect_describe(seq["images"][f][c], seq["intrinsics"][c], seq["extraction_dir"][f][c])
            for p in frame_match_problems(seq, feats, f, (f + 1) % nd):
                orc.fe_match(p)
            nf += 1
        cdt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": nf / cdt, "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": f"{nf} stereo frames (detect+describe both cameras + 17 match calls), "
                                         f"single-thread CPU restatement (oracle/), {cdt:.1f} s; not brisk"}
    fe.close()
    return res


def run_preprocess(args, local):
    """Image pre-processing in front of the detector (Subscriber::imageCallback): config_stereorig_v1 shape,
    1600x1200 raw -> resize 0.5 -> CLAHE(clip 1.0, 2x2 tiles) -> 800x600, batched over `--frames` stereo frames."""
    from svin_b200.preprocess import Preprocessor
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import preprocess_oracle as po
    F = args.frames
    rng = np.random.default_rng(20260925)
    y, x = np.mgrid[0:1200, 0:1600]
    base = [(120 + 90 * np.sin(x / (31.0 + k)) * np.cos(y / 17.0) + rng.normal(0, 20, x.shape)).clip(0, 255)
            .astype(np.uint8) for k in range(4)]
    imgs = [base[k % 4] for k in range(2 * F)]
    kw = dict(resizeFactor=0.5, histogramMethod="CLAHE", claheClipLimit=1.0, claheTilesGridSize=2)
    with Preprocessor(1600, 1200, max_images=2 * F, device=local, **kw) as pre:
        pre.upload(imgs)
        for _ in range(3):
            pre.run()
        dev_ms, kms = 0.0, {}
        for _ in range(args.steps):
            pre.run()
            t = pre.timings()
            dev_ms += t["run_ms"]
            for n, v in t["kernel_ms"].items():
                kms[n] = kms.get(n, 0.0) + v
        out = pre.download()
        pre.process(imgs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pre.process(imgs)
        t_e2e = time.perf_counter() - t0
        tm = pre.timings()
    ok = bool(np.array_equal(out[0], po.preprocess(base[0], 0.5, False, po.HIST_CLAHE, 1.0, 2)))
    t0 = time.perf_counter()
    n_cpu = 0
    while time.perf_counter() - t0 < 3.0:
        po.preprocess(base[n_cpu % 4], 0.5, False, po.HIST_CLAHE, 1.0, 2)
        n_cpu += 1
    cdt = time.perf_counter() - t0
    S, D = 1600 * 1200, 800 * 600
    alg = S + D + D + D + D          # decimate: read S write D; histogram: read D; apply: read D write D
    peak, peak_src = measured_peak_hbm()
    gbs = alg * 2 * F / (dev_ms / args.steps * 1e-3) / 1e9
    return {"metric": "pre-processed stereo frames/s (1600x1200 -> 800x600, CLAHE 2x2, config_stereorig_v1)",
            "value": F * args.steps / (dev_ms * 1e-3), "unit": "frames/s", "frames_per_step": F,
            "device_ms_per_step": dev_ms / args.steps, "bit_exact_vs_oracle": ok,
            "kernels_ms_per_step": {k: v / args.steps for k, v in kms.items()},
            "e2e": {"value": F * args.steps / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(tm["h2d_bytes"]),
                    "d2h_bytes_per_step": int(tm["d2h_bytes"])},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "algorithmic_bytes_per_image": alg, "peak_source": peak_src, "traffic": None},
            "cpu_baseline": {"value": n_cpu / 2 / cdt, "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": f"{n_cpu} images, numpy restatement of the OpenCV chain (oracle/), {cdt:.1f} s; "
                                       "not OpenCV's SIMD code"}}


ORACLE_FLAGS = "-O2 -ffp-contract=off (parity build)"


def oracle_solver():
    """The CPU restatement (oracle/) for the CPU legs: the -O3 -march=native timing build of BASELINE.md §3, compiled on
    the box that runs the bench (falls back to the prebuilt parity build if g++ is not there)."""
    global ORACLE_FLAGS
    if "oracle_lib" not in sys.modules and "SVIN_ORACLE_SO" not in os.environ:
        so = os.path.join(ROOT, "oracle", "_native", "libsvin_oracle_native.so")
        try:
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "native"],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            os.environ["SVIN_ORACLE_SO"] = so
            ORACLE_FLAGS = "-O3 -march=native"
        except Exception:  # noqa: BLE001
            pass
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    return oracle_lib


def latency_section(eng, local):
    """ONE window per svin_ba_optimize call with host buffers - the shape ThreadedKFVio::optimizationLoop produces
    (ThreadedKFVio.cpp:1086: one Estimator::optimize per frame) - for BASELINE configs[1] and configs[0], next to the
    single-thread CPU restatement on the same window."""
    orc = oracle_solver()
    opt = default_options()
    out = {}
    for label, kw in (("configs[1] 10 KF + 3, 2k landmarks", dict(num_keyframes=10, num_landmarks=2000)),
                      ("configs[0] 5 KF + 3, 800 landmarks", dict(num_keyframes=5, num_landmarks=800))):
        w = make_window(seed=20260925, num_imu_frames=3, mode="steady", **kw)[0]
        for _ in range(10):
            eng.optimize([w.copy()], opt)
        wall, dev, up = [], [], []
        for _ in range(100):
            x = w.copy()
            x.c_struct()
            t0 = time.perf_counter()
            summ, _ = eng.optimize([x], opt)
            wall.append(1e3 * (time.perf_counter() - t0))
            tm = eng.timings()
            dev.append(tm["solve_ms"])
            up.append(tm["host_upload_ms"])
        cpu = []
        for _ in range(5):
            x = w.copy()
            x.c_struct()
            t0 = time.perf_counter()
            orc.solve(x, opt, quality=True)
            cpu.append(1e3 * (time.perf_counter() - t0))
        p = lambda a, q: float(np.percentile(a, q))
        out[label] = {"observations": int(w.num_obs), "iterations": int(summ[0]["iterations"]),
                      "e2e_ms": {"p50": p(wall, 50), "p10": p(wall, 10), "p90": p(wall, 90)},
                      "device_solve_ms_p50": p(dev, 50), "host_upload_ms_p50": p(up, 50),
                      "h2d_bytes": int(tm["h2d_bytes"]), "d2h_bytes": int(tm["d2h_bytes"]),
                      "cpu_1thread_ms_p50": p(cpu, 50), "speedup_vs_cpu_1thread": p(cpu, 50) / p(wall, 50)}
    out["note"] = ("one window per call, host buffers in and out, upload + graph replay + download inside the timed "
                   "region; CPU = single-thread restatement (oracle/, %s), not Ceres - it has no intra-solve threading, "
                   "so the reference's 2-thread setting (ThreadedKFVio.cpp:1086) is reported as throughput in "
                   "cpu_baseline.threads" % ORACLE_FLAGS)
    return out


def marginalization_section(eng):
    """B9: svin_ba_marginalize (once per frame on the optimisation thread, ThreadedKFVio.cpp:1115) on the steady-state
    workload - the window + spec of the last applyMarginalizationStrategy call of a closed-loop chain (10 keyframes: the
    oldest keyframe's pose, the oldest speed/bias and the landmarks only they see are folded into the existing prior) -
    and on round 1's larger synthetic case (first frame of a 13-frame 'initial' window, 195-dim), next to the CPU
    restatement."""
    from svin_b200.marginalization import MargSpec, marginalization_subwindow
    from svin_b200.sequence import CudaBackend
    from svin_b200.synthetic_sequence import run_chain
    orc = oracle_solver()

    def timed(sub, spec):
        for _ in range(3):
            eng.upload([sub])
            eng.marginalize(spec)
        t_up, t_m = [], []
        for _ in range(30):
            t0 = time.perf_counter()
            eng.upload([sub])
            t1 = time.perf_counter()
            eng.marginalize(spec)
            t_m.append(1e3 * (time.perf_counter() - t1))
            t_up.append(1e3 * (t1 - t0))
        cpu = []
        for _ in range(5):
            t0 = time.perf_counter()
            orc.marginalize(sub, spec)
            cpu.append(1e3 * (time.perf_counter() - t0))
        return {"landmarks_marginalised": int(sub.num_landmarks), "observations": int(sub.num_obs),
                "dense_dim": int(sub.dense_dim()), "prior_dim": int(spec.c.prior_dim),
                "upload_ms_p50": float(np.median(t_up)), "marginalize_ms_p50": float(np.median(t_m)),
                "cpu_1thread_ms_p50": float(np.median(cpu))}

    sw = run_chain(CudaBackend(eng), n_frames=18)
    out = {"steady_state": timed(sw.last_marg["sub"], sw.last_marg["spec"])}
    w = make_window(seed=20260925, num_keyframes=10, num_imu_frames=3, num_landmarks=2000, mode="initial")[0]
    sub, mp, ms = marginalization_subwindow(w, frames_removed=1)
    out["initial_13_frames"] = timed(sub, MargSpec(sub, mp, ms))
    out["note"] = ("eigen-decompositions on a 4-CTA cluster (one-sided Jacobi, warp per column pair); once per frame, not "
                   "on the per-iteration loop")
    return out


def cpu_baseline_timed(batch, seconds: float):
    """Single-thread oracle (CPU restatement, not Ceres) over the batch's windows until `seconds` have passed."""
    orc = oracle_solver()
    opt = default_options()
    n, t0 = 0, time.perf_counter()
    while True:
        w = batch[n % len(batch)].copy()
        w.c_struct()
        orc.solve(w, opt, quality=True)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n / dt, dt, n


def cpu_baseline(sample, threads: int):
    """Oracle (CPU restatement, not Ceres) on `sample` windows using `threads` host threads -> windows/s."""
    from concurrent.futures import ThreadPoolExecutor
    orc = oracle_solver()
    opt = default_options()
    work = [w.copy() for w in sample]
    for w in work:
        w.c_struct()
    t0 = time.perf_counter()
    if threads == 1:
        for w in work:
            orc.solve(w, opt, quality=True)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda w: orc.solve(w, opt, quality=True), work))
    dt = time.perf_counter() - t0
    return len(work) / dt, dt


def run_reference(args):
    world, rank, local, dist = dist_setup(args.gpus)
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    cores = os.cpu_count() or 1
    n = max(cores, 8)
    orc = oracle_solver()      # builds the -O3 -march=native timing library outside the timed region
    from svin_b200.sequence import OracleBackend
    batch = make_batch(n, min(n, 4), prior=args.prior, backend=OracleBackend(orc))
    for _ in range(args.warmup):
        cpu_baseline(batch[:cores], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_baseline(batch, cores)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "windows_per_step": n, "max_num_iterations": 10,
                   "windows_from": args.prior},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} windows per step, one solve per host thread ({cores} threads); CPU restatement "
                                   f"of the reference path (oracle/, {ORACLE_FLAGS}), not Ceres — Ceres/Eigen are absent "
                                   "from the image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.frames > 0:
        line["frontend"] = reference_frontend(cores)
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def reference_frontend(cores: int, seconds: float = 6.0):
    """The front-end metric on the CPU path: detect + describe (both cameras) + the 17 match calls per stereo frame
    with the oracle, one frame per host thread."""
    from concurrent.futures import ThreadPoolExecutor
    from svin_b200.synthetic_images import make_stereo_sequence
    orc = oracle_solver()
    nd = 4
    seq = make_stereo_sequence(seed=20260925, n_frames=nd)
    feats = {(f, c): orc.fe_detect_describe(seq["images"][f][c], seq["intrinsics"][c], seq["extraction_dir"][f][c])
             for f in range(nd) for c in range(2)}
    probs = [frame_match_problems(seq, feats, f, (f + 1) % nd) for f in range(nd)]

    def one_frame(k):
        f = k % nd
        for c in range(2):
            orc.fe_detect_describe(seq["images"][f][c], seq["intrinsics"][c], seq["extraction_dir"][f][c])
        for p in probs[f]:
            orc.fe_match(p)

    done, t0 = 0, time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        while time.perf_counter() - t0 < seconds:
            list(ex.map(one_frame, range(done, done + cores)))
            done += cores
    dt = time.perf_counter() - t0
    return {"metric": "BRISK detect+describe+match stereo frames/s (752x480, <=400 kp/image, 17 match calls/frame)",
            "value": done / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{done} stereo frames, one per host thread ({cores} threads), CPU restatement (oracle/), "
                      f"{dt:.1f} s; not brisk"}


def run_gpu(args):
    import torch
    world, rank, local, dist = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the svin_b200 engine has no CPU fallback "
                         "(use --impl reference for the CPU path)")
    from svin_b200.engine import BaEngine
    B = args.windows
    opt = default_options()
    eng = BaEngine(local)
    from svin_b200.sequence import CudaBackend
    batch = make_batch(B, args.distinct, seed0=20260925 + 1000 * rank, prior=args.prior, backend=CudaBackend(eng))
    for w in batch:
        w.c_struct()
    n_obs_total = sum(w.num_obs for w in batch)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local)

    # ---------------- HBM-resident solve ("value")
    eng.upload(batch)
    for _ in range(max(args.warmup, 3)):
        eng.reset()
        summaries = eng.solve(opt)
    l0 = eng.timings()["kernel_launches"]
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        eng.reset()
        summaries = eng.solve(opt)
        dev_ms += eng.timings()["solve_ms"]
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    launches = eng.timings()["kernel_launches"] - l0
    dt = max_over_ranks(dist, dt, local)
    value = world * B * args.steps / dt

    e2e_line = None
    if not args.skip_e2e:
        # ---------------- end to end through the C ABI with host buffers ("e2e")
        # serial: one blocking svin_ba_optimize per step.  pipelined: two contexts, each driven by its own host
        # thread (ctypes releases the GIL), so that packing + H2D of step k+1 overlaps the solve of step k.
        from svin_b200.engine import BaPipeline
        pipe_depth = int(os.environ.get("SVIN_PIPE_DEPTH", "3"))
        pipe = BaPipeline(local, depth=pipe_depth)
        for _ in range(2):
            eng.optimize([w.copy() for w in batch], opt)
        pipe.optimize_many([[w.copy() for w in batch] for _ in range(2 * pipe_depth)], opt)
        # enough steps to amortise the pipeline fill (one upload) and drain; fewer when many ranks share the host's memory
        n_e2e = max(args.steps, 24 if world <= 2 else 12)

        def fresh_sets(n=None):
            sets = [[w.copy() for w in batch] for _ in range(n or n_e2e)]
            for s in sets:
                for w in s:
                    w.c_struct()
            return sets

        n_serial = max(args.steps, 8)
        e2e_sets = fresh_sets(n_serial)
        barrier()
        t0 = time.perf_counter()
        for s in e2e_sets:
            eng.optimize(s, opt)
        barrier()
        dt_serial = max_over_ranks(dist, time.perf_counter() - t0, local) * n_e2e / n_serial
        del e2e_sets
        tm = eng.timings()
        e2e_sets = fresh_sets()
        gc.collect()
        gc.disable()
        barrier()
        t0 = time.perf_counter()
        pipe.optimize_many(e2e_sets, opt)
        barrier()
        dt_e2e = max_over_ranks(dist, time.perf_counter() - t0, local)
        gc.enable()
        pipe_trace = pipe.trace
        e2e_value = world * B * n_e2e / dt_e2e
        e2e_serial = world * B * n_e2e / dt_serial
        pipe.close()
        del e2e_sets          # ~7 k windows x 45 arrays: a generation-2 collection over them takes > 100 ms
        gc.collect()
        e2e_line = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(tm["h2d_bytes"]),
                    "d2h_bytes_per_step": int(tm["d2h_bytes"]), "ms_per_step": 1e3 * dt_e2e / n_e2e,
                    "steps": n_e2e,
                    "pipeline": f"BaPipeline: {pipe_depth} contexts x 1 driver thread + {pipe.host_threads - 1} packing workers each, uploads of the next steps overlap the solve of step k", "host_cores": os.cpu_count(),
                    "pipeline_trace_ms": pipe_trace,
                    "serial": {"value": e2e_serial, "ms_per_step": 1e3 * dt_serial / n_e2e},
                    "last_step_breakdown_ms": {k: round(float(tm[k]), 3) for k in (
                        "host_order_ms", "host_fill_ms", "host_upload_ms", "h2d_ms", "solve_ms", "d2h_ms",
                        "host_scatter_ms")}}

    # ---------------- per-kernel times (profiling pass, not part of the timed numbers)
    eng.upload(batch)
    eng.set_profiling(True)
    eng.solve(opt)
    kt = eng.kernel_times()
    prof_summaries = eng.solve(opt)
    kt = eng.kernel_times()
    eng.set_profiling(False)
    iters = np.array([s["iterations"] for s in prof_summaries])
    succ = np.array([s["num_successful_steps"] for s in prof_summaries])
    obs = np.array([w.num_obs for w in batch])
    units = {"linearize": float((obs * (iters + 1)).sum()),
             "schur": float((obs * np.minimum(succ + 1, np.maximum(iters, 1))).sum())}
    units["backsub"] = units["schur"]
    units["step_lm"] = float((obs * iters).sum())
    total_ms = sum(v["ms"] for v in kt.values())
    dominant = max(BYTES_PER_OBS, key=lambda k: kt[k]["ms"])
    peak, peak_src = measured_peak_hbm()
    kern = {}
    for k in BYTES_PER_OBS:
        if kt[k]["launches"]:
            per_launch_bytes = BYTES_PER_OBS[k] * units[k] / kt[k]["launches"]
            avg_ms = kt[k]["ms"] / kt[k]["launches"]
            kern[k] = {"ms_total": kt[k]["ms"], "launches": kt[k]["launches"], "share": kt[k]["ms"] / total_ms,
                       "gbs": per_launch_bytes / (avg_ms * 1e-3) / 1e9}
    for k in kt:
        if k not in kern:
            kern[k] = {"ms_total": kt[k]["ms"], "launches": kt[k]["launches"], "share": kt[k]["ms"] / total_ms}
    achieved = kern[dominant]["gbs"]
    traffic, traffic_src = ncu_traffic(dominant, B, float(obs.mean()))

    frontend = run_frontend(args, local, rank, world, dist, barrier) if args.frames > 0 else None
    sharded = sharded_section(args, local, rank, world, dist)
    loop = None
    if not args.skip_e2e:
        try:
            loop = loop_closure_section(args, local, rank, world, dist)
        except Exception as exc:  # noqa: BLE001 - side section
            loop = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        cpu_val, cpu_dt, cpu_n = cpu_baseline_timed(batch, args.cpu_seconds) if world == 1 else (None, None, 0)
        cpu_threads = None
        if world == 1 and args.cpu_seconds >= 4:
            cores = os.cpu_count() or 1
            n2 = max(2, 2 * int(cpu_val * args.cpu_seconds / 4))          # ~cpu_seconds/2 of work on 2 threads
            nall = max(cores, int(cores * cpu_val * args.cpu_seconds / 4))
            v2, _ = cpu_baseline(batch[:1] * n2, 2)
            vall, _ = cpu_baseline(batch[:1] * nall, cores)
            cpu_threads = {"1": cpu_val, "2": v2, str(cores): vall, "unit": UNIT,
                           "note": "one window per thread; 2 = the reference's Ceres thread count (ThreadedKFVio.cpp:1086)"}
        latency = marg = None
        if world == 1 and not args.skip_e2e:
            try:
                latency = latency_section(eng, local)
                marg = marginalization_section(eng)
            except Exception as exc:  # noqa: BLE001 - side sections never lose the headline line
                latency = {"error": f"{type(exc).__name__}: {exc}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "windows_per_gpu_per_step": B, "distinct_seeds": args.distinct,
                       "windows_from": ("closed-loop chain: %d consecutive windows as Estimator::optimize receives them, prior = "
                                        "svin_ba_marginalize output" % args.distinct) if args.prior == "chain" else
                                       "round-1 generator: random PSD prior, truth (+) noise",
                       "landmarks_per_window": float(np.mean([w.num_landmarks for w in batch])),
                       "observations_per_window": float(obs.mean()), "max_num_iterations": 10,
                       "iterations_mean": float(iters.mean()), "successful_steps_mean": float(succ.mean()),
                       "parallelism": f"{world} independent replica ranks, no collective",
                       "l2": "working set per step (~%.1f GB of Jacobian/state buffers) exceeds the 126 MB L2"
                             % (B * 5.3e-3)},
            "device_ms_per_step": dev_ms / args.steps,
            "e2e": e2e_line,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_OBS[dominant] * units[dominant]
                         / kt[dominant]["launches"],
                         "algorithmic_bytes_per_observation": BYTES_PER_OBS[dominant],
                         "linearisation": "compact (r + J_lm planes, 64 B/obs)" if FUSED else "materialised (160 B/obs)",
                         "note": "the schur family is several launches per slot (one per chunk lane mapping); "
                                 "achieved = algorithmic bytes of a slot / summed duration of its launches"},
            "kernels": kern,
        }
        if frontend is not None:
            line["frontend"] = frontend
            try:   # rank-local side section: never lose the headline line over it
                line["preprocess"] = run_preprocess(args, local)
            except Exception as exc:  # noqa: BLE001
                line["preprocess"] = {"error": f"{type(exc).__name__}: {exc}"}
        if cpu_val is not None:
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{cpu_n} windows of the same batch, single-thread CPU "
                                              f"restatement (oracle/, {ORACLE_FLAGS}), {cpu_dt:.1f} s; not Ceres",
                                    "threads": cpu_threads}
        if latency is not None:
            line["latency"] = latency
        if marg is not None:
            line["marginalization"] = marg
        if sharded is not None:
            line["sharded"] = sharded
        if loop is not None:
            line["loop_closure"] = loop
        emit(line)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def loop_closure_section(args, local, rank, world, dist):
    """BASELINE configs[4] (query part): a 5k-keyframe DBoW2 database (k = 10, L = 6 vocabulary of the shape of
    brief_k10L6.bin, 500 BRIEF-256 descriptors per keyframe), db.query(bowVec, ret, 4, frame_index - 50) per new keyframe.
    Entries are sharded id % N over the ranks; the per-rank top-4 are exchanged with ONE all_gather of 4 (id, score) pairs
    per rank and merged.  The PGO half of configs[4] is not built."""
    import torch
    from svin_b200.loop import LoopEngine, merge_shards
    from svin_b200.synthetic_loop import make_keyframes, random_vocabulary
    n = 5000
    voc = random_vocabulary(10, 6, seed=1)
    frames, _ = make_keyframes(n, 1200, per_image=500, seed=21, revisit_after=1700)
    eng = LoopEngine(voc["first_child"], voc["num_children"], voc["descriptor"], voc["weight"], voc["word_id"], device=local,
                     rank=rank, world=world)
    t0 = time.perf_counter()
    for s_ in range(0, n, 500):
        eng.add(frames[s_:s_ + 500])
    t_add = time.perf_counter() - t0
    qs = list(range(n - 1, n - 201, -1))
    for q in qs[:5]:
        eng.query(frames[q], 4, q - 50)
    dev, merged = [], []
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for q in qs:
        r = eng.query(frames[q], 4, q - 50)
        dev.append(eng.stats()["last_device_ms"])
        if dist is not None:
            mine = torch.full((4, 2), -1.0, dtype=torch.float64, device=f"cuda:{local}")
            for k, (e, sc) in enumerate(r):
                mine[k, 0], mine[k, 1] = e, sc
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            r = merge_shards([[(int(e), float(sc)) for e, sc in t.cpu().tolist() if e >= 0] for t in allr], 4)
        merged.append(r)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dt = max_over_ranks(dist, dt, local)
    out = {"database_keyframes": n, "descriptors_per_keyframe": 500, "vocabulary": "k=10 L=6 synthetic (1,111,111 nodes)",
           "entries_this_rank": eng.stats()["entries_local"], "add_keyframes_per_s": n / t_add,
           "query_ms_e2e_mean": 1e3 * dt / len(qs), "query_device_ms_p50": float(np.median(dev)),
           "queries": len(qs), "exchange": None if dist is None else "one all_gather of 4 (id, score) pairs per rank per query",
           "note": "e2e = host descriptors in, merged top-4 out; device = inverted-file scoring + top-k kernels only"}
    eng.close()
    if rank == 0 and world == 1:
        from oracle import loop_oracle as lo      # CPU leg only
        db = lo.Database(lo.Vocabulary(voc["first_child"], voc["num_children"], voc["descriptor"], voc["weight"],
                                       voc["word_id"]), fast=True)
        m = 1000
        for f in frames[:m]:
            db.add(f)
        t0 = time.perf_counter()
        for q in range(m - 1, m - 6, -1):
            exp = db.query(frames[q], 4, q - 50)
        out["cpu_port"] = {"query_ms": 1e3 * (time.perf_counter() - t0) / 5, "database_keyframes": m, "cores": 1,
                           "kind": "port", "note": "numpy/python restatement on a 1000-keyframe database (building 5k "
                                                   "entries takes 30 s of CPU); an interpreter loop, NOT representative "
                                                   "of C++ DBoW2 - reported for completeness only"}
    return out


def sharded_section(args, local, rank, world, dist):
    """BASELINE configs[3] inside the default run so that the driver's scaling sweep records it: ONE 20-keyframe / 8k-
    landmark window, landmark blocks sharded over the N ranks (l -> rank l % N), three packed all-reduces per iteration
    (NCCL).  Every rank also solves the whole window alone and asserts that the sharded
    solution equals it - the hardware proof of the NCCL path (the 1-GPU test box can only run it through the in-process
    communicator).  N = 1: the single-GPU time only."""
    import torch
    from svin_b200.engine import BaEngine
    from svin_b200.sharding import landmark_owner, shard_window
    try:
        full = make_window(seed=20260925, num_keyframes=20, num_imu_frames=3, num_landmarks=8000, mode="steady")[0]
        opt = default_options()
        steps = max(args.steps, 10)

        def timed(eng, wins):
            eng.upload(wins)
            for _ in range(3):
                eng.reset()
                summ = eng.solve(opt)
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize(local)
            dev = 0.0
            t0 = time.perf_counter()
            for _ in range(steps):
                eng.reset()
                summ = eng.solve(opt)
                dev += eng.timings()["solve_ms"]
            torch.cuda.synchronize(local)
            wall = time.perf_counter() - t0
            eng.download_all()
            return summ[0], max_over_ranks(dist, dev / steps, local), max_over_ranks(dist, 1e3 * wall / steps, local)

        ref = full.copy()
        with BaEngine(local) as e1:
            s1, dev1, wall1 = timed(e1, [ref])
        out = {"workload": "20-KF window (+3 IMU frames), 8k landmarks (BASELINE configs[3])", "n_ranks": world,
               "observations": int(full.num_obs), "iterations": int(s1["iterations"]),
               "single_gpu_ms_per_solve": {"device": dev1, "wall": wall1}}
        if world == 1:
            return out
        mine = shard_window(full, rank, world)
        eng = BaEngine(local)
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            uid = torch.from_numpy(BaEngine.nccl_unique_id().copy()).to(f"cuda:{local}")
        dist.broadcast(uid, src=0)
        eng.comm_init(uid.cpu().numpy(), rank, world)
        sN, devN, wallN = timed(eng, [mine])
        eng.close()
        rel = lambda a, b: float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))
        own = landmark_owner(full.num_landmarks, world) == rank
        errs = [rel(mine.pose_blocks, ref.pose_blocks), rel(mine.speedbias, ref.speedbias),
                rel(mine.landmarks, ref.landmarks[own])]
        ok = float(sN["iterations"] == s1["iterations"] and max(errs) < 1e-7)
        t = torch.tensor([ok, -max(errs)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok_all, worst = bool(t[0].item() > 0.5), float(-t[1].item())
        assert ok_all, f"sharded NCCL solve differs from the single-GPU solve (worst relative difference {worst:.2e})"
        out.update({"sharded_ms_per_solve": {"device": devN, "wall": wallN}, "speedup_device": dev1 / devN,
                    "exchanges_per_iteration": 3, "collective": "ncclAllReduce(sum, f64) x3 per iteration on the solve stream (one packed buffer each)",
                    "parity_vs_single_gpu": {"max_relative_difference": worst, "iterations_equal": True, "tolerance": 1e-7},
                    "scaling": "strong"})
        return out
    except AssertionError:
        raise
    except Exception as exc:  # noqa: BLE001 - never lose the headline line over the side section
        return {"error": f"{type(exc).__name__}: {exc}"}


def run_sharded(args):
    """BASELINE configs[3]: ONE 20-keyframe / 8k-landmark window, landmark blocks sharded over the ranks, the
    reduced system all-reduced with NCCL every iteration (strong scaling of a single solve)."""
    import torch
    world, rank, local, dist = dist_setup(args.gpus)
    from svin_b200.engine import BaEngine
    from svin_b200.sharding import shard_window
    nwin = args.windows if args.windows != 296 else 1
    base = [make_window(seed=20260925 + i, num_keyframes=20, num_imu_frames=3, num_landmarks=8000, mode="steady")[0]
            for i in range(min(nwin, 2))]
    wins = [shard_window(base[i % len(base)], rank, world) for i in range(nwin)]
    for w in wins:
        w.c_struct()
    eng = BaEngine(local)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            uid = torch.from_numpy(BaEngine.nccl_unique_id().copy()).to(f"cuda:{local}")
        dist.broadcast(uid, src=0)
        eng.comm_init(uid.cpu().numpy(), rank, world)
    opt = default_options()
    eng.upload(wins)
    for _ in range(max(args.warmup, 3)):
        eng.reset()
        summ = eng.solve(opt)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(local)
    t0 = time.perf_counter()
    dev = 0.0
    for _ in range(args.steps):
        eng.reset()
        summ = eng.solve(opt)
        dev += eng.timings()["solve_ms"]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(local)
    dt = max_over_ranks(dist, time.perf_counter() - t0, local)
    if rank == 0:
        emit({
            "metric": METRIC.replace("10 KF, 2k", "20 KF, 8k"), "value": nwin * args.steps / dt, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
            "device_ms_per_step": dev / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "20-KF window (+3 IMU frames), 8k landmarks, landmark-block Schur sharded over the "
                                   "ranks, NCCL all-reduce of the reduced system per iteration (BASELINE configs[3])",
                       "windows": nwin, "observations_per_rank": int(np.mean([w.num_obs for w in wins])),
                       "iterations": summ[0]["iterations"], "parallelism": f"landmark shards x{world}"}})
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="svin_b200", choices=["svin_b200", "reference"])
    ap.add_argument("--windows", type=int, default=296, help="windows per GPU per step (2 per SM on a 148-SM B200)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic seeds replicated to fill the batch")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline: solve windows of the batch for this long")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "sharded"],
                    help="replicas: independent windows per GPU (headline); sharded: one window's landmarks split over the GPUs")
    ap.add_argument("--skip-e2e", action="store_true", help="resident solve only (used for the ncu launch list)")
    ap.add_argument("--frames", type=int, default=64, help="stereo frames per front-end step (0 = skip the front-end)")
    ap.add_argument("--prior", default="chain", choices=["chain", "random"],
                    help="chain: windows + marginalisation prior from a closed-loop run (SURVEY 8(d)); random: round 1's "
                         "generator (random PSD prior)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "sharded":
        run_sharded(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
