"""Thin Python handle on the CUDA BA engine (ctypes over include/svin_b200.h).

Mirrors the call sequence of okvis::Estimator::optimize (Estimator.cpp:876-929): the window is
flattened, solved on the device and the solution plus the per-landmark quality is written back.
There is no CPU path here: every method raises SvinError if the CUDA library or device is missing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .window import BaWindow, default_options


class BaEngine:
    def __init__(self, device: int = 0):
        self._lib = capi.load()
        self._ctx = C.c_void_p()
        capi.check(self._lib.svin_ba_create(device, C.byref(self._ctx)), self._lib)
        self._windows = None

    def close(self):
        if self._ctx:
            self._lib.svin_ba_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- staged API -----------------------------------------------------------------------
    def upload(self, windows: list[BaWindow]):
        arr = (capi.SvinBaWindow * len(windows))(*[w.c_struct() for w in windows])
        capi.check(self._lib.svin_ba_upload(self._ctx, arr, len(windows)), self._lib)
        self._windows = windows
        self._arr = arr

    def reset(self):
        capi.check(self._lib.svin_ba_reset(self._ctx), self._lib)

    def solve(self, options: capi.SvinBaOptions | None = None) -> list[dict]:
        opt = options or default_options()
        n = len(self._windows)
        summ = (capi.SvinBaSummary * n)()
        capi.check(self._lib.svin_ba_solve(self._ctx, C.byref(opt), summ), self._lib)
        return [s.as_dict() for s in summ]

    def download(self, index: int, into: BaWindow | None = None):
        w = into or self._windows[index]
        q = np.zeros(w.num_landmarks)
        s = w.c_struct()
        capi.check(self._lib.svin_ba_download(self._ctx, index, C.byref(s), q.ctypes.data_as(capi.c_double_p)),
                   self._lib)
        return w, q

    def download_all(self):
        """All windows of the uploaded batch with one D2H copy -> list of landmark-quality arrays."""
        ws = self._windows
        qall = np.zeros(sum(w.num_landmarks for w in ws))
        ends = np.cumsum([w.num_landmarks for w in ws])
        quals = [qall[e - w.num_landmarks:e] for w, e in zip(ws, ends)]
        qptr = (capi.c_double_p * len(ws))(*[q.ctypes.data_as(capi.c_double_p) for q in quals])
        capi.check(self._lib.svin_ba_download_all(self._ctx, self._arr, len(ws), qptr), self._lib)
        return quals

    def evaluate(self, index: int = 0) -> dict:
        w = self._windows[index]
        n, m = w.num_obs, len(w.imu_pose0)
        out = dict(reproj_residuals=np.zeros((n, 2)), reproj_J_pose=np.zeros((n, 2, 6)),
                   reproj_J_landmark=np.zeros((n, 2, 3)), reproj_J_extrinsics=np.zeros((n, 2, 6)),
                   imu_residuals=np.zeros((m, 15)), imu_J_pose0=np.zeros((m, 15, 6)),
                   imu_J_speedbias0=np.zeros((m, 15, 9)), imu_J_pose1=np.zeros((m, 15, 6)),
                   imu_J_speedbias1=np.zeros((m, 15, 9)), cost=np.zeros(1))
        ev = capi.SvinBaEvaluation()
        for k, v in out.items():
            setattr(ev, k, v.ctypes.data_as(capi.c_double_p))
        capi.check(self._lib.svin_ba_evaluate(self._ctx, index, C.byref(ev)), self._lib)
        return out

    def marginalize(self, spec, index: int = 0) -> dict:
        """MarginalizationError::{addResidualBlock, marginalizeOut, updateErrorComputation} on uploaded window
        `index` (okvis_ceres/src/MarginalizationError.cpp:126-758) -> dict(dim, kind, index, H, b0, J, e0)."""
        from .marginalization import MargResult
        res = MargResult(self._windows[index])
        capi.check(self._lib.svin_ba_marginalize(self._ctx, index, C.byref(spec.c), C.byref(res.c)), self._lib)
        return res.unpack()

    def timings(self) -> dict:
        t = capi.SvinBaTimings()
        capi.check(self._lib.svin_ba_timings(self._ctx, C.byref(t)), self._lib)
        return {n: getattr(t, n) for n, _ in t._fields_}

    # ---- sharded single-window mode -----------------------------------------------------------------
    @staticmethod
    def nccl_unique_id() -> np.ndarray:
        lib = capi.load()
        out = np.zeros(128, dtype=np.uint8)
        capi.check(lib.svin_nccl_unique_id(out.ctypes.data_as(capi.c_uint8_p)), lib)
        return out

    def comm_init(self, unique_id: np.ndarray, rank: int, world: int):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        capi.check(self._lib.svin_ba_comm_init(self._ctx, uid.ctypes.data_as(capi.c_uint8_p), rank, world), self._lib)

    def set_profiling(self, enable: bool):
        capi.check(self._lib.svin_ba_set_profiling(self._ctx, int(enable)), self._lib)

    def kernel_times(self) -> dict:
        t = capi.SvinBaKernelTimes()
        capi.check(self._lib.svin_ba_kernel_times(self._ctx, C.byref(t)), self._lib)
        return {n: dict(ms=t.ms[i], launches=t.launches[i]) for i, n in enumerate(capi.BA_KERNEL_NAMES)}

    # ---- one-shot API: the drop-in for Estimator::optimize ---------------------------------------
    def optimize(self, windows: list[BaWindow], options: capi.SvinBaOptions | None = None):
        """upload + solve + download (host buffers in, host buffers out).  Returns (summaries, qualities)."""
        opt = options or default_options()
        n = len(windows)
        arr = (capi.SvinBaWindow * n)(*[w.c_struct() for w in windows])
        summ = (capi.SvinBaSummary * n)()
        qall = np.zeros(sum(w.num_landmarks for w in windows))     # one allocation, sliced per window
        ends = np.cumsum([w.num_landmarks for w in windows])
        quals = [qall[e - w.num_landmarks:e] for w, e in zip(windows, ends)]
        qptr = (capi.c_double_p * n)(*[q.ctypes.data_as(capi.c_double_p) for q in quals])
        capi.check(self._lib.svin_ba_optimize(self._ctx, arr, n, C.byref(opt), summ, qptr), self._lib)
        self._windows = windows
        self._arr = arr
        return [s.as_dict() for s in summ], quals


def solve_sharded_local(shards: list[BaWindow], options: capi.SvinBaOptions | None = None, device: int = 0):
    """The sharded solve (BASELINE configs[3]) with every rank's context in THIS process on one device: shards[k] is
    rank k's window (svin_b200.sharding.shard_window).  The exchange steps run through svin_ba_comm_init_local; one host
    thread per rank drives upload / solve / download.  Solutions are written into the shards; returns the summaries."""
    import threading
    opt = options or default_options()
    lib = capi.load()
    engines = [BaEngine(device) for _ in shards]
    out: list = [None] * len(shards)
    errors: list = []
    try:
        ctxs = (C.c_void_p * len(engines))(*[e._ctx for e in engines])
        capi.check(lib.svin_ba_comm_init_local(ctxs, len(engines)), lib)

        def drive(k):
            try:
                out[k] = engines[k].optimize([shards[k]], opt)[0][0]
            except Exception as exc:  # noqa: BLE001 - surfaced below
                errors.append(exc)

        threads = [threading.Thread(target=drive, args=(k,)) for k in range(len(engines))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        for e in engines:
            e.close()
    if errors:
        raise errors[0]
    return out


class BaPipeline:
    """Depth-2 software pipeline over the staged C ABI (svin_ba_upload / svin_ba_solve / svin_ba_download_all).

    Two contexts, each driven by its own host thread; a lock serialises the solves so that the packing + H2D copy
    of batch k+1 (context B) runs while batch k (context A) is being solved, instead of both contexts solving at the
    same time.  This is the host-side analogue of ThreadedKFVio's overlap of frontend and optimisation threads
    (okvis_multisensor_processing/src/ThreadedKFVio.cpp:1071-1141), applied across independent windows."""

    def __init__(self, device: int = 0, depth: int = 2, host_threads: int | None = None):
        import os
        import threading
        # every context packs its uploads on its own worker pool; `host_threads` sizes it
        prev = os.environ.get("SVIN_HOST_THREADS")
        # (measured on a 16-core box, 3 contexts: full pools per context 16.7k windows/s, cores / depth 14.9k)
        if host_threads is None:
            # the ranks of a node (torchrun: LOCAL_WORLD_SIZE) share its cores
            host_threads = max(2, min(32, os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
        os.environ["SVIN_HOST_THREADS"] = str(host_threads)
        try:
            self._engines = [BaEngine(device) for _ in range(depth)]
        finally:
            if prev is None:
                del os.environ["SVIN_HOST_THREADS"]
            else:
                os.environ["SVIN_HOST_THREADS"] = prev
        self.host_threads = host_threads
        self._gpu = threading.Lock()

    def close(self):
        for e in self._engines:
            e.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def optimize_many(self, batches: list[list[BaWindow]], options: capi.SvinBaOptions | None = None):
        """Every batch is solved as by BaEngine.optimize; returns [(summaries, qualities)] in batch order."""
        import threading
        import time
        opt = options or default_options()
        out: list = [None] * len(batches)
        errors: list = []
        self.trace = [None] * len(batches)   # per batch: (upload start, upload end, solve start, solve end, done) [s]
        t_origin = time.perf_counter()

        def drive(j):
            eng = self._engines[j]
            try:
                for k in range(j, len(batches), len(self._engines)):
                    t0 = time.perf_counter()
                    eng.upload(batches[k])
                    t1 = time.perf_counter()
                    with self._gpu:
                        t2 = time.perf_counter()
                        summ = eng.solve(opt)
                        t3 = time.perf_counter()
                    out[k] = (summ, eng.download_all())
                    self.trace[k] = tuple(round(1e3 * (t - t_origin), 2) for t in (t0, t1, t2, t3, time.perf_counter()))
            except Exception as exc:  # surfaced to the caller below
                errors.append(exc)

        threads = [threading.Thread(target=drive, args=(j,)) for j in range(len(self._engines))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out
