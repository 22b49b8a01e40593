"""Host-side sliding-window book-keeping: the part of okvis::Estimator that stays on the host.

The engine solves *windows* and computes *marginalisation priors*; which frames, landmarks and residuals make up
the next window is graph book-keeping that the reference does in
  Estimator::addStates                   (okvis_ceres/src/Estimator.cpp:98-411)
  Estimator::addObservation              (okvis_ceres/include/okvis/implementation/Estimator.hpp:48-87)
  Estimator::applyMarginalizationStrategy (okvis_ceres/src/Estimator.cpp:495-814)
and that a drop-in adapter keeps doing there.  This module restates that book-keeping over plain
numpy so that a whole *sequence* - add frame, optimise, marginalise, repeat, as ThreadedKFVio's optimisation loop
does (okvis_multisensor_processing/src/ThreadedKFVio.cpp:1086,1115) and okvis_ceres/test/TestEstimator.cpp:141-192
exercises - can be driven through either the CUDA engine or the CPU oracle from the tests and the bench.

No arithmetic of the hot path lives here: `optimize` and `marginalize` hand a flattened BaWindow to the backend.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .marginalization import MargSpec
from .synthetic import delta_q, quat_mul, quat_to_rot
from .window import BaWindow, default_options

POSE, SB = capi.SVIN_BLOCK_POSE, capi.SVIN_BLOCK_SPEEDBIAS


class Frame:
    def __init__(self, fid, t_ns, keyframe, pose_id, sb_id, ext_ids):
        self.id, self.t_ns, self.keyframe = fid, int(t_ns), bool(keyframe)
        self.pose_id, self.sb_id, self.ext_ids = pose_id, sb_id, list(ext_ids)


class CudaBackend:
    """optimize / marginalize through the C ABI (svin_ba_optimize, svin_ba_marginalize)."""

    def __init__(self, engine):
        self.eng = engine

    def solve(self, w: BaWindow, opt):
        s, q = self.eng.optimize([w], opt)
        return s[0], q[0]

    def marginalize(self, sub: BaWindow, spec: MargSpec):
        self.eng.upload([sub])
        return self.eng.marginalize(spec)


class OracleBackend:
    """The CPU oracle (test infrastructure): only tests and bench.py's CPU legs construct this."""

    def __init__(self, oracle_lib):
        self.o = oracle_lib

    def solve(self, w: BaWindow, opt):
        return self.o.solve(w, opt)

    def marginalize(self, sub: BaWindow, spec: MargSpec):
        return self.o.marginalize(sub, spec)


class SlidingWindow:
    """States, landmarks, error terms and the linear prior of one estimator instance."""

    def __init__(self, intrinsics, T_SC, imu_params, estimate_extrinsics=False, sigma_abs=(1e-3, 1e-4),
                 sigma_rel=(0.0, 0.0), loss_type=capi.SVIN_LOSS_CAUCHY, loss_scale=1.0):
        self.intrinsics = np.asarray(intrinsics, dtype=np.float64)
        self.T_SC = [np.asarray(t, dtype=np.float64) for t in T_SC]
        self.ncam = len(self.T_SC)
        self.imu_params = dict(imu_params)
        self.estimate_extrinsics = bool(estimate_extrinsics)
        self.sigma_abs, self.sigma_rel = sigma_abs, sigma_rel
        self.loss_type, self.loss_scale = loss_type, loss_scale
        self._next_id = 1
        self.pose = {}        # block id -> [x y z qx qy qz qw]
        self.pose_fixed = {}
        self.sb = {}          # block id -> 9
        self.frames: list[Frame] = []
        self.landmarks = {}   # id -> hp(4)
        self.obs = []         # [lm_id, frame_id, cam, z(2), info(4)]
        self.imu_terms = []   # dict(f0, f1, t0, t1, t, gyro, accel)
        self.pose_priors = []  # (pose block id, measurement(7), information(36))
        self.sb_priors = []    # (sb block id, measurement(9), information(81))
        self.rel_pose = []     # (block0, block1, information(36))
        self.prior = None      # dict(blocks [(kind, id)], H, b0, J, e0, lin {(kind, id): point})
        self.quality = {}

    def _new_id(self):
        self._next_id += 1
        return self._next_id - 1

    # ------------------------------------------------------------------ Estimator::addStates
    def add_states(self, t_ns, keyframe, pose_guess, sb_guess, imu=None):
        """New frame.  `imu` = (t_ns[], gyro[][3], accel[][3]) covering [previous frame, this frame] (with margin), linking
        the previous frame's states to this one by an ImuError (Estimator.cpp:196-206).  The very first frame gets
        the PoseError / SpeedAndBiasError priors of Estimator.cpp:319-361 and, if the extrinsics are estimated, their
        absolute priors (:329-344); later frames with relative sigmas > 0 get new extrinsics blocks linked by
        RelativePoseError (:385-403)."""
        fid = self._new_id()
        pid, sid = self._new_id(), self._new_id()
        self.pose[pid] = np.array(pose_guess, dtype=np.float64)
        self.pose_fixed[pid] = 0
        self.sb[sid] = np.array(sb_guess, dtype=np.float64)
        first = not self.frames
        rel = self.estimate_extrinsics and self.sigma_rel[0] > 0.0
        if first:
            ext_ids = []
            for c in range(self.ncam):
                eid = self._new_id()
                self.pose[eid] = self.T_SC[c].copy()
                self.pose_fixed[eid] = 0 if self.estimate_extrinsics else 1
                ext_ids.append(eid)
                if self.estimate_extrinsics:
                    tv, rv = self.sigma_abs[0] ** 2, self.sigma_abs[1] ** 2
                    self.pose_priors.append((eid, self.T_SC[c].copy(), np.diag([1 / tv] * 3 + [1 / rv] * 3).ravel()))
            info = np.zeros((6, 6))
            info[0, 0] = info[1, 1] = info[2, 2] = info[5, 5] = 1.0e8          # Estimator.cpp:321-326
            self.pose_priors.append((pid, self.pose[pid].copy(), info.ravel()))
            sinfo = np.diag([1.0] * 3 + [1.0 / 0.03 ** 2] * 3 + [1.0 / 0.1 ** 2] * 3)   # Estimator.cpp:351-355
            self.sb_priors.append((sid, self.sb[sid].copy(), sinfo.ravel()))
        else:
            prev = self.frames[-1]
            if rel:
                ext_ids = []
                dt = (int(t_ns) - prev.t_ns) * 1e-9
                for c in range(self.ncam):
                    eid = self._new_id()
                    self.pose[eid] = self.pose[prev.ext_ids[c]].copy()
                    self.pose_fixed[eid] = 0
                    ext_ids.append(eid)
                    tv, rv = self.sigma_rel[0] ** 2 * dt, self.sigma_rel[1] ** 2 * dt
                    self.rel_pose.append((prev.ext_ids[c], eid, np.diag([1 / tv] * 3 + [1 / rv] * 3).ravel()))
            else:
                ext_ids = list(prev.ext_ids)
            assert imu is not None
            t, g, a = imu
            self.imu_terms.append(dict(f0=prev.id, f1=fid, t0=prev.t_ns, t1=int(t_ns), t=np.asarray(t, np.int64),
                                       gyro=np.asarray(g, np.float64), accel=np.asarray(a, np.float64)))
        self.frames.append(Frame(fid, t_ns, keyframe, pid, sid, ext_ids))
        return fid

    def add_landmark(self, hp):
        lid = self._new_id()
        self.landmarks[lid] = np.array(hp, dtype=np.float64)
        return lid

    def add_observation(self, lm_id, frame_id, cam, z, size=8.0):
        i = 64.0 / (size * size)                                              # Estimator.hpp impl:64-67
        self.obs.append([lm_id, frame_id, cam, np.array(z, dtype=np.float64), np.array([i, 0, 0, i])])

    def frame(self, fid) -> Frame:
        return next(f for f in self.frames if f.id == fid)

    # ------------------------------------------------------------------ flatten = the adapter's Map walk
    def flatten(self, obs=None, imu_terms=None, pose_priors=None, sb_priors=None, rel_pose=None, prior=None,
                pose_ids=None, sb_ids=None, lm_ids=None, points=None):
        """BaWindow over the given subsets (default: everything = the window Estimator::optimize solves).
        `points`: {(kind, id): value} overriding block values (linearisation points for the marginalisation window).
        Returns (window, maps) with maps = dict(pose=[ids], sb=[ids], lm=[ids])."""
        obs = self.obs if obs is None else obs
        imu_terms = self.imu_terms if imu_terms is None else imu_terms
        pose_priors = self.pose_priors if pose_priors is None else pose_priors
        sb_priors = self.sb_priors if sb_priors is None else sb_priors
        rel_pose = self.rel_pose if rel_pose is None else rel_pose
        prior = self.prior if prior is None else prior
        points = points or {}
        if pose_ids is None:
            pose_ids = [f.pose_id for f in self.frames if f.pose_id is not None]
            seen = set(pose_ids)
            for f in self.frames:
                for e in f.ext_ids:
                    if e not in seen:
                        seen.add(e)
                        pose_ids.append(e)
        if sb_ids is None:
            sb_ids = [f.sb_id for f in self.frames if f.sb_id is not None]
        if lm_ids is None:
            used = {o[0] for o in obs}
            lm_ids = [l for l in self.landmarks if l in used]
        pi = {b: k for k, b in enumerate(pose_ids)}
        si = {b: k for k, b in enumerate(sb_ids)}
        li = {b: k for k, b in enumerate(lm_ids)}
        fr = {f.id: f for f in self.frames}
        w = BaWindow()
        w.imu_params = dict(self.imu_params)
        w.loss_type, w.loss_scale = self.loss_type, self.loss_scale
        w.pose_blocks = np.array([points.get((POSE, b), self.pose[b]) for b in pose_ids]).reshape(-1, 7)
        w.pose_fixed = np.array([self.pose_fixed[b] for b in pose_ids], dtype=np.uint8)
        w.speedbias = np.array([points.get((SB, b), self.sb[b]) for b in sb_ids]).reshape(-1, 9)
        w.speedbias_fixed = np.zeros(len(sb_ids), dtype=np.uint8)
        w.landmarks = np.array([self.landmarks[l] for l in lm_ids]).reshape(-1, 4)
        w.intrinsics = self.intrinsics
        if obs:
            # the order the C++ adapter emits (adapters/EstimatorB200.cpp sorts by (landmark, pose, camera)); svin_ba_upload
            # plans such windows on the device (csrc/ba_plan.cu)
            obs = sorted(obs, key=lambda o: (li[o[0]], pi[fr[o[1]].pose_id], o[2]))
            w.obs_pose = [pi[fr[o[1]].pose_id] for o in obs]
            w.obs_landmark = [li[o[0]] for o in obs]
            w.obs_extrinsics = [pi[fr[o[1]].ext_ids[o[2]]] for o in obs]
            w.obs_camera = [o[2] for o in obs]
            w.obs_measurement = np.array([o[3] for o in obs])
            w.obs_information = np.array([o[4] for o in obs])
        off, mt, mg, ma = [0], [], [], []
        for t in imu_terms:
            mt.append(t["t"]); mg.append(t["gyro"]); ma.append(t["accel"])
            off.append(off[-1] + len(t["t"]))
        w.imu_pose0 = [pi[fr[t["f0"]].pose_id] for t in imu_terms]
        w.imu_pose1 = [pi[fr[t["f1"]].pose_id] for t in imu_terms]
        w.imu_speedbias0 = [si[fr[t["f0"]].sb_id] for t in imu_terms]
        w.imu_speedbias1 = [si[fr[t["f1"]].sb_id] for t in imu_terms]
        w.imu_t0_ns = [t["t0"] for t in imu_terms]
        w.imu_t1_ns = [t["t1"] for t in imu_terms]
        w.imu_meas_offset = off
        if imu_terms:
            w.imu_meas_t_ns, w.imu_meas_gyro, w.imu_meas_accel = np.concatenate(mt), np.vstack(mg), np.vstack(ma)
        w.pose_prior_block = [pi[b] for b, _, _ in pose_priors]
        w.pose_prior_measurement = [m for _, m, _ in pose_priors]
        w.pose_prior_information = [i for _, _, i in pose_priors]
        w.speedbias_prior_block = [si[b] for b, _, _ in sb_priors]
        w.speedbias_prior_measurement = [m for _, m, _ in sb_priors]
        w.speedbias_prior_information = [i for _, _, i in sb_priors]
        w.relative_pose_block0 = [pi[a] for a, _, _ in rel_pose]
        w.relative_pose_block1 = [pi[b] for _, b, _ in rel_pose]
        w.relative_pose_information = [i for _, _, i in rel_pose]
        if prior:
            w.marg_block_kind = [k for k, _ in prior["blocks"]]
            w.marg_block_index = [pi[b] if k == POSE else si[b] for k, b in prior["blocks"]]
            w.marg_linearization_points = np.concatenate([prior["lin"][kb] for kb in prior["blocks"]])
            w.marg_J, w.marg_e0 = prior["J"].ravel(), prior["e0"]
            w.marg_dim = len(prior["e0"])
        w.finalize()
        return w, dict(pose=pose_ids, sb=sb_ids, lm=lm_ids)

    # ------------------------------------------------------------------ Estimator::optimize
    def optimize(self, backend, options=None):
        opt = options or default_options()
        w, maps = self.flatten()
        summary, quality = backend.solve(w, opt)
        for k, b in enumerate(maps["pose"]):
            self.pose[b] = w.pose_blocks[k].copy()
        for k, b in enumerate(maps["sb"]):
            self.sb[b] = w.speedbias[k].copy()
        for k, l in enumerate(maps["lm"]):
            self.landmarks[l] = w.landmarks[k].copy()
            self.quality[l] = float(quality[k])
        return summary, w

    # ------------------------------------------------------------------ Estimator::applyMarginalizationStrategy
    def apply_marginalization_strategy(self, backend, num_keyframes, num_imu_frames):
        """Estimator.cpp:495-814.  Returns the ids of the removed landmarks (None if nothing was to do)."""
        if len(self.frames) <= num_imu_frames:
            return None                                                       # :499-506
        older = self.frames[:-num_imu_frames][::-1]                            # newest first, like the reverse iterator
        remove_frames, counted = [], 0
        for f in older:                                                       # :528-538
            if (not f.keyframe) or counted >= num_keyframes:
                remove_frames.append(f)
            else:
                counted += 1
        linearized = {f.id for f in older}
        removed_ids = {f.id for f in remove_frames}
        order = {f.id: k for k, f in enumerate(self.frames)}
        marg_pose, marg_sb = [], []
        lin_imu, lin_sbp, lin_pp, lin_rp = [], [], [], []

        def take(lst, pred, into):
            keep = []
            for x in lst:
                (into if pred(x) else keep).append(x)
            lst[:] = keep

        fr = {f.id: f for f in self.frames}
        # ---- everything but the pose of every frame older than the IMU window (:541-614)
        for f in older:
            if f.sb_id is not None:
                sid = f.sb_id
                marg_sb.append(sid)
                take(self.imu_terms, lambda t: fr[t["f0"]].sb_id == sid or fr[t["f1"]].sb_id == sid, lin_imu)
                take(self.sb_priors, lambda p: p[0] == sid, lin_sbp)
        # ---- poses (and per-frame extrinsics) of the frames that leave (:616-672)
        redo_fixation = False
        for f in remove_frames:
            marg_pose.append(f.pose_id)
            before = len(self.pose_priors)
            self.pose_priors[:] = [p for p in self.pose_priors if p[0] != f.pose_id]   # PoseError: removed, not linearised
            redo_fixation |= len(self.pose_priors) != before
            take(self.imu_terms, lambda t: t["f0"] == f.id or t["f1"] == f.id, lin_imu)
            nxt = self.frames[order[f.id] + 1]
            for c, eid in enumerate(f.ext_ids):
                if self.pose_fixed[eid] or nxt.ext_ids[c] == eid:
                    continue
                marg_pose.append(eid)
                take(self.pose_priors, lambda p: p[0] == eid, lin_pp)
                take(self.rel_pose, lambda r: r[0] == eid or r[1] == eid, lin_rp)
        # ---- observations (:674-772)
        current_kf = older[0].id
        by_lm = {}
        for o in self.obs:
            by_lm.setdefault(o[0], []).append(o)
        keep_obs, lin_obs, removed_landmarks = [], [], []
        for lid in list(self.landmarks):
            res = by_lm.get(lid, [])
            if not res:
                if remove_frames:
                    del self.landmarks[lid]
                    removed_landmarks.append(lid)
                continue
            if not remove_frames or not any(o[1] in removed_ids for o in res):
                keep_obs += res
                continue
            has_new = any(o[1] >= current_kf for o in res)
            marginalize = not has_new
            obs_count = sum(1 for o in res if o[1] in linearized)
            kept, added = [], []
            for o in res:
                if (o[1] in removed_ids and has_new) or (o[1] not in linearized and marginalize):
                    continue                                                  # removeObservation
                if marginalize and o[1] in linearized:
                    if obs_count >= 2:
                        added.append(o)
                    continue
                kept.append(o)
            if not kept and not added:
                del self.landmarks[lid]                                        # justDelete
                removed_landmarks.append(lid)
            elif marginalize and added:
                lin_obs += added
                removed_landmarks.append(lid)                                  # erased after the numeric step below
            else:
                keep_obs += kept
        self.obs = keep_obs
        # ---- numeric core: MarginalizationError::{addResidualBlock, marginalizeOut, updateErrorComputation}
        if marg_pose or marg_sb:
            self._marginalize_numeric(backend, marg_pose, marg_sb, lin_obs, lin_imu, lin_pp, lin_sbp, lin_rp)
        for lid in removed_landmarks:
            self.landmarks.pop(lid, None)
        for f in older:
            f.sb_id = None if f.sb_id in marg_sb else f.sb_id
        for b in marg_sb:
            self.sb.pop(b, None)
        self.frames = [f for f in self.frames if f.id not in removed_ids]
        live = {e for f in self.frames for e in f.ext_ids} | {f.pose_id for f in self.frames}
        for b in marg_pose:
            if b not in live:
                self.pose.pop(b, None)
        if redo_fixation:                                                      # :800-811
            p0 = self.frames[0].pose_id
            info = np.zeros((6, 6))
            info[0, 0] = info[1, 1] = info[2, 2] = info[5, 5] = 1.0e14
            self.pose_priors.append((p0, self.pose[p0].copy(), info.ravel()))
        return removed_landmarks

    def _marginalize_numeric(self, backend, marg_pose, marg_sb, lin_obs, lin_imu, lin_pp, lin_sbp, lin_rp):
        fr = {f.id: f for f in self.frames}
        prior = self.prior
        pose_ids, sb_ids = [], []

        def need_pose(b):
            if b not in pose_ids:
                pose_ids.append(b)

        def need_sb(b):
            if b not in sb_ids:
                sb_ids.append(b)

        if prior:
            for k, b in prior["blocks"]:
                (need_pose if k == POSE else need_sb)(b)
        for t in lin_imu:
            need_pose(fr[t["f0"]].pose_id); need_pose(fr[t["f1"]].pose_id)
            need_sb(fr[t["f0"]].sb_id); need_sb(fr[t["f1"]].sb_id)
        for b, _, _ in lin_pp:
            need_pose(b)
        for b, _, _ in lin_sbp:
            need_sb(b)
        for a, b, _ in lin_rp:
            need_pose(a); need_pose(b)
        for o in lin_obs:
            need_pose(fr[o[1]].pose_id); need_pose(fr[o[1]].ext_ids[o[2]])
        for b in marg_pose:
            need_pose(b)
        for b in marg_sb:
            need_sb(b)
        lm_ids = []
        for o in lin_obs:
            if o[0] not in lm_ids:
                lm_ids.append(o[0])
        points = dict(prior["lin"]) if prior else {}        # first-estimate points of the blocks already in the prior
        sub, maps = self.flatten(obs=lin_obs, imu_terms=lin_imu, pose_priors=lin_pp, sb_priors=lin_sbp, rel_pose=lin_rp,
                                 prior={}, pose_ids=pose_ids, sb_ids=sb_ids, lm_ids=lm_ids, points=points)
        mp = np.array([1 if b in marg_pose else 0 for b in pose_ids], dtype=np.uint8)
        ms = np.array([1 if b in marg_sb else 0 for b in sb_ids], dtype=np.uint8)
        pi = {b: k for k, b in enumerate(pose_ids)}
        si = {b: k for k, b in enumerate(sb_ids)}
        if prior:
            spec = MargSpec(sub, mp, ms, [k for k, _ in prior["blocks"]],
                            [pi[b] if k == POSE else si[b] for k, b in prior["blocks"]], prior["H"], prior["b0"])
        else:
            spec = MargSpec(sub, mp, ms)
        res = backend.marginalize(sub, spec)
        self.last_marg = dict(sub=sub, spec=spec, result=res)
        if res["dim"] == 0:
            self.prior = None
            return
        blocks = [(int(k), (pose_ids if k == POSE else sb_ids)[int(i)]) for k, i in zip(res["kind"], res["index"])]
        lin = {}
        for kb in blocks:
            k, b = kb
            lin[kb] = points[kb].copy() if kb in points else (self.pose[b] if k == POSE else self.sb[b]).copy()
        self.prior = dict(blocks=blocks, H=res["H"], b0=res["b0"], J=res["J"], e0=res["e0"], lin=lin)


# ---------------------------------------------------------------------------------------------- synthetic sequences
def propagate(pose, sb, t, gyro, accel, t0, t1, g):
    """Initial guess of a new frame's states from the IMU (the role of ImuError::propagation in addStates,
    Estimator.cpp:118-134); plain mid-point integration - only a starting value, not part of the parity surface."""
    r, q, v = pose[:3].copy(), pose[3:7].copy(), sb[:3].copy()
    bg, ba = sb[3:6], sb[6:9]
    sel = np.nonzero((t >= t0) & (t <= t1))[0]
    tt = np.concatenate([[t0], t[sel], [t1]]) if len(sel) else np.array([t0, t1])
    for k in range(len(tt) - 1):
        dt = (tt[k + 1] - tt[k]) * 1e-9
        if dt <= 0:
            continue
        i = int(np.clip(np.searchsorted(t, tt[k], side="right") - 1, 0, len(t) - 1))
        w, a = gyro[i] - bg, accel[i] - ba
        C = quat_to_rot(q)
        acc_w = C @ a - np.array([0, 0, g])
        r = r + v * dt + 0.5 * acc_w * dt * dt
        v = v + acc_w * dt
        q = quat_mul(q, delta_q(w * dt))
        q = q / np.linalg.norm(q)
    out_sb = sb.copy()
    out_sb[:3] = v
    return np.concatenate([r, q]), out_sb
