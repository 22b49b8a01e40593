"""TEST INFRASTRUCTURE - CPU oracle (numpy) of the RANSAC row.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product path is
svin_b200/csrc/ransac_kernels.cu behind svin_ransac_* (include/svin_b200.h).

What is restated from the reference tree (exactly, file:line):
  * the consensus scores  - FrameAbsolutePoseSacProblem::getSelectedDistancesToModel
                            (okvis_frontend/include/opengv/sac_problems/absolute_pose/FrameAbsolutePoseSacProblem.hpp:131-160),
                            FrameRelativePoseSacProblem (…/relative_pose/FrameRelativePoseSacProblem.hpp:121-157),
                            FrameRotationOnlySacProblem (…/relative_pose/FrameRotationOnlySacProblem.hpp:116-137)
  * the correspondence set - bearing = backProject(keypoint).normalized(), sigmaAngle = sqrt(2) (0.8 size / 12)^2 / fu^2
                            (okvis_frontend/src/FrameNoncentralAbsoluteAdapter.cpp:64-131, FrameRelativeAdapter.cpp:60-190)
  * the callers' decisions - Frontend::runRansac3d2d / runRansac2d2d / runRansac2d2dToRefineScale
                            (okvis_frontend/src/Frontend.cpp:617-676, 832-980, 680-830): threshold 9, 50 iterations,
                            >= 10 inliers, rotation-only vs relative-pose ratio rule.

What lives in OpenGV (pinned at cc32b16 by okvis_ros/okvis/CMakeLists.txt:188-189, NOT under /root/reference) and is
restated from its published description - PARITY UNPINNED against OpenGV itself:
  * sac::Ransac::computeModel: iterate while iterations < k, skip failed models (at most 10 x max_iterations),
    keep the hypothesis with the most inliers (score < threshold), k = log(1 - 0.99) / log(1 - w^sampleSize),
    stop after max_iterations + 1 iterations, inliers = selectWithinDistance(best).
  * minimal solvers.  Absolute pose: sample size 4 (3 + 1 to disambiguate).  The reference asks for GP3P (Kneip's
    generalised P3P, a generated Groebner-basis solver); here the three points of a sample must come from ONE camera and
    the central P3P problem is solved in that camera (Grunert's quartic, Haralick et al. 1994) and moved to the body frame
    with the camera's extrinsics - the case GP3P degenerates to; samples that mix cameras yield no model.  Consensus is
    always scored over ALL correspondences of ALL cameras with the non-central score above.
    Relative pose: sample size 8 (OpenGV: 5 + 3 to disambiguate); OpenGV's STEWENIUS five-point Groebner solver is
    replaced by the EIGHTPT algorithm the same OpenGV problem class offers (Longuet-Higgins / Hartley), on the same 8
    samples.  Rotation only: sample size 2, closed-form alignment of the two bearing pairs (OpenGV twopt_rotationOnly).
  * random sampling: the sample index sets are an INPUT (the caller draws them), so that CPU and GPU evaluate the
    same hypotheses; OpenGV draws them with rand().
This oracle deliberately uses different numerics than the CUDA kernels (np.roots for the quartic, SVD for the
eight-point null space, essential-matrix projection and absolute orientation), so agreement is a real check.
"""
from __future__ import annotations

import numpy as np

PROBABILITY = 0.99


# ------------------------------------------------------------------------------------------ sac::Ransac::computeModel
def ransac_replay(valid, counts, n_corr, sample_size, max_iterations):
    """The sequential loop over pre-scored hypotheses -> index of the winning hypothesis (-1: none)."""
    it, skipped, k, best, best_j = 0, 0, 1.0, -2 ** 31, -1
    max_skip = max_iterations * 10
    eps = np.finfo(np.float64).eps
    for j in range(len(valid)):
        if not (it < k and skipped < max_skip):
            break
        if not valid[j]:
            skipped += 1
            continue
        if counts[j] > best:
            best, best_j = int(counts[j]), j
            w = counts[j] / float(n_corr)
            p_no = 1.0 - w ** sample_size
            p_no = min(max(eps, p_no), 1.0 - eps)
            k = np.log(1.0 - PROBABILITY) / np.log(p_no)
        it += 1
        if it > max_iterations:
            break
    return best_j


# ------------------------------------------------------------------------------------------ scores (in-tree)
def score_absolute(R, t, points, bearings, cam_index, cam_R, cam_t, sigma):
    """FrameAbsolutePoseSacProblem.hpp:131-160; model (R | t) = body pose in the world."""
    body = (points - t) @ R                                    # R^T (p - t)
    rep = np.einsum("nij,ni->nj", cam_R[cam_index], body - cam_t[cam_index])   # R_c^T (.)
    rep = rep / np.linalg.norm(rep, axis=1, keepdims=True)
    e = rep - bearings
    return (e * e).sum(axis=1) / sigma


def triangulate2(R12, t12, f1, f2):
    """opengv::triangulation::triangulate2 (the routine stereo_triangulation.cpp:62-63 says it was adapted from)."""
    f2u = f2 @ R12.T
    b0, b1 = f1 @ t12, f2u @ t12
    a00, a10 = (f1 * f1).sum(1), (f1 * f2u).sum(1)
    a01, a11 = -a10, -(f2u * f2u).sum(1)
    det = a00 * a11 - a01 * a10
    l0 = (a11 * b0 - a01 * b1) / det
    l1 = (-a10 * b0 + a00 * b1) / det
    xm = l0[:, None] * f1
    xn = t12 + l1[:, None] * f2u
    return 0.5 * (xm + xn)


def score_relative(R12, t12, f1, f2, sigma1, sigma2):
    """FrameRelativePoseSacProblem.hpp:121-157."""
    p = triangulate2(R12, t12, f1, f2)
    r1 = p / np.linalg.norm(p, axis=1, keepdims=True)
    r2 = (p - t12) @ R12                                        # R12^T (p - t12)
    r2 = r2 / np.linalg.norm(r2, axis=1, keepdims=True)
    e1, e2 = r1 - f1, r2 - f2
    return (e1 * e1).sum(1) * 0.5 / sigma1 + (e2 * e2).sum(1) * 0.5 / sigma2


def score_rotation(R12, f1, f2, sigma1, sigma2):
    """FrameRotationOnlySacProblem.hpp:116-137."""
    e1 = f2 @ R12.T - f1
    e2 = f1 @ R12 - f2
    return (e1 * e1).sum(1) * 0.5 / sigma1 + (e2 * e2).sum(1) * 0.5 / sigma2


# ------------------------------------------------------------------------------------------ minimal solvers
def p3p_grunert(X, f):
    """Central P3P.  X (3,3) world points, f (3,3) unit bearings in the camera -> list of (R_cw, t_cw) with
    x_cam = R_cw X + t_cw.  Grunert's formulation (Haralick, Lee, Ottenberg, Noelle, IJCV 1994, section 2.1)."""
    a, b, c = np.linalg.norm(X[1] - X[2]), np.linalg.norm(X[0] - X[2]), np.linalg.norm(X[0] - X[1])
    ca, cb, cg = f[1] @ f[2], f[0] @ f[2], f[0] @ f[1]
    if min(a, b, c) < 1e-12:
        return []
    K = (a * a - c * c) / (b * b)
    # s2 = u s1, s3 = v s1;  u = N(v) / D(v)
    N = np.array([K - 1.0, -2.0 * K * cb, 1.0 + K])            # v^2, v, 1
    D = np.array([-2.0 * ca, 2.0 * cg])                        # v, 1
    Q = np.array([1.0, -2.0 * cb, 1.0])
    # 1 + u^2 - 2 u cos(gamma) = (c^2 / b^2) Q   ->   D^2 + N^2 - 2 cos(gamma) N D - (c^2/b^2) Q D^2 = 0
    D2 = np.polymul(D, D)
    P = np.polyadd(np.polyadd(D2, np.polymul(N, N)),
                   np.polyadd(-2.0 * cg * np.polymul(N, D), -(c * c) / (b * b) * np.polymul(Q, D2)))
    out = []
    for v in np.roots(P):
        if abs(v.imag) > 1e-9 * max(1.0, abs(v.real)) or v.real <= 0:
            continue
        v = v.real
        den = np.polyval(D, v)
        if abs(den) < 1e-14:
            continue
        u = np.polyval(N, v) / den
        q = 1.0 + v * v - 2.0 * v * cb
        if u <= 0 or q <= 0:
            continue
        s1 = b / np.sqrt(q)
        Y = np.stack([s1 * f[0], u * s1 * f[1], v * s1 * f[2]])
        # absolute orientation of the two congruent triangles (orthonormal frames)
        def frame(P3):
            e1 = P3[1] - P3[0]
            e1 = e1 / np.linalg.norm(e1)
            e3 = np.cross(e1, P3[2] - P3[0])
            n3 = np.linalg.norm(e3)
            if n3 < 1e-14:
                return None
            e3 = e3 / n3
            return np.stack([e1, np.cross(e3, e1), e3], axis=1)
        Fw, Fc = frame(X), frame(Y)
        if Fw is None or Fc is None:
            continue
        R = Fc @ Fw.T
        out.append((R, Y[0] - R @ X[0]))
    return out


def solve_absolute(sample, points, bearings, cam_index, cam_R, cam_t, sigma):
    """One hypothesis: P3P on sample[:3] (one camera), 4th point picks the solution -> (R, t) body pose in the world."""
    i3 = sample[:3]
    c = cam_index[i3[0]]
    if not (cam_index[i3] == c).all() or len(set(int(x) for x in sample)) < 4:
        return None
    best, best_score = None, np.inf
    for R_cw, t_cw in p3p_grunert(points[i3], bearings[i3]):
        R_bw = cam_R[c] @ R_cw
        t_bw = cam_R[c] @ t_cw + cam_t[c]
        R, t = R_bw.T, -R_bw.T @ t_bw
        k = sample[3:4]
        s = score_absolute(R, t, points[k], bearings[k], cam_index[k], cam_R, cam_t, sigma[k])[0]
        if s < best_score:
            best, best_score = (R, t), s
    return best


def solve_rotation(sample, f1, f2):
    """Rotation aligning two bearing pairs, f1 = R12 f2 (orthonormal triads)."""
    a, b = sample[0], sample[1]
    if a == b:
        return None

    def triad(x, y):
        e3 = np.cross(x, y)
        n = np.linalg.norm(e3)
        if n < 1e-12:
            return None
        e3 = e3 / n
        return np.stack([x, np.cross(e3, x), e3], axis=1)
    B1, B2 = triad(f1[a], f1[b]), triad(f2[a], f2[b])
    if B1 is None or B2 is None:
        return None
    return B1 @ B2.T


def solve_relative(sample, f1, f2):
    """Eight-point algorithm on the 8 samples -> (R12, t12), |t12| = 1, chosen among the four decompositions by the
    number of sample points in front of both cameras."""
    if len(set(int(x) for x in sample)) < 8:
        return None
    a, b = f1[sample], f2[sample]
    A = np.einsum("ni,nj->nij", a, b).reshape(-1, 9)           # f1^T E f2 = 0
    _, _, Vt = np.linalg.svd(A)
    E = Vt[-1].reshape(3, 3)
    U, S, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0, -1.0, 0], [1.0, 0, 0], [0, 0, 1.0]])
    best, best_n = None, -1
    for R in (U @ W @ Vt, U @ W.T @ Vt):
        for t in (U[:, 2], -U[:, 2]):
            p = triangulate2(R, t, a, b)
            z1 = (p * a).sum(1)
            z2 = (((p - t) @ R) * b).sum(1)
            n = int(((z1 > 0) & (z2 > 0)).sum())
            if n > best_n:
                best, best_n = (R, t), n
    return best


# ------------------------------------------------------------------------------------------ whole problems
def ransac_absolute(points, bearings, cam_index, cam_R, cam_t, sigma, samples, threshold=9.0, max_iterations=50):
    n = len(points)
    models, valid, counts = [], [], []
    for s in samples:
        m = solve_absolute(s, points, bearings, cam_index, cam_R, cam_t, sigma)
        models.append(m)
        valid.append(m is not None)
        counts.append(int((score_absolute(m[0], m[1], points, bearings, cam_index, cam_R, cam_t, sigma) < threshold).sum())
                      if m is not None else 0)
    j = ransac_replay(valid, counts, n, 4, max_iterations)
    if j < 0:
        return dict(best=-1, num_inliers=0, inliers=np.zeros(n, bool), R=np.eye(3), t=np.zeros(3), counts=counts, valid=valid)
    R, t = models[j]
    inl = score_absolute(R, t, points, bearings, cam_index, cam_R, cam_t, sigma) < threshold
    return dict(best=j, num_inliers=int(inl.sum()), inliers=inl, R=R, t=t, counts=counts, valid=valid)


def ransac_relative(f1, f2, sigma1, sigma2, samples_rot, samples_rel, threshold=9.0, max_iterations=50):
    """Both RANSACs of runRansac2d2d on one correspondence set -> (rotation-only result, relative-pose result)."""
    n = len(f1)
    out = []
    for kind, samples, size in (("rot", samples_rot, 2), ("rel", samples_rel, 8)):
        models, valid, counts = [], [], []
        for s in samples:
            m = solve_rotation(s, f1, f2) if kind == "rot" else solve_relative(s, f1, f2)
            models.append(m)
            valid.append(m is not None)
            if m is None:
                counts.append(0)
            elif kind == "rot":
                counts.append(int((score_rotation(m, f1, f2, sigma1, sigma2) < threshold).sum()))
            else:
                counts.append(int((score_relative(m[0], m[1], f1, f2, sigma1, sigma2) < threshold).sum()))
        j = ransac_replay(valid, counts, n, size, max_iterations)
        if j < 0:
            out.append(dict(best=-1, num_inliers=0, inliers=np.zeros(n, bool), R=np.eye(3), t=np.zeros(3), counts=counts,
                            valid=valid))
            continue
        if kind == "rot":
            R, t = models[j], np.zeros(3)
            inl = score_rotation(R, f1, f2, sigma1, sigma2) < threshold
        else:
            R, t = models[j]
            inl = score_relative(R, t, f1, f2, sigma1, sigma2) < threshold
        out.append(dict(best=j, num_inliers=int(inl.sum()), inliers=inl, R=R, t=t, counts=counts, valid=valid))
    return out[0], out[1]


def decide_2d2d(rot, rel, n):
    """Frontend.cpp:876-905: which RANSAC wins and whether it counts as a success -> (inlier mask, rotation_only, success)."""
    rr, pr = np.float32(rot["num_inliers"]) / np.float32(n), np.float32(rel["num_inliers"]) / np.float32(n)
    if rr > pr or rr > np.float32(0.8):
        return rot["inliers"], True, rot["num_inliers"] > 10
    return rel["inliers"], False, rel["num_inliers"] > 10
