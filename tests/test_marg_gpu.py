"""GPU parity of the marginalisation core against the oracle.

H and b0 are compared entry-wise; J_ and e0_ come out of an eigen-decomposition whose null-space basis is not
unique, so they are compared through what the reference uses them for: J^T J (== H) and J^T e0 (== -b0)."""
import numpy as np
import pytest

import oracle_lib
from svin_b200.marginalization import MargSpec, marginalization_subwindow
from svin_b200.synthetic import make_window

pytestmark = pytest.mark.gpu


def _compare(out, ref, tol_h=1e-9, tol_b=1e-7):
    assert out["dim"] == ref["dim"]
    assert (out["kind"] == ref["kind"]).all() and (out["index"] == ref["index"]).all()
    sH, sb = np.abs(ref["H"]).max(), max(1.0, np.abs(ref["b0"]).max())
    assert np.abs(out["H"] - ref["H"]).max() < tol_h * sH
    assert np.abs(out["b0"] - ref["b0"]).max() < tol_b * sb
    J, e0 = out["J"], out["e0"]
    assert np.abs(J.T @ J - ref["J"].T @ ref["J"]).max() < 1e-8 * sH
    assert np.abs(J.T @ e0 - ref["J"].T @ ref["e0"]).max() < 1e-6 * sb
    assert abs(e0 @ e0 - ref["e0"] @ ref["e0"]) < 1e-6 * max(1.0, ref["e0"] @ ref["e0"])


@pytest.mark.parametrize("seed,extr", [(51, "fixed"), (52, "random_walk")])
def test_marginalize_first_frame(seed, extr):
    from svin_b200.engine import BaEngine
    w, _ = make_window(seed=seed, num_keyframes=4, num_imu_frames=3, num_landmarks=400, mode="initial",
                       extrinsics=extr)
    sub, mp, ms = marginalization_subwindow(w, frames_removed=1)
    assert sub.num_landmarks > 10
    if extr == "random_walk":          # frame 0's two extrinsics blocks leave with it (Estimator.cpp:560-610)
        P = len(w.speedbias)
        mp[P], mp[P + 1] = 1, 1
    spec = MargSpec(sub, mp, ms)
    ref = oracle_lib.marginalize(sub, spec)
    with BaEngine(0) as eng:
        eng.upload([sub])
        out = eng.marginalize(spec)
        _compare(out, ref)
        # landmark part only (nothing dense is marginalised)
        spec0 = MargSpec(sub, np.zeros_like(mp), np.zeros_like(ms))
        _compare(eng.marginalize(spec0), oracle_lib.marginalize(sub, spec0))


def test_marginalize_with_existing_prior():
    from svin_b200.engine import BaEngine
    w, _ = make_window(seed=53, num_keyframes=4, num_imu_frames=3, num_landmarks=300, mode="initial")
    sub, mp, ms = marginalization_subwindow(w, frames_removed=1)
    first = oracle_lib.marginalize(sub, MargSpec(sub, mp, ms))
    # feed the result back as the existing prior of a window that still holds all its blocks
    spec = MargSpec(sub, np.zeros_like(mp), np.zeros_like(ms), first["kind"], first["index"], first["H"], first["b0"])
    ref = oracle_lib.marginalize(sub, spec)
    with BaEngine(0) as eng:
        eng.upload([sub])
        _compare(eng.marginalize(spec), ref)
