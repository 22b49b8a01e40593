"""include/svin_b200.h from C: the header compiles as strict C99, every struct has the layout the ctypes
mirror (svin_b200/capi.py) assumes, and a plain C program pushes one window through svin_ba_optimize."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from svin_b200 import capi
from svin_b200.synthetic import make_window

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    exe = str(tmp_path_factory.mktemp("cabi") / "c_abi_harness")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                           os.path.join(ROOT, "tests", "c_abi_harness.c"), "-I", os.path.join(ROOT, "include"),
                           "-L", libdir, "-lsvin_b200", f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


def test_header_is_valid_c_and_struct_layouts_equal_the_ctypes_mirror(harness):
    out = subprocess.check_output([harness, "layout"], text=True)
    sizes, fields = {}, {}
    for line in out.splitlines():
        t = line.split()
        if len(t) == 2:
            sizes[t[0]] = int(t[1])
        else:
            s, f = t[0].split(".")
            fields.setdefault(s, {})[f] = (int(t[1]), int(t[2]))
    assert len(sizes) >= 19
    checked = 0
    for name, size in sizes.items():
        cls = getattr(capi, name, None)
        assert cls is not None, f"{name} is declared in the header but has no ctypes mirror"
        assert C.sizeof(cls) == size, name
        mirror = {f[0]: getattr(cls, f[0]) for f in cls._fields_}
        assert set(mirror) == set(fields[name]), (name, set(mirror) ^ set(fields[name]))
        for f, (off, sz) in fields[name].items():
            assert (mirror[f].offset, mirror[f].size) == (off, sz), (name, f)
            checked += 1
    assert checked > 200


def _dump(w, path):
    w.finalize()
    cnt = [len(w.pose_blocks), len(w.speedbias), len(w.landmarks), len(w.intrinsics), w.num_obs, int(w.loss_type),
           len(w.imu_pose0), len(w.pose_prior_block), len(w.speedbias_prior_block), len(w.relative_pose_block0),
           len(w.sonar_pose), len(w.depth_pose), len(w.marg_block_kind), int(w.marg_dim), 0, 0]
    ip = capi.SvinImuParams()
    for k, v in w.imu_params.items():
        setattr(ip, k, float(v))
    names = ["pose_blocks", "speedbias", "landmarks", "pose_fixed", "speedbias_fixed", "landmark_fixed", "intrinsics",
             "obs_pose", "obs_landmark", "obs_extrinsics", "obs_camera", "obs_measurement", "obs_information",
             "imu_pose0", "imu_speedbias0", "imu_pose1", "imu_speedbias1", "imu_t0_ns", "imu_t1_ns", "imu_meas_offset",
             "imu_meas_t_ns", "imu_meas_gyro", "imu_meas_accel", "pose_prior_block", "pose_prior_measurement",
             "pose_prior_information", "speedbias_prior_block", "speedbias_prior_measurement",
             "speedbias_prior_information", "marg_block_kind", "marg_block_index", "marg_linearization_points",
             "marg_J", "marg_e0"]
    with open(path, "wb") as f:
        f.write(struct.pack("16i", *cnt))
        f.write(struct.pack("d", float(w.loss_scale)))
        f.write(bytes(ip))
        for n in names:
            b = np.ascontiguousarray(getattr(w, n)).tobytes()
            f.write(struct.pack("q", len(b)))
            f.write(b)


@pytest.mark.gpu
def test_c_program_solves_a_window_through_the_abi(harness, tmp_path):
    from svin_b200.engine import BaEngine
    w, _ = make_window(seed=31, num_keyframes=5, num_imu_frames=3, num_landmarks=400, mode="steady")
    _dump(w, tmp_path / "in.bin")
    r = subprocess.run([harness, "solve", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    summ = capi.SvinBaSummary.from_buffer_copy(raw[:C.sizeof(capi.SvinBaSummary)])
    vals = np.frombuffer(raw[C.sizeof(capi.SvinBaSummary):], dtype=np.float64)
    P, S, L = len(w.pose_blocks), len(w.speedbias), len(w.landmarks)
    pose, sb = vals[:7 * P].reshape(P, 7), vals[7 * P:7 * P + 9 * S].reshape(S, 9)
    lm = vals[7 * P + 9 * S:7 * P + 9 * S + 4 * L].reshape(L, 4)
    q = vals[7 * P + 9 * S + 4 * L:]
    with BaEngine(0) as eng:
        s_py, q_py = eng.optimize([w])
    assert summ.iterations == s_py[0]["iterations"] and summ.termination == s_py[0]["termination"]
    # the same library on the same input; only the order of the fp64 atomics differs from run to run
    assert np.allclose(pose, w.pose_blocks, rtol=1e-8, atol=1e-10) and np.allclose(sb, w.speedbias, rtol=1e-8, atol=1e-10)
    assert np.allclose(lm, w.landmarks, rtol=1e-8, atol=1e-10) and np.allclose(q, q_py[0], rtol=1e-6, atol=1e-9)
