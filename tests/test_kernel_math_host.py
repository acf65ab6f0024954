"""The Float64 math of the kernels (device_utils.cuh / se3_common.cuh: small-angle sin/cos, angle wrapping, the seeded
square root, polynomial quaternion Exp, series Log, rotation-vector representatives) compiled for the HOST from the very
same source text and swept densely over its whole domain, including the thresholds where the kernels switch between the
polynomial fast paths and the general paths.  The GPU parity tests sample these paths; this sweeps them."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HM = os.path.join(ROOT, "tests", "host_math")
sys.path.insert(0, HM)


@pytest.fixture(scope="module")
def report(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    from extract import extract
    csrc = os.path.join(ROOT, "rome.jl_b200", "csrc")
    du = open(os.path.join(csrc, "device_utils.cuh")).read()
    se = open(os.path.join(csrc, "se3_common.cuh")).read()
    body = extract(du, ["kPi", "kTwoPi", "kTwoPiLo", "kInvTwoPi", "wrap_pi", "sym_rem", "kSinC", "kCosC", "kOddInv",
                        "kSmallAngle", "sincos_small", "sqrt_seeded", "closest_rotvec", "sincos_anchored"])
    body += extract(se, ["Quat", "kPi2", "exp_scale_general", "exp_scale_poly", "quat_exp", "qmul", "qconj",
                         "log_scale_general", "kSmallLog", "log_scale_small", "quat_pos", "quat_log_any", "quat_rotate"])
    d = tmp_path_factory.mktemp("host_math")
    src = d / "host_math.cpp"
    src.write_text(open(os.path.join(HM, "shim.inc")).read() + body + open(os.path.join(HM, "harness_main.inc")).read())
    exe = d / "host_math"
    # -ffp-contract=off: only the fma() calls written in the source fuse, as in the device build of these functions
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-D_GNU_SOURCE", str(src), "-o", str(exe), "-lm"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    rep = {}
    for ln in out.splitlines():
        k, *v = ln.split()
        rep[k] = v
    return rep


def test_small_angle_sincos_and_wrapping(report):
    s, c = map(float, report["sincos_small_abs"])
    assert s < 2.3e-16 and c < 2.3e-16          # fdlibm kernels: < 1 ulp on |x| <= 0.78
    assert float(report["sincos_anchored_abs"][0]) < 4.5e-16
    assert float(report["wrap_pi_abs"][0]) < 5e-13 and float(report["wrap_pi_abs"][2]) <= 3.1415926535897936
    pi = 3.141592653589793
    got = [float(v) for v in report["sym_rem"]]
    # Manifolds.sym_rem: x ~ pi (isapprox, rtol sqrt(eps)) -> -pi; elsewhere the nearest-multiple remainder
    assert got[0] == -pi and got[1] == -pi and abs(got[2] - (pi - 1e-6)) < 1e-15 and abs(abs(got[3]) - pi) < 1e-15
    assert abs(abs(got[4]) - pi) < 1e-14


def test_seeded_sqrt(report):
    assert float(report["sqrt_seeded_rel"][0]) < 3.4e-16   # ~1.5 ulp with the seed anywhere inside MUFU.RSQ's error bound


def test_quaternion_exp_log(report):
    k, c = map(float, report["exp_scale_abs"])
    assert k < 2.3e-16 and c < 4.5e-16
    assert float(report["log_scale_rel"][0]) < 6e-16
    assert float(report["exp_log_roundtrip_abs"][0]) < 5e-15   # rotation vectors up to |w| = 3.1415, q and -q alike
    assert float(report["exp_general_abs"][0]) < 1e-14         # |w| in (pi, 2 pi): same rotation as the principal vector
    assert float(report["qmul_rotate_abs"][0]) < 2e-14


def test_closest_rotation_vector(report):
    err, _, moved = report["closest_rotvec_abs"]
    assert float(err) < 2e-14 and int(moved) > 10000   # the far-side half of the cases had to move
