"""Compiles and runs tests/cpp/test_host_classes.cpp: the reference-shaped C++ host classes (ImagePyramid, Frame,
FeatureSelection, ImageAlignment, FeatureAlignment) over the C ABI, checked against the oracle on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semi-direct-visual-odometry_b200")


def _compile(tmp_path):
    exe = str(tmp_path / "test_host_classes")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(PKG, "host"), "-I", os.path.join(ROOT, "oracle"),
           os.path.join(ROOT, "tests", "cpp", "test_host_classes.cpp"), "-o", exe, "-L", PKG, "-lsvo_b200",
           "-L", os.path.join(ROOT, "oracle"), "-lsvo_oracle", "-Wl,-rpath," + PKG, "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-pthread"]
    subprocess.check_call(cmd)
    return exe


def test_host_classes_compile(tmp_path, orc):
    """CPU: the host headers are valid C++17 against the C ABI and link against libsvo_b200.so."""
    _compile(tmp_path)


@pytest.mark.gpu
def test_host_classes_on_gpu(tmp_path, orc, pair_cache):
    exe = _compile(tmp_path)
    pair = pair_cache(4, 500)
    pair["ref"].tofile(tmp_path / "ref.u8")
    pair["cur"].tofile(tmp_path / "cur.u8")
    args = [exe, str(tmp_path / "ref.u8"), str(tmp_path / "cur.u8"), str(pair["w"]), str(pair["h"])]
    args += ["%.17g" % v for v in pair["T_cur_true"]]
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout
    assert "ALL HOST-CLASS CHECKS PASSED" in r.stdout
