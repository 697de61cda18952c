"""The single-process multi-GPU layer of libsvo_b200.so (svo_multi_*, csrc/multi.cu): a C++ driver (tests/cpp/test_multi.cpp)
and the ctypes binding, both against the same batch on one context.  On a one-GPU box the layer runs with one device and
with the same device twice (two contexts, two worker threads: the sharding and the gather are exercised all the same)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semi-direct-visual-odometry_b200")


def _compile(tmp_path):
    exe = str(tmp_path / "test_multi")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_multi.cpp"), "-o", exe,
                           "-L", PKG, "-lsvo_b200", "-Wl,-rpath," + PKG, "-pthread"])
    return exe


def test_multi_driver_compiles(tmp_path, pkg):
    _compile(tmp_path)


def _batch(pkg, n, nfeat):
    b = pkg.synth.make_batch(n, nfeat, workers=1)
    jobs = pkg.capi.make_jobs(n)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    jobs["ref_slot"], jobs["kf_slot"], jobs["cur_slot"] = np.arange(n), np.arange(n), np.arange(n) + n
    jobs["n_ref"], jobs["n_kf"], jobs["feat_offset"] = b["n_feat"], 0, b["feat_offset"]
    jobs["T_ref"], jobs["T_kf"], jobs["T_cur"] = ident, ident, ident
    return b, jobs


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
def test_multi_cpp_driver(tmp_path, pkg):
    exe = _compile(tmp_path)
    n = 7
    b, jobs = _batch(pkg, n, 200)
    b["ref"].tofile(tmp_path / "ref.u8")
    b["cur"].tofile(tmp_path / "cur.u8")
    jobs.tofile(tmp_path / "jobs.bin")
    b["feats"].tofile(tmp_path / "feats.bin")
    for d in sorted({1, _gpus()}):
        r = subprocess.run([exe, str(tmp_path), str(n), str(b["w"]), str(b["h"]), str(d)], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        print(r.stdout)
        assert r.returncode == 0 and "MULTI-GPU LAYER OK" in r.stdout, r.stdout


@pytest.mark.gpu
def test_multi_binding_matches_one_context(pkg):
    n = 9
    b, jobs = _batch(pkg, n, 150)
    with pkg.Context(b["w"], b["h"], b["K"], levels=4, max_frames=2 * n, max_jobs=n, max_features=512, max_fa_items=16) as ctx:
        ctx.upload(0, b["ref"])
        ctx.upload(n, b["cur"])
        want, wstats = ctx.sparse_align(jobs, b["feats"], mode=pkg.capi.GN, max_iter=30)
    ng = _gpus()
    for devices in ([0], [0, 0], list(range(ng)) if ng > 1 else [0, 0, 0]):   # the same device several times is legal
        D = len(devices)
        per = max(pkg.capi.shard(n, D, i)[1] - pkg.capi.shard(n, D, i)[0] for i in range(D))
        with pkg.MultiContext(b["w"], b["h"], b["K"], D, devices=devices, max_frames=2 * per, max_jobs=per, max_features=512) as m:
            m.upload(0, b["ref"])
            m.upload(per, b["cur"], prefetch=True)
            mj = jobs.copy()
            for i in range(D):
                lo, hi = pkg.capi.shard(n, D, i)
                mj["ref_slot"][lo:hi] = mj["kf_slot"][lo:hi] = np.arange(hi - lo)
                mj["cur_slot"][lo:hi] = per + np.arange(hi - lo)
            got, gstats = m.sparse_align(mj, b["feats"], mode=pkg.capi.GN, max_iter=30)
            assert got.tobytes() == want.tobytes(), devices
            assert gstats.tobytes() == wstats.tobytes(), devices
            m.stage(mj, b["feats"], mode=pkg.capi.GN, max_iter=30)
            ms = m.time_launches(1, 2)
            assert ms > 0 and m.fetch().tobytes() == want.tobytes()
