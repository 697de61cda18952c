"""CPU-side checks of the boundary: libsvo_b200.so loads, exports every symbol include/svo_b200.h declares, the
ctypes mirrors have the C layouts, and without a CUDA device the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "svo_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    L = pkg.load()
    decl = _declared()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(L, name), "libsvo_b200.so does not export %s" % name
    assert sorted(pkg.capi.SYMBOLS) == decl
    assert b"sm_100a" in L.svo_version()


def test_struct_layouts_match_header(pkg, tmp_path):
    """sizeof/offsetof from a C compile of the header == the numpy/ctypes mirrors."""
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "svo_b200.h"\nint main(void){'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(svo_config), sizeof(svo_feature_px),'
                    'sizeof(svo_align_feature), sizeof(svo_align_job), sizeof(svo_align_params), sizeof(svo_align_result),'
                    'sizeof(svo_align_level_stats), sizeof(svo_fa_item), sizeof(svo_fa_params), sizeof(svo_fa_result),'
                    'offsetof(svo_align_level_stats, pose_after), sizeof(svo_frontend_params), sizeof(svo_frontend_result),'
                    'sizeof(svo_epi_item), sizeof(svo_epi_params), sizeof(svo_epi_result), offsetof(svo_epi_item, depth));'
                    'printf("%zu %zu\\n", sizeof(svo_reproj_candidate), sizeof(svo_reproj_match));'
                    'printf("%zu %zu\\n", sizeof(svo_klt_params), offsetof(svo_klt_params, epsilon)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    c = pkg.capi
    want = [C.sizeof(c.Config), c.FEATURE_PX_DTYPE.itemsize, c.ALIGN_FEATURE_DTYPE.itemsize, c.ALIGN_JOB_DTYPE.itemsize,
            C.sizeof(c.AlignParams), c.ALIGN_RESULT_DTYPE.itemsize, c.ALIGN_STATS_DTYPE.itemsize, c.FA_ITEM_DTYPE.itemsize,
            C.sizeof(c.FaParams), c.FA_RESULT_DTYPE.itemsize, c.ALIGN_STATS_DTYPE.fields["pose_after"][1],
            C.sizeof(c.FrontendParams), c.FRONTEND_RESULT_DTYPE.itemsize, c.EPI_ITEM_DTYPE.itemsize, C.sizeof(c.EpiParams),
            c.EPI_RESULT_DTYPE.itemsize, c.EPI_ITEM_DTYPE.fields["depth"][1], c.REPROJ_CAND_DTYPE.itemsize,
            c.REPROJ_MATCH_DTYPE.itemsize, C.sizeof(c.KltParams), c.KltParams.epsilon.offset]
    assert got == want


def test_oracle_feature_layout_is_shared(pkg, orc):
    assert pkg.capi.ALIGN_FEATURE_DTYPE == orc.FEATURE_DTYPE == pkg.synth.FEATURE_DTYPE


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_device_is_a_loud_error(pkg):
    with pytest.raises(pkg.SvoError) as e:
        pkg.Context(1241, 376, (721.5, 721.5, 609.5, 172.8))
    assert e.value.code == pkg.capi.ERR_NO_DEVICE


def test_invalid_config_rejected(pkg):
    L = pkg.load()
    cfg = pkg.capi.Config(0, 4, 4, 4, 1, 1, 1, 1, 0, None, (C.c_double * 4)(1, 1, 1, 1))
    h = C.c_void_p()
    assert L.svo_create(C.byref(cfg), C.byref(h)) == pkg.capi.ERR_INVALID and not h.value
    assert L.svo_create(None, C.byref(h)) == pkg.capi.ERR_INVALID
    assert L.svo_sync(None) == pkg.capi.ERR_INVALID and L.svo_launch_count(None) == 0


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (only tests, smoke and bench's cpu legs may)."""
    pk = os.path.join(ROOT, "semi-direct-visual-odometry_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                assert "svo_oracle" not in txt and "import oracle" not in txt and "orc_" not in txt, os.path.join(dp, f)
    out = subprocess.run(["ldd", os.path.join(pk, "libsvo_b200.so")], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out


def test_synthetic_pair_is_deterministic(pkg):
    a = pkg.synth.make_pair(index=5, n_features=50)
    b = pkg.synth.make_pair(index=5, n_features=50)
    assert np.array_equal(a["ref"], b["ref"]) and np.array_equal(a["cur"], b["cur"])
    assert np.array_equal(a["feats"], b["feats"]) and a["ref"].shape == (376, 1241)
    n = np.linalg.norm(a["feats"]["bearing"], axis=1)
    assert np.allclose(n, 1.0)  # PinholeCamera::inverseProject2d returns unit vectors (src/pinhole_camera.cpp:100)
