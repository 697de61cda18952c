"""The bounds arithmetic of the alignment kernel's combined median + MAD selection (csrc/select4.cuh), pinned on the CPU:
tests/model_select4.py is the executable statement the CUDA code transcribes; here it is compared with the sorted-array
definition of the reference's median rule (src/algorithm.cpp:834-872, MEDIAN_EXACT) on adversarial key distributions,
window predictions that are off by a fraction of a window, even and odd row counts, ties, tiny and huge scales."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import model_select4 as ms

BIAS = 1 << 25


def _keys(rng, kind, n):
    if kind == "gauss":
        r = rng.normal(rng.uniform(-0.4, 0.4), rng.uniform(0.05, 12), n)
    elif kind == "heavy":
        r = rng.standard_t(1.5, n) * rng.uniform(0.1, 5) + rng.uniform(-0.3, 0.3)
    elif kind == "ties":
        r = rng.integers(-3, 4, n).astype(float) * rng.choice([1.0, 0.25, 1 / 65536])
    elif kind == "zeros":
        r = np.where(rng.random(n) < 0.7, 0.0, rng.normal(0, 5, n))
    elif kind == "bimodal":
        r = np.where(rng.random(n) < 0.5, rng.normal(-8, 0.5, n), rng.normal(9, 0.7, n))
    else:
        r = rng.normal(0, 3e-4, n)
    return np.rint(np.clip(r, -255, 255) * 65536).astype(np.int64) + BIAS


KINDS = ["gauss", "heavy", "ties", "zeros", "bimodal", "tiny"]


@pytest.mark.parametrize("kind", KINDS)
def test_one_round_plus_lists_is_exact_whenever_it_answers(kind):
    rng = np.random.default_rng(KINDS.index(kind))
    hits = 0
    for _ in range(400):
        NB = int(rng.choice([64, 128, 512]))
        nfeat, area = int(rng.integers(1, NB + 1)), int(rng.choice([16, 25]))
        n = nfeat * area
        keys = _keys(rng, kind, int(rng.integers(1, nfeat + 1)) * area)
        k = len(keys) // 2
        need = (n % 2 == 0) and k > 0
        ref = ms.reference(keys, k, need)
        s = int(rng.integers(0, 12))
        W = NB << s
        m0 = ref[0] + int(rng.normal(0, W / 6))
        d0 = max(0, ref[2] // 2 + int(rng.normal(0, W / 6)))
        out = ms.select(keys, n, ms.Win.predicted(m0, d0, d0, s, NB, NB), capM=10**9, capD=10**9)
        if out != ms.MISS:
            hits += 1
            assert out == ref, (kind, NB, s, out, ref)
    assert hits > 100  # the windows do catch the targets most of the time


@pytest.mark.parametrize("kind", ["gauss", "heavy", "bimodal", "ties"])
def test_coarse_round_then_refined_round_is_exact(kind):
    rng = np.random.default_rng(10 + KINDS.index(kind))
    hits = 0
    for _ in range(300):
        NB = int(rng.choice([64, 128, 512]))
        nfeat = int(rng.integers(max(1, NB // 4), NB + 1))
        n = nfeat * 25
        keys = _keys(rng, kind, int(rng.integers(max(1, nfeat // 2), nfeat + 1)) * 25)
        k = len(keys) // 2
        need = (n % 2 == 0) and k > 0
        ref = ms.reference(keys, k, need)
        dprev = int((ref[2] // 2) * rng.uniform(0.62, 2.0))
        dlo, dhi = int(0.45 * dprev), int(1.75 * dprev)
        sA = max(0, int(np.ceil(np.log2(max(1, (dhi - dlo) / 22)))))
        m0 = ref[0] + int(rng.normal(0, 0.15 * 65536))
        out = ms.select_cold(keys, n, m0, dlo, dhi, sA, NB, capM=10**9, capD=10**9)
        if out != ms.MISS:
            hits += 1
            assert out == ref, (kind, NB, out, ref)
    assert hits > (100 if kind != "ties" else 30)


def test_list_overflow_is_reported_not_guessed():
    keys = np.full(2000, BIAS + 7, dtype=np.int64)  # 2,000 equal residuals: every candidate list overflows
    out = ms.select(keys, 2000, ms.Win.predicted(BIAS, 0, 0, 3, 512, 512))
    assert out == ms.OVERFLOW
