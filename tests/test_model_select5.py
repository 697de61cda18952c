"""The selection logic of the alignment kernel's robust scale (csrc/select5.cuh), pinned on the CPU: tests/model_select5.py
restates its passes in numpy float32; here every route through them -- predicted brackets (the fused pass), count passes,
the bisection safety net -- is compared with the sorted-array definition (MEDIAN_EXACT) on the distributions
that stress it: heavy ties, two clusters, a median far from zero, tiny and huge scales, predictions that miss."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import model_select5 as ms

F = np.float32


def reference(x, n_total):
    x = np.sort(np.asarray(x, F))
    n, k = len(x), len(x) // 2
    even = n_total % 2 == 0 and k > 0
    med = F(F(0.5) * x[k] + F(0.5) * x[k - 1]) if even else x[k]
    d = np.sort(np.abs((x - med).astype(F)))
    mad = F(F(0.5) * d[k] + F(0.5) * d[k - 1]) if even else d[k]
    return med, mad


def laplace(rng, n, scale, shift=0.0):
    return (rng.laplace(shift, scale, n) * 65536.0).astype(F)


CASES = {
    "laplace": lambda rng: laplace(rng, 12475, 5.0),
    "laplace_even": lambda rng: laplace(rng, 12500, 5.0),
    "shifted": lambda rng: laplace(rng, 9999, 3.0, 90.0),
    "negative": lambda rng: laplace(rng, 10000, 8.0, -120.0),
    "tiny_scale": lambda rng: laplace(rng, 5001, 1e-3),
    "huge_scale": lambda rng: laplace(rng, 5000, 60.0),
    "ties_zero": lambda rng: np.where(rng.random(8000) < 0.7, 0.0, laplace(rng, 8000, 4.0)).astype(F),
    "all_equal": lambda rng: np.full(3000, 77.0 * 65536.0, F),
    "two_values": lambda rng: np.where(rng.random(6001) < 0.5, -40.0 * 65536, 160.0 * 65536).astype(F),
    "two_clusters": lambda rng: np.concatenate([laplace(rng, 3000, 0.5, -30.0), laplace(rng, 3001, 0.5, 30.0)]),
    "quantised": lambda rng: (np.round(rng.laplace(0, 5, 7000)) * 65536.0).astype(F),
    "few": lambda rng: laplace(rng, 25, 5.0),
    "two": lambda rng: np.array([3.0, -1.0], F),
    "one": lambda rng: np.array([2.5], F),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("force", [0, 1, 2])
def test_cold_selection_is_exact(name, force):
    rng = np.random.default_rng(sum(map(ord, name)))
    x = CASES[name](rng)
    for n_total in (len(x), len(x) + 1):   # both median rules
        med, mad, st = ms.sigma(x, n_total, force=force)
        want = reference(x, n_total)
        assert (med, mad) == want, (name, force, n_total, med, mad, want, st)
        if force == 2:
            assert st["bisection"] == 2 and st["bracket"] == st["count"] == 0


def test_sequence_of_evaluations_uses_the_prediction_and_stays_exact():
    """A converging sequence (as the Gauss-Newton iterations of a level): the first evaluation counts, the later ones find
    both statistics in ONE fused pass; a jump (as at a level change without the reset) misses and recovers."""
    rng = np.random.default_rng(5)
    base = rng.laplace(0, 5, 12475)
    drift = rng.normal(0, 1, 12475)
    pred = None
    fused_hits = 0
    for it, eps in enumerate([1.0, 0.3, 0.05, 0.01, 0.002, 0.0004, 0.6, 0.59]):
        x = ((base * (1 + 0.2 * eps) + eps * drift) * 65536.0).astype(F)
        if pred is None:
            pred = {"v": [F(0), F(0)], "moved": [F(0), F(0)], "rho": [F(0), F(0)], "have": False, "have_move": False}
        med, mad, st = ms.sigma(x, len(x), pred)
        assert (med, mad) == reference(x, len(x)), (it, st)
        if it == 0:
            assert st["fused"] == 0 and st["count"] >= 2
        fused_hits += st["fused"] == 1 and st["bracket"] == st["count"] == 0
    assert fused_hits >= 2


@pytest.mark.parametrize("seed", range(6))
def test_wrong_predictions_never_give_wrong_answers(seed):
    """brackets that miss low / high, that are far too narrow or hold thousands of keys, a deviation bracket that swallows the
    median: the rank arithmetic sends every one of them to the count passes"""
    rng = np.random.default_rng(100 + seed)
    x = laplace(rng, 12000 + seed, 4.0, rng.uniform(-20, 20))
    med, mad = reference(x, len(x))
    for dm, dd, moved, rho in [(0.0, 0.0, 10.0, 0.01), (3e4, 0.0, 10.0, 0.01), (-3e4, 0.0, 10.0, 0.01), (0.0, 5e4, 10.0, 0.01),
                               (0.0, -5e4, 10.0, 0.01), (100.0, -100.0, 50.0, 0.0134), (0.0, 0.0, 2e5, 1e-5), (2e5, 2e5, 1.0, 0.02),
                               (0.0, -0.99 * float(mad), 1.0, 0.01)]:
        pred = {"v": [F(med + dm), F(mad + dd)], "moved": [F(moved), F(moved)], "rho": [F(rho), F(rho)], "have": True,
                "have_move": True}
        got = ms.sigma(x, len(x), pred)
        assert (got[0], got[1]) == (med, mad), (dm, dd, moved, rho, got)
