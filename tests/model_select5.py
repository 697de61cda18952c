"""Executable statement of the robust-scale selection of the alignment kernel (csrc/select5.cuh) in numpy float32: the
same passes, the same rank arithmetic, the same acceptance tests -- without threads.  tests/test_model_select5.py compares it
with sorted arrays; the CUDA code is compared with the oracle through the C ABI (tests/test_gpu_parity.py, every tier forced).

Keys are FP32 residuals x 2^16.  k = n // 2 is the rank of the median, `need_pred` (the even rule) asks for element k - 1
too.  Every pass recounts the keys below its own range, so the target's presence in a list is decided by integers only.
"""
import numpy as np

F = np.float32
CAP = 256        # list entries a bracket pass may hold
PASSES = 12      # passes a phase may take before the bisection


def _avg(hi, lw, need_pred):
    return F(np.float32(0.5) * hi + np.float32(0.5) * lw) if need_pred else F(hi)


def rank_list(lst, idx, need_pred):
    """element idx (and idx - 1): the LARGEST value with at most idx entries strictly below it (s5_rank)"""
    lst = np.asarray(lst, F)
    below = (lst[None, :] < lst[:, None]).sum(1)
    hi = lst[below <= idx].max()
    lw = lst[below + 1 <= idx].max() if need_pred else hi
    return hi, lw


def bracket_pass(v, lo, W):
    """keys below lo, keys inside [lo, lo + W) -- by the sign / the unsigned bit pattern of u = v - lo"""
    u = (v - F(lo)).astype(F)
    bits = u.view(np.uint32)
    inside = bits < np.array(W, F).view(np.uint32)
    return int((bits >> 31).sum()), v[inside]


def count_pass(v, lo, W):
    """16 bins: inner bins 1..14 tile [lo, lo + W), bin 0 / 15 everything below / above (bin = rint(60 t) >> 2)"""
    sc = F(F(14.0) / (F(15.0) * F(W)))
    of = F(-(F(lo) - F(W) * F(1.0 / 16.0)) * sc)
    t = np.clip((v * sc + of).astype(F), F(0), F(1))          # (fma in the kernel: the bins differ by an ulp at most, and
    j = np.rint((t * F(60.0)).astype(F)).astype(np.int64)      #  nothing depends on where exactly a bin ends)
    return np.bincount(j >> 2, minlength=16)[:16]


def locate(T, lo, W, k, k_low):
    cum = np.cumsum(T[:15])
    reach_hi, reach_lo = np.nonzero(k < cum)[0], np.nonzero(k_low < cum)[0]
    b_hi = int(reach_hi[0]) if len(reach_hi) else 15
    b_lo = int(reach_lo[0]) if len(reach_lo) else 15
    c_lo = int(cum[b_lo - 1]) if b_lo > 0 else 0
    c_hi_end = int(cum[b_hi]) if b_hi < 15 else int(T.sum())
    m = c_hi_end - c_lo
    bw = F(W) * F(1.0 / 14.0)
    nlo, nhi = F(lo + F(b_lo - 1) * bw), F(lo + F(b_hi) * bw)
    if b_lo == 0:
        nlo = F(nhi - max(F(16.0) * F(W), F(4.0) * abs(nhi)))
    if b_hi == 15:
        nhi = F(nlo + max(F(16.0) * F(W), F(4.0) * abs(nlo)))
    mg = F(F(0.02) * bw + F(1.0e-6) * max(abs(nlo), abs(nhi)))
    bracket = b_lo > 0 and b_hi < 15 and m <= 192
    ties = (not bracket) and b_lo > 0 and b_hi < 15 and not (bw > F(1.0e-2))
    return F(nlo - mg), F((nhi - nlo) + F(2.0) * mg), bracket, ties


def bisection(v, k, need_pred):
    """the safety net: bitwise bisection over the ordered-integer image of the keys"""
    b = v.view(np.uint32).astype(np.uint64)
    o = np.where(b >> 31, b ^ 0xffffffff, b ^ 0x80000000).astype(np.uint64)
    T = np.uint64(0)
    for bit in range(31, -1, -1):
        cand = T | np.uint64(1 << bit)
        if int((o < cand).sum()) <= k:
            T = cand

    def un(x):
        x = int(x)
        x = x ^ 0x80000000 if x & 0x80000000 else (~x) & 0xffffffff
        return np.array(x, np.uint32).view(F)
    hi = un(T)
    less = o < T
    lw = un(o[less].max()) if need_pred and int(less.sum()) > k - 1 else hi
    return hi, lw


def select_phase(v, k, need_pred, lo, W, bracket, stats):
    """one target (median of v): the loop of passes of s5_sigma; returns (value, density around it)"""
    k_low = k - 1 if need_pred else k
    for _ in range(PASSES):
        if bracket:
            stats["bracket"] += 1
            C0, lst = bracket_pass(v, lo, W)
            inside = C0 <= k_low and k < C0 + len(lst)
            if inside and len(lst) <= CAP:
                hi, lw = rank_list(lst, k - C0, need_pred)
                return _avg(hi, lw, need_pred), F(max(len(lst), 1)) / F(W)
            bracket = False
            if not inside:
                lo, W = F(lo - F(6.5) * F(W)), F(F(W) * F(14.0))
        else:
            stats["count"] += 1
            T = count_pass(v, lo, W)
            lo, W, bracket, ties = locate(T, lo, W, k, k_low)
            if ties:
                break
    stats["bisection"] += 1
    hi, lw = bisection(v, k, need_pred)
    return _avg(hi, lw, need_pred), F(0)


def fused(x, k, need_pred, pm, pd, hm, hd, stats):
    """both targets from ONE pass (the fused pass of s5_sigma); returns (median or None, deviation or None, rhos)"""
    k_low = k - 1 if need_pred else k
    mA, wM, dA, wD0 = F(pm - hm), F(F(2) * hm), F(pd - hd), F(F(2) * hd)
    gLo, wD = F(dA - hm), F(wD0 + F(2) * hm)
    a = (x - mA).astype(F)
    g = (np.abs((a - hm).astype(F)) - gLo).astype(F)
    ab, gb = a.view(np.uint32), g.view(np.uint32)
    cM, cI = int((ab >> 31).sum()), int((gb >> 31).sum())
    inM, inD = ab < np.array(wM, F).view(np.uint32), gb < np.array(wD, F).view(np.uint32)
    push = inM | inD
    lst = x[push]
    stats["fused"] += 1
    if len(lst) > CAP:
        return None, None, None
    Lb, nM, nD = int((ab[push] >> 31).sum()), int(inM.sum()), int(inD.sum())
    if not (cM <= k_low and k < cM + nM):
        return None, None, None
    hi, lw = rank_list(lst, k - cM + Lb, need_pred)
    med = _avg(hi, lw, need_pred)
    rho0 = F(max(nM, 1)) / wM
    dev = np.where(inD[push], np.abs((lst - med).astype(F)), F(np.inf)).astype(F)
    if not (cI <= k_low and k < cI + nD):
        return med, None, (rho0, None)
    hi, lw = rank_list(dev, k - cI, need_pred)
    mgn = F(F(2) + F(1.0e-6) * (dA + wD0))
    if lw >= dA + mgn and hi < dA + wD0 - mgn:   # the proof: inner keys deviate by less than dA, outer ones by dB or more
        return med, _avg(hi, lw, need_pred), (rho0, F(max(nD, 1)) / wD)
    return med, None, (rho0, None)


def sigma(x, n_total, pred=None, force=0):
    """median and median absolute deviation of the float32 keys x as s5_sigma finds them.  pred: dict with v (2), moved (2),
    rho (2), have, have_move (updated in place).  Returns (median, mad, stats)."""
    x = np.asarray(x, F)
    n = len(x)
    k = n // 2
    need_pred = (n_total % 2 == 0) and k > 0
    stats = {"bracket": 0, "count": 0, "bisection": 0, "fused": 0}
    if pred is None:
        pred = {"v": [F(0), F(0)], "moved": [F(0), F(0)], "rho": [F(0), F(0)], "have": False, "have_move": False}
    res = [None, None]
    predicted = pred["have"] and pred["have_move"] and force == 0
    fused_tried = False
    if predicted and pred["rho"][0] > 0 and pred["rho"][1] > 0:
        hm = F(min(max(F(1.25) * pred["moved"][0], F(8) / pred["rho"][0], F(16)), F(1.0e7)))
        hd = F(min(max(F(1.25) * pred["moved"][1], F(8) / pred["rho"][1], F(16)), F(1.0e7)))
        if F(2) * hm * pred["rho"][0] + F(2) * (hd + hm) * pred["rho"][1] <= F(200):
            fused_tried = True
            med, mad, rhos = fused(x, k, need_pred, pred["v"][0], pred["v"][1], hm, hd, stats)
            res = [med, mad]
            if med is not None:
                pred["rho"][0] = rhos[0]
            if mad is not None:
                pred["rho"][1] = rhos[1]
    for phase in range(2):
        if res[phase] is not None:
            continue
        v = x if phase == 0 else np.abs((x - res[0]).astype(F))
        pv, pdv = pred["v"][phase], pred["v"][1]
        bracket = False
        if force == 2:
            hi, lw = bisection(v, k, need_pred)
            res[phase] = _avg(hi, lw, need_pred)
            stats["bisection"] += 1
            continue
        if predicted:
            rho = pred["rho"][phase]
            h = F(1.25) * pred["moved"][phase]
            if rho > 0:
                h = max(h, F(8) / rho)
            h = F(min(max(h, F(16)), F(1.0e7)))
            lo, W = F(pv - h), F(F(2) * h)
            bracket = rho > 0 and W * rho <= F(200) and not fused_tried
        elif pred["have"] and force != 1:
            s = max(pdv, F(1024))
            lo, W = (F(pv - s), F(F(2) * s)) if phase == 0 else (F(F(0.5) * s), F(F(1.5) * s))
        else:
            lo, W = (F(-1048576.0), F(2097152.0)) if phase == 0 else (F(0), F(2097152.0))
        res[phase], pred["rho"][phase] = select_phase(v, k, need_pred, lo, W, bracket, stats)
    if pred["have"]:
        pred["moved"] = [abs(F(res[0] - pred["v"][0])), abs(F(res[1] - pred["v"][1]))]
        pred["have_move"] = True
    pred["v"] = [F(res[0]), F(res[1])]
    pred["have"] = True
    return F(res[0]), F(res[1]), stats
