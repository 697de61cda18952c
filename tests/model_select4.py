"""Executable model (numpy, integer keys) of the combined median + MAD selection of the alignment kernel
(csrc/select4.cuh).  The CUDA code is a transcription of `locate` / `lists` below; tests/test_model_select4.py checks
the model against the sorted-array definitions (src/algorithm.cpp:834-872, MEDIAN_EXACT) on adversarial
distributions, so the bounds arithmetic is pinned on the CPU before it runs on a GPU.

Keys are unsigned fixed-point residuals (rint(r 2^16) + 2^25).  k = numValid / 2; the statistic is the mean of the
elements k-1 and k when the TOTAL row count N is even, else element k.  Deviations are kept doubled:
|2 key - med2| with med2 = kHi + kLo.

A ROUND histograms the keys over three windows of one global grid (bin = key >> s): M around the predicted median,
L and R around median -/+ deviation.  From the three histograms and the counts below each window `locate` derives
  * the bin(s) [bMin, bMax] of the median (and of its predecessor when the even rule needs it),
  * j0 < j1 with  j0 2^s < d* <= j1 2^s  for the k-th smallest deviation d* (and its predecessor), by counting the
    keys that MUST / CAN lie within d of any median in those bins,
  * the candidate bins L' = [bMin - j1, bMax - j0], R' = [bMin + j0, bMax + j1] and `base`, the number of keys that
    are closer to the median than every candidate.
`lists` then ranks the few keys of those bins exactly.  A round can follow a round (windows = the previous candidate
bins at a smaller s); windows come from a prediction (previous evaluation) or from a coarse round.
"""
import numpy as np

MISS, OVERFLOW = "miss", "overflow"


class Win:
    """three windows on the grid of bins of width 2^s: starts gM, gL, gR and sizes nbM, nbL, nbR (bins)"""

    def __init__(self, s, gM, nbM, gL, nbL, gR, nbR):
        self.s, self.gM, self.nbM, self.gL, self.nbL, self.gR, self.nbR = s, gM, nbM, gL, nbL, gR, nbR
        assert gL + nbL <= gM and gM + nbM <= gR, "windows are ordered and disjoint"
        self.contig = (gL + nbL == gM) and (gM + nbM == gR)

    @staticmethod
    def predicted(m0, dlo, dhi, s, nbM, nbLR):
        """M centred on m0; R covers m0 + [dlo, dhi] (centred when narrower than the window), L mirrors it"""
        gM = (m0 >> s) - nbM // 2
        c = (dlo + dhi) // 2
        gR = ((m0 + c) >> s) - nbLR // 2
        gL = ((m0 - c) >> s) - (nbLR - 1) // 2
        if gR < gM + nbM or gL + nbLR > gM:  # near: contiguous triple
            gR, gL = gM + nbM, gM - nbLR
        return Win(s, gM, nbM, gL, nbLR, gR, nbLR)


def histograms(keys, w):
    gb = keys >> w.s
    out = []
    for g, nb in ((w.gM, w.nbM), (w.gL, w.nbL), (w.gR, w.nbR)):
        t = gb - g
        below = int((t < 0).sum())
        h = np.bincount(t[(t >= 0) & (t < nb)], minlength=nb)
        out.append((below, np.concatenate([[0], np.cumsum(h)]), h))
    return out


def first_true(lo, hi, pred):
    """smallest j in [lo, hi] with pred(j) (monotone), hi + 1 if none -- the kernel does this 32 probes at a time"""
    for j in range(lo, hi + 1):
        if pred(j):
            return j
    return hi + 1


def locate(hists, k, need_pred, w):
    (cM, PM, hM), (cL, PL, hL), (cR, PR, hR) = hists

    def Cf(b):  # keys with grid bin < b; b must be an edge of the window the rule below picks
        if b >= w.gR:
            i = b - w.gR
            assert 0 <= i <= w.nbR
            return cR + int(PR[i])
        if b >= w.gM:
            i = b - w.gM
            assert 0 <= i <= w.nbM
            return cM + int(PM[i])
        i = b - w.gL
        assert 0 <= i <= w.nbL
        return cL + int(PL[i])

    # ---- median ----
    kM = k - cM
    if kM < 0 or kM >= int(PM[w.nbM]):
        return MISS
    iM = int(np.searchsorted(PM, kM, side="right")) - 1
    rM = kM - int(PM[iM])
    bMax = w.gM + iM
    bMin = bMax
    if need_pred and rM == 0:
        nz = np.nonzero(hM[:iM])[0]
        if len(nz) == 0:
            return MISS  # the predecessor lies below the window
        bMin = w.gM + int(nz[-1])
    # ---- deviation bracket ----
    # Glo(j) = keys that lie within j 2^s of EVERY median in [bMin, bMax]; Ghi(j) = keys that CAN lie within
    Glo = lambda j: Cf(bMin + j) - Cf(bMax - j + 1)
    Ghi = lambda j: Cf(bMax + j + 1) - Cf(bMin - j)
    hi1 = min(w.gR + w.nbR - bMin, bMax + 1 - w.gL)
    hi0 = min(w.gR + w.nbR - bMax - 1, bMin - w.gL)
    if w.contig:
        lo1, lo0 = 1, 0
    else:
        lo1 = max(w.gR - bMin, bMax + 1 - w.gL - w.nbL)
        lo0 = max(w.gR - bMax - 1, bMin - w.gL - w.nbL)
        lo1, lo0 = max(lo1, 1), max(lo0, 0)
    kk = k if need_pred else k + 1
    # The kernel probes with one thread per bin: thread t of window X takes the UPPER edge gX + t + 1 of its bin as the
    # upper argument of Glo / Ghi (only R's edges without contiguity).  The lowest edge of a window is therefore no
    # probe; an unprobed lo1 only loosens j1, the unprobed lo0 is checked explicitly.
    edges = []
    for X, (g, nb) in enumerate(((w.gM, w.nbM), (w.gL, w.nbL), (w.gR, w.nbR))):
        if X != 2 and not w.contig:
            continue
        edges += [g + t + 1 for t in range(nb)]
    c1 = [b - bMin for b in edges if lo1 <= b - bMin <= hi1 and Glo(b - bMin) >= k + 1]
    if not c1:
        return MISS
    j1 = min(c1)
    c0 = [b - bMax - 1 for b in edges if lo0 <= b - bMax - 1 <= hi0 and Ghi(b - bMax - 1) >= kk]
    jz = min(c0) if c0 else hi0 + 1
    if lo0 <= hi0 and jz > lo0 and Ghi(lo0) >= kk:
        jz = lo0
    jz = min(jz, min(hi0, j1) + 1)
    j0 = jz - 1
    if j0 < lo0:
        if not w.contig:
            return MISS  # the lower bound lies below the windows
        j0 = -1
    base = 0
    if j0 >= 0:
        if j0 < lo1 and not (w.contig and j0 >= 0):
            return MISS
        base = max(0, Glo(j0))
    Ll, Lh, Rl, Rh = bMin - j1, bMax - j0, bMin + j0, bMax + j1
    if w.contig:
        if Ll < w.gL or Rh >= w.gR + w.nbR:
            return MISS
    elif Ll < w.gL or Lh >= w.gL + w.nbL or Rl < w.gR or Rh >= w.gR + w.nbR:
        return MISS
    return dict(bMin=bMin, bMax=bMax, hMb=int(hM[iM]), rM=rM, j0=j0, j1=j1, base=base, Ll=Ll, Lh=Lh, Rl=Rl, Rh=Rh)


def lists(keys, k, need_pred, s, r1, capM=32, capD=64, bias=1 << 25):
    gb = keys >> s
    inM = (gb == r1["bMax"]) | (gb == r1["bMin"])
    listM = np.sort(keys[inM])
    if len(listM) > capM:
        return OVERFLOW
    idx = len(listM) - r1["hMb"] + r1["rM"]
    kHi = int(listM[idx])
    kLo = int(listM[idx - 1]) if need_pred else kHi
    med2 = (kHi - bias) + (kLo - bias)
    inD = ((gb >= r1["Ll"]) & (gb <= r1["Lh"])) | ((gb >= r1["Rl"]) & (gb <= r1["Rh"]))
    dev = np.sort(np.abs(2 * (keys[inD] - bias) - med2))
    if len(dev) > capD:
        return OVERFLOW
    tD = k - r1["base"]
    if not (0 <= tD < len(dev)) or (need_pred and tD < 1):
        raise AssertionError("bounds arithmetic broken: target %d of %d candidates" % (tD, len(dev)))
    dHi = int(dev[tD])
    dLo = int(dev[tD - 1]) if need_pred else dHi
    return kHi, kLo, dHi, dLo


def reference(keys, k, need_pred, bias=1 << 25):
    """sorted-array definition"""
    srt = np.sort(keys)
    kHi = int(srt[k])
    kLo = int(srt[k - 1]) if need_pred else kHi
    med2 = (kHi - bias) + (kLo - bias)
    dev = np.sort(np.abs(2 * (keys - bias) - med2))
    dHi = int(dev[k])
    dLo = int(dev[k - 1]) if need_pred else dHi
    return kHi, kLo, dHi, dLo


def refine_windows(r1, s, NB, margin=4):
    """windows of the next round: the candidate bins of this round plus `margin` bins on either side (the bounds of the
    finer round probe up to two bins beyond the targets), on the finest grid s2 <= s where every window has at most NB
    bins (None if not even s does)."""
    for s2 in range(0, s + 1):
        f = s - s2
        mlo, mhi = (r1["bMin"] << f) - margin, ((r1["bMax"] + 1) << f) + margin
        llo, lhi = (r1["Ll"] << f) - margin, ((r1["Lh"] + 1) << f) + margin
        rlo, rhi = (r1["Rl"] << f) - margin, ((r1["Rh"] + 1) << f) + margin
        if mhi - mlo > NB or lhi - llo > NB or rhi - rlo > NB:
            continue
        if lhi <= mlo and mhi <= rlo:  # separate windows
            return Win(s2, mlo, mhi - mlo, llo, lhi - llo, rlo, rhi - rlo)
        # they touch or overlap: a contiguous triple around the M window
        nbL, nbR = max(mlo - min(llo, mlo - 1), 1), max(max(rhi, mhi + 1) - mhi, 1)
        if nbL <= NB and nbR <= NB:
            return Win(s2, mlo, mhi - mlo, mlo - nbL, nbL, mhi, nbR)
    return None


def select(keys, n_total, w, capM=32, capD=64):
    """one round + lists.  keys: visible keys; n_total: row count N (its parity picks the even rule)"""
    k = len(keys) // 2
    need_pred = (n_total % 2 == 0) and k > 0
    r1 = locate(histograms(keys, w), k, need_pred, w)
    if r1 == MISS:
        return MISS
    return lists(keys, k, need_pred, w.s, r1, capM, capD)


def select_cold(keys, n_total, m0, dlo, dhi, sA, NB, capM=32, capD=64, stats=None):
    """coarse round with 16 + 24 + 24 bins (the kernel counts them with thread-private counters), then one round of NB
    bins per window over the candidate bins, then lists"""
    k = len(keys) // 2
    need_pred = (n_total % 2 == 0) and k > 0
    w0 = Win.predicted(m0, dlo, dhi, sA, 16, 24)
    r0 = locate(histograms(keys, w0), k, need_pred, w0)
    if r0 == MISS:
        return MISS
    w1 = refine_windows(r0, sA, NB)
    if w1 is None:
        return MISS
    hs = histograms(keys, w1)
    if stats is not None:
        stats.append(sum(int(h[1][-1]) for h in hs))
    r1 = locate(hs, k, need_pred, w1)
    if r1 == MISS:
        raise AssertionError("a refinement round cannot miss")
    return lists(keys, k, need_pred, w1.s, r1, capM, capD)
