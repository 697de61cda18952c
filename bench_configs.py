#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configurations other than the headline batch (bench.py measures that one):

  config 0  one 1241x376 frame pair, 4 levels, 500 features: ImageAlignment::align, faithful / iterated LM / GN, patch
            5x5 and the 4x4 extension: device-only us (pyramids resident) and end-to-end us (new frame H2D from pinned
            memory + pyramids + alignment + pose D2H), against the oracle on ONE host thread (the reference's path is
            single-threaded)
  config 1  FeatureAlignment batch: 2,000 patches, 7x7 (reference) and 8x8, identity and affine-warped templates,
            one step (what the reference executes) and <= 30 iterations
  config 2  the whole per-frame front end as one CUDA graph launch (svo_frontend_run)
  config 3  1,024 pairs x 1,000 features (bench.py --features 1000)
  config 5  (the first "next" component) depth-filter epipolar search: 2,000 seeds, 7x7 patches
  config 8   computeOpticalFlowSparse's tracker: cv::calcOpticalFlowPyrLK, 500 / 2,000 points, 11x11, 4 levels
  config 7   Map::reprojectMap: 1,000 candidates -> per-cell choice -> <= 151 feature alignments
  config 6   FeatureSelection::gradientMagnitudeWithSSC on one frame, 250 / 1,000 candidates

Prints one JSON line per measurement; every line carries `roofline` (algorithmic bytes of SURVEY 8d / device time
against the measured HBM peak) and `cpu_baseline` (the oracle port, bounded sample).  Needs a GPU.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def roof(bytes_, us):
    peak, src = peaks()
    ach = bytes_ / (us * 1e-6) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src,
            "algorithmic_bytes": bytes_, "traffic": None, "note": "latency-bound: one small problem cannot fill the device"}


def ev_time(torch, stream, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def wall_time(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="0,1,2,3,5,6,7,8,9")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    want = set(int(x) for x in a.configs.split(","))
    import torch
    import oracle as orc
    orc.build()
    pkg = importlib.import_module("semi-direct-visual-odometry_b200")
    capi, synth = pkg.capi, pkg.synth
    pkg.load()
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    lines = []

    def emit(d):
        d.setdefault("data", "synthetic")
        d.setdefault("n_gpus", 1)
        lines.append(d)
        print(json.dumps(d), flush=True)

    pair = synth.make_pair(0, 500)
    w, h, K = pair["w"], pair["h"], pair["K"]
    n = len(pair["feats"])
    with torch.cuda.stream(stream):
        ctx = pkg.Context(w, h, K, levels=4, max_frames=4, max_jobs=1, max_features=512, max_fa_items=4096,
                          stream=stream.cuda_stream)
        pin = ctx.pinned(2 * h * w)
        fr = pin.array.reshape(2, h, w)
        fr[0], fr[1] = pair["ref"], pair["cur"]
        ctx.upload(0, fr)
        job = capi.make_jobs(1)
        job[0]["ref_slot"], job[0]["kf_slot"], job[0]["cur_slot"] = 0, 0, 1
        job[0]["n_ref"], job[0]["n_kf"] = pair["n_ref"], 0
        job[0]["T_ref"], job[0]["T_kf"], job[0]["T_cur"] = pair["T_ref"], pair["T_kf"], pair["T_cur_init"]
        pyr = tuple(orc.build_pyramid(pair[k], 4)[0] for k in ("ref", "kf", "cur"))

        # ---------------- config 0 ----------------
        if 0 in want:
            for patch in (5, 4):
                for mname, mode, omode in (("faithful", capi.LM_FAITHFUL, orc.LM_FAITHFUL), ("lm_iterated", capi.LM_ITERATED, orc.LM_ITERATED),
                                           ("gn", capi.GN, orc.GN)):
                    kw = dict(patch_size=patch, min_level=0, max_level=3, mode=mode, max_iter=30)
                    res, st = ctx.sparse_align(job, pair["feats"], want_stats=True, **kw)
                    ctx.sparse_align_stage(job, pair["feats"], **kw)
                    ctx.sparse_align_h2d()
                    for _ in range(5):
                        ctx.sparse_align_launch()
                    dev_us = ev_time(torch, stream, ctx.sparse_align_launch, 50)

                    def e2e():
                        ctx.upload(1, fr[1])
                        return ctx.sparse_align(job, pair["feats"], want_stats=False, **kw)

                    for _ in range(3):
                        e2e()
                    e2e_us = wall_time(e2e, 30)
                    nvis = st[0]["n_px"].astype(np.float64) / (patch * patch)
                    evs = st[0]["evaluations"].astype(np.float64)
                    bytes_ = float((nvis * ((patch + 3) ** 2 + evs * (patch + 1) ** 2)).sum() + 32.0 * n + 256.0 * 4)
                    ts = []
                    for _ in range(10):
                        t0 = time.perf_counter()
                        rmse, T, status, lv = orc.sparse_align(*pyr, w, h, pair["feats"], pair["n_ref"], 0, pair["T_ref"], pair["T_kf"],
                                                               K, pair["T_cur_init"], patch_size=patch, mode=omode, max_iter=30)
                        ts.append((time.perf_counter() - t0) * 1e6)
                    cpu_us = float(np.median(ts))
                    emit({"config": {"workload": "config 0: one 1241x376 frame pair, 4 levels, %d features, patch %dx%d, %s <= 30 "
                                                 "iterations/level" % (n, patch, patch, mname)},
                          "metric": "us_per_frame_pair_sparse_align", "unit": "us", "higher_is_better": False,
                          "value": dev_us, "dtype": "f32/f64 mixed",
                          "evaluations": int(res[0]["evaluations"]), "status": int(res[0]["status"]),
                          "pose_diff_vs_oracle": float(np.abs(res[0]["T_cur"] - T).max()),
                          "e2e": {"value": e2e_us, "unit": "us", "h2d_bytes_per_step": int(h * w + job.nbytes + pair["feats"].nbytes),
                                  "d2h_bytes_per_step": int(capi.ALIGN_RESULT_DTYPE.itemsize),
                                  "what": "svo_frames_upload(new frame, pinned) + svo_sparse_align, host wall clock, median of 30"},
                          "roofline": roof(bytes_, dev_us),
                          "cpu_baseline": {"value": cpu_us, "unit": "us", "cores": 1, "kind": "port",
                                           "sample": "the same pair, oracle port, median of 10"}})

        # ---------------- config 1 ----------------
        if 1 in want:
            rng = np.random.default_rng(123)
            nfa = 2000
            idx = np.arange(nfa) % n
            R, t = synth.se3_Rt(pair["T_cur_true"])
            pc = pair["feats"]["point"][idx] @ R.T + t
            uv = np.stack([K[0] * pc[:, 0] / pc[:, 2] + K[2], K[1] * pc[:, 1] / pc[:, 2] + K[3]], 1)
            start = uv + rng.uniform(-2, 2, (nfa, 2))
            ref_g, cur_g = ctx.download(0, 0, 1), ctx.download(1, 0, 1)
            for patch in (7, 8):
                for affine in (False, True):
                    for mname, mode, omode, cap in (("faithful (1 step)", capi.LM_FAITHFUL, orc.LM_FAITHFUL, 1),
                                                    ("gn", capi.GN, orc.GN, 30)):
                        items = np.zeros(nfa, capi.FA_ITEM_DTYPE)
                        items["ref_slot"], items["cur_slot"] = 0, 1
                        items["ref_px"], items["px"] = pair["feats"]["px"][idx], start
                        items["A"] = (1, 0, 0, 1)
                        if affine:
                            items["A"] = np.array([1, 0, 0, 1]) + rng.uniform(-0.05, 0.05, (nfa, 4))
                            items["use_affine"] = 1
                        r = ctx.feature_align(items, patch_size=patch, mode=mode, max_iter=30)
                        ctx.feature_align_stage(items, patch_size=patch, mode=mode, max_iter=30)
                        ctx.feature_align_h2d()
                        for _ in range(3):
                            ctx.feature_align_launch()
                        dev_us = ev_time(torch, stream, ctx.feature_align_launch, 30)
                        e2e_us = wall_time(lambda: ctx.feature_align(items, patch_size=patch, mode=mode, max_iter=30), 20)
                        its = np.maximum(1, r["iterations"]).astype(np.float64)
                        bytes_ = float(((patch + 3) ** 2 + (its + 1) * (patch + 1) ** 2 + 48).sum())
                        ns = 200
                        t0 = time.perf_counter()
                        worst = 0.0
                        for i in range(ns):
                            rm, px, stt, it = orc.feature_align(ref_g, cur_g, items["ref_px"][i], items["px"][i],
                                                                A=items["A"][i] if affine else None, patch_size=patch, mode=omode,
                                                                max_iter=30)
                            if np.isfinite(px).all() and np.isfinite(r["px"][i]).all():
                                worst = max(worst, float(np.abs(px - r["px"][i]).max()))
                        cpu_us = (time.perf_counter() - t0) * 1e6 / ns * nfa
                        emit({"config": {"workload": "config 1: FeatureAlignment batch, %d patches %dx%d on gradient level 0, %s template, "
                                                     "%s" % (nfa, patch, patch, "affine-warped" if affine else "identity", mname)},
                              "metric": "us_per_batch_feature_align", "unit": "us", "higher_is_better": False, "value": dev_us,
                              "dtype": "f64 (float bilinear taps)", "features_per_sec": nfa / (dev_us * 1e-6),
                              "mean_iterations": float(its.mean()), "max_px_diff_vs_oracle_sample": worst,
                              "e2e": {"value": e2e_us, "unit": "us", "h2d_bytes_per_step": int(items.nbytes),
                                      "d2h_bytes_per_step": int(nfa * capi.FA_RESULT_DTYPE.itemsize),
                                      "what": "svo_feature_align (items H2D, kernel, results D2H), host wall clock, median of 20"},
                              "roofline": roof(bytes_, dev_us),
                              "cpu_baseline": {"value": cpu_us, "unit": "us", "cores": 1, "kind": "port",
                                               "sample": "first %d of the %d items through the oracle port, scaled to the batch" % (ns, nfa)}})

        # ---------------- config 2 ----------------
        if 2 in want:
            for mname, mode in (("faithful", capi.LM_FAITHFUL), ("gn", capi.GN)):
                kw = dict(cell=30, thr=50, max_features=512, mode=mode, max_iter=30, fa_patch=7, fa_mode=capi.LM_FAITHFUL)
                for _ in range(3):
                    out, sel, ref = ctx.frontend_run(pair["cur"], job, pair["feats"], 0, 0, 2, **kw)
                l0 = ctx.launches
                us = wall_time(lambda: ctx.frontend_run(pair["cur"], job, pair["feats"], 0, 0, 2, **kw), 50)
                kernels = (ctx.launches - l0) // 50
                # the same stages called one by one (every call synchronises and round-trips through the host)
                def composed():
                    ctx.upload(1, fr[1])
                    ctx.select_grid(1, 30, 50)
                    r, _ = ctx.sparse_align(job, pair["feats"], mode=mode, max_iter=30, want_stats=False)
                    items = np.zeros(n, capi.FA_ITEM_DTYPE)
                    items["ref_slot"], items["cur_slot"] = 0, 1
                    items["ref_px"], items["px"] = pair["feats"]["px"], pair["feats"]["px"]
                    items["A"] = (1, 0, 0, 1)
                    ctx.feature_align(items, patch_size=7, mode=capi.LM_FAITHFUL)
                for _ in range(3):
                    composed()
                comp_us = wall_time(composed, 30)
                # CPU: oracle stages on one thread
                t0 = time.perf_counter()
                ip, gp = orc.build_pyramid(pair["cur"], 4)
                g0 = orc.unpack_pyramid(gp, w, h, 4)[0]
                orc.grid_select(g0, 30, 50)
                rmse, T, status, lv = orc.sparse_align(pyr[0], pyr[1], ip, w, h, pair["feats"], pair["n_ref"], 0, pair["T_ref"], pair["T_kf"], K,
                                                       pair["T_cur_init"], mode=getattr(orc, "GN" if mode == capi.GN else "LM_FAITHFUL"), max_iter=30)
                rg = orc.unpack_pyramid(orc.build_pyramid(pair["ref"], 4)[1], w, h, 4)[0]
                for i in range(n):
                    orc.feature_align(rg, g0, pair["feats"]["px"][i], ref["px"][i] if np.isfinite(ref["px"][i]).all() else pair["feats"]["px"][i])
                cpu_us = (time.perf_counter() - t0) * 1e6
                bytes_ = 1239860.0 + 471000.0 + 0.44e6 + n * 212.0
                emit({"config": {"workload": "config 2: per-frame front end in one CUDA graph: frame H2D + pyramids + grid argmax (cell 30) + "
                                             "sparse alignment (%d features, %s) + reprojection + feature alignment (7x7, 1 step) + D2H" % (n, mname)},
                      "metric": "us_per_frame_front_end", "unit": "us", "higher_is_better": False, "value": us, "dtype": "mixed",
                      "kernels_per_graph": int(kernels), "n_selected": int(out["n_selected"]), "n_candidates": int(out["n_candidates"]),
                      "composed_calls_us": comp_us,
                      "e2e": {"value": us, "unit": "us", "h2d_bytes_per_step": int(h * w + 192 + 72 * 512 + 546),
                              "d2h_bytes_per_step": int(80 + 4 + 12 * 546 + 32 * 512),
                              "what": "svo_frontend_run: host copies into pinned mirrors, ONE cudaGraphLaunch, sync, host copies out; wall clock, median of 50"},
                      "roofline": roof(bytes_, us),
                      "cpu_baseline": {"value": cpu_us, "unit": "us", "cores": 1, "kind": "port",
                                       "sample": "the same frame through the oracle stages (pyramid, grid argmax, alignment, %d feature alignments), once" % n}})
        # ---------------- "next" row f3: epipolar search of the depth filter ----------------
        if 5 in want:
            wide = synth.make_pair(8, 500, motion_scale=3.0)   # a longer baseline: segments of tens of pixels
            ctx.upload(2, np.stack([wide["ref"], wide["cur"]]))
            rng = np.random.default_rng(5)
            ns = 2000
            idx = np.arange(ns) % len(wide["feats"])
            f = wide["feats"][idx]
            d_true = np.linalg.norm(f["point"], axis=1)
            items = np.zeros(ns, capi.EPI_ITEM_DTYPE)
            items["ref_slot"], items["cur_slot"] = 2, 3
            items["T_rel"] = synth.se3_mul(wide["T_cur_true"], synth.se3_inv(wide["T_ref"]))
            items["px"], items["bearing"] = f["px"], f["bearing"]
            items["depth"] = d_true * rng.uniform(0.8, 1.25, ns)
            items["min_depth"], items["max_depth"] = d_true * 0.5, d_true * 2.0
            r = ctx.epipolar_match(items)
            for _ in range(3):
                ctx.epipolar_match(items)
            e2e_us = wall_time(lambda: ctx.epipolar_match(items), 20)
            steps = r["steps"].astype(np.float64)
            bytes_ = float((49 * 4 + steps * 49 * 4 + 128 + 40).sum())  # 4 taps per patch pixel and step, item in, result out
            nsamp = 300
            t0 = time.perf_counter()
            agree = 0
            for i in range(nsamp):
                it = items[i]
                o = orc.epipolar_match(wide["ref"], wide["cur"], K, it["T_rel"], it["px"], it["bearing"], it["depth"], it["min_depth"],
                                       it["max_depth"])
                agree += int(o["found"] == bool(r[i]["found"]) and (not o["found"] or abs(o["depth"] - r[i]["depth"]) <= 1e-9 * o["depth"]))
            cpu_us = (time.perf_counter() - t0) * 1e6 / nsamp * ns
            ok = r["found"] == 1
            emit({"config": {"workload": "next row f3: depth-filter epipolar search (algorithm::matchEpipolarConstraint), %d seeds, 7x7 "
                                         "patches, depth interval [0.5, 2] x true depth, %.1f one-pixel steps per seed" % (ns, steps.mean())},
                  "metric": "us_per_batch_epipolar_match", "unit": "us", "higher_is_better": False, "value": e2e_us, "dtype": "f64 geometry, "
                  "float bilinear taps, uint8 patches", "seeds_per_sec": ns / (e2e_us * 1e-6), "found_fraction": float(ok.mean()),
                  "median_rel_depth_error": float(np.median(np.abs(r["depth"][ok] - d_true[ok]) / d_true[ok])),
                  "oracle_agreement_sample": "%d of %d" % (agree, nsamp),
                  "e2e": {"value": e2e_us, "unit": "us", "h2d_bytes_per_step": int(items.nbytes), "d2h_bytes_per_step": int(r.nbytes),
                          "what": "svo_epipolar_match (seeds H2D, one kernel, results D2H), host wall clock, median of 20"},
                  "roofline": roof(bytes_, e2e_us),
                  "cpu_baseline": {"value": cpu_us, "unit": "us", "cores": 1, "kind": "port",
                                   "sample": "first %d of the %d seeds through the oracle port, scaled to the batch" % (nsamp, ns)}})
        # ---------------- "next" row f4: pyramidal Lucas-Kanade (initialisation) ----------------
        if 8 in want:
            try:
                import cv2
            except ImportError:
                cv2 = None
            rng = np.random.default_rng(81)
            for npts in (500, 2000):
                base = pair["feats"]["px"][: pair["n_ref"]].astype(np.float32)
                pts = base if npts <= len(base) else np.concatenate(
                    [base, rng.uniform([8, 8], [w - 9, h - 9], size=(npts - len(base), 2)).astype(np.float32)])
                pts = np.ascontiguousarray(pts[:npts])
                got, gst, gerr = ctx.klt_track(0, 1, pts, pts, win=11)
                for _ in range(3):
                    ctx.klt_track(0, 1, pts, pts, win=11)
                us = wall_time(lambda: ctx.klt_track(0, 1, pts, pts, win=11), 30)
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    ow, ost, oerr, _ = orc.klt_track(pair["ref"], pair["cur"], pts, pts, win=11)
                    ts.append((time.perf_counter() - t0) * 1e6)
                both = (gst == 1) & (ost == 1)
                d = np.abs(got[both] - ow[both]).max(axis=1)
                cpu = {"value": float(np.median(ts)), "unit": "us", "cores": 1, "kind": "port",
                       "sample": "the same call through the oracle port (pyramids + derivative images rebuilt per call, as OpenCV does), median of 3"}
                if cv2 is not None:
                    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 1e-4)
                    tc = []
                    for _ in range(7):
                        t0 = time.perf_counter()
                        cw, cst, _ = cv2.calcOpticalFlowPyrLK(pair["ref"], pair["cur"], pts.copy(), pts.copy(), winSize=(11, 11), maxLevel=3,
                                                              criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
                        tc.append((time.perf_counter() - t0) * 1e6)
                    cpu = {"value": float(np.median(tc)), "unit": "us", "cores": cv2.getNumThreads(), "kind": "reference",
                           "sample": "cv2.calcOpticalFlowPyrLK %s of this image (the library call the reference makes), median of 7; "
                                     "the oracle port takes %.0f us on one thread" % (cv2.__version__, float(np.median(ts)))}
                    cb = (gst == 1) & (cst.reshape(-1) == 1)
                    cd = np.abs(got[cb] - cw[cb]).max(axis=1)
                # algorithmic bytes: per point and level the (win+3)^2 reference neighbourhood once + (win+1)^2 current-image bytes
                # per iteration (~8 iterations per level on this data); all of it L1/L2 resident
                emit({"config": {"workload": "next row f4: calcOpticalFlowPyrLK, %d points, 11x11 window, 4 levels, <= 30 iterations, eps 1e-4"
                                             % npts},
                      "metric": "us_per_frame_klt", "unit": "us", "higher_is_better": False, "value": us, "dtype": "i32/f32",
                      "parity": {"status_mismatches_vs_oracle": int((gst != ost).sum()), "tracked": int(both.sum()),
                                 "px_diff_q99_vs_oracle": float(np.quantile(d, 0.99)), "px_diff_max_vs_oracle": float(d.max()),
                                 "px_diff_q99_vs_cv2": float(np.quantile(cd, 0.99)) if cv2 is not None else None},
                      "e2e": {"value": us, "unit": "us", "h2d_bytes_per_step": int(2 * pts.nbytes), "d2h_bytes_per_step": int(pts.nbytes + 5 * npts),
                              "what": "svo_klt_track (points H2D, one kernel for all levels, positions/status/error D2H), host wall clock, "
                                      "median of 30; the pyramids are the frame slots' own"},
                      "roofline": roof(float(npts * 4 * (14 * 14 + 8 * 12 * 12)), us),
                      "cpu_baseline": cpu})
        # ---------------- "next" row f1: Map::reprojectMap ----------------
        if 7 in want:
            rng = np.random.default_rng(71)
            f = pair["feats"][pair["feats"]["has_point"] != 0]
            cands = np.zeros(2 * len(f), capi.REPROJ_CAND_DTYPE)
            cands["ref_slot"] = 0
            cands["ref_px"], cands["point"] = np.tile(f["px"], (2, 1)), np.tile(f["point"], (2, 1))
            cands["type"] = rng.choice([0, 1, 2, 3], size=len(cands), p=[0.4, 0.1, 0.2, 0.3])
            ncell = -(-w // 30) * -(-h // 30)
            order = rng.permutation(ncell).astype(np.int32)
            got, proj = ctx.reproject_map(1, pair["T_cur_true"], cands, 30, order)
            for _ in range(3):
                ctx.reproject_map(1, pair["T_cur_true"], cands, 30, order)
            us = wall_time(lambda: ctx.reproject_map(1, pair["T_cur_true"], cands, 30, order), 30)
            g0, g1 = ctx.download(0, 0, 1), ctx.download(1, 0, 1)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                ow, op = orc.reproject_map([g0, g1], g1, K, pair["T_cur_true"], cands, 30, order)
                ts.append((time.perf_counter() - t0) * 1e6)
            same = len(ow) == len(got) and np.array_equal(got["candidate"], ow[:, 1].astype(np.int32)) and \
                np.abs(got["px"] - ow[:, 2:4]).max() < 1e-7
            emit({"config": {"workload": "next row f1: Map::reprojectMap, %d candidates, cell 30, %d cells, %d matches (one 7x7 feature "
                                         "alignment each)" % (len(cands), ncell, len(got))},
                  "metric": "us_per_frame_reproject_map", "unit": "us", "higher_is_better": False, "value": us, "dtype": "f64",
                  "identical_to_oracle": bool(same),
                  "e2e": {"value": us, "unit": "us", "h2d_bytes_per_step": int(cands.nbytes + order.nbytes),
                          "d2h_bytes_per_step": int(151 * 40 + len(cands)),
                          "what": "svo_reproject_map (candidates + cell order H2D, four kernels, matches D2H), host wall clock, median of 30"},
                  "roofline": roof(float(48 * len(cands) + 212 * len(got)), us),
                  "cpu_baseline": {"value": float(np.median(ts)), "unit": "us", "cores": 1, "kind": "port",
                                   "sample": "the same call through the oracle port, median of 3"}})
        # ---------------- "next" row f2: SSC feature selection ----------------
        if 6 in want:
            grad0 = ctx.download(0, 0, 1)
            for thr, kc in ((50, 250), (20, 1000), (100, 100)):
                got, gi = ctx.select_ssc(0, thr, kc, 30)
                for _ in range(3):
                    ctx.select_ssc(0, thr, kc, 30)
                us = wall_time(lambda: ctx.select_ssc(0, thr, kc, 30), 20)
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    ow, oi = orc.select_ssc(grad0, thr, kc, 30)
                    ts.append((time.perf_counter() - t0) * 1e6)
                same = oi == gi and np.array_equal(np.stack([got["x"], got["y"], got["magnitude"]], 1), ow)
                bytes_ = float(h * w * (1 + gi["iterations"]) + 12 * len(got))   # the gradient image once per width + output
                emit({"config": {"workload": "next row f2: FeatureSelection::gradientMagnitudeWithSSC, one 1241x376 frame, threshold %d, %d "
                                             "candidates, cell 30, bucketing: %d keypoints, %d SSC widths, %d features" %
                                             (thr, kc, gi["keypoints"], gi["iterations"], len(got))},
                      "metric": "us_per_frame_select_ssc", "unit": "us", "higher_is_better": False, "value": us, "dtype": "u8 / int",
                      "identical_to_oracle": bool(same),
                      "e2e": {"value": us, "unit": "us", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(12 * 4096 + 24),
                              "what": "svo_select_ssc on a resident frame (one kernel, features D2H), host wall clock, median of 20"},
                      "roofline": roof(bytes_, us),
                      "cpu_baseline": {"value": float(np.median(ts)), "unit": "us", "cores": 1, "kind": "port",
                                       "sample": "the same frame through the oracle port (threshold scan, stable sort, SSC, bucketing), median of 3"}})
        pin.free()
        ctx.close()

    # ---------------- config 9: rows a1 + a2 + a5 on a batch of frames ----------------
    if 9 in want:
        nb = 64          # distinct frames (the selection line works on these)
        npyr = 1024      # resident frames of the pyramid line: the batch size of the headline workload
        batch = synth.make_batch(nb, 500)
        with torch.cuda.stream(stream):
            with pkg.Context(w, h, K, levels=4, max_frames=npyr, max_jobs=1, max_features=512, max_fa_items=16, stream=stream.cuda_stream) as c9:
                for i0 in range(0, npyr, nb):
                    c9.upload(i0, batch["cur"])
                c9.sync()
                for _ in range(3):
                    c9.rebuild(0, npyr)
                pyr_us = ev_time(torch, stream, lambda: c9.rebuild(0, npyr), 10)
                sel = [c9.select_grid(i, 30, 50) for i in range(nb)]

                def sel_all():
                    for i in range(nb):
                        c9.select_grid(i, 30, 50)

                sel_all()
                sel_us = wall_time(sel_all, 5)
                t0 = time.perf_counter()
                same = True
                for i in range(8):
                    op = orc.unpack_pyramid(orc.build_pyramid(batch["cur"][i], 4)[1], w, h, 4)
                    want_sel = orc.grid_select(op[0], 30, 50)
                    same = same and np.array_equal(np.stack([sel[i]["x"], sel[i]["y"], sel[i]["magnitude"]], 1), want_sel)
                cpu_us = (time.perf_counter() - t0) * 1e6 / 8
        dims = [(w, h)]
        for _ in range(3):
            dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))
        pyr_bytes = float(npyr * (2 * w * h + 2 * sum(a_ * b_ for a_, b_ in dims[1:]) + sum(2 * a_ * b_ for a_, b_ in dims[1:3])))  # read L0, write G0 + both stacks; levels 1-2 re-read
        emit({"config": {"workload": "rows a1 + a2: ImagePyramid::createImagePyramid (AbsGradientSaturatedSum + pyrDown of both stacks, 4 levels) of %d "
                                     "resident 1241x376 frames (svo_frames_rebuild)" % npyr},
              "metric": "us_per_frame_pyramid", "unit": "us", "higher_is_better": False, "value": pyr_us / npyr, "dtype": "u8",
              "roofline": dict(roof(pyr_bytes, pyr_us), note="instruction-bound u8 stencil (DESIGN.md 4): ~66 instructions per lane and source row for both stacks"),
              "e2e": {"value": pyr_us / npyr, "unit": "us", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                                         "what": "device time of the pyramid kernels over the batch (k_pyr_march: level 0 -> 1 with the gradient fused, then one launch pair per level; CUDA events), frames resident"},
              "cpu_baseline": {"value": None, "unit": "us", "cores": 1, "kind": "port", "sample": "see the next line (pyramid + selection timed together)"}})
        emit({"config": {"workload": "rows a4 + a5: FeatureSelection::gradientMagnitudeByValue (grid argmax, cell 30, threshold 50) on each of %d resident "
                                     "frames, one svo_select_grid call per frame (as Frame by Frame in the reference)" % nb},
              "metric": "us_per_frame_select_grid", "unit": "us", "higher_is_better": False, "value": sel_us / nb, "dtype": "u8",
              "identical_to_oracle": bool(same),
              "roofline": roof(float(nb * (w * h + 12 * 546)), sel_us),
              "e2e": {"value": sel_us / nb, "unit": "us", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(12 * 546 + 16),
                      "what": "svo_select_grid per frame: two kernels, features written zero-copy to mapped host memory, host wall clock"},
              "cpu_baseline": {"value": cpu_us, "unit": "us", "cores": 1, "kind": "port",
                               "sample": "oracle port on one host thread: pyramid (gradient + pyrDown x 2 stacks) + grid selection per frame, mean of 8 frames"}})

    # ---------------- config 3 ----------------
    if 3 in want:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--features", "1000", "--steps", "5", "--warmup", "3",
                            "--cpu-sample", "64"], capture_output=True, text=True)
        for l in p.stdout.splitlines():
            if l.startswith("{"):
                d = json.loads(l)
                d["config"]["workload"] = "config 3: " + d["config"]["workload"]
                lines.append(d)
                print(json.dumps(d), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
