#!/usr/bin/env python
"""hot_lines.py source_page.csv [N] -- top source lines of `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`
by warp-stall samples and by executed warp instructions (per file:line, inlined code attributed to its own line)."""
import csv
import sys
from collections import defaultdict


def main():
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    fpath, hdr = None, None
    samples, insts, text = defaultdict(float), defaultdict(float), {}
    stall_cols = {}
    stalls = defaultdict(lambda: defaultdict(float))
    for r in csv.reader(open(sys.argv[1])):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
            continue
        if hdr is None or r[0] in ("", "Function Name") or not r[0].isdigit():
            continue
        key = (fpath, int(r[0]))
        try:
            samples[key] += float(r[i_s])
            insts[key] += float(r[i_i])
        except (ValueError, IndexError):
            continue
        text[key] = r[1].strip()[:90]
        for i, h in stall_cols.items():
            try:
                stalls[key][h] += float(r[i])
            except (ValueError, IndexError):
                pass
    ts, ti = sum(samples.values()), sum(insts.values())
    print("total samples %.0f, warp instructions %.0f" % (ts, ti))
    print("\n== top lines by stall samples")
    for key, v in sorted(samples.items(), key=lambda kv: -kv[1])[:n]:
        top = sorted(stalls[key].items(), key=lambda kv: -kv[1])[:3]
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s   [%s]" % (100 * v / ts, 100 * insts[key] / ti, key[0], key[1], text[key],
                                                          ", ".join("%s %.0f" % (h[6:], x) for h, x in top if x > 0)))
    print("\n== per file")
    pf_s, pf_i = defaultdict(float), defaultdict(float)
    for k, v in samples.items():
        pf_s[k[0]] += v
        pf_i[k[0]] += insts[k]
    for f in sorted(pf_s, key=lambda f: -pf_s[f]):
        print("%5.1f%% smp %5.1f%% inst  %s" % (100 * pf_s[f] / ts, 100 * pf_i[f] / ti, f))


if __name__ == "__main__":
    main()
