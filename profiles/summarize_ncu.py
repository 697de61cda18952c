#!/usr/bin/env python
"""summarize_ncu.py raw.csv [out.json] -- condense `ncu -i X.ncu-rep --page raw --csv` into the handful of
counters DESIGN.md and bench.py quote (per kernel launch): duration, DRAM bytes, issue activity, occupancy limits
and the top warp-stall reasons.  The .ncu-rep files themselves stay in gpurun_out/ (scratch, too large to commit)."""
import csv
import json
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_op_shared_atom.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return s


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": d.get("Kernel Name"), "id": d.get("ID")}
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                rec[k] = {"value": num(d[k]), "unit": units[hdr.index(k)]}
        stalls = []
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                if d[k] not in ("", "n/a"):
                    stalls.append((num(d[k]), k))
        stalls.sort(reverse=True)
        rec["top_stalls"] = [{"metric": k, "value": v} for v, k in stalls[:8]]
        out.append(rec)
    js = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(js + "\n")
    else:
        print(js)


if __name__ == "__main__":
    main()
