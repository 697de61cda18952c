# ncu --set full capture of one k_align_v5 launch + hot source lines + per-opcode stall samples
# usage: prof_v3.sh <tag> <launch-skip> [env...]
set -x
tag=$1; skip=$2; shift 2
B="python bench.py --pairs 148 --steps 1 --warmup 1 --no-cpu-baseline"
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_align_v5 --launch-skip $skip --launch-count 1 -f -o gpurun_out/prof_$tag $B > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_$tag.csv 2>/dev/null
python profiles/hot_lines.py gpurun_out/src_$tag.csv 70 > gpurun_out/hot_$tag.txt
python - <<PY > gpurun_out/opcodes_$tag.txt
import csv
from collections import defaultdict
smp=defaultdict(float); ins=defaultdict(float); hdr=None
for r in csv.reader(open("gpurun_out/src_$tag.csv")):
    if not r: continue
    if r[0]=="Line No": hdr=r; i_s=hdr.index("# Samples"); i_i=hdr.index("Instructions Executed"); continue
    if hdr and r[0]=="" and len(r)>i_i and r[2].startswith("0x"):
        op=r[3].strip().split()[0]
        if op.startswith("@"): op=r[3].strip().split()[1]
        op=op.split(".")[0]
        try: smp[op]+=float(r[i_s]); ins[op]+=float(r[i_i])
        except ValueError: pass
ts=sum(smp.values()); ti=sum(ins.values())
print("opcode  samples%  instructions%  samples/instr(rel)")
for op,v in sorted(smp.items(), key=lambda kv:-kv[1])[:35]:
    print("%-10s %6.2f %6.2f %6.2f" % (op, 100*v/ts, 100*ins[op]/ti, (v/ts)/max(ins[op]/ti,1e-9)))
PY
rm -f gpurun_out/src_$tag.csv
