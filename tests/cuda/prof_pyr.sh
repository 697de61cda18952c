B="python tests/cuda/e2e_probe.py"
ncu --set full --clock-control none --import-source on -k regex:k_pyr_level --launch-skip 6 --launch-count 1 -f -o gpurun_out/prof_r1s_pyr $B > gpurun_out/ncu_r1s_pyr.log 2>&1
ncu -i gpurun_out/prof_r1s_pyr.ncu-rep --page raw --csv > gpurun_out/raw_r1s_pyr.csv 2>/dev/null
ncu -i gpurun_out/prof_r1s_pyr.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_r1s_pyr.csv 2>/dev/null
python profiles/hot_lines.py gpurun_out/src_r1s_pyr.csv 25 > gpurun_out/hot_r1s_pyr.txt
rm -f gpurun_out/src_r1s_pyr.csv
python profiles/summarize_ncu.py gpurun_out/raw_r1s_pyr.csv | python -c "
import sys,json
d=json.load(sys.stdin)[0]
print(d['kernel'][:60])
for k,v in d.items():
    if isinstance(v,dict): print('  ',k,v['value'],v['unit'])
for s in d['top_stalls'][:6]: print('     ',s['metric'][34:-23], s['value'])
"
head -24 gpurun_out/hot_r1s_pyr.txt | cut -c1-150
