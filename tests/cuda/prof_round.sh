# Round profile under gpurun (one GPU): (1) plain bench, (2) ncu launch list of the same command, (3) ncu --set full of
# the alignment kernel and of the fused pyramid kernel.  Outputs in gpurun_out/, summaries copied to profiles/ by hand.
set -x
tag=$1
B="python bench.py --pairs 148 --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_$tag.json 2> gpurun_out/plain_$tag.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
for k in k_align_cluster k_pyr_level; do
  skip=2; [ $k = k_pyr_level ] && skip=0
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip --launch-count 1 -f -o gpurun_out/prof_${tag}_$k $B > gpurun_out/ncu_${tag}_$k.log 2>&1
  ncu -i gpurun_out/prof_${tag}_$k.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_$k.csv 2>/dev/null
  ncu -i gpurun_out/prof_${tag}_$k.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_${tag}_$k.csv 2>/dev/null
  python profiles/hot_lines.py gpurun_out/src_${tag}_$k.csv 40 > gpurun_out/hot_${tag}_$k.txt
  rm -f gpurun_out/src_${tag}_$k.csv
done
