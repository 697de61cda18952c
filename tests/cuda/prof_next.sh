# ncu --set full of the next-row kernels (one launch each), summaries copied to profiles/ by hand
set -x
tag=$1
C="python bench_configs.py --configs 6,8,5,7"
$C > gpurun_out/plain_next_$tag.jsonl 2> gpurun_out/plain_next_$tag.err || exit 1
for k in k_select_ssc k_klt_track k_epipolar_match k_reproject_bin; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_${tag}_$k $C > gpurun_out/ncu_${tag}_$k.log 2>&1
  ncu -i gpurun_out/prof_${tag}_$k.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_$k.csv 2>/dev/null
  rm -f gpurun_out/prof_${tag}_$k.ncu-rep
done
