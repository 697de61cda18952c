set -x
B="python bench.py --pairs 148 --steps 1 --warmup 1 --no-cpu-baseline"
prof() { # name skip env
  env $3 ncu --set full --clock-control none --import-source on -k regex:k_align_cluster --launch-skip $2 --launch-count 1 -f -o gpurun_out/prof_$1 $B > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/raw_$1.csv 2>/dev/null
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_$1.csv 2>/dev/null
  python profiles/hot_lines.py gpurun_out/src_$1.csv 60 > gpurun_out/hot_$1.txt
  rm -f gpurun_out/src_$1.csv
}
prof r1f_nt256_c2 2 "SVO_ALIGN_NT=256"
prof r1f_single_nt64_c8 10 "A=1"
prof r1f_nt512_c1 2 "SVO_ALIGN_NT=512"
ls -la gpurun_out/
