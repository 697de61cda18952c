# ncu --set full capture of one k_align_cluster launch + hot source lines; usage: prof_v3.sh <tag> <launch-skip> [env...]
set -x
tag=$1; skip=$2; shift 2
B="python bench.py --pairs 148 --steps 1 --warmup 1 --no-cpu-baseline"
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_align_cluster --launch-skip $skip --launch-count 1 -f -o gpurun_out/prof_$tag $B > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_$tag.csv 2>/dev/null
python profiles/hot_lines.py gpurun_out/src_$tag.csv 70 > gpurun_out/hot_$tag.txt
rm -f gpurun_out/src_$tag.csv
