# Round-2 measurement set on one GPU: bench line, per-config lines, ncu launch list, ncu --set full of the alignment kernel.
# usage: bash tests/cuda/measure_r2.sh <tag>     (outputs under gpurun_out/, condensed copies are committed under profiles/)
tag=$1
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench_configs.py --out gpurun_out/${tag}_bench_configs.jsonl > gpurun_out/${tag}_bench_configs.log 2>&1
# launch list (cold caches, serialised launches: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
bash tests/cuda/prof_v5.sh $tag 1 > gpurun_out/${tag}_prof.log 2>&1
python profiles/summarize_ncu.py gpurun_out/raw_$tag.csv gpurun_out/${tag}_ncu_full_k_align_v5.json
ls -la gpurun_out | grep $tag
