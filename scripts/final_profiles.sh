#!/bin/bash
# Evidence pass (run under gpurun): launch list of the bench command + one --set full capture per kernel family.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
for w in "sqeuclidean 32768 32768 128:pw:expanded_tc" "nn 131072 262144 96:nnscreen:screen_tc" "nn 131072 262144 96:nnexact:expanded_tc" "cityblock 16384 16384 256:ux:unexpanded"; do
  IFS=: read a t kern <<< "$w"
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$kern" -s 2 -c 1 \
      -o gpurun_out/prof_final_$t python scripts/prof_pairwise.py $a 1 > gpurun_out/ncu_final_$t.log 2>&1
done
python bench.py --steps 10 --warmup 3 2>/dev/null > gpurun_out/bench_n1.json
cut -c1-300 gpurun_out/bench_n1.json
