#!/bin/bash
# Evidence pass for round 2 (run under gpurun, 1 GPU): pipe peaks, launch list of the bench command, one --set full
# capture per kernel family AT THE BASELINE SHAPES.  Outputs under gpurun_out/; summaries are copied to profiles/ here.
mkdir -p gpurun_out
./scripts/probes/peaks > gpurun_out/r02_peaks.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
for w in "sqeuclidean 100000 100000 128:pw_l2:expanded_tc" "cosine 100000 100000 128:pw_cos:expanded_tc" \
         "correlation 100000 100000 128:pw_corr:expanded_tc" "cityblock 50000 50000 256:ux_l1:unexpanded" \
         "nn 1000000 1000000 96:nnscreen:screen_tc" "nn 1000000 1000000 96:nnexact:expanded_tc"; do
  IFS=: read a t kern <<< "$w"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$kern" -s 2 -c 1 -f \
      -o gpurun_out/r02_prof_$t python scripts/prof_pairwise.py $a 1 > gpurun_out/r02_ncu_$t.log 2>&1
done
# fp16-in 200000 x 200000 x 64 (one 50000-row block): the k <= 64 full-width store path
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expanded_tc -s 2 -c 1 -f \
    -o gpurun_out/r02_prof_pw_fp16 python scripts/prof_pairwise.py sqeuclidean_f16 50000 200000 64 1 > gpurun_out/r02_ncu_pw_fp16.log 2>&1
tail -8 gpurun_out/r02_peaks.txt
