#!/bin/bash
# Evidence pass for round 2 (run under gpurun, 1 GPU): pipe peaks, launch list of the bench command, one --set full
# capture per kernel family AT THE BASELINE SHAPES.  The .ncu-rep files are summarised on the box (scripts/ncu_top.py +
# the traffic / shared-memory-pipe metrics) and deleted: gpurun_out/ brings back at most 64 MiB.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
./scripts/probes/peaks > $O/r02_peaks.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 > $O/r02_bench_under_ncu.log 2>&1
summ() {  # $1 = tag
  python scripts/ncu_top.py /tmp/r02_prof_$1.ncu-rep 30 > $O/r02_ncu_$1_summary.txt 2>&1
  ncu -i /tmp/r02_prof_$1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
keys=['dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','gpu__time_duration.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fma_realtime.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fmaheavy_realtime.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fmalite_realtime.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__cycles_active.avg']
for a,b,c in zip(h,u,v):
    if any(a.endswith(k) for k in keys): print(a,'[',b,'] =',c)
" >> $O/r02_ncu_$1_summary.txt
  rm -f /tmp/r02_prof_$1.ncu-rep
}
for w in "sqeuclidean 100000 100000 128:pw_l2:expanded_tc" "cosine 100000 100000 128:pw_cos:expanded_tc" \
         "correlation 100000 100000 128:pw_corr:expanded_tc" "cityblock 50000 50000 256:ux_l1:unexpanded" \
         "nn 1000000 1000000 96:nnscreen:screen_tc" "nn 1000000 1000000 96:nnexact:expanded_tc" \
         "sqeuclidean_f16 50000 200000 64:pw_fp16:expanded_tc"; do
  IFS=: read a t kern <<< "$w"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$kern" -s 2 -c 1 -f \
      -o /tmp/r02_prof_$t python scripts/prof_pairwise.py $a 1 > $O/r02_ncu_$t.log 2>&1
  summ $t
done
du -sh $O; ls $O
