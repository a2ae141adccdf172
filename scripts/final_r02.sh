#!/bin/bash
# Final verification pass of round 2 (one gpurun call): GPU test tier, N=1 bench line, 1-GPU fusedL2NN time,
# compute-sanitizer over every kernel family, ncu capture of the large coarse-pass launch.
mkdir -p gpurun_out/fin
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/fin/pytest_gpu.txt
python bench.py > gpurun_out/fin/bench_n1.json 2> gpurun_out/fin/bench_n1.err
python scripts/nn_time.py 8388608 0 > gpurun_out/fin/nn_8m.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize.py > gpurun_out/fin/sanitizer_memcheck.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize.py > gpurun_out/fin/sanitizer_racecheck.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:screen_tc_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/fin/screen_full -f python scripts/nn_time.py 1048576 0 > gpurun_out/fin/screen_full.log 2>&1
python scripts/ncu_top.py gpurun_out/fin/screen_full.ncu-rep 40 > gpurun_out/fin/ncu_nnscreen_summary.txt 2>&1
ncu -i gpurun_out/fin/screen_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for a,b,c in zip(h,u,v):
    if ('mem_shared' in a and 'wavefronts' in a and 'pct' in a) or 'dram__bytes_read.sum' == a or 'dram__bytes_write.sum' == a: print(a,b,c)
" >> gpurun_out/fin/ncu_nnscreen_summary.txt
rm -f gpurun_out/fin/screen_full.ncu-rep
tail -3 gpurun_out/fin/pytest_gpu.txt; tail -c 600 gpurun_out/fin/bench_n1.json; cat gpurun_out/fin/nn_8m.txt; tail -3 gpurun_out/fin/sanitizer_memcheck.txt; tail -3 gpurun_out/fin/sanitizer_racecheck.txt
