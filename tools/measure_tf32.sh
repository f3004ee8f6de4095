#!/bin/bash
# ncu captures of the tf32 batch kernel's main phase (the last, largest contraction launch of a batch) at 1024 and 128 queries.
# Usage (under gpurun): bash tools/measure_tf32.sh [tag]
R=${1:-r02}
O=gpurun_out
mkdir -p $O
export CSGPU_GEMM_MIN_BATCH=2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 4 -c 1 -o $O/${R}_tf32_main_1024 -f python tools/bench_batch.py --dtype fp32 --cases 1024:100 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_tf32_main_1024.ncu-rep > $O/${R}_ncu_tf32_main_1024_summary.txt 2>&1; cat $O/${R}_ncu_tf32_main_1024_summary.txt
rm -f $O/${R}_tf32_main_1024.ncu-rep   # reports are 10-50 MB each; gpurun brings back at most 64 MiB
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 4 -c 1 -o $O/${R}_tf32_main_128 -f python tools/bench_batch.py --dtype fp32 --cases 128:100 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_tf32_main_128.ncu-rep > $O/${R}_ncu_tf32_main_128_summary.txt 2>&1; cat $O/${R}_ncu_tf32_main_128_summary.txt
rm -f $O/${R}_tf32_main_128.ncu-rep   # reports are 10-50 MB each; gpurun brings back at most 64 MiB
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_tf32_1024.csv python tools/bench_batch.py --dtype fp32 --cases 1024:100 --reps 1 > /dev/null 2>&1
grep -c . $O/${R}_launches_tf32_1024.csv
