#!/bin/bash
# round 2, kernel changes: tests, then SIMT retile + 64-query tile, dynamic split for the filtered and multi-query scans (A/B)
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/r02_t3.txt; cat $O/r02_t3.txt
( echo "== default routing"; timeout 300 python tools/bench_batch.py --dtype fp32 --cases 16:100,24:100,32:100,64:100,128:100,256:100,1024:100,1024:10 --reps 4;
  echo "== CSGPU_GEMM_MIN_BATCH=10 (SIMT tile from 10 queries on)"; CSGPU_GEMM_MIN_BATCH=10 timeout 300 python tools/bench_batch.py --dtype fp32 --cases 16:100,24:100,32:100,40:100,48:100,64:100 --reps 4 ) > $O/r02_bench_batch_simt.txt 2>&1; cat $O/r02_bench_batch_simt.txt
true
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_simt -s 4 -c 1 -o $O/r02_simt_main python tools/bench_batch.py --dtype fp32 --cases 1024:100 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py $O/r02_simt_main.ncu-rep > $O/r02_ncu_simt_main_summary.txt 2>&1; cat $O/r02_ncu_simt_main_summary.txt
