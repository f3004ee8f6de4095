#!/bin/bash
# One GPU box, what profiles/ needs for round 2: tests, smoke, bench line + reference arm, launch list, ncu captures, sanitizer.
# Usage (under gpurun): bash tools/measure_r02.sh
R=r02
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${R}_gputests.txt; cat $O/${R}_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/${R}_smoke.txt; cat $O/${R}_smoke.txt
timeout 600 python bench.py 2>&1 | grep '^{' > $O/${R}_bench_n1.json; cut -c1-600 $O/${R}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | grep '^{' > $O/${R}_bench_reference.json; cut -c1-400 $O/${R}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 5 -c 2 -o $O/${R}_scan_k10 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-parity > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_scan_k10.ncu-rep > $O/${R}_ncu_scan_k10_summary.txt 2>&1
rm -f $O/${R}_scan_k10.ncu-rep   # reports are 10-50 MB each; gpurun brings back at most 64 MiB
CSGPU_BATCH_SIMT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_simt -s 4 -c 1 -o $O/${R}_simt_main python tools/bench_batch.py --dtype fp32 --cases 1024:100 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_simt_main.ncu-rep > $O/${R}_ncu_simt_main_summary.txt 2>&1; cat $O/${R}_ncu_simt_main_summary.txt
rm -f $O/${R}_simt_main.ncu-rep   # reports are 10-50 MB each; gpurun brings back at most 64 MiB
bash tools/measure_tf32.sh $R
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_multi -s 2 -c 1 -o $O/${R}_multi16 env CSGPU_GEMM_MIN_BATCH=100000 python tools/bench_batch.py --dtype fp32 --cases 16:100 --reps 2 > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_multi16.ncu-rep > $O/${R}_ncu_multi16_summary.txt 2>&1; cat $O/${R}_ncu_multi16_summary.txt
rm -f $O/${R}_multi16.ncu-rep   # reports are 10-50 MB each; gpurun brings back at most 64 MiB
bash tools/sanitize.sh $R
( for t in memcheck racecheck synccheck; do echo "# compute-sanitizer --tool $t python tools/sanitize_driver.py"; grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver ok|Error|hazard' $O/${R}_sanitize_$t.txt | head -12; done ) > $O/${R}_sanitizer.txt; cat $O/${R}_sanitizer.txt
ls -la $O | tail -30
