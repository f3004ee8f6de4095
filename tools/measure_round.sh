#!/bin/bash
# One GPU box, everything profiles/ needs for a round: tests, bench line, launch list, ncu captures, secondary benches.
# Usage (under gpurun): bash tools/measure_round.sh r01
R=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${R}_gputests.txt; cat $O/${R}_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/${R}_smoke.txt; cat $O/${R}_smoke.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/${R}_clocks_bench_n1.csv &
SMI=$!
timeout 600 python bench.py 2>&1 | grep '^{' > $O/${R}_bench_n1.json; cat $O/${R}_bench_n1.json
kill $SMI
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | grep '^{' > $O/${R}_bench_reference.json; cat $O/${R}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 5 -c 2 -o $O/${R}_scan_k10 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
( timeout 300 python tools/bench_batch.py --dtype fp32 --cases 1:10,8:10,9:200,8:100,16:100,64:100,128:100,1024:100,1024:10 --reps 4; timeout 300 python tools/bench_batch.py --dtype bf16 --cases 1:10,128:100,1024:100,1024:10 --reps 4 --recall; timeout 300 python tools/bench_batch.py --dtype fp32 --prefilter --cases 9:200,64:100,128:100,1024:100,1024:10 --reps 6 ) > $O/${R}_bench_batch.txt 2>&1; cat $O/${R}_bench_batch.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_simt -s 4 -c 1 -o $O/${R}_simt_main python tools/bench_batch.py --dtype fp32 --cases 1024:100 --reps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel -s 4 -c 1 -o $O/${R}_bf16_main python tools/bench_batch.py --dtype bf16 --cases 1024:100 --reps 1 > /dev/null 2>&1
( timeout 300 python tools/bench_single.py --dim 384 --rows 10000000 --ks 10,100,200,1000 --densities 1.0,0.25,0.01; timeout 300 python tools/bench_single.py --dim 768 --rows 5000000 --ks 10,200 --densities 1.0,0.25,0.01; timeout 300 python tools/bench_hybrid.py ) > $O/${R}_bench_single_k_filter.txt 2>&1; cat $O/${R}_bench_single_k_filter.txt
( timeout 300 python tools/bench_single.py --dim 384 --rows 10000000 --ks 1,10,32,100,128 --byte-prefilter --reps 60; timeout 300 python tools/bench_single.py --dim 768 --rows 5000000 --ks 10,100 --byte-prefilter --reps 60; timeout 200 python tools/bench_single.py --dim 384 --rows 1000000 --ks 10 --byte-prefilter --reps 60; timeout 120 python tools/i8_timing.py 10000000 10 2>&1 | tail -3 ) > $O/${R}_bench_byte_prefilter.txt 2>&1; cat $O/${R}_bench_byte_prefilter.txt
timeout 200 python tools/bench_variants.py > $O/${R}_bench_variants.txt 2>&1; cat $O/${R}_bench_variants.txt
timeout 300 python bench.py --byte-prefilter --steps 200 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{' > $O/${R}_bench_n1_byte_prefilter.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_bench_n1_byte_prefilter.csv python bench.py --byte-prefilter --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 120 ./build/tune_scan 10000000 20 > $O/${R}_tune_scan.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_i8 -s 70 -c 1 -o $O/${R}_scan_i8 python tools/bench_single.py --rows 10000000 --dim 384 --ks 10 --byte-prefilter --reps 4 > /dev/null 2>&1
( for m in 0 1 0 1; do echo "CSGPU_SCAN_STATIC=$m"; CSGPU_SCAN_STATIC=$m timeout 300 python tools/bench_single.py --rows 10000000 --dim 384 --ks 10,100 --reps 60 2>&1 | tail -2; done ) > $O/${R}_scan_static_vs_dynamic.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:select_sorted -s 3 -c 1 -o $O/${R}_select_rescore python tools/bench_batch.py --dtype fp32 --prefilter --cases 1024:100 --reps 1 > /dev/null 2>&1
ls -la $O | tail -20
