#!/bin/bash
# The part of tools/measure_round.sh that the byte prefilter / headline bench line depend on (one GPU box, ~4 min).
R=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${R}_gputests.txt; cat $O/${R}_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/${R}_smoke.txt; cat $O/${R}_smoke.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/${R}_clocks_bench_n1.csv &
SMI=$!
timeout 600 python bench.py 2>&1 | grep '^{' > $O/${R}_bench_n1.json; cut -c1-600 $O/${R}_bench_n1.json
kill $SMI
timeout 300 python bench.py --byte-prefilter --steps 200 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{' > $O/${R}_bench_n1_byte_prefilter.json; cut -c1-300 $O/${R}_bench_n1_byte_prefilter.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
( timeout 300 python tools/bench_single.py --dim 384 --rows 10000000 --ks 1,10,32,100,128,256 --byte-prefilter --reps 60; timeout 300 python tools/bench_single.py --dim 768 --rows 5000000 --ks 10,100,200 --byte-prefilter --reps 60; timeout 200 python tools/bench_single.py --dim 384 --rows 1000000 --ks 10 --byte-prefilter --reps 60; timeout 120 python tools/i8_timing.py 10000000 10 2>&1 | tail -3; timeout 120 python tools/i8_timing.py 10000000 100 2>&1 | tail -2 ) > $O/${R}_bench_byte_prefilter.txt 2>&1; cut -c1-260 $O/${R}_bench_byte_prefilter.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_i8 -s 70 -c 1 -o $O/${R}_scan_i8 python tools/bench_single.py --rows 10000000 --dim 384 --ks 10 --byte-prefilter --reps 4 > /dev/null 2>&1
ls -la $O | tail -8
