#!/bin/bash
# 8 GPUs of one box (gpurun --gpus 8): BASELINE configs[3] — the 400M-row corpus, 50M rows per GPU — through the fp32
# scan and through the byte prefilter, plus the 80M-row weak-scaling point with the byte prefilter.
R=${1:-r01}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29541 bench.py --gpus 8 --rows-per-gpu 50000000 --steps 60 --warmup 3 2>&1 | grep '^{' > $O/${R}_bench_n8_400M.json; cut -c1-400 $O/${R}_bench_n8_400M.json
timeout 400 $TR --master-port 29542 bench.py --gpus 8 --rows-per-gpu 50000000 --steps 100 --warmup 3 --byte-prefilter 2>&1 | grep '^{' > $O/${R}_bench_n8_400M_byte_prefilter.json; cut -c1-400 $O/${R}_bench_n8_400M_byte_prefilter.json
timeout 300 $TR --master-port 29543 bench.py --gpus 8 --steps 200 --warmup 5 --byte-prefilter 2>&1 | grep '^{' > $O/${R}_bench_n8_byte_prefilter.json; cut -c1-400 $O/${R}_bench_n8_byte_prefilter.json
timeout 300 $TR --master-port 29544 bench.py --gpus 8 --steps 200 --warmup 5 2>&1 | grep '^{' > $O/${R}_bench_n8_fused.json; cut -c1-400 $O/${R}_bench_n8_fused.json
