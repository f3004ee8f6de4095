#!/bin/bash
# A/B of the batch path's phase growth (CSGPU_BATCH_GROWTH) on the tf32 route, one box.
for g in ${@:-0 32 64 0 64}; do
  echo "CSGPU_BATCH_GROWTH=$g (0 = default rule)"
  for rows in 100000 1000000 10000000; do
    if [ "$g" = "0" ]; then python tools/tf32_probe.py $rows 384 "16:10,128:10,1024:10,64:32" 2>&1 | cut -c1-200
    else CSGPU_BATCH_GROWTH=$g python tools/tf32_probe.py $rows 384 "16:10,128:10,1024:10,64:32" 2>&1 | cut -c1-200; fi
  done
done
