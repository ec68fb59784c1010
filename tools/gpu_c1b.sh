#!/usr/bin/env bash
cd "$(dirname "$0")/.."
for wb in 8 4 2 1; do for tr in 16384 65536 200000; do
  echo "== RT_WAVE_BATCH=$wb RT_TAIL_RAYS=$tr"; RT_WAVE_BATCH=$wb RT_TAIL_RAYS=$tr python tools/c1_timing.py
done; done 2>&1 | tee gpurun_out/c1_timing_b.txt
