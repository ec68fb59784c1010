#!/usr/bin/env bash
# tools/gpu_c1.sh — run ON THE GPU BOX: C1-exact timing over hand-over thresholds and host-check intervals
cd "$(dirname "$0")/.."
for wb in 8 4 3 2; do for tr in 16384 65536 200000; do
  echo "== shipped RT_WAVE_BATCH=$wb RT_TAIL_RAYS=$tr"; RT_WAVE_BATCH=$wb RT_TAIL_RAYS=$tr python tools/c1_timing.py
done; done 2>&1 | tee gpurun_out/c1_timing.txt
