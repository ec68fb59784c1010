#!/usr/bin/env bash
# tools/gpu_slots.sh — run ON THE GPU BOX: paths in flight (RT_SLOTS) against job size on C4
cd "$(dirname "$0")/.."
for spp in 30 125 1000; do for sl in 4194304 8388608 16777216 33554432 67108864; do
  echo "== spp $spp slots $sl"; RT_SLOTS=$sl python tools/prof_cmd.py $spp; RT_SLOTS=$sl python tools/prof_cmd.py $spp
done; done 2>&1 | tee gpurun_out/slots_sweep.txt
