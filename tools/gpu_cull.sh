#!/usr/bin/env bash
# tools/gpu_cull.sh — run ON THE GPU BOX: parity tests with the RT_STACK_CULL build, then A/B of shipped + lib/variants/*.so
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=$PWD/accelerated-ray-tracer_b200/lib/variants
RT_LIB=$V/cull.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_rng_mode or random_scene or primary_hit or scale_up or statistics" 2>&1 | tail -5 | tee $O/cull_tests.txt
for cfg in "300 9 800 800" "1000 7 600 600" "64 1 3840 2160 500" "100 1 1200 600"; do
  echo "== cfg $cfg shipped"; python tools/prof_cmd.py $cfg; python tools/prof_cmd.py $cfg
  for v in $V/*.so; do
    echo "== cfg $cfg $(basename $v)"; RT_LIB=$v python tools/prof_cmd.py $cfg; RT_LIB=$v python tools/prof_cmd.py $cfg
  done
done 2>&1 | tee $O/ab_cull.txt
