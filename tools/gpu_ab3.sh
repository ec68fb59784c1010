#!/usr/bin/env bash
# tools/gpu_ab3.sh TAG — run ON THE GPU BOX: quick parity subset, packed-instruction microbenchmark, stats, A/B on C4 / C3 / C5-1M / C1
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${1:-ab3}
timeout 900 python -m pytest tests -q -m gpu -x -k "not converged" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
[[ -x tools/packed_bench ]] && tools/packed_bench | tee $O/packed_bench_$TAG.json
[[ -f accelerated-ray-tracer_b200/lib/variants/stats.so ]] && python tools/stats_cmd.py 40 | tee $O/stats_$TAG.txt
for cfg in "300 9 800 800" "200 8 600 600" "16 1 3840 2160 500" "100 1 1200 600"; do
  echo "== cfg $cfg shipped"; python tools/prof_cmd.py $cfg; python tools/prof_cmd.py $cfg
  for v in accelerated-ray-tracer_b200/lib/variants/*.so; do
    [[ $v == *stats* ]] && continue
    echo "== cfg $cfg $v"; RT_LIB=$PWD/$v python tools/prof_cmd.py $cfg; RT_LIB=$PWD/$v python tools/prof_cmd.py $cfg
  done
done 2>&1 | tee $O/ab_$TAG.txt
