#!/usr/bin/env bash
# tools/gpu_ab2.sh [SPP] — run ON THE GPU BOX: pytest -m gpu with the shipped library, then the A/B of tools/ab.sh on C4, and C3 / C5-1M lines
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${2:-ab}
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
bash tools/ab.sh ${1:-300} 2>&1 | tee $O/ab_$TAG.txt
for cfg in "200 8 600 600" "16 1 3840 2160 500"; do
  echo "== cfg $cfg shipped"; python tools/prof_cmd.py $cfg
  for v in accelerated-ray-tracer_b200/lib/variants/*.so; do
    [[ $v == *stats* ]] && continue
    echo "== cfg $cfg $v"; RT_LIB=$PWD/$v python tools/prof_cmd.py $cfg
  done
done 2>&1 | tee -a $O/ab_$TAG.txt
