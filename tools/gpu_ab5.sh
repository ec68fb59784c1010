#!/usr/bin/env bash
# tools/gpu_ab5.sh TAG — run ON THE GPU BOX: A/B on C4 / C5-1M / C5-100k, C1-exact tail sweep (RT_TAIL_RAYS), quick parity
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${1:-ab5}
timeout 600 python -m pytest tests -q -m gpu -x -k "not converged" 2>&1 | tail -3
bash tools/gpu_ab4.sh $TAG "300 9 800 800;16 1 3840 2160 500;16 1 3840 2160 158"
for tr in 0 4096 16384 65536 262144 1000000; do
  echo "== cfg 10 1 400 225 env:RT_TAIL_RAYS=$tr"; for i in 1 2 3; do RT_TAIL_RAYS=$tr python tools/prof_cmd.py 10 1 400 225; done
done 2>&1 | tee -a $O/ab_$TAG.txt
