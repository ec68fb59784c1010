#!/usr/bin/env bash
# tools/gpu_r02h.sh TAG — run ON THE GPU BOX: converged-image tests, per-ray work counters (stats build), ncu --set full with source
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${1:-r02h}
timeout 600 python -m pytest tests -q -m gpu -x -k "converged" > $O/pytest_conv.log 2>&1; echo "pytest rc=$?" >> $O/pytest_conv.log; tail -4 $O/pytest_conv.log
timeout 300 python tools/stats_cmd.py 40 > $O/stats_$TAG.txt 2>&1; cat $O/stats_$TAG.txt
timeout 300 python tools/stats_cmd.py 16 1 3840 2160 500 > $O/stats_c5_$TAG.txt 2>&1; cat $O/stats_c5_$TAG.txt
timeout 300 python tools/prof_cmd.py 40 > $O/plain.log 2>&1 || { cat $O/plain.log; exit 1; }
cat $O/plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 24 -c 4 -f -o $O/prof_${TAG} python tools/prof_cmd.py 40 > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 $O/ncu_full.log
