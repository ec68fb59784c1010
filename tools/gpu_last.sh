#!/usr/bin/env bash
# tools/gpu_last.sh — run ON THE GPU BOX: final bench line, C1 timing, ncu of C5 at 10^6 spheres
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; cut -c1-300 $O/final_bench.json
python tools/c1_timing.py | tee $O/c1_timing_final.txt; python tools/c1_timing.py | tee -a $O/c1_timing_final.txt
timeout 300 python tools/prof_cmd.py 8 1 3840 2160 500 | tee $O/c5_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 8 -c 4 -f -o $O/prof_r02z_c5 python tools/prof_cmd.py 8 1 3840 2160 500 > $O/ncu_c5.log 2>&1
echo "ncu c5 rc=$?"; tail -2 $O/ncu_c5.log
