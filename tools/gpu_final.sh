#!/usr/bin/env bash
# tools/gpu_final.sh — run ON THE GPU BOX: what the driver runs at round end (GPU tests, smoke, bench both arms) plus the ncu
# launch list of a short bench.py command (share of each kernel in a step).
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; cat $O/final_smoke.log
timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; tail -c 600 $O/final_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference.json 2>> $O/final_bench.err; echo "bench ref rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --spp-per-step 8 --no-e2e --no-cpu-baseline --no-reference-cuda"
timeout 300 $CMD > $O/final_bench_short.json 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/launches_bench.csv $CMD > $O/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
