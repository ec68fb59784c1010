#!/usr/bin/env bash
# tools/gpu_capture_bench.sh TAG — run ON THE GPU BOX: ncu --set full of k_trace / k_shade in the STEADY STATE of a long job
# (300 spp of C4 = 192 M samples on 32 Mi paths in flight: every wave is refilled by regeneration; 12 waves skipped),
# summarised on the box into profiles/TAG_traffic.json, then bench.py, whose roofline reads that capture.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${1:-r02zg}
rm -f $O/launches.csv
timeout 300 python tools/prof_cmd.py 300 > $O/plain.log 2>&1 || { cat $O/plain.log; exit 1; }
cat $O/plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 96 -c 4 -f -o $O/prof_${TAG} python tools/prof_cmd.py 300 > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 $O/ncu_full.log
timeout 300 python tools/summarize_profile.py $TAG; cp profiles/${TAG}_traffic.json $O/box_${TAG}_traffic.json
timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; tail -c 300 $O/final_bench.json
