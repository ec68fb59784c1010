#!/usr/bin/env bash
# tools/gpu_round_end.sh TAG — run ON THE GPU BOX: what the driver runs at round end (GPU tests, smoke, bench both arms), after an
# ncu capture of the shipped kernels (launch list + --set full of k_trace / k_shade at full wave size) that is summarised ON THE
# BOX into profiles/TAG_traffic.json, so that bench.py's roofline reads the per-ray figures of the tree it times. The raw
# report comes back in gpurun_out/ and the summary is regenerated from it at home (tools/summarize_profile.py TAG).
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${1:-r02zf}
timeout 900 python -m pytest tests -q -m gpu > $O/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; cat $O/final_smoke.log
timeout 300 python tools/prof_cmd.py 40 > $O/plain.log 2>&1 || { cat $O/plain.log; exit 1; }
cat $O/plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python tools/prof_cmd.py 40 > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 24 -c 4 -f -o $O/prof_${TAG} python tools/prof_cmd.py 40 > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 $O/ncu_full.log
timeout 300 python tools/summarize_profile.py $TAG; cp profiles/${TAG}_traffic.json $O/box_${TAG}_traffic.json
timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; tail -c 600 $O/final_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference.json 2>> $O/final_bench.err; echo "bench ref rc=$?"
python tools/c1_timing.py | tee $O/c1_timing_final.txt
