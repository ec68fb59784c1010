#!/usr/bin/env bash
# tools/gpu_ncu2.sh TAG [VARIANT.so] — run ON THE GPU BOX: ncu --set full of k_trace (2 launches) with the shipped library or a variant
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${1:-ncu2}
[[ -n "${2:-}" ]] && export RT_LIB=$PWD/accelerated-ray-tracer_b200/lib/variants/$2
timeout 300 python tools/prof_cmd.py 40 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace' -s 24 -c 2 -f -o $O/prof_${TAG} python tools/prof_cmd.py 40 > $O/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_full.log
