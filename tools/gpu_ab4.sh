#!/usr/bin/env bash
# tools/gpu_ab4.sh TAG "cfg;cfg;..." — run ON THE GPU BOX: stats of every lib/stats/*.so, then A/B of shipped + lib/variants/*.so on the given prof_cmd configs
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${1:-ab4}; CFGS=${2:-"300 9 800 800"}
for s in accelerated-ray-tracer_b200/lib/variants/stats.so accelerated-ray-tracer_b200/lib/stats/*.so; do
  [[ -f $s ]] || continue
  echo "== stats $s"; STATS_LIB=$PWD/$s python tools/stats_cmd.py 40
done 2>&1 | tee $O/stats_$TAG.txt
IFS=';' read -ra CF <<< "$CFGS"
for cfg in "${CF[@]}"; do
  echo "== cfg $cfg shipped"; python tools/prof_cmd.py $cfg; python tools/prof_cmd.py $cfg
  for v in accelerated-ray-tracer_b200/lib/variants/*.so; do
    [[ $v == *stats* ]] && continue
    echo "== cfg $cfg $v"; RT_LIB=$PWD/$v python tools/prof_cmd.py $cfg; RT_LIB=$PWD/$v python tools/prof_cmd.py $cfg
  done
done 2>&1 | tee $O/ab_$TAG.txt
# env sweeps on the shipped library (RT_SLOTS / RT_POOLS / RT_WAVE_BATCH)
if [[ -n "${ENVSWEEP:-}" ]]; then
  IFS=';' read -ra EV <<< "$ENVSWEEP"
  for ev in "${EV[@]}"; do
    echo "== cfg 300 9 800 800 env:${ev// /,}"; env $ev python tools/prof_cmd.py 300; env $ev python tools/prof_cmd.py 300
  done 2>&1 | tee -a $O/ab_$TAG.txt
fi
