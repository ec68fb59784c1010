#!/usr/bin/env bash
# tools/ab.sh — run ON THE GPU BOX: prof_cmd with the shipped library and every variant in lib/variants.
cd "$(dirname "$0")/.."
SPP=${1:-200}
echo "== shipped"; python tools/prof_cmd.py $SPP; python tools/prof_cmd.py $SPP
for v in accelerated-ray-tracer_b200/lib/variants/*.so; do
  [[ $v == *stats* ]] && continue
  echo "== $v"; RT_LIB=$PWD/$v python tools/prof_cmd.py $SPP; RT_LIB=$PWD/$v python tools/prof_cmd.py $SPP
done
