#!/usr/bin/env bash
# tools/gpu_ncu.sh TAG [variant.so ...] — run ON THE GPU BOX: ncu --set full of k_trace / k_shade for the shipped library and variants.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TAG=$1; shift
timeout 300 python tools/prof_cmd.py 12 > $O/plain.log 2>&1 || { cat $O/plain.log; exit 1; }
cat $O/plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 16 -c 4 -f -o $O/prof_${TAG} python tools/prof_cmd.py 12 > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 $O/ncu_full.log
for v in "$@"; do
  n=$(basename $v .so)
  RT_LIB=$PWD/$v timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace' -s 8 -c 2 -f -o $O/prof_${TAG}_$n python tools/prof_cmd.py 12 > $O/ncu_full_$n.log 2>&1
  echo "ncu $n rc=$?"
done
ls -la $O/*.ncu-rep
