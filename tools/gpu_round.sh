#!/usr/bin/env bash
# tools/gpu_round.sh — run ON THE GPU BOX (gpurun): tests, bench, launch list, ncu capture of the hot kernels.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt
WHAT=${1:-all}
if [[ $WHAT == all || $WHAT == *tests* ]]; then
  timeout 1500 python -m pytest tests -q -m gpu -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
  tail -25 $O/pytest_gpu.log
fi
if [[ $WHAT == all || $WHAT == *bench* ]]; then
  timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err; cat $O/bench_reference.json
fi
if [[ $WHAT == all || $WHAT == *wavelog* ]]; then
  rm -f $O/wavelog.txt
  RT_WAVE_LOG=$O/wavelog.txt timeout 300 python tools/prof_cmd.py 100 > $O/wavelog.out 2>&1; cat $O/wavelog.out
fi
if [[ $WHAT == all || $WHAT == *ncu* ]]; then
  timeout 300 python tools/prof_cmd.py 12 > $O/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv python tools/prof_cmd.py 12 > $O/ncu_launches.log 2>&1
  echo "launch list rc=$?"; cat $O/plain.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 16 -c 4 -f -o $O/prof_${TAG:-r01} python tools/prof_cmd.py 12 > $O/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 $O/ncu_full.log
fi
ls -la $O | head -40
if [[ $WHAT == *slots* ]]; then
  for n in 262144 524288 1048576 2097152 4194304; do RT_SLOTS=$n timeout 300 python tools/prof_cmd.py 300; done > $O/slots_sweep.txt 2>&1; cat $O/slots_sweep.txt
fi
