#!/usr/bin/env bash
# tools/gpu_r02.sh WHAT TAG — run ON THE GPU BOX: round-2 measurement set.
#   tests      pytest -m gpu
#   bench      bench.py (N = 1), both arms
#   micro      tools/microbench.cu (measured FP32 / issue / L2 / L1 / HBM / atomic peaks)
#   ncu        launch list + ncu --set full of k_trace / k_shade at FULL wave size (prof_cmd 40 spp: 8 Mi paths in flight)
#   ab         prof_cmd with the shipped library and every variant in lib/variants
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
WHAT=${1:-all}; TAG=${2:-r02}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
if [[ $WHAT == *tests* ]]; then
  timeout 1500 python -m pytest tests -q -m gpu -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
  tail -8 $O/pytest_gpu.log
fi
if [[ $WHAT == *micro* ]]; then
  timeout 300 accelerated-ray-tracer_b200/lib/rt_microbench > $O/microbench_$TAG.json 2> $O/microbench.err; echo "microbench rc=$?"; cat $O/microbench_$TAG.json
fi
if [[ $WHAT == *bench* ]]; then
  timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench_$TAG.json; tail -3 $O/bench.err
fi
if [[ $WHAT == *ab* ]]; then
  bash tools/ab.sh ${SPP:-300} 2>&1 | tee $O/ab_$TAG.txt
fi
if [[ $WHAT == *ncu* ]]; then
  SPPN=${SPPN:-40}
  timeout 300 python tools/prof_cmd.py $SPPN > $O/plain.log 2>&1 || { cat $O/plain.log; exit 1; }
  cat $O/plain.log
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python tools/prof_cmd.py $SPPN > $O/ncu_launches.log 2>&1
  echo "launch list rc=$?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_shade' -s 24 -c 4 -f -o $O/prof_${TAG} python tools/prof_cmd.py $SPPN > $O/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -2 $O/ncu_full.log
fi
ls -la $O | tail -8
if [[ $WHAT == *configs* ]]; then
  timeout 900 python tools/all_configs.py 2>&1 | tee $O/all_configs_$TAG.jsonl
fi
