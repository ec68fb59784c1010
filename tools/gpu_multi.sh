#!/usr/bin/env bash
# tools/gpu_multi.sh N WHAT TAG — run ON AN N-GPU BOX: multi-GPU correctness (dist_check, the 2-GPU pytest, rt_cli) and bench lines.
set -uo pipefail
cd "$(dirname "$0")/.."
N=${1:-2}; WHAT=${2:-check,bench}; TAG=${3:-r02}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
FAST="--no-cpu-baseline --no-reference-cuda"
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus_n$N.txt 2>&1
if [[ $WHAT == *check* ]]; then
  timeout 900 $TR --nproc-per-node $N --master-port 29571 tools/dist_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tee $O/dist_check_n${N}_$TAG.txt
  timeout 900 python -m pytest tests -q -m gpu -k "two_gpu" -rs 2>&1 | tail -5 | tee $O/pytest_two_gpu_$TAG.log
fi
if [[ $WHAT == *bench* ]]; then
  for n in ${NS:-$N}; do
    timeout 900 $TR --nproc-per-node $n --master-port 29572 bench.py --gpus $n $FAST > $O/bench_n${n}_spp_$TAG.json 2> $O/bench_n$n.err; echo "bench spp n=$n rc=$?"; cut -c1-700 $O/bench_n${n}_spp_$TAG.json
  done
fi
if [[ $WHAT == *tile* ]]; then
  timeout 900 $TR --nproc-per-node $N --master-port 29573 bench.py --gpus $N --split tile $FAST --no-e2e > $O/bench_n${N}_tile_$TAG.json 2>> $O/bench_n$N.err; echo "bench tile rc=$?"; cut -c1-500 $O/bench_n${N}_tile_$TAG.json
fi
if [[ $WHAT == *c5* ]]; then
  for n in ${NS5:-1 $N}; do
    if [[ $n == 1 ]]; then timeout 900 python bench.py --config c5-1m --steps 4 --warmup 2 --no-e2e > $O/bench_n1_c5_1m_$TAG.json 2>> $O/bench_n$N.err
    else timeout 900 $TR --nproc-per-node $n --master-port 29574 bench.py --gpus $n --config c5-1m --steps 4 --warmup 2 --no-e2e > $O/bench_n${n}_c5_1m_$TAG.json 2>> $O/bench_n$N.err; fi
    echo "bench c5-1m n=$n rc=$?"; cut -c1-500 $O/bench_n${n}_c5_1m_$TAG.json
  done
fi
[[ -f $O/bench_n$N.err ]] && tail -5 $O/bench_n$N.err; true
