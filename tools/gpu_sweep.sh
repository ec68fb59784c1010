#!/usr/bin/env bash
# tools/gpu_sweep.sh — run ON THE GPU BOX: slots x pools x variant sweep of prof_cmd (C4, 1000 spp)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
SPP=${SPP:-1000}
for v in ${VARIANTS:-shipped}; do
  for slots in ${SLOTS:-2097152 4194304 8388608 16777216}; do
    for pools in ${POOLS:-1 2 4}; do
      lib=$PWD/accelerated-ray-tracer_b200/lib/variants/$v.so
      [[ $v == shipped ]] && lib=$PWD/accelerated-ray-tracer_b200/lib/librt_b200.so
      echo -n "$v slots=$slots pools=$pools: "
      RT_LIB=$lib RT_SLOTS=$slots RT_POOLS=$pools timeout 120 python tools/prof_cmd.py $SPP 2>&1 | tail -1
    done
  done
done | tee $O/sweep.txt
