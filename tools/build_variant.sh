#!/usr/bin/env bash
# tools/build_variant.sh NAME "-DFOO=1 ..." — A/B builds of the library for tuning runs (lib/variants/NAME.so; never shipped).
set -e
cd "$(dirname "$0")/.."
C=accelerated-ray-tracer_b200/csrc
mkdir -p accelerated-ray-tracer_b200/lib/variants
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off $2 \
  -I $C -I include -shared -o accelerated-ray-tracer_b200/lib/variants/$1.so $C/rt_host.cu $C/scene_builder.cpp $C/generators.cpp $C/jpeg_baseline.cpp
