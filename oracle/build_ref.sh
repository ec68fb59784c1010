#!/usr/bin/env bash
# oracle/build_ref.sh — builds the two harnesses that run the UNMODIFIED reference code.
#   oracle/_ref/ref_cpu      reference headers + main.cu device part, g++ behind oracle/shim (CPU)
#   baseline/_ref/ref_gpu    same sources, nvcc -arch=sm_100 (the reference's own CUDA build)
#   oracle/_ref/textures/*.ppm  textures decoded with the reference's vendored stb_image
# Reference sources are compiled from a throw-away staging copy under /tmp (never copied into the
# repo). Staging edits, all mechanical:
#   * hittable.cuh:26  `= default;` -> `{}`   (nvcc 12.9 "execution space mismatch" otherwise)
#   * main.cu cut before the first host function (the host drivers use <<<>>> and MSVC-era flow)
#   * GRID_MIN/GRID_MAX (main.cu:140-141) read a run-time variable so config C5 can scale the grid
# Needs /root/reference; on the GPU box only the prebuilt outputs are used.
set -euo pipefail
REF=${REF:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
[ -d "$REF/src" ] || { echo "build_ref: $REF not present, keeping prebuilt outputs"; exit 0; }
STAGE=$(mktemp -d /tmp/rt_ref_stage.XXXXXX)
trap 'rm -rf "$STAGE"' EXIT
cp -r "$REF/src" "$REF/external" "$STAGE/"
chmod -R u+w "$STAGE"
sed -i 's/__device__ virtual ~hittable() = default;/__device__ virtual ~hittable() {}/' "$STAGE/src/hittable.cuh"
awk '/^void bouncing_spheres\(\)/{exit} {print}' "$STAGE/src/main.cu" > "$STAGE/src/ref_device_part.inc"
sed -i -e 's/^#define GRID_MIN .*/#define GRID_MIN (-g_grid_half)/' \
       -e 's/^#define GRID_MAX .*/#define GRID_MAX (g_grid_half)/' "$STAGE/src/ref_device_part.inc"
mkdir -p "$ROOT/oracle/_ref/textures" "$ROOT/baseline/_ref"
WHAT=${1:-all}
if [ "$WHAT" = all ] || [ "$WHAT" = cpu ]; then
  g++ -std=c++17 -O2 -ffp-contract=off -fopenmp -w \
      -DREF_DEVICE_PART='"ref_device_part.inc"' \
      -I "$ROOT/oracle/shim" -I "$STAGE/src" -I "$ROOT/include" -I "$ROOT/oracle/ref_harness" \
      -o "$ROOT/oracle/_ref/ref_cpu" "$ROOT/oracle/ref_harness/ref_cpu.cpp"
  for t in earthmap poolball porcelain 8ball; do
    "$ROOT/oracle/_ref/ref_cpu" --decode "$REF/textures/$t.jpg" "$ROOT/oracle/_ref/textures/$t.ppm"
  done
  "$ROOT/oracle/_ref/ref_cpu" --decode "$REF/textures/hardwood.jpg" "$ROOT/oracle/_ref/textures/hardwood.ppm"
  # the reference's own .jpg files next to what ITS decoder makes of them: input and expected output of the product's
  # JPEG decoder test (oracle/_ref is git-ignored; it travels to the GPU box like the other built files)
  mkdir -p "$ROOT/oracle/_ref/textures_jpg"
  cp "$REF"/textures/*.jpg "$ROOT/oracle/_ref/textures_jpg/"
  echo "built oracle/_ref/ref_cpu + textures"
fi
if [ "$WHAT" = all ] || [ "$WHAT" = gpu ]; then
  nvcc -std=c++17 -arch=sm_100 -O3 --expt-relaxed-constexpr --expt-extended-lambda -diag-suppress=20199 -w \
      -DREF_DEVICE_PART='"ref_device_part.inc"' \
      -I "$STAGE/src" -I "$ROOT/include" -I "$ROOT/oracle/ref_harness" \
      -o "$ROOT/baseline/_ref/ref_gpu" "$ROOT/oracle/ref_harness/ref_gpu.cu"
  echo "built baseline/_ref/ref_gpu"
fi
