#!/usr/bin/env bash
# tools/gpu_ref_goldens.sh — run ON THE GPU BOX (via gpurun). Runs the reference's own CUDA build
# (baseline/_ref/ref_gpu, built here by oracle/build_ref.sh) and leaves scene dumps, primary-hit ID
# buffers, framebuffers and timings under gpurun_out/ref/. tools/pack_goldens.py packs them into
# tests/golden/.
set -uo pipefail
cd "$(dirname "$0")/.."
B=baseline/_ref/ref_gpu
T=oracle/_ref/textures
O=gpurun_out/ref
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
run() { echo "+ $*" >> $O/log.txt; timeout 900 "$@" >> $O/results.jsonl 2>> $O/log.txt || echo "FAILED($?): $*" >> $O/log.txt; }
# --- parity goldens (low spp, full dumps; packed to .npz, raw files removed: gpurun_out <= 64 MiB) ---
run $B --scene 1 --nx 400 --ny 225 --ns 10 --ids 1 --reps 1 --textures $T --out $O/c1_400x225_10
run $B --scene 7 --nx 300 --ny 300 --ns 16 --ids 1 --reps 1 --textures $T --out $O/c2_300x300_16
run $B --scene 8 --nx 300 --ny 300 --ns 16 --ids 1 --reps 1 --textures $T --out $O/c3_300x300_16
run $B --scene 9 --nx 400 --ny 400 --ns 16 --ids 1 --reps 1 --textures $T --out $O/c4_400x400_16
for s in 2 3 4 5 6 10; do
  run $B --scene $s --nx 300 --ny 150 --ns 8 --ids 1 --reps 1 --textures $T --out $O/s${s}_300x150_8
done
python tools/pack_goldens.py $O
# full-resolution primary-hit ID buffers (no framebuffer)
run $B --scene 7 --nx 600 --ny 600 --ns 1 --ids 1 --reps 0 --count 0 --textures $T --out $O/c2_600x600_ids
run $B --scene 8 --nx 600 --ny 600 --ns 1 --ids 1 --reps 0 --count 0 --textures $T --out $O/c3_600x600_ids
run $B --scene 9 --nx 800 --ny 800 --ns 1 --ids 1 --reps 0 --count 0 --textures $T --out $O/c4_800x800_ids
python tools/pack_goldens.py $O
# converged images for the PSNR test: tools/gpu_ref_converged.sh (separate call: ~10 GPU-minutes)
rm -f $O/*.sd $O/*.ids $O/*.fb
cat $O/results.jsonl
tail -5 $O/log.txt
du -sh gpurun_out
