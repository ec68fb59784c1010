#!/usr/bin/env bash
# tools/gpu_ref_converged.sh — run ON THE GPU BOX: high-spp renders of the reference's own CUDA build for the
# PSNR >= 40 dB test (the reference needs ~2e4 spp for its own Monte-Carlo noise to drop below -40 dB).
set -uo pipefail
cd "$(dirname "$0")/.."
B=baseline/_ref/ref_gpu; T=oracle/_ref/textures; O=gpurun_out/ref; mkdir -p $O
run() { echo "+ $*" >> $O/log.txt; timeout 1200 "$@" >> $O/results_converged.jsonl 2>> $O/log.txt || echo "FAILED($?): $*" >> $O/log.txt; }
run $B --scene 7 --nx 160 --ny 160 --ns 30000 --reps 1 --count 0 --textures $T --out $O/c2_160x160_30000
run $B --scene 8 --nx 160 --ny 160 --ns 30000 --reps 1 --count 0 --textures $T --out $O/c3_160x160_30000
run $B --scene 9 --nx 160 --ny 160 --ns 24000 --reps 1 --count 0 --textures $T --out $O/c4_160x160_24000
python tools/pack_goldens.py $O
rm -f $O/*.sd $O/*.ids $O/*.fb
cat $O/results_converged.jsonl; tail -3 $O/log.txt
