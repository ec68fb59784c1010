#!/usr/bin/env bash
# tools/gpu_ref_converged.sh — run ON THE GPU BOX: high-spp renders of the reference's own CUDA build for the
# PSNR >= 40 dB test (the reference needs ~2e4 spp for its own Monte-Carlo noise to drop below -40 dB).
set -uo pipefail
cd "$(dirname "$0")/.."
B=baseline/_ref/ref_gpu; T=oracle/_ref/textures; O=gpurun_out/ref; mkdir -p $O
run() { echo "+ $*" >> $O/log.txt; timeout 1200 "$@" >> $O/results_converged.jsonl 2>> $O/log.txt || echo "FAILED($?): $*" >> $O/log.txt; }
# The reference loops a pixel's samples inside one thread, so its run time is ~ns x (time of one sample) unless the
# image has enough pixels to fill the GPU: render MORE pixels at moderate spp and box-downsample in linear radiance
# (pack_goldens.py, suffix _dsK) instead of few pixels at huge spp. Same estimator, K*K*ns samples per final pixel.
WHAT=${1:-c2,c3,c4}
[[ $WHAT == *c2* ]] && run $B --scene 7 --nx 640 --ny 640 --ns 2000 --reps 1 --count 0 --textures $T --out $O/c2_640x640_2000_ds4
[[ $WHAT == *c3* ]] && run $B --scene 8 --nx 640 --ny 640 --ns 2000 --reps 1 --count 0 --textures $T --out $O/c3_640x640_2000_ds4
[[ $WHAT == *c4* ]] && run $B --scene 9 --nx 800 --ny 800 --ns 1000 --reps 1 --count 0 --textures $T --out $O/c4_800x800_1000_ds5
python tools/pack_goldens.py $O
rm -f $O/*.sd $O/*.ids $O/*.fb
cat $O/results_converged.jsonl; tail -3 $O/log.txt
