#!/usr/bin/env bash
# tools/gpu_ref_round2.sh — run ON THE GPU BOX: the reference's own CUDA build (baseline/_ref/ref_gpu) under the
# reference's own per-scene stack / heap limits and managed framebuffer (round-2 harness):
#   1. C1 and C4 parity dumps again, compared with the committed goldens (made under the round-1 limits): must be identical;
#   2. converged images for the Philox-mode checks of C1 (400x225, 5000 spp) and C5 at 10 004 spheres (320x180, 2000 spp);
#   3. the timing that bench.py quotes as reference_cuda_sm100 (scene 9, 800x800, 8 spp).
set -uo pipefail
cd "$(dirname "$0")/.."
B=baseline/_ref/ref_gpu; T=oracle/_ref/textures; O=gpurun_out/ref2; mkdir -p $O; rm -f $O/*
run() { echo "+ $*" >> $O/log.txt; timeout 1500 "$@" >> $O/results.jsonl 2>> $O/log.txt || echo "FAILED($?): $*" >> $O/log.txt; }
run $B --scene 1 --nx 400 --ny 225 --ns 10 --ids 1 --reps 1 --textures $T --out $O/c1_400x225_10
run $B --scene 9 --nx 400 --ny 400 --ns 16 --ids 1 --reps 1 --textures $T --out $O/c4_400x400_16
run $B --scene 8 --nx 300 --ny 300 --ns 16 --ids 1 --reps 1 --textures $T --out $O/c3_300x300_16
python tools/pack_goldens.py $O
python - <<'PY'
import numpy as np
for n in ("c1_400x225_10", "c4_400x400_16", "c3_300x300_16"):
    a, b = np.load("gpurun_out/ref2/%s.npz" % n), np.load("tests/golden/ref_gpu/%s.npz" % n)
    same = all(np.array_equal(a[k].view(np.uint8) if a[k].dtype.kind == "f" else a[k], b[k].view(np.uint8) if b[k].dtype.kind == "f" else b[k]) for k in ("sd", "ids_obj", "ids_t", "ids_mat", "fb"))
    print("golden %s under the reference's own limits: %s" % (n, "IDENTICAL to the committed golden" if same else "DIFFERS"))
PY
rm -f $O/c1_400x225_10.npz $O/c4_400x400_16.npz $O/c3_300x300_16.npz
run $B --scene 1 --nx 400 --ny 225 --ns 5000 --reps 1 --count 0 --textures $T --out $O/c1_400x225_5000
run $B --scene 1 --grid 50 --nx 320 --ny 180 --ns 2000 --reps 1 --count 0 --textures $T --out $O/c5_10k_320x180_2000
python tools/pack_goldens.py $O
rm -f $O/*.sd $O/*.ids $O/*.fb
run $B --scene 9 --nx 800 --ny 800 --ns 8 --reps 2 --count 0 --textures $T
cat $O/results.jsonl | cut -c1-400; tail -3 $O/log.txt; ls -la $O
