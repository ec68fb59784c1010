#!/usr/bin/env bash
# tools/gpu_dbg.sh — run ON THE GPU BOX: small diagnostics (adaptive error map, C1 wave log, pool count on C1)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
rm -f $O/adaptive_log.txt $O/wl_c1.txt
RT_ADAPTIVE_LOG=$O/adaptive_log.txt python - <<'PY'
import sys; sys.path.insert(0, "accelerated-ray-tracer_b200")
import pyrt
with pyrt.Scene(1, 160, 96) as sc:
    a = sc.render_adaptive(max_spp=128, threshold=0.0, pass_spp=32, tile=16)
    print("probe", a.err_min, a.err_max, a.err_spp)
PY
head -16 $O/adaptive_log.txt
RT_WAVE_LOG=$O/wl_c1.txt python tools/prof_cmd.py 10 1 400 225
awk '{t+=$3; s+=$4} NR<=12 || NR%8==0 {print} END {print "trace ms", t, "shade ms", s}' $O/wl_c1.txt
for p in 1 2 4; do for i in 1 2 3; do RT_POOLS=$p python tools/prof_cmd.py 10 1 400 225; done; done
for i in 1 2; do python tools/prof_cmd.py 100 1 1200 600; done
for t in 0 4096 16384 65536; do for i in 1 2 3; do echo -n "tail=$t "; RT_TAIL_RAYS=$t python tools/prof_cmd.py 10 1 400 225; done; done
for t in 0 16384; do echo -n "tail=$t "; RT_TAIL_RAYS=$t python tools/prof_cmd.py 1000; done
