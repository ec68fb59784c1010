#!/usr/bin/env bash
# tools/gpu_ab.sh — run ON THE GPU BOX: GPU tests, then prof_cmd with the shipped library and every variant.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
if [[ ${1:-tests} == tests ]]; then
  timeout 900 python -m pytest tests -q -m gpu -x -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
  tail -15 $O/pytest_gpu.log
fi
bash tools/ab.sh ${2:-300} 2>&1 | tee $O/ab.txt
rm -f $O/wavelog.txt
RT_WAVE_LOG=$O/wavelog.txt timeout 300 python tools/prof_cmd.py 100 > $O/wavelog.out 2>&1; cat $O/wavelog.out
python - <<'PY'
import sys
rows=[l.split() for l in open('gpurun_out/wavelog.txt')]
tr=sum(float(r[2]) for r in rows); sh=sum(float(r[3]) for r in rows); rays=sum(int(r[1]) for r in rows)
print("wavelog: waves %d rays %d trace %.2f ms shade %.2f ms -> trace %.1f Mrays/s shade %.1f Mrays/s" % (len(rows), rays, tr, sh, rays/tr/1e3, rays/sh/1e3))
PY
