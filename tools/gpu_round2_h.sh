#!/bin/bash
# round 2, GPU call H: warp-per-row burst thresholds, second version (2 rows per 128-thread CTA, ring loads up front) + ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
o=gpurun_out/h
timeout 900 python -m pytest tests -m gpu -x -q -k "burst or c4 or rawnorm or raw_normal or stream" > ${o}_pytest.log 2>&1; tail -2 ${o}_pytest.log
for a in "default 256 60" "c4 32 300"; do
  timeout 600 python tools/profile_families.py $a >> ${o}_families.txt 2>&1
done
cat ${o}_families.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"nm_convx_kernel|nm_burst" -s 162 -c 12 -o /tmp/h_def \
    python tools/profile_families.py default 256 60 > ${o}_ncu.log 2>&1
tail -3 ${o}_ncu.log
ncu -i /tmp/h_def.ncu-rep --page raw --csv > ${o}_raw.csv
python tools/ncu_summary.py ${o}_raw.csv > ${o}_summary.txt
ncu -i /tmp/h_def.ncu-rep --page source --csv --print-source cuda,sass > /tmp/h_lines.csv
python tools/ncu_lines.py /tmp/h_lines.csv 30 > ${o}_lines.txt
ls -la ${o}_*
