#!/bin/bash
# round 2, GPU call F: ncu --set full of the slow kernels of the default feature set (current build: single-buffer banks on 1536 /
# 2048 points, split burst thresholds), steady state (full history ring: chunks 7 and 8 of a 591-window run)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
o=gpurun_out/f
ncu --set full --import-source on --clock-control none -k regex:"Sharpwave|Bursts|nm_burst_thr_kernel|nm_burst_feat_kernel" -s 162 -c 12 -o /tmp/f_def \
    python tools/profile_families.py default 256 60 > ${o}_ncu.log 2>&1
tail -5 ${o}_ncu.log
ncu -i /tmp/f_def.ncu-rep --page raw --csv > ${o}_raw.csv
python tools/ncu_summary.py ${o}_raw.csv > ${o}_summary.txt
ncu -i /tmp/f_def.ncu-rep --page source --csv --print-source cuda,sass > /tmp/f_lines.csv
python tools/ncu_lines.py /tmp/f_lines.csv 30 > ${o}_lines.txt
ls -la ${o}_*
