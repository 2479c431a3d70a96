#!/bin/bash
# round 2, GPU call G: warp-per-row burst-threshold kernel -- GPU tests, per-family times, ncu --set full of the default-set kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
o=gpurun_out/g
timeout 1200 python -m pytest tests -m gpu -x -q > ${o}_pytest.log 2>&1; tail -3 ${o}_pytest.log
for a in "default 256 60" "c4 32 300" "default 128 60 2000" "default 1024 20"; do
  timeout 600 python tools/profile_families.py $a >> ${o}_families.txt 2>&1
done
cat ${o}_families.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"Sharpwave|Bursts|nm_burst_thr_kernel|nm_burst_feat_kernel" -s 162 -c 12 -o /tmp/g_def \
    python tools/profile_families.py default 256 60 > ${o}_ncu.log 2>&1
tail -3 ${o}_ncu.log
ncu -i /tmp/g_def.ncu-rep --page raw --csv > ${o}_raw.csv
python tools/ncu_summary.py ${o}_raw.csv > ${o}_summary.txt
ncu -i /tmp/g_def.ncu-rep --page source --csv --print-source cuda,sass > /tmp/g_lines.csv
python tools/ncu_lines.py /tmp/g_lines.csv 30 > ${o}_lines.txt
ls -la ${o}_*
