#!/bin/bash
# round 2, GPU call K: ncu --set full of the sorted-queue warp-per-row threshold kernel (one steady-state launch, 64 rows, no split)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
o=gpurun_out/k
NMB200_BURST_SPLIT_OCC=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"nm_burst_thr" -s 27 -c 1 -o /tmp/k_thr \
    python tools/profile_families.py c4 32 60 > ${o}_ncu.log 2>&1
tail -3 ${o}_ncu.log
ncu -i /tmp/k_thr.ncu-rep --page raw --csv > ${o}_raw.csv
python tools/ncu_summary.py ${o}_raw.csv > ${o}_summary.txt
ncu -i /tmp/k_thr.ncu-rep --page source --csv --print-source cuda,sass > /tmp/k_lines.csv
python tools/ncu_lines.py /tmp/k_lines.csv 40 > ${o}_lines.txt
python - <<'PY' > gpurun_out/k_sass_top.txt
import csv, sys
csv.field_size_limit(1 << 30)
rows = []
col = {}
for r in csv.reader(open('/tmp/k_lines.csv', errors='replace')):
    if not r: continue
    if r[0] in ('Line No', '#', 'Address') or 'Source' in r[:2]:
        col = {h: i for i, h in enumerate(r)}
        continue
    if col and '# Samples' in col and len(r) > col['# Samples']:
        try:
            s = int(r[col['# Samples']])
        except ValueError:
            continue
        rows.append((s, r[:3]))
rows.sort(key=lambda t: -t[0])
for s, r in rows[:60]:
    print(s, ' | '.join(x[:110] for x in r))
PY
cat ${o}_summary.txt | head -40
