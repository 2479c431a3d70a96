"""Aggregate the warp-stall samples of an `ncu --page source --csv --print-source sass` export between consecutive
synchronisation instructions (phases of a barrier-structured kernel).  usage: ncu_segments.py file.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hdr = rows[1]
si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
seg, cur, idx = [], dict(samples=0, n=0, exe=0, start=0, fp64=0, lds=0, ldg=0, top=(0, '')), 0
for r in rows[2:]:
    if len(r) <= ii:
        continue
    src = r[1]
    s, e = int(r[si] or 0), int(r[ii] or 0)
    cur['samples'] += s; cur['n'] += 1; cur['exe'] += e
    if s > cur['top'][0]:
        cur['top'] = (s, src.strip()[:50])
    if any(x in src for x in ('DADD', 'DMUL', 'DFMA', 'DSETP')): cur['fp64'] += 1
    if 'LDS' in src or 'STS' in src: cur['lds'] += 1
    if 'LDG' in src or 'LDL' in src or 'STL' in src: cur['ldg'] += 1
    idx += 1
    if 'BAR.SYNC' in src or 'SYNCS' in src or 'WARPSYNC' in src:
        cur['end'] = src.strip()[:24]; seg.append(cur)
        cur = dict(samples=0, n=0, exe=0, start=idx, fp64=0, lds=0, ldg=0, top=(0, ''))
seg.append(cur)
tot = sum(s['samples'] for s in seg)
print("total samples", tot, "instructions", idx)
for s in seg:
    if s['samples'] > minpct / 100 * tot:
        print(f"{s['start']:6d} n={s['n']:5d} fp64={s['fp64']:4d} lds={s['lds']:3d} ldg={s['ldg']:3d} exe/inst={s['exe']/max(s['n'],1):9.0f} "
              f"{100*s['samples']/tot:5.1f}%  top={s['top']}  {s.get('end','')}")
