"""Summarise an `ncu --page source --csv` export per barrier-delimited phase of each kernel.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_phases.py src.csv

For every kernel: instructions executed, stall samples and the top stall reasons between consecutive BAR.SYNC
instructions (a "phase" of the persistent item loop), so that hot phases can be mapped back to the source.
"""
import csv
import sys
from collections import Counter


def main(path):
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            body = []
            i += 2
            while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
                if len(rows[i]) == len(hdr):
                    body.append(rows[i])
                i += 1
            report(name, hdr, body)
        else:
            i += 1


def report(name, hdr, body):
    col = {h: k for k, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    print("=" * 100)
    print(name)
    tot_s = sum(int(r[col["# Samples"]]) for r in body)
    tot_i = sum(int(r[col["Instructions Executed"]]) for r in body)
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    phase, ph_rows = 0, []

    def flush():
        nonlocal ph_rows, phase
        if not ph_rows:
            return
        s = sum(int(r[col["# Samples"]]) for r in ph_rows)
        ins = sum(int(r[col["Instructions Executed"]]) for r in ph_rows)
        ops = Counter()
        for r in ph_rows:
            op = r[col["Source"]].split()
            op = [t for t in op if not t.startswith("@")]
            if op:
                ops[op[0].split(".")[0]] += int(r[col["Instructions Executed"]])
        st = Counter()
        for h in stall_cols:
            st[h[6:]] = sum(int(r[col[h]] or 0) for r in ph_rows)
        top = ", ".join(f"{k} {v * 100 // max(1, s)}%" for k, v in st.most_common(4))
        topo = ", ".join(f"{k} {v * 100 // max(1, ins)}%" for k, v in ops.most_common(5))
        print(f"phase {phase:3d}: {len(ph_rows):5d} sass  inst {ins * 100 / max(1, tot_i):5.1f}%  samples {s * 100 / max(1, tot_s):5.1f}%  [{top}]  ops[{topo}]")
        ph_rows = []
        phase += 1

    for r in body:
        ph_rows.append(r)
        if "BAR.SYNC" in r[col["Source"]]:
            flush()
    flush()


if __name__ == "__main__":
    main(sys.argv[1])
