"""Hot source lines of a kernel from `ncu -i rep --page source --csv --print-source cuda,sass` (needs -lineinfo + --import-source on).

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > lines.csv ; python tools/ncu_lines.py lines.csv [top_n] [kernel substring]

Per kernel: stall samples per CUDA source line (file:line), the share of the kernel's samples and the source text.
"""
import csv
import sys
from collections import defaultdict

csv.field_size_limit(1 << 30)


def main(path, top=40, want=""):
    kern, fname = "", ""
    per = defaultdict(lambda: defaultdict(lambda: [0, 0, ""]))  # kernel -> (file, line) -> [samples, instructions, text]
    col = {}
    for r in csv.reader(open(path, errors="replace")):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            kern = r[1]
        elif r[0] == "Line No":
            col = {h: i for i, h in enumerate(r)}
        elif r[0] and r[0].isdigit() and col:
            try:
                s = int(r[col["# Samples"]])
                n = int(r[col["Instructions Executed"]])
            except (ValueError, KeyError, IndexError):
                continue
            e = per[kern][(fname, int(r[0]))]
            e[0] += s
            e[1] += n
            e[2] = r[1].strip()
    for k, lines in per.items():
        if want and want not in k:
            continue
        tot = sum(v[0] for v in lines.values()) or 1
        print("=" * 120)
        print(k[:200])
        print(f"total samples {tot}")
        by_file = defaultdict(int)
        for (f, _), v in lines.items():
            by_file[f] += v[0]
        print("  per file: " + ", ".join(f"{f} {100.0 * s / tot:.1f}%" for f, s in sorted(by_file.items(), key=lambda kv: -kv[1])))
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"  {100.0 * v[0] / tot:5.1f}%  {v[1]:>10d} inst  {f}:{ln:<5d} {v[2][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else "")
