"""Hot source lines of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`: stall samples and
executed instructions per CUDA source line.   python tools/ncu_hot_lines.py dump.csv [top_n]"""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    agg, tot = defaultdict(lambda: [0.0, 0.0, ""]), 0.0
    for r in rows[hi + 1:]:
        if len(r) < 8 or not r[0].strip():
            continue
        try:
            ln, s, inst = int(r[0]), float(r[6] or 0), float(r[7] or 0)
        except ValueError:
            continue
        agg[ln][0] += s; agg[ln][1] += inst; agg[ln][2] = r[1].strip()[:110]
        tot += s
    print("total samples", tot)
    for ln, (s, inst, src) in sorted(sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]):
        print(f"{ln:4d} {100 * s / tot:5.1f}% inst {int(inst):7d}  {src}")


if __name__ == "__main__":
    main()
