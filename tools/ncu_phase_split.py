"""Split an `ncu --page source --csv` (SASS view) dump of one kernel into phases by address range and print, per phase:
executed warp instructions, share, stall-sample breakdown, shared-memory wavefronts (ideal vs actual) and an opcode
histogram weighted by executed count.   python tools/ncu_phase_split.py dump.csv name:lo:hi [name:lo:hi ...]"""
import csv
import sys
from collections import Counter, defaultdict


def main():
    path, specs = sys.argv[1], sys.argv[2:]
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    phases = []
    for s in specs:
        n, lo, hi = s.split(":")
        phases.append((n, int(lo, 16), int(hi, 16)))
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    base = None
    agg = defaultdict(lambda: defaultdict(float))
    ops = defaultdict(Counter)
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
        if base is None:
            base = addr
        off = addr - base
        ph = next((n for n, lo, hi in phases if lo <= off < hi), "other")
        f = lambda k: float(r[col[k]] or 0)
        a = agg[ph]
        a["inst"] += f("Instructions Executed")
        a["thread_inst"] += f("Thread Instructions Executed")
        a["samples"] += f("# Samples")
        a["wf"] += f("L1 Wavefronts Shared")
        a["wf_ideal"] += f("L1 Wavefronts Shared Ideal")
        for s in stall_cols:
            a[s] += f(s)
        src = r[col["Source"]].split()
        op = src[1] if src and src[0].startswith("@") else (src[0] if src else "?")
        ops[ph][op.split(".")[0]] += f("Instructions Executed")
    tot = sum(a["inst"] for a in agg.values())
    tots = sum(a["samples"] for a in agg.values())
    print(f"total warp instructions {tot:.0f}, stall samples {tots:.0f}")
    for ph in [p[0] for p in phases] + ["other"]:
        a = agg.get(ph)
        if not a:
            continue
        print(f"\n== {ph}: {a['inst']:.0f} warp inst ({100 * a['inst'] / tot:.1f} %), samples {a['samples']:.0f} "
              f"({100 * a['samples'] / max(tots, 1):.1f} %), smem wavefronts {a['wf']:.0f} (ideal {a['wf_ideal']:.0f})")
        st = sorted(((a[s], s) for s in stall_cols if a[s] > 0), reverse=True)
        print("   stalls: " + ", ".join(f"{s[6:]} {100 * v / max(a['samples'], 1):.1f}%" for v, s in st[:8]))
        print("   opcodes: " + ", ".join(f"{o} {100 * c / a['inst']:.1f}%" for o, c in ops[ph].most_common(12)))


if __name__ == "__main__":
    main()
