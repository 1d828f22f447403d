"""`ncu --csv --log-file X` (one row per launch and metric) -> one row per launch: id, kernel, grid, time_ns, DRAM bytes
read / written, warp instructions; prints the share of every kernel in one serial pass of the step.
    python tools/ncu_launch_list.py ncu_long.csv out.csv"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    launches = OrderedDict()
    for r in rows[1:]:
        if r[col["ID"]] == "ID":
            continue
        d = launches.setdefault(int(r[col["ID"]]), {"kernel": r[col["Kernel Name"]][:48], "grid": r[col["Grid Size"]]})
        unit, val = r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
        name = r[col["Metric Name"]]
        if name == "gpu__time_duration.sum":
            val *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        if name.startswith("dram__bytes"):
            val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[name] = val
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "time_ns", "dram_read_bytes", "dram_write_bytes", "warp_inst"])
        for i, d in launches.items():
            w.writerow([i, d["kernel"], d["grid"], d.get("gpu__time_duration.sum"), int(d.get("dram__bytes_read.sum", 0)),
                        int(d.get("dram__bytes_write.sum", 0)), int(d.get("smsp__inst_executed.sum", 0))])
    # shares of one serial pass: the first launch of each of the step's kernels after the warm-up passes (full batch)
    names = ["k1_fast_kernel", "k2_decode_kernel", "k2_box_kernel", "k3_nms_kernel", "k4_units_kernel", "k5_measure_kernel"]
    per = defaultdict(list)
    for d in launches.values():
        m = re.search(r"k\d_\w+", d["kernel"])
        if m and m.group(0) in names:
            per[m.group(0)].append((d["grid"], d["gpu__time_duration.sum"]))
    full_grid = {k: v[0][0] for k, v in per.items()}
    means = {k: [t for g, t in v if g == full_grid[k]] for k, v in per.items()}
    n_full = min(len(v) for v in means.values())
    means = {k: sum(v[:n_full]) / n_full for k, v in means.items()}
    tot = sum(means.values())
    for k in names:
        if k in means:
            print(f"{k:20s} {full_grid[k]:14s} mean of {n_full:2d} launches {means[k] / 1e3:7.1f} us  share {100 * means[k] / tot:5.1f} %")


if __name__ == "__main__":
    main()
