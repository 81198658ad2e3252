"""Summarises an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel-name totals, shares of one step and (when captured) DRAM traffic and achieved GB/s.

    python tools/summarize_launches.py gpurun_out/launches.csv [first_id last_id] > profiles/rNN_launches_summary.txt

Without an id window the last complete step is the span between the last two launches of the stem input kernel pair
(the first kernels of every step; stem_pack since round 2)."""
from __future__ import annotations

import collections
import csv
import re
import sys

MUL = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("cstp::", "").replace("at::native::", "at::")
    return name[:70]


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = by.setdefault(int(r["ID"]), {"name": short(r["Kernel Name"]), "t": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r["Metric Value"].replace(",", "")) * MUL.get(r["Metric Unit"], 1.0)
        key = {"gpu__time_duration.sum": "t", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr"}.get(r["Metric Name"])
        if key:
            d[key] = v
    return by


def main():
    by = load(sys.argv[1])
    ids = list(by)
    if len(sys.argv) >= 4:
        lo, hi = int(sys.argv[2]), int(sys.argv[3])
    else:
        stems = [i for i in ids if by[i]["name"].startswith(("stem_im2col", "stem_pack"))]
        starts = [s for k, s in enumerate(stems) if k == 0 or s - stems[k - 1] > 1]
        lo, hi = starts[-2], starts[-1]
    step = [by[i] for i in ids if lo <= i < hi]
    tot = sum(d["t"] for d in step)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in step:
        a = agg[d["name"]]
        a[0] += 1
        a[1] += d["t"]
        a[2] += d["rd"] + d["wr"]
    has_dram = any(a[2] > 0 for a in agg.values())
    print(f"# launches [{lo},{hi}) = {len(step)} kernels, sum of kernel durations {tot / 1e6:.3f} ms "
          "(ncu-serialised, cold cache: compare shares, not absolutes)")
    print(f"{'kernel':70s} {'launches':>8s} {'ms':>9s} {'share':>7s}" + (f" {'DRAM GB':>9s} {'GB/s':>8s}" if has_dram else ""))
    for n, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"{n:70s} {c:8d} {t / 1e6:9.3f} {100 * t / tot:6.1f}%"
        if has_dram:
            line += f" {b / 1e9:9.2f} {b / t if t else 0:8.0f}"
        print(line)


if __name__ == "__main__":
    main()
