"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel-name totals and shares of one step.

    python tools/summarize_launches.py gpurun_out/launches.csv [first_id last_id] > profiles/rNN_launches_summary.txt

Without an id window, the last complete step is located automatically as the span between the last two launches of the
stem im2col kernel pair (the first kernels of every step)."""
from __future__ import annotations

import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^cstp::", "", name)
    name = re.sub(r"^at::native::", "at::", name)
    return name[:90]


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return [(int(r["ID"]), short(r["Kernel Name"]), float(r["Metric Value"])) for r in csv.DictReader(lines)]


def main():
    rows = load(sys.argv[1])
    if len(sys.argv) >= 4:
        lo, hi = int(sys.argv[2]), int(sys.argv[3])
    else:
        stems = [i for i, n, _ in rows if n.startswith("stem_im2col")]
        starts = [s for k, s in enumerate(stems) if k == 0 or s - stems[k - 1] > 1]
        lo, hi = starts[-2], starts[-1]
    step = [r for r in rows if lo <= r[0] < hi]
    tot = sum(t for _, _, t in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, n, t in step:
        agg[n][0] += 1
        agg[n][1] += t
    print(f"# launches [{lo},{hi}) = {len(step)} kernels, sum of kernel durations {tot / 1e6:.3f} ms "
          "(ncu-serialised, cold cache: compare shares, not absolutes)")
    print(f"{'kernel':90s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:90s} {c:8d} {t / 1e6:10.3f} {100 * t / tot:6.1f}%")


if __name__ == "__main__":
    main()
