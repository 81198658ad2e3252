"""Per-launch table of an `ncu --set full` report (read with `ncu -i <rep> --page raw --csv`): duration, tensor-pipe activity,
DRAM / L2 / shared-memory throughput, DRAM bytes, issue-slot use -- one row per launch plus per-kernel-name averages.

    ncu -i gpurun_out/ev2/ncu_full_step_b60.ncu-rep --page raw --csv | python tools/summarize_ncu_full.py > profiles/rNN_...txt
"""
import collections
import csv
import re
import sys

COLS = [("ms", "gpu__time_duration.sum", 1e-6),                                   # ns -> ms
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("l1/smem%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("issue%", "sm__inst_executed.avg.per_cycle_elapsed", 25.0),             # inst/cycle of 4 -> %
        ("rd MB", "dram__bytes_read.sum", 1e-6),
        ("wr MB", "dram__bytes_write.sum", 1e-6)]
UNIT = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
        "Gbyte": 1e9, "%": 1.0, "inst/cycle": 1.0}


def short(name):
    name = re.sub(r"^void ", "", name).replace("cstp::", "")
    m = re.match(r"([A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")).replace("(bool)", "")[:44] if m else name[:44]


rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
print(f"{'id':>5s} {'kernel':44s} " + " ".join(f"{c[0]:>9s}" for c in COLS) + "  grid x block")
agg = collections.OrderedDict()
for r in rows[2:]:
    vals = []
    for _, metric, scale in COLS:
        if metric not in ix or r[ix[metric]] in ("", "n/a"):
            vals.append(float("nan"))
            continue
        v = float(r[ix[metric]].replace(",", "")) * UNIT.get(units[ix[metric]], 1.0) * scale
        vals.append(v)
    name = short(r[ix["Kernel Name"]])
    print(f"{r[ix['ID']]:>5s} {name:44s} " + " ".join(f"{v:9.3f}" for v in vals) +
          f"  {r[ix['launch__grid_size']]} x {r[ix['launch__block_size']]}")
    a = agg.setdefault(name, [0] + [0.0] * len(COLS))
    a[0] += 1
    for i, v in enumerate(vals):
        a[i + 1] += v
print()
print("# per kernel name: launches; ms and MB are sums over the launches, the percentages plain averages")
print(f"{'kernel':44s} {'n':>4s} " + " ".join(f"{c[0]:>9s}" for c in COLS))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    n = a[0]
    out = []
    for i, (label, _, _) in enumerate(COLS):
        out.append(a[i + 1] if label in ("ms", "rd MB", "wr MB") else a[i + 1] / n)
    print(f"{name:44s} {n:4d} " + " ".join(f"{v:9.3f}" for v in out))
