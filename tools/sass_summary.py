"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (tcgen05 -> UTC*MMA, tcgen05.ld -> LDTM,
TMA -> UTMALDG / UTMAPF, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, setmaxnreg -> USETMAXREG) in libcstp_b200.so.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt        (no GPU needed: cuobjdump -sass)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cstp_b200", "libcstp_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "USETMAXREG", "FENCE.VIEW.ASYNC",
       "STG.E.ENL2.256", "LDS.128", "STS.128", "HMMA", "BAR.SYNC"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*$", "", cur).replace("cstp::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for p in PAT:
        if re.search(r"\b" + re.escape(p), line):
            counts[cur][p] += 1
used = [p for p in PAT if any(c[p] for c in counts.values())]
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): occurrences per kernel of the instructions named in "
      "/opt/skills/guides/B200_PROFILING.md 'What proves a Blackwell-native kernel'")
print(f"{'kernel':58s} " + " ".join(f"{p[:10]:>10s}" for p in used))
tot = collections.Counter()
for k, c in counts.items():
    if any(c[p] for p in ("UTCHMMA", "LDTM", "UTMALDG", "USETMAXREG", "FENCE.VIEW.ASYNC")):
        print(f"{k[:58]:58s} " + " ".join(f"{c[p]:10d}" for p in used))
        tot.update(c)
print(f"{'TOTAL (tensor-core kernels)':58s} " + " ".join(f"{tot[p]:10d}" for p in used))
print(f"# {len(counts)} kernels in the library; the others (BatchNorm streaming, losses, optimiser, clip pipeline) use no tensor-core / TMA path")
