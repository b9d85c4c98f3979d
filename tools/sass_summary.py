#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, UTMALDG / UTMASTG = TMA loads / stores, LDTM / STTM = tcgen05.ld / st, HMMA = legacy mma.sync.
    python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt        (runs on the CPU box: cuobjdump only)"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "autognothi_b200/lib/libautognothi_b200.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "IMMA", "DFMA", "MUFU",
         "LDGSTS", "REDG", "ATOMG"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                kernels[cur][w] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {lib}: instruction counts per kernel (static code, not executed counts)")
print(f"# {'kernel':100s} {'instr':>6s}  " + " ".join(f"{w:>7s}" for w in WATCH))
rows = []
for (name, cnt), dn in zip(kernels.items(), demangled):
    short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("agb::", "")
    rows.append((short, cnt))
for short, cnt in sorted(rows, key=lambda r: (-(r[1]["UTCHMMA"] + r[1]["UTCQMMA"] + r[1]["UTCIMMA"]), -r[1]["HMMA"], r[0])):
    print(f"{short[:102]:102s} {cnt['_total']:6d}  " + " ".join(f"{cnt[w]:7d}" if cnt[w] else f"{'.':>7s}" for w in WATCH))
tc = sum(1 for _, c in rows if c["UTCHMMA"] + c["UTCQMMA"] + c["UTCIMMA"])
print(f"# {len(rows)} kernels; {tc} issue tcgen05.mma (UTC*MMA); {sum(1 for _, c in rows if c['UTMALDG'])} use TMA loads; "
      f"{sum(1 for _, c in rows if c['HMMA'])} use legacy mma.sync (HMMA)")
