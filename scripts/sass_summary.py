"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the built library uses (tcgen05 MMA = UTCHMMA,
TMA tensor loads = UTMALDG, bulk async copies = UBLKCP, TMEM loads = LDTM, packed f32x2 = FFMA2 / FMUL2 / FADD2, ...).

    python scripts/sass_summary.py [lib] > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mini_mcmc_b200", "libminimcmc.so")
WATCH = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "DFMA", "HMMA", "SHFL", "LDS", "STS",
         "LDG", "STG", "ATOM", "RED", "MUFU", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, total, name = collections.OrderedDict(), {}, None
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("mmc::", "").replace("(anonymous namespace)::", "")
        counts[name] = collections.Counter()
        total[name] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        total[name] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR") and op.startswith(w)):
                counts[name][w] += 1
                break
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} (sm_100a): instructions per kernel and counts of the watched mnemonics")
print(f"# {'kernel':100s} {'instr':>7s}  " + " ".join(f"{w}" for w in WATCH))
for k, c in counts.items():
    if total[k] < 50:
        continue
    print(f"{k[:100]:100s} {total[k]:7d}  " + " ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
