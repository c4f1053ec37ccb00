#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove what the step kernels do on sm_100a (B200_PROFILING.md):
UBLKCP (cp.async.bulk, the 1-D TMA engine copies: .S.G = global->shared loads, .G.S = shared->global stores), SYNCS (mbarrier
arrive / try_wait / expect_tx), UTMA* (tensor-map TMA: none expected, the tiles are contiguous), UTC*MMA / HMMA (tensor cores:
none expected, nothing on the path is a contraction), ACQBULK / griddepcontrol (programmatic dependent launch), MEMBAR, RED / ATOM.

    python scripts/sass_ops.py > profiles/r02_sass_ops.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "marlon_b200", "libcbx.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([
    ("instructions", r"^\s+/\*[0-9a-f]{4,}\*/"), ("UBLKCP.S.G (bulk load)", r"UBLKCP\.S\.G"), ("UBLKCP.G.S (bulk store)", r"UBLKCP\.G\.S"),
    ("UBLKCP (other)", r"UBLKCP(?!\.S\.G|\.G\.S)"), ("SYNCS (mbarrier)", r"SYNCS"), ("UTMALDG/UTMASTG", r"UTMA(LDG|STG)"),
    ("UTCMMA/HMMA/IMMA (tensor core)", r"UTC\w*MMA|HMMA|IMMA|QGMMA"), ("ACQBULK / CCTL / PDL (griddep)", r"ACQBULK|PREEXIT|ACQSHMINIT"),
    ("MEMBAR", r"MEMBAR"), ("FENCE.VIEW.ASYNC", r"FENCE\.VIEW\.ASYNC"), ("RED / ATOM(S/G)", r"\b(RED|ATOMG|ATOMS|ATOM)\b"),
    ("STG (plain global store)", r"\bSTG\b"), ("LDG (plain global load)", r"\bLDG\b"), ("STS", r"\bSTS\b"), ("LDS", r"\bLDS\b")])
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print("cuobjdump -sass marlon_b200/libcbx.so: cubins for", ", ".join(arch))
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for name, pat in pats.items():
            if re.search(pat, line):
                counts[cur][name] += 1
for fn, c in counts.items():
    demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print(f"\n{demangled}")
    for name in pats:
        print(f"    {name:38s} {c.get(name, 0):7d}")
