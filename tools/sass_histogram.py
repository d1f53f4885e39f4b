"""SASS opcode histogram of libgim_b200.so (cuobjdump -sass), per kernel family: proves which kernels carry tcgen05 / TMA / TMEM / cluster
instructions.  python tools/sass_histogram.py > profiles/sass_histogram_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "optimalstrategiesagainstgenerativeattacks_b200", "libgim_b200.so")
KEY = re.compile(r"^(UTCHMMA|UTCQMMA|UTCBAR|UTCCP|UTMALDG|UTMASTG|UTMAPF|UTMACCTL|UTMACMDFLUSH|LDTM|STTM|UTCATOMSWS|SYNCS|UCGABAR|HMMA|LDSM|LDGSTS|RED|ATOM|REDG|ATOMG|MUFU|BAR|LDG|STG|LDS|STS|FFMA|HFMA2|CCTL|ELECT|UBLKCP|UBLKPF|CS2R|ACQBULK|ENDCOLLECTIVE|FENCE)")

out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        if KEY.match(op):
            cur[op] += 1
print("SASS opcode histogram of libgim_b200.so (sm_100a), `cuobjdump -sass`; selected opcodes per kernel (full mnemonic with modifiers)\n")
tot = collections.Counter()
for name, c in per.items():
    t5 = {k: v for k, v in c.items() if k.startswith(("UTC", "UTMA", "LDTM", "STTM", "UBLK"))}
    if not t5 and not any(k.startswith(("HMMA", "LDSM")) for k in c):
        continue
    print("%s   (%d instructions)" % (name[:150], c["_total"]))
    for k, v in sorted(c.items()):
        if k != "_total" and (k.startswith(("UTC", "UTMA", "LDTM", "STTM", "UBLK", "HMMA", "LDSM", "SYNCS", "UCGABAR", "RED", "ELECT", "LDGSTS", "FENCE"))):
            print("    %-46s %5d" % (k, v))
            tot[k] += v
    print()
print("whole library, tcgen05 / TMA / TMEM / mma.sync families:")
for k, v in sorted(tot.items()):
    print("    %-46s %5d" % (k, v))
