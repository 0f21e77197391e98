"""Developer tool (CPU): histogram of the Blackwell-specific SASS opcodes per kernel of the shipped library
(`cuobjdump -sass mtl_ssl_b200/libmtlssl.so`): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG /
UTMAREDG (TMA tensor loads / stores / reduce-adds), UTCBAR (tcgen05.commit), SYNCS (mbarrier), LDGSTS (cp.async), HMMA
(legacy tensor path: must be absent).  usage: sass_opcodes.py [lib.so] > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mtl_ssl_b200", "libmtlssl.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG(?:\.[0-9A-Z.]+)?|UTMASTG(?:\.[0-9A-Z.]+)?|UTMAREDG(?:\.[0-9A-Z.]+)?|"
                 r"UTMAPF|UTCBAR|UTCCP|SYNCS\.[A-Z.0-9]+|LDGSTS(?:\.[A-Z.0-9]+)?|UBLKCP|HMMA|HGMMA|QGMMA|IGMMA|ELECT|"
                 r"REDUX|SHFL\.[A-Z]+|VOTE\.[A-Z]+|ATOMG?\.[A-Z.0-9]+|RED\.[A-Z.0-9]+)\b")
per = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "")
        per[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = PAT.search(line)
    if m:
        op = m.group(1)
        op = re.sub(r"^(SYNCS\.[A-Z]+).*", r"\1", op)
        op = re.sub(r"^(ATOMG?|RED)\.([A-Z.]*?)(\.[0-9A-Z]+)*$", r"\1", op)
        per[name][op] += 1
tot = collections.Counter()
print("SASS opcode histogram of %s (sm_100a), `python tools/sass_opcodes.py`\n" % os.path.relpath(lib, ROOT))
for k, c in per.items():
    if not c:
        continue
    tot.update(c)
    print("%s\n    %s" % (k, "  ".join("%s %d" % kv for kv in sorted(c.items()))))
print("\nTOTAL\n    %s" % "  ".join("%s %d" % kv for kv in sorted(tot.items())))
print("\nlegacy tensor-core opcodes (HMMA / HGMMA / QGMMA / IGMMA): %d" % sum(tot[k] for k in ("HMMA", "HGMMA", "QGMMA", "IGMMA")))
