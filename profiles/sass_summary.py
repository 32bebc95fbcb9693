#!/usr/bin/env python3
"""Per-kernel static SASS counts (LOP3, POPC, IADD/VIADD, IMAD, SHF, LDG, LDS/STS, UBLKCP, REDUX, atomics, branches) plus registers,
spills and static shared memory from ptxas -v.  Usage: python profiles/sass_summary.py >> profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "xmp_b200", "libxmp_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kern[cur][m.group(2)] += 1
        kern[cur]["_all"] += 1
info = {}
log = open(os.path.join(ROOT, "xmp_b200", "csrc", "ptxas.log")).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", log):
    info[m.group(1)] = (int(m.group(5)), int(m.group(2)), int(m.group(3)) + int(m.group(4)), int(m.group(7) or 0))
print("| kernel | regs | stack B | spill B | static smem B | SASS | LOP3 | POPC | IADD3+VIADD | IMAD | SHF | LDG | LDS+STS | UBLKCP | REDUX | ATOM/RED | BRA |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for k, c in kern.items():
    name = re.sub(r"\(.*", "", demangle(k)).replace("void ", "").replace("xmp::", "")
    r = info.get(k, (0, 0, 0, 0))
    at = sum(v for o, v in c.items() if o.startswith(("ATOM", "RED")) and o != "REDUX")
    print("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |" % (
        name, r[0], r[1], r[2], r[3], c["_all"], c["LOP3"], c["POPC"], c["IADD3"] + c["VIADD"], c["IMAD"], c["SHF"], c["LDG"],
        c["LDS"] + c["STS"], c["UBLKCP"], c["REDUX"], at, c["BRA"]))
