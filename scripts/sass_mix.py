#!/usr/bin/env python
"""Static instruction mix of the kernels in an object file (cuobjdump -sass), to judge a kernel
change on the CPU before spending GPU time: per function the opcode histogram, the FP64 /
other split and 2 x FP64 + other, the issue-port cycles that bound kernel 2 on B200 (an FP64
instruction holds its scheduler's port for two cycles). Static counts include fall-back code
that rarely runs and count loop bodies once -- compare variants, do not read them as times.

    python scripts/sass_mix.py [object=perturb_b200/csrc/sgp4_prop.o] [name regex=ILi0ELi0]
"""
import collections
import re
import subprocess
import sys

obj = sys.argv[1] if len(sys.argv) > 1 else "perturb_b200/csrc/sgp4_prop.o"
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else "ILi0ELi0")
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")
name, hist = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        hist[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and name:
        ops = m.group(1).split()
        op = ops[1] if ops[0].startswith("@") and len(ops) > 1 else ops[0]
        hist[name][op.split(".")[0]] += 1
for fn, c in hist.items():
    if not pat.search(fn):
        continue
    total = sum(c.values())
    f64 = sum(c[k] for k in FP64)
    print("%s\n  %d instructions: %d FP64 (incl. DSETP) + %d other -> 2 x FP64 + other = %d"
          % (fn, total, f64, total - f64, 2 * f64 + total - f64))
    print("  " + ", ".join("%s %d" % kv for kv in c.most_common(18)))
