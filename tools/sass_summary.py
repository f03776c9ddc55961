#!/usr/bin/env python
"""SASS opcode histogram of one kernel of libKMCUDA.so (evidence for profiles/: which Blackwell instructions the
headline kernel is made of).  No GPU needed.

    python tools/sass_summary.py [mangled-name-substring] > profiles/r02_sass_tc_assign_4_0.txt
"""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kmcuda_b200", "libKMCUDA.so")
want = sys.argv[1] if len(sys.argv) > 1 else "tc_assign_kernelILi4ELi0"
names = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, text=True).stdout
fn = [ln.split()[1].rstrip(":") for ln in names.splitlines() if ln.strip().startswith("Function") and want in ln]
assert fn, "no kernel matches " + want
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn[0], LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
res = [ln for ln in names.splitlines() if fn[0] in ln or ln.strip().startswith("REG:")]
ops = collections.Counter()
total = 0
for ln in sass.splitlines():
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m:
        ops[m.group(2)] += 1
        total += 1
print("kernel:", fn[0])
i = names.splitlines().index([ln for ln in names.splitlines() if fn[0] in ln][0])
print("resources:", names.splitlines()[i + 1].strip())
print("static SASS instructions: %d (%.1f KB)" % (total, total * 16 / 1024))
full = collections.Counter(re.findall(r"\b(UTCHMMA[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UCGABAR_[A-Z]*|UTCATOMSWS[.A-Z0-9_]*)", sass))
print()
print("CTA-pair (cta_group::2) forms, with modifiers:")
for k in sorted(full):
    print("  %-36s %5d" % (k, full[k]))
print()
print("Blackwell-native markers (B200_PROFILING.md table):")
for key, what in [("UTCHMMA", "tcgen05.mma kind::f16"), ("UTCBAR", "tcgen05.commit"), ("LDTM", "tcgen05.ld"),
                  ("STTM", "tcgen05.st"), ("UTMALDG", "cp.async.bulk.tensor (TMA)"), ("UBLKCP", "cp.async.bulk"),
                  ("SYNCS", "mbarrier"), ("NANOSLEEP", "try_wait suspend hint"), ("FFMA2", "fma.f32x2"),
                  ("FADD2", "add/sub.f32x2"), ("FMUL2", "mul.f32x2"), ("FMNMX3", "3-input max"), ("HMMA", "legacy mma.sync"),
                  ("LDS", "ld.shared"), ("STS", "st.shared"), ("LD", "generic ld"), ("ST", "generic st")]:
    print("  %-10s %5d   %s" % (key, ops.get(key, 0), what))
print()
print("all opcodes:")
for op, n in ops.most_common():
    print("  %-14s %5d" % (op, n))
