"""Counts the Blackwell-specific SASS mnemonics per kernel of the built library (cuobjdump -sass), the evidence that the
hot kernels are tcgen05 / TMA / TMEM code and not mma.sync fallbacks.  python scripts/sass_evidence.py > profiles/<tag>_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "genpercept_b200", "libgenpercept_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "ELECT", "SYNCS", "HMMA", "MUFU.EX2", "MUFU.TANH", "MUFU.RCP"]
fn, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::", "", fn)
        fn = re.sub(r"\(.*", "", fn)
        counts.setdefault(fn, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[fn][k if k != "SYNCS" else op.split(".")[0] + "." + op.split(".")[1]] += 1
print("SASS mnemonics per kernel (cuobjdump -sass genpercept_b200/libgenpercept_b200.so, sm_100a):")
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st,")
print("SYNCS.* = mbarrier ops, ELECT = elect.sync; HMMA (mma.sync) must not appear in the GEMM / attention kernels.\n")
for fn, c in counts.items():
    if not c:
        continue
    print(f"{fn}")
    print("    " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
