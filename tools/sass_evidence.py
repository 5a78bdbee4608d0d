"""Per-kernel SASS mnemonic counts of the built library objects (cuobjdump -sass clipbert_b200/lib/obj/*.o): which kernels issue
tcgen05 / TMA / TMEM / mma.sync / multimem instructions. Usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import glob
import os
import re
import subprocess
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMACCTL", "UTMACMDFLUSH", "SYNCS", "HMMA", "LDSM", "LDGMC", "REDG", "STG", "LDG"]
print("SASS evidence (cuobjdump -sass of clipbert_b200/lib/obj/*.o, nvcc 12.9 -gencode arch=compute_100a,code=sm_100a), generated on the build host.")
print("Mnemonics: UTCHMMA = tcgen05.mma; UTCBAR = tcgen05.commit; LDTM = tcgen05.ld (TMEM -> registers); UTCATOMSWS = tcgen05 alloc/dealloc;")
print("UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA load / store); UBLKCP = cp.async.bulk (1-D bulk copy: shift vectors); UTMACCTL = prefetch.tensormap;")
print("SYNCS = mbarrier ops; HMMA + LDSM = mma.sync + ldmatrix (attention); LDGMC = multimem.ld_reduce (NVLS all-reduce); REDG = red.global.add (wgrad).\n")
for obj in sorted(glob.glob(os.path.join(ROOT, "clipbert_b200", "lib", "obj", "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    per, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and m.group(1) in KEYS:
            per[cur][m.group(1)] += 1
    tot = Counter()
    for c in per.values():
        tot.update(c)
    print("%s: %d kernels | %s" % (os.path.basename(obj), len(per), "  ".join("%s %d" % (k, tot[k]) for k in KEYS if tot[k])))
    for fn, c in per.items():
        if any(c[k] for k in ("UTCHMMA", "HMMA", "LDGMC", "UTMALDG", "UBLKCP")):
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            print("    %-72s %s" % (name[:72], "  ".join("%s %d" % (k, c[k]) for k in KEYS[:13] if c[k])))
