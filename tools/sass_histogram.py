"""SASS mnemonic histogram of the built libraries (cuobjdump -sass): evidence that the tcgen05 / TMEM / TMA instructions
made it into the binaries.  python tools/sass_histogram.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "ACQBULK",
       "ELECT", "REDG", "RED.", "ATOM", "HMMA", "IMMA", "LDGSTS", "LDSM", "MUFU", "ERRBAR", "UCGABAR", "USETMAXREG")
for lib in ("libsalun.so", "libsalun_split.so"):
    path = os.path.join(ROOT, "unlearn_saliency_b200", "csrc", lib)
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    per_kernel = collections.OrderedDict()
    cur = None
    total = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_kernel[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            per_kernel[cur][op] += 1
            total[op] += 1
    print(f"==== {lib}: {len(per_kernel)} kernels, {sum(total.values())} SASS instructions")
    fam = collections.Counter()
    for op, c in total.items():
        for k in KEY:
            if op.startswith(k):
                fam[op] += c
    for op, c in sorted(fam.items(), key=lambda kv: -kv[1]):
        print(f"  {op:40s} {c}")
    print("  -- tensor-core / TMA instructions per kernel")
    for kname, cnt in per_kernel.items():
        sel = {op: c for op, c in cnt.items() if op.startswith(("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"))}
        if sel:
            short = subprocess.run(["c++filt", kname], capture_output=True, text=True).stdout.strip()[:110]
            print(f"  {short:110s} " + " ".join(f"{op}={c}" for op, c in sorted(sel.items())))
