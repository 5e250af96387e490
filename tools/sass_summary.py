"""Per-kernel SASS opcode histogram of libfidm_b200.so (cuobjdump -sass): the Blackwell-native instructions each
kernel actually contains (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier).  Usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "face-inpainting-diffusion-models_b200", "libfidm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS",
        "MUFU", "HMMA", "FFMA", "LDS", "STS", "LDG", "STG", "BAR", "MEMBAR", "FENCE", "F2FP", "F2F", "SHFL", "STL", "LDL")
name, hist, total = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("fidm::", "")
        hist[name], total[name] = collections.Counter(), 0
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        total[name] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[name][op if k.startswith("UTC") or k in ("LDTM", "UTMALDG", "UTMASTG", "MUFU") else k] += 1
                break
print(f"# SASS opcode histogram per kernel: {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a)")
print("# tcgen05.mma -> UTCHMMA (kind::f16) / UTCQMMA (kind::f8f6f4), .2CTA = cta_group::2; tcgen05.ld -> LDTM; TMA -> UTMALDG / UTMASTG")
for k in sorted(hist, key=lambda k: -total[k]):
    if total[k] == 0:
        continue
    items = ", ".join(f"{op} {n}" for op, n in sorted(hist[k].items(), key=lambda kv: (-kv[1], kv[0])))
    print(f"{k[:110]}\n    {total[k]} instructions: {items}")
grand = collections.Counter()
for h in hist.values():
    grand.update(h)
print("# whole library:", ", ".join(f"{op} {n}" for op, n in sorted(grand.items(), key=lambda kv: -kv[1]) if op.startswith(("UTC", "LDTM", "UTMA"))))
