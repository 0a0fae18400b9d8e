#!/usr/bin/env python
"""Opcode histogram per kernel of libcng_b200.so (cuobjdump -sass), written to profiles/sass_opcodes_<tag>.txt.
Shows which kernels carry the Blackwell tensor-core / TMA instructions: UTCHMMA (tcgen05.mma kind::f16), UTCBAR (tcgen05.commit),
LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk), SYNCS (mbarrier), MUFU.SIN / COS, RED (red.global.add).
    python tools/sass_opcodes.py [tag]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "conditioned_nerf_gan_b200", "libcng_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMA", "SYNCS", "MUFU", "RED", "ATOM", "HMMA", "BAR", "MEMBAR", "FENCE", "SHFL", "DADD", "STS", "LDS", "LDG", "STG")
out = [f"# opcode histograms of {os.path.basename(lib)} (cuobjdump -sass, sm_100a), {len(kernels)} kernels", ""]
for name, c in kernels.items():
    total = sum(c.values())
    fam = collections.Counter()
    for op, n in c.items():
        for k in KEY:
            if op.startswith(k):
                fam[op if k in ("MUFU", "UTCHMMA", "LDTM", "SYNCS", "UBLKCP", "RED") else k] += n
                break
    d = demangle(name)
    d = re.sub(r"\(.*", "", d)
    out.append(f"{d}  [{total} instructions]")
    out.append("    " + ", ".join(f"{k} {v}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))
    out.append("    top: " + ", ".join(f"{k} {v}" for k, v in c.most_common(8)))
path = os.path.join(ROOT, "profiles", f"sass_opcodes_{tag}.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path, len(kernels), "kernels")
