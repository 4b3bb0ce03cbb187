#!/usr/bin/env python
"""SASS / resource evidence of the SHIPPED library, regenerated from the .so itself.
usage: python tools/sass_evidence.py [out.md]      (default: profiles/r02_sass_evidence.md)"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "aliasfree-diffusion-models-pytorch_b200", "libafr_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_evidence.md")

sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
MNEMONICS = [
    ("UTMALDG.3D", "TMA cp.async.bulk.tensor.3d (haloed tile -> shared)"),
    ("SYNCS.ARRIVE.TRANS64", "mbarrier arrive.expect_tx"),
    ("SYNCS.PHASECHK.TRANS64.TRYWAIT", "mbarrier try_wait.parity"),
    ("FFMA2", "packed fma.rn.f32x2 (new on sm_100)"),
    ("FMUL2", "packed mul.rn.f32x2"),
    ("FADD2", "packed add.rn.f32x2"),
    ("MUFU.EX2", "ex2.approx (one per GELU / GELU')"),
    ("SHFL.UP", "column-0 exchange between strips; row / column neighbours in the warp-shuffle plane kernels"),
    ("SHFL.DOWN", "row below / right neighbour in the warp-shuffle plane kernels"),
    ("STG.E.EF.128", "128-bit streaming stores"),
    ("STG.E.EF.64", "64-bit streaming stores (down3_warp_kernel: 2 outputs per lane, contiguous across the warp)"),
    ("LDG.E.128.CONSTANT", "128-bit read-only loads"),
    ("LDS.128", "128-bit shared loads of the staged tile"),
    ("CCTL", "prefetch.global"),
    ("HMMA", "legacy tensor path (none expected)"),
    ("UTCHMMA", "tcgen05 (none: depthwise stencil)"),
]
lines = sass.split("\n")
count = collections.Counter()
for l in lines:
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        op = m.group(1)
        for key, _ in MNEMONICS:
            if op == key or op.startswith(key + ".") or (key in ("FFMA2", "FMUL2", "FADD2", "HMMA", "UTCHMMA", "CCTL") and op.split(".")[0] == key):
                count[key] += 1

fam = collections.defaultdict(list)
cur = None
for l in res.split("\n"):
    m = re.search(r"Function (\S+):", l)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", l)
    if m and cur:
        hits = re.findall(r"\d+((?:fgelu|up|down|gelu|rotate|ddpm|groupnorm|affine)[a-z0-9_]*_kernel)", cur)
        base = hits[-1] if hits else cur
        fam[base].append(int(m.group(1)))
        cur = None

with open(out, "w") as f:
    f.write("# r02 SASS evidence (cuobjdump -sass of the shipped libafr_b200.so)\n\n")
    f.write("Regenerate: `python tools/sass_evidence.py`.  Library sha256[:16] = `%s`, arch in the image: %s; built with\n"
            "`nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -cudart shared` (nvcc 12.9).\n\n"
            % (hashlib.sha256(open(SO, "rb").read()).hexdigest()[:16], ", ".join(arch)))
    f.write("| mnemonic | occurrences | meaning |\n|---|---|---|\n")
    for key, what in MNEMONICS:
        f.write(f"| `{key}` | {count[key]} | {what} |\n")
    f.write("\nRegisters per kernel family (min-max over the template instantiations, `cuobjdump -res-usage`):\n\n")
    f.write("| kernel family | instantiations | registers/thread |\n|---|---|---|\n")
    for base in sorted(fam):
        r = fam[base]
        f.write(f"| `{base}` | {len(r)} | {min(r)}" + (f"-{max(r)}" if max(r) != min(r) else "") + " |\n")
print(open(out).read())
