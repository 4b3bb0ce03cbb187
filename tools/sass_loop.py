"""Instruction mix of the innermost hot loop of a kernel in libafr_b200.so (static SASS).
usage: python tools/sass_loop.py <substring of mangled kernel name>"""
import re, subprocess, sys
from collections import Counter
so = "/root/repo/aliasfree-diffusion-models-pytorch_b200/libafr_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
blk = [b for b in blocks if sys.argv[1] in b.split("\n")[0]][0]
ins = []
for l in blk.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:.*?)0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a: loops.append((int(m.group(1), 16), a))
def count(lo, hi):
    c = Counter()
    for a, t in ins:
        if lo <= a <= hi:
            op = t.split()[1] if t.startswith("@") else t.split()[0]
            c[op.split(".")[0]] += 1
    return c
print(blk.split("\n")[0][:100])
for lo, hi in sorted(loops, key=lambda x: x[1] - x[0]):
    c = count(lo, hi)
    if c["MUFU"] >= 8 and (len(sys.argv) < 3 or True):
        print(f"loop {lo:#x}-{hi:#x}: {sum(c.values())} instr", c.most_common(14))
