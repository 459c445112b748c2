"""SASS excerpts of the built library (no GPU needed): what backs the instruction-mix claims of DESIGN.md.
usage: python profiles/sass_excerpts.py > profiles/r2_sass_excerpts.txt"""
import collections
import re
import subprocess
import sys

SO = sys.argv[1] if len(sys.argv) > 1 else "jpeglibrary_b200/lib/libjpegb200.so"
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        funcs[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())


def demangle(n):
    return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n


def mix(lines):
    c = collections.Counter()
    for ln in lines:
        t = ln.split("*/", 1)[1].split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op.split(".")[0].rstrip(";")] += 1
    return c


def pick(sub):
    return [(n, l) for n, l in funcs.items() if sub in n]


print("# SASS excerpts of", SO, "(cuobjdump -sass; sm_100a)\n")
print("## kernels and their static instruction counts")
for n, l in funcs.items():
    print("%6d  %s" % (len(l), demangle(n)[:150]))

name, lines = [x for x in pick("jb_k2_idct_color_warpILi0ELi3ELi2ELi2E")][0]
c = mix(lines)
print("\n## K2 fast renderer, RGB24 4:2:0 instance:", demangle(name)[:120])
print("packed fp32: FADD2 %d, FFMA2 %d, FMUL2 %d; scalar FADD %d, FMUL %d, FFMA %d; I2F %d; VIADDMNMX (DPX clamp) %d"
      % (c["FADD2"], c["FFMA2"], c["FMUL2"], c["FADD"], c["FMUL"], c["FFMA"], c["I2F"], c["VIADDMNMX"]))
print("(one 2-D IDCT = 16 1-D transforms of 32 additions + 12 products, two transforms per packed instruction: 8 x 32 = 256 FADD2,")
print(" 8 x 12 = 96 FFMA2 as products (addend: the -0.0 pair) + 32 FFMA2 dequantisation + 32 FFMA2 scale / round / level shift = 160)")
print("TMA tensor store / bulk copy / cp.async lines:")
for ln in lines:
    if re.search(r"UTMASTG|UBLKCP|UTMACMDFLUSH|LDGSTS|UTMALDG", ln):
        print(ln)
print("first packed-arithmetic lines (jb_idct8x2: FFMA2 with the -0.0 pair in a register as the product, FADD2 as the sums):")
shown = 0
for ln in lines:
    if re.search(r"FADD2|FFMA2", ln):
        print(ln)
        shown += 1
        if shown >= 24:
            break

name, lines = [x for x in pick("jb_k0_restart_scan")][0]
print("\n## K0 restart-marker index: TMA bulk loads of the compressed bytes (cp.async.bulk + mbarrier)")
for ln in lines:
    if re.search(r"UBLKCP|SYNCS|UTMA", ln):
        print(ln)

name, lines = [x for x in pick("jb_k1_huff_flatILb0E")][0]
c = mix(lines)
print("\n## K1 flat segment decoder (restart segments):", demangle(name)[:100])
print("static mix:", ", ".join("%s %d" % kv for kv in c.most_common(14)))
# the decode round: from the loop head (first VOTE.ANY after the table set-up barriers) to the backward branch
idx = [i for i, ln in enumerate(lines) if "VOTE.ANY" in ln]
bars = [i for i, ln in enumerate(lines) if "BAR.SYNC" in ln]
start = next(i for i in idx if i > bars[-1])
end = idx[-1] + 2
print("decode round, SASS lines %d..%d of %d (refill with fused un-stuffing, up to three symbol steps, ranked block hand-off,"
      " block change):" % (start, end, len(lines)))
for ln in lines[start:end]:
    print(ln)
