"""Per-kernel counts of the tcgen05 / TMEM / TMA / mbarrier / packed-fp32 mnemonics in the built library's SASS.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gcn-string_b200", "libgcnstring_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
want = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTC[A-Z]+|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|SYNCS|FADD2|REDUX|ATOMS|ATOMG)\b")
print("# SASS evidence, round 2: `cuobjdump -sass gcn-string_b200/libgcnstring_b200.so` (scripts/sass_summary.py), per kernel: instruction\n"
      "# count and the tcgen05 / TMEM / TMA / mbarrier / packed-fp32 mnemonics (UTCHMMA = tcgen05.mma kind::f16/tf32, .2CTA = cta_group::2;\n"
      "# LDTM/STTM = tcgen05.ld/st; UTMALDG/UTMASTG/UTMAREDG = cp.async.bulk.tensor load/store/reduce; UBLKCP = cp.async.bulk;\n"
      "# SYNCS = mbarrier ops; FADD2 = add.f32x2; ATOMG in the slab kernel = the work-queue ticket).  Only kernels with at least one of them.\n")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for name, blk in zip(names, blocks):
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk, flags=re.M)
    c = collections.Counter()
    for i in ins:
        m = want.match(i.split(".")[0] + (".2CTA" if ".2CTA" in i and i.startswith("UTCHMMA") else ""))
        if m:
            c[m.group(1) if not i.startswith("UTCHMMA") else ("UTCHMMA.2CTA" if ".2CTA" in i else "UTCHMMA")] += 1
    if any(k not in ("REDUX", "ATOMS", "ATOMG") for k in c):
        print(name[:150])
        print(f"    instructions {len(ins)}: " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
