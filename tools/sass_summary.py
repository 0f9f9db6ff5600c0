"""Per-kernel count of the SASS mnemonics that prove tcgen05 / TMA code paths in libmsunet_sm100.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), UTMALDG / UTMASTG (TMA tensor loads / stores), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
LDGSTS (cp.async).  Usage: sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "semantic_segmentation_of_stylegan2_artifacts_b200", "libmsunet_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "LDGSTS", "SYNCS", "MUFU.TANH", "MUFU.EX2"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in keys:
        if re.search(r"\b" + re.escape(k) + r"\b", line):
            per[cur][k] += 1
            if k == "UTCHMMA.2CTA":
                break   # counted once, not again as plain UTCHMMA
demangled = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.basename(lib)}: occurrences per kernel (static SASS, not launches)")
print("kernel," + ",".join(keys))
tot = collections.Counter()
fam = collections.OrderedDict()      # kernels without tensor-core / TMA instructions (the cp.async LayerNorm templates) by family
for (name, c), dn in zip(per.items(), demangled):
    if not any(c[k] for k in keys[:7]):
        continue
    short = re.sub(r"\(.*", "", dn)
    tot.update(c)
    if any(c[k] for k in keys[:6]):
        print(short + "," + ",".join(str(c[k]) for k in keys))
    else:
        f = fam.setdefault(re.sub(r"<.*", "<...>", short), [0, collections.Counter()])
        f[0] += 1
        f[1].update(c)
for f, (n, c) in fam.items():
    print(f"{f} x{n} instantiations," + ",".join(str(c[k]) for k in keys))
print("TOTAL," + ",".join(str(tot[k]) for k in keys))
