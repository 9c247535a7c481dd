"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md), from the built library.

    python scripts/sass_summary.py [tag]      -> profiles/sass_<tag>.md
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xsdeepfwfm_deprecated_b200 import _lib
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
MN = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "SYNCS", "USETMAXREG", "LDGSTS", "LDG", "STS", "ST.E", "FFMA", "F2F", "STL", "LDL"]
sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
counts, cur, size = collections.OrderedDict(), None, collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        size[cur] += 1
        for mn in MN:
            if re.search(r"\b" + re.escape(mn) + r"\b", line) or (mn.endswith(".E") and mn in line):
                counts[cur][mn] += 1
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
out = os.path.join(ROOT, "profiles", f"sass_{tag}.md")
with open(out, "w") as f:
    f.write(f"# SASS mnemonic counts per kernel of libdeepfwfm_sm100a.so ({tag}; `cuobjdump -sass`, sm_100a)\n\n")
    f.write("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), USETMAXREG = setmaxnreg, "
            "SYNCS = mbarrier ops, LDGSTS = cp.async, STL/LDL = local-memory (spill) traffic.\n\n")
    f.write("| kernel | instructions | " + " | ".join(MN) + " |\n|---|---|" + "---|" * len(MN) + "\n")
    tot = collections.Counter()
    for (k, c), name in zip(counts.items(), demangle):
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        f.write(f"| `{short[-70:]}` | {size[k] // 1} | " + " | ".join(str(c[m]) for m in MN) + " |\n")
        tot.update(c)
    f.write(f"| **whole library** | {sum(size.values())} | " + " | ".join(str(tot[m]) for m in MN) + " |\n")
print(open(out).read()[:3000])
