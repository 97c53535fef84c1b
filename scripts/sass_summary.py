"""profiles/r02_sass_summary.csv: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use
(cuobjdump -sass of the built library; B200_PROFILING.md names the mnemonics).  python scripts/sass_summary.py [out.csv]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "FFMA", "HMMA", "MUFU"]
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "flowcompare_b200", "libflowcompare_b200.so")], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur:
        for op in OPS:
            if re.search(r"\b" + op + r"\b", line):
                counts[cur][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
rows, tot = [], collections.Counter()
for (k, c), n in zip(counts.items(), names):
    if sum(c.values()):
        n = re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", "").replace("void ", ""))
        rows.append('"' + n + '",' + ",".join(str(c[o]) for o in OPS)); tot.update(c)
out = ["# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, SYNCS = mbarrier ops",
       "kernel," + ",".join(OPS)] + sorted(rows) + ["TOTAL," + ",".join(str(tot[o]) for o in OPS)]
open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.csv"), "w").write("\n".join(out) + "\n")
print(out[-1])
