"""Per-shape breakdown of the tcgen05 GEMM launches of one instrumented step (file written by FC_PROFILE_DUMP=<path> python bench.py)."""
import sys
from collections import defaultdict
d = defaultdict(lambda: [0, 0.0, 0.0])
for l in open(sys.argv[1]):
    c, tag, fl, ms = l.split(); c = int(c); tag = int(tag)
    if c != 1: continue
    k = (tag & 0xffff, (tag >> 16) & 0xffff, (tag >> 32) & 0xf, (tag >> 36) & 0xf, (tag >> 40) & 1)
    d[k][0] += 1; d[k][1] += float(ms); d[k][2] += float(fl)
tot = sum(v[1] for v in d.values())
print(f"tcgen05 GEMM launches: {sum(v[0] for v in d.values())}, {tot:.2f} ms")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"K={k[0]:4d} N={k[1]:4d} epi={k[2]} act={k[3]} res={k[4]}  n={v[0]:4d}  ms={v[1]:7.2f} ({v[1]/tot*100:4.1f}%)  {v[2]/v[1]/1e9:6.1f} TF/s  avg {v[1]/v[0]*1e3:6.1f} us")
