import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import pointops as fpo
from oracle import pointops_refcuda as ref, port_paconv
g = torch.Generator().manual_seed(0)
x = (torch.rand(2, 1250, 3, generator=g) * 2 - 1).cuda()
q = x[:, :312].contiguous()
a, ad = fpo.knnquery_heap(32, x, q, return_dist2=True)
b, bd = ref.knnquery_heap(32, x, q)
c = port_paconv.knnquery_heap(32, x.cpu(), q.cpu())
print("knn ours vs ref idx equal", torch.equal(a, b), "mismatch frac", (a != b).float().mean().item(), "ours vs cpu", torch.equal(a.cpu(), c), "ref vs cpu", torch.equal(b.cpu(), c))
print(" ours[0,0,:8]", a[0, 0, :8].tolist(), "ref", b[0, 0, :8].tolist(), "d2 ours", ad[0, 0, :4].tolist(), "ref", bd[0, 0, :4].tolist())
d, i = fpo.nearestneighbor(x, q)
rd2, ri = ref.nearestneighbor(x, q)
cd, ci = port_paconv.nearestneighbor(x.cpu(), q.cpu())
print("3nn ours vs ref idx", torch.equal(i, ri), (i != ri).float().mean().item(), "ours vs cpu", torch.equal(i.cpu(), ci), "ref vs cpu", torch.equal(ri.cpu(), ci))
print(" ours", i[0, 5].tolist(), d[0, 5].tolist(), "ref", ri[0, 5].tolist(), rd2[0, 5].sqrt().tolist(), "cpu", ci[0, 5].tolist(), cd[0, 5].tolist())
xd = x.clone(); xd[:, 625:] = xd[:, :625]
f1 = fpo.furthestsampling(xd, 312); f2 = ref.furthestsampling(xd, 312); f3 = port_paconv.furthestsampling(xd.cpu(), 312)
print("fps dup ours vs ref", torch.equal(f1, f2), "first mismatch", (f1 != f2).nonzero()[:3].tolist(), "ours vs cpu", torch.equal(f1.cpu(), f3), "ref vs cpu", torch.equal(f2.cpu(), f3))
if not torch.equal(f1, f2):
    j = (f1 != f2).nonzero()[0]
    print("  at", j.tolist(), "ours", f1[j[0], j[1]].item(), "ref", f2[j[0], j[1]].item())
print("---- fps determinism / tie rule")
g = torch.Generator().manual_seed(1250 + 312)
xx = torch.rand(2, 1250, 3, generator=g) * 2 - 1
xx[:, 625:] = xx[:, :625]
xx = xx.contiguous()
w = port_paconv.furthestsampling(xx, 312)
xc = xx.cuda()
for rep in range(3):
    a = fpo.furthestsampling(xc, 312).cpu(); b = ref.furthestsampling(xc, 312).cpu()
    print(rep, "ours==cpu", torch.equal(a, w), "ref==cpu", torch.equal(b, w), "ours[0,:8]", a[0, :8].tolist(), "ref[0,:8]", b[0, :8].tolist(), "cpu", w[0, :8].tolist())
a1 = fpo.furthestsampling(xc[:1].contiguous(), 312).cpu()
print("B=1 ours==cpu", torch.equal(a1, w[:1]))
for tb in (1024, 512, 2048, -1):
    a = fpo.furthestsampling(xc, 312, tie_block=tb).cpu()
    print("tie_block", tb, a[0, :8].tolist())
