"""How evenly does the star test of a lattice generation split over G contiguous id ranges?
(one GPU; times every rank's grid range and centroid range separately)"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from shgo_b200 import hotpath as hp, lattice, functors as F
from shgo_b200.dist import split_range
d, gen, G = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
b = [(-32.768, 32.768)] * d
dev = torch.device("cuda")
tab = torch.from_numpy(lattice.coordinate_tables(b, gen)).to(dev)
nv, ngrid = lattice.counts(d, gen)
f, feas = hp.lattice_eval(F.F_ACKLEY, 0, d, gen, tab, 0, nv)
def t(fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / k
whole = t(lambda: hp.lattice_star(d, gen, f))
ev = t(lambda: hp.lattice_eval(F.F_ACKLEY, 0, d, gen, tab, 0, nv, f=f, feasible=feas))
print("whole generation: eval %.3f ms  star %.3f ms  (%d vertices)" % (ev, whole, nv))
tot = []
for r in range(G):
    g0, g1 = split_range(ngrid, G, r)
    c0, c1 = split_range(nv - ngrid, G, r)
    a = t(lambda: hp.lattice_star(d, gen, f, g0, g1 - g0))
    c = t(lambda: hp.lattice_star(d, gen, f, ngrid + c0, c1 - c0))
    e = t(lambda: (hp.lattice_eval(F.F_ACKLEY, 0, d, gen, tab, g0, g1 - g0, f=f[g0:g1], feasible=feas[g0:g1]),
                   hp.lattice_eval(F.F_ACKLEY, 0, d, gen, tab, ngrid + c0, c1 - c0, f=f[ngrid + c0:ngrid + c1],
                                   feasible=feas[ngrid + c0:ngrid + c1])))
    tot.append(a + c + e)
    print("rank %d: eval %.3f  star grid %.3f  star centroids %.3f  sum %.3f ms" % (r, e, a, c, a + c + e))
print("max %.3f  mean %.3f  -> balance %.0f %%" % (max(tot), sum(tot) / G, 100 * sum(tot) / G / max(tot)))
