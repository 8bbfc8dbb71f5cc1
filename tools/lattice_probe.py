import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from shgo_b200 import functors as F, hotpath as hp, lattice
d, gen = int(sys.argv[1]), int(sys.argv[2])
b = [(-32.768, 32.768)] * d
tab = torch.from_numpy(lattice.coordinate_tables(b, gen)).cuda()
nv, _ = lattice.counts(d, gen)
f, feas = hp.lattice_eval(F.F_ACKLEY, 0, d, gen, tab, 0, nv)
for _ in range(3):
    m = hp.lattice_star(d, gen, f)
torch.cuda.synchronize()
print("nv", nv, "minimisers", int(m.sum()))
