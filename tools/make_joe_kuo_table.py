"""Regenerate shgo_b200/data/joe_kuo_dirnums.txt.

The Sobol direction numbers are the published Joe & Kuo (2008) table
"new-joe-kuo-6.21201" (free for any use, see their web page).  Both scipy
(`scipy/stats/_sobol_direction_numbers.npz`) and the reference
(`shgo/_shgo_lib/sobol_vec.gz`) ship it.  We re-export the first DMAX
dimensions from the *installed scipy's* copy in the original text format
`d s a m_1 ... m_s` so that the product does not reach into scipy privates at
run time.  tests/test_sobol_table.py cross-checks the rebuilt 30-bit direction
matrix against scipy's own `_sv`.
"""
import os
import numpy as np
import scipy.stats._sobol as _sobol

DMAX = 1024
src = os.path.join(os.path.dirname(_sobol.__file__), "_sobol_direction_numbers.npz")
z = np.load(src)
poly, vinit = z["poly"], z["vinit"]
out = os.path.join(os.path.dirname(__file__), "..", "shgo_b200", "data", "joe_kuo_dirnums.txt")
with open(out, "w") as fh:
    fh.write("# Joe-Kuo new-joe-kuo-6 direction numbers, dimensions 2..%d: d s a m_i\n" % DMAX)
    for j in range(1, DMAX):           # row 0 is dimension 1 (all m_i = 1, no polynomial)
        p = int(poly[j])
        s = p.bit_length() - 1         # polynomial degree
        a = (p >> 1) & ((1 << (s - 1)) - 1) if s > 1 else 0   # interior coefficients
        m = [int(v) for v in vinit[j, :s]]
        fh.write("%d %d %d %s\n" % (j + 1, s, a, " ".join(map(str, m))))
print("wrote", out)
