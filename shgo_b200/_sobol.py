"""Host side of the Sobol generator: direction numbers.

The reference draws its points from scipy's qmc.Sobol(d, scramble=False, seed=0)
(/root/reference/shgo/_shgo.py:703-704), i.e. the 30-bit Joe-Kuo "new-joe-kuo-6.21201" generator.
The table itself is public data; data/joe_kuo_dirnums.txt holds its first 1024 dimensions in the
original `d s a m_i` format (identical to the reference's shgo/_shgo_lib/sobol_vec.gz rows; see
tools/make_joe_kuo_table.py).  direction_numbers(d) expands it into the [d][30] uint32 matrix the
kernels consume (scipy calls it `_sv`); tests/test_host.py checks it against scipy's own.
"""
import functools
import os

import numpy as np

BITS = 30
MAX_DIM_TABLE = 1024
_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "joe_kuo_dirnums.txt")


@functools.lru_cache(maxsize=1)
def _rows():
    rows = {}
    with open(_TABLE) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            t = [int(v) for v in line.split()]
            rows[t[0]] = (t[1], t[2], t[3:3 + t[1]])
    return rows


@functools.lru_cache(maxsize=64)
def direction_numbers(d):
    """[d][30] uint32 direction matrix of the unscrambled Sobol sequence."""
    if not 1 <= d <= MAX_DIM_TABLE:
        raise ValueError("Sobol dimension %d outside 1..%d" % (d, MAX_DIM_TABLE))
    sv = np.zeros((d, BITS), np.uint32)
    sv[0] = [1 << (BITS - 1 - i) for i in range(BITS)]
    rows = _rows()
    for j in range(1, d):
        s, a, m = rows[j + 1]
        v = [0] * BITS
        for i in range(min(s, BITS)):
            v[i] = m[i] << (BITS - 1 - i)
        for i in range(s, BITS):
            x = v[i - s] ^ (v[i - s] >> s)
            for k in range(1, s):
                if (a >> (s - 1 - k)) & 1:
                    x ^= v[i - k]
            v[i] = x
        sv[j] = v
    sv.setflags(write=False)
    return sv
