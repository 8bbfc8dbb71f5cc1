"""Near-tie diagnosis (acceptance criterion: "any minimiser-pool differences are traced to documented
exact near-ties").

Objectives that go through cos / sin / exp differ from numpy's evaluation by a few ulp (CUDA
libdevice vs glibc / numpy SIMD kernels).  The strict star test f[v] < f[w] can then flip on an edge
whose two values are equal to within those few ulp.  `explain_pool_difference` takes the two f
vectors and the graph and, for every vertex classified differently, lists the deciding edges and how
close their values are -- so that a pool difference is either explained by a near-tie or flagged.
Host-side numpy utility; it is a report, not part of the classification.
"""
import numpy as np


def ulp_distance(a, b):
    """|a - b| measured in units in the last place of the larger magnitude (inf -> inf)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b) / np.spacing(np.maximum(np.abs(a), np.abs(b)))
    d = np.where(np.isinf(a) & np.isinf(b) & (a == b), 0.0, d)
    return d


def explain_pool_difference(row_ptr, col, f_a, f_b, is_min_a, is_min_b, max_ulps=8.0):
    """Returns (explained, report).  report: one dict per vertex whose flag differs, with the edges
    (v, w) on which the two evaluations order the end points differently and their ulp distances.
    explained is True when every differing vertex has at least one such edge and all of them are
    within max_ulps in BOTH evaluations."""
    row_ptr, col = np.asarray(row_ptr), np.asarray(col)
    f_a, f_b = np.asarray(f_a, np.float64), np.asarray(f_b, np.float64)
    diff = np.flatnonzero(np.asarray(is_min_a, bool) != np.asarray(is_min_b, bool))
    report, explained = [], True
    for v in diff.tolist():
        nb = col[row_ptr[v]:row_ptr[v + 1]]
        flip = nb[(f_a[v] < f_a[nb]) != (f_b[v] < f_b[nb])]
        ua = ulp_distance(f_a[v], f_a[flip])
        ub = ulp_distance(f_b[v], f_b[flip])
        ok = len(flip) > 0 and bool(np.all(ua <= max_ulps) and np.all(ub <= max_ulps))
        explained = explained and ok
        report.append({"vertex": v, "edges": [(v, int(w)) for w in flip.tolist()],
                       "ulps_a": ua.tolist(), "ulps_b": ub.tolist(), "near_tie": ok})
    return explained, report
