import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


SOBOL_CASES = [
    "sobol_rastrigin_6d_256",
    "sobol_rastrigin_3d_4096",
    "sobol_cattle_4d_1024",
    "sobol_cattle_4d_4096",
    "sobol_styblinski_8d_128",
    "sobol_styblinski_3d_4096",
    "sobol_eggholder_2d_4096",
    "sobol_ackley_3d_2048",
    "sobol_rosenbrock_4d_512",
    "sobol_sphere_2d_256_cons",
    "sobol_sphere_1d_64",
]
LATTICE_CASES = [
    "lattice_unit_2d",
    "lattice_unit_3d",
    "lattice_rastrigin_4d",
    "lattice_rosen_5d",
    "lattice_ackley_3d",
    "lattice_unit_1d",
]
# pinned against the oracle on the CPU only for now (the GPU list above is what ran on a B200)
LATTICE_CASES_HIGH_D = ["lattice_unit_6d", "lattice_unit_7d"]
SIMPLICIAL_POOL_CASES = [
    "simplicial_ackley_3d_it3",
    "simplicial_rastrigin_2d_it4",
    "simplicial_styblinski_4d_it3",
    "simplicial_sphere_2d_cons_it3",
]
# recorded from the reference in 10 minutes (10-D generation 1); pinned against the oracle on the CPU
SIMPLICIAL_POOL_CASES_HIGH_D = ["simplicial_ackley_10d_it2"]
# objectives whose oracle restatement is bit-identical to the numpy evaluation
BIT_EXACT_F = {"rosenbrock", "cattle_feed", "sphere"}


def gset_of(g):
    s = str(g["gset"])
    return None if s == "None" else s


@pytest.fixture(scope="session")
def golden():
    return load_golden


def pytest_terminal_summary(terminalreporter):
    """worst observed f error per functor (tests/tolerance.py), also left in gpurun_out/ on a GPU box"""
    import tolerance
    if not tolerance.WORST:
        return
    text = tolerance.report()
    terminalreporter.write_line(text)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        try:
            import torch
            tag = "gpu" if torch.cuda.is_available() else "cpu"
            with open(os.path.join(out, "f_tolerance_%s.txt" % tag), "w") as fh:
                fh.write(text + "\n")
        except Exception:
            pass
