"""The driver package's OWN shgo test-suite with DeviceComplex injected at the SHGO.HC seam.

scipy ships a vendored near-copy of the reference's driver and of its test module
(scipy/optimize/tests/test__shgo.py = reference shgo/tests/test__shgo.py with scipy imports); it is
present on the GPU box, the reference checkout is not.  Passing it unchanged is the drop-in claim.
(In the build container the reference's own module was run the same way: 68 passed, 3 skipped.)
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(fake):
    import scipy.optimize.tests.test__shgo as t
    env = dict(os.environ, SHGO_B200_DRIVER="scipy.optimize._shgo", PYTHONPATH=HERE + os.pathsep +
               os.environ.get("PYTHONPATH", ""))
    env["SHGO_B200_FAKE"] = "1" if fake else "0"
    r = subprocess.run([sys.executable, "-m", "pytest", t.__file__, "-q", "-x", "-p", "driver_suite_plugin",
                        "-p", "no:cacheprovider", "-W", "ignore::DeprecationWarning",
                        "-W", "ignore::pytest.PytestUnknownMarkWarning"],
                       env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


def test_scipy_shgo_suite_host_logic_on_cpu():
    _run(fake=True)


@pytest.mark.gpu
def test_scipy_shgo_suite_on_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _run(fake=False)
