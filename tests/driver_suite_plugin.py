"""pytest plugin: run a DRIVER package's own shgo test module with DeviceComplex injected at the
SHGO.HC seam (python -m pytest <path to test__shgo.py> -p driver_suite_plugin).

    SHGO_B200_DRIVER   driver module to patch (default scipy.optimize._shgo; the reference's is
                       shgo._shgo)
    SHGO_B200_FAKE=1   use the oracle-backed CPU stand-in for the kernels (no GPU container)
"""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


@pytest.fixture(autouse=True)
def _device_complex(monkeypatch):
    if os.environ.get("SHGO_B200_FAKE") == "1":
        import fake_hotpath
        fake_hotpath.install(monkeypatch)
    from shgo_b200 import api
    drv = os.environ.get("SHGO_B200_DRIVER", "scipy.optimize._shgo")
    api.install(drv)
    yield
    api.uninstall(drv)
