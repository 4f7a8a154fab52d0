import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built libraries (they are git-ignored): build them once, like the driver's build()
    need = [os.path.join(ROOT, "cpppathtracer_b200", "libpt_b200.so"), os.path.join(ROOT, "cpppathtracer_b200", "libpt_b200_fma.so"), os.path.join(ROOT, "cpppathtracer_b200", "libpt_b200_grid.so"),
            os.path.join(ROOT, "oracle", "libpt_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def reference_sky():
    """RGBA8 1280x1280 whose left quarter is the reference's sky.png (the only part its upload keeps)."""
    g = golden("sky_quarter.npz")
    w = int(g["full_width"])
    q = g["rgb"]
    sky = np.zeros((q.shape[0], w, 4), np.uint8)
    sky[:, : q.shape[1], :3] = q
    sky[..., 3] = 255
    return sky


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import ref_host
    if not ref_host.port_available():
        import __graft_entry__ as g
        g.build()
    return ref_host.PortOracle()


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import ref_host
    if not ref_host.available():
        pytest.skip("oracle/_ref/libpt_ref_host.so not built (needs /root/reference)")
    return ref_host.RefHost()


@pytest.fixture(scope="session")
def sky():
    return reference_sky()
