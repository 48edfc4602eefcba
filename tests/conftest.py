import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EMUL_LIB = os.path.join(ROOT, "tests", "emul", "libbpe_emul.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    # the checker (oracle), the corpus generator and the kernel emulation are built on demand on
    # the CPU box; on the GPU box the prebuilt files travel with the snapshot
    need = [os.path.join(ROOT, "oracle", "liboracle.so"), os.path.join(ROOT, "tools", "libsynthcorpus.so"),
            os.path.join(ROOT, "zig-bpe_b200", "lib", "libbpe_b200.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session")
def zb():
    return importlib.import_module("zig-bpe_b200")


@pytest.fixture(scope="session")
def ora():
    from oracle import oracle_py
    return oracle_py


@pytest.fixture(scope="session")
def synth():
    from tools import synthcorpus
    return synthcorpus


@pytest.fixture(scope="session")
def taylor():
    return open(os.path.join(GOLDEN, "taylorswift.txt"), "rb").read()


@pytest.fixture(scope="session")
def golden_merges():
    return [tuple(int(x) for x in line.split(",")) for line in open(os.path.join(GOLDEN, "merges_300.txt"))]


@pytest.fixture(scope="session")
def emu(zb):
    """Engine over the CPU emulation of the kernels (logic tests without a GPU)."""
    if not os.path.exists(EMUL_LIB):
        import __graft_entry__ as g
        g.build()
    e = zb.Engine(lib_path=EMUL_LIB)
    e.set_option("table_log2", 13)
    return e


@pytest.fixture(scope="session")
def gpu(zb):
    """Engine on cuda:0 through the real C ABI. Fails (not skips) when the library or GPU is missing."""
    return zb.Engine(device=0)


def merges_array(m):
    import numpy as np
    return np.stack([m["first"], m["second"], m["new_token"]], axis=1) if len(m) else np.zeros((0, 3), dtype=np.uint16)
