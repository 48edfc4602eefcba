"""The emulated kernels must not depend on the order in which the threads of a block (or the blocks of a grid) run
between barriers: tests/emul permutes both under EMUL_ORDER=reverse|random. A cheap stand-in, on the CPU, for the
concurrency of a real GPU (it caught nothing in the kernels, but it pins the emulator's __syncwarp semantics)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("order", ["reverse", "random"])
def test_emulated_kernels_are_order_independent(order, emu):  # `emu` makes sure the emulation library is built
    env = dict(os.environ, EMUL_ORDER=order)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_emul.py"), "-x", "-q", "-p", "no:cacheprovider",
                        "-k", "encode or train_small or golden_prefix or resident or queueless"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
