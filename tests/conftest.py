import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "rgbd-recon_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def bits_equal(a, b):
    """Bit-exact float comparison that treats NaN == NaN (payload-insensitive) and +0 == +0 only."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    an, bn = np.isnan(a), np.isnan(b)
    return (an == bn) & (an | (a.view(np.uint32) == b.view(np.uint32)))


def mismatch_report(name, got, want, limit=5):
    eq = bits_equal(got, want)
    bad = np.argwhere(~eq)
    lines = [f"{name}: {len(bad)} of {eq.size} values differ"]
    for idx in bad[:limit]:
        t = tuple(idx)
        lines.append(f"  at {t}: got {got[t]!r} want {want[t]!r}")
    return "\n".join(lines)


@pytest.fixture(scope="session")
def small_scene():
    from rrpy import synth
    return synth.make_scene(N=2, W=128, H=106, CW=160, CH=135, cv_res=(32, 32, 64))
