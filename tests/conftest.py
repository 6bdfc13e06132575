import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files():
    """Fixtures of the main render path (volume_render / backward / depth / query)."""
    return sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith(("x_", "y_", "ref_")))


def golden_variant_files():
    """Fixtures of the march variants (opacity_render, motion_render)."""
    return sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("x_ball"))


def golden_fmt_file():
    """Fixture of the view-dependent formats + motion-feature render (tests/golden/make_golden_fmt.py)."""
    return os.path.join(GOLDEN_DIR, "y_fmt_ball_L4.npz")


def fmt_case_names(z):
    return sorted(k[:-5] for k in z.files if k.endswith("_meta"))


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")
