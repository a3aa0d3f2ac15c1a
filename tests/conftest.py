"""pytest configuration: markers, repo root on sys.path, golden-fixture loader."""
import gzip
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_CACHE = {}


def load_golden(name):
    """Cases recorded from the unmodified reference by oracle/make_golden.py."""
    if name not in _CACHE:
        with gzip.open(os.path.join(GOLDEN_DIR, f"{name}.pt.gz"), "rb") as fh:
            _CACHE[name] = torch.load(fh, weights_only=False)["cases"]
    return _CACHE[name]


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Bit-for-bit equality: distinguishes -0.0 from 0.0; NaNs must sit at the same positions
    (their payload bits are not compared -- GPUs return the canonical NaN)."""
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    a, b = a.detach().contiguous().cpu(), b.detach().contiguous().cpu()
    if a.numel() == 0:
        return True
    if a.dtype.is_floating_point:
        na, nb = torch.isnan(a), torch.isnan(b)
        if not torch.equal(na, nb):
            return False
        if na.any():
            a, b = a.masked_fill(na, 0), b.masked_fill(nb, 0)
    return torch.equal(a.view(torch.uint8), b.view(torch.uint8))


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(autouse=True)
def _restore_process_wide_flags():
    """strict_quantization / export_mode are process-wide switches (as in the reference, flags.py:49-98); a test that
    flips one (the quick-start recipe does, bench_workloads.quantize_for_w8a8) must not change what later tests see."""
    try:
        import fastforward_b200 as ff
    except Exception:                       # library not built: the tests that need it fail on their own import
        yield
        return
    strict, export = ff.get_strict_quantization(), ff.get_export_mode()
    yield
    ff.set_strict_quantization(strict)
    ff.set_export_mode(export)
