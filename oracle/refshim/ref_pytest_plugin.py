"""pytest plugin that runs the reference's OWN test files with the B200 kernels underneath
--  TEST INFRASTRUCTURE ONLY.

Loaded with ``-p ref_pytest_plugin --noconftest`` by ``tools/run_ref_tests.py``.  It replaces the
reference's ``tests/conftest.py`` (which imports the libcst-based code generator,
``tests/conftest.py:120``) with the two fixtures the selected tests use (``tests/conftest.py:80-90``)
and the ``slow`` / ``benchmark`` markers, imports the unmodified reference through ``ffshim`` and --
unless ``FFQ_REF_PLUGIN=0`` -- calls ``fastforward_b200.plugin.install()`` so every CUDA tensor that
reaches ``torch.ops.fastforward.*`` or the dispatcher's ``linear`` runs this repository's kernels.

Environment:
  FFQ_REF_PLUGIN=0|1          install the B200 backend (default 1)
  FFQ_REF_ESTIMATORS=0|1      also swap in the sync-free running_minmax estimator (default 0)
  FFQ_REF_DEFAULT_DEVICE=cuda make ``cuda`` the default device of every factory call in the tests, so the
                              reference's CPU-written tests drive the CUDA kernels unchanged
"""
import os
import random

import pytest
import torch

import ffshim  # noqa: F401  (must precede the first import of fastforward)
import fastforward  # noqa: E402
from fastforward.testing import seed_prngs  # noqa: E402

_PLUGIN = os.environ.get("FFQ_REF_PLUGIN", "1") != "0"
_ESTIMATORS = os.environ.get("FFQ_REF_ESTIMATORS", "0") == "1"
_DEFAULT_DEVICE = os.environ.get("FFQ_REF_DEFAULT_DEVICE", "")

if _PLUGIN:
    from fastforward_b200 import plugin as _plugin

    _plugin.install(fastforward, patch_estimators=_ESTIMATORS)


def pytest_configure(config):
    config.addinivalue_line("markers", "slow: reference marker (deselected by default upstream)")
    config.addinivalue_line("markers", "benchmark: reference marker (deselected by default upstream)")
    if _DEFAULT_DEVICE:
        torch.set_default_device(_DEFAULT_DEVICE)


def pytest_report_header(config):
    lines = [f"reference: {os.path.dirname(fastforward.__file__)}", f"default device: {_DEFAULT_DEVICE or 'cpu'}"]
    if _PLUGIN:
        from fastforward_b200 import _cabi

        lines.append(f"B200 backend installed (estimators patched: {_ESTIMATORS}); library: {_cabi.LIB_PATH}")
    else:
        lines.append("B200 backend NOT installed: the reference's own eager path")
    return lines


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    if _PLUGIN:
        from fastforward_b200 import _cabi

        terminalreporter.write_line(f"ffq kernel launches during this session: {_cabi.launch_count()}")


@pytest.fixture(scope="session", name="random_seed")
def random_seed_fixture() -> int:
    return random.randint(0, 2**64 - 1)


@pytest.fixture(name="_seed_prngs")
def seed_prngs_fixture(random_seed: int) -> int:
    seed_prngs(random_seed)
    return random_seed
