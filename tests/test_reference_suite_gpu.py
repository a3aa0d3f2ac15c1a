"""The reference's OWN test files, unchanged, on the B200 with ``fastforward_b200.plugin.install()`` underneath
("the cuda-parametrised cases become the parity suite for the new kernels unchanged").

The files are the staged copy under oracle/_ref/tests (``python oracle/fetch_ref.py``); they run in a subprocess with
the reference on PYTHONPATH (tools/run_ref_tests.py).  Logs go to gpurun_out/ (copied to profiles/ when committed)."""
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import ref_loader  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not staged (python oracle/fetch_ref.py)")]


def _counts(out: str) -> dict:
    tail = out.strip().splitlines()[-1]
    return {k: int(v) for v, k in re.findall(r"(\d+) (passed|failed|error|errors|skipped|deselected|xfailed)", tail)}


def _launches(out: str) -> int:
    m = re.search(r"ffq kernel launches during this session: (\d+)", out)
    return int(m.group(1)) if m else -1


def test_reference_cuda_tests_pass_on_the_b200_backend():
    import run_ref_tests

    log = os.path.join(ROOT, "gpurun_out", "ref_tests_cuda.log")
    rc, out = run_ref_tests.run("cuda", log=log)
    c = _counts(out)
    assert rc == 0 and c.get("failed", 0) == 0 and c.get("error", 0) + c.get("errors", 0) == 0, out[-4000:]
    assert c.get("passed", 0) >= 500, out[-2000:]          # 544 cuda-parametrised / cuda-only cases in the staged files
    assert _launches(out) > 1000, "the reference's CUDA tests did not reach the ffq kernels"


def test_reference_cpu_written_tests_drive_the_kernels_with_cuda_as_default_device():
    """Every staged hot-path test file with torch.set_default_device('cuda'): tensors the tests create land on the GPU,
    so LinearQuantizer / QuantizedTensor / estimate_ranges / fuse / freeze / gptq run on our kernels through the
    reference's public API.  Some of these CPU-written tests cannot pass with a CUDA default device whatever the backend
    (quantizers are constructed with device="cpu", results are compared with CPU literals): the same files are first run
    with the reference ALONE, and the backend must not fail a single test the reference alone passes."""
    import run_ref_tests

    out_dir = os.path.join(ROOT, "gpurun_out")
    _, base = run_ref_tests.run("default-cuda", log=os.path.join(out_dir, "ref_tests_default_cuda_noplugin.log"),
                                extra=["-rf"], no_plugin=True)
    _, out = run_ref_tests.run("default-cuda", log=os.path.join(out_dir, "ref_tests_default_cuda.log"), extra=["-rf"])
    failed_alone = set(re.findall(r"^FAILED (\S+)", base, flags=re.M))
    failed = set(re.findall(r"^FAILED (\S+)", out, flags=re.M))
    regressions = sorted(failed - failed_alone)
    assert not regressions, "\n".join(regressions) + "\n" + out[-6000:]
    assert _counts(out).get("passed", 0) >= max(700, _counts(base).get("passed", 0)), out[-2000:]
    assert _launches(out) > 2000
