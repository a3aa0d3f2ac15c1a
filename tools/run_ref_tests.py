#!/usr/bin/env python
"""Run the reference's own test files against this backend ("the cuda-parametrised
cases become the parity suite for the new kernels unchanged").

    python tools/run_ref_tests.py [--mode cuda|all|default-cuda|cpu] [--log FILE] [extra pytest args]

  cuda          the tests the reference itself writes for CUDA: every ``cuda``-parametrised case plus the
                tests that hard-code ``device="cuda"`` (tests/test_dispatcher.py:79,101,121,
                tests/test_quantized_tensor.py::test_quantized_tensor_to)
  all           every staged hot-path test file, unchanged (CPU cases run the reference's eager chain, CUDA cases ours)
  default-cuda  the same files with ``torch.set_default_device("cuda")``: the reference's CPU-written tests then
                create their tensors on the GPU and exercise the kernels through the public API
  cpu           ``-k "not cuda"`` (harness check in the build container, which has no GPU)

The reference is the staged copy under oracle/_ref (``python oracle/fetch_ref.py``)."""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
SHIM = os.path.join(ROOT, "oracle", "refshim")

# the hot-path test files
HOT = [
    "tests/quantization/test_tiled_affine.py", "tests/quantization/test_dynamic.py", "tests/quantization/test_tiled_tensor.py",
    "tests/quantization/test_ste.py", "tests/quantization/test_granularity.py", "tests/quantization/affine",
    "tests/quantization/test_fuse.py", "tests/quantization/test_freeze.py", "tests/quantization/test_gptq.py",
    "tests/quantization/test_function.py", "tests/quantization/test_strict_quantization.py",
    "tests/nn/test_linear_quantizer.py", "tests/nn/test_fallback.py", "tests/nn/test_linear_quantized_ops.py",
    "tests/nn/test_quantizer.py", "tests/nn/test_dynamic_linear_quantizer.py",
    "tests/range_setting/test_minmax.py", "tests/range_setting/test_minerror.py",
    "tests/test_range_setting.py", "tests/test_quantized_tensor.py", "tests/test_dispatcher.py", "tests/test_overrides.py",
]
CUDA_K = "cuda or test_dispatch or test_quantized_tensor_to"



def build_command(mode: str, extra=(), no_plugin: bool = False, estimators: bool = False):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REF, "src"), SHIM, ROOT, env.get("PYTHONPATH", "")])
    env["FFQ_REF_PLUGIN"] = "0" if no_plugin else "1"
    env["FFQ_REF_ESTIMATORS"] = "1" if estimators else "0"
    files = [f for f in HOT if os.path.exists(os.path.join(REF, f))]
    cmd = [sys.executable, "-m", "pytest", "-p", "ref_pytest_plugin", "--noconftest", "-c", os.devnull, "--rootdir", REF,
           "-p", "no:cacheprovider", "-q", "-m", "not benchmark"]
    if mode == "cuda":
        cmd += ["-k", CUDA_K]
    elif mode == "cpu":
        cmd += ["-k", "not cuda"]
    elif mode == "default-cuda":
        env["FFQ_REF_DEFAULT_DEVICE"] = "cuda"
    return cmd + files + list(extra), env


def run(mode: str, log=None, extra=(), no_plugin=False, estimators=False):
    if not os.path.isdir(os.path.join(REF, "src", "fastforward")):
        raise FileNotFoundError("oracle/_ref is missing: run `python oracle/fetch_ref.py` in the build container")
    cmd, env = build_command(mode, extra, no_plugin, estimators)
    p = subprocess.run(cmd, cwd=REF, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    out = f"$ (cwd oracle/_ref, mode {mode}) {' '.join(cmd[1:])}\n" + p.stdout
    if log:
        os.makedirs(os.path.dirname(os.path.abspath(log)), exist_ok=True)
        with open(log, "w") as f:
            f.write(out)
    return p.returncode, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="cuda", choices=["cuda", "all", "default-cuda", "cpu"])
    ap.add_argument("--log", default=None)
    ap.add_argument("--no-plugin", action="store_true", help="run the reference alone (its eager path)")
    ap.add_argument("--estimators", action="store_true", help="also patch in the sync-free running_minmax")
    a, rest = ap.parse_known_args()
    rc, out = run(a.mode, a.log, rest, a.no_plugin, a.estimators)
    print("\n".join(out.splitlines()[-60:]))
    sys.exit(rc)


if __name__ == "__main__":
    main()
