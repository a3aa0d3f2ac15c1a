#!/usr/bin/env python
"""Stage the UNMODIFIED reference where the GPU box can see it  --  TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/fetch_ref.py            # /root/reference  ->  oracle/_ref/   (git-ignored, travels with gpurun)

The reference is pure Python, so nothing is compiled: the package (``src/fastforward``), the test files
of the quantization hot path and the tutorial's quantized-Llama helpers are copied
byte for byte.  ``oracle/_ref/`` is listed in .gitignore: reference sources never enter this
repository's history; the staged copy only exists so that (1) the reference's own tests can run on the
B200 with ``fastforward_b200.plugin.install()`` underneath, and (2) bench.py's reference arms time the
real reference instead of a port.  ``oracle/_ref/MANIFEST.json`` records what was staged (sha256 of
every file), so a run can state which reference it used."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference"

# (source relative to the reference root, destination relative to oracle/_ref)
TREES = [
    ("src/fastforward", "src/fastforward"),
    ("tests/quantization", "tests/quantization"),
    ("tests/nn", "tests/nn"),
    ("tests/range_setting", "tests/range_setting"),
    ("docs/examples/doc_helpers", "doc_helpers"),
]
FILES = [
    "tests/__init__.py", "tests/test_dispatcher.py", "tests/test_quantized_tensor.py", "tests/test_range_setting.py",
    "tests/test_overrides.py", "tests/test_flags.py", "tests/test_forward_override.py", "tests/test_gen_fallback.py",
    "LICENSE",
]
KEEP_EXT = (".py", ".pyi", ".json", ".yaml", ".yml", ".txt", ".typed", ".md", "")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def fetch(src_root: str = SRC_DEFAULT, dest: str = DEST) -> dict:
    if not os.path.isdir(os.path.join(src_root, "src", "fastforward")):
        raise FileNotFoundError(f"no reference tree at {src_root}")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    manifest = {}
    for rel_src, rel_dst in TREES:
        top = os.path.join(src_root, rel_src)
        for dirpath, dirnames, filenames in os.walk(top):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", "__snapshots__")]
            for fn in filenames:
                if os.path.splitext(fn)[1] not in KEEP_EXT:
                    continue
                s = os.path.join(dirpath, fn)
                d = os.path.join(dest, rel_dst, os.path.relpath(s, top))
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                manifest[os.path.relpath(d, dest)] = _sha(d)
    for rel in FILES:
        s = os.path.join(src_root, rel)
        if os.path.exists(s):
            d = os.path.join(dest, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            manifest[rel] = _sha(d)
    meta = {"source": src_root, "files": len(manifest), "sha256": manifest}
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(meta, f, indent=0, sort_keys=True)
    return meta


if __name__ == "__main__":
    m = fetch(sys.argv[1] if len(sys.argv) > 1 else SRC_DEFAULT)
    print(f"staged {m['files']} reference files under {DEST}")
