"""Import the staged, UNMODIFIED reference package  --  TEST / BASELINE INFRASTRUCTURE ONLY.

``load_reference()`` returns the real ``fastforward`` module from ``oracle/_ref/src`` (staged by
``oracle/fetch_ref.py``; git-ignored, shipped to the GPU box by gpurun) after installing the import
shim for its two absent, off-path dependencies (``oracle/refshim``).  Nothing under
``fastforward_b200/`` imports this; users are tests/, bench.py's reference arms and tools/."""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
REF_SRC = os.path.join(REF_ROOT, "src")
SHIM = os.path.join(HERE, "refshim")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "fastforward", "__init__.py"))


def manifest() -> dict:
    try:
        with open(os.path.join(REF_ROOT, "MANIFEST.json")) as f:
            m = json.load(f)
        return {"files": m.get("files"), "source": m.get("source")}
    except OSError:
        return {}


def load_reference():
    """The reference's top-level module (``import fastforward as ff`` of the staged copy)."""
    if not available():
        raise ImportError("the staged reference is missing: run `python oracle/fetch_ref.py` where /root/reference exists")
    for p in (SHIM, REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ffshim  # noqa: F401  (stubs fastforward.autoquant before the package imports it)
    import fastforward

    if not os.path.abspath(fastforward.__file__).startswith(os.path.abspath(REF_SRC)):
        raise ImportError(f"`fastforward` resolved to {fastforward.__file__}, not the staged reference")
    return fastforward
