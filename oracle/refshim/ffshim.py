"""Import shim for the UNMODIFIED reference package  --  TEST / BASELINE INFRASTRUCTURE ONLY.

``import ffshim`` BEFORE ``import fastforward``: it pre-seeds the two modules of the reference that
need libcst / mypy (absent from this image) with empty stand-ins, and makes the ``optree`` stand-in
next to this file importable.  Both are off the quantization hot path:
``fastforward.autoquant`` is the source-to-source code generator (src/fastforward/__init__.py:15) and
``fastforward.testing.autoquant`` its test helper (src/fastforward/testing/__init__.py:8)."""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


def _na(*a, **k):
    raise NotImplementedError("autoquant is stubbed in the reference shim (needs libcst/mypy)")


_stub("fastforward.autoquant", autoquantize=_na)
_stub("fastforward.testing.autoquant")
