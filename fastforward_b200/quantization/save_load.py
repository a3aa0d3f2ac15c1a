"""On-disk formats of a quantized model (reference: quantization/save_load.py:59-746).

Two layouts, both byte-compatible with the reference's so that files move freely between this package and an installed
``fastforward`` (pinned by tests/test_save_load.py against files the unmodified reference wrote and reads):

* **quantization state**  ``<cache>/quantization-state/<name>/<tag>/{config.yaml, model.safetensors}``
  -- the quantizers only: their configuration as ``!ff.obj`` YAML (``serialization.py``) and their parameters in one
  safetensors file whose string metadata maps ``quantizer name -> "param=tensor_key[::lazy],..."``;
* **quantized-model artifact**  ``<dir>/{config.yaml, quantizer_state.safetensors, weights.safetensors, manifest.json}``
  -- the above plus the model's own weights (tied tensors stored once, aliases in the manifest).

Host-only code: nothing here touches the hot path; tensors are written from wherever they live (safetensors copies
device tensors to the host itself)."""

from __future__ import annotations

import json
import logging
import os
from operator import attrgetter
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple, Union

import torch

from .. import serialization
from ..exceptions import QuantizationError
from ..nn.quantized_module import named_quantizers
from ..nn.quantizer import Quantizer, QuantizerStub

logger = logging.getLogger(__name__)

FORMAT_VERSION = "1.0"
EMPTY_MARKER = "__ff_empty_quantizer_state__"    # safetensors cannot re-read an empty dict with empty metadata
_LAZY = "::lazy"
PathLike = Union[str, Path]


def _ff_version() -> str:
    from .. import __version__
    return str(__version__)


# ---------------------------------------------------------------------------------------------------------------
# where things live
# ---------------------------------------------------------------------------------------------------------------
def get_assets_path(namespace: str, tag: str, *, cache_dir: Optional[PathLike] = None) -> Path:
    """``<cache>/<namespace>/<tag>`` with the reference's cache-root rules (cache.py:9-49): explicit ``cache_dir``,
    else ``$FF_CACHE``, else ``$XDG_CACHE_HOME/fastforward``, else ``~/.cache/fastforward``; characters that cannot
    be part of one path component are replaced by ``--``."""
    root = cache_dir or os.getenv("FF_CACHE")
    if root is None:
        xdg = os.getenv("XDG_CACHE_HOME")
        root = Path(xdg) / "fastforward" if xdg else Path.home() / ".cache" / "fastforward"
    root = Path(root).expanduser().resolve()
    for bad in " /\\:<>|":
        namespace, tag = namespace.replace(bad, "--"), tag.replace(bad, "--")
    path = root / namespace / tag
    if path.is_file():
        raise ValueError(f"The asset path '{path}' points to an existing file (not a directory)")
    return path


def _model_identifier(model: torch.nn.Module, given: Optional[PathLike]) -> PathLike:
    name = given if given is not None else getattr(getattr(model, "config", None), "name_or_path", None)
    if name is None:
        raise RuntimeError("Unable to detect the model identifier. Please provide it manually if there is no "
                           "`config.name_or_path` property in the model")
    return name


def _make_dir(path: Path) -> Path:
    try:
        path.mkdir(exist_ok=True, parents=True, mode=0o775)
    except (FileExistsError, NotADirectoryError) as e:
        raise ValueError(f"Cannot create directory {path} because of an existing file.") from e
    return path


# ---------------------------------------------------------------------------------------------------------------
# quantizer state <-> (tensors, metadata, config)
# ---------------------------------------------------------------------------------------------------------------
def _gather(model: torch.nn.Module, allow_lazy: bool):
    """(tensors, metadata, quantizers by name).  A quantizer instance attached in several places is stored once, under
    the lexicographically first of its names; every name gets a metadata entry pointing at those tensors."""
    names_of: Dict[Quantizer, List[str]] = {}
    for name, quantizer in named_quantizers(model, remove_duplicate=False):
        names_of.setdefault(quantizer, []).append(name)
    tensors: Dict[str, torch.Tensor] = {}
    metadata: Dict[str, str] = {}
    by_name: Dict[str, Quantizer] = {}
    for quantizer, names in names_of.items():
        owner = min(names)
        entries = []
        lazy = []
        for key, value in quantizer.state_dict(keep_vars=True).items():
            if torch.nn.parameter.is_lazy(value):
                lazy.append(key)
                entries.append(f"{key}={owner}.{key}{_LAZY}")
            else:
                tensors[f"{owner}.{key}"] = value.detach()
                entries.append(f"{key}={owner}.{key}")
        if lazy:
            msg = (f"A quantizer having lazy parameters (UninitializedParameter or UninitializedBuffer) was found. "
                   f"Parameters: {set(lazy)}.\nTip: quantizers normally materialize the uninitialized parameters during "
                   "range estimation.")
            if not allow_lazy:
                logger.error(msg)
                raise ValueError(msg)
            logger.warning(msg)
        for name in names:
            metadata[name] = ",".join(entries)
            by_name[name] = quantizer
    return tensors, metadata, by_name


def _write_state(model: torch.nn.Module, directory: Path, identifier: PathLike, tensor_file: str, allow_lazy: bool) -> Path:
    from safetensors.torch import save_file

    tensors, metadata, by_name = _gather(model, allow_lazy)
    if not tensors and not metadata:
        metadata = {EMPTY_MARKER: "true"}
    save_file({k: v.contiguous() for k, v in tensors.items()}, str(directory / tensor_file), metadata=metadata)
    config = {
        "version": FORMAT_VERSION,
        "name_or_path": str(identifier),
        "transformers_version": str(getattr(getattr(model, "config", None), "transformers_version", None)),
        "fastforward_version": _ff_version(),
        "quantizers": by_name,
    }
    config_path = directory / "config.yaml"
    with open(config_path, "w") as f:
        serialization.dump(config, f)
    return config_path


def _parse_entries(spec: str) -> Tuple[Dict[str, str], List[str]]:
    present: Dict[str, str] = {}
    lazy: List[str] = []
    for item in spec.split(","):
        if not item:
            continue
        key, _, target = item.partition("=")
        target, _, decoration = target.partition("::")
        if "lazy" in decoration:
            lazy.append(key)
        else:
            present[key] = target
    return present, lazy


def _attach(model: torch.nn.Module, name: str, quantizer: Quantizer, policy: str) -> None:
    parent_path, _, attribute = name.rpartition(".")
    parent = attrgetter(parent_path)(model) if parent_path else model
    current = getattr(parent, attribute, None)
    if not isinstance(current, Quantizer):
        raise ValueError(f"'{name}' is not a quantizer or was overwritten by a non-quantizer object")
    if not isinstance(current, QuantizerStub):
        if policy == "skip":
            return
        if policy == "error":
            raise QuantizationError(
                f"'{name}' is a quantizer, but is already initialized. If you want to overwrite the existing quantizer, "
                'use overwrite_policy="overwrite" or if you want to skip loading existing quantizers use '
                'overwrite_policy="skip"')
        if policy != "overwrite":
            raise QuantizationError(
                f"Encountered a quantizer that was already initialized. Since overwrite_policy={policy} is illegal "
                "cannot resolve conflict.please use 'error', 'skip', or 'overwrite")
    if quantizer.quant_metadata is None:
        quantizer.quant_metadata = current.quant_metadata        # metadata belongs to the slot, not to the file
    setattr(parent, attribute, quantizer)


def _read_state(model: torch.nn.Module, config_path: Path, tensor_path: Path, expected_name: Optional[str],
                policy: str, allow_lazy: bool) -> None:
    from safetensors import safe_open

    with open(config_path) as f:
        config = serialization.load(f)
    if config.get("version") != FORMAT_VERSION:
        raise ValueError(f"Unsupported quantization state version: {config.get('version')}")
    if expected_name is not None and str(config.get("name_or_path")) != str(expected_name):
        msg = f"Model identifier mismatch: expected '{expected_name}', found '{config.get('name_or_path')}' in saved state"
        logger.error(msg)
        raise RuntimeError(msg)
    quantizers: Dict[str, Quantizer] = config.get("quantizers") or {}
    if quantizers:
        restored = set()
        with safe_open(str(tensor_path), framework="pt") as f:
            metadata = f.metadata() or {}
            for name, quantizer in quantizers.items():
                present, lazy = _parse_entries(metadata[name])
                if id(quantizer) not in restored:         # a shared instance appears under each of its names
                    restored.add(id(quantizer))
                    missing, unexpected = quantizer.load_state_dict(
                        {key: f.get_tensor(target) for key, target in present.items()}, strict=False)
                else:
                    missing, unexpected = [], []
                if lazy:
                    msg = f"Lazy parameters were found in quantization state and cannot be loaded. Parameters: {lazy}."
                    if not allow_lazy:
                        logger.error(msg)
                        raise ValueError(msg)
                    logger.warning(msg)
                missing = sorted(set(missing) - set(lazy))
                if missing or unexpected:
                    msg = (f"There are some missing ({missing}) or unexpected ({list(unexpected)}) keys during loading "
                           "state_dict")
                    logger.error(msg)
                    raise RuntimeError(msg)
    for name, quantizer in quantizers.items():
        _attach(model, name, quantizer, policy)


# ---------------------------------------------------------------------------------------------------------------
# public: quantization state
# ---------------------------------------------------------------------------------------------------------------
def save_quantization_state(model: torch.nn.Module, *, tag: str = "main", name_or_path: Optional[PathLike] = None,
                            cache_dir: Optional[PathLike] = None, allow_lazy_params: bool = False) -> Path:
    """Write the configuration and parameters of every (non-stub) quantizer of ``model`` below the asset cache;
    returns the path of ``config.yaml`` (save_load.py:277-330)."""
    identifier = _model_identifier(model, name_or_path)
    directory = _make_dir(get_assets_path(f"quantization-state/{identifier}", tag, cache_dir=cache_dir))
    return _write_state(model, directory, identifier, "model.safetensors", allow_lazy_params)


def load_quantization_state(model: torch.nn.Module, *, tag: str = "main", name_or_path: Optional[PathLike] = None,
                            cache_dir: Optional[PathLike] = None, overwrite_policy: str = "error",
                            allow_lazy_params: bool = False) -> None:
    """Rebuild the saved quantizers and put them where the stubs of ``model`` are (save_load.py:333-397).
    ``name_or_path`` may be the path of a ``config.yaml`` directly; then the identifier check is the model's own."""
    name = getattr(getattr(model, "config", None), "name_or_path", None)
    as_path = Path(name_or_path) if name_or_path is not None else None
    if as_path is not None and not as_path.exists():
        name = str(name_or_path)
    if name is None:
        raise RuntimeError("Unable to detect the model identifier. Please provide it manually if there is no "
                           "`config.name_or_path` property in the model")
    if as_path is not None and as_path.exists():
        config_path = as_path
    else:
        config_path = get_assets_path(f"quantization-state/{name}", tag, cache_dir=cache_dir) / "config.yaml"
    tensor_path = config_path.parent / "model.safetensors"
    if not config_path.exists():
        raise FileNotFoundError(f"Quantization state config not found at {config_path}")
    if not tensor_path.exists():
        raise FileNotFoundError(f"Quantization state model not found at {tensor_path}")
    _read_state(model, config_path, tensor_path, name, overwrite_policy, allow_lazy_params)


# ---------------------------------------------------------------------------------------------------------------
# public: self-contained artifact
# ---------------------------------------------------------------------------------------------------------------
def _split_tied(tensors: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, str]]:
    """safetensors refuses tensors that share memory: keep one name (the lexicographically first) per storage view and
    remember the others as aliases of it."""
    groups: Dict[tuple, List[str]] = {}
    for name, t in tensors.items():
        groups.setdefault((t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.device), []).append(name)
    kept: Dict[str, torch.Tensor] = {}
    aliases: Dict[str, str] = {}
    for names in groups.values():
        first = min(names)
        kept[first] = tensors[first]
        aliases.update({n: first for n in names if n != first})
    return kept, aliases


def _quantizer_prefixes(model: torch.nn.Module) -> tuple:
    return tuple(f"{name}." for name, _ in named_quantizers(model, remove_duplicate=False, skip_stubs=False))


def _own_weights(model: torch.nn.Module) -> Iterable[Tuple[str, torch.Tensor]]:
    prefixes = _quantizer_prefixes(model)
    for name, value in model.state_dict().items():
        if not (prefixes and name.startswith(prefixes)):
            yield name, value


def save_quantized_model(model: torch.nn.Module, path: PathLike, *, name_or_path: Optional[PathLike] = None,
                         allow_lazy_params: bool = False) -> Path:
    """Snapshot of the model exactly as it is -- quantizers and weights -- in one directory (save_load.py:540-640).
    Grid-snapping or stubbing the weights first is the caller's choice (``fuse_qdq_weights``)."""
    from safetensors.torch import save_file

    identifier = _model_identifier(model, name_or_path)
    directory = _make_dir(Path(path))
    _write_state(model, directory, identifier, "quantizer_state.safetensors", allow_lazy_params)
    weights, tied = _split_tied({name: value.detach() for name, value in _own_weights(model)})
    save_file({k: v.contiguous() for k, v in weights.items()}, str(directory / "weights.safetensors"))
    with open(directory / "manifest.json", "w") as f:
        json.dump({"version": FORMAT_VERSION, "name_or_path": str(identifier), "fastforward_version": _ff_version(),
                   "tied_weights": tied}, f, indent=2)
    return directory


def load_quantized_model(model: torch.nn.Module, path: PathLike, *, overwrite_policy: str = "error",
                         allow_lazy_params: bool = False, expected_name: Optional[str] = None) -> None:
    """Restore what ``save_quantized_model`` wrote: the quantizers replace the model's stubs, then the weights are
    loaded; weight keys must match the model's exactly (save_load.py:643-746).  ``expected_name=""`` skips the
    identifier check."""
    from safetensors import safe_open

    directory = Path(path)
    files = {n: directory / n for n in ("config.yaml", "quantizer_state.safetensors", "weights.safetensors", "manifest.json")}
    for p in files.values():
        if not p.exists():
            raise FileNotFoundError(f"Quantized model artifact file not found: {p}")
    with open(files["manifest.json"]) as f:
        manifest = json.load(f)
    if manifest.get("version") != FORMAT_VERSION:
        raise ValueError(f"Unsupported quantized model artifact version: {manifest.get('version')}")
    if expected_name is None:
        expected_name = getattr(getattr(model, "config", None), "name_or_path", None)
    _read_state(model, files["config.yaml"], files["quantizer_state.safetensors"], expected_name or None,
                overwrite_policy, allow_lazy_params)
    with safe_open(str(files["weights.safetensors"]), framework="pt") as f:
        weights = {key: f.get_tensor(key) for key in f.keys()}
    for alias, kept in (manifest.get("tied_weights") or {}).items():
        if kept not in weights:
            raise RuntimeError(f"Tied weight '{alias}' references '{kept}', which is missing from the saved weights.")
        weights[alias] = weights[kept]
    expected = {name for name, _ in _own_weights(model)}
    missing, unexpected = sorted(expected - weights.keys()), sorted(weights.keys() - expected)
    if missing or unexpected:
        msg = f"Saved weights do not match this model: missing {missing}, unexpected {unexpected}."
        logger.error(msg)
        raise RuntimeError(msg)
    model.load_state_dict(weights, strict=False)
