"""fastforward_b200 -- B200-native backend for FastForward's quantization hot path.

The public names mirror ``import fastforward as ff`` for the path this backend covers
(SURVEY.md section 8): ``quantize_model``, ``find_quantizers(...).initialize``, ``estimate_ranges``,
``nn.LinearQuantizer``, ``QuantizedTensor``, the dispatcher and the flags.  All arithmetic is
hand-written CUDA (sm_100a) behind the C ABI in ``include/ffq_b200.h``; importing this package
fails loudly if that library has not been built, and no op has a CPU fallback."""

from . import _cabi as _cabi          # loads lib/libffq_b200.so (ImportError if missing)
from . import dispatcher as dispatcher
from . import exceptions as exceptions
from . import flags as flags
from . import mpath as mpath
from . import nn as nn
from . import ops as ops
from . import quantization as quantization
from . import range_setting as range_setting
from .dispatcher import Predicate as Predicate
from .dispatcher import register as register
from .exceptions import QuantizationError as QuantizationError
from .flags import (  # noqa: F401
    compiled_quant_funcs, export_mode, get_compiled_quant_funcs, get_export_mode, get_strict_quantization,
    set_compiled_quant_funcs, set_export_mode, set_strict_quantization, strict_quantization,
)
from .nn.quantized_module import quantize_model as quantize_model
from .nn.quantized_module import quantized_module_map as quantized_module_map
from .nn.quantized_module import surrogate_quantized_modules as surrogate_quantized_modules
from .overrides import disable_quantization as disable_quantization
from .overrides import enable_quantization as enable_quantization
from .quant_init import QuantizationConfig as QuantizationConfig
from .quant_init import QuantizerCollection as QuantizerCollection
from .quant_init import find_quantizers as find_quantizers
from .quantization.granularity import PerBlock as PerBlock
from .quantization.granularity import PerChannel as PerChannel
from .quantization.granularity import PerTensor as PerTensor
from .quantization.granularity import PerTile as PerTile
from .quantized_tensor import QuantizedTensor as QuantizedTensor
from .range_setting import estimate_ranges as estimate_ranges
from .quantization import fuse as _fuse  # noqa: E402  (after nn: it needs LinearQuantizer)
from .quantization import view_ops as _view_ops  # noqa: E402,F401  (registers the per-tensor view ops with the dispatcher)
from .quantization.fuse import fuse_qdq_weights as fuse_qdq_weights
from .quantization import save_load as _save_load  # noqa: E402  (after nn: it walks the quantizers of a model)

# on-disk formats (reference quantization/__init__.py:16-19, nn/quantized_module.py:357-363): functions of the
# ``quantization`` namespace and, equivalently, methods of every QuantizedModule
for _name in ("save_quantization_state", "load_quantization_state", "save_quantized_model", "load_quantized_model"):
    setattr(quantization, _name, getattr(_save_load, _name))
    setattr(nn.QuantizedModule, _name, getattr(_save_load, _name))
del _name

# the rest of the reference's `fastforward.quantization` namespace (quantization/__init__.py:6-19) and the two module
# aliases of the top level (__init__.py:33-34); bound here because these modules need `nn` to be importable first
from .quantization import freeze as _freeze  # noqa: E402
from .quantization import gptq as _gptq  # noqa: E402,F401  (ff.quantization.gptq is the module AND callable, see its end)
from .quantization.affine import dynamic as _dynamic, static as _static  # noqa: E402
from .quantization.function import create_quantization_function as _create_quantization_function  # noqa: E402

quantization.static, quantization.dynamic = _static, _dynamic
quantization.freeze_parameters = _freeze.freeze_parameters
quantization.create_quantization_function = _create_quantization_function
quantization.QuantizationConfig, quantization.QuantizerCollection = QuantizationConfig, QuantizerCollection
for _name in ("fuse_qdq_weights", "find_weight_quantizers", "stub_weight_quantizers", "ConventionDiscovery",
              "WeightQuantizerDiscovery"):
    setattr(quantization, _name, getattr(_fuse, _name))
del _name
affine, granularity = quantization.affine, quantization.granularity

__version__ = "0.1.0"
