"""LPBQ encodings of per-block quantized weights (reference: export/_lpbq.py:15-198).

LPBQ ("low-power blockwise quantization") stores the fp32 per-block scales of a symmetric ``PerBlock`` weight quantizer
as one float scale per channel and a ``compressed_bw``-bit integer per block.  The arithmetic -- a per-channel max, two
divisions, a round and a clamp over every scale of the model -- is the CUDA kernel ``ffq_lpbq_encode`` (one pass);
this module is the part around it: which quantizers qualify, the 2-D view of the flat scale vector and the encoding
dictionary with the reference's keys.  There is no CPU path: scales that live on the host are rejected."""

from __future__ import annotations

from typing import Any, Dict, Tuple

import torch

from .. import ops
from . import granularity as G


class LPBQProcessor:
    """``LPBQProcessor(compressed_bw=4, decompressed_bw=8)``: PerBlock quantization parameters -> LPBQ encoding."""

    def __init__(self, compressed_bw: int = 4, decompressed_bw: int = 8) -> None:
        if compressed_bw <= 0 or decompressed_bw <= 0:
            raise ValueError(f"Bitwidths cannot be 0 or negative (compressed_bitwidth={compressed_bw}, "
                             f"decompressed_bitwdith={decompressed_bw})")
        if compressed_bw >= decompressed_bw:
            raise ValueError("Compressed bitwidth cannot be larger than decompressed bitwidth "
                             f"(compressed_bitwidth={compressed_bw}, decompressed_bitwdith={decompressed_bw})")
        if compressed_bw > 8:
            raise ValueError(f"Compressed bitwidth can be max 8, got compressed_bitwidth={compressed_bw}")
        if decompressed_bw > 32:
            raise ValueError(f"Decompressed bitwidth can be max 32, got decompressed_bitwidth={decompressed_bw}")
        self.compressed_bw = compressed_bw
        self.decompressed_bw = decompressed_bw

    # ---- which parameters qualify, and how the flat scale vector folds into [channels x blocks] ------------------
    def _layout(self, tensor_name: str, data_shape, tile_size, bitwidth: int, is_symmetric: bool) -> Tuple[int, Tuple[int, int], int]:
        """(block size, 2-D shape of the scales, channel axis) or ValueError when LPBQ does not apply: exactly one
        block dimension and one per-channel dimension, symmetric, and the quantizer's bit width == compressed_bw."""
        gran = G.granularity_from_sizes(torch.Size(data_shape), torch.Size(tile_size))
        if not (isinstance(gran, G.PerBlock) and len(gran.block_dims) == 1 and len(gran.per_channel_dims) == 1
                and is_symmetric is True and bitwidth == self.compressed_bw):
            raise ValueError(f"Parameters for {tensor_name} not suitable for LPBQ")
        block = gran.block_sizes[0]
        if gran.block_dims == (1,) and gran.per_channel_dims == (0,):
            return block, (data_shape[0], data_shape[1] // block), 0        # [out_channels, blocks]: a channel is a row
        if gran.block_dims == (0,) and gran.per_channel_dims == (1,):
            return block, (data_shape[0] // block, data_shape[1]), 1        # [blocks, in_channels]: a channel is a column
        raise ValueError(f"Parameters for {tensor_name} not suitable for LPBQ")

    def grouped_dynamic_quantize(self, scale_2d: torch.Tensor, block_grouping, bitwidth: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """``block_grouping`` as in the reference: ``[1, -1]`` groups along dimension 1 (a channel is a row),
        ``[-1, 1]`` along dimension 0.  Returns (per-block integers shaped like ``scale_2d``, per-channel float scale
        with the grouped dimension kept as size 1)."""
        grouping = list(block_grouping)
        if grouping == [1, -1]:
            axis = 0
        elif grouping == [-1, 1]:
            axis = 1
        else:
            raise NotImplementedError(f"block_grouping {grouping}: only [1, -1] and [-1, 1] occur for LPBQ")
        iq, fs = ops.lpbq_encode(scale_2d, axis, bitwidth)
        return iq, (fs.reshape(-1, 1) if axis == 0 else fs.reshape(1, -1))

    def generate_lpbq_encoding(self, tensor_name: str, scale: torch.Tensor, data_shape, tile_size, bitwidth: int,
                               is_symmetric: bool = True) -> Dict[str, Any]:
        """The encoding dictionary of one tensor (export/_lpbq.py:76-129): ``scale`` is the quantizer's flat per-tile
        scale vector."""
        block, shape2d, axis = self._layout(tensor_name, tuple(data_shape), tuple(tile_size), bitwidth, is_symmetric)
        iq, fs = ops.lpbq_encode(scale.detach().reshape(shape2d), axis, self.compressed_bw)
        floats = fs.reshape(-1).tolist()
        return {
            "name": tensor_name, "dtype": "INT", "enc_type": "LPBQ", "is_sym": True,
            "compressed_bw": self.compressed_bw, "bw": self.decompressed_bw, "block_size": block,
            "per_block_int_scale": iq.reshape(-1).tolist(), "scale": floats,
            "offset": [float(-(2 ** (self.decompressed_bw - 1)))] * len(floats),
        }

    def encode_quantizer(self, tensor_name: str, quantizer, data_shape) -> Dict[str, Any]:
        """Convenience over ``generate_lpbq_encoding`` for a calibrated ``LinearQuantizer`` and the weight's shape."""
        tile = quantizer.granularity.tile_size(torch.Size(data_shape))
        if isinstance(tile, str):
            tile = tuple(data_shape)
        return self.generate_lpbq_encoding(tensor_name, quantizer.scale, data_shape, tuple(tile), int(quantizer.num_bits),
                                           bool(quantizer.symmetric))
