"""Tile <-> row layout helpers (reference: quantization/tiled_tensor.py:19-144).

Pure view/permute bookkeeping used by host code and tests; the CUDA kernels never materialise
the row layout -- they index tiles in place (csrc/ffq_api.cu: make_plan)."""

from __future__ import annotations

import math
from typing import Sequence

import torch


def check_tile_compatibility(input_size: Sequence[int], tile_size: Sequence[int]) -> None:
    if len(input_size) != len(tile_size):
        raise ValueError(
            "Input dimensionality must match tile_size dimensionality got "
            f"{len(input_size)} and {len(tile_size)}"
        )
    bad = [i for i, (d, t) in enumerate(zip(input_size, tile_size)) if t > 0 and d % t != 0]
    if bad:
        raise ValueError(
            "Each dimension of tile_size must divide the corresponding input dimension. Got "
            + ", ".join(f"{input_size[i]} and {tile_size[i]} for dimension {i}" for i in bad)
            + "."
        )


def _resolve(data_shape: Sequence[int], tile_size) -> tuple:
    return tuple(data_shape) if isinstance(tile_size, str) else tuple(int(t) for t in tile_size)


def tiles_to_rows(data: torch.Tensor, tile_size) -> torch.Tensor:
    if data.numel() == 0:
        return data.reshape(1, 0)
    tile = _resolve(data.shape, tile_size)
    check_tile_compatibility(tuple(data.shape), tile)
    n = len(tile)
    split = [v for d, t in zip(data.shape, tile) for v in (d // t, t)]
    order = list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2))
    return data.reshape(split).permute(order).reshape(data.numel() // math.prod(tile), -1)


def rows_to_tiles(tiled_data: torch.Tensor, data_size: Sequence[int], tile_size) -> torch.Tensor:
    data_size = tuple(int(s) for s in data_size)
    if tiled_data.numel() == 0:
        return tiled_data.reshape(data_size)
    tile = _resolve(data_size, tile_size)
    check_tile_compatibility(data_size, tile)
    rows, cols = math.prod(data_size) // math.prod(tile), math.prod(tile)
    if tuple(tiled_data.shape) != (rows, cols):
        raise ValueError(
            f"tiled_data is expected to be of size {torch.Size((rows, cols))} but found {tiled_data.size()}"
        )
    n = len(tile)
    grid = [d // t for d, t in zip(data_size, tile)]
    back = [v for i in range(n) for v in (i, n + i)]
    return tiled_data.reshape(grid + list(tile)).permute(back).reshape(data_size)
