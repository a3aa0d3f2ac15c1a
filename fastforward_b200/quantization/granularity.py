"""Granularities: how quantization parameters are shared over a tensor
(reference: quantization/granularity.py:20-332).  A granularity only answers
``tile_size(data_shape)``; everything downstream works on tiles."""

from __future__ import annotations

import abc
import logging
from typing import Any, Sequence

import torch

from ..serialization import remember_init_args
from .tiled_tensor import check_tile_compatibility

logger = logging.getLogger(__name__)


def _tuple(v) -> tuple:
    return (v,) if isinstance(v, int) else tuple(v)


@remember_init_args
class Granularity(abc.ABC):
    _fields: tuple = ()

    @abc.abstractmethod
    def tile_size(self, data_shape: torch.Size):
        """``torch.Size`` of one tile, or the literal ``"data_shape"`` for per-tensor."""

    def parameter_dimensionality(self, data_shape: torch.Size) -> int:
        tile = self.tile_size(data_shape)
        if isinstance(tile, str):
            return 1
        return torch.Size(data_shape).numel() // torch.Size(tile).numel()

    def repr_args(self) -> dict:
        return {}

    def __repr__(self) -> str:
        return f"{type(self).__name__}({', '.join(f'{k}={v}' for k, v in self.repr_args().items())})"

    def __eq__(self, other: object) -> bool:
        return type(self) is type(other) and all(getattr(self, f) == getattr(other, f) for f in self._fields)

    def __hash__(self) -> int:
        return hash((type(self).__name__,) + tuple(getattr(self, f) for f in self._fields))


class PerTensor(Granularity):
    def tile_size(self, data_shape: torch.Size):
        return "data_shape"


class PerChannel(Granularity):
    _fields = ("channel_dims",)
    __match_args__ = ("channel_dims",)

    def __init__(self, channel_dim: int | Sequence[int] = 0) -> None:
        self.channel_dims = _tuple(channel_dim)

    def tile_size(self, data_shape: torch.Size) -> torch.Size:
        tile = list(data_shape)
        for d in self.channel_dims:
            tile[d] = 1
        return torch.Size(tile)

    def repr_args(self) -> dict:
        return {"channel": self.channel_dims[0] if len(self.channel_dims) == 1 else self.channel_dims}


class PerBlock(Granularity):
    _fields = ("block_dims", "block_sizes", "per_channel_dims", "strict_blocks")

    def __init__(self, block_dims, block_sizes, per_channel_dims=(), strict_blocks: bool = True) -> None:
        self.block_dims, self.block_sizes = _tuple(block_dims), _tuple(block_sizes)
        self.per_channel_dims, self.strict_blocks = _tuple(per_channel_dims), strict_blocks
        if len(self.block_dims) != len(self.block_sizes):
            raise ValueError("block_sizes and block_dims must be of equal length")
        both = [str(d) for d in self.per_channel_dims if d in self.block_dims]
        if both:
            logger.warning(
                f"Dimensions {', '.join(both)} are in both 'block_dims' and 'per_channel_dims'. "
                "They will be quantized as per-block following 'block_sizes'"
            )

    def tile_size(self, data_shape: torch.Size) -> torch.Size:
        tile = list(data_shape)
        for d in self.per_channel_dims:
            tile[d] = 1
        for d, size in zip(self.block_dims, self.block_sizes):
            if size > data_shape[d]:
                raise ValueError(
                    f"Can't apply per block quantization using block-size={size} over dimension {d} "
                    f"for a tensor with shape {data_shape}. "
                )
            if self.strict_blocks and data_shape[d] % size != 0:
                raise ValueError(
                    f"Block dim {d} of size {size} does not divide the data dim {data_shape[d]} exactly. "
                    "This is required because strict_blocks=True"
                )
            tile[d] = size
        return torch.Size(tile)

    def repr_args(self) -> dict:
        return {f: getattr(self, f) for f in self._fields}


class PerTile(Granularity):
    _fields = ("tile_shape",)
    __match_args__ = ("tile_shape",)

    def __init__(self, tile_shape: Sequence[int]) -> None:
        self.tile_shape = torch.Size(tile_shape)

    def tile_size(self, data_shape: torch.Size) -> torch.Size:
        check_tile_compatibility(data_shape, self.tile_shape)
        return self.tile_shape

    def repr_args(self) -> dict:
        return {"tile_shape": self.tile_shape}


def is_per_tensor(g: Granularity) -> bool:
    return isinstance(g, PerTensor)


def is_per_channel(g: Granularity) -> bool:
    return isinstance(g, PerChannel)


def is_per_block(g: Granularity) -> bool:
    return isinstance(g, PerBlock)


def granularity_from_sizes(data_size: torch.Size, tile_size: torch.Size) -> Granularity:
    """Simplest granularity with ``g.tile_size(data_size) == tile_size``."""
    data_size, tile_size = torch.Size(data_size), torch.Size(tile_size)
    if data_size == tile_size:
        return PerTensor()
    dims = range(len(data_size))
    if all(t == d or t == 1 for d, t in zip(data_size, tile_size)):
        return PerChannel(tuple(i for i in dims if tile_size[i] == 1 and data_size[i] > 1))
    block_dims = tuple(i for i in dims if tile_size[i] not in (1, data_size[i]))
    return PerBlock(
        block_dims,
        tuple(tile_size[i] for i in block_dims),
        tuple(i for i in dims if tile_size[i] == 1 and data_size[i] > 1),
        strict_blocks=all(d % t == 0 for d, t in zip(data_size, tile_size)),
    )
