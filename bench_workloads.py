"""Synthetic workloads for bench.py: Llama-3-shaped decoder stacks (random init, no checkpoint)
driven through the public API exactly as the reference's quick-start does
(docs/examples/quick_start_quantize_llms.nb.py:140-260): quantize_model -> find_quantizers(...)
.initialize(LinearQuantizer, ...) -> estimate_ranges(model, running_minmax) -> forward.

Only the 7 linears per decoder layer are quantized (W: per-channel, A: per-tensor asymmetric on
their inputs); RMSNorm / RoPE / SDPA / SiLU are plain library ops, as they are in the reference
when their quantizers are left as stubs."""

from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch


@dataclasses.dataclass
class LlamaShape:
    name: str
    hidden: int
    ffn: int
    layers: int
    q_heads: int
    kv_heads: int
    head_dim: int = 128
    vocab: int = 128256

    @property
    def linear_params_per_layer(self) -> int:
        kv = self.kv_heads * self.head_dim
        return 2 * self.hidden * self.hidden + 2 * kv * self.hidden + 3 * self.hidden * self.ffn

    @property
    def act_elems_per_token_per_layer(self) -> int:
        # inputs of q,k,v (hidden each), o (hidden), gate, up (hidden each), down (ffn)
        return 6 * self.hidden + self.ffn


LLAMA3_8B = LlamaShape("llama-3-8b-shape", 4096, 14336, 32, 32, 8)
LLAMA3_70B = LlamaShape("llama-3-70b-shape", 8192, 28672, 80, 64, 8)
TINY = LlamaShape("tiny-llama-shape", 256, 512, 2, 4, 2, head_dim=64, vocab=1024)


class RMSNorm(torch.nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5, dtype=None, device=None) -> None:
        super().__init__()
        self.weight = torch.nn.Parameter(torch.ones(dim, dtype=dtype, device=device))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.rms_norm(x, (x.shape[-1],), self.weight, self.eps)   # one library kernel


_ROPE_CACHE: dict = {}


def rope_tables(s: int, d: int, device, dtype, theta: float = 500000.0):
    """(cos, sin_signed), each [S, 1, D] in the half-split layout; sin_signed = (-sin | +sin) so that
    rotate_half(x) * sin == roll(x, D/2) * sin_signed bit for bit (a sign flip commutes with rounding)."""
    key = (s, d, str(device), dtype)
    if key not in _ROPE_CACHE:
        pos = torch.arange(s, device=device, dtype=torch.float32)
        inv = 1.0 / (theta ** (torch.arange(0, d, 2, device=device, dtype=torch.float32) / d))
        ang = torch.cat((pos[:, None] * inv[None, :],) * 2, dim=-1)          # [S, D], half-split layout
        cos, sin = ang.cos().to(dtype), ang.sin().to(dtype)
        sin_signed = torch.cat((-sin[:, : d // 2], sin[:, d // 2:]), dim=-1)
        _ROPE_CACHE[key] = (cos[:, None, :].contiguous(), sin_signed[:, None, :].contiguous())
    return _ROPE_CACHE[key]


def rope(x: torch.Tensor) -> torch.Tensor:
    """x: [B, S, H, D] (as the projection wrote it).  The rotate-half formulation of the reference's model code,
    q*cos + rotate_half(q)*sin, with the rotation done by one roll instead of neg + two slices + cat."""
    cos, sin_signed = rope_tables(x.shape[1], x.shape[-1], x.device, x.dtype)
    return x * cos + torch.roll(x, x.shape[-1] // 2, dims=-1) * sin_signed


class Attention(torch.nn.Module):
    def __init__(self, sh: LlamaShape, dtype, device) -> None:
        super().__init__()
        kv = sh.kv_heads * sh.head_dim
        self.sh = sh
        self.q_proj = torch.nn.Linear(sh.hidden, sh.q_heads * sh.head_dim, bias=False, dtype=dtype, device=device)
        self.k_proj = torch.nn.Linear(sh.hidden, kv, bias=False, dtype=dtype, device=device)
        self.v_proj = torch.nn.Linear(sh.hidden, kv, bias=False, dtype=dtype, device=device)
        self.o_proj = torch.nn.Linear(sh.q_heads * sh.head_dim, sh.hidden, bias=False, dtype=dtype, device=device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, s, _ = x.shape
        sh = self.sh
        q = rope(self.q_proj(x).view(b, s, sh.q_heads, sh.head_dim)).transpose(1, 2)
        k = rope(self.k_proj(x).view(b, s, sh.kv_heads, sh.head_dim)).transpose(1, 2)
        v = self.v_proj(x).view(b, s, sh.kv_heads, sh.head_dim).transpose(1, 2)
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
        return self.o_proj(o.transpose(1, 2).reshape(b, s, sh.q_heads * sh.head_dim))


class MLP(torch.nn.Module):
    def __init__(self, sh: LlamaShape, dtype, device) -> None:
        super().__init__()
        self.gate_proj = torch.nn.Linear(sh.hidden, sh.ffn, bias=False, dtype=dtype, device=device)
        self.up_proj = torch.nn.Linear(sh.hidden, sh.ffn, bias=False, dtype=dtype, device=device)
        self.down_proj = torch.nn.Linear(sh.ffn, sh.hidden, bias=False, dtype=dtype, device=device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.down_proj(torch.nn.functional.silu(self.gate_proj(x)) * self.up_proj(x))


class DecoderLayer(torch.nn.Module):
    def __init__(self, sh: LlamaShape, dtype, device) -> None:
        super().__init__()
        self.input_layernorm = RMSNorm(sh.hidden, dtype=dtype, device=device)
        self.self_attn = Attention(sh, dtype, device)
        self.post_attention_layernorm = RMSNorm(sh.hidden, dtype=dtype, device=device)
        self.mlp = MLP(sh, dtype, device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.self_attn(self.input_layernorm(x))
        return x + self.mlp(self.post_attention_layernorm(x))


class DecoderStack(torch.nn.Module):
    """Embedding + N decoder layers + final norm (no lm_head: calibration of the decoder linears
    does not need logits)."""

    def __init__(self, sh: LlamaShape, layers: Optional[int] = None, dtype=torch.bfloat16, device="cpu") -> None:
        super().__init__()
        n = sh.layers if layers is None else layers
        self.sh = sh
        self.embed_tokens = torch.nn.Embedding(sh.vocab, sh.hidden, dtype=dtype, device=device)
        self.layers = torch.nn.ModuleList(DecoderLayer(sh, dtype, device) for _ in range(n))
        self.norm = RMSNorm(sh.hidden, dtype=dtype, device=device)

    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        x = self.embed_tokens(tokens)
        for layer in self.layers:
            x = layer(x)
        return self.norm(x)


def init_weights_(model: torch.nn.Module, seed: int = 0, std: float = 0.02) -> None:
    """normal(0, 0.02) for every matrix, generated on the parameter's device with a fixed seed."""
    with torch.no_grad():
        for i, (name, p) in enumerate(model.named_parameters()):
            if p.dim() >= 2:
                g = torch.Generator(device=p.device).manual_seed(seed * 100003 + i)
                p.copy_(torch.empty(p.shape, dtype=torch.float32, device=p.device).normal_(0, std, generator=g).to(p.dtype))


def quantize_for_w8a8(ff, model: torch.nn.Module, w_bits: int = 8, a_bits: int = 8, int8_codes: bool = True,
                      w_granularity=None):
    """The quick-start recipe: W per-channel symmetric, A per-tensor asymmetric on the linears' inputs."""
    import bench_workloads as bw  # noqa: F401  (class names resolved by mpath in this module's namespace)

    extra = ff.surrogate_quantized_modules(model)
    ff.quantize_model(model, extra_conversion=extra)
    ff.set_strict_quantization(False)      # as the quick-start does (quick_start_quantize_llms.nb.py:145): stubs stay stubs
    qdt = torch.int8 if int8_codes else None
    ff.find_quantizers(model, "**/layers/**/[quantizer:parameter/weight]").initialize(
        ff.nn.LinearQuantizer, num_bits=w_bits, granularity=w_granularity or ff.PerChannel(0), quantized_dtype=qdt)
    ff.find_quantizers(model, "**/layers/**/[quantizer:activation/input]").initialize(
        ff.nn.LinearQuantizer, num_bits=a_bits, symmetric=False, granularity=ff.PerTensor(), quantized_dtype=qdt)
    return model
