"""oracle.llama — one Llama decode step on CPU, op by op.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates the operator sequence of
src/models/meta/llama.h:61-64 (MLP), :123-133 (attention block), :170-177
(decoder layer: x + attn(norm(x)); h + mlp(norm(h))), :220-232 (model) and
:281-289 (lm_head) with the oracle ops, against a paged KV cache laid out like
src/memory/kv_cache.cpp:15-27 ([n_slots, n_kv_heads, head_dim] per layer).

Used by tests (small shapes) and by bench.py's cpu_baseline / --impl reference
leg (Llama-3-8B shapes, a bounded number of layers).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import ops, quant


@dataclass
class LlamaConfig:
    hidden: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    n_kv_heads: int = 8
    head_dim: int = 128
    inter: int = 14336
    vocab: int = 128256
    rope_theta: float = 500000.0
    rms_eps: float = 1e-5
    max_pos: int = 8192
    # llama3 rope scaling (models/meta/llama.h:362-368); None = unscaled
    rope_scaling: Optional[Dict[str, float]] = field(
        default_factory=lambda: dict(factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                     original_max_position_embeddings=8192))
    group_size: int = 128
    quant: str = "awq"  # "awq" | "gptq" | "none"


def inv_freq_for(cfg: LlamaConfig) -> torch.Tensor:
    f = ops.compute_default_inv_freq(cfg.head_dim, cfg.rope_theta)
    if cfg.rope_scaling:
        rs = cfg.rope_scaling
        f = ops.apply_llama3_rope_scaling(f, rs["factor"], rs["low_freq_factor"],
                                          rs["high_freq_factor"],
                                          int(rs["original_max_position_embeddings"]))
    return f


class Linear:
    """dense bf16 weight W[K,N] (already dequantised for the quantised variants)."""

    def __init__(self, w: torch.Tensor):
        self.w = w
        self._w32 = None   # fp32 copy made on first use: the CPU path computes in fp32
                           # (llm_engine.cpp:28-31), so the conversion is not part of a step

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self._w32 is None:
            self._w32 = self.w.to(torch.float32)
        # == quant.w4a16_gemm(x, self.w): bf16 inputs, fp32 accumulation, one rounding
        return (x.to(torch.float32) @ self._w32).to(x.dtype)


def decoder_layer(x, positions, layer, cfg: LlamaConfig, cos_sin, k_cache, v_cache,
                  slot_ids, meta):
    """One llama.h:170-177 layer.  `layer` = dict of Linear + norm weights.  Mutates caches."""
    T = x.shape[0]
    H, Hkv, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
    n1 = ops.rms_norm(x, layer["input_norm"], cfg.rms_eps)
    qkv = layer["qkv"](n1)
    q = qkv[:, : H * D].reshape(T, H, D)
    k = qkv[:, H * D: (H + Hkv) * D].reshape(T, Hkv, D)
    v = qkv[:, (H + Hkv) * D:].reshape(T, Hkv, D)
    q, k = ops.rope(q, k, positions, cos_sin, D, interleaved=False)
    ops.kv_write(slot_ids, k, v, k_cache, v_cache)
    attn = ops.paged_attention(q, k_cache, v_cache, meta["q_cu_lens"], meta["kv_cu_lens"],
                               meta["block_table"], meta["block_cu_lens"], meta["block_size"],
                               sm_scale=D ** -0.5)
    o = layer["o"](attn.reshape(T, H * D))
    h = (x.float() + o.float()).to(x.dtype)
    n2 = ops.rms_norm(h, layer["post_norm"], cfg.rms_eps)
    gu = layer["gate_up"](n2)
    act = ops.silu_mul(gu[:, : cfg.inter], gu[:, cfg.inter:])
    d = layer["down"](act)
    return (h.float() + d.float()).to(x.dtype)


def decode_step(tokens, positions, model, cfg: LlamaConfig, k_caches, v_caches, slot_ids, meta):
    """models/meta/llama.h:220-232 + logits (:281-289).  Returns logits [T, vocab]."""
    h = model["embed"][tokens.long()]
    for i, layer in enumerate(model["layers"]):
        h = decoder_layer(h, positions, layer, cfg, model["cos_sin"], k_caches[i], v_caches[i],
                          slot_ids, meta)
    h = ops.rms_norm(h, model["final_norm"], cfg.rms_eps)
    return quant.w4a16_gemm(h, model["lm_head"])
