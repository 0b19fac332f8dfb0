"""Plugin-level face of the B200 decode path: mirrors of the `src/layers` interfaces the
reference's models are written against, backed by scalellm_b200.kernels.

    InputParameters          src/models/parameters.h:11-56
    KVCache                  src/memory/kv_cache.{h,cpp}
    AttentionHandler         src/layers/attention/handler.h:15-65
    B200AttnHandler          drop-in for ScaleAttnHandler (scale_attn_handler.cpp:24-84)
    Attention                src/layers/attention/attention.cpp:22-46
    RMSNorm                  src/layers/normalization.h:114-139
    RotaryEmbedding          src/layers/pos_embedding.cpp:183-215 (kernel variant)
    QuantArgs                src/layers/quantization/quant_args.h:10-33
    Column/RowParallelQLinear  qlinear_awq_marlin_impl.cpp:129-365, qlinear_gptq_marlin_impl.cpp
    Column/RowParallelLinear   src/layers/linear/parallel_linear.cpp:221-308 (dense bf16 -> cuBLASLt)
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import kernels
from .model_parallel import (ParallelArgs, gather_from_model_parallel_region,
                             reduce_from_model_parallel_region, scatter_to_model_parallel_region,
                             shard_range)


# ---------------------------------------------------------------------------
# step metadata + KV cache
# ---------------------------------------------------------------------------
@dataclass
class InputParameters:
    num_sequences: int = 0
    q_cu_seq_lens: Optional[torch.Tensor] = None    # int32 [n_seq+1]
    kv_cu_seq_lens: Optional[torch.Tensor] = None   # int32 [n_seq+1]
    kv_max_seq_len: int = 0
    q_max_seq_len: int = 0
    new_cache_slots: Optional[torch.Tensor] = None  # int32 [n_tokens]
    block_tables: Optional[torch.Tensor] = None     # int32 [n_blocks] first-slot ids (batch.cpp:206-209)
    cu_block_lens: Optional[torch.Tensor] = None    # int32 [n_seq+1]

    def to(self, device) -> "InputParameters":
        mv = lambda t: None if t is None else t.to(device)
        return InputParameters(self.num_sequences, mv(self.q_cu_seq_lens), mv(self.kv_cu_seq_lens),
                               self.kv_max_seq_len, self.q_max_seq_len, mv(self.new_cache_slots),
                               mv(self.block_tables), mv(self.cu_block_lens))


class KVCache:
    """Per-layer K and V tensors [n_blocks*block_size, n_kv_heads, head_dim] (kv_cache.cpp:15-27)."""

    def __init__(self, n_blocks: int, block_size: int, n_kv_heads: int, head_dim: int,
                 dtype: torch.dtype, device):
        self._block_size = block_size
        self.key_cache = torch.empty((n_blocks * block_size, n_kv_heads, head_dim), dtype=dtype,
                                     device=device)
        self.value_cache = torch.empty_like(self.key_cache)

    def block_size(self) -> int:
        return self._block_size

    def empty(self) -> bool:
        return self.key_cache.numel() == 0

    def get_kv_cache(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.key_cache, self.value_cache

    def set_kv_cache(self, slot_ids: torch.Tensor, keys: torch.Tensor, values: torch.Tensor) -> None:
        kernels.set_kv_cache(slot_ids, keys, values, self.key_cache, self.value_cache)

    def get_kv_by_slots(self, slot_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return kernels.get_kv_cache(slot_ids, self.key_cache, self.value_cache)


# ---------------------------------------------------------------------------
# rotary embedding
# ---------------------------------------------------------------------------
def compute_default_inv_freq(rotary_dim: int, theta: float) -> torch.Tensor:
    """pos_embedding.cpp:75-81"""
    sl = torch.arange(0, rotary_dim, 2, dtype=torch.float32)
    return 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float32), sl / rotary_dim)


def apply_llama3_rope_scaling(inv_freq: torch.Tensor, factor: float, low_freq_factor: float,
                              high_freq_factor: float, old_context_len: int) -> torch.Tensor:
    """pos_embedding.cpp:83-109 (vectorised, float32)."""
    f = inv_freq.to(torch.float32)
    wavelen = (2 * math.pi) / f
    low_wl = old_context_len / low_freq_factor
    high_wl = old_context_len / high_freq_factor
    smooth = (old_context_len / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor)
    mid = (1 - smooth) * f / factor + smooth * f
    out = torch.where(wavelen < high_wl, f, torch.where(wavelen > low_wl, f / factor, mid))
    return out.to(torch.float32)


class RotaryEmbedding:
    """RotaryEmbeddingKernel: cos|sin cache in the model dtype, in-place kernel."""

    def __init__(self, rotary_dim: int, max_position_embeddings: int, inv_freq: torch.Tensor,
                 interleaved: bool, dtype: torch.dtype, device):
        self.rotary_dim, self.interleaved = rotary_dim, interleaved
        t = torch.arange(0, max_position_embeddings, dtype=torch.float32)
        freqs = torch.einsum("i,j->ij", t, inv_freq.to(torch.float32).cpu())
        self.cos_sin_cache = torch.cat([freqs.cos(), freqs.sin()], dim=-1).to(dtype).to(device)

    def forward(self, query: torch.Tensor, key: torch.Tensor,
                positions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        kernels.apply_rotary_pos_emb(query, key, positions, self.cos_sin_cache, self.rotary_dim,
                                     self.interleaved)
        return query, key

    __call__ = forward


# ---------------------------------------------------------------------------
# attention handler plugin
# ---------------------------------------------------------------------------
class AttentionHandler:
    """handler.h:15-65"""

    def get_estimate_workspace_size(self) -> int:
        return -1

    def set_workspace(self, workspace: torch.Tensor) -> None:
        pass

    def apply_pos_emb(self, query, key, positions):
        raise NotImplementedError

    def batch_decode(self, query, kv_cache: KVCache, input_params: InputParameters,
                     sliding_window: int, output: torch.Tensor) -> None:
        raise NotImplementedError

    def append_kv_cache(self, kv_cache: KVCache, key, value, input_params: InputParameters) -> None:
        raise NotImplementedError


class B200AttnHandler(AttentionHandler):
    """The `--attention_handler=b200` plugin.  Same three calls as ScaleAttnHandler; when the
    caller uses `Attention.forward` the rope + cache write are fused into one launch."""

    def __init__(self, sm_scale: float, logits_soft_cap: float = 0.0,
                 alibi_slopes: Optional[torch.Tensor] = None,
                 pos_emb: Optional[RotaryEmbedding] = None):
        self.sm_scale, self.logits_soft_cap = sm_scale, logits_soft_cap
        self.alibi_slopes, self.pos_emb = alibi_slopes, pos_emb
        self._workspace: Optional[torch.Tensor] = None

    @classmethod
    def create_handler_with_rope(cls, head_dim: int, rotary_dim: int, max_position: int,
                                 inv_freq: torch.Tensor, interleaved: bool, dtype, device,
                                 logits_soft_cap: float = 0.0) -> "B200AttnHandler":
        rope = RotaryEmbedding(rotary_dim, max_position, inv_freq, interleaved, dtype, device)
        return cls(1.0 / math.sqrt(head_dim), logits_soft_cap, None, rope)

    def set_workspace(self, workspace: torch.Tensor) -> None:
        self._workspace = workspace

    def apply_pos_emb(self, query, key, positions):
        if positions is not None and self.pos_emb is not None:
            return self.pos_emb(query, key, positions)
        return query, key

    def append_kv_cache(self, kv_cache, key, value, input_params) -> None:
        if not kv_cache.empty():
            kv_cache.set_kv_cache(input_params.new_cache_slots, key, value)

    def apply_pos_emb_and_append(self, query, key, value, positions, kv_cache: KVCache,
                                 input_params: InputParameters) -> None:
        if self.pos_emb is None or positions is None or kv_cache.empty():
            self.apply_pos_emb(query, key, positions)
            self.append_kv_cache(kv_cache, key, value, input_params)
            return
        kernels.rope_and_set_kv_cache(query, key, value, positions, self.pos_emb.cos_sin_cache,
                                      input_params.new_cache_slots, kv_cache.key_cache,
                                      kv_cache.value_cache, self.pos_emb.rotary_dim,
                                      self.pos_emb.interleaved)

    def batch_decode(self, query, kv_cache, input_params, sliding_window, output) -> None:
        kc, vc = kv_cache.get_kv_cache()
        kernels.paged_kv_varlen_mha(output, query, kc, vc, input_params.q_cu_seq_lens,
                                    input_params.kv_cu_seq_lens, input_params.block_tables,
                                    input_params.cu_block_lens, self.alibi_slopes,
                                    kv_cache.block_size(), input_params.q_max_seq_len,
                                    input_params.kv_max_seq_len, self.sm_scale,
                                    self.logits_soft_cap, sliding_window,
                                    workspace=self._workspace)


class Attention:
    """attention.cpp:22-46"""

    def __init__(self, n_heads: int, n_kv_heads: int, head_dim: int, handler: AttentionHandler,
                 sliding_window: int = -1, fuse_rope_kv: bool = True):
        assert n_heads % n_kv_heads == 0
        self.n_heads, self.n_kv_heads, self.head_dim = n_heads, n_kv_heads, head_dim
        self.handler, self.sliding_window, self.fuse = handler, sliding_window, fuse_rope_kv

    def forward(self, query, key, value, positions, kv_cache: KVCache,
                input_params: InputParameters) -> torch.Tensor:
        T = query.size(0)
        q = query.view(T, self.n_heads, self.head_dim)
        k = key.view(T, self.n_kv_heads, self.head_dim)
        v = value.view(T, self.n_kv_heads, self.head_dim)
        if self.fuse and isinstance(self.handler, B200AttnHandler):
            self.handler.apply_pos_emb_and_append(q, k, v, positions, kv_cache, input_params)
        else:
            q, k = self.handler.apply_pos_emb(q, k, positions)
            self.handler.append_kv_cache(kv_cache, k, v, input_params)
        out = torch.empty((T, self.n_heads, self.head_dim), dtype=q.dtype, device=q.device)
        self.handler.batch_decode(q, kv_cache, input_params, self.sliding_window, out)
        return out.view(T, -1)

    def supports_partials(self) -> bool:
        h = self.handler
        return (self.fuse and isinstance(h, B200AttnHandler) and h.pos_emb is not None)

    def forward_partials(self, qkv_partials, positions, kv_cache: KVCache,
                         input_params: InputParameters, dtype=torch.bfloat16) -> torch.Tensor:
        """Same as forward, with q | k | v delivered as the qkv GEMM's stream-K partials: their
        reduction, RoPE and the KV-slot write are one launch."""
        h = self.handler
        qkv = kernels.rope_and_set_kv_cache_splitk(
            qkv_partials, self.n_heads, self.n_kv_heads, self.head_dim, positions,
            h.pos_emb.cos_sin_cache, input_params.new_cache_slots, kv_cache.key_cache,
            kv_cache.value_cache, h.pos_emb.rotary_dim, h.pos_emb.interleaved, dtype)
        T = qkv.size(0)
        q = qkv[:, : self.n_heads * self.head_dim].view(T, self.n_heads, self.head_dim)
        out = torch.empty((T, self.n_heads, self.head_dim), dtype=qkv.dtype, device=qkv.device)
        h.batch_decode(q, kv_cache, input_params, self.sliding_window, out)
        return out.view(T, -1)

    __call__ = forward


# ---------------------------------------------------------------------------
# norm
# ---------------------------------------------------------------------------
class RMSNorm:
    def __init__(self, dim: int, eps: float, dtype: torch.dtype, device):
        self.weight = torch.ones(dim, dtype=dtype, device=device)
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(x)
        kernels.rms_norm(out, x, self.weight, self.eps)
        return out

    def forward_residual(self, x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        """residual += x (in place), returns norm(residual): kernel::rms_norm_residual."""
        out = torch.empty_like(x)
        kernels.rms_norm_residual(out, residual, x, self.weight, self.eps)
        return out

    def forward_residual_partials(self, partials, residual: torch.Tensor, pa=None) -> torch.Tensor:
        """Same, with x delivered as the producing GEMM's stream-K partials (kernels.W4Partials):
        the cross-CTA reduction of the GEMM is fused into this kernel.  With a tensor-parallel
        group (pa.world_size > 1) the partials are this rank's share and the TP all-reduce is
        fused in as well."""
        if pa is not None and pa.world_size > 1:
            return pa.process_group.allreduce_partials_norm(partials, residual, self.weight, self.eps)
        out = torch.empty_like(residual)
        kernels.rms_norm_residual_splitk(out, residual, partials, self.weight, self.eps)
        return out

    __call__ = forward


# ---------------------------------------------------------------------------
# linear layers
# ---------------------------------------------------------------------------
@dataclass
class QuantArgs:
    """quant_args.h:10-33"""
    quant_method: str = ""      # "awq" | "gptq" | ""
    bits: int = 0
    group_size: int = 0
    desc_act: bool = False
    is_sym: bool = False
    zero_point: bool = True


_DENSE_PREFILL_ROWS = 256   # above this many rows: dequant + bf16 GEMM (B200_W4_PREFILL_DENSE=0: streaming kernel)


def _check_quant(qa: QuantArgs, in_features: int, out_features: int) -> None:
    # qlinear_awq_marlin_impl.cpp:28-31,150-151
    if qa.bits != 4:
        raise NotImplementedError("B200 W4A16 path supports 4-bit weights only")
    if qa.group_size not in (-1, 32, 64, 128):
        raise NotImplementedError(f"group_size {qa.group_size} not in (-1, 32, 64, 128)")
    if in_features % 128 or out_features % 128:
        raise ValueError("W4A16 needs in_features % 128 == 0 and out_features % 128 == 0 per shard")


class _QLinearBase:
    def __init__(self, in_features: int, out_features: int, bias: bool, qa: QuantArgs,
                 device):
        _check_quant(qa, in_features, out_features)
        self.K, self.N, self.qa, self.device = in_features, out_features, qa, device
        self.packed: Optional[torch.Tensor] = None
        self.perm: Optional[torch.Tensor] = None     # GPTQ act-order: activation column order
        self.bias: Optional[torch.Tensor] = None
        self._has_bias = bias
        self._ckpt: Dict[str, torch.Tensor] = {}

    # weights arrive in checkpoint format; packing is lazy like repack_weight
    # (qlinear_awq_marlin_impl.cpp:99-125,232-235)
    def _set_shard(self, qweight, qzeros, scales, bias=None, g_idx=None) -> None:
        self._ckpt = dict(qweight=qweight.contiguous().to(self.device),
                          qzeros=None if qzeros is None else qzeros.contiguous().to(self.device),
                          scales=scales.contiguous().to(self.device).to(torch.bfloat16),
                          g_idx=None if g_idx is None else g_idx.contiguous().to(self.device))
        self.packed, self.perm = None, None
        if bias is not None:
            self.bias = bias.to(self.device).to(torch.bfloat16)

    def verify_loaded_weights(self) -> None:
        if not self._ckpt and self.packed is None:
            raise RuntimeError("qweight / qzeros / scales not loaded")

    def _ensure_packed(self) -> None:
        if self.packed is not None:
            return
        self.verify_loaded_weights()
        c = self._ckpt
        g = self.qa.group_size
        if self.qa.quant_method == "awq":
            self.packed = kernels.w4a16_prepack_awq(c["qweight"], c["qzeros"], c["scales"], g)
        else:
            qz = None if self.qa.is_sym else c["qzeros"]
            if self.qa.desc_act and c.get("g_idx") is not None:
                # act-order: rows sorted by group at pack time, activation columns gathered with the
                # same perm at run time (qlinear_gptq_marlin_impl.cpp:43-56, gptq_gemm.cu:66-104)
                self.packed, self.perm = kernels.w4a16_prepack_gptq_actorder(
                    c["qweight"], qz, c["scales"], c["g_idx"], g, zeros_plus_one=True)
            else:
                self.packed = kernels.w4a16_prepack_gptq(c["qweight"], qz, c["scales"], g,
                                                         zeros_plus_one=True)
        # The first GEMM follows on the same stream with the programmatic-launch attribute and its
        # weight producer skips griddepcontrol.wait (weights are constants): make the prepack's
        # writes visible first.  One-time, never inside a graph capture (it allocates).
        torch.cuda.current_stream().synchronize()
        self._ckpt = {}

    def _gemm(self, x: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
        self._ensure_packed()
        x2 = x.reshape(-1, x.shape[-1])
        if self.perm is not None:
            x2 = kernels.permute_cols(x2, self.perm)
        if x2.shape[0] > _DENSE_PREFILL_ROWS and os.environ.get("B200_W4_PREFILL_DENSE", "1") != "0":
            # Prefill-sized batches are compute bound: the streaming kernel would re-read the int4
            # weights once per 128 rows.  Dequantise once (same bf16 values as the fused kernel,
            # b200_w4a16_dequant) and let the library bf16 GEMM do the rest (a plain library GEMM, like
            # the reference's dense layers; TTFT path: 51.6 -> 31.9 ms p50 on B200).
            w = kernels.w4a16_dequant(self.packed, self.K, self.N, self.qa.group_size)
            out = torch.matmul(x2, w)
            if bias is not None:
                out = out + bias
            return out.view(*x.shape[:-1], self.N)
        out = kernels.w4a16_gemm(x2, self.packed, self.N, self.qa.group_size, bias=bias)
        return out.view(*x.shape[:-1], self.N)


def _shard_qtensors(sd: Dict[str, torch.Tensor], qa: QuantArgs, dim: int, rank: int, world: int,
                    K: int, N: int):
    """Shard checkpoint tensors.  dim=1: column parallel (split N); dim=0: row parallel (split K).
    AWQ: qweight [K, N/8], qzeros [K/g, N/8], scales [K/g, N];
    GPTQ: qweight [K/8, N], qzeros [K/g, N/8], scales [K/g, N]."""
    qw, qz, sc = sd["qweight"], sd.get("qzeros"), sd["scales"]
    if world == 1:
        return qw, qz, sc
    g = K if qa.group_size <= 0 else qa.group_size
    if dim == 1:
        r_n = shard_range(N, rank, world)
        r_n8 = slice(r_n.start // 8, r_n.stop // 8)
        qw_s = qw[:, r_n8] if qa.quant_method == "awq" else qw[:, r_n]
        return qw_s, None if qz is None else qz[:, r_n8], sc[:, r_n]
    r_k = shard_range(K, rank, world)
    assert (K // world) % g == 0, "row-parallel shard must align to quant groups"  # :287
    r_g = slice(r_k.start // g, r_k.stop // g)
    qw_s = qw[r_k, :] if qa.quant_method == "awq" else qw[r_k.start // 8: r_k.stop // 8, :]
    return qw_s, None if qz is None else qz[r_g, :], sc[r_g, :]


class ColumnParallelQLinear(_QLinearBase):
    """Y = X W (+b), W split along N (qlinear_awq_marlin_impl.cpp:129-257)."""

    def __init__(self, in_features: int, out_features: int, bias: bool, gather_output: bool,
                 qa: QuantArgs, pa: ParallelArgs, device):
        assert out_features % pa.world_size == 0
        super().__init__(in_features, out_features // pa.world_size, bias, qa, device)
        self.full_N, self.gather_output, self.pa = out_features, gather_output, pa

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        qw, qz, sc = _shard_qtensors(sd, self.qa, 1, self.pa.rank, self.pa.world_size, self.K,
                                     self.full_N)
        b = sd.get("bias")
        if b is not None:
            b = b[shard_range(self.full_N, self.pa.rank, self.pa.world_size)]
        self._set_shard(qw, qz, sc, b, sd.get("g_idx") if self.qa.desc_act else None)

    def supports_partials(self, n_rows: int) -> bool:
        """Partials output for a fused consumer (rope / silu*mul): the column shard is rank-local,
        so this also holds under tensor parallelism (no gather, no bias)."""
        return (self.bias is None and 0 < n_rows <= 128 and not self.qa.desc_act
                and not (self.pa.world_size > 1 and self.gather_output))

    def forward_partials(self, x: torch.Tensor) -> "kernels.W4Partials":
        self._ensure_packed()
        x2 = x.reshape(-1, x.shape[-1])
        return kernels.w4a16_gemm_splitk(x2, self.packed, self.N, self.qa.group_size)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = self._gemm(x, self.bias)
        if self.pa.world_size > 1 and self.gather_output:
            out = gather_from_model_parallel_region(out, self.pa)
        return out

    __call__ = forward


class RowParallelQLinear(_QLinearBase):
    """Y = sum_ranks X_r W_r (+b), W split along K (qlinear_awq_marlin_impl.cpp:260-365)."""

    def __init__(self, in_features: int, out_features: int, bias: bool,
                 input_is_parallelized: bool, qa: QuantArgs, pa: ParallelArgs, device):
        assert in_features % pa.world_size == 0
        super().__init__(in_features // pa.world_size, out_features, bias, qa, device)
        self.full_K, self.input_is_parallelized, self.pa = in_features, input_is_parallelized, pa

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        qw, qz, sc = _shard_qtensors(sd, self.qa, 0, self.pa.rank, self.pa.world_size,
                                     self.full_K, self.N)
        g_idx = sd.get("g_idx") if self.qa.desc_act else None
        if g_idx is not None and self.pa.world_size > 1:
            raise NotImplementedError("act-order (desc_act) on a K-sharded (row-parallel, TP > 1) weight: "
                                      "its groups are split across ranks (is_k_full = false)")
        self._set_shard(qw, qz, sc, sd.get("bias"), g_idx)

    def supports_partials(self, n_rows: int) -> bool:
        """Partials output (the GEMM's cross-CTA reduction fused into the consumer norm; under TP
        the all-reduce is fused into the same kernel) — no bias."""
        if self.bias is not None or not 0 < n_rows <= 128 or self.qa.desc_act:
            return False
        if self.pa.world_size == 1:
            return True
        pg = self.pa.process_group
        return (os.environ.get("B200_FUSE_AR_NORM", "1") != "0" and hasattr(pg, "supports_partials_norm")
                and pg.supports_partials_norm(n_rows, self.N, torch.bfloat16))

    def forward_partials(self, x: torch.Tensor) -> "kernels.W4Partials":
        self._ensure_packed()
        x2 = x.reshape(-1, x.shape[-1])
        return kernels.w4a16_gemm_splitk(x2, self.packed, self.N, self.qa.group_size)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.input_is_parallelized:
            x = scatter_to_model_parallel_region(x, self.pa)
        if self.pa.world_size == 1:
            return self._gemm(x, self.bias)
        x2 = x.reshape(-1, x.shape[-1])
        pg = self.pa.process_group
        if 0 < x2.shape[0] <= 128 and hasattr(pg, "allreduce_partials") and x2.is_cuda:
            # split-K partials go straight into the NVLink all-reduce (no GEMM fix-up pass)
            self._ensure_packed()
            parts = kernels.w4a16_gemm_splitk(x2, self.packed, self.N, self.qa.group_size)
            out = pg.allreduce_partials(parts, x.dtype).view(*x.shape[:-1], self.N)
        else:
            out = self._gemm(x, None)
            out = reduce_from_model_parallel_region(out, self.pa)
        if self.bias is not None:  # bias after the reduction (:360-363)
            out = out + self.bias
        return out

    __call__ = forward


def _dense_linear(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """F::linear of the reference's dense layers (parallel_linear.cpp:258,299): the tcgen05 stream
    kernel of csrc/dense.cu for bf16 on CUDA (B200_DENSE_IMPL=cublas: the library GEMM, for A/B)."""
    if os.environ.get("B200_DENSE_IMPL") != "cublas" and kernels.dense_supported(x, weight):
        x2 = x.reshape(-1, x.shape[-1])
        return kernels.dense_gemm(x2, weight).view(*x.shape[:-1], weight.shape[0])
    return F.linear(x, weight)


def _dense_partials_ok(x_rows: int, weight: torch.Tensor) -> bool:
    return (os.environ.get("B200_DENSE_IMPL") != "cublas" and weight.is_cuda
            and weight.dtype == torch.bfloat16 and 0 < x_rows <= 128
            and weight.shape[0] % 128 == 0 and weight.shape[1] % 128 == 0)


class ColumnParallelLinear:
    """Dense bf16 column-parallel linear (parallel_linear.cpp:221-263): csrc/dense.cu."""

    def __init__(self, in_features: int, out_features: int, gather_output: bool, pa: ParallelArgs,
                 dtype, device):
        assert out_features % pa.world_size == 0
        self.weight = torch.empty((out_features // pa.world_size, in_features), dtype=dtype,
                                  device=device)
        self.full_N, self.gather_output, self.pa = out_features, gather_output, pa

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        w = sd["weight"][shard_range(self.full_N, self.pa.rank, self.pa.world_size)]
        self.weight.copy_(w)

    def forward_local(self, x: torch.Tensor) -> torch.Tensor:
        """This rank's column shard of the output (no gather)."""
        return _dense_linear(x, self.weight)

    def supports_partials(self, n_rows: int) -> bool:
        """Partials for a fused consumer (rope / silu*mul), like ColumnParallelQLinear."""
        return _dense_partials_ok(n_rows, self.weight) and not (self.pa.world_size > 1 and self.gather_output)

    def forward_partials(self, x: torch.Tensor) -> "kernels.W4Partials":
        return kernels.dense_gemm_splitk(x.reshape(-1, x.shape[-1]), self.weight)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = _dense_linear(x, self.weight)
        if self.pa.world_size > 1 and self.gather_output:
            out = gather_from_model_parallel_region(out, self.pa)
        return out

    __call__ = forward


class RowParallelLinear:
    """Dense bf16 row-parallel linear (parallel_linear.cpp:266-308)."""

    def __init__(self, in_features: int, out_features: int, input_is_parallelized: bool,
                 pa: ParallelArgs, dtype, device):
        assert in_features % pa.world_size == 0
        self.weight = torch.empty((out_features, in_features // pa.world_size), dtype=dtype,
                                  device=device)
        self.full_K, self.input_is_parallelized, self.pa = in_features, input_is_parallelized, pa

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        w = sd["weight"][:, shard_range(self.full_K, self.pa.rank, self.pa.world_size)]
        self.weight.copy_(w)

    def supports_partials(self, n_rows: int) -> bool:
        """Partials for the fused residual + RMSNorm consumer (under TP: fused into the all-reduce),
        like RowParallelQLinear."""
        if not _dense_partials_ok(n_rows, self.weight):
            return False
        if self.pa.world_size == 1:
            return True
        pg = self.pa.process_group
        return (os.environ.get("B200_FUSE_AR_NORM", "1") != "0" and hasattr(pg, "supports_partials_norm")
                and pg.supports_partials_norm(n_rows, self.weight.shape[0], torch.bfloat16))

    def forward_partials(self, x: torch.Tensor) -> "kernels.W4Partials":
        return kernels.dense_gemm_splitk(x.reshape(-1, x.shape[-1]), self.weight)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.input_is_parallelized:
            x = scatter_to_model_parallel_region(x, self.pa)
        out = _dense_linear(x, self.weight)
        return reduce_from_model_parallel_region(out, self.pa)

    __call__ = forward
