"""Operator-level face of the B200 decode path: the reference's `src/kernels` free
functions, same names / argument order / in-place conventions, on torch CUDA tensors.

    llm::kernel::rms_norm, rms_norm_residual     src/kernels/layernorm_kernels.h:6-19
    llm::kernel::apply_rotary_pos_emb            src/kernels/pos_embedding_kernels.h:7-13
    llm::kernel::set_kv_cache                    src/kernels/kv_cache_kernels.h:6-11
    llm::kernel::silu, silu_with_mul             src/kernels/activation_kernels.h:6-14
    llm::paged_kv_varlen_mha                     src/kernels/attention/attn_api.h:12-27
    marlin::awq_repack / gptq_repack / gptq_gemm src/kernels/quantization/marlin.h:17-37
        -> w4a16_prepack_awq / w4a16_prepack_gptq / w4a16_gemm (our own packed layout)

Every call goes through the C ABI of libb200decode.so on the current CUDA stream.
CPU tensors are rejected — there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import B200Error, check  # noqa: F401

_DT = {torch.bfloat16: _lib.B200_BF16, torch.float16: _lib.B200_FP16, torch.float32: _lib.B200_FP32}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}") from None


def _cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("scalellm_b200 kernels require CUDA tensors (no CPU fallback)")


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_i32(*ts: torch.Tensor) -> None:
    for t in ts:
        if t.dtype != torch.int32:
            raise TypeError("index / length tensors must be int32 (reference: torch::kInt)")
        if not t.is_contiguous():
            raise ValueError("index / length tensors must be contiguous")


# ---------------------------------------------------------------------------
# workspace cache (per device).  Allocate BEFORE CUDA-graph capture by calling the
# op once eagerly, exactly like the reference warms its kernels before capture.
# ---------------------------------------------------------------------------
_WS: Dict[Tuple[str, int], torch.Tensor] = {}
_WS_RETIRED: list = []   # outgrown workspaces: a captured CUDA graph may still hold their addresses


def _workspace(kind: str, device: torch.device, nbytes: int, zero: bool = False) -> torch.Tensor:
    """Grow-only: a larger request allocates a new buffer, but the old one is never freed — a
    CUDA graph captured earlier (GraphedStep) has its raw pointer baked in and would otherwise
    replay into memory the caching allocator has handed to someone else."""
    key = (kind, device.index if device.index is not None else torch.cuda.current_device())
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _WS_RETIRED.append(ws)
        ws = (torch.zeros if zero else torch.empty)(max(nbytes, 1), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


# ---------------------------------------------------------------------------
# norms
# ---------------------------------------------------------------------------
def rms_norm(out: torch.Tensor, input: torch.Tensor, weight: torch.Tensor, epsilon: float) -> None:
    _cuda(out, input, weight)
    assert input.is_contiguous() and out.is_contiguous(), "tensors must be contiguous"
    n = input.shape[-1]
    rows = input.numel() // n
    if rows == 0:
        return
    check(_lib.load().b200_rms_norm(_p(out), _p(input), _p(weight), rows, n, epsilon, _dt(input),
                                    _stream()))


def rms_norm_residual(out: torch.Tensor, residual: torch.Tensor, input: torch.Tensor,
                      weight: torch.Tensor, epsilon: float) -> None:
    _cuda(out, residual, input, weight)
    assert input.is_contiguous() and out.is_contiguous() and residual.is_contiguous()
    n = input.shape[-1]
    rows = input.numel() // n
    if rows == 0:
        return
    check(_lib.load().b200_rms_norm_residual(_p(out), _p(residual), _p(input), _p(weight), rows, n,
                                             epsilon, _dt(input), _stream()))


# ---------------------------------------------------------------------------
# rotary embedding / kv cache
# ---------------------------------------------------------------------------
def apply_rotary_pos_emb(querys: torch.Tensor, keys: torch.Tensor, positions: torch.Tensor,
                         cos_sin: torch.Tensor, rotary_dim: int, interleaved: bool) -> None:
    """In place on querys [T,H,D] and keys [T,Hkv,D] (heads dense, token stride free)."""
    _cuda(querys, keys, positions, cos_sin)
    _check_i32(positions)
    assert querys.stride(-1) == 1 and querys.stride(-2) == querys.size(-1)
    assert keys.stride(-1) == 1 and keys.stride(-2) == keys.size(-1)
    assert cos_sin.is_contiguous() and cos_sin.dtype == querys.dtype
    T, H, D = querys.shape[-3], querys.shape[-2], querys.shape[-1]
    check(_lib.load().b200_rope_inplace(_p(querys), _p(keys), _p(positions), _p(cos_sin), T, H,
                                        keys.shape[-2], D, rotary_dim, querys.stride(-3),
                                        keys.stride(-3), int(interleaved), _dt(querys), _stream()))


def set_kv_cache(slot_ids: torch.Tensor, keys: torch.Tensor, values: torch.Tensor,
                 key_cache: torch.Tensor, value_cache: torch.Tensor) -> None:
    _cuda(slot_ids, keys, values, key_cache, value_cache)
    _check_i32(slot_ids)
    assert keys.stride(-1) == 1 and keys.stride(-2) == keys.size(-1)
    assert values.stride(-1) == 1 and values.stride(-2) == values.size(-1)
    assert key_cache.is_contiguous() and value_cache.is_contiguous()
    T, Hkv, D = keys.shape[-3], keys.shape[-2], keys.shape[-1]
    check(_lib.load().b200_kv_write(_p(slot_ids), _p(keys), _p(values), _p(key_cache),
                                    _p(value_cache), T, Hkv, D, keys.stride(-3), values.stride(-3),
                                    _dt(keys), _stream()))


def get_kv_cache(slot_ids: torch.Tensor, key_cache: torch.Tensor,
                 value_cache: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """KVCache::get_kv_cache (src/memory/kv_cache.cpp:60-98) as a device gather."""
    _cuda(slot_ids, key_cache, value_cache)
    _check_i32(slot_ids)
    T = slot_ids.numel()
    Hkv, D = key_cache.shape[-2], key_cache.shape[-1]
    k = torch.empty((T, Hkv, D), dtype=key_cache.dtype, device=key_cache.device)
    v = torch.empty_like(k)
    check(_lib.load().b200_kv_gather(_p(slot_ids), _p(key_cache), _p(value_cache), _p(k), _p(v), T,
                                     Hkv, D, _dt(key_cache), _stream()))
    return k, v


def rope_and_set_kv_cache(querys: torch.Tensor, keys: torch.Tensor, values: torch.Tensor,
                          positions: torch.Tensor, cos_sin: torch.Tensor, slot_ids: torch.Tensor,
                          key_cache: torch.Tensor, value_cache: torch.Tensor, rotary_dim: int,
                          interleaved: bool) -> None:
    """Fused apply_rotary_pos_emb + set_kv_cache (bit-identical to the two calls)."""
    _cuda(querys, keys, values, positions, cos_sin, slot_ids, key_cache, value_cache)
    _check_i32(positions, slot_ids)
    assert querys.stride(-1) == 1 and querys.stride(-2) == querys.size(-1)
    assert keys.stride(-1) == 1 and keys.stride(-2) == keys.size(-1)
    assert values.stride(-1) == 1 and values.stride(-2) == values.size(-1)
    assert key_cache.is_contiguous() and value_cache.is_contiguous()
    T, H, D = querys.shape[-3], querys.shape[-2], querys.shape[-1]
    check(_lib.load().b200_rope_kv_write(_p(querys), _p(keys), _p(values), _p(positions),
                                         _p(cos_sin), _p(slot_ids), _p(key_cache), _p(value_cache),
                                         T, H, keys.shape[-2], D, rotary_dim, querys.stride(-3),
                                         keys.stride(-3), values.stride(-3), int(interleaved),
                                         _dt(querys), _stream()))


# ---------------------------------------------------------------------------
# activations
# ---------------------------------------------------------------------------
def silu(input: torch.Tensor) -> torch.Tensor:
    _cuda(input)
    assert input.dim() == 2 and input.stride(1) == 1
    out = torch.empty((input.shape[0], input.shape[1]), dtype=input.dtype, device=input.device)
    check(_lib.load().b200_silu(_p(out), _p(input), input.shape[0], input.shape[1],
                                input.stride(0), _dt(input), _stream()))
    return out


def silu_with_mul(input: torch.Tensor) -> torch.Tensor:
    _cuda(input)
    assert input.dim() == 2 and input.is_contiguous()
    n = input.shape[1] // 2
    out = torch.empty((input.shape[0], n), dtype=input.dtype, device=input.device)
    check(_lib.load().b200_silu_mul(_p(out), _p(input), input.shape[0], n, _dt(input), _stream()))
    return out


# ---------------------------------------------------------------------------
# dense bf16 linear (A8): F::linear -> cuBLASLt of the reference, parallel_linear.cpp:256-263,294-308
# ---------------------------------------------------------------------------
def dense_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """Shapes / dtypes the tcgen05 dense kernel takes (rank-invariant: sizes, dtypes and strides of
    freshly allocated activations only)."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
            and x.dim() >= 2 and x.shape[-1] % 64 == 0 and weight.shape[0] % 8 == 0
            and weight.stride(1) == 1 and weight.stride(0) % 8 == 0)


def dense_gemm(a: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = A @ W^T (+ bias): A [M, K] bf16 (rows dense in K), W [N, K] bf16 as nn.Linear stores it."""
    _cuda(a, weight, bias)
    assert a.dim() == 2 and a.stride(1) == 1 and weight.dim() == 2 and weight.stride(1) == 1
    M, K = a.shape
    N = weight.shape[0]
    assert weight.shape[1] == K
    if a.stride(0) % 8 or a.data_ptr() % 16:
        a = a.contiguous()
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    if M == 0:
        return out
    ws = _workspace("dense", a.device, int(_lib.load().b200_dense_workspace_bytes(M, N, K)))
    check(_lib.load().b200_dense_gemm(_p(out), _p(a), _p(weight), _p(bias), M, N, K, a.stride(0),
                                      weight.stride(0), out.stride(0), _p(ws), ws.numel(), _stream()))
    return out


def dense_gemm_splitk(a: torch.Tensor, weight: torch.Tensor, poison: bool = False) -> "W4Partials":
    """Partials mode of the dense GEMM: fp32 partials [slots, M, N] in the int4 GEMM's stream-K
    format, for the same fused consumers.  M <= 128, N and K multiples of 128."""
    _cuda(a, weight)
    assert a.dim() == 2 and a.dtype == torch.bfloat16 and a.stride(1) == 1
    assert weight.dim() == 2 and weight.dtype == torch.bfloat16 and weight.stride(1) == 1
    M, K = a.shape
    N = weight.shape[0]
    if a.stride(0) % 8 or a.data_ptr() % 16:
        a = a.contiguous()
    slots = int(_lib.load().b200_dense_splitk_splits(M, N, K))
    assert slots > 0, "dense_gemm_splitk: N and K must be multiples of 128"
    partials = torch.empty((slots, M, N), dtype=torch.float32, device=a.device)
    if poison:
        partials.fill_(float("nan"))
    check(_lib.load().b200_dense_gemm_splitk(_p(partials), _p(a), _p(weight), M, N, K, a.stride(0),
                                             weight.stride(0), slots, _stream()))
    return W4Partials(partials, K)


# ---------------------------------------------------------------------------
# sampling tail: logits processors (src/kernels/sampling/sampling_kernels.h:7-29), in place
# ---------------------------------------------------------------------------
def apply_temperature_penalty(logits: torch.Tensor, temperatures: torch.Tensor) -> None:
    _cuda(logits, temperatures)
    assert logits.dim() == 2 and logits.is_contiguous() and temperatures.is_contiguous()
    assert temperatures.dtype == logits.dtype and temperatures.numel() == logits.shape[0]
    check(_lib.load().b200_apply_temperature(_p(logits), _p(temperatures), logits.shape[0], logits.shape[1],
                                             _dt(logits), _stream()))


def apply_repetition_penalty(logits: torch.Tensor, token_ids: torch.Tensor, token_ids_lens: torch.Tensor,
                             penalties: torch.Tensor) -> None:
    _cuda(logits, token_ids, token_ids_lens, penalties)
    assert logits.dim() == 2 and logits.is_contiguous() and token_ids.is_contiguous()
    assert token_ids.dtype == torch.int64 and token_ids_lens.dtype == torch.int32 and penalties.dtype == logits.dtype
    check(_lib.load().b200_apply_repetition_penalty(_p(logits), _p(token_ids), _p(token_ids_lens), _p(penalties),
                                                    logits.shape[0], logits.shape[1], token_ids.shape[1],
                                                    _dt(logits), _stream()))


def apply_frequency_presence_penalty(logits: torch.Tensor, token_ids: torch.Tensor, token_counts: torch.Tensor,
                                     token_ids_lens: torch.Tensor, frequency_penalties: torch.Tensor,
                                     presence_penalties: torch.Tensor) -> None:
    _cuda(logits, token_ids, token_counts, token_ids_lens, frequency_penalties, presence_penalties)
    assert logits.dim() == 2 and logits.is_contiguous() and token_ids.is_contiguous() and token_counts.is_contiguous()
    assert token_ids.dtype == torch.int64 and token_counts.dtype == torch.int32 and token_ids_lens.dtype == torch.int32
    check(_lib.load().b200_apply_frequency_presence_penalty(
        _p(logits), _p(token_ids), _p(token_counts), _p(token_ids_lens), _p(frequency_penalties),
        _p(presence_penalties), logits.shape[0], logits.shape[1], token_ids.shape[1], _dt(logits), _stream()))


def invoke_softmax(logits: torch.Tensor) -> None:
    _cuda(logits)
    assert logits.dim() == 2 and logits.is_contiguous()
    check(_lib.load().b200_softmax(_p(logits), logits.shape[0], logits.shape[1], _dt(logits), _stream()))


def apply_top_k_top_p(logits: torch.Tensor, top_k: Optional[torch.Tensor], top_p: Optional[torch.Tensor]) -> None:
    """TopKTopPLogitsProcessor::forward (src/sampling/logits_processor.h:243-276) in place, one launch:
    everything outside the top-k / top-p cut of its row becomes -inf.  top_k [batch] int64 (<= 0: off),
    top_p [batch] float32 (>= 1: off); either may be None."""
    _cuda(logits, top_k, top_p)
    assert logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype in (torch.bfloat16, torch.float16)
    assert top_k is None or (top_k.dtype == torch.int64 and top_k.is_contiguous() and top_k.numel() == logits.shape[0])
    assert top_p is None or (top_p.dtype == torch.float32 and top_p.is_contiguous() and top_p.numel() == logits.shape[0])
    check(_lib.load().b200_topk_topp_filter(_p(logits), _p(top_k), _p(top_p), logits.shape[0], logits.shape[1],
                                            logits.stride(0), _dt(logits), _stream()))


def _gelu(input: torch.Tensor, act: int, with_mul: bool) -> torch.Tensor:
    _cuda(input)
    assert input.dim() == 2 and input.stride(1) == 1 and (not with_mul or input.is_contiguous())
    n = input.shape[1] // 2 if with_mul else input.shape[1]
    out = torch.empty((input.shape[0], n), dtype=input.dtype, device=input.device)
    check(_lib.load().b200_gelu(_p(out), _p(input), input.shape[0], n, input.stride(0), act,
                                1 if with_mul else 0, _dt(input), _stream()))
    return out


def gelu_new(input: torch.Tensor) -> torch.Tensor:
    """activation_kernels.h:8 (tanh form, tanh.approx)"""
    return _gelu(input, 1, False)


def gelu_fast(input: torch.Tensor) -> torch.Tensor:
    return _gelu(input, 2, False)


def gelu_new_with_mul(input: torch.Tensor) -> torch.Tensor:
    return _gelu(input, 1, True)


def gelu_fast_with_mul(input: torch.Tensor) -> torch.Tensor:
    return _gelu(input, 2, True)


def gemma_rms_norm(out: torch.Tensor, input: torch.Tensor, weight: torch.Tensor, epsilon: float) -> None:
    """layernorm_kernels.h:11-14: out = (T)(x * rstd * (1 + w))"""
    _cuda(out, input, weight)
    assert input.is_contiguous() and out.is_contiguous()
    n = input.shape[-1]
    check(_lib.load().b200_gemma_rms_norm(_p(out), _p(input), _p(weight), input.numel() // n, n, epsilon,
                                          _dt(input), _stream()))


def layer_norm(out: torch.Tensor, input: torch.Tensor, weight: torch.Tensor,
               bias: Optional[torch.Tensor], epsilon: float) -> None:
    """layernorm_kernels.h:21-25 (bias may be None / undefined)"""
    _cuda(out, input, weight)
    assert input.is_contiguous() and out.is_contiguous()
    n = input.shape[-1]
    check(_lib.load().b200_layer_norm(_p(out), _p(input), _p(weight), None if bias is None else _p(bias),
                                      input.numel() // n, n, epsilon, _dt(input), _stream()))


def silu_mul(gate: torch.Tensor, up: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act_func(gate) * up on strided views of the fused gate_up output (llama.h:61-64)."""
    _cuda(gate, up)
    assert gate.dim() == 2 and gate.shape == up.shape and gate.stride(1) == 1 and up.stride(1) == 1
    if out is None:
        out = torch.empty(gate.shape, dtype=gate.dtype, device=gate.device)
    check(_lib.load().b200_silu_mul_strided(_p(out), _p(gate), _p(up), gate.shape[0], gate.shape[1],
                                            gate.stride(0), up.stride(0), _dt(gate), _stream()))
    return out


# ---------------------------------------------------------------------------
# paged attention
# ---------------------------------------------------------------------------
def paged_attn_workspace_bytes(batch: int, max_q_len: int, max_kv_len: int, n_heads: int,
                               n_kv_heads: int, head_dim: int) -> int:
    return int(_lib.load().b200_paged_attn_workspace_bytes(batch, max_q_len, max_kv_len, n_heads,
                                                           n_kv_heads, head_dim))


def paged_kv_varlen_mha(out: torch.Tensor, query: torch.Tensor, key_cache: torch.Tensor,
                        value_cache: torch.Tensor, q_cu_lens: torch.Tensor,
                        kv_cu_lens: torch.Tensor, block_table: torch.Tensor,
                        block_cu_lens: torch.Tensor, alibi_slopes: Optional[torch.Tensor],
                        block_size: int, max_q_len: int, max_kv_len: int, sm_scale: float,
                        logits_soft_cap: float, sliding_window: int,
                        workspace: Optional[torch.Tensor] = None) -> None:
    _cuda(out, query, key_cache, value_cache, q_cu_lens, kv_cu_lens, block_table, block_cu_lens,
          alibi_slopes)
    _check_i32(q_cu_lens, kv_cu_lens, block_table, block_cu_lens)
    assert query.stride(-1) == 1 and out.stride(-1) == 1
    assert key_cache.stride(-1) == 1 and value_cache.stride(-1) == 1
    assert key_cache.stride() == value_cache.stride() and key_cache.shape == value_cache.shape
    if alibi_slopes is not None:
        assert alibi_slopes.dtype == torch.float32 and alibi_slopes.is_contiguous()
    batch = q_cu_lens.numel() - 1
    H, D = query.shape[-2], query.shape[-1]
    Hkv = key_cache.shape[-2]
    need = paged_attn_workspace_bytes(batch, max_q_len, max_kv_len, H, Hkv, D)
    if workspace is None and need > 0:
        workspace = _workspace("attn", query.device, need)
    check(_lib.load().b200_paged_attn_decode(
        _p(out), _p(query), _p(key_cache), _p(value_cache), _p(q_cu_lens), _p(kv_cu_lens),
        _p(block_table), _p(block_cu_lens), _p(alibi_slopes), batch, H, Hkv, D,
        key_cache.shape[0], query.stride(0), query.stride(1), out.stride(0), out.stride(1),
        key_cache.stride(0), key_cache.stride(1), block_size, max_q_len, max_kv_len, sm_scale,
        logits_soft_cap, sliding_window, _p(workspace),
        0 if workspace is None else workspace.numel() * workspace.element_size(), _dt(query),
        _stream()))


# ---------------------------------------------------------------------------
# W4A16
# ---------------------------------------------------------------------------
def w4a16_packed_bytes(K: int, N: int, group_size: int) -> int:
    n = int(_lib.load().b200_w4a16_packed_bytes(K, N, group_size))
    if n < 0:
        raise ValueError(f"W4A16 needs K % 128 == 0 and N % 128 == 0, got K={K} N={N}")
    return n


def w4a16_prepack_awq(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor,
                      group_size: int) -> torch.Tensor:
    """AWQ checkpoint tensors -> opaque packed weight (replaces marlin::awq_repack + the
    zero-point / scale permutations of qlinear_awq_marlin_impl.cpp:62-125)."""
    _cuda(qweight, qzeros, scales)
    assert qweight.dtype == torch.int32 and qzeros.dtype == torch.int32
    assert scales.dtype == torch.bfloat16, "W4A16 path computes in bf16"
    assert qweight.is_contiguous() and qzeros.is_contiguous() and scales.is_contiguous()
    K, N = qweight.shape[0], qweight.shape[1] * 8
    packed = torch.empty(w4a16_packed_bytes(K, N, group_size), dtype=torch.uint8,
                         device=qweight.device)
    check(_lib.load().b200_w4a16_prepack_awq(_p(packed), _p(qweight), _p(qzeros), _p(scales), K, N,
                                             group_size, _stream()))
    return packed


def w4a16_prepack_gptq(qweight: torch.Tensor, qzeros: Optional[torch.Tensor], scales: torch.Tensor,
                       group_size: int, zeros_plus_one: bool = True) -> torch.Tensor:
    """GPTQ checkpoint tensors -> packed weight.  qzeros=None means the symmetric zero point 8
    used by the reference's Marlin path (qlinear_gptq_marlin_impl.cpp:18-20)."""
    _cuda(qweight, qzeros, scales)
    assert qweight.dtype == torch.int32 and scales.dtype == torch.bfloat16
    assert qweight.is_contiguous() and scales.is_contiguous()
    K, N = qweight.shape[0] * 8, qweight.shape[1]
    packed = torch.empty(w4a16_packed_bytes(K, N, group_size), dtype=torch.uint8,
                         device=qweight.device)
    check(_lib.load().b200_w4a16_prepack_gptq(_p(packed), _p(qweight), _p(qzeros), _p(scales), K,
                                              N, group_size, int(zeros_plus_one), _stream()))
    return packed


def w4a16_prepack_gptq_actorder(qweight: torch.Tensor, qzeros: Optional[torch.Tensor], scales: torch.Tensor,
                                g_idx: torch.Tensor, group_size: int, zeros_plus_one: bool = True):
    """GPTQ desc_act checkpoint -> (packed weight, perm): rows sorted by quant group
    (perm = argsort(g_idx), qlinear_gptq_marlin_impl.cpp:43-56); the caller gathers the activation
    columns with the same perm (permute_cols) before the GEMM.  Whole K only."""
    _cuda(qweight, qzeros, scales, g_idx)
    assert qweight.dtype == torch.int32 and scales.dtype == torch.bfloat16 and group_size > 0
    K, N = qweight.shape[0] * 8, qweight.shape[1]
    assert g_idx.numel() == K
    perm = torch.argsort(g_idx.to(torch.int64), stable=True).to(torch.int32).contiguous()
    g_sorted = g_idx.to(torch.int32)[perm.long()].contiguous()
    # whole K: every group has exactly group_size rows, so sorted rows fall into aligned blocks
    want = torch.arange(K, device=g_idx.device, dtype=torch.int32) // group_size
    if not torch.equal(g_sorted, want):
        raise ValueError("act-order: g_idx does not cover K with groups of exactly group_size rows "
                         "(a K-sharded act-order weight is not supported)")
    packed = torch.empty(w4a16_packed_bytes(K, N, group_size), dtype=torch.uint8, device=qweight.device)
    check(_lib.load().b200_w4a16_prepack_gptq_actorder(_p(packed), _p(qweight), _p(qzeros), _p(scales),
                                                       _p(perm), _p(g_sorted), K, N, group_size,
                                                       int(zeros_plus_one), _stream()))
    return packed, perm


def permute_cols(x: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """out[:, j] = x[:, perm[j]] (permute_cols_kernel, marlin/gptq_gemm.cu:66-104)"""
    _cuda(x, perm)
    assert x.dim() == 2 and x.stride(1) == 1 and perm.dtype == torch.int32 and perm.numel() == x.shape[1]
    out = torch.empty((x.shape[0], x.shape[1]), dtype=x.dtype, device=x.device)
    check(_lib.load().b200_permute_cols(_p(out), _p(x), _p(perm), x.shape[0], x.shape[1], x.stride(0),
                                        out.stride(0), _dt(x), _stream()))
    return out


def w4a16_repack_nibbles(qweight: torch.Tensor, method: str, perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The operator-level drop-in's repack (marlin::awq_repack / gptq_repack with the reference's
    signature): nibble tiles only, [K/16, N*16/8] int32."""
    _cuda(qweight, perm)
    if method == "awq":
        K, N = qweight.shape[0], qweight.shape[1] * 8
        out = torch.empty((K // 16, N * 2), dtype=torch.int32, device=qweight.device)
        check(_lib.load().b200_w4a16_repack_awq(_p(out), _p(qweight), K, N, _stream()))
    else:
        K, N = qweight.shape[0] * 8, qweight.shape[1]
        out = torch.empty((K // 16, N * 2), dtype=torch.int32, device=qweight.device)
        check(_lib.load().b200_w4a16_repack_gptq(_p(out), _p(qweight), _p(perm), K, N, _stream()))
    return out


def w4a16_assemble_marlin(nibbles: torch.Tensor, scales_marlin: torch.Tensor,
                          zeros_marlin: Optional[torch.Tensor], K: int, N: int, group_size: int) -> torch.Tensor:
    """nibble tiles + Marlin-order scales (+ Marlin-packed zero points; None = symmetric 8) -> tile blobs"""
    _cuda(nibbles, scales_marlin, zeros_marlin)
    packed = torch.empty(w4a16_packed_bytes(K, N, group_size), dtype=torch.uint8, device=nibbles.device)
    check(_lib.load().b200_w4a16_assemble_marlin(_p(packed), _p(nibbles), _p(scales_marlin), _p(zeros_marlin),
                                                 K, N, group_size, _stream()))
    return packed


def w4a16_dequant(packed: torch.Tensor, K: int, N: int, group_size: int) -> torch.Tensor:
    _cuda(packed)
    w = torch.empty((K, N), dtype=torch.bfloat16, device=packed.device)
    check(_lib.load().b200_w4a16_dequant(_p(w), _p(packed), K, N, group_size, _stream()))
    return w


def w4a16_workspace(device: torch.device, M: int, N: int, K: int) -> torch.Tensor:
    need = int(_lib.load().b200_w4a16_workspace_bytes(M, N, K))
    return _workspace("w4a16", device, need, zero=False)


def w4a16_gemm(a: torch.Tensor, packed: torch.Tensor, N: int, group_size: int,
               bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
               workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[M,N] = A[M,K] @ dequant(packed).  A bf16 with unit inner stride."""
    _cuda(a, packed, bias, out)
    assert a.dim() == 2 and a.dtype == torch.bfloat16 and a.stride(1) == 1
    M, K = a.shape
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    assert out.dtype == torch.bfloat16 and out.stride(1) == 1 and out.shape == (M, N)
    if workspace is None:
        workspace = w4a16_workspace(a.device, M, N, K)
    check(_lib.load().b200_w4a16_gemm(_p(out), _p(a), _p(packed), _p(bias), M, N, K, a.stride(0),
                                      out.stride(0), group_size, _p(workspace),
                                      workspace.numel(), _stream()))
    return out


def w4a16_splitk_splits(M: int, N: int, K: int) -> int:
    return int(_lib.load().b200_w4a16_splitk_splits(M, N, K))


class W4Partials(NamedTuple):
    """fp32 stream-K partials of a W4A16 GEMM: data [slots, M, N]; K is the GEMM's reduction
    dimension (the consumer recomputes the tile -> contributor-slot partition from (N, K))."""
    data: torch.Tensor
    K: int


def w4a16_gemm_splitk(a: torch.Tensor, packed: torch.Tensor, N: int, group_size: int,
                      poison: bool = False) -> W4Partials:
    """Partials mode: the GEMM writes fp32 partials [slots, M, N] and the consumer
    (rms_norm_residual_splitk, the TP all-reduce, w4a16_reduce_partials) performs the reduction.
    M <= 128."""
    _cuda(a, packed)
    assert a.dim() == 2 and a.dtype == torch.bfloat16 and a.stride(1) == 1
    M, K = a.shape
    slots = w4a16_splitk_splits(M, N, K)
    partials = torch.empty((slots, M, N), dtype=torch.float32, device=a.device)
    if poison:  # tests: slots a tile does not use must never be read
        partials.fill_(float("nan"))
    check(_lib.load().b200_w4a16_gemm_splitk(_p(partials), _p(a), _p(packed), M, N, K, a.stride(0),
                                             group_size, slots, _stream()))
    return W4Partials(partials, K)


def w4a16_reduce_partials(partials: W4Partials, out: Optional[torch.Tensor] = None,
                          bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M, N] (bf16) = bf16(sum over slots of partials) (+ bias)."""
    data = partials.data
    S, M, N = data.shape
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=data.device)
    _cuda(data, out, bias)
    assert out.dtype == torch.bfloat16 and out.stride(-1) == 1
    check(_lib.load().b200_w4a16_reduce_partials(_p(out), _p(data), S, partials.K, _p(bias), M, N,
                                                 out.stride(0), _stream()))
    return out


def rope_and_set_kv_cache_splitk(partials: W4Partials, n_heads: int, n_kv_heads: int, head_dim: int,
                                 positions: torch.Tensor, cos_sin: torch.Tensor,
                                 slot_ids: torch.Tensor, key_cache: torch.Tensor,
                                 value_cache: torch.Tensor, rotary_dim: int, interleaved: bool,
                                 dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """qkv GEMM partials -> qkv [T, (H + 2 Hkv) D] with q, k rotated, rotated k and v written to
    their cache slots: the GEMM's reduction + RoPE + set_kv_cache in one launch."""
    data = partials.data
    _cuda(data, positions, cos_sin, slot_ids, key_cache, value_cache)
    _check_i32(positions, slot_ids)
    S, T, n = data.shape
    assert n == (n_heads + 2 * n_kv_heads) * head_dim and data.is_contiguous()
    assert key_cache.is_contiguous() and value_cache.is_contiguous()
    qkv = torch.empty((T, n), dtype=dtype, device=data.device)
    check(_lib.load().b200_rope_kv_write_splitk(_p(qkv), _p(data), S, partials.K, _p(positions),
                                                _p(cos_sin), _p(slot_ids), _p(key_cache),
                                                _p(value_cache), T, n_heads, n_kv_heads, head_dim,
                                                rotary_dim, int(interleaved), _dt(qkv), _stream()))
    return qkv


def silu_mul_splitk(partials: W4Partials, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """gate_up GEMM partials [slots, rows, 2*inter] -> silu(gate) * up [rows, inter]."""
    data = partials.data
    _cuda(data)
    S, rows, n2 = data.shape
    assert n2 % 2 == 0 and data.is_contiguous()
    out = torch.empty((rows, n2 // 2), dtype=dtype, device=data.device)
    check(_lib.load().b200_silu_mul_splitk(_p(out), _p(data), S, partials.K, rows, n2 // 2,
                                           _dt(out), _stream()))
    return out


def rms_norm_residual_splitk(out: torch.Tensor, residual: torch.Tensor, partials: W4Partials,
                             weight: torch.Tensor, epsilon: float) -> None:
    """residual += T(sum over slots of partials); out = rms_norm(residual) * weight."""
    data = partials.data
    _cuda(out, residual, data, weight)
    assert data.dtype == torch.float32 and data.is_contiguous() and data.dim() == 3
    assert out.is_contiguous() and residual.is_contiguous()
    S, rows, n = data.shape
    if rows == 0:
        return
    check(_lib.load().b200_rms_norm_residual_splitk(_p(out), _p(residual), _p(data), S, partials.K,
                                                    _p(weight), rows, n, epsilon, _dt(out),
                                                    _stream()))


def argmax(logits: torch.Tensor) -> torch.Tensor:
    """Greedy sampling: torch.argmax(logits, dim=-1) (first maximal index) in one launch."""
    _cuda(logits)
    assert logits.dim() == 2 and logits.stride(1) == 1
    out = torch.empty(logits.shape[0], dtype=torch.int64, device=logits.device)
    check(_lib.load().b200_argmax(_p(out), _p(logits), logits.shape[0], logits.shape[1],
                                  logits.stride(0), _dt(logits), _stream()))
    return out


def launch_count() -> int:
    return int(_lib.load().b200_launch_count())


def launch_count_reset() -> None:
    _lib.load().b200_launch_count_reset()
