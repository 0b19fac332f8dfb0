"""oracle.ops — norm / rope / kv-cache / activation / paged attention on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  torch-CPU, fp32 math, the
reference's rounding points made explicit with ``_r`` (round to the tensor dtype
and come back to fp32, which is what c10::BFloat16 / c10::Half operators do:
torch/headeronly/util/BFloat16.h "operator*" etc. compute in float and round).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch


def _r(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """One rounding to `dtype`, result kept in fp32."""
    return x.to(dtype).to(torch.float32)


# ----------------------------------------------------------------------------
# RMSNorm — src/kernels/layernorm_kernels.cu:15-41 (kernel rounding) and
# src/layers/normalization.h:17-52 (torch formula)
# ----------------------------------------------------------------------------
def _sumsq_kernel_order(xf: torch.Tensor) -> torch.Tensor:
    """Row-wise sum of squares in the reference KERNEL's order, fp32 (layernorm_kernels.cu:30-35 +
    reduce_kernel_utils.cuh:15-64): a block of BD = min(n, 1024) threads; thread t accumulates
    x[t], x[t+BD], ... with one fused multiply-add each (`variance += x * x` compiles to FFMA),
    every warp of 32 consecutive threads reduces by an xor butterfly (16, 8, 4, 2, 1), and a last
    butterfly runs over the <= 32 warp sums.  Returns [rows, 1] float32."""
    import numpy as np
    x = xf.detach().to(torch.float32).reshape(-1, xf.shape[-1]).numpy()
    rows, n = x.shape
    BD = min(n, 1024)
    nvw = (BD + 31) // 32
    v = np.zeros((rows, nvw * 32), dtype=np.float32)
    for k in range((n + BD - 1) // BD):
        seg = x[:, k * BD: min(n, (k + 1) * BD)].astype(np.longdouble)
        w = seg.shape[1]
        # fma: exact product (48 bits fit in the 64-bit significand), one rounding to fp32
        v[:, :w] = (seg * seg + v[:, :w].astype(np.longdouble)).astype(np.float32)

    def butterfly(a):                      # a: [..., 32] float32 -> every lane holds the total
        lane = np.arange(32)
        for m in (16, 8, 4, 2, 1):
            a = (a + a[..., lane ^ m]).astype(np.float32)
        return a

    warp = butterfly(v.reshape(rows, nvw, 32))[..., 0]           # lane 0 parks the warp's sum
    red = np.zeros((rows, 32), dtype=np.float32)
    red[:, :nvw] = warp
    tot = butterfly(red)[:, 0]
    return torch.from_numpy(tot.astype(np.float32)).reshape(*xf.shape[:-1], 1)


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    xf = x.to(torch.float32)
    var = _sumsq_kernel_order(xf) / x.shape[-1]
    rstd = torch.rsqrt(var + eps)              # the GPU's MUFU.RSQ may differ by an ulp from this
    y = _r(xf * rstd, dt)                      # (T)(x * s_variance)   layernorm_kernels.cu:39
    return (y * weight.to(torch.float32)).to(dt)  # ... * weight[i] in T


def rms_norm_residual(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor,
                      eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (out, new_residual).  layernorm_kernels.cu:125-155: the variance
    uses the unrounded fp32 sum r+x, the output re-reads the ROUNDED residual."""
    dt = x.dtype
    s = residual.to(torch.float32) + x.to(torch.float32)
    new_res = s.to(dt)
    var = _sumsq_kernel_order(s) / x.shape[-1]
    rstd = torch.rsqrt(var + eps)
    y = _r(new_res.to(torch.float32) * rstd, dt)
    return (y * weight.to(torch.float32)).to(dt), new_res


# ----------------------------------------------------------------------------
# Rotary embedding — src/layers/pos_embedding.cpp:75-121,183-197 (cache),
# src/kernels/pos_embedding_kernels.cu:10-31 (per-op rounding)
# ----------------------------------------------------------------------------
def compute_default_inv_freq(rotary_dim: int, theta: float) -> torch.Tensor:
    sl = torch.arange(0, rotary_dim, 2, dtype=torch.float32)
    return 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float32), sl / rotary_dim)


def apply_llama3_rope_scaling(inv_freq: torch.Tensor, factor: float, low_freq_factor: float,
                              high_freq_factor: float, old_context_len: int) -> torch.Tensor:
    """pos_embedding.cpp:83-109 — float32 arithmetic element by element."""
    import numpy as np
    f32 = np.float32
    low_wl = f32(old_context_len) / f32(low_freq_factor)
    high_wl = f32(old_context_len) / f32(high_freq_factor)
    out = []
    for freq in inv_freq.to(torch.float32).numpy():
        freq = f32(freq)
        new = freq
        wavelen = f32(2 * math.pi / float(freq))
        if wavelen < high_wl:
            pass
        elif wavelen > low_wl:
            new = freq / f32(factor)
        else:
            smooth = (f32(old_context_len) / wavelen - f32(low_freq_factor)) / (
                f32(high_freq_factor) - f32(low_freq_factor))
            new = (f32(1) - smooth) * freq / f32(factor) + smooth * freq
        out.append(f32(new))
    return torch.tensor(np.array(out, dtype=np.float32))


def build_cos_sin_cache(rotary_dim: int, max_pos: int, inv_freq: torch.Tensor,
                        dtype: torch.dtype) -> torch.Tensor:
    """[max_pos, rotary_dim] = [cos | sin] in the MODEL dtype (pos_embedding.cpp:190-197)."""
    t = torch.arange(0, max_pos, dtype=torch.float32)
    freqs = torch.einsum("i,j->ij", t, inv_freq.to(torch.float32))
    return torch.cat([freqs.cos(), freqs.sin()], dim=-1).to(dtype)


def rope(q: torch.Tensor, k: torch.Tensor, positions: torch.Tensor, cos_sin: torch.Tensor,
         rotary_dim: int, interleaved: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """q:[T,H,D] k:[T,Hkv,D]; returns rotated copies.  x' = x*c - y*s, y' = x*s + y*c,
    every product and the add/sub rounded to T (pos_embedding_kernels.cu:24-29)."""
    dt = q.dtype
    half = rotary_dim // 2
    cs = cos_sin[positions.long()].to(torch.float32)  # [T, rotary_dim]
    c = cs[:, None, :half]
    s = cs[:, None, half:]

    def one(x: torch.Tensor) -> torch.Tensor:
        out = x.clone()
        xf = x.to(torch.float32)
        if interleaved:
            a, b = xf[..., 0:rotary_dim:2], xf[..., 1:rotary_dim:2]
        else:
            a, b = xf[..., :half], xf[..., half:rotary_dim]
        na = _r(a * c, dt) - _r(b * s, dt)
        nb = _r(a * s, dt) + _r(b * c, dt)
        if interleaved:
            out[..., 0:rotary_dim:2] = na.to(dt)
            out[..., 1:rotary_dim:2] = nb.to(dt)
        else:
            out[..., :half] = na.to(dt)
            out[..., half:rotary_dim] = nb.to(dt)
        return out

    return one(q), one(k)


# ----------------------------------------------------------------------------
# KV cache slot scatter / gather — src/memory/kv_cache.cpp:60-98,
# src/kernels/kv_cache_kernels.cu:9-41   (bit exact copies)
# ----------------------------------------------------------------------------
def kv_write(slot_ids: torch.Tensor, k: torch.Tensor, v: torch.Tensor, k_cache: torch.Tensor,
             v_cache: torch.Tensor) -> None:
    ids = slot_ids.long()
    k_cache[ids] = k
    v_cache[ids] = v


def kv_gather(slot_ids: torch.Tensor, k_cache: torch.Tensor,
              v_cache: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    ids = slot_ids.long()
    return k_cache[ids].clone(), v_cache[ids].clone()


# ----------------------------------------------------------------------------
# Activations — src/kernels/activation_kernels.cu:44-50,84-95,
# src/layers/activation.cpp:41-44,73-77
# ----------------------------------------------------------------------------
def silu(x: torch.Tensor) -> torch.Tensor:
    xf = x.to(torch.float32)
    return (xf / (1.0 + torch.exp(-xf))).to(x.dtype)


def silu_mul(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """T(silu(gate)) * up with the product rounded again (llama.h:61-64)."""
    dt = gate.dtype
    return (silu(gate).to(torch.float32) * up.to(torch.float32)).to(dt)


def silu_with_mul(x: torch.Tensor) -> torch.Tensor:
    n = x.shape[-1] // 2
    return silu_mul(x[..., :n], x[..., n:])


# ----------------------------------------------------------------------------
# Paged-KV variable-length attention — src/layers/attention/ref_handler.cpp:12-127,
# src/kernels/attention/tests/mha_ref.h:71-168, tests/kernels/attention/ref_attention.py
# ----------------------------------------------------------------------------
def slot_ids_for_sequence(block_table: torch.Tensor, block_cu_lens: torch.Tensor, b: int,
                          kv_len: int, block_size: int) -> torch.Tensor:
    """block_table holds FIRST-SLOT ids (block_id * block_size): sm80_kernel_mha.cuh:148-152."""
    idx = torch.arange(kv_len, dtype=torch.long)
    base = int(block_cu_lens[b])
    first = block_table[base + idx // block_size].long()
    return first + (idx % block_size)


def mha_ref(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, sm_scale: float,
            alibi_slopes: Optional[torch.Tensor], logits_soft_cap: float,
            sliding_window: int) -> torch.Tensor:
    """q:[q_len,H,D] k,v:[kv_len,Hkv,D] -> [q_len,H,D] (mha_ref.h:71-133)."""
    q_len, n_heads, _ = q.shape
    kv_len, n_kv_heads, _ = k.shape
    assert kv_len >= q_len
    if n_heads != n_kv_heads:
        g = n_heads // n_kv_heads
        k = k.repeat_interleave(g, dim=1)
        v = v.repeat_interleave(g, dim=1)
    scores = torch.einsum("qhd,khd->hqk", q.to(torch.float32), k.to(torch.float32)) * sm_scale
    if logits_soft_cap > 0.0:  # ref_handler.cpp:36
        scores = torch.tanh(scores / logits_soft_cap) * logits_soft_cap
    if alibi_slopes is not None:
        dist = torch.arange(kv_len, dtype=torch.float32)
        scores = scores + dist.view(1, 1, kv_len) * alibi_slopes.to(torch.float32).view(n_heads, 1, 1)
    mask = torch.ones(q_len, kv_len, dtype=torch.bool)
    if sliding_window >= 0:
        mask = torch.triu(mask, diagonal=kv_len - q_len - sliding_window)
    mask = torch.tril(mask, diagonal=kv_len - q_len)
    scores = scores.masked_fill(~mask, float("-inf"))
    p = torch.softmax(scores, dim=-1)
    return torch.einsum("hqk,khd->qhd", p, v.to(torch.float32)).to(q.dtype)


def paged_attention(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                    q_cu_lens: Sequence[int], kv_cu_lens: Sequence[int],
                    block_table: torch.Tensor, block_cu_lens: Sequence[int], block_size: int,
                    sm_scale: float, alibi_slopes: Optional[torch.Tensor] = None,
                    logits_soft_cap: float = 0.0, sliding_window: int = -1) -> torch.Tensor:
    """q:[T,H,D]; caches [n_slots,Hkv,D]; semantics of llm::paged_kv_varlen_mha
    (src/kernels/attention/attn_api.h:12-27)."""
    outs = []
    n_seqs = len(q_cu_lens) - 1
    bcl = torch.as_tensor(block_cu_lens)
    for b in range(n_seqs):
        qs, qe = int(q_cu_lens[b]), int(q_cu_lens[b + 1])
        kv_len = int(kv_cu_lens[b + 1]) - int(kv_cu_lens[b])
        if qe == qs:
            continue
        slots = slot_ids_for_sequence(block_table, bcl, b, kv_len, block_size)
        outs.append(mha_ref(q[qs:qe], k_cache[slots], v_cache[slots], sm_scale, alibi_slopes,
                            logits_soft_cap, sliding_window))
    return torch.cat(outs, dim=0) if outs else q.new_zeros(q.shape)


# ----------------------------------------------------------------------------
# Sampling tail: logits processors — src/kernels/sampling/penalty_kernels.cu:9-33,52-75,107-140,
# src/kernels/sampling/softmax_kernels.cu:11-54 (kernel semantics: every store rounds to T)
# ----------------------------------------------------------------------------
def apply_temperature_penalty(logits: torch.Tensor, temperatures: torch.Tensor) -> torch.Tensor:
    t = temperatures.to(torch.float32)
    inv = torch.where(t == 0, torch.ones_like(t), 1.0 / t)
    # `logits[i] *= inv` is c10's operator*=(T&, const T&): the float inverse is converted to T first
    inv = inv.to(logits.dtype).to(torch.float32)
    return (logits.to(torch.float32) * inv[:, None]).to(logits.dtype)


def apply_repetition_penalty(logits: torch.Tensor, token_ids: torch.Tensor, lens: torch.Tensor,
                             penalties: torch.Tensor) -> torch.Tensor:
    out = logits.clone()
    for b in range(logits.shape[0]):
        ids = token_ids[b, : int(lens[b])].long()
        x = out[b, ids].to(torch.float32)
        p = float(penalties[b].to(torch.float32))
        out[b, ids] = torch.where(x < 0, x * p, x / p).to(logits.dtype)
    return out


def apply_frequency_presence_penalty(logits: torch.Tensor, token_ids: torch.Tensor, counts: torch.Tensor,
                                     lens: torch.Tensor, freq: torch.Tensor, pres: torch.Tensor) -> torch.Tensor:
    out = logits.clone()
    for b in range(logits.shape[0]):
        n = int(lens[b])
        ids, c = token_ids[b, :n].long(), counts[b, :n]
        keep = c > 0
        ids, c = ids[keep], c[keep].to(torch.float64)
        x = out[b, ids].to(torch.float64)
        # `logit -= count * freq` compiles to one fused multiply-add (exact product, one rounding):
        # evaluated in double and rounded to fp32; then the presence term, a plain fp32 subtract
        x = (x - c * float(freq[b].to(torch.float32))).to(torch.float32)
        x = x - float(pres[b].to(torch.float32))
        out[b, ids] = x.to(logits.dtype)
    return out


def softmax_inplace_semantics(logits: torch.Tensor) -> torch.Tensor:
    """exp(x - max) is STORED in T and read back for the sum; sum + 1e-6 divides (softmax_kernels.cu:30-53).
    The sum's order is the kernel's (strided by min(vocab, 1024) threads, then two butterflies)."""
    import numpy as np
    dt = logits.dtype
    x = logits.to(torch.float32)
    e = torch.exp(x - x.max(dim=-1, keepdim=True).values).to(dt)
    ef = e.to(torch.float32).numpy()
    rows, n = ef.shape
    BD = min(n, 1024)
    nvw = (BD + 31) // 32
    v = np.zeros((rows, nvw * 32), dtype=np.float32)
    for k in range((n + BD - 1) // BD):
        seg = ef[:, k * BD: min(n, (k + 1) * BD)]
        v[:, : seg.shape[1]] = (v[:, : seg.shape[1]] + seg).astype(np.float32)
    lane = np.arange(32)

    def butterfly(a):
        for m in (16, 8, 4, 2, 1):
            a = (a + a[..., lane ^ m]).astype(np.float32)
        return a

    warp = butterfly(v.reshape(rows, nvw, 32))[..., 0]
    red = np.zeros((rows, 32), dtype=np.float32)
    red[:, :nvw] = warp
    denom = torch.from_numpy((butterfly(red)[:, 0] + np.float32(1e-6)).astype(np.float32))
    denom = denom.to(dt).to(torch.float32)     # `logits[i] /= sum` is operator/=(T&, const T&): the divisor is rounded to T
    return (e.to(torch.float32) / denom[:, None]).to(dt)


def top_k_top_p_filter(logits: torch.Tensor, top_k, top_p) -> torch.Tensor:
    """TopKTopPLogitsProcessor::forward (src/sampling/logits_processor.h:243-276), restated:
        sort each row descending; positions >= top_k -> -inf (top_k <= 0: no limit, :232-234);
        probs = softmax of what is left; positions with (cumsum(probs) - probs) > top_p -> -inf;
        scatter back to vocabulary order.
    Restated with the choices the reference leaves open made explicit: the sort is STABLE (of equal
    logits the lower vocabulary index comes first — torch.sort's order among ties is unspecified and the
    reference's test, logits_processor_test.cpp:263-357, only compares sorted values), and softmax /
    cumsum run in float64 where the reference's tensors are of the logits' dtype (bf16: every prob and
    every partial sum rounded to 8 bits — a token whose exclusive cumulative probability is within
    that rounding of top_p can fall on either side there).  top_p >= 1 means "no top-p limit": in the
    reference `(cumsum - probs) > 1.0` is false in exact arithmetic and true only where the rounding of its
    cumulative sum overshoots 1 (tail tokens of negligible mass, tests/test_sampling_oracle.py).
    logits [batch, vocab]; top_k int64 [batch] or None; top_p float [batch] or None.  Returns a new tensor
    of the logits' dtype."""
    x = logits.detach().cpu()
    out = x.clone()
    B, V = x.shape
    for b in range(B):
        v = x[b].double()
        order = torch.sort(v, descending=True, stable=True).indices
        keep = torch.ones(V, dtype=torch.bool)
        k = V
        if top_k is not None:
            kk = int(top_k[b])
            if 0 < kk < V:
                k = kk
        keep[k:] = False
        if top_p is not None and float(top_p[b]) < 1.0:   # top_p >= 1: no limit (see the docstring)
            p = float(top_p[b])
            vs = v[order].clone()
            vs[k:] = float("-inf")
            probs = torch.softmax(vs, dim=-1)
            excl = torch.cumsum(probs, dim=-1) - probs
            keep &= ~(excl > p)
        dropped = order[~keep]
        out[b, dropped] = float("-inf")
    return out
