"""oracle.quant — AWQ / GPTQ int4 checkpoint formats, dequantisation, W4A16 matmul.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy for the bit twiddling,
torch-CPU for the arithmetic.

Restates:
  * packing  : tests/kernels/quant_utils.py:101-197 (pack_rows / pack_cols / AWQ interleave)
  * GPTQ v1  : src/layers/quantization/qlinear_impl.cpp:21-99  (zeros + 1)
  * numerics : src/kernels/quantization/marlin/numeric_conversion.h:144-167 (int4 -> bf16 exact),
               :221-240 (sub_zp exact, scale = one bf16 multiply), marlin/mma.h:15-43 (fp32 accumulate)
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

AWQ_ORDER = np.array([0, 2, 4, 6, 1, 3, 5, 7])  # quant_utils.py:165-175


# ----------------------------------------------------------------------------
# checkpoint formats
# ----------------------------------------------------------------------------
def unpack_cols(packed: torch.Tensor) -> np.ndarray:
    """[K, N/8] int32 -> [K, N] (nibble i of a word = column 8*j+i).  quant_utils.py:140-160"""
    p = packed.cpu().numpy().astype(np.uint32)
    k, n8 = p.shape
    out = np.zeros((k, n8 * 8), dtype=np.int32)
    for i in range(8):
        out[:, i::8] = (p >> (4 * i)) & 0xF
    return out


def unpack_rows(packed: torch.Tensor) -> np.ndarray:
    """[K/8, N] int32 -> [K, N] (nibble i of a word = row 8*j+i).  quant_utils.py:118-137"""
    p = packed.cpu().numpy().astype(np.uint32)
    k8, n = p.shape
    out = np.zeros((k8 * 8, n), dtype=np.int32)
    for i in range(8):
        out[i::8, :] = (p >> (4 * i)) & 0xF
    return out


def pack_cols(q: np.ndarray) -> torch.Tensor:
    q = q.astype(np.uint32)
    k, n = q.shape
    out = np.zeros((k, n // 8), dtype=np.uint32)
    for i in range(8):
        out |= q[:, i::8] << np.uint32(4 * i)
    return torch.from_numpy(out.astype(np.int32))


def pack_rows(q: np.ndarray) -> torch.Tensor:
    q = q.astype(np.uint32)
    k, n = q.shape
    out = np.zeros((k // 8, n), dtype=np.uint32)
    for i in range(8):
        out |= q[i::8, :] << np.uint32(4 * i)
    return torch.from_numpy(out.astype(np.int32))


def pack_awq(q: np.ndarray) -> torch.Tensor:
    """AWQ: columns of every group of 8 reordered [0,2,4,6,1,3,5,7] then packed along N
    (quant_utils.py:181-184)."""
    k, n = q.shape
    qi = q.reshape(-1, 8)[:, AWQ_ORDER].reshape(k, n)
    return pack_cols(qi)


def unpack_awq(packed: torch.Tensor) -> np.ndarray:
    u = unpack_cols(packed)
    k, n = u.shape
    inv = np.argsort(AWQ_ORDER)
    return u.reshape(-1, 8)[:, inv].reshape(k, n)


def pack_gptq(q: np.ndarray) -> torch.Tensor:
    return pack_rows(q)  # quant_utils.py:177-178


def unpack_gptq(packed: torch.Tensor) -> np.ndarray:
    return unpack_rows(packed)


def unpack_gptq_zeros(qzeros: torch.Tensor, plus_one: bool) -> np.ndarray:
    """[K/g, N/8] int32, natural order along N; GPTQ-v1 stores zero-1 (qlinear_impl.cpp:44)."""
    z = unpack_cols(qzeros)
    return z + 1 if plus_one else z


# ----------------------------------------------------------------------------
# dequantisation and matmul
# ----------------------------------------------------------------------------
def dequant(q: np.ndarray, z: np.ndarray, scales: torch.Tensor, group_size: int) -> torch.Tensor:
    """W[k,n] = bf16_mul(bf16(q) - bf16(z), s[k/g, n]).  q:[K,N] ints, z:[K/g,N] ints (or a
    scalar), scales:[K/g,N] bf16.  The subtraction is exact in bf16 (|q-z| <= 16) and the
    product of an 8-bit-mantissa scale with a small integer is exact in fp32, so a single
    rounding fp32->bf16 reproduces __hmul2 (numeric_conversion.h:221-229)."""
    K, N = q.shape
    g = K if group_size <= 0 else group_size
    gi = np.arange(K) // g
    zf = np.broadcast_to(np.asarray(z), (K // g, N))[gi] if np.ndim(z) else np.full((K, N), z)
    d = torch.from_numpy((q - zf).astype(np.float32))
    s = scales.to(torch.float32)[torch.from_numpy(gi)]
    return (d * s).to(scales.dtype)


def w4a16_gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = A @ W with bf16 inputs, fp32 accumulation, one rounding at the end
    (marlin/mma.h:15-43, gemm_kernel.cuh use_fp32_reduce).  Bias is added as a second
    rounded op like the layer does (qlinear_awq_marlin_impl.cpp:360-363)."""
    c = (a.to(torch.float32) @ w.to(torch.float32)).to(a.dtype)
    if bias is not None:
        c = (c.to(torch.float32) + bias.to(torch.float32)).to(a.dtype)
    return c


def construct_gptq_weights(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor,
                           g_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """detail::construct_weights (qlinear_impl.cpp:21-56): scales[g_idx] * (w - (zeros+1)[g_idx]),
    evaluated in the scales' dtype like the reference (HalfTensor arithmetic)."""
    w = torch.from_numpy(unpack_gptq(qweight))
    z = torch.from_numpy(unpack_gptq_zeros(qzeros, plus_one=True))
    K = w.shape[0]
    if g_idx is None:
        g = K // scales.shape[0]
        g_idx = torch.arange(K) // g
    g_idx = g_idx.long()
    return scales[g_idx] * (w - z[g_idx]).to(scales.dtype)


# ----------------------------------------------------------------------------
# synthetic checkpoints for tests / bench
# ----------------------------------------------------------------------------
def random_awq_checkpoint(K: int, N: int, group_size: int, seed: int):
    """q,z ~ U{0..15}, s = |randn| * 0.01 (bf16).  Returns dict of checkpoint tensors + q,z."""
    g = K if group_size <= 0 else group_size
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, size=(K, N), dtype=np.int32)
    z = rng.integers(0, 16, size=(K // g, N), dtype=np.int32)
    gen = torch.Generator().manual_seed(seed)
    s = (torch.randn(K // g, N, generator=gen).abs() * 0.01 + 1e-4).to(torch.bfloat16)
    return {"qweight": pack_awq(q), "qzeros": pack_awq(z), "scales": s, "q": q, "z": z}


def random_gptq_checkpoint(K: int, N: int, group_size: int, seed: int):
    """symmetric (zero point 8) GPTQ checkpoint, desc_act = false."""
    g = K if group_size <= 0 else group_size
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, size=(K, N), dtype=np.int32)
    gen = torch.Generator().manual_seed(seed)
    s = (torch.randn(K // g, N, generator=gen).abs() * 0.01 + 1e-4).to(torch.bfloat16)
    z = np.full((K // g, N), 8, dtype=np.int32)
    # stored zeros follow the v1 convention (zero - 1 = 7)
    return {"qweight": pack_gptq(q), "qzeros": pack_cols(z - 1), "scales": s, "q": q, "z": z}


# ----------------------------------------------------------------------------
# Marlin layouts — what the reference's own GEMM kernel (marlin::gptq_gemm) consumes.  Restated
# from tests/kernels/quant_utils.py:178-291 (weights, scales) and
# src/layers/quantization/qlinear_awq_marlin_impl.cpp:62-97 (zero points); pinned against
# tests/golden/marlin_golden.npz (written by the reference's quant_utils).  Used by the GPU tests
# that run the reference kernel beside ours, and (the inverse maps) to check the drop-in shim that
# accepts Marlin-layout scales / zero points.  4-bit only.
# ----------------------------------------------------------------------------
_INTERLEAVE4 = np.array([0, 2, 4, 6, 1, 3, 5, 7])


def marlin_weight_perm() -> np.ndarray:
    """quant_utils.py:201-228: element order inside a [16 k x 64 n] slab (1024 entries), then the
    pairs-interleave of fast_conversion_interleave (:178-187)."""
    perm = []
    for i in range(32):
        perm1 = []
        col = i // 4
        for block in (0, 1):
            for row in (2 * (i % 4), 2 * (i % 4) + 1, 2 * (i % 4 + 4), 2 * (i % 4 + 4) + 1):
                perm1.append(16 * row + col + 8 * block)
        for j in range(4):
            perm.extend(p + 256 * j for p in perm1)
    perm = np.array(perm)
    return perm.reshape(-1, 8)[:, _INTERLEAVE4].ravel()


def marlin_scales_perm():
    """quant_utils.py:231-241"""
    scale_perm = [i + 8 * j for i in range(8) for j in range(8)]
    scale_perm_single = [2 * i + j for i in range(4) for j in (0, 1, 8, 9, 16, 17, 24, 25)]
    return np.array(scale_perm), np.array(scale_perm_single)


def pack_marlin_weights(q: np.ndarray) -> torch.Tensor:
    """[K, N] ints in 0..15 -> int32 [K/16, N*16/8] (quant_utils.py:245-278)."""
    k, n = q.shape
    assert k % 16 == 0 and n % 64 == 0
    t = q.reshape(k // 16, 16, n // 16, 16).transpose(0, 2, 1, 3).reshape(k // 16, n * 16)
    perm = marlin_weight_perm()
    res = t.reshape(-1, perm.size)[:, perm].reshape(t.shape).astype(np.uint32)
    packed = np.zeros((res.shape[0], res.shape[1] // 8), dtype=np.uint32)
    for i in range(8):
        packed |= res[:, i::8] << np.uint32(4 * i)
    return torch.from_numpy(packed.view(np.int32).copy())


def unpack_marlin_weights(packed: torch.Tensor, K: int, N: int) -> np.ndarray:
    """inverse of pack_marlin_weights -> [K, N] ints"""
    p = packed.cpu().numpy().view(np.uint32)
    res = np.zeros((p.shape[0], p.shape[1] * 8), dtype=np.int64)
    for i in range(8):
        res[:, i::8] = (p >> np.uint32(4 * i)) & 0xF
    perm = marlin_weight_perm()
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    t = res.reshape(-1, perm.size)[:, inv].reshape(K // 16, N * 16)
    return t.reshape(K // 16, N // 16, 16, 16).transpose(0, 2, 1, 3).reshape(K, N)


def permute_marlin_scales(s: torch.Tensor) -> torch.Tensor:
    """[G, N] -> Marlin column order (quant_utils.py:282-291; qlinear_awq_marlin_impl.cpp:116-124)."""
    g, n = s.shape
    perm, single = marlin_scales_perm()
    p = single if g == 1 else perm
    return s.reshape(-1, len(p))[:, torch.from_numpy(p)].reshape(-1, n).contiguous()


def unpermute_marlin_scales(s: torch.Tensor) -> torch.Tensor:
    g, n = s.shape
    perm, single = marlin_scales_perm()
    p = single if g == 1 else perm
    inv = np.empty_like(p)
    inv[p] = np.arange(len(p))
    return s.reshape(-1, len(p))[:, torch.from_numpy(inv)].reshape(-1, n).contiguous()


def marlin_zero_points(z: np.ndarray) -> torch.Tensor:
    """Natural-order zero points [G, N] (ints) -> the packed int32 [G, N/8] tensor the Marlin kernel
    reads with has_zp=true: columns permuted like the scales, then the 4-bit interleave, then packed
    along N (qlinear_awq_marlin_impl.cpp:84-96, after its AWQ un-interleave :64-76)."""
    g, n = z.shape
    perm, _ = marlin_scales_perm()
    m = z.reshape(-1, len(perm))[:, perm]
    m = m.reshape(-1, 8)[:, _INTERLEAVE4].reshape(g, n)
    return pack_cols(m)


def unpack_marlin_zero_points(packed: torch.Tensor) -> np.ndarray:
    """inverse of marlin_zero_points -> [G, N] ints"""
    m = unpack_cols(packed)
    g, n = m.shape
    inv8 = np.empty(8, dtype=np.int64)
    inv8[_INTERLEAVE4] = np.arange(8)
    m = m.reshape(-1, 8)[:, inv8]
    perm, _ = marlin_scales_perm()
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    return m.reshape(-1, len(perm))[:, inv].reshape(g, n)
