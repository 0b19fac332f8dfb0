"""GPU parity: AWQ / GPTQ prepack (bit exact through the dequant round trip) and the tcgen05
W4A16 GEMM vs the CPU oracle.  Bars: prepack bit exact; GEMM mean relative error < 1e-3 (the
reference's own Marlin bar, tests/kernels/marlin_gemm_test.py:104-107) and per-element within
bf16 rounding of the fp32 oracle result."""
import os

import numpy as np
import pytest
import torch

from oracle import quant
from scalellm_b200 import kernels
from tests.util import bf16_from_bits, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev_ckpt(ck):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in ck.items()}


@pytest.mark.parametrize("g", [128, 64, 32, -1])
@pytest.mark.parametrize("K,N", [(128, 128), (512, 256), (4096, 384)])
def test_awq_prepack_dequant_bit_exact(K, N, g):
    ck = quant.random_awq_checkpoint(K, N, g, seed=K + N + g)
    w_ref = quant.dequant(ck["q"], ck["z"], ck["scales"], g)
    d = dev_ckpt(ck)
    packed = kernels.w4a16_prepack_awq(d["qweight"], d["qzeros"], d["scales"], g)
    w = kernels.w4a16_dequant(packed, K, N, g)
    assert torch.equal(w.cpu().view(torch.int16), w_ref.view(torch.int16))


@pytest.mark.parametrize("g", [128, 32, -1])
@pytest.mark.parametrize("sym", [True, False])
def test_gptq_prepack_dequant_bit_exact(g, sym):
    K, N = 1024, 256
    ck = quant.random_gptq_checkpoint(K, N, g, seed=5 + g)
    d = dev_ckpt(ck)
    if sym:  # Marlin path: qzeros ignored, zero point 8 (qlinear_gptq_marlin_impl.cpp:18-20)
        packed = kernels.w4a16_prepack_gptq(d["qweight"], None, d["scales"], g)
        w_ref = quant.dequant(ck["q"], 8, ck["scales"], g)
    else:    # v1 checkpoint zeros (stored zero-1), including the zero+1 == 16 corner
        rng = np.random.default_rng(1)
        ng = 1 if g <= 0 else K // g
        zs = rng.integers(0, 16, size=(ng, N), dtype=np.int32)   # stored values 0..15 -> zeros 1..16
        qz = quant.pack_cols(zs).to(DEV)
        packed = kernels.w4a16_prepack_gptq(d["qweight"], qz, d["scales"], g, zeros_plus_one=True)
        w_ref = quant.dequant(ck["q"], zs + 1, ck["scales"], g)
    w = kernels.w4a16_dequant(packed, K, N, g)
    assert torch.equal(w.cpu().view(torch.int16), w_ref.view(torch.int16))


def test_prepack_of_reference_quant_utils_golden(golden_dir):
    """Checkpoints packed by the reference's quant_utils.pack_awq_weights / pack_gptq_weights."""
    gld = np.load(os.path.join(golden_dir, "quant_golden.npz"))
    for tag in "abc":
        K, N, g = (int(x) for x in gld[f"{tag}_shape"])
        s = bf16_from_bits(gld[f"{tag}_scales_bf16"]).to(DEV)
        w_ref = bf16_from_bits(gld[f"{tag}_wref_bf16"])
        ng = s.shape[0]
        z8 = quant.pack_awq(np.full((ng, N), 8, dtype=np.int32)).to(DEV)
        p_awq = kernels.w4a16_prepack_awq(torch.from_numpy(gld[f"{tag}_awq_packed"]).to(DEV), z8, s, g)
        p_gptq = kernels.w4a16_prepack_gptq(torch.from_numpy(gld[f"{tag}_gptq_packed"]).to(DEV),
                                            None, s, g)
        assert torch.equal(p_awq, p_gptq), "both checkpoint formats describe the same matrix"
        w = kernels.w4a16_dequant(p_awq, K, N, g)
        assert torch.equal(w.cpu().view(torch.int16), w_ref.view(torch.int16))


def gemm_case(M, K, N, g, seed, method="awq"):
    ck = (quant.random_awq_checkpoint if method == "awq" else quant.random_gptq_checkpoint)(
        K, N, g, seed)
    w_ref = quant.dequant(ck["q"], ck["z"], ck["scales"], g)
    gen = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=gen).bfloat16()
    d = dev_ckpt(ck)
    if method == "awq":
        packed = kernels.w4a16_prepack_awq(d["qweight"], d["qzeros"], d["scales"], g)
    else:
        packed = kernels.w4a16_prepack_gptq(d["qweight"], None, d["scales"], g)
    return a, w_ref, packed


def check_gemm(out, a, w_ref, what):
    ref32 = a.float() @ w_ref.float()
    ref = ref32.bfloat16()
    o = out.float().cpu()
    assert not torch.isnan(o).any(), what
    # marlin_gemm_test.py:104-107: mean relative error of the bf16 output vs the bf16 reference
    err = rel_err(o, ref.float())
    assert err < 1e-3, f"{what}: mean rel err {err:.3e}"
    # element-wise: |out - fp32 ref| <= one bf16 ulp of the value + fp32 accumulation noise
    tol = ref32.abs() * 2 ** -7 + 1e-3 * ref32.abs().mean()
    bad = ((o - ref32).abs() > tol).float().mean().item()
    assert bad == 0.0, f"{what}: {bad:.3e} of elements outside bf16 rounding of the oracle"
    frac_equal = (out.cpu().view(torch.int16) == ref.view(torch.int16)).float().mean().item()
    assert frac_equal > 0.9, f"{what}: only {frac_equal:.3f} bit-identical to the oracle"


@pytest.mark.parametrize("M", [1, 7, 16, 17, 32, 64, 65, 128])
def test_gemm_m_sweep(M):
    a, w_ref, packed = gemm_case(M, 1024, 512, 128, seed=M)
    out = kernels.w4a16_gemm(a.to(DEV), packed, 512, 128)
    torch.cuda.synchronize()
    check_gemm(out, a, w_ref, f"M={M}")


@pytest.mark.parametrize("g", [128, 64, 32, -1])
@pytest.mark.parametrize("method", ["awq", "gptq"])
def test_gemm_group_sizes(g, method):
    a, w_ref, packed = gemm_case(48, 2048, 384, g, seed=3 + g, method=method)
    out = kernels.w4a16_gemm(a.to(DEV), packed, 384, g)
    check_gemm(out, a, w_ref, f"g={g} {method}")


@pytest.mark.parametrize("K,N", [(4096, 6144), (4096, 4096), (4096, 28672), (14336, 4096),
                                 (128, 128), (256, 18944),
                                 # the same projections as TP=8 shards (column: N/8, row: K/8)
                                 (4096, 768), (512, 4096), (4096, 3584), (1792, 4096)])
def test_gemm_llama_shapes_m64(K, N):
    """The four Llama-3-8B decoder projections at the benchmark batch (+ edge shapes): every
    stream-K partition (full tiles, head/tail partials, multi-CTA reductions) is exercised."""
    a, w_ref, packed = gemm_case(64, K, N, 128, seed=K // 128 + N // 128)
    ws = kernels.w4a16_workspace(torch.device(DEV), 64, N, K)
    ws.view(torch.float32)[: ws.numel() // 4].fill_(float("nan"))  # no initialisation contract
    out = kernels.w4a16_gemm(a.to(DEV), packed, N, 128, workspace=ws)
    torch.cuda.synchronize()
    check_gemm(out, a, w_ref, f"K={K} N={N}")
    # deterministic: the fixed-order reduction gives bit-identical results run to run
    out2 = kernels.w4a16_gemm(a.to(DEV), packed, N, 128, workspace=ws)
    assert torch.equal(out, out2)


def test_gemm_bias_strided_and_large_m():
    a, w_ref, packed = gemm_case(200, 512, 256, 128, seed=9)      # M > 128: two passes
    bias = torch.randn(256).bfloat16()
    buf = torch.zeros(200, 1024, dtype=torch.bfloat16, device=DEV)
    buf[:, 256:768] = a.to(DEV)
    out = kernels.w4a16_gemm(buf[:, 256:768], packed, 256, 128, bias=bias.to(DEV))
    ref = quant.w4a16_gemm(a, w_ref, bias)
    assert rel_err(out, ref) < 1e-3
    assert (out.cpu().view(torch.int16) == ref.view(torch.int16)).float().mean() > 0.85


@pytest.mark.parametrize("method", ["awq", "gptq"])
def test_dense_prefill_path_of_the_int4_linear(method, monkeypatch):
    """Above 256 rows (prefill) the int4 linear dequantises once (bit-exact bf16 weights) and runs the
    library bf16 GEMM (default; B200_W4_PREFILL_DENSE=0: the streaming kernel in 128-row passes) — same
    bar as the fused kernel vs the oracle."""
    from scalellm_b200.layers import ColumnParallelQLinear, QuantArgs
    from scalellm_b200.model_parallel import ParallelArgs
    K, N, M, g = 1024, 768, 300, 128
    ck = (quant.random_awq_checkpoint if method == "awq" else quant.random_gptq_checkpoint)(K, N, g, seed=5)
    qa = QuantArgs(method, 4, g, is_sym=(method == "gptq"))
    lin = ColumnParallelQLinear(K, N, False, False, qa, ParallelArgs(0, 1, None), DEV)
    lin.load_state_dict({k: ck[k] for k in ("qweight", "qzeros", "scales") if ck.get(k) is not None})
    a = (torch.randn(M, K, generator=torch.Generator().manual_seed(6)) * 0.5).bfloat16()
    w_ref = quant.dequant(ck["q"], ck["z"], ck["scales"], g)
    ref = quant.w4a16_gemm(a, w_ref)
    monkeypatch.delenv("B200_W4_PREFILL_DENSE", raising=False)
    kernels.launch_count_reset()
    dense = lin(a.to(DEV))
    n_dense = kernels.launch_count()
    monkeypatch.setenv("B200_W4_PREFILL_DENSE", "0")
    fused = lin(a.to(DEV))
    assert n_dense >= 1                      # the dequant kernel of ours ran (then the library GEMM)
    assert rel_err(dense, ref) < 1e-3 and rel_err(fused, ref) < 1e-3
    assert (dense.cpu().view(torch.int16) == ref.view(torch.int16)).float().mean() > 0.85


def test_gemm_linearity_full_size():
    """Size-independent property at the benchmark shape: C(a1 + a2) == C(a1) + C(a2) when all
    terms are exactly representable (activations are small integers, weights q-z with s=2^-6)."""
    K, N, M, g = 4096, 4096, 64, 128
    rng = np.random.default_rng(0)
    q = rng.integers(0, 16, size=(K, N), dtype=np.int32)
    z = rng.integers(0, 16, size=(K // g, N), dtype=np.int32)
    s = torch.full((K // g, N), 2.0 ** -6, dtype=torch.bfloat16)
    packed = kernels.w4a16_prepack_awq(quant.pack_awq(q).to(DEV), quant.pack_awq(z).to(DEV),
                                       s.to(DEV), g)
    a1 = torch.from_numpy(rng.integers(-1, 2, size=(M, K)).astype(np.float32)).bfloat16().to(DEV)
    a2 = torch.from_numpy(rng.integers(-1, 2, size=(M, K)).astype(np.float32)).bfloat16().to(DEV)
    # every product and partial sum is an integer multiple of 2^-6 below 2^24: fp32 exact
    c1 = kernels.w4a16_gemm(a1, packed, N, g).float()
    c2 = kernels.w4a16_gemm(a2, packed, N, g).float()
    c12 = kernels.w4a16_gemm(a1 + a2, packed, N, g).float()
    w = torch.from_numpy((q - np.repeat(z, g, axis=0)).astype(np.float32)) * 2.0 ** -6
    exact = (a1.float().cpu() @ w)
    # results are exact up to the single final bf16 rounding
    assert torch.equal(c1.cpu(), exact.bfloat16().float())
    # each term carries one bf16 rounding (2^-9 relative) of its own magnitude
    tol = 2.0 ** -8 * (c1.abs() + c2.abs() + c12.abs())
    assert bool(((c12 - (c1 + c2)).abs() <= tol).all())


@pytest.mark.parametrize("K,N,M", [(4096, 4096, 64), (14336, 4096, 64), (1024, 512, 17), (512, 256, 128)])
def test_splitk_partials_fused_into_rms_norm_residual(K, N, M):
    """W4A16 split-K partial mode + b200_rms_norm_residual_splitk == GEMM -> residual add -> RMSNorm
    (models/meta/llama.h:174-176) with the oracle ops; the residual stream must come out within
    one bf16 rounding flip of the unfused oracle (summation order differs), the norm within 2 ulp."""
    from oracle import ops
    from tests.util import assert_ulp_or_abs
    a, w_ref, packed = gemm_case(M, K, N, 128, seed=K + N)
    gen = torch.Generator().manual_seed(1)
    res = torch.randn(M, N, generator=gen).bfloat16()
    wn = (1 + 0.1 * torch.randn(N, generator=gen)).bfloat16()
    # unused slots of the partials buffer must never be read: poison them
    partials = kernels.w4a16_gemm_splitk(a.to(DEV), packed, N, 128, poison=True)
    S = partials.data.shape[0]
    assert S == kernels.w4a16_splitk_splits(M, N, K) and 1 <= S <= 8 and partials.K == K
    # the partials sum to the GEMM result
    c = kernels.w4a16_reduce_partials(partials).cpu()
    ref32 = a.float() @ w_ref.float()
    assert torch.isfinite(c.float()).all()
    assert rel_err(c, ref32.bfloat16().float()) < 1e-3
    assert torch.equal(c, kernels.w4a16_gemm(a.to(DEV), packed, N, 128).cpu())
    d_res = res.to(DEV).clone()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    kernels.rms_norm_residual_splitk(out, d_res, partials, wn.to(DEV), 1e-5)
    ref_out, ref_res = ops.rms_norm_residual(ref32.bfloat16(), res, wn, 1e-5)
    # a one-ulp flip of the GEMM output (fp32 summation order) is many ulps of a residual sum that
    # cancels towards 0, so bound the ABSOLUTE error by one bf16 ulp of the largest operand
    assert_ulp_or_abs(d_res, ref_res, max_ulp=2, abs_frac=2 ** -7, what="residual")
    assert_ulp_or_abs(out, ref_out, max_ulp=3, abs_frac=2 ** -6, what="norm")
    # bit-identical to the unfused B200 path (same slot order, same rounding point)
    r2 = res.to(DEV).clone()
    out2 = torch.empty_like(out)
    kernels.rms_norm_residual(out2, r2, c.to(DEV), wn.to(DEV), 1e-5)
    assert torch.equal(d_res, r2) and torch.equal(out, out2)
    # deterministic
    p2 = kernels.w4a16_gemm_splitk(a.to(DEV), packed, N, 128)
    assert torch.equal(c, kernels.w4a16_reduce_partials(p2).cpu())


@pytest.mark.parametrize("K,inter,M", [(4096, 14336, 64), (512, 256, 7)])
def test_silu_mul_splitk_matches_unfused(K, inter, M):
    """gate_up partials -> silu(gate) * up in one launch == reduce -> silu_mul (bit-identical:
    same slot order, same rounding points; models/meta/llama.h:61-64)."""
    a, w_ref, packed = gemm_case(M, K, 2 * inter, 128, seed=K + inter)
    parts = kernels.w4a16_gemm_splitk(a.to(DEV), packed, 2 * inter, 128, poison=True)
    fused = kernels.silu_mul_splitk(parts)
    gu = kernels.w4a16_reduce_partials(parts)
    unfused = kernels.silu_with_mul(gu)
    assert torch.isfinite(fused.float()).all()
    assert torch.equal(fused, unfused)


@pytest.mark.parametrize("M,H,Hkv,D,bs", [(64, 32, 8, 128, 8), (5, 4, 2, 64, 16)])
def test_rope_kv_write_splitk_matches_unfused(M, H, Hkv, D, bs):
    """qkv partials -> (sum, RoPE, KV-slot write) in one launch == reduce -> rope_and_set_kv_cache:
    bit-identical q/k/v rows and caches (models/meta/llama.h:123-133)."""
    from scalellm_b200.layers import RotaryEmbedding
    K, n = 512, (H + 2 * Hkv) * D
    a, w_ref, packed = gemm_case(M, K, n, 128, seed=n + M)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    rope = RotaryEmbedding(D, 4096, inv_freq, False, torch.bfloat16, DEV)
    gen = torch.Generator().manual_seed(3)
    positions = torch.randint(0, 4096, (M,), generator=gen, dtype=torch.int32).to(DEV)
    n_slots = 40 * bs
    slots = torch.randperm(n_slots, generator=gen)[:M].to(torch.int32).to(DEV)
    caches = [torch.zeros(n_slots, Hkv, D, dtype=torch.bfloat16, device=DEV) for _ in range(4)]
    parts = kernels.w4a16_gemm_splitk(a.to(DEV), packed, n, 128, poison=True)
    qkv_f = kernels.rope_and_set_kv_cache_splitk(parts, H, Hkv, D, positions, rope.cos_sin_cache,
                                                 slots, caches[0], caches[1], rope.rotary_dim, False)
    qkv_u = kernels.w4a16_reduce_partials(parts)
    q = qkv_u[:, : H * D].view(M, H, D)
    k = qkv_u[:, H * D: (H + Hkv) * D].view(M, Hkv, D)
    v = qkv_u[:, (H + Hkv) * D:].view(M, Hkv, D)
    kernels.rope_and_set_kv_cache(q, k, v, positions, rope.cos_sin_cache, slots, caches[2],
                                  caches[3], rope.rotary_dim, False)
    assert torch.isfinite(qkv_f.float()).all()
    assert torch.equal(qkv_f, qkv_u)
    assert torch.equal(caches[0], caches[2]) and torch.equal(caches[1], caches[3])


# ---------------------------------------------------------------------------------------------
# GPTQ act-order (desc_act) and the operator-level drop-in layouts
# ---------------------------------------------------------------------------------------------
def _actorder_checkpoint(K, N, g, seed):
    """A GPTQ desc_act checkpoint as AutoGPTQ writes it: the rows of the original order carry the
    group of their position in a random permutation (every group has exactly g rows)."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, size=(K, N))
    order = rng.permutation(K)                      # quantisation order
    g_idx = np.empty(K, dtype=np.int32)
    g_idx[order] = np.arange(K) // g                # row order[i] was quantised in group i // g
    z = rng.integers(0, 15, size=(K // g, N))       # stored minus one (GPTQ v1): <= 14 keeps z + 1 <= 15
    s = (torch.randn(K // g, N, generator=torch.Generator().manual_seed(seed)).abs() * 0.01 + 1e-3).bfloat16()
    return dict(q=q, qweight=quant.pack_gptq(q), qzeros=quant.pack_cols(z), scales=s,
                g_idx=torch.from_numpy(g_idx))


@pytest.mark.parametrize("K,N,M", [(1024, 512, 48), (256, 256, 5)])
def test_gptq_act_order_matches_construct_weights(K, N, M):
    """desc_act through the plugin layer: packed rows sorted by group + activation columns gathered
    by the same perm == x @ construct_weights(qweight, qzeros, scales, g_idx) (qlinear_impl.cpp:21-56,
    the reference's own definition of the dequantised act-order weight)."""
    from scalellm_b200.layers import ColumnParallelQLinear, QuantArgs
    from scalellm_b200.model_parallel import ParallelArgs
    ck = _actorder_checkpoint(K, N, 128, seed=K + N)
    w_ref = quant.construct_gptq_weights(ck["qweight"], ck["qzeros"], ck["scales"], ck["g_idx"])   # bf16 [K, N]
    qa = QuantArgs(quant_method="gptq", bits=4, group_size=128, desc_act=True, is_sym=False)
    lin = ColumnParallelQLinear(K, N, False, False, qa, ParallelArgs(0, 1, None), torch.device(DEV))
    lin.load_state_dict({k: ck[k] for k in ("qweight", "qzeros", "scales", "g_idx")})
    a = torch.randn(M, K, generator=torch.Generator().manual_seed(3)).bfloat16()
    out = lin(a.to(DEV))
    assert lin.perm is not None and not lin.supports_partials(M)
    # dequantised weights are bit-identical to the reference definition, row for row after the perm
    w_ours = kernels.w4a16_dequant(lin.packed, K, N, 128).cpu()
    assert torch.equal(w_ours, w_ref[lin.perm.cpu().long()])
    want = quant.w4a16_gemm(a, w_ref)
    assert rel_err(out, want) < 1e-3
    # without g_idx the same tensors mean a different weight: the perm matters
    lin2 = ColumnParallelQLinear(K, N, False, False, QuantArgs("gptq", 4, 128, False, False),
                                 ParallelArgs(0, 1, None), torch.device(DEV))
    lin2.load_state_dict({k: ck[k] for k in ("qweight", "qzeros", "scales")})
    assert rel_err(lin2(a.to(DEV)), want) > 1e-2


def test_real_gptq_fixture_with_g_idx(golden_dir):
    """src/layers/quantization/data/gptq_small.safetensors (K = N = 256, g128, carries g_idx): the
    reference's own checkpoint tensors through the act-order path."""
    from scalellm_b200.layers import ColumnParallelQLinear, QuantArgs
    from scalellm_b200.model_parallel import ParallelArgs
    d = np.load(os.path.join(golden_dir, "gptq_small.npz"))
    qweight, qzeros = torch.from_numpy(d["qweight"]), torch.from_numpy(d["qzeros"])
    scales = torch.from_numpy(d["scales"]).view(torch.float16).to(torch.bfloat16)
    g_idx = torch.from_numpy(d["g_idx"])
    K, N = qweight.shape[0] * 8, qweight.shape[1]
    w_ref = quant.construct_gptq_weights(qweight, qzeros, scales, g_idx)
    qa = QuantArgs(quant_method="gptq", bits=4, group_size=128, desc_act=True, is_sym=False)
    lin = ColumnParallelQLinear(K, N, False, False, qa, ParallelArgs(0, 1, None), torch.device(DEV))
    lin.load_state_dict(dict(qweight=qweight, qzeros=qzeros, scales=scales, g_idx=g_idx))
    a = torch.randn(16, K, generator=torch.Generator().manual_seed(1)).bfloat16()
    out = lin(a.to(DEV))
    assert torch.equal(kernels.w4a16_dequant(lin.packed, K, N, 128).cpu(), w_ref[lin.perm.cpu().long()])
    assert rel_err(out, quant.w4a16_gemm(a, w_ref)) < 1e-3


@pytest.mark.parametrize("method,g", [("awq", 128), ("gptq", 128), ("awq", -1), ("gptq", 64)])
def test_nibble_repack_plus_marlin_order_scales_assemble_to_the_same_blobs(method, g):
    """The operator-level drop-in's two steps — repack(q_weight) -> nibble tiles, then
    assemble(nibbles, Marlin-order scales, Marlin-packed zero points) — give byte for byte the blobs
    the one-step prepack builds from the checkpoint tensors."""
    K, N = 512, 384 if g != -1 else 256
    ck = quant.random_awq_checkpoint(K, N, g, seed=5) if method == "awq" else quant.random_gptq_checkpoint(K, N, g, seed=5)
    sc = ck["scales"]
    nib = kernels.w4a16_repack_nibbles(ck["qweight"].to(DEV), method)
    assert nib.shape == (K // 16, N * 2) and nib.dtype == torch.int32
    if method == "awq":
        z = quant.unpack_awq(ck["qzeros"])
        want = kernels.w4a16_prepack_awq(ck["qweight"].to(DEV), ck["qzeros"].to(DEV), sc.to(DEV), g)
        mz = quant.marlin_zero_points(z).to(DEV)
    else:
        want = kernels.w4a16_prepack_gptq(ck["qweight"].to(DEV), None, sc.to(DEV), g)
        mz = None                                                      # symmetric: zero point 8
    got = kernels.w4a16_assemble_marlin(nib, quant.permute_marlin_scales(sc).to(DEV), mz, K, N, g)
    assert torch.equal(got, want)
