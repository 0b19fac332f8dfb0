"""GPU: libb200decode against the reference's OWN CUDA kernels, compiled for sm_100a from
/root/reference by oracle/ref/Makefile into oracle/_ref/_ref_kernels.so (test infrastructure; the
.so is built in the dev container and travels to the GPU box).  This is the strongest pin the
parity claim can have: same inputs, the reference's kernel vs ours.

  rms_norm / rms_norm_residual   src/kernels/layernorm_kernels.cu:43-63,157-180      bit-exact
  apply_rotary_pos_emb           src/kernels/pos_embedding_kernels.cu:84-119         bit-exact
  set_kv_cache                   src/kernels/kv_cache_kernels.cu:43-78               bit-exact
  silu / silu_with_mul           src/kernels/activation_kernels.cu:121-154           bit-exact
  paged_kv_varlen_mha            src/kernels/attention/attn_api.cpp:14-73            both within the
                                 fp32-restatement tolerance, and within 2 bf16 ulp of each other

Skipped only when oracle/_ref/_ref_kernels.so is genuinely absent (build() makes it wherever
/root/reference exists; it travels to the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

from scalellm_b200 import kernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "_ref_kernels.so")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref not built (needs /root/reference)")]
DEV = "cuda"


@pytest.fixture(scope="module")
def ref():
    # only head_dim 128 of the reference attention is instantiated: resolve symbols lazily
    sys.path.insert(0, os.path.dirname(SO))
    old = sys.getdlopenflags()
    sys.setdlopenflags(os.RTLD_LAZY | os.RTLD_LOCAL)
    try:
        import _ref_kernels
    finally:
        sys.setdlopenflags(old)
        sys.path.pop(0)
    return _ref_kernels


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,n", [(64, 4096), (7, 1024), (1, 8192), (33, 320)])
def test_rms_norm_bit_exact_vs_reference_kernel(ref, dtype, rows, n):
    g = torch.Generator().manual_seed(rows * n)
    x = (torch.randn(rows, n, generator=g) * 3).to(dtype).to(DEV)
    w = (1 + 0.2 * torch.randn(n, generator=g)).to(dtype).to(DEV)
    res = torch.randn(rows, n, generator=g).to(dtype).to(DEV)
    o_ref, o_b = torch.empty_like(x), torch.empty_like(x)
    ref.rms_norm(o_ref, x, w, 1e-5)
    kernels.rms_norm(o_b, x, w, 1e-5)
    assert torch.equal(o_b, o_ref)
    r_ref, r_b = res.clone(), res.clone()
    ref.rms_norm_residual(o_ref, r_ref, x, w, 1e-5)
    kernels.rms_norm_residual(o_b, r_b, x, w, 1e-5)
    assert torch.equal(r_b, r_ref) and torch.equal(o_b, o_ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("interleaved", [False, True])
def test_rope_bit_exact_vs_reference_kernel(ref, dtype, interleaved):
    T, H, Hkv, D = 64, 32, 8, 128
    g = torch.Generator().manual_seed(3)
    q = torch.randn(T, H, D, generator=g).to(dtype).to(DEV)
    k = torch.randn(T, Hkv, D, generator=g).to(dtype).to(DEV)
    pos = torch.randint(0, 4096, (T,), generator=g, dtype=torch.int32).to(DEV)
    inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    f = torch.outer(torch.arange(4096, dtype=torch.float32), inv)
    cs = torch.cat([f.cos(), f.sin()], -1).to(dtype).to(DEV)
    q1, k1, q2, k2 = q.clone(), k.clone(), q.clone(), k.clone()
    ref.apply_rotary_pos_emb(q1, k1, pos, cs, D, interleaved)
    kernels.apply_rotary_pos_emb(q2, k2, pos, cs, D, interleaved)
    assert torch.equal(q2, q1) and torch.equal(k2, k1)


def test_set_kv_cache_bit_exact_vs_reference_kernel(ref):
    T, Hkv, D, n_slots = 64, 8, 128, 4096
    g = torch.Generator().manual_seed(5)
    k = torch.randn(T, Hkv, D, generator=g).bfloat16().to(DEV)
    v = torch.randn(T, Hkv, D, generator=g).bfloat16().to(DEV)
    slots = torch.randperm(n_slots, generator=g)[:T].to(torch.int32).to(DEV)
    c = [torch.zeros(n_slots, Hkv, D, dtype=torch.bfloat16, device=DEV) for _ in range(4)]
    ref.set_kv_cache(slots, k, v, c[0], c[1])
    kernels.set_kv_cache(slots, k, v, c[2], c[3])
    assert torch.equal(c[2], c[0]) and torch.equal(c[3], c[1])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_silu_bit_exact_vs_reference_kernel(ref, dtype):
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(64, 2 * 14336, generator=g) * 4).to(dtype).to(DEV)
    assert torch.equal(kernels.silu(x), ref.silu(x))
    assert torch.equal(kernels.silu_with_mul(x), ref.silu_with_mul(x))


@pytest.mark.parametrize("case", ["decode", "ragged", "alibi", "softcap", "window", "multi_token"])
def test_paged_attention_vs_reference_kernel(ref, case):
    from oracle import ops
    from tests.test_gpu_attention import make_case
    H, Hkv, D, bs = 32, 8, 128, 8
    kv = [2048, 77, 300, 9, 1025, 513, 64, 2000] if case != "decode" else [2048] * 8
    ql = [1, 3, 2, 1, 4, 1, 2, 1] if case == "multi_token" else [1] * 8
    c = make_case(ql, kv, H, Hkv, D, bs, torch.bfloat16, seed=len(case))
    slopes = (0.5 ** torch.arange(1, H + 1, dtype=torch.float32) * 0.1).to(DEV) if case == "alibi" else None
    cap = 30.0 if case == "softcap" else 0.0
    window = 128 if case == "window" else -1
    args = [c[k].to(DEV) for k in ("q", "kc", "vc", "q_cu", "kv_cu", "table", "blk_cu")]
    scale = D ** -0.5
    o_b = torch.empty_like(args[0])
    o_r = torch.empty_like(args[0])
    kernels.paged_kv_varlen_mha(o_b, *args, slopes, bs, c["max_q"], c["max_kv"], scale, cap, window)
    ref.paged_kv_varlen_mha(o_r, *args, slopes, bs, c["max_q"], c["max_kv"], scale, cap, window)
    torch.cuda.synchronize()
    want = ops.paged_attention(c["q"], c["kc"], c["vc"], c["q_cu"].tolist(), c["kv_cu"].tolist(),
                               c["table"], c["blk_cu"].tolist(), bs, scale, logits_soft_cap=cap,
                               sliding_window=window,
                               alibi_slopes=None if slopes is None else slopes.cpu())
    for name, got in (("b200", o_b), ("reference", o_r)):       # each against the fp32 restatement
        err = (got.float().cpu() - want.float()).abs().max().item()
        assert err <= 2e-2, (name, err)
    d = (o_b.float() - o_r.float()).abs()
    tol = 2 * 2.0 ** -8 * o_r.float().abs() + 2e-3 * o_r.float().abs().max()
    assert bool((d <= tol).all()), float(d.max())


def test_marlin_weight_arithmetic_known_answers(ref):
    """Every (q, z) pair x a spread of bf16 scales through the reference's own device functions
    (numeric_conversion.h dequant<bf16,4,has_zp> -> sub_zp -> scale, in gemm_kernel.cuh's order)
    == the oracle's restatement (oracle/quant.py dequant) == b200_w4a16_dequant.  Pins the AWQ
    zero-point path, for which the reference has no known-answer test of its own."""
    from oracle import quant
    g = torch.Generator().manual_seed(1)
    scales = torch.cat([torch.randn(61, generator=g).abs() * 0.02 + 1e-4,
                        torch.tensor([1.0, 0.5, 3.0]),
                        torch.tensor([2.0 ** -20, 65280.0])]).bfloat16()
    S = scales.numel()
    zp_tab, sym_tab = ref.marlin_dequant_table(scales.to(DEV))
    torch.cuda.synchronize()
    # oracle on the same grid: K = 16 values of q as rows, one group, N = S columns per z
    q = np.repeat(np.arange(16)[:, None], S, axis=1)
    for z in range(16):
        want = quant.dequant(q, z, scales[None, :], -1)            # [16, S]
        assert torch.equal(zp_tab[:, z, :].cpu(), want), z
    assert torch.equal(sym_tab.cpu(), quant.dequant(q, 8, scales[None, :], -1))
    # and our prepack + dequant kernel on a weight that contains every (q, z) pair
    K, N = 128, 256
    qw = np.zeros((K, N), dtype=np.int64)
    zz = np.zeros((1, N), dtype=np.int64)
    for n in range(N):
        zz[0, n] = n % 16
        qw[:, n] = (np.arange(K) + n // 16) % 16
    sc = scales[torch.arange(N) % S][None, :].contiguous()
    packed = kernels.w4a16_prepack_awq(quant.pack_awq(qw).to(DEV), quant.pack_awq(zz).to(DEV), sc.to(DEV), 128)
    ours = kernels.w4a16_dequant(packed, K, N, 128).cpu()
    for n in range(N):
        col = zp_tab[:, n % 16, n % S].cpu()                         # indexed by q
        assert torch.equal(ours[:, n], col[torch.from_numpy(qw[:, n])]), n


@pytest.mark.parametrize("M", [17, 33, 64])
def test_w4a16_gemm_vs_reference_marlin_kernel(ref, golden_dir, M):
    """The reference's own Marlin GEMM (marlin::gptq_gemm, src/kernels/quantization/marlin.h:17-28,
    fed exactly as tests/kernels/marlin_gemm_test.py:82-112 feeds it) and b200_w4a16_gemm on the same
    symmetric 4-bit / group-128 weights (tests/golden/marlin_golden.npz: both packings written by the
    reference's quant_utils) and the same activations.  Both accumulate bf16 x bf16 products in fp32
    and round once, in different orders: each must sit within the oracle's tolerance, and they must
    agree with each other to a bf16 ulp (or 2e-3 of the output scale where values cancel)."""
    from oracle import quant
    d = np.load(os.path.join(golden_dir, "marlin_golden.npz"))
    K, N, g = (int(x) for x in d["shape"])
    scales = torch.from_numpy(d["scales_bf16"]).view(torch.bfloat16)
    gen = torch.Generator().manual_seed(M)
    a = torch.randn(M, K, generator=gen).bfloat16()
    # ours: GPTQ checkpoint packing -> tile blobs (symmetric: zero point 8)
    packed = kernels.w4a16_prepack_gptq(torch.from_numpy(d["gptq_packed"]).to(DEV), None, scales.to(DEV), g)
    ours = kernels.w4a16_gemm(a.to(DEV), packed, N, g)
    # reference: Marlin packing + permuted scales, no zero points, no act-order
    empty_i = torch.empty(0, dtype=torch.int32, device=DEV)
    out_ref = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    workspace = torch.zeros(N // 64 * 16, dtype=torch.int32, device=DEV)
    ref.marlin_gemm(a.to(DEV), torch.from_numpy(d["marlin_packed"]).to(DEV), out_ref,
                    torch.from_numpy(d["marlin_scales_bf16"]).view(torch.bfloat16).to(DEV), empty_i,
                    empty_i, empty_i, workspace, 4, True, False, True)
    torch.cuda.synchronize()
    assert int(workspace.abs().sum()) == 0                       # Marlin returns its locks zeroed
    w = quant.dequant(d["q"].astype(np.int64), 8, scales, g)     # oracle weights
    want = quant.w4a16_gemm(a, w)
    scale = want.float().abs().max().item()
    for name, got in (("b200", ours), ("reference", out_ref)):
        err = (got.float().cpu() - want.float()).abs().max().item()
        assert err <= 2e-3 * scale + 2.0 ** -8 * scale, (name, err, scale)
    diff = (ours.float() - out_ref.float()).abs()
    tol = 2.0 ** -8 * out_ref.float().abs() + 2e-3 * scale
    assert bool((diff <= tol).all()), float(diff.max())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("rows,n", [(64, 4096), (7, 1024), (5, 2304), (3, 320)])
def test_gemma_rms_norm_and_layer_norm_bit_exact_vs_reference_kernel(ref, dtype, rows, n):
    """The rest of the reference's :kernels norm surface (layernorm_kernels.cu:66-123,185-260):
    same loop shape, same FFMA accumulation, the (1.0 + w) factor in double for Gemma."""
    g = torch.Generator().manual_seed(rows + n)
    x = (torch.randn(rows, n, generator=g) * 2 + 0.3).to(dtype).to(DEV)
    w = (0.2 * torch.randn(n, generator=g)).to(dtype).to(DEV)
    b = (0.1 * torch.randn(n, generator=g)).to(dtype).to(DEV)
    o_ref, o_b = torch.empty_like(x), torch.empty_like(x)
    ref.gemma_rms_norm(o_ref, x, w, 1e-6)
    kernels.gemma_rms_norm(o_b, x, w, 1e-6)
    assert torch.equal(o_b, o_ref)
    for bias in (b, None):
        ref.layer_norm(o_ref, x, w, bias, 1e-5)
        kernels.layer_norm(o_b, x, w, bias, 1e-5)
        assert torch.equal(o_b, o_ref), bias is None


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_gelu_bit_exact_vs_reference_kernel(ref, dtype):
    """gelu_new / gelu_fast (+ fused multiply), activation_kernels.cu:13-41,84-145"""
    g = torch.Generator().manual_seed(17)
    x = (torch.randn(33, 2 * 3072, generator=g) * 3).to(dtype).to(DEV)
    half = x[:, :3072].contiguous()
    assert torch.equal(kernels.gelu_new(half), ref.gelu_new(half))
    assert torch.equal(kernels.gelu_fast(half), ref.gelu_fast(half))
    assert torch.equal(kernels.gelu_new_with_mul(x), ref.gelu_new_with_mul(x))
    assert torch.equal(kernels.gelu_fast_with_mul(x), ref.gelu_fast_with_mul(x))
    view = x[:, 1000:1512]                                   # row stride != n (activation_kernel's `stride`)
    assert torch.equal(kernels.gelu_new(view), ref.gelu_new(view))
