"""GPU: the dense bf16 linear (csrc/dense.cu, SURVEY 8a A8 — the reference's F::linear -> cuBLASLt,
src/layers/linear/parallel_linear.cpp:256-263,294-308) against a plain PyTorch fp32 reference of the
same op: bf16 inputs, fp32 accumulation, one rounding.  Tolerance: one bf16 ulp of the output or 2e-3
of the output scale (the summation order differs from any reference's), mean relative error < 1e-3 —
the reference's own criterion for its GEMMs (tests/kernels/marlin_gemm_test.py:104-107)."""
import pytest
import torch

from scalellm_b200 import kernels

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check(got, a, w, bias=None):
    want = a.float() @ w.float().t()
    if bias is not None:
        want = (want.bfloat16().float() + bias.float())
    want = want.bfloat16()
    d = (got.float() - want.float()).abs()
    scale = want.float().abs().max().item()
    assert bool((d <= 2.0 ** -7 * want.float().abs() + 2e-3 * scale).all()), (float(d.max()), scale)   # ulp(x) <= 2^-7 |x|
    assert d.mean().item() / want.float().abs().mean().item() < 1e-3
    return (got.view(torch.int16) == want.view(torch.int16)).float().mean().item()


@pytest.mark.parametrize("M", [1, 16, 17, 64, 100, 128])
@pytest.mark.parametrize("N,K", [(4096, 4096), (6144, 4096), (16032, 4096), (264, 256), (1024, 14336), (8, 64)])
def test_dense_gemm_matches_fp32_reference(M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g, device=DEV).bfloat16()
    w = (torch.randn(N, K, generator=g, device=DEV) * 0.05).bfloat16()
    same = _check(kernels.dense_gemm(a, w), a, w)
    assert same > 0.9
    # deterministic run to run (fixed stream-K partition, fixed summation order)
    assert torch.equal(kernels.dense_gemm(a, w), kernels.dense_gemm(a, w))


def test_dense_gemm_lm_head_shape_bias_strides_and_large_batches():
    g = torch.Generator(device=DEV).manual_seed(3)
    w = (torch.randn(128256, 4096, generator=g, device=DEV) * 0.02).bfloat16()
    a = torch.randn(64, 4096, generator=g, device=DEV).bfloat16()
    _check(kernels.dense_gemm(a, w), a, w)
    # bias, a view of a wider activation buffer, and an output written in place
    bias = torch.randn(4096, generator=g, device=DEV).bfloat16()
    w2 = w[:4096]
    buf = torch.randn(40, 3 * 4096, generator=g, device=DEV).bfloat16()
    view = buf[:, 4096: 2 * 4096]
    out = torch.empty(40, 4096, dtype=torch.bfloat16, device=DEV)
    _check(kernels.dense_gemm(view, w2, bias=bias, out=out), view, w2, bias)
    # more than 128 rows: one pass over W per 128 rows
    a3 = torch.randn(300, 4096, generator=g, device=DEV).bfloat16()
    _check(kernels.dense_gemm(a3, w2), a3, w2)
    # against the library GEMM the reference uses: same tolerance class
    lib = torch.nn.functional.linear(a, w2)
    d = (kernels.dense_gemm(a, w2).float() - lib.float()).abs()
    assert bool((d <= 2 * 2.0 ** -7 * lib.float().abs() + 2e-3 * lib.float().abs().max()).all())


def test_dense_layers_use_the_kernel_and_match_the_library(monkeypatch):
    from scalellm_b200.layers import ColumnParallelLinear, RowParallelLinear
    from scalellm_b200.model_parallel import ParallelArgs
    pa = ParallelArgs(0, 1, None)
    col = ColumnParallelLinear(4096, 1024, False, pa, torch.bfloat16, torch.device(DEV))
    row = RowParallelLinear(1024, 4096, True, pa, torch.bfloat16, torch.device(DEV))
    g = torch.Generator().manual_seed(5)
    col.load_state_dict({"weight": (torch.randn(1024, 4096, generator=g) * 0.03).bfloat16()})
    row.load_state_dict({"weight": (torch.randn(4096, 1024, generator=g) * 0.03).bfloat16()})
    x = torch.randn(64, 4096, generator=g).bfloat16().to(DEV)
    kernels.launch_count_reset()
    y = row(col(x))
    assert kernels.launch_count() >= 4            # two GEMMs + two reduction passes of ours
    monkeypatch.setenv("B200_DENSE_IMPL", "cublas")
    y_lib = row(col(x))
    d = (y.float() - y_lib.float()).abs()
    assert bool((d <= 4 * 2.0 ** -8 * y_lib.float().abs() + 4e-3 * y_lib.float().abs().max()).all())


@pytest.mark.parametrize("M", [1, 64, 128])
def test_dense_partials_feed_the_int4_gemms_consumers(M):
    """b200_dense_gemm_splitk writes the int4 GEMM's stream-K partials format: the plain reduction
    gives the bf16 result of b200_dense_gemm bit for bit, and the fused consumers (sum + residual +
    RMSNorm, sum + SiLU*mul) equal "reduce, then the unfused kernel" with the unused slots poisoned."""
    g = torch.Generator(device=DEV).manual_seed(11 + M)
    a = torch.randn(M, 4096, generator=g, device=DEV).bfloat16()
    w = (torch.randn(4096, 4096, generator=g, device=DEV) * 0.03).bfloat16()
    parts = kernels.dense_gemm_splitk(a, w, poison=True)
    full = kernels.dense_gemm(a, w)
    assert torch.equal(kernels.w4a16_reduce_partials(parts), full)
    # residual + RMSNorm consumer
    res = torch.randn(M, 4096, generator=g, device=DEV).bfloat16()
    wn = torch.randn(4096, generator=g, device=DEV).bfloat16()
    r1, r2 = res.clone(), res.clone()
    o1 = torch.empty_like(res)
    kernels.rms_norm_residual_splitk(o1, r1, parts, wn, 1e-5)
    o2 = torch.empty_like(res)
    kernels.rms_norm_residual(o2, r2, full, wn, 1e-5)
    assert torch.equal(o1, o2) and torch.equal(r1, r2)
    # SiLU * mul consumer over a gate_up-shaped weight
    w2 = (torch.randn(2048, 4096, generator=g, device=DEV) * 0.03).bfloat16()
    p2 = kernels.dense_gemm_splitk(a, w2, poison=True)
    f2 = kernels.dense_gemm(a, w2)
    got = kernels.silu_mul_splitk(p2, torch.bfloat16)
    want = kernels.silu_mul(f2[:, :1024], f2[:, 1024:])
    assert torch.equal(got, want)
