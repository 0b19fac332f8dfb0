"""GPU parity: RMSNorm / RoPE / KV write / SiLU through the C ABI vs the CPU oracle.
Tolerances: copies bit exact; RoPE exact (per-op rounding restated); RMSNorm / SiLU <= 1 ulp on a
tiny fraction of elements (fp32 sum order and rsqrtf/__expf approximations)."""
import pytest
import torch

from oracle import ops
from scalellm_b200 import kernels
from tests.util import assert_ulp

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,n", [(64, 4096), (100, 1038), (3, 8192), (7, 128), (2, 20560),
                                    (1, 40000), (5, 1024)])
def test_rms_norm(dtype, rows, n):
    torch.manual_seed(rows * 7 + n)
    x = torch.randn(rows, n).to(dtype)
    w = torch.randn(n).to(dtype)
    ref = ops.rms_norm(x, w, 1e-5)
    out = torch.empty_like(x, device=DEV)
    kernels.rms_norm(out, x.to(DEV), w.to(DEV), 1e-5)
    # two roundings ((T)(x*rstd), then *w): a 1-ulp flip of the first can show as 2 ulp
    assert_ulp(out, ref, max_ulp=2, max_frac=2e-3, what="rms_norm")
    # reference's own bar (normalization_test.cpp:129-132)
    assert torch.allclose(out.float().cpu(), ref.float(), rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,n", [(64, 4096), (9, 1038), (4, 16384)])
def test_rms_norm_residual(dtype, rows, n):
    torch.manual_seed(n)
    x = torch.randn(rows, n).to(dtype)
    r = torch.randn(rows, n).to(dtype)
    w = torch.randn(n).to(dtype)
    ref_out, ref_res = ops.rms_norm_residual(x, r, w, 1e-5)
    res = r.to(DEV).clone()
    out = torch.empty_like(x, device=DEV)
    kernels.rms_norm_residual(out, res, x.to(DEV), w.to(DEV), 1e-5)
    assert torch.equal(res.cpu(), ref_res), "residual stream must be bit exact"
    assert_ulp(out, ref_out, max_ulp=2, max_frac=2e-3, what="rms_norm_residual")


def test_rms_norm_fp32_and_empty():
    x = torch.randn(5, 512)
    w = torch.randn(512)
    out = torch.empty_like(x, device=DEV)
    kernels.rms_norm(out, x.to(DEV), w.to(DEV), 1e-6)
    assert torch.allclose(out.cpu(), ops.rms_norm(x, w, 1e-6), rtol=1e-5, atol=1e-6)
    e = torch.empty(0, 512, device=DEV)
    kernels.rms_norm(torch.empty_like(e), e, w.to(DEV), 1e-6)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("interleaved", [False, True])
@pytest.mark.parametrize("H,Hkv,D,rot", [(32, 8, 128, 128), (6, 3, 64, 32), (4, 4, 80, 20), (8, 1, 256, 256)])
def test_rope_inplace_exact(dtype, interleaved, H, Hkv, D, rot):
    torch.manual_seed(H + D)
    T = 17
    qkv = torch.randn(T, (H + 2 * Hkv) * D).to(dtype)      # strided views like the fused qkv output
    q = qkv[:, : H * D].view(T, H, D)
    k = qkv[:, H * D:(H + Hkv) * D].view(T, Hkv, D)
    inv = ops.compute_default_inv_freq(rot, 10000.0)
    cs = ops.build_cos_sin_cache(rot, 512, inv, dtype)
    pos = torch.randint(0, 512, (T,), dtype=torch.int32)
    rq, rk = ops.rope(q, k, pos, cs, rot, interleaved)
    dq = qkv.to(DEV)
    gq = dq[:, : H * D].view(T, H, D)
    gk = dq[:, H * D:(H + Hkv) * D].view(T, Hkv, D)
    kernels.apply_rotary_pos_emb(gq, gk, pos.to(DEV), cs.to(DEV), rot, interleaved)
    assert torch.equal(gq.cpu(), rq) and torch.equal(gk.cpu(), rk)
    # V part of the fused buffer untouched
    assert torch.equal(dq[:, (H + Hkv) * D:].cpu(), qkv[:, (H + Hkv) * D:])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_kv_write_and_gather_bit_exact(dtype):
    """kv_cache_test.cpp:16,55: round trip through random slots is exact."""
    torch.manual_seed(0)
    n_slots, Hkv, D, T = 4096, 8, 128, 300
    k = torch.randn(T, Hkv, D).to(dtype)
    v = torch.randn(T, Hkv, D).to(dtype)
    slots = torch.randperm(n_slots)[:T].to(torch.int32)
    kc = torch.zeros(n_slots, Hkv, D, dtype=dtype, device=DEV)
    vc = torch.zeros_like(kc)
    kernels.set_kv_cache(slots.to(DEV), k.to(DEV), v.to(DEV), kc, vc)
    rkc = torch.zeros(n_slots, Hkv, D, dtype=dtype)
    rvc = torch.zeros_like(rkc)
    ops.kv_write(slots, k, v, rkc, rvc)
    assert torch.equal(kc.cpu(), rkc) and torch.equal(vc.cpu(), rvc)
    gk, gv = kernels.get_kv_cache(slots.to(DEV), kc, vc)
    assert torch.equal(gk.cpu(), k) and torch.equal(gv.cpu(), v)


def test_kv_write_strided_inputs_odd_shapes():
    T, Hkv, D = 5, 3, 40   # row bytes not a multiple of 16 -> scalar path
    buf = torch.randn(T, 3 * Hkv * D).bfloat16()
    k = buf[:, Hkv * D: 2 * Hkv * D].view(T, Hkv, D)
    v = buf[:, 2 * Hkv * D:].view(T, Hkv, D)
    slots = torch.tensor([9, 0, 3, 7, 4], dtype=torch.int32)
    kc = torch.zeros(10, Hkv, D, dtype=torch.bfloat16, device=DEV)
    vc = torch.zeros_like(kc)
    d = buf.to(DEV)
    kernels.set_kv_cache(slots.to(DEV), d[:, Hkv * D: 2 * Hkv * D].view(T, Hkv, D),
                         d[:, 2 * Hkv * D:].view(T, Hkv, D), kc, vc)
    assert torch.equal(kc.cpu()[slots.long()], k) and torch.equal(vc.cpu()[slots.long()], v)


@pytest.mark.parametrize("H,Hkv,D,rot", [(32, 8, 128, 128), (4, 2, 64, 32)])
def test_fused_rope_kv_write_equals_two_step(H, Hkv, D, rot):
    torch.manual_seed(1)
    T, n_slots = 64, 1024
    qkv = torch.randn(T, (H + 2 * Hkv) * D).bfloat16().to(DEV)
    inv = ops.compute_default_inv_freq(rot, 500000.0)
    cs = ops.build_cos_sin_cache(rot, 4096, inv, torch.bfloat16).to(DEV)
    pos = torch.randint(0, 4096, (T,), dtype=torch.int32, device=DEV)
    slots = torch.randperm(n_slots)[:T].to(torch.int32).to(DEV)

    def views(t):
        return (t[:, : H * D].view(T, H, D), t[:, H * D:(H + Hkv) * D].view(T, Hkv, D),
                t[:, (H + Hkv) * D:].view(T, Hkv, D))

    a, b = qkv.clone(), qkv.clone()
    kc1 = torch.zeros(n_slots, Hkv, D, dtype=torch.bfloat16, device=DEV)
    vc1, kc2, vc2 = torch.zeros_like(kc1), torch.zeros_like(kc1), torch.zeros_like(kc1)
    q1, k1, v1 = views(a)
    kernels.apply_rotary_pos_emb(q1, k1, pos, cs, rot, False)
    kernels.set_kv_cache(slots, k1, v1, kc1, vc1)
    q2, k2, v2 = views(b)
    kernels.rope_and_set_kv_cache(q2, k2, v2, pos, cs, slots, kc2, vc2, rot, False)
    assert torch.equal(a, b) and torch.equal(kc1, kc2) and torch.equal(vc1, vc2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_silu_and_silu_mul(dtype):
    torch.manual_seed(0)
    rows, n = 64, 14336
    x = (torch.randn(rows, 2 * n) * 3).to(dtype)
    d = x.to(DEV)
    ref = ops.silu(x[:, :n])
    out = kernels.silu(d[:, :n])                       # strided view, like llama.h:63
    if dtype == torch.float32:
        assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-6)
    else:
        assert_ulp(out, ref, max_ulp=1, max_frac=1e-3, what="silu")
    ref2 = ops.silu_with_mul(x)
    out2 = kernels.silu_with_mul(d)
    out3 = kernels.silu_mul(d[:, :n], d[:, n:])
    assert torch.equal(out2, out3)
    if dtype == torch.float32:
        assert torch.allclose(out2.cpu(), ref2, rtol=1e-5, atol=1e-6)
    else:
        assert_ulp(out2, ref2, max_ulp=2, max_frac=1e-3, what="silu_mul")
        # the fused op equals kernel::silu followed by a torch multiply (two roundings)
        assert torch.equal(out2, out * d[:, n:])


def test_silu_odd_width_scalar_path():
    x = torch.randn(3, 2 * 37).bfloat16()
    out = kernels.silu_with_mul(x.to(DEV))
    assert_ulp(out, ops.silu_with_mul(x), max_ulp=1, max_frac=0.05)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("rows,n", [(64, 128256), (3, 1000), (1, 7), (5, 4099), (2, 16387), (7, 50264)])
def test_argmax_matches_torch(dtype, rows, n):
    """Greedy sampling tail: first index of the maximum (ties are common in bf16 logits), NaN wins."""
    from scalellm_b200 import kernels
    gen = torch.Generator().manual_seed(n)
    x = torch.randn(rows, n, generator=gen).to(dtype)
    x[0, n // 2] = x[0].max()            # a tie: the first occurrence must win
    x[0, n - 1] = x[0].max()
    if rows > 1:
        x[1, min(5, n - 1)] = float("nan")
    d = x.to(DEV)
    got = kernels.argmax(d)
    assert got.dtype == torch.int64
    assert torch.equal(got.cpu(), torch.argmax(x.float(), dim=-1))
    # a strided view (row stride > n), as the lm_head output of a padded buffer would be
    buf = torch.zeros(rows, n + 8, dtype=dtype, device=DEV)
    buf[:, :n] = d
    assert torch.equal(kernels.argmax(buf[:, :n]).cpu(), got.cpu())
