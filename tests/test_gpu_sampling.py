"""GPU: the sampling tail's logits processors (csrc/sampling.cu) against the oracle's restatement of
src/kernels/sampling/{penalty,softmax}_kernels.cu and — where oracle/_ref is built — bit for bit
against the reference's own kernels."""
import os
import sys

import pytest
import torch

from oracle import ops
from scalellm_b200 import kernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "_ref_kernels.so")
pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref():
    if not os.path.exists(SO):
        return None
    sys.path.insert(0, os.path.dirname(SO))
    old = sys.getdlopenflags()
    sys.setdlopenflags(os.RTLD_LAZY | os.RTLD_LOCAL)
    try:
        import _ref_kernels
    finally:
        sys.setdlopenflags(old)
        sys.path.pop(0)
    return _ref_kernels if hasattr(_ref_kernels, "invoke_softmax") else None


def _case(dtype, B, V, L, seed):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(B, V, generator=g) * 4).to(dtype)
    ids = torch.stack([torch.randperm(V, generator=g)[:L] for _ in range(B)]).to(torch.int64)
    lens = torch.randint(0, L + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = L
    counts = torch.randint(0, 5, (B, L), generator=g, dtype=torch.int32)
    return logits, ids, lens, counts


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("B,V,L", [(64, 128256, 300), (3, 1000, 17), (1, 32000, 1)])
def test_logits_processors_match_the_oracle_and_the_reference_kernels(dtype, B, V, L):
    ref = _ref()
    logits, ids, lens, counts = _case(dtype, B, V, L, seed=B + L)
    g = torch.Generator().manual_seed(1)
    temps = (torch.rand(B, generator=g) * 1.5 + 0.1).to(dtype)
    temps[0] = 0.0                                                   # 0 means "leave the row alone"
    rep = (torch.rand(B, generator=g) + 1.0).to(dtype)
    freq = (torch.rand(B, generator=g) * 0.5).to(dtype)
    pres = (torch.rand(B, generator=g) * 0.5).to(dtype)
    d = lambda t: t.to(DEV)

    x = d(logits).clone()
    kernels.apply_temperature_penalty(x, d(temps))
    assert torch.equal(x.cpu(), ops.apply_temperature_penalty(logits, temps))
    if ref is not None:
        r = d(logits).clone()
        ref.apply_temperature_penalty(r, d(temps))
        assert torch.equal(x, r)

    y = x.clone()
    kernels.apply_repetition_penalty(y, d(ids), d(lens), d(rep))
    assert torch.equal(y.cpu(), ops.apply_repetition_penalty(x.cpu(), ids, lens, rep))
    if ref is not None:
        r = x.clone()
        ref.apply_repetition_penalty(r, d(ids), d(lens), d(rep))
        assert torch.equal(y, r)

    z = y.clone()
    kernels.apply_frequency_presence_penalty(z, d(ids), d(counts), d(lens), d(freq), d(pres))
    assert torch.equal(z.cpu(), ops.apply_frequency_presence_penalty(y.cpu(), ids, counts, lens, freq, pres))
    if ref is not None:
        r = y.clone()
        ref.apply_frequency_presence_penalty(r, d(ids), d(counts), d(lens), d(freq), d(pres))
        assert torch.equal(z, r)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("B,V", [(64, 128256), (5, 1000), (2, 50257)])
def test_softmax_in_place(dtype, B, V):
    """The GPU's __expf (ex2.approx of x * log2 e) differs from the host's exp by up to ~2e-5 relative
    at the arguments met here (|x - max| up to ~25), so the oracle comparison carries 2 ulp of T plus
    that; against the reference's own kernel (same instruction) it is bit for bit."""
    logits = (torch.randn(B, V, generator=torch.Generator().manual_seed(V)) * 3).to(dtype)
    x = logits.to(DEV).clone()
    kernels.invoke_softmax(x)
    want = ops.softmax_inplace_semantics(logits)
    ulp = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11, torch.float32: 2.0 ** -22}[dtype]
    tiny = 2.0 ** -23 if dtype == torch.float16 else 1e-12      # fp16: probabilities down in the subnormals
    assert bool(((x.cpu().float() - want.float()).abs() <= (2 * ulp + 4e-5) * want.float().abs() + tiny).all())
    assert abs(float(x.float().sum(-1).mean()) - 1.0) < 2e-2
    ref = _ref()
    if ref is not None:
        r = logits.to(DEV).clone()
        ref.invoke_softmax(r)
        assert torch.equal(x, r)
