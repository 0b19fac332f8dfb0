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
    that — and 2 more because the divisor itself is rounded to T: an exponential that rounds the other
    way on the host (its exp differs by CPU) can move the row sum across a rounding boundary of T, which
    shifts every output of the row by one ulp of the divisor (seen on one GPU box, not on the others).
    Against the reference's own kernel (same instruction) it is bit for bit."""
    logits = (torch.randn(B, V, generator=torch.Generator().manual_seed(V)) * 3).to(dtype)
    x = logits.to(DEV).clone()
    kernels.invoke_softmax(x)
    want = ops.softmax_inplace_semantics(logits)
    ulp = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11, torch.float32: 2.0 ** -22}[dtype]
    tiny = 2.0 ** -23 if dtype == torch.float16 else 1e-12      # fp16: probabilities down in the subnormals
    assert bool(((x.cpu().float() - want.float()).abs() <= (4 * ulp + 4e-5) * want.float().abs() + tiny).all())
    assert abs(float(x.float().sum(-1).mean()) - 1.0) < 2e-2
    ref = _ref()
    if ref is not None:
        r = logits.to(DEV).clone()
        ref.invoke_softmax(r)
        assert torch.equal(x, r)


def _kept(t):
    return torch.isfinite(t.float())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("vocab", [128256, 50257, 1000])
def test_top_k_top_p_filter_matches_the_restated_processor(dtype, vocab):
    """b200_topk_topp_filter against oracle.ops.top_k_top_p_filter (TopKTopPLogitsProcessor::forward,
    logits_processor.h:243-276, stable sort, float64 sums): the same tokens survive, the survivors keep
    their bits.  Rows cover: top-k only, top-p only, both, neither, k > vocab, p tiny (only the arg max
    survives), p >= 1, and rows of few distinct values (many exact ties at the cut)."""
    g = torch.Generator().manual_seed(vocab)
    B = 10
    x = (torch.randn(B, vocab, generator=g) * 3).to(dtype)
    x[6] = torch.randint(-2, 3, (vocab,), generator=g).to(dtype)          # 5 distinct values: ties everywhere
    x[7] = x[7].float().round().to(dtype)                                  # integer logits: ties at the cut
    top_k = torch.tensor([50, 0, 40, 0, vocab + 5, 1, 300, 77, 0, 5], dtype=torch.int64)
    top_p = torch.tensor([1.0, 0.9, 0.5, 1.0, 0.95, 0.3, 0.8, 0.6, 1e-6, 1.5], dtype=torch.float32)
    want = ops.top_k_top_p_filter(x, top_k, top_p)
    got = x.to(DEV).clone()
    kernels.apply_top_k_top_p(got, top_k.to(DEV), top_p.to(DEV))
    got = got.cpu()
    for b in range(B):
        kg, kw = _kept(got[b]), _kept(want[b])
        # the kept counts agree up to the summation arithmetic at the top-p boundary (fixed point vs float64:
        # identical unless a cumulative probability lands within ~1e-6 of p)
        assert abs(int(kg.sum()) - int(kw.sum())) <= (0 if top_p[b] >= 1 else 1), (b, int(kg.sum()), int(kw.sum()))
        if int(kg.sum()) == int(kw.sum()) and not torch.equal(kg, kw):
            # the only freedom left: +0 and -0 compare equal for the oracle's sort (index order) but are
            # different keys for the kernel (+0 above -0); the surviving VALUES must still be the same
            assert torch.equal(x[b][kg].float().sort().values, x[b][kw].float().sort().values), b
            assert bool(((x[b][kg ^ kw]).float() == 0).all()), b
        assert torch.equal(got[b][kg].view(torch.int16), x[b][kg].view(torch.int16))   # survivors untouched
        assert bool((got[b][~kg] == float("-inf")).all())
        # a prefix of the sorted order: every survivor >= every dropped logit
        if (~kg).any():
            assert float(x[b][kg].float().min()) >= float(x[b][~kg].float().max())
    # either argument may be absent; a strided view works; run to run identical
    only_k = x.to(DEV).clone()
    kernels.apply_top_k_top_p(only_k, top_k.to(DEV), None)
    assert torch.equal(_kept(only_k.cpu()), _kept(ops.top_k_top_p_filter(x, top_k, None)))
    buf = torch.zeros(B, vocab + 24, dtype=dtype, device=DEV)
    buf[:, :vocab] = x.to(DEV)
    kernels.apply_top_k_top_p(buf[:, :vocab], top_k.to(DEV), top_p.to(DEV))
    assert torch.equal(buf[:, :vocab].cpu().view(torch.int16), got.view(torch.int16))
    assert bool((buf[:, vocab:] == 0).all())


def test_top_k_top_p_filter_against_the_library_pipeline():
    """The reference's own sequence of library calls (sort, masked_fill, softmax, cumsum, gather) on the
    same bf16 logits on the GPU: same survivors except tokens whose exclusive cumulative probability is
    within bf16 rounding of top_p, and except the order among exact ties."""
    g = torch.Generator().manual_seed(3)
    B, V = 16, 32000
    x = (torch.randn(B, V, generator=g) * 4).bfloat16().to(DEV)
    top_k = torch.tensor([0, 100, 1000, 20] * 4, dtype=torch.int64, device=DEV)
    top_p = torch.tensor([0.9, 0.95, 0.5, 0.99] * 4, dtype=torch.float32, device=DEV)
    # logits_processor.h:243-276 verbatim in torch
    ls, li = x.sort(dim=-1, descending=True)
    kk = torch.where(top_k <= 0, torch.full_like(top_k, 2 ** 62), top_k).unsqueeze(1)
    ls = ls.masked_fill(torch.arange(V, device=DEV).expand_as(ls) >= kk, float("-inf"))
    ps = ls.softmax(dim=-1)
    ls = ls.masked_fill((ps.cumsum(dim=-1) - ps) > top_p.unsqueeze(1), float("-inf"))
    lib = ls.gather(-1, li.argsort(-1))
    got = x.clone()
    kernels.apply_top_k_top_p(got, top_k, top_p)
    n_lib, n_got = _kept(lib).sum(-1), _kept(got).sum(-1)
    # bf16 cumsum saturates near 1: allow a slack proportional to the kept count for the top-p rows
    assert bool(((n_lib - n_got).abs() <= 2 + n_got // 8).all()), (n_lib.tolist(), n_got.tolist())
    # both keep a prefix of the sorted order containing the arg max
    am = x.float().argmax(-1)
    assert bool(_kept(got)[torch.arange(B), am].all()) and bool(_kept(lib)[torch.arange(B), am].all())
