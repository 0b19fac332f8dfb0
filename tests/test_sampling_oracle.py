"""CPU: oracle.ops.top_k_top_p_filter against the reference's own sequence of library calls
(TopKTopPLogitsProcessor::forward, src/sampling/logits_processor.h:243-276, transcribed op for op in
torch on float32 logits, where rounding of the cumulative sums is not an issue and random logits have
no ties): the oracle is that algorithm."""
import torch

from oracle import ops


def _reference_ops(logits, top_k, top_p):
    V = logits.shape[-1]
    ls, li = logits.sort(dim=-1, descending=True)
    if top_k is not None:
        kk = torch.where(top_k <= 0, torch.full_like(top_k, 2 ** 62), top_k).unsqueeze(1)
        ls = ls.masked_fill(torch.arange(V).expand_as(ls) >= kk, float("-inf"))
    if top_p is not None:
        ps = ls.softmax(dim=-1)
        ls = ls.masked_fill((ps.cumsum(dim=-1) - ps) > top_p.unsqueeze(1), float("-inf"))
    return ls.gather(-1, li.argsort(-1))


def test_oracle_top_k_top_p_is_the_references_processor():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(12, 5000, generator=g) * 3
    top_k = torch.tensor([0, 5, 50, 500, 0, 1, 7000, 3, 0, 10, 20, 30], dtype=torch.int64)
    top_p = torch.tensor([0.9, 1.0, 0.5, 0.95, 0.1, 0.7, 0.99, 0.2, 1e-4, 0.8, 0.6, 2.0], dtype=torch.float32)
    for k, p in ((top_k, top_p), (top_k, None), (None, top_p)):
        want = _reference_ops(x, k, p)
        got = ops.top_k_top_p_filter(x, k, p)
        kg, kw = torch.isfinite(got), torch.isfinite(want)
        # top_p == 1.0 exactly: the reference's float32 cumulative sum can round to just above 1 and drop
        # tail tokens (here ~100 tokens carrying < 1e-6 of the mass); the oracle (and the kernel) read
        # top_p >= 1 as "no limit".  Everywhere else: the same tokens.
        exact = torch.ones(12, dtype=torch.bool) if p is None else (p != 1.0)
        assert torch.equal(kg[exact], kw[exact])
        if p is not None and bool((~exact).any()):
            rows = ~exact
            assert bool((kg[rows] | ~kw[rows]).all())                      # the oracle keeps a superset
            lost = torch.softmax(x[rows], -1)[kg[rows] & ~kw[rows]].sum()
            assert float(lost) < 1e-5
        assert torch.equal(got[kg], x[kg])
    # the arg max always survives, also at top_p -> 0
    assert bool(torch.isfinite(ops.top_k_top_p_filter(x, None, torch.zeros(12)))[torch.arange(12), x.argmax(-1)].all())
