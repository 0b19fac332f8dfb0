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


# ---------------------------------------------------------------------------------------------
# The decision logic of csrc/sampling.cu's topk_topp_kernel (two-level radix histogram of the 16-bit
# key with fixed-point probability mass, the top-k boundary, Z, the top-p boundary, ties by index),
# restated in numpy statement for statement and run against the oracle on many random rows — the
# corner cases (all-equal rows, -inf entries, a single finite logit, k = 1, p -> 0, k > vocab) are cheap
# to reach here and expensive to reach on a GPU box.
# ---------------------------------------------------------------------------------------------
import numpy as np

FIX = 2.0 ** 40

def keys_of(x_bf16):
    u = x_bf16.view(torch.int16).numpy().astype(np.int64) & 0xFFFF
    key = np.where(u & 0x8000, (~u) & 0xFFFF, u | 0x8000)
    return u, key

def emulate_row(x, k_in, p):
    """mirror of topk_topp_kernel's decisions for one row (x: bf16 tensor [n])"""
    n = x.numel()
    k = int(k_in)
    if k <= 0 or k > n:
        k = n
    if k == n and not (p < 1.0):
        return x.clone()
    xf = x.float().numpy()
    xmax = np.nanmax(xf) if np.isfinite(xf).any() or True else -np.inf
    xmax = xf.max()
    if xmax == -np.inf:
        return x.clone()
    u, key = keys_of(x)
    with np.errstate(all='ignore'):
        e = np.exp((xf - np.float32(xmax)).astype(np.float32)).astype(np.float32)
    q = np.where(e >= 0, (e.astype(np.float64) * FIX).astype(np.uint64), 0).astype(np.uint64)
    hi, lo = key >> 8, key & 255
    cnt1 = np.bincount(hi, minlength=256)
    mass1 = np.zeros(256, dtype=np.uint64)
    np.add.at(mass1, hi, q)
    # pass A
    c = 0; m = 0; b = 255
    while b > 0:
        if c + cnt1[b] >= k: break
        c += int(cnt1[b]); m += int(mass1[b]); b -= 1
    bin_k, c_above, m_above = b, c, m
    def level2(bin_hi):
        sel = hi == bin_hi
        cnt2 = np.bincount(lo[sel], minlength=256)
        mass2 = np.zeros(256, dtype=np.uint64)
        np.add.at(mass2, lo[sel], q[sel])
        return cnt2, mass2
    cnt2, mass2 = level2(bin_k)
    c = c_above; m = m_above; l = 255
    while l > 0:
        if c + cnt2[l] >= k: break
        c += int(cnt2[l]); m += int(mass2[l]); l -= 1
    key_k = (bin_k << 8) | l
    need_k = k - c
    q_k = int(mass2[l]) // int(cnt2[l]) if cnt2[l] else 0
    z_fix = m + need_k * q_k
    key_t, ties_kept, ties_total = key_k, need_k, int(cnt2[l])
    if p < 1.0:
        t_fix = int(np.float64(np.float32(p)) * np.float64(z_fix)) if p > 0 else 0
        M = 0; C = 0; found = -1
        for bb in range(255, bin_k, -1):
            if M + int(mass1[bb]) > t_fix:
                found = bb; break
            M += int(mass1[bb]); C += int(cnt1[bb])
        if found >= 0:
            bin_p = found; cnt2, mass2 = level2(bin_p); Mabove = M
        else:
            bin_p = bin_k; Mabove = m_above
        M = Mabove
        in_k_bin = bin_p == bin_k
        for ll in range(255, -1, -1):
            kk = (bin_p << 8) | ll
            if in_k_bin and kk < key_k: break
            cc = int(cnt2[ll])
            if cc == 0: continue
            qq = int(mass2[ll]) // cc
            if in_k_bin and kk == key_k: cc = need_k
            if M + cc * qq > t_fix:
                kept = (t_fix - M) // qq + 1 if qq else cc
                kept = min(kept, cc)
                key_t, ties_kept, ties_total = kk, int(kept), int(cnt2[ll])
                break
            M += cc * qq
            if in_k_bin and kk == key_k: break
    keep = key > key_t
    tie_idx = np.nonzero(key == key_t)[0]
    keep[tie_idx[:ties_kept]] = True
    out = x.clone()
    out[torch.from_numpy(~keep)] = float('-inf')
    return out



def test_radix_cut_logic_of_the_filter_kernel_matches_the_processor():
    rng = np.random.default_rng(0)
    for it in range(320):
        n = int(rng.choice([7, 33, 256, 1000, 5000]))
        kind = it % 8
        g = torch.Generator().manual_seed(it)
        if kind == 0:
            x = torch.randn(n, generator=g) * 3
        elif kind == 1:
            x = torch.randint(-2, 3, (n,), generator=g).float()
        elif kind == 2:
            x = torch.full((n,), 1.5)
        elif kind == 3:
            x = torch.randn(n, generator=g) * 3
            x[torch.rand(n, generator=g) < 0.3] = float("-inf")
        elif kind == 4:
            x = (torch.randn(n, generator=g) * 3).round()
        elif kind == 5:
            x = torch.randn(n, generator=g) * 30
        elif kind == 6:
            x = torch.full((n,), float("-inf"))
            x[int(rng.integers(n))] = 2.0
        else:
            x = -torch.rand(n, generator=g) * 1e-3
        k = int(rng.choice([0, 1, 2, 5, n // 2, n, n + 3]))
        p = float(rng.choice([1.0, 0.9, 0.5, 0.1, 1e-6, 0.999, 2.0]))
        x = x.bfloat16()
        got = emulate_row(x, k, p)
        want = ops.top_k_top_p_filter(x[None], torch.tensor([k]), torch.tensor([p], dtype=torch.float32))[0]
        kg, kw = torch.isfinite(got.float()), torch.isfinite(want.float())
        fin = torch.isfinite(x.float())
        dg, dw = int((kg & fin).sum()), int((kw & fin).sum())
        assert abs(dg - dw) <= (0 if p >= 1 else 1), (it, kind, k, p, n, dg, dw)
        if dg == dw and not torch.equal(kg, kw):   # +0 / -0: equal for the sort, distinct keys for the radix
            assert torch.equal(x[kg].float().sort().values, x[kw].float().sort().values), (it, kind, k, p)
            assert bool((x[kg ^ kw].float() == 0).all()), (it, kind, k, p)
