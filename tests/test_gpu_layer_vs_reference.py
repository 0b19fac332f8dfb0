"""GPU: one Llama-3-8B-shaped decoder layer (h 4096, 32/8 heads of 128, inter 14336, AWQ int4 g128)
at the benchmark shape — batch 64, kv_len 2048, block_size 8, shuffled block ids — through OUR fused
decode path, against the same layer assembled from the reference's OWN kernels compiled for sm_100a
(oracle/_ref): rms_norm -> marlin::gptq_gemm(has_zp=true) -> apply_rotary_pos_emb -> set_kv_cache ->
paged_kv_varlen_mha -> gptq_gemm -> add -> rms_norm -> gptq_gemm -> silu * up -> gptq_gemm -> add
(models/meta/llama.h:61-64,123-133,170-177).

Bars (north star: "within 1e-3 rtol bf16, bit-exact KV indexing"):
  * every op fed the REFERENCE's own intermediate: bit-exact for RMSNorm, RoPE, KV write, SiLU*mul;
    the two accumulating ops (int4 GEMM, attention) within one bf16 ulp-or-2e-3 of the output scale
    of the reference kernel (different fp32 summation orders);
  * end to end (our layer on its own intermediates), the north star's "within 1e-3 rtol" read as
    an error relative to the output scale — mean |ours - ref| <= 1e-3 * max |ref| — because two bf16
    pipelines that round at ten points in different summation orders differ by about one bf16 ulp
    (2^-8 = 3.9e-3 relative) on most elements, so no element-wise 1e-3 can hold between ANY two
    correct bf16 implementations (measured here: mean |ours - ref| / mean |ref| = 3.7e-3, i.e. one
    ulp).  To show that this one ulp is rounding and not error, both are also compared with an
    fp32 evaluation of the same layer (same bf16 inputs and dequantised weights, no intermediate
    rounding): ours must be no farther from it than the reference's kernels are (x 1.25);
    max |ours - ref| <= 2 bf16 ulp + 4e-3 of the output scale; the KV slots the step wrote are
    exactly the reference's slots, V and K within the qkv GEMM's tolerance."""
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


def _ref():
    sys.path.insert(0, os.path.dirname(SO))
    old = sys.getdlopenflags()
    sys.setdlopenflags(os.RTLD_LAZY | os.RTLD_LOCAL)
    try:
        import _ref_kernels
    finally:
        sys.setdlopenflags(old)
        sys.path.pop(0)
    return _ref_kernels


def _awq_linear(K, N, g, seed):
    """Random AWQ weight in every layout: checkpoint tensors for us, Marlin tensors for the reference."""
    from oracle import quant
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, size=(K, N), dtype=np.int64).astype(np.uint8)
    z = rng.integers(0, 16, size=(K // g, N), dtype=np.int64).astype(np.uint8)
    s = (torch.randn(K // g, N, generator=torch.Generator().manual_seed(seed)).abs() * 0.01 + 1e-4).bfloat16()
    ours = dict(qweight=quant.pack_awq(q.astype(np.int64)), qzeros=quant.pack_awq(z.astype(np.int64)), scales=s)
    marlin = dict(B=quant.pack_marlin_weights(q.astype(np.int64)).to(DEV),
                  scales=quant.permute_marlin_scales(s).to(DEV),
                  zeros=quant.marlin_zero_points(z.astype(np.int64)).to(DEV))
    return ours, marlin


def _marlin(ref, a, m, N):
    out = torch.empty(a.shape[0], N, dtype=torch.bfloat16, device=DEV)
    ws = torch.zeros(N // 64 * 16, dtype=torch.int32, device=DEV)
    e = torch.empty(0, dtype=torch.int32, device=DEV)
    ref.marlin_gemm(a.contiguous(), m["B"], out, m["scales"], m["zeros"], e, e, ws, 4, True, True, True)
    return out


def _close(got, want, what, ulps=1.0, rel=2e-3):
    d = (got.float() - want.float()).abs()
    scale = want.float().abs().max().item()
    tol = ulps * 2.0 ** -8 * want.float().abs() + rel * scale
    assert bool((d <= tol).all()), (what, float(d.max()), scale)
    return d.mean().item() / max(want.float().abs().mean().item(), 1e-30)


def test_llama3_8b_layer_vs_the_references_own_kernels():
    from scalellm_b200.decode_step import (BlockPool, LlamaArgs, LlamaDecoder, StepBuffers, build_decode_batch)
    from scalellm_b200.layers import QuantArgs
    from scalellm_b200.model_parallel import ParallelArgs
    ref = _ref()
    B, S, bs, g = 64, 2048, 8, 128
    args = LlamaArgs.llama3_8b()
    args.n_layers, args.vocab_size = 1, 1024
    h, H, Hkv, D, I = args.hidden_size, args.n_heads, args.n_kv_heads, args.head_dim, args.intermediate_size
    qa = QuantArgs(quant_method="awq", bits=4, group_size=g)
    model = LlamaDecoder(args, qa, ParallelArgs(0, 1, None), torch.device(DEV))
    shapes = dict(qkv=(h, (H + 2 * Hkv) * D), o=(H * D, h), gate_up=(h, 2 * I), down=(I, h))
    ours, marl = {}, {}
    for i, (name, (K, N)) in enumerate(shapes.items()):
        ours[name], marl[name] = _awq_linear(K, N, g, seed=10 + i)
    gen = torch.Generator().manual_seed(1)
    w_in = (1 + 0.1 * torch.randn(h, generator=gen)).bfloat16().to(DEV)
    w_post = (1 + 0.1 * torch.randn(h, generator=gen)).bfloat16().to(DEV)
    model.load_layer(0, dict(ours, input_norm=w_in, post_norm=w_post))
    x = (torch.randn(B, h, generator=gen) * 0.5).bfloat16().to(DEV)
    model.embed[:B].copy_(x)
    model.final_norm.weight.fill_(1.0)
    model.lm_head.weight.zero_()

    nblk = (S + bs - 1) // bs + 1
    n_blocks = B * nblk + 8
    pool = BlockPool(n_blocks, bs, seed=2)
    for _ in range(B):
        pool.add_sequence(S + 4)
    model.alloc_kv(n_blocks, bs, randomize=True, seed=3)
    kc0, vc0 = model.kv_caches[0].key_cache.clone(), model.kv_caches[0].value_cache.clone()
    hb = build_decode_batch(pool, [S] * B, [1] * B, args.vocab_size)
    hb.tokens[:] = np.arange(B)                       # token t -> hidden state x[t]
    bufs = StepBuffers(torch.device(DEV), B, B, B * nblk)
    tokens, positions, params = bufs.upload(hb)

    # ---------------- ours: the fused product path ----------------
    _, h_ours = model.forward(tokens, positions, params, return_hidden=True)
    torch.cuda.synchronize()
    kc_ours, vc_ours = model.kv_caches[0].key_cache, model.kv_caches[0].value_cache

    # ---------------- the reference's kernels, op by op ----------------
    cs = model.handler.pos_emb.cos_sin_cache
    sm_scale = D ** -0.5
    n1 = torch.empty_like(x)
    ref.rms_norm(n1, x, w_in, args.rms_norm_eps)
    qkv = _marlin(ref, n1, marl["qkv"], (H + 2 * Hkv) * D)
    qkv_pre = qkv.clone()
    q = qkv[:, : H * D].view(B, H, D)
    k = qkv[:, H * D: (H + Hkv) * D].view(B, Hkv, D)
    v = qkv[:, (H + Hkv) * D:].view(B, Hkv, D)
    ref.apply_rotary_pos_emb(q, k, positions, cs, D, False)
    kc_ref, vc_ref = kc0.clone(), vc0.clone()
    ref.set_kv_cache(params.new_cache_slots, k, v, kc_ref, vc_ref)
    attn = torch.empty(B, H, D, dtype=torch.bfloat16, device=DEV)
    ref.paged_kv_varlen_mha(attn, q, kc_ref, vc_ref, params.q_cu_seq_lens, params.kv_cu_seq_lens,
                            params.block_tables, params.cu_block_lens, None, bs, 1, hb.kv_max, sm_scale, 0.0, -1)
    o = _marlin(ref, attn.view(B, H * D), marl["o"], h)
    h1 = x + o
    n2 = torch.empty_like(h1)
    ref.rms_norm(n2, h1, w_post, args.rms_norm_eps)
    gu = _marlin(ref, n2, marl["gate_up"], 2 * I)
    act = ref.silu(gu[:, :I].contiguous()) * gu[:, I:]        # llama.h:63: silu kernel, torch mul
    dn = _marlin(ref, act, marl["down"], h)
    h_ref = h1 + dn
    torch.cuda.synchronize()

    # ---------------- (1) every op of ours on the reference's intermediates ----------------
    o_b = torch.empty_like(x)
    kernels.rms_norm(o_b, x, w_in, args.rms_norm_eps)
    assert torch.equal(o_b, n1)
    L = model.layers[0]
    rel = {}
    rel["qkv"] = _close(L["qkv"](n1), qkv_pre, "qkv gemm")
    q2, k2 = qkv_pre[:, : H * D].clone().view(B, H, D), qkv_pre[:, H * D: (H + Hkv) * D].clone().view(B, Hkv, D)
    kernels.apply_rotary_pos_emb(q2, k2, positions, cs, D, False)
    assert torch.equal(q2, q) and torch.equal(k2, k)
    kc_b, vc_b = kc0.clone(), vc0.clone()
    kernels.set_kv_cache(params.new_cache_slots, k, v, kc_b, vc_b)
    assert torch.equal(kc_b, kc_ref) and torch.equal(vc_b, vc_ref)
    a_b = torch.empty_like(attn)
    kernels.paged_kv_varlen_mha(a_b, q, kc_ref, vc_ref, params.q_cu_seq_lens, params.kv_cu_seq_lens,
                                params.block_tables, params.cu_block_lens, None, bs, 1, hb.kv_max,
                                sm_scale, 0.0, -1)
    rel["attention"] = _close(a_b, attn, "attention", ulps=2.0)
    rel["o"] = _close(L["o"](attn.view(B, H * D)), o, "o gemm")
    kernels.rms_norm(o_b, h1, w_post, args.rms_norm_eps)
    assert torch.equal(o_b, n2)
    rel["gate_up"] = _close(L["gate_up"](n2), gu, "gate_up gemm")
    assert torch.equal(kernels.silu_mul(gu[:, :I], gu[:, I:]), act)
    rel["down"] = _close(L["down"](act), dn, "down gemm")

    # ---------------- (2) the whole layer, our own intermediates ----------------
    d = (h_ours.float() - h_ref.float()).abs()
    scale = h_ref.float().abs().max().item()
    mean_rel = d.mean().item() / h_ref.float().abs().mean().item()
    assert d.mean().item() <= 1e-3 * scale, (d.mean().item(), scale)     # north star, relative to the output scale
    assert bool((d <= 2 * 2.0 ** -8 * h_ref.float().abs() + 4e-3 * scale).all()), (float(d.max()), scale)
    # fp32 evaluation of the same layer: neither pipeline is closer to it than the other
    W = {name: kernels.w4a16_dequant(L[name].packed, *shapes[name], g).float() for name in shapes}
    f = lambda t: t.float()
    rms = lambda t, w: t * torch.rsqrt((t * t).mean(-1, keepdim=True) + args.rms_norm_eps) * f(w)
    t_qkv = rms(f(x), w_in) @ W["qkv"]
    tq, tk, tv = (t_qkv[:, : H * D].view(B, H, D), t_qkv[:, H * D: (H + Hkv) * D].view(B, Hkv, D),
                  t_qkv[:, (H + Hkv) * D:].view(B, Hkv, D))
    cos, sin = f(cs[positions.long(), : D // 2])[:, None, :], f(cs[positions.long(), D // 2:])[:, None, :]
    rot = lambda t: torch.cat([t[..., : D // 2] * cos - t[..., D // 2:] * sin,
                               t[..., : D // 2] * sin + t[..., D // 2:] * cos], -1)
    tq, tk = rot(tq), rot(tk)
    pos = torch.arange(S, device=DEV)
    tab = torch.from_numpy(pool._table[:B].astype(np.int64)).to(DEV)             # [B, blocks] first-slot ids
    slot_of = tab[:, pos // bs] + (pos % bs)[None, :]                            # [B, S]
    Kf, Vf = f(kc0[slot_of]), f(vc0[slot_of])                                    # [B, S, Hkv, D]
    Kf[:, S - 1], Vf[:, S - 1] = tk, tv                                          # the step's own token
    G = H // Hkv
    sc_ = torch.einsum("bhgd,bshd->bhgs", tq.view(B, Hkv, G, D), Kf) * sm_scale
    t_attn = torch.einsum("bhgs,bshd->bhgd", torch.softmax(sc_, -1), Vf).reshape(B, H * D)
    t_h1 = f(x) + t_attn @ W["o"]
    t_gu = rms(t_h1, w_post) @ W["gate_up"]
    t_h2 = t_h1 + (torch.nn.functional.silu(t_gu[:, :I]) * t_gu[:, I:]) @ W["down"]
    err_ours = (h_ours.float() - t_h2).abs().mean().item()
    err_ref = (h_ref.float() - t_h2).abs().mean().item()
    assert err_ours <= 1.25 * err_ref + 1e-6 * scale, (err_ours, err_ref)
    print(f"fp32 evaluation: mean |ours - fp32| = {err_ours:.3e}, mean |reference kernels - fp32| = {err_ref:.3e}, "
          f"output scale {scale:.3f}")
    # KV: exactly the reference's slots were written (everything else untouched), V and K within
    # the qkv GEMM's tolerance of the reference's values
    slots = params.new_cache_slots.long()
    mask = torch.ones(kc0.shape[0], dtype=torch.bool, device=DEV)
    mask[slots] = False
    assert torch.equal(kc_ours[mask], kc0[mask]) and torch.equal(vc_ours[mask], vc0[mask])
    _close(vc_ours[slots], vc_ref[slots], "v cache")
    _close(kc_ours[slots], kc_ref[slots], "k cache", ulps=2.0)
    print("mean rel err per accumulating op vs the reference kernel:", {k: f"{v:.2e}" for k, v in rel.items()},
          "layer:", f"{mean_rel:.2e}")
