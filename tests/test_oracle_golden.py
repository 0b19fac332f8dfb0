"""CPU: pin the oracle against golden vectors produced by the reference's own Python
(tests/golden/make_golden.py) and the reference tests' known answers."""
import json
import os

import numpy as np
import torch

from oracle import ops, quant
from tests.util import bf16_from_bits, f16_from_bits, assert_ulp


def test_awq_gptq_packing_matches_quant_utils(golden_dir):
    g = np.load(os.path.join(golden_dir, "quant_golden.npz"))
    for tag in "abc":
        q = g[f"{tag}_q"].astype(np.int32)
        awq = torch.from_numpy(g[f"{tag}_awq_packed"])
        gptq = torch.from_numpy(g[f"{tag}_gptq_packed"])
        # unpack(reference pack) == q, and our pack == reference pack, bit exact
        assert np.array_equal(quant.unpack_awq(awq), q)
        assert np.array_equal(quant.unpack_gptq(gptq), q)
        assert torch.equal(quant.pack_awq(q), awq)
        assert torch.equal(quant.pack_gptq(q), gptq)


def test_dequant_matches_quant_utils_wref(golden_dir):
    g = np.load(os.path.join(golden_dir, "quant_golden.npz"))
    for tag in "abc":
        K, N, gs = (int(x) for x in g[f"{tag}_shape"])
        q = g[f"{tag}_q"].astype(np.int32)
        s = bf16_from_bits(g[f"{tag}_scales_bf16"])
        w_ref = bf16_from_bits(g[f"{tag}_wref_bf16"])  # (q_w - 8).to(bf16) * s, quant_utils.py:73
        w = quant.dequant(q, 8, s, gs)
        assert torch.equal(w.view(torch.int16), w_ref.view(torch.int16)), tag


def test_gptq_small_construct_weights_two_ways(golden_dir):
    """qlinear_impl_test.cpp:10-40: construct_weights with and without g_idx agree on the
    reference's real GPTQ tensors."""
    g = np.load(os.path.join(golden_dir, "gptq_small.npz"))
    qweight = torch.from_numpy(g["qweight"])
    qzeros = torch.from_numpy(g["qzeros"])
    scales = f16_from_bits(g["scales"])
    g_idx = torch.from_numpy(g["g_idx"])
    w1 = quant.construct_gptq_weights(qweight, qzeros, scales, g_idx)
    w2 = quant.construct_gptq_weights(qweight, qzeros, scales, None)
    assert w1.shape == (256, 256)
    assert torch.equal(w1, w2)
    # and the explicit loop (the "slow" path of the reference)
    q = quant.unpack_gptq(qweight)
    z = quant.unpack_gptq_zeros(qzeros, plus_one=True)
    for k in (0, 17, 128, 255):
        for n in (0, 5, 255):
            gi = int(g_idx[k])
            ref = scales[gi, n] * torch.tensor(float(q[k, n] - z[gi, n]), dtype=torch.float16)
            assert w1[k, n] == ref


def _run_attn_case(g, tag):
    H, Hkv, D, bs, win = (int(x) for x in g[f"{tag}_meta"])
    cap, sm_scale = (float(x) for x in g[f"{tag}_cap"])
    q_lens = [int(x) for x in g[f"{tag}_q_lens"]]
    kv_lens = [int(x) for x in g[f"{tag}_kv_lens"]]
    block_ids = g[f"{tag}_block_ids"]
    q = bf16_from_bits(g[f"{tag}_q"])
    kc = bf16_from_bits(g[f"{tag}_kc"]).reshape(-1, Hkv, D)   # [n_blocks*bs, Hkv, D]
    vc = bf16_from_bits(g[f"{tag}_vc"]).reshape(-1, Hkv, D)
    slopes = torch.from_numpy(g[f"{tag}_slopes"]) if f"{tag}_slopes" in g.files else None
    # reference python uses block ids [B, max_blocks]; the C++ API uses flattened first-slot ids
    table, blk_cu = [], [0]
    for b, kv in enumerate(kv_lens):
        nb = (kv + bs - 1) // bs
        table.extend((block_ids[b, :nb].astype(np.int64) * bs).tolist())
        blk_cu.append(blk_cu[-1] + nb)
    q_cu = np.concatenate([[0], np.cumsum(q_lens)])
    kv_cu = np.concatenate([[0], np.cumsum(kv_lens)])
    out = ops.paged_attention(q, kc, vc, q_cu, kv_cu, torch.tensor(table, dtype=torch.int32),
                              blk_cu, bs, sm_scale, slopes, cap, win)
    return out, bf16_from_bits(g[f"{tag}_out"])


def test_paged_attention_matches_ref_attention_py(golden_dir):
    g = np.load(os.path.join(golden_dir, "attn_golden.npz"))
    for tag in ("decode_gqa", "mixed_window", "alibi_cap"):
        out, ref = _run_attn_case(g, tag)
        assert out.shape == ref.shape
        # both are fp32-softmax references rounded once to bf16
        assert_ulp(out, ref, max_ulp=1, max_frac=0.01, what=tag)


def test_llama3_rope_scaling_known_answers(golden_dir):
    """RopeScalingTest.Llama3 (src/layers/pos_embedding_test.cpp:98-138)."""
    k = json.load(open(os.path.join(golden_dir, "llama3_rope_inv_freq.json")))
    inv = ops.compute_default_inv_freq(k["rotary_dim"], k["theta"])
    assert torch.allclose(inv, torch.tensor(k["expected_inv_freq"]), rtol=k["rtol"], atol=0)
    scaled = ops.apply_llama3_rope_scaling(inv, k["factor"], k["low_freq_factor"],
                                           k["high_freq_factor"], k["old_context_len"])
    assert torch.allclose(scaled, torch.tensor(k["expected_scaled_inv_freq"]), rtol=k["rtol"],
                          atol=0)
    # the product's vectorised version agrees with the oracle
    from scalellm_b200.layers import apply_llama3_rope_scaling, compute_default_inv_freq
    inv2 = compute_default_inv_freq(k["rotary_dim"], k["theta"])
    sc2 = apply_llama3_rope_scaling(inv2, k["factor"], k["low_freq_factor"],
                                    k["high_freq_factor"], k["old_context_len"])
    assert torch.allclose(sc2, torch.tensor(k["expected_scaled_inv_freq"]), rtol=k["rtol"], atol=0)


def test_rope_oracle_equals_torch_ops_in_dtype():
    """The oracle's per-op rounding == RotaryEmbeddingGeneric evaluated with torch ops in bf16
    (pos_embedding.cpp:39-53: q*cos + rotate_half(q)*sin), the comparison the reference test makes."""
    torch.manual_seed(0)
    T, H, Hkv, D = 5, 4, 2, 64
    q = torch.randn(T, H, D).bfloat16()
    k = torch.randn(T, Hkv, D).bfloat16()
    inv = ops.compute_default_inv_freq(D, 10000.0)
    cs = ops.build_cos_sin_cache(D, 128, inv, torch.bfloat16)
    pos = torch.tensor([0, 3, 17, 99, 127], dtype=torch.int32)
    qo, ko = ops.rope(q, k, pos, cs, D, interleaved=False)

    def generic(x):
        c, s = cs[pos.long()].chunk(2, dim=-1)
        c = torch.cat([c, c], -1)[:, None, :]
        s = torch.cat([s, s], -1)[:, None, :]
        x1, x2 = x.chunk(2, dim=-1)
        return (x * c) + (torch.cat([-x2, x1], -1) * s)

    assert torch.equal(qo, generic(q))
    assert torch.equal(ko, generic(k))


def test_rms_norm_oracle_matches_layer_formula():
    """normalization.h:17-52 (detail::rms_norm) within the reference's own tolerance."""
    torch.manual_seed(0)
    x = torch.randn(7, 1038).bfloat16()
    w = torch.randn(1038).bfloat16()
    xf = x.float()
    ref = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).bfloat16() * w
    out = ops.rms_norm(x, w, 1e-5)
    assert torch.allclose(out.float(), ref.float(), rtol=1e-2, atol=1e-3)
    o2, r2 = ops.rms_norm_residual(x, x.clone(), w, 1e-5)
    assert torch.equal(r2, (x.float() * 2).bfloat16())


def test_marlin_golden_fixture_is_consistent_with_oracle_packers(golden_dir):
    """The fixture that feeds the reference's own Marlin GEMM on the GPU (marlin_golden.npz, written
    by the reference's quant_utils): its GPTQ packing must be what our oracle packs from the same
    integers, and the Marlin tensors must have the layout marlin::gptq_gemm expects
    (src/kernels/quantization/marlin.h:17-28: B (k/16, n*16/8), scales (k/g, n))."""
    import numpy as np
    import torch
    from oracle import quant
    d = np.load(os.path.join(golden_dir, "marlin_golden.npz"))
    K, N, g = (int(x) for x in d["shape"])
    q = d["q"].astype(np.int64)
    assert q.shape == (K, N) and q.min() >= 0 and q.max() <= 15
    assert np.array_equal(quant.pack_gptq(q).numpy(), d["gptq_packed"])
    assert d["marlin_packed"].shape == (K // 16, N * 16 // 8) and d["marlin_packed"].dtype == np.int32
    assert d["marlin_scales_bf16"].shape == d["scales_bf16"].shape == (K // g, N)
    # the permuted scales are a permutation of the original ones, row by row
    assert np.array_equal(np.sort(d["marlin_scales_bf16"], axis=1), np.sort(d["scales_bf16"], axis=1))


def test_marlin_layout_restatement_matches_the_reference_packers(golden_dir):
    """oracle/quant.py's Marlin weight / scale layouts == what the reference's quant_utils wrote
    (tests/golden/marlin_golden.npz), and the inverse maps (used to check the drop-in shim that takes
    Marlin-layout tensors) invert them; the zero-point map (qlinear_awq_marlin_impl.cpp:62-97) is a
    bijection over the same column order as the scales."""
    d = np.load(os.path.join(golden_dir, "marlin_golden.npz"))
    K, N, g = (int(x) for x in d["shape"])
    q = d["q"].astype(np.int64)
    mp = quant.pack_marlin_weights(q)
    assert np.array_equal(mp.numpy(), d["marlin_packed"])
    assert np.array_equal(quant.unpack_marlin_weights(mp, K, N), q)
    s = torch.from_numpy(d["scales_bf16"]).view(torch.bfloat16)
    ms = quant.permute_marlin_scales(s)
    assert torch.equal(ms.view(torch.int16).reshape(-1),
                       torch.from_numpy(d["marlin_scales_bf16"]).view(torch.int16).reshape(-1))
    assert torch.equal(quant.unpermute_marlin_scales(ms), s)
    z = np.random.default_rng(0).integers(0, 16, size=(K // g, N))
    mz = quant.marlin_zero_points(z)
    assert np.array_equal(quant.unpack_marlin_zero_points(mz), z)
    # column c of the natural order lands where scale column c lands (same permutation), then the
    # 4-bit interleave inside every group of 8: spot-check one row against the C++ recipe
    perm, _ = quant.marlin_scales_perm()
    want = z.reshape(-1, 64)[:, perm].reshape(-1, 8)[:, [0, 2, 4, 6, 1, 3, 5, 7]].reshape(z.shape)
    assert np.array_equal(quant.unpack_cols(mz), want)
