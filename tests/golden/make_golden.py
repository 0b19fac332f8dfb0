"""Generate the committed golden fixtures by running the REFERENCE's own Python.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
Imports, unmodified:
    /root/reference/tests/kernels/quant_utils.py            (quantize / pack AWQ, GPTQ)
    /root/reference/tests/kernels/attention/ref_attention.py (paged var-len attention reference)
and reads /root/reference/src/layers/quantization/data/gptq_small.safetensors (real GPTQ tensors).
Outputs (small, committed): quant_golden.npz, marlin_golden.npz, attn_golden.npz, gptq_small.npz,
llama3_rope_inv_freq.json (numbers transcribed from src/layers/pos_embedding_test.cpp:98-138).
Nothing under tests/ or the product reads /root/reference at run time.
"""
import json
import zlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "tests", "kernels"))
sys.path.insert(0, os.path.join(REF, "tests", "kernels", "attention"))
import quant_utils as qu  # noqa: E402
import ref_attention as ra  # noqa: E402


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).view(torch.int16).numpy().copy()


def make_quant():
    out = {}
    for tag, (K, N, g) in {"a": (256, 256, 128), "b": (128, 128, 32), "c": (256, 128, -1)}.items():
        torch.manual_seed(100 + K + N)
        w = torch.randn(K, N, dtype=torch.float32).to(torch.bfloat16)
        w_ref, q_w, s, _, _ = qu.quantize_weights(w, num_bits=4, group_size=g, act_order=False)
        out[f"{tag}_shape"] = np.array([K, N, g])
        out[f"{tag}_q"] = q_w.numpy().astype(np.int8)
        out[f"{tag}_scales_bf16"] = bf16_bits(s)
        out[f"{tag}_wref_bf16"] = bf16_bits(w_ref)           # (q - 8) * s evaluated in bf16
        out[f"{tag}_awq_packed"] = qu.pack_awq_weights(q_w, 4).numpy()
        out[f"{tag}_gptq_packed"] = qu.pack_gptq_weights(q_w, 4).numpy()
    np.savez_compressed(os.path.join(HERE, "quant_golden.npz"), **out)


def make_attn():
    out = {}
    cases = {
        # tag: (q_lens, kv_lens, H, Hkv, D, bs, soft_cap, window, alibi)
        "decode_gqa": ([1, 1, 1], [37, 100, 64], 8, 2, 128, 8, 0.0, -1, False),
        "mixed_window": ([1, 3, 2], [29, 50, 17], 4, 4, 64, 4, 0.0, 10, False),
        "alibi_cap": ([2, 1], [40, 23], 6, 3, 64, 16, 30.0, -1, True),
    }
    for tag, (q_lens, kv_lens, H, Hkv, D, bs, cap, win, alibi) in cases.items():
        torch.manual_seed(zlib.crc32(tag.encode()) % 1000)
        n_seqs = len(q_lens)
        max_blocks = max((kv + bs - 1) // bs for kv in kv_lens)
        n_blocks = n_seqs * max_blocks + 3
        perm = torch.randperm(n_blocks)
        block_tables = perm[: n_seqs * max_blocks].view(n_seqs, max_blocks).to(torch.int32)
        kc = torch.randn(n_blocks, bs, Hkv, D).to(torch.bfloat16)
        vc = torch.randn(n_blocks, bs, Hkv, D).to(torch.bfloat16)
        q = torch.randn(sum(q_lens), H, D).to(torch.bfloat16)
        slopes = torch.rand(H, dtype=torch.float32) if alibi else None
        sm_scale = D ** -0.5
        o = ra.varlen_masked_self_attention(q, kc, vc, q_lens, kv_lens, block_tables, sm_scale,
                                            logits_soft_cap=cap, sliding_window=win,
                                            alibi_slopes=slopes)
        out[f"{tag}_meta"] = np.array([H, Hkv, D, bs, win], dtype=np.int64)
        out[f"{tag}_cap"] = np.array([cap, sm_scale], dtype=np.float64)
        out[f"{tag}_q_lens"] = np.array(q_lens)
        out[f"{tag}_kv_lens"] = np.array(kv_lens)
        out[f"{tag}_block_ids"] = block_tables.numpy()
        out[f"{tag}_q"] = bf16_bits(q)
        out[f"{tag}_kc"] = bf16_bits(kc)
        out[f"{tag}_vc"] = bf16_bits(vc)
        out[f"{tag}_out"] = bf16_bits(o)
        if alibi:
            out[f"{tag}_slopes"] = slopes.numpy()
    np.savez_compressed(os.path.join(HERE, "attn_golden.npz"), **out)


def make_marlin():
    """Symmetric 4-bit, group 128 weights in BOTH formats from the reference's utilities: the GPTQ
    checkpoint packing our prepack consumes and the Marlin packing + scale permutation the
    reference's own GEMM kernel consumes (tests/kernels/marlin_gemm_test.py:82-95).  Used by
    tests/test_gpu_vs_reference_kernels.py to run marlin::gptq_gemm and b200_w4a16_gemm on the same
    weights.  K x N = 512 x 512 keeps the fixture small."""
    K, N, g = 512, 512, 128
    torch.manual_seed(77)
    w = torch.randn(K, N, dtype=torch.float32).to(torch.bfloat16)
    w_ref, q_w, s, _, _ = qu.quantize_weights(w, num_bits=4, group_size=g, act_order=False)
    out = dict(shape=np.array([K, N, g]),
               q=q_w.numpy().astype(np.int8),
               scales_bf16=bf16_bits(s),
               gptq_packed=qu.pack_gptq_weights(q_w, 4).numpy(),
               marlin_packed=qu.pack_marlin_weights(q_w, num_bits=4).numpy(),
               marlin_scales_bf16=bf16_bits(qu.permute_marlin_scales(s)))
    np.savez_compressed(os.path.join(HERE, "marlin_golden.npz"), **out)


def make_gptq_small():
    from safetensors.torch import load_file
    sd = load_file(os.path.join(REF, "src/layers/quantization/data/gptq_small.safetensors"))
    out = {}
    for k, v in sd.items():
        out[k.replace(".", "_")] = (v.view(torch.int16).numpy() if v.dtype == torch.float16
                                    else v.numpy())
        print("gptq_small:", k, tuple(v.shape), v.dtype)
    np.savez_compressed(os.path.join(HERE, "gptq_small.npz"), **out)


def make_rope():
    # src/layers/pos_embedding_test.cpp:98-138 (RopeScalingTest.Llama3), rtol 1e-4
    default = [1.0000e+00, 8.1462e-01, 6.6360e-01, 5.4058e-01, 4.4037e-01, 3.5873e-01, 2.9223e-01,
               2.3805e-01, 1.9392e-01, 1.5797e-01, 1.2869e-01, 1.0483e-01, 8.5397e-02, 6.9566e-02,
               5.6670e-02, 4.6164e-02, 3.7606e-02, 3.0635e-02, 2.4955e-02, 2.0329e-02, 1.6560e-02,
               1.3490e-02, 1.0990e-02, 8.9523e-03, 7.2927e-03, 5.9407e-03, 4.8394e-03, 3.9423e-03,
               3.2114e-03, 2.6161e-03, 2.1311e-03, 1.7360e-03, 1.4142e-03, 1.1520e-03, 9.3847e-04,
               7.6450e-04, 6.2277e-04, 5.0732e-04, 4.1327e-04, 3.3666e-04, 2.7425e-04, 2.2341e-04,
               1.8199e-04, 1.4825e-04, 1.2077e-04, 9.8381e-05, 8.0143e-05, 6.5286e-05, 5.3183e-05,
               4.3324e-05, 3.5292e-05, 2.8750e-05, 2.3420e-05, 1.9078e-05, 1.5542e-05, 1.2660e-05,
               1.0313e-05, 8.4015e-06, 6.8440e-06, 5.5752e-06, 4.5417e-06, 3.6997e-06, 3.0139e-06,
               2.4551e-06]
    scaled = default[:29] + [
        2.1666e-03, 1.3719e-03, 8.5675e-04, 5.2485e-04, 3.1269e-04, 1.7851e-04, 9.5562e-05,
        7.7847e-05, 6.3415e-05, 5.1659e-05, 4.2082e-05, 3.4281e-05, 2.7926e-05, 2.2749e-05,
        1.8532e-05, 1.5096e-05, 1.2298e-05, 1.0018e-05, 8.1607e-06, 6.6479e-06, 5.4155e-06,
        4.4115e-06, 3.5937e-06, 2.9275e-06, 2.3848e-06, 1.9427e-06, 1.5826e-06, 1.2892e-06,
        1.0502e-06, 8.5550e-07, 6.9690e-07, 5.6771e-07, 4.6247e-07, 3.7673e-07, 3.0689e-07]
    assert len(default) == 64 and len(scaled) == 64
    json.dump(dict(rotary_dim=128, theta=500000.0, factor=8.0, low_freq_factor=1.0,
                   high_freq_factor=4.0, old_context_len=8192, rtol=1e-4,
                   expected_inv_freq=default, expected_scaled_inv_freq=scaled),
              open(os.path.join(HERE, "llama3_rope_inv_freq.json"), "w"), indent=1)


if __name__ == "__main__":
    make_quant()
    make_marlin()
    make_attn()
    make_gptq_small()
    make_rope()
    print("golden fixtures written to", HERE)
