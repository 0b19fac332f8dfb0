"""The C++ host side (shim/): the reference's operator signatures over the C ABI.
CPU: the extension loads and exposes the reference's Python kernel surface
(scalellm/csrc/kernels.cu:9-55 names).  GPU: results are bit-identical to the ctypes path."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_shim():
    sys.path.insert(0, os.path.join(ROOT, "scalellm_b200"))
    try:
        import _b200_shim
    finally:
        sys.path.pop(0)
    return _b200_shim


def test_shim_loads_and_exports_reference_surface():
    m = load_shim()
    for name in ("rms_norm", "rms_norm_residual", "apply_rotary_pos_emb", "set_kv_cache", "silu",
                 "silu_with_mul", "paged_kv_varlen_mha", "marlin_awq_repack", "marlin_gptq_repack",
                 "marlin_gemm"):
        assert hasattr(m, name), name
    assert m.packed_bytes(4096, 4096, 128) == 32 * 32 * (8192 + 256 + 128)


@pytest.mark.gpu
def test_shim_matches_ctypes_path():
    from scalellm_b200 import kernels
    from oracle import quant
    m = load_shim()
    dev = "cuda"
    torch.manual_seed(0)
    x = torch.randn(64, 4096, device=dev).bfloat16()
    w = torch.randn(4096, device=dev).bfloat16()
    o1, o2 = torch.empty_like(x), torch.empty_like(x)
    kernels.rms_norm(o1, x, w, 1e-5)
    m.rms_norm(o2, x, w, 1e-5)
    assert torch.equal(o1, o2)
    g = x[:, :2048].contiguous()
    assert torch.equal(kernels.silu(g), m.silu(g))
    assert torch.equal(kernels.silu_with_mul(x), m.silu_with_mul(x))
    # W4A16 through the marlin:: names
    K, N, M = 1024, 512, 48
    ck = quant.random_awq_checkpoint(K, N, 128, seed=3)
    qw, qz, sc = ck["qweight"].to(dev), ck["qzeros"].to(dev), ck["scales"].to(dev)
    packed = torch.empty(m.packed_bytes(K, N, 128), dtype=torch.uint8, device=dev)
    m.marlin_awq_repack(qw, qz, sc, packed, 128)
    assert torch.equal(packed, kernels.w4a16_prepack_awq(qw, qz, sc, 128))
    a = torch.randn(M, K, device=dev).bfloat16()
    c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ws = torch.zeros(m.workspace_bytes(M, N, K), dtype=torch.uint8, device=dev)
    empty = torch.empty(0, dtype=torch.int32, device=dev)
    m.marlin_gemm(a, packed, c, sc, empty, empty, empty, ws, 4, True, True, True)
    assert torch.equal(c, kernels.w4a16_gemm(a, packed, N, 128))
    # paged attention through llm::paged_kv_varlen_mha
    H, Hkv, D, bs = 8, 2, 128, 8
    kv_lens = [100, 37]
    nblk = [(k + bs - 1) // bs for k in kv_lens]
    table = (torch.randperm(sum(nblk) + 3)[: sum(nblk)] * bs).to(torch.int32).to(dev)
    kc = torch.randn((sum(nblk) + 3) * bs, Hkv, D, device=dev).bfloat16()
    vc = torch.randn_like(kc)
    q = torch.randn(2, H, D, device=dev).bfloat16()
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    args = (q, kc, vc, i32([0, 1, 2]), i32([0, 100, 137]), table, i32([0, nblk[0], sum(nblk)]), None, bs,
            1, 100, D ** -0.5, 0.0, -1)
    out1, out2 = torch.empty_like(q), torch.empty_like(q)
    kernels.paged_kv_varlen_mha(out1, *args)
    m.paged_kv_varlen_mha(out2, *args)
    assert torch.equal(out1, out2)


def test_process_group_bindings_fail_loudly_without_gpu():
    """The C++ ProcessGroup plumbing (shim/b200_process_group) is part of the shim; creating the
    communicators needs GPUs and must raise, not abort, without them."""
    import torch
    shim = load_shim()
    assert hasattr(shim, "ProcessGroup") and hasattr(shim, "create_process_groups")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="ar_create_all"):
            shim.create_process_groups([0, 1])


def test_cpp_decode_demo_builds_and_refuses_to_run_without_a_gpu():
    """shim/decode_demo.cpp: a C++ main() that links the shim + libb200decode and drives the decode
    step; without a CUDA device it must say so and exit non-zero (no CPU path)."""
    import os
    import subprocess
    import torch
    import __graft_entry__ as g
    g._build_shim()                                      # also links the demo next to the extension
    exe = os.path.join(os.path.dirname(__file__), "..", "scalellm_b200", "decode_demo")
    assert os.path.exists(exe)
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "1", "2", "16", "1"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 2 and "no CUDA device" in r.stderr
