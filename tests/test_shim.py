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
    # the top-k / top-p filter through llm::kernel::apply_top_k_top_p (top_p handed over as fp16: converted)
    lg = (torch.randn(5, 3000, device=dev) * 3).bfloat16()
    tk = torch.tensor([10, 0, 5, 0, 50], dtype=torch.int64, device=dev)
    tp = torch.tensor([0.9, 0.5, 1.0, 1.0, 0.25], dtype=torch.float32, device=dev)
    f1, f2, f3 = lg.clone(), lg.clone(), lg.clone()
    kernels.apply_top_k_top_p(f1, tk, tp)
    m.apply_top_k_top_p(f2, tk, tp)
    assert torch.equal(f1.view(torch.int16), f2.view(torch.int16))
    m.apply_top_k_top_p(f3, None, tp)
    kernels.apply_top_k_top_p(lg, None, tp)
    assert torch.equal(f3.view(torch.int16), lg.view(torch.int16))


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


_REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(_REF, "src", "kernels")),
                    reason="needs the reference tree (dev container only)")
def test_a_translation_unit_against_the_references_own_headers_links_with_the_shim(tmp_path):
    """SURVEY 8b "operator-level signatures": a caller compiled against the REFERENCE's headers
    (layernorm_kernels.h, pos_embedding_kernels.h, kv_cache_kernels.h, activation_kernels.h,
    attention/attn_api.h, quantization/marlin.h — untouched, from /root/reference) links against
    _b200_shim.so with every symbol it uses resolved: the link-time swap of :kernels,
    :attention.kernels and :marlin.kernels needs no source edit."""
    import subprocess
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    import __graft_entry__ as g
    g._build_shim()
    src = tmp_path / "caller.cpp"
    src.write_text(r'''
#include "layernorm_kernels.h"
#include "pos_embedding_kernels.h"
#include "kv_cache_kernels.h"
#include "activation_kernels.h"
#include "attention/attn_api.h"
#include "quantization/marlin.h"
// take the address of every function the reference's layers call, with the reference's types
void* table[] = {
  (void*)static_cast<void (*)(torch::Tensor&, torch::Tensor, torch::Tensor, float)>(&llm::kernel::rms_norm),
  (void*)static_cast<void (*)(torch::Tensor&, torch::Tensor, torch::Tensor, float)>(&llm::kernel::gemma_rms_norm),
  (void*)static_cast<void (*)(torch::Tensor&, torch::Tensor&, torch::Tensor, torch::Tensor, float)>(&llm::kernel::rms_norm_residual),
  (void*)static_cast<void (*)(torch::Tensor&, torch::Tensor, torch::Tensor, torch::Tensor, float)>(&llm::kernel::layer_norm),
  (void*)static_cast<void (*)(torch::Tensor&, torch::Tensor&, const torch::Tensor&, const torch::Tensor&, int, bool)>(&llm::kernel::apply_rotary_pos_emb),
  (void*)static_cast<void (*)(const torch::Tensor&, const torch::Tensor&, const torch::Tensor&, torch::Tensor&, torch::Tensor&)>(&llm::kernel::set_kv_cache),
  (void*)static_cast<torch::Tensor (*)(torch::Tensor)>(&llm::kernel::gelu_new),
  (void*)static_cast<torch::Tensor (*)(torch::Tensor)>(&llm::kernel::gelu_fast),
  (void*)static_cast<torch::Tensor (*)(torch::Tensor)>(&llm::kernel::silu),
  (void*)static_cast<torch::Tensor (*)(torch::Tensor)>(&llm::kernel::gelu_new_with_mul),
  (void*)static_cast<torch::Tensor (*)(torch::Tensor)>(&llm::kernel::gelu_fast_with_mul),
  (void*)static_cast<torch::Tensor (*)(torch::Tensor)>(&llm::kernel::silu_with_mul),
  (void*)&llm::paged_kv_varlen_mha,
  (void*)&marlin::gptq_gemm,
  (void*)&marlin::gptq_repack,
  (void*)&marlin::awq_repack,
};
int main() { return table[0] == nullptr; }
''')
    try:
        inc = ce.include_paths(device_type="cuda")
    except TypeError:
        inc = ce.include_paths(cuda=True)
    tl = os.path.join(os.path.dirname(torch.__file__), "lib")
    so_dir = os.path.join(ROOT, "scalellm_b200")
    cmd = (["g++", "-std=c++17", "-O0", str(src), "-o", str(tmp_path / "caller"),
            "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI)),
            "-I" + os.path.join(_REF, "src", "kernels")] + ["-I" + i for i in inc] +
           ["-I" + sysconfig.get_paths()["include"], "-I/usr/local/cuda/include",
            os.path.join(so_dir, "_b200_shim.so"), "-L" + so_dir, "-lb200decode",
            "-L" + tl, "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
            "-L" + sysconfig.get_config_var("LIBDIR"), "-lpython" + sysconfig.get_config_var("LDVERSION"),
            "-Wl,-rpath," + so_dir, "-Wl,-rpath," + tl, "-Wl,--no-undefined"])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.gpu
def test_marlin_drop_in_with_the_references_exact_signatures():
    """marlin::awq_repack(q_weight, out, num_bits) / gptq_repack(q_weight, perm, out, num_bits) /
    gptq_gemm(..., scales and zeros in Marlin's order, has_zp, g_idx, perm ...) exactly as
    qlinear_awq_marlin_impl.cpp:99-125,332-365 and qlinear_gptq_marlin_impl.cpp:43-72 call them:
    results bit-identical to the one-step B200 prepack path, zeros honoured, act-order honoured,
    fp16 refused instead of reinterpreted, the weight assembled once."""
    import numpy as np
    from scalellm_b200 import kernels
    from oracle import quant
    m = load_shim()
    dev = "cuda"
    K, N, M, g = 1024, 512, 48, 128
    e = torch.empty(0, dtype=torch.int32, device=dev)
    ws = torch.zeros(N // 64 * 16, dtype=torch.int32, device=dev)      # Marlin's lock workspace
    a = torch.randn(M, K, device=dev).bfloat16()
    # ---- AWQ: has_zp = true ----
    ck = quant.random_awq_checkpoint(K, N, g, seed=11)
    qw, qz, sc = ck["qweight"].to(dev), ck["qzeros"].to(dev), ck["scales"].to(dev)
    want = kernels.w4a16_gemm(a, kernels.w4a16_prepack_awq(qw, qz, sc, g), N, g)
    out = torch.empty(K // 16, N * 2, dtype=torch.int32, device=dev)
    m.marlin_awq_repack_ref(qw, out, 4)
    ms = quant.permute_marlin_scales(ck["scales"]).to(dev)
    mz = quant.marlin_zero_points(quant.unpack_awq(ck["qzeros"])).to(dev)
    n0 = m.assembled_weights()
    c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        m.marlin_gemm(a, out, c, ms, mz, e, e, ws, 4, True, True, True)
    assert torch.equal(c, want) and m.assembled_weights() == n0 + 1 and int(ws.abs().sum()) == 0
    # the zero points are really read: has_zp = false means the symmetric 8
    m.marlin_gemm(a, out, c, ms, e, e, e, ws, 4, True, False, True)
    sym = kernels.w4a16_gemm(a, kernels.w4a16_prepack_awq(
        qw, quant.pack_awq(np.full((K // g, N), 8)).to(dev), sc, g), N, g)
    assert torch.equal(c, sym) and not torch.equal(c, want)
    # fp16 activations are refused, not reinterpreted
    with pytest.raises(RuntimeError, match="bf16"):
        m.marlin_gemm(a.half(), out, c, ms, mz, e, e, ws, 4, True, True, True)
    # ---- GPTQ act-order: gptq_repack(perm) + gptq_gemm(g_idx sorted, perm) ----
    rng = np.random.default_rng(3)
    q = rng.integers(0, 16, size=(K, N))
    order = rng.permutation(K)
    g_idx = np.empty(K, dtype=np.int32)
    g_idx[order] = np.arange(K) // g
    s2 = (torch.randn(K // g, N).abs() * 0.01 + 1e-3).bfloat16()
    perm = torch.argsort(torch.from_numpy(g_idx).long(), stable=True).to(torch.int32)
    # the act-order weight by definition: row k uses the scale of group g_idx[k], zero point 8
    w_ref = (s2[torch.from_numpy(g_idx).long()].float() * torch.from_numpy(q - 8).float()).bfloat16()
    out2 = torch.empty(K // 16, N * 2, dtype=torch.int32, device=dev)
    m.marlin_gptq_repack_ref(quant.pack_gptq(q).to(dev), perm.to(dev), out2, 4)
    g_sorted = torch.from_numpy(g_idx)[perm.long()].to(torch.int32).to(dev)
    m.marlin_gemm(a, out2, c, quant.permute_marlin_scales(s2).to(dev), e, g_sorted, perm.to(dev), ws, 4,
                  True, False, True)
    want2 = quant.w4a16_gemm(a.cpu(), w_ref)
    from tests.util import rel_err
    assert rel_err(c, want2) < 1e-3
    with pytest.raises(RuntimeError, match="K-sharded"):
        m.marlin_gemm(a, out2, c, quant.permute_marlin_scales(s2).to(dev), e, g_sorted, perm.to(dev), ws, 4,
                      False, False, True)
