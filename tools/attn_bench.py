"""Kernel-only timing of the paged attention decode op at the benchmark shape (B=64, S=2048,
H=32/8, D=128), rotating over several caches so every launch streams from HBM."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from scalellm_b200 import kernels  # noqa: E402
from _timing import time_us  # noqa: E402

DEV = "cuda"


def main():
    shapes = [(64, 2048, 32, 8, 128)]
    if len(sys.argv) > 1 and sys.argv[1] == "ragged":
        # bench.py's situation: kv_len not a multiple of 16 and max_kv_len > kv_len
        one_shape(64, 2071, 32, 8, 128, max_kv=2102, bss=(8,))
        one_shape(64, 2071, 32, 8, 128, max_kv=2071, bss=(8,))
        one_shape(64, 2064, 32, 8, 128, max_kv=2064, bss=(8,))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "variant":
        # A/B of the opt-in instantiations (B200_ATTN_OCC / B200_ATTN_TR, read once per process):
        # the default plan only, benchmark shape, block_size 8, no reference kernel
        one_shape(64, 2048, 32, 8, 128, bss=(8,), only_auto=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "prefill":
        # prefill / chunked-prefill shapes (csrc/prefill_attn.cu; B200_ATTN_PREFILL=0 in the
        # environment times the decode stream kernel on the same problem), the reference's kernel beside
        for B, q_len, kv_len in ((1, 128, 2048), (1, 512, 2048), (1, 2048, 2048), (8, 512, 2048), (4, 2048, 2048)):
            prefill_shape(B, q_len, kv_len, 32, 8, 128)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "tp":
        # the per-rank attention problem under tensor parallelism (heads / kv heads divided by the
        # world size, same sequences): where the fixed cost of a launch shows
        for H, Hkv in ((16, 4), (8, 2), (4, 1)):
            one_shape(64, 2048, H, Hkv, 128, bss=(8,), only_auto=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "locality":
        shapes = [(64, 2048, 32, 8, 128), (512, 2048, 4, 1, 128), (128, 2048, 16, 4, 128)]
    for shp in shapes:
        one_shape(*shp)


def one_shape(B, S, H, Hkv, D, max_kv=None, bss=(8, 128), only_auto=False):
    max_kv = max_kv or S
    print(f"--- B={B} S={S} max_kv={max_kv} H={H} Hkv={Hkv} D={D}")
    for bs in bss:
        nblk = (S + bs - 1) // bs
        n_blocks = B * nblk + 8
        L = 10
        caches = [(torch.randn(n_blocks * bs, Hkv, D, device=DEV).bfloat16(),
                   torch.randn(n_blocks * bs, Hkv, D, device=DEV).bfloat16()) for _ in range(L)]
        perm = torch.randperm(n_blocks)[: B * nblk]
        table = (perm * bs).to(torch.int32).to(DEV)
        i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=DEV)
        q_cu, kv_cu, blk_cu = i32(np.arange(B + 1)), i32(np.arange(B + 1) * S), i32(np.arange(B + 1) * nblk)
        q = torch.randn(B, H, D, device=DEV).bfloat16()
        out = torch.empty_like(q)

        def launch(kc, vc):
            kernels.paged_kv_varlen_mha(out, q, kc, vc, q_cu, kv_cu, table, blk_cu, None, bs, 1, max_kv,
                                        D ** -0.5, 0.0, -1)
        for kc, vc in caches[:3]:
            launch(kc, vc)
        torch.cuda.synchronize()
        configs = [("mma W4 s=2", {"B200_ATTN_IMPL": "mma", "B200_ATTN_WARPS": "4", "B200_ATTN_SPLITS": "2"}),
                   ("stream t32", {"B200_ATTN_TPS": "32"}),
                   ("stream auto", {})]
        if only_auto:
            configs = [(f"occ={os.environ.get('B200_ATTN_OCC', '0')} tr={os.environ.get('B200_ATTN_TR', '0')}", {})]
        for tag, env in configs:
            for k in ("B200_ATTN_WARPS", "B200_ATTN_SPLITS", "B200_ATTN_TPS", "B200_ATTN_IMPL"):
                os.environ.pop(k, None)
            os.environ.update(env)
            us, mode = time_us(lambda: [launch(kc, vc) for kc, vc in caches], L)
            byts = 2 * B * S * Hkv * D * 2 + 2 * B * H * D * 2
            print(f"attn bs={bs} {tag:12s}: {us:7.1f} us/launch "
                  f"{byts / us / 1e3:7.1f} GB/s ({byts / us / 1e3 / 6576.4:.3f} of measured HBM peak) [{mode}]",
                  flush=True)
        for k in ("B200_ATTN_WARPS", "B200_ATTN_SPLITS", "B200_ATTN_TPS", "B200_ATTN_IMPL"):
            os.environ.pop(k, None)
        ref = None if only_auto else load_reference_kernels()
        if ref is not None and D == 128:
            # GPU baseline beside ours: the reference's own sm80 mma.sync
            # kernel, compiled for sm_100a from /root/reference (oracle/ref/Makefile), same tensors
            def launch_ref(kc, vc):
                ref.paged_kv_varlen_mha(out, q, kc, vc, q_cu, kv_cu, table, blk_cu, None, bs, 1, max_kv,
                                        D ** -0.5, 0.0, -1)
            us, mode = time_us(lambda: [launch_ref(kc, vc) for kc, vc in caches], L)
            byts = 2 * B * S * Hkv * D * 2 + 2 * B * H * D * 2
            print(f"attn bs={bs} {'REFERENCE':12s}: {us:7.1f} us/launch "
                  f"{byts / us / 1e3:7.1f} GB/s ({byts / us / 1e3 / 6576.4:.3f} of measured HBM peak) [{mode}]",
                  flush=True)


def prefill_shape(B, q_len, kv_len, H, Hkv, D, bs=8):
    nblk = (kv_len + bs - 1) // bs
    n_blocks = B * nblk + 8
    L = 4
    caches = [(torch.randn(n_blocks * bs, Hkv, D, device=DEV).bfloat16(),
               torch.randn(n_blocks * bs, Hkv, D, device=DEV).bfloat16()) for _ in range(L)]
    table = (torch.randperm(n_blocks)[: B * nblk] * bs).to(torch.int32).to(DEV)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=DEV)
    q_cu, kv_cu, blk_cu = i32(np.arange(B + 1) * q_len), i32(np.arange(B + 1) * kv_len), i32(np.arange(B + 1) * nblk)
    q = torch.randn(B * q_len, H, D, device=DEV).bfloat16()
    out = torch.empty_like(q)
    # causal: query i of the chunk sees kv_len - q_len + i + 1 keys
    keys = q_len * (kv_len - q_len) + q_len * (q_len + 1) // 2
    flops = 4.0 * B * keys * H * D
    tag = "decode stream kernel" if os.environ.get("B200_ATTN_PREFILL") == "0" else "tcgen05 prefill kernel"

    def run(mod):
        def launch(kc, vc):
            mod.paged_kv_varlen_mha(out, q, kc, vc, q_cu, kv_cu, table, blk_cu, None, bs, q_len, kv_len,
                                    D ** -0.5, 0.0, -1)
        for kc, vc in caches[:2]:
            launch(kc, vc)
        torch.cuda.synchronize()
        us, mode = time_us(lambda: [launch(kc, vc) for kc, vc in caches], L)
        return us, mode
    us, mode = run(kernels)
    print(f"prefill B={B} q={q_len} kv={kv_len} ours ({tag}): {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s [{mode}]", flush=True)
    ref = load_reference_kernels()
    if ref is not None:
        us_r, mode = run(ref)
        print(f"prefill B={B} q={q_len} kv={kv_len} REFERENCE kernel (sm80 mma.sync, built for sm_100a): "
              f"{us_r:8.1f} us  {flops / us_r / 1e6:7.1f} TFLOP/s  -> ours {us_r / us:.2f}x [{mode}]", flush=True)


_REF = []


def load_reference_kernels():
    """oracle/_ref/_ref_kernels.so if it was built (imports lazily: only head_dim 128 of the
    reference attention is instantiated); timing tool only, never the product path."""
    if _REF:
        return _REF[0]
    so_dir = os.path.join(ROOT, "oracle", "_ref")
    mod = None
    if os.path.exists(os.path.join(so_dir, "_ref_kernels.so")):
        sys.path.insert(0, so_dir)
        old = sys.getdlopenflags()
        sys.setdlopenflags(os.RTLD_LAZY | os.RTLD_LOCAL)
        try:
            import _ref_kernels as mod
        except Exception as e:  # noqa: BLE001
            print(f"(reference kernels not loadable: {e})")
            mod = None
        finally:
            sys.setdlopenflags(old)
            sys.path.pop(0)
    _REF.append(mod)
    return mod


if __name__ == "__main__":
    main()
