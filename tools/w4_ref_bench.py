"""GPU baseline beside ours: the reference's own Marlin GEMM kernel, compiled
for sm_100a from /root/reference into oracle/_ref (oracle/ref/Makefile), against b200_w4a16_gemm on
the four Llama-3-8B projection shapes at M = 64.  Timing only: the weights are random words in each
library's own layout (any word is a valid packed int4 weight), rotated over several copies so every
launch streams from HBM.  Never on the product path."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from scalellm_b200 import kernels  # noqa: E402
from attn_bench import load_reference_kernels  # noqa: E402

DEV = "cuda"


def timeit(fn, n_copies):
    """us per launch over the rotating copies: >= 100 CUDA-graph replays after 10 warm-ups"""
    from _timing import time_us
    return time_us(lambda: [fn(i) for i in range(n_copies)], n_copies)   # (us, "graph" | "eager")


def main():
    ref = load_reference_kernels()
    M, g, L = 64, 128, 8
    gen = torch.Generator(device=DEV).manual_seed(0)
    ri = lambda *s: torch.randint(-2 ** 31, 2 ** 31 - 1, s, generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    for name, K, N in (("qkv", 4096, 6144), ("o", 4096, 4096), ("gate_up", 4096, 28672), ("down", 14336, 4096)):
        a = torch.randn(M, K, device=DEV).bfloat16()
        sc = (torch.rand(K // g, N, device=DEV) * 0.01 + 1e-3).bfloat16()
        ours_w = [kernels.w4a16_prepack_gptq(ri(K // 8, N), None, sc, g) for _ in range(L)]
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        (t_full, m1), (t_part, m2) = (timeit(lambda i: kernels.w4a16_gemm(a, ours_w[i], N, g, out=out), L),
                                      timeit(lambda i: kernels.w4a16_gemm_splitk(a, ours_w[i], N, g), L))
        line = (f"{name:8s} K={K:5d} N={N:5d}: b200 gemm+reduce {t_full:6.1f} us ({m1}), "
                f"partials only {t_part:6.1f} us ({m2})")
        if ref is not None:
            marlin_w = [ri(K // 16, N * 16 // 8) for _ in range(L)]
            ws = torch.zeros(N // 64 * 16, dtype=torch.int32, device=DEV)
            e = torch.empty(0, dtype=torch.int32, device=DEV)
            t_ref, m3 = timeit(lambda i: ref.marlin_gemm(a, marlin_w[i], out, sc, e, e, e, ws, 4, True, False, True), L)
            line += f", reference Marlin {t_ref:6.1f} us ({m3})"
        print(line, flush=True)


if __name__ == "__main__":
    main()
