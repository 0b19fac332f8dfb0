"""GEMM-only A/B of the W4A16 kernel variants (B200_W4_VARIANT, read once per process): the four
Llama-3-8B projection shapes at M = 64, launches replayed from a CUDA graph, weights rotated over
L copies so every launch streams from HBM.  One JSON line: {"variant": v, "us": {...}, "sum_us": s}.
Quick (seconds): tools/gpu_ci.sh runs it per variant and the full bench only for the fastest."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from scalellm_b200 import kernels  # noqa: E402
from _timing import time_us  # noqa: E402

DEV = "cuda"


def main():
    M, g, L = 64, 128, 16
    gen = torch.Generator(device=DEV).manual_seed(0)
    ri = lambda *s: torch.randint(-2 ** 31, 2 ** 31 - 1, s, generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    res = {}
    for name, K, N in (("qkv", 4096, 6144), ("o", 4096, 4096), ("gate_up", 4096, 28672), ("down", 14336, 4096)):
        a = torch.randn(M, K, device=DEV).bfloat16()
        sc = (torch.rand(K // g, N, device=DEV) * 0.01 + 1e-3).bfloat16()
        ws = [kernels.w4a16_prepack_gptq(ri(K // 8, N), None, sc, g) for _ in range(L)]

        def sweep():
            for w in ws:
                kernels.w4a16_gemm_splitk(a, w, N, g)
        res[name], mode = time_us(sweep, L, replays=20, warmup=3)
        assert mode == "graph", "GEMM launches could not be captured"
        del ws
    print(json.dumps({"variant": os.environ.get("B200_W4_VARIANT", "0"), "us": res,
                      "sum_us": sum(res.values())}), flush=True)


if __name__ == "__main__":
    main()
