"""GPU bring-up probe for the tcgen05 W4A16 GEMM: runs small cases, prints error stats and dumps
(A, W_ref, out) to gpurun_out/w4_probe_*.pt for offline analysis on the CPU box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import quant  # noqa: E402
from scalellm_b200 import kernels  # noqa: E402

DEV = "cuda"


def case(tag, M, K, N, g=128, onehot=False):
    ck = quant.random_awq_checkpoint(K, N, g, seed=1)
    w_ref = quant.dequant(ck["q"], ck["z"], ck["scales"], g)
    if onehot:
        a = torch.zeros(M, K)
        for m in range(M):
            a[m, (m * 37 + 5) % K] = 1.0
        a = a.bfloat16()
    else:
        a = torch.randn(M, K, generator=torch.Generator().manual_seed(0)).bfloat16()
    packed = kernels.w4a16_prepack_awq(ck["qweight"].to(DEV), ck["qzeros"].to(DEV),
                                       ck["scales"].to(DEV), g)
    wd = kernels.w4a16_dequant(packed, K, N, g).cpu()
    out = kernels.w4a16_gemm(a.to(DEV), packed, N, g)
    torch.cuda.synchronize()
    ref = (a.float() @ w_ref.float())
    o = out.float().cpu()
    err = (o - ref).abs()
    print(f"[{tag}] M={M} K={K} N={N} g={g} onehot={onehot} impl={os.environ.get('B200_W4A16_IMPL','tcgen05')}: "
          f"prepack_exact={torch.equal(wd.view(torch.int16), w_ref.view(torch.int16))} "
          f"max_abs_err={err.max():.4e} mean_rel={err.mean() / ref.abs().mean():.3e} "
          f"nan={torch.isnan(o).any().item()} zero_frac={(o == 0).float().mean():.3f}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    torch.save(dict(a=a, w_ref=w_ref, out=out.cpu(), ref=ref),
               os.path.join(ROOT, "gpurun_out", f"w4_probe_{tag}.pt"))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    cases = {
        "t1": dict(M=16, K=128, N=128, onehot=True),
        "t2": dict(M=16, K=128, N=128),
        "t3": dict(M=64, K=256, N=128),
        "t4": dict(M=64, K=1024, N=256),
        "t5": dict(M=64, K=4096, N=4096),
        "t6": dict(M=128, K=512, N=384, g=32),
    }
    for tag, kw in cases.items():
        if which in ("all", tag):
            case(tag, **kw)
