"""Launch the gate_up-shaped W4A16 GEMM a few times (for ncu --set full source-level sampling)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scalellm_b200 import kernels
DEV = "cuda"
K, N, M, g = 4096, 28672, 64, 128
gen = torch.Generator(device=DEV).manual_seed(0)
qw = torch.randint(-2**31, 2**31 - 1, (K, N // 8), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
qz = torch.randint(-2**31, 2**31 - 1, (K // g, N // 8), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
sc = (torch.randn(K // g, N, generator=gen, device=DEV).abs() * 0.01 + 1e-4).bfloat16()
packed = kernels.w4a16_prepack_awq(qw, qz, sc, g)
a = torch.randn(M, K, device=DEV).bfloat16()
out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
for _ in range(6):
    kernels.w4a16_gemm(a, packed, N, g, out=out)
torch.cuda.synchronize()
