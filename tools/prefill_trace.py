"""Where a prefill-attention CTA's time goes (csrc/prefill_attn.cu, debug trace): clock64 totals per
role — softmax thread of row 0 (wait for S, pass 1, wait for PV + fold, pass 2), MMA issuer (waits for
K/V, for P, issue), producer (waits for a free stage, table lookup + TMA issue) — per tile, median
over the CTAs of one launch.  Debug tool; not a bench value."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scalellm_b200 import _lib, kernels  # noqa: E402

DEV = "cuda"
NAMES = {0: "CTA lifetime", 2: "softmax: wait S", 3: "softmax: pass 1 (max)", 4: "softmax: wait PV + fold O",
         5: "softmax: pass 2 (exp, P -> smem)", 6: "mma: wait K/V", 7: "mma: wait P", 11: "mma: issue + commit",
         10: "mma: wait Q", 8: "producer: wait free stage", 9: "producer: table + TMA issue"}


def one(B, q_len, kv_len, H=32, Hkv=8, D=128, bs=8):
    nblk = (kv_len + bs - 1) // bs
    n_blocks = B * nblk + 8
    kc = torch.randn(n_blocks * bs, Hkv, D, device=DEV).bfloat16()
    vc = torch.randn(n_blocks * bs, Hkv, D, device=DEV).bfloat16()
    table = (torch.randperm(n_blocks)[: B * nblk] * bs).to(torch.int32).to(DEV)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=DEV)
    q_cu, kv_cu, blk_cu = i32(np.arange(B + 1) * q_len), i32(np.arange(B + 1) * kv_len), i32(np.arange(B + 1) * nblk)
    q = torch.randn(B * q_len, H, D, device=DEV).bfloat16()
    out = torch.empty_like(q)

    def launch():
        kernels.paged_kv_varlen_mha(out, q, kc, vc, q_cu, kv_cu, table, blk_cu, None, bs, q_len, kv_len,
                                    D ** -0.5, 0.0, -1)
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    trace = torch.zeros(1024 * 16, dtype=torch.int64, device=DEV)
    lib = _lib.load()
    lib.b200_debug_set_trace(trace.data_ptr())
    launch()
    torch.cuda.synchronize()
    lib.b200_debug_set_trace(None)
    t = trace.cpu().view(-1, 16).double()
    t = t[t[:, 1] > 0]
    tiles = t[:, 1]
    print(f"== B={B} q={q_len} kv={kv_len} bs={bs}: {t.shape[0]} CTAs traced, tiles per CTA median {tiles.median():.0f} "
          f"(min {tiles.min():.0f}, max {tiles.max():.0f}); cycles (~1.9 per ns)")
    for k, nm in NAMES.items():
        col = t[:, k]
        per = col / tiles
        print(f"   {nm:36s} total median {col.median():10.0f}   per tile median {per.median():8.0f}  (max {per.max():8.0f})")


if __name__ == "__main__":
    one(1, 128, 2048)
    one(4, 2048, 2048)
    one(1, 128, 2048, bs=128)
