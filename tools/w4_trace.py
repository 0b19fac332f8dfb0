"""Timeline probe for the W4A16 GEMM: per-CTA clock64 milestones (see W4_TRACE in w4a16.cu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scalellm_b200 import _lib, kernels  # noqa: E402

DEV = "cuda"
NAMES = ["start", "setup_done", "deq_first_raw", "deq_g0_done", "mma_first", "mma_last_issued",
         "epi_first_full", "epi_done", "end", "SUM mma wait act_full", "SUM mma wait deq_full",
         "SUM mma wait acc_empty", "SUM deq(g0) wait raw_full", "SUM deq(g0) wait slot"]


def run(K, N, M=64, g=128):
    gen = torch.Generator(device=DEV).manual_seed(0)
    qw = torch.randint(-2**31, 2**31 - 1, (K, N // 8), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    qz = torch.randint(-2**31, 2**31 - 1, (K // g, N // 8), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
    sc = (torch.randn(K // g, N, generator=gen, device=DEV).abs() * 0.01 + 1e-4).bfloat16()
    packed = kernels.w4a16_prepack_awq(qw, qz, sc, g)
    a = torch.randn(M, K, device=DEV).bfloat16()
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    for _ in range(3):
        kernels.w4a16_gemm_splitk(a, packed, N, g)
    torch.cuda.synchronize()
    # flush L2 so the traced launch streams weights from HBM
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    flush.zero_()
    trace = torch.zeros(1024 * 16, dtype=torch.int64, device=DEV)
    lib = _lib.load()
    lib.b200_debug_set_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kernels.w4a16_gemm_splitk(a, packed, N, g)   # the GEMM launch alone (partials out)
    e1.record()
    torch.cuda.synchronize()
    lib.b200_debug_set_trace(None)
    t = trace.cpu().view(-1, 16)
    t = t[t[:, 0] != 0]
    rel = (t - t[:, :1]).float()
    rel[:, 9:] = t[:, 9:].float()        # wait totals are already durations
    ns0, ns1 = t[:, 14], t[:, 15]        # %globaltimer at CTA entry / exit
    rel[t == 0] = float("nan")
    print(f"== K={K} N={N} M={M}: event time {e0.elapsed_time(e1)*1e3:.1f} us, {t.shape[0]} CTAs "
          f"(cycles; ~1.9 cycles/ns)")
    if int(ns0.min()) > 0:
        print(f"  grid span {(ns1.max() - ns0.min()).item() / 1e3:.2f} us (first CTA entry -> last CTA exit); "
              f"CTA entries spread over {(ns0.max() - ns0.min()).item() / 1e3:.2f} us; "
              f"median CTA lifetime {(ns1 - ns0).float().median().item() / 1e3:.2f} us")
    for i, nm in enumerate(NAMES):
        col = rel[:, i]
        col = col[~torch.isnan(col)]
        if col.numel() == 0 or nm == "-":
            continue
        print(f"  {nm:20s} n={col.numel():4d} median={col.median():9.0f} min={col.min():9.0f} "
              f"max={col.max():9.0f}")
    torch.save(t, os.path.join(ROOT, "gpurun_out", f"w4_trace_{K}x{N}.pt"))


def back_to_back(K, N, M=64, g=128, n_launch=6):
    """%globaltimer view of consecutive GEMM launches over distinct weights (as in a decode step):
    how long each grid lives, and the dead time between one grid's last CTA exit and the next
    grid's first CTA entry (negative = the next grid started early under programmatic launch)."""
    gen = torch.Generator(device=DEV).manual_seed(1)
    packs = []
    for _ in range(n_launch):
        qw = torch.randint(-2**31, 2**31 - 1, (K, N // 8), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
        qz = torch.randint(-2**31, 2**31 - 1, (K // g, N // 8), generator=gen, device=DEV, dtype=torch.int64).to(torch.int32)
        sc = (torch.randn(K // g, N, generator=gen, device=DEV).abs() * 0.01 + 1e-4).bfloat16()
        packs.append(kernels.w4a16_prepack_awq(qw, qz, sc, g))
    a = torch.randn(M, K, device=DEV).bfloat16()
    for pk in packs[:2]:
        kernels.w4a16_gemm_splitk(a, pk, N, g)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    flush.zero_()
    traces = [torch.zeros(1024 * 16, dtype=torch.int64, device=DEV) for _ in packs]
    lib = _lib.load()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for pk, tr in zip(packs, traces):
        lib.b200_debug_set_trace(tr.data_ptr())
        kernels.w4a16_gemm_splitk(a, pk, N, g)
    e1.record()
    torch.cuda.synchronize()
    lib.b200_debug_set_trace(None)
    ent, ext = [], []
    for tr in traces:
        t = tr.cpu().view(-1, 16)
        t = t[t[:, 14] != 0]
        ent.append((int(t[:, 14].min()), int(t[:, 14].max())))
        ext.append((int(t[:, 15].min()), int(t[:, 15].max())))
    print(f"== back to back, K={K} N={N} M={M}, B200_PDL={os.environ.get('B200_PDL', 'default')}: "
          f"{n_launch} launches in {e0.elapsed_time(e1) * 1e3:.1f} us (eager, traced kernel)")
    for i in range(n_launch):
        gap = (ent[i][0] - ext[i - 1][1]) / 1e3 if i else float("nan")
        print(f"  launch {i}: grid span {(ext[i][1] - ent[i][0]) / 1e3:6.2f} us, entries spread "
              f"{(ent[i][1] - ent[i][0]) / 1e3:5.2f} us, gap after previous grid {gap:6.2f} us")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for K, N in [(4096, 4096), (4096, 6144), (4096, 28672), (14336, 4096)]:
        run(K, N)
    for K, N in [(4096, 4096), (4096, 28672)]:
        back_to_back(K, N)
