"""Latency of the tensor-parallel collectives in isolation (torchrun --nproc-per-node N tools/ar_bench.py).

Every rank issues the same sequence back to back, so there is no inter-rank skew to absorb: the
per-call time is the collective's own latency.  Device-timed with CUDA events, max over ranks.
  * plain all-reduce of [64, 4096] bf16 (512 KiB, the decode step's message): ours (B200_AR_ALGO
    as set) vs NCCL
  * the fused form the decoder uses: GEMM partials -> all-reduce -> residual add -> RMSNorm
  * all-gather of [64, 512] bf16 along the last dim (embedding), sharded argmax of [64, vocab / N]
and, with the debug trace on (b200_debug_set_trace), where a fused call's time goes per phase
(%globaltimer stamps of every block: entry -> contribution pushed -> all contributions here ->
reduced row pushed / received -> done).  Timing only; correctness lives in tests/test_gpu_allreduce.py.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scalellm_b200 import _lib, kernels  # noqa: E402
from scalellm_b200.model_parallel import ProcessGroup  # noqa: E402


def timed(fn, iters=200, warm=20):
    """us per call: `inner` calls captured into one CUDA graph, replayed; device-timed, max over
    ranks.  (Issued eagerly from Python these 5-20 us kernels measure the host's launch rate.)"""
    inner = 20
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    reps = max(1, iters // inner)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / (reps * inner)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    pg = ProcessGroup(rank, world, dev)
    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)
    algo = os.environ.get("B200_AR_ALGO", "") or "twoshot"
    say(f"== world {world}, all-reduce algorithm {algo}")
    for rows, n in ((64, 4096), (32, 8192), (8, 4096)):
        x = torch.randn(rows, n, device=dev).bfloat16()
        y = x.clone()
        t_ours = timed(lambda: pg.allreduce(y))
        t_nccl = timed(lambda: dist.all_reduce(y))
        say(f"all-reduce [{rows}, {n}] bf16 ({rows * n * 2 >> 10} KiB): ours {t_ours:6.2f} us   NCCL {t_nccl:6.2f} us")
    # the fused form: stream-K partials of a row-parallel GEMM -> reduce -> +residual -> RMSNorm
    M, K, N = 64, 4096 // world if world <= 8 else 512, 4096
    K = max(128, K // 128 * 128)
    rng = np.random.default_rng(1 + rank)
    qw = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=(K, N // 8), dtype=np.int64).astype(np.int32)).to(dev)
    qz = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=(K // 128, N // 8), dtype=np.int64).astype(np.int32)).to(dev)
    sc = (torch.rand(K // 128, N) * 0.01 + 1e-3).bfloat16().to(dev)
    packed = kernels.w4a16_prepack_awq(qw, qz, sc, 128)
    a = torch.randn(M, K, device=dev).bfloat16()
    parts = kernels.w4a16_gemm_splitk(a, packed, N, 128)
    res = torch.randn(M, N, device=dev).bfloat16()
    wn = torch.ones(N, device=dev).bfloat16()
    if pg.supports_partials_norm(M, N, torch.bfloat16):
        t = timed(lambda: pg.allreduce_partials_norm(parts, res, wn, 1e-5))
        say(f"fused partials({parts.data.shape[0]} slots) -> all-reduce -> +residual -> RMSNorm [{M}, {N}]: {t:6.2f} us")
        t1 = timed(lambda: kernels.rms_norm_residual_splitk(torch.empty_like(res), res, parts, wn, 1e-5))
        say(f"   the same consumer without the exchange (1-GPU kernel):                 {t1:6.2f} us")
        # per-phase trace of ONE fused call (all ranks in lockstep after a barrier)
        trace = torch.zeros(256 * 8, dtype=torch.int64, device=dev)
        lib = _lib.load()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        lib.b200_debug_set_trace(trace.data_ptr())
        pg.allreduce_partials_norm(parts, res, wn, 1e-5)
        torch.cuda.synchronize()
        lib.b200_debug_set_trace(None)
        t = trace.cpu().view(-1, 8)[:M].double()
        t0 = t[:, 0].min()
        names = {0: "entry (after griddepcontrol.wait)", 1: "contribution stored + flag sent",
                 2: "one-shot: all contributions arrived",
                 4: "reduced row complete (own chunk reduced + pushed, the others arrived)",
                 5: "row done (residual + norm written)"}
        for r in range(world):
            if rank == r:
                print(f"   rank {r} phases, us since the first block's entry (median / max over the {M} blocks):", flush=True)
                for slot, nm in names.items():
                    col = t[:, slot]
                    col = col[col > 0]
                    if col.numel():
                        print(f"     {nm:42s} {float((col - t0).median()) / 1e3:7.2f} / {float((col - t0).max()) / 1e3:7.2f}  (n={col.numel()})", flush=True)
            dist.barrier()
    x = torch.randn(64, 4096 // world, device=dev).bfloat16()
    t = timed(lambda: pg.allgather_lastdim(x))
    outs = [torch.empty_like(x) for _ in range(world)]
    t2 = timed(lambda: torch.cat((dist.all_gather(outs, x), outs)[1], dim=-1))
    say(f"all-gather [64, {4096 // world}] -> [64, 4096]: ours {t:6.2f} us   NCCL + cat {t2:6.2f} us")
    V = 128256 // world
    lg = torch.randn(64, V, device=dev).bfloat16()
    t = timed(lambda: pg.argmax_sharded(lg))
    big = [torch.empty_like(lg) for _ in range(world)]
    t2 = timed(lambda: kernels.argmax(torch.cat((dist.all_gather(big, lg), big)[1], dim=-1)), iters=50)
    say(f"greedy over sharded logits [64, {V}] x {world}: ours {t:6.2f} us   NCCL gather + cat + argmax {t2:6.2f} us")
    dist.barrier()
    torch.cuda.synchronize()
    pg.close()
    os._exit(0)


if __name__ == "__main__":
    main()
