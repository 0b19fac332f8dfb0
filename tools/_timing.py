"""Timing helper shared by the kernel-only benchmarks (CUDA events over >= 100 graph
replays after 10 warm-ups): `sweep()` launches the kernel once per rotating input copy; the sweep is
captured into a CUDA graph so the host's launch rate is not part of the number.  Falls back to
eager launches (and says so) if the kernel cannot be captured."""
import torch


def time_us(sweep, n_launches: int, replays: int = 100, warmup: int = 10):
    """-> (microseconds per launch, "graph" | "eager")"""
    sweep()
    torch.cuda.synchronize()
    mode, run = "eager", sweep
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            sweep()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            sweep()
        mode, run = "graph", g.replay
    except Exception:  # noqa: BLE001  (e.g. a library call that synchronises)
        torch.cuda.synchronize()
    for _ in range(warmup):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (replays * n_launches), mode
