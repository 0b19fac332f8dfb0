"""Device timeline of the graphed decode step (the benchmark configuration), kernel by kernel.

ncu serialises launches and flushes caches, and CUDA events around single launches cannot see
inside a graph replay.  This tool records the replay with the CUPTI activity trace that
torch.profiler drives and reports, per kernel name, the launches, busy time and share of the step,
plus where the step's wall time is NOT covered by any kernel (gaps) and how much of it has two
kernels resident at once (overlap under programmatic dependent launch).

  python tools/step_timeline.py [--layers 32] [--batch 64] [--seqlen 2048] [--out gpurun_out/step_timeline.md]
  torchrun --nproc-per-node N ... tools/step_timeline.py   # tensor parallel: rank 0's timeline

A number printed here is taken under a tracer: use it for shares and gaps, never as a bench value.
"""
import argparse
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\((?:int|bool|unsigned int)\)", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("b200::", "")
    return name if len(name) <= 70 else name[:67] + "..."


def summarise(events, n_steps: int):
    """events: list of (name, start_us, dur_us) of every kernel inside the traced replays, any
    order.  Returns (rows, stats): per-kernel aggregates and the coverage of the traced window."""
    ev = sorted(events, key=lambda e: e[1])
    if not ev:
        return [], {}
    agg = {}
    for name, _, dur in ev:
        a = agg.setdefault(short(name), [0, 0.0])
        a[0] += 1
        a[1] += dur
    t0 = ev[0][1]
    t1 = max(s + d for _, s, d in ev)
    # union of busy intervals -> covered time; sum of durations - union = doubly-covered time
    covered, cur_s, cur_e = 0.0, ev[0][1], ev[0][1] + ev[0][2]
    gaps = []
    for _, s, d in ev[1:]:
        if s > cur_e:
            covered += cur_e - cur_s
            gaps.append(s - cur_e)
            cur_s, cur_e = s, s + d
        else:
            cur_e = max(cur_e, s + d)
    covered += cur_e - cur_s
    busy = sum(d for _, _, d in ev)
    rows = sorted(((n, c, t) for n, (c, t) in agg.items()), key=lambda r: -r[2])
    gaps.sort()
    stats = {"window_us": t1 - t0, "covered_us": covered, "busy_us": busy,
             "overlap_us": busy - covered, "gap_us": (t1 - t0) - covered, "n_gaps": len(gaps),
             "gap_median_us": gaps[len(gaps) // 2] if gaps else 0.0,
             "gap_max_us": gaps[-1] if gaps else 0.0, "kernels": len(ev), "steps": n_steps}
    return rows, stats


def render(rows, st, title: str) -> str:
    n = max(1, st["steps"])
    out = [f"# {title}", "",
           f"{st['kernels']} kernel records over {n} replay(s); per step: window "
           f"{st['window_us'] / n:.1f} us, covered by at least one kernel {st['covered_us'] / n:.1f} us, "
           f"idle between kernels {st['gap_us'] / n:.1f} us in {st['n_gaps'] // n} gaps (median "
           f"{st['gap_median_us']:.2f} us, max {st['gap_max_us']:.2f} us), two kernels resident "
           f"{st['overlap_us'] / n:.1f} us.",
           "Taken under the CUPTI tracer: shares and gaps only, not a bench value.", "",
           "| kernel | launches / step | busy us / step | mean us | share of window |",
           "|---|---:|---:|---:|---:|"]
    for name, c, t in rows:
        out.append(f"| `{name}` | {c / n:.1f} | {t / n:.1f} | {t / c:.2f} | {100 * t / st['window_us']:.1f}% |")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seqlen", type=int, default=2048)
    ap.add_argument("--block-size", type=int, default=8)
    ap.add_argument("--quant", default="awq", choices=["awq", "gptq"])
    ap.add_argument("--replays", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_timeline.md"))
    a = ap.parse_args()

    import torch
    from torch.profiler import ProfilerActivity, profile

    from scalellm_b200.decode_step import (BlockPool, GraphedStep, LlamaArgs, LlamaDecoder, StepBuffers,
                                           build_decode_batch)
    from scalellm_b200.layers import QuantArgs
    from scalellm_b200.model_parallel import ParallelArgs

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        import torch.distributed as dist
        from scalellm_b200.model_parallel import ProcessGroup
        dist.init_process_group("nccl", device_id=dev)
        pg = ProcessGroup(rank, world, dev)
    args = LlamaArgs.llama3_8b()
    args.n_layers = a.layers
    qa = QuantArgs(quant_method=a.quant, bits=4, group_size=128, is_sym=(a.quant == "gptq"))
    model = LlamaDecoder(args, qa, ParallelArgs(rank, world, pg), dev)
    model.init_random(seed=0)
    B, S, bs = a.batch, a.seqlen, a.block_size
    cap = S + 16
    blocks_per_seq = (cap + bs - 1) // bs
    n_blocks = B * blocks_per_seq + 16
    pool = BlockPool(n_blocks, bs, seed=2)
    for _ in range(B):
        pool.add_sequence(cap)
    model.alloc_kv(n_blocks, bs, randomize=True, seed=1)
    bufs = StepBuffers(dev, B, B, B * blocks_per_seq)
    hb = build_decode_batch(pool, [S] * B, [1] * B, args.vocab_size)
    hb.kv_max = cap
    step = GraphedStep(model, bufs, hb, greedy=True)
    for _ in range(3):
        step.replay()
    torch.cuda.synchronize()

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.replays):
            step.replay()
        torch.cuda.synchronize()
    events = []
    for e in prof.events():
        # kernel records only (device side); memcpy/memset records carry these names
        if getattr(e, "device_type", None) is None or "cuda" not in str(e.device_type).lower():
            continue
        if e.name.startswith(("Memcpy", "Memset")):
            continue
        tr = e.time_range
        events.append((e.name, float(tr.start), float(tr.end - tr.start)))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        if rank != 0:
            os._exit(0)     # communicators captured into graphs can block at tear-down
    rows, st = summarise(events, a.replays)
    # one layer in the middle of the last replay, kernel by kernel: start relative to the layer's
    # first kernel, duration, and the gap / overlap against the previous kernel's end
    seq = sorted(events, key=lambda e: e[1])
    per_step = len(seq) // max(1, a.replays)
    last = seq[-per_step:]
    attn_idx = [i for i, e in enumerate(last) if "paged_attn_persist" in e[0] or "paged_attn_decode" in e[0]]
    layer_txt = ""
    if len(attn_idx) > 4:
        mid = attn_idx[len(attn_idx) // 2]
        nxt = attn_idx[len(attn_idx) // 2 + 1]
        lo = mid - (nxt - mid) + 1 if mid - (nxt - mid) + 1 > 0 else 0
        win = last[lo: nxt + 1]
        t0 = win[0][1]
        lines = ["", "## one layer (middle of the last replay), in start order", "",
                 "| kernel | start us | dur us | end us | start - prev end us |", "|---|---:|---:|---:|---:|"]
        prev_end = None
        for name, s0, d in win:
            gap = "" if prev_end is None else f"{s0 - prev_end:+.2f}"
            lines.append(f"| `{short(name)}` | {s0 - t0:.2f} | {d:.2f} | {s0 - t0 + d:.2f} | {gap} |")
            prev_end = max(prev_end or 0.0, s0 + d)
        layer_txt = "\n".join(lines) + "\n"
    if not rows:
        print("no device kernel records in the trace (CUPTI unavailable?)")
        sys.exit(1)
    text = render(rows, st, f"decode-step timeline: Llama-3-8B {a.quant} int4, batch {B}, kv_len {S}, "
                            f"block_size {bs}, {a.layers} layers, TP={world}, CUDA graph replay")
    text += layer_txt
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        f.write(text)
    print(text, flush=True)
    if world > 1:
        os._exit(0)


if __name__ == "__main__":
    main()
