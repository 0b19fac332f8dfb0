#!/usr/bin/env python
"""Turn raw ncu output (gpurun_out/, scratch) into the small text summaries kept under profiles/.

  ncu_summary.py launches <launches.csv> <out.md> [title]
      per-kernel count / total / mean / share of a `--metrics gpu__time_duration.sum` launch list
  ncu_summary.py stalls <report.ncu-rep> <out.md> [launch index] [title]
      warp-state (stall reason) shares of one launch and its hottest instructions
  ncu_summary.py full <report.ncu-rep> <out.md> [traffic_key]
      key metrics of every launch in a `--set full` capture; with traffic_key also records
      dram read+write bytes per launch in profiles/traffic.json (read by bench.py)
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__cycles_active.avg",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__grid_size",
    "launch__block_size",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\((?:int|bool|unsigned int)\)", "", name)   # template value casts: (int)128 -> 128
    name = re.sub(r"\(.*$", "", name)
    return name.replace("b200::", "")


def launches(path, out, title):
    rows = []
    with open(path, newline="") as f:
        text = f.read()
    start = text.find('"ID"')
    for r in csv.DictReader(io.StringIO(text[start:])):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") == "us":
                v *= 1e3
            rows.append((short(r["Kernel Name"]), v, r["Grid Size"], r["Block Size"]))
    # drop model construction (weight init, prepack): the first decode step starts with the
    # embedding lookup, the only gather kernel in the run
    first = next((i for i, r in enumerate(rows) if "vectorized_gather_kernel" in r[0]), 0)
    dropped, rows = first, rows[first:]
    agg = OrderedDict()
    for n, v, g, b in rows:
        a = agg.setdefault(n, [0, 0.0, g, b])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for n, a in agg.items() if not n.startswith(("at::", "cutlass", "nvjet", "void at")))
    with open(out, "w") as f:
        f.write(f"# {title}\n\n")
        f.write(f"source: `{os.path.relpath(path, ROOT)}` (ncu --metrics gpu__time_duration.sum "
                f"--clock-control none); {len(rows)} launches from the first decode step on "
                f"({dropped} model-construction launches dropped), {tot / 1e3:.1f} us total.\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | mean us | share | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.2f} | "
                    f"{100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} |\n")
        f.write(f"\nlibb200decode kernels: {100 * ours / tot:.1f}% of the GPU time in the list.\n")
    print(open(out).read())


def full(rep, out, traffic_key):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    start = txt.find('"ID"')
    rd = list(csv.reader(io.StringIO(txt[start:])))
    hdr, units, data = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full: `{os.path.basename(rep)}`\n\n")
        traffic = []
        for r in data:
            f.write(f"## launch {r[col['ID']]}: `{short(r[col['Kernel Name']])}` grid {r[col['Grid Size']]} "
                    f"block {r[col['Block Size']]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in col:
                    f.write(f"| {k} | {r[col[k]]} | {units[col[k]]} |\n")
            try:
                def num(k):
                    v = float(r[col[k]].replace(",", ""))
                    u = units[col[k]].lower()
                    mul = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                    return v * mul
                rdb, wrb = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
                dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
                du = units[col["gpu__time_duration.sum"]].lower()
                dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(du, 1e-9)
                f.write(f"| dram read+write | {(rdb + wrb) / 1e6:.2f} | MB |\n")
                f.write(f"| dram GB/s (under ncu) | {(rdb + wrb) / dur_s / 1e9:.0f} | GB/s |\n")
                traffic.append(rdb + wrb)
            except Exception as e:  # noqa: BLE001
                f.write(f"| (traffic unavailable: {e}) | | |\n")
            f.write("\n")
    print(open(out).read()[:6000])
    if traffic_key and traffic:
        tj = os.path.join(ROOT, "profiles", "traffic.json")
        d = json.load(open(tj)) if os.path.exists(tj) else {}
        d[traffic_key] = {"dram_bytes_per_launch": sum(traffic) / len(traffic),
                          "launches": len(traffic), "source": os.path.basename(out)}
        json.dump(d, open(tj, "w"), indent=1)


def stalls(rep, out, launch_skip="0", title=""):
    """Warp-state sampling of one launch of a `--set full --import-source on` capture: share of
    each stall reason over all sampled warps, and the instructions that collect the most samples."""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                          "--launch-skip", str(launch_skip), "--launch-count", "1"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    name = next((r[1] for r in rows if r and r[0] == "Kernel Name"), "?")
    h = rows[hdr[0]]
    end = hdr[1] - 1 if len(hdr) > 1 else len(rows)
    col = {n: i for i, n in enumerate(h)}
    data = [r for r in rows[hdr[0] + 1:end] if len(r) > 10 and r[0].startswith("0x")]
    names = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    base = int(data[0][0], 16)
    tot = sum(int(r[col["# Samples"]]) for r in data)
    agg = {}
    for r in data:
        for n in names:
            v = int(r[col[n]] or 0)
            if v:
                agg[n] = agg.get(n, 0) + v
    with open(out, "w") as f:
        f.write(f"# {title or 'warp-state sampling'}\n\n`{short(name)}` — launch {launch_skip} of "
                f"`{os.path.basename(rep)}`; {len(data)} SASS instructions, {tot} warp samples.\n"
                "`stall_selected` = the warp issued; `stall_wait` = fixed-latency dependency; "
                "`stall_short_sb` = shared-memory / MUFU / shuffle result; `stall_long_sb` = global "
                "memory, TMA or mbarrier result.\n\n| state | samples | share |\n|---|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            f.write(f"| {k} | {v} | {100 * v / tot:.1f}% |\n")
        f.write("\nInstructions with the most samples:\n\n| offset | SASS | samples | executed | top states |\n|---|---|---:|---:|---|\n")
        for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:20]:
            st = sorted(((n, int(r[col[n]] or 0)) for n in names), key=lambda kv: -kv[1])[:2]
            f.write(f"| {int(r[0], 16) - base:#x} | `{r[1].strip()[:70]}` | {r[col['# Samples']]} | "
                    f"{r[col['Instructions Executed']]} | {', '.join(f'{n} {v}' for n, v in st if v)} |\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "stalls":
        stalls(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "0",
               sys.argv[5] if len(sys.argv) > 5 else "")
    elif mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu launch list")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
