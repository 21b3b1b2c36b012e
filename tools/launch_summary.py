"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares,
split by phase (ViT+projector / prefill / decode).  Usage: python tools_launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    order = []
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        if row["Metric Unit"] == "ns":
            t /= 1e3
        elif row["Metric Unit"] == "ms":
            t *= 1e3
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        order.append((name, t, row.get("Grid Size", "")))
    return order


def table(rows, label, top=12):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, _ in rows:
        agg[n][0] += 1
        agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"\n== {label}: {tot / 1e3:.2f} ms over {len(rows)} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {v[1] / tot * 100:6.2f}%  {v[1] / 1e3:9.3f} ms  n={v[0]:6d}  avg={v[1] / v[0]:9.2f} us  {k}")


if __name__ == "__main__":
    order = load(sys.argv[1])
    table(order, "whole step", 16)
    i_sp = next(i for i, o in enumerate(order) if "splice_embed" in o[0])
    i_dec = next(i for i, o in enumerate(order) if "decode_attn" in o[0])
    i_dec0 = max(i for i in range(i_dec) if "splice_embed" in order[i][0])   # embed gather that opens the first decode step
    table(order[:i_sp], "ViT + projector")
    table(order[i_sp:i_dec0], "prefill (+ first argmax)")
    table(order[i_dec0:], "decode")
    steps = sum(1 for o in order[i_dec0:] if "argmax" in o[0])
    dec = sum(t for _, t, _ in order[i_dec0:])
    print(f"\ndecode: {steps} steps, {dec / steps / 1e3:.3f} ms per step (serialised, cold-cache)")
