"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_bench_launches.md
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) == len(hdr):
            rows.append(dict(zip(hdr, r)))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"<.*", "", name) if not name.startswith(("void smt", "smt", "void unnamed")) and "smt" not in name else re.sub(r"\(.*", "", r["Kernel Name"])[:110]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        agg[name][0] += 1
        agg[name][1] += us
    total = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# Launch list summary ({n} launches, {total / 1e3:.2f} ms of kernel time; ncu per-launch times are cold-cache and "
          "serialised: compare SHARES)\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    ours = 0.0
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        tag = " **(ours)**" if ("smt" in name or "block_grad" in name or "compact_adam" in name or "sqnorm" in name or "splitk" in name or "block_s" in name or "block_copy" in name or "topk" in name) else ""
        if tag:
            ours += t
        print(f"| `{name[:100]}`{tag} | {c} | {t:.1f} | {100 * t / total:.2f}% | {t / c:.2f} |")
    print(f"\nSMT kernels (ours) among the top 40: {ours:.1f} us = {100 * ours / total:.2f}% of kernel time")


if __name__ == "__main__":
    main(sys.argv[1])
