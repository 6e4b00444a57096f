"""ncu --csv launch list (gpu__time_duration.sum) -> per-kernel totals, markdown on stdout."""
import csv
import re
import sys
from collections import defaultdict


def main(path, until=None):
    """until: kernel-name substring; only launches up to (and including) its first occurrence are summarised
    (e.g. `guided_step` = the first CFG denoise step of the iteration, i.e. one no-grad UNet forward)."""
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void |leco::", "", name)
        rows.append((name, ns))
    if until:
        for i, (n, _) in enumerate(rows):
            if until in n:
                rows = rows[:i + 1]
                break
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    print(f"launches: {len(rows)}   total device time (serialised, cold cache): {tot / 1e6:.3f} ms\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---|---|---|---|")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.1f} |")


if __name__ == "__main__":
    main(*sys.argv[1:])
