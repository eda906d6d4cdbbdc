"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (markdown)."""
import collections
import csv
import re
import sys


def main(path, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[mu], 1.0)
        short = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[kn]))
        agg[short][0] += 1
        agg[short][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}\n")
    print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e6:.2f} ms of kernel time (cold-cache, serialised under ncu: compare shares)\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k[:120]}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
