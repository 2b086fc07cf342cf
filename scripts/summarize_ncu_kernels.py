"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (any workload).
Usage: python scripts/summarize_ncu_kernels.py LAUNCHES.csv [top_n]"""
import collections
import csv
import re
import sys


def short(name):
    name = name.replace("void ", "")
    m = re.match(r"([\w:]+)(<[^(]*>)?", name)
    return ((m.group(1) + (m.group(2) or "")) if m else name)[:110]


def main():
    rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
        k = short(r["Kernel Name"])
        tot[k] += ms
        cnt[k] += 1
    total = sum(tot.values())
    print(f"{len(rows)} launches, {total:.2f} ms of kernel time (cold-cache, serialised: compare shares, not absolutes)\n")
    print("| kernel | launches | ms | share | avg us |\n|---|---|---|---|---|")
    for k, v in tot.most_common(top):
        print(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / total:.1f} % | {1000 * v / cnt[k]:.1f} |")


if __name__ == "__main__":
    main()
