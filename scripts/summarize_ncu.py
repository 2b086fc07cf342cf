"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-kernel table (markdown).

Usage: python scripts/summarize_ncu.py LAUNCHES.csv STEPS [bench.json] > profiles/rN_ncu_launches_summary.md
STEPS = number of forward steps the profiled process ran (bench.py --steps 2 --warmup 3 runs 3 warm-up + 2 timed + 3 end-to-end
+ 2 instrumented = 10).  Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py, not absolutes."""
import collections
import csv
import json
import re
import sys


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = name.replace("void ", "")
    m = re.match(r"([\w:]+)(<[^(]*>)?", name)
    base = m.group(1) if m else name
    targs = m.group(2) if m and m.group(2) else ""
    targs = re.sub(r"\(bool\)", "", targs)
    return (base + targs)[:90]


def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    bench = json.load(open(sys.argv[3])) if len(sys.argv) > 3 else None
    rows = load(path)
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        k = short(r["Kernel Name"])
        tot[k] += ms
        cnt[k] += 1
    total = sum(tot.values())
    print(f"# ncu launch list summary ({path})\n")
    print(f"{len(rows)} launches over {steps} forward steps; per-step = total / {steps}.  Times under ncu are cold-cache and serialised:")
    print("compare SHARES with bench.py, not absolutes.\n")
    print("| kernel | launches/step | ms/step | share |\n|---|---|---|---|")
    for k, v in tot.most_common(28):
        print(f"| `{k}` | {cnt[k] / steps:.1f} | {v / steps:.3f} | {100 * v / total:.1f} % |")
    conv = sum(v for k, v in tot.items() if re.search(r"conv_tc_kernel|conv_halo_kernel|conv_stem|im2col|stem_pad", k))
    print(f"\nSum {total / steps:.2f} ms/step; convolution kernels (conv_tc + conv_halo + conv_stem_dense + im2col + stem_pad) "
          f"{conv / steps:.2f} ms/step = {100 * conv / total:.1f} % of kernel time.")
    if bench:
        r = bench["roofline"]
        print(f"bench.py on the same build: {bench['ms_per_step']:.2f} ms/step, conv share {r['conv_share_of_step']:.2f} "
              f"({r['conv_ms_per_step']:.2f} ms from CUDA events around every conv launch).")


if __name__ == "__main__":
    main()
