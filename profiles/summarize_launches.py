"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/r1_launches.md
(ncu times are cold-cache and serialised: compare SHARES, not absolutes.)
"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v * 1e6 if unit == "s" else v
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("fsmg::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    print(f"launch list `{path}`: {sum(a[0] for a in agg.values())} launches, {total / 1e3:.2f} ms of kernel time\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---|---|---|---|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {us:.1f} | {100 * us / total:.1f}% | {us / n:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
