"""Digest of `ncu --page source --csv --print-source sass` (per-instruction warp-stall samples) into a small markdown file:
per kernel the stall-reason totals and the instructions that collect the most samples.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > rep.source.csv
    python profiles/stall_summary.py rep.source.csv [top_n] > profiles/rN_stalls_summary.md
"""
import csv
import sys

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], hdr=None, ins=[])
        kernels.append(cur)
    elif r and r[0] == "Address":
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:                       # second view of the same kernel (repeated table): ignore
            cur = dict(name=cur["name"] + " (repeat)", hdr=r, ins=[])
    elif cur is not None and r:
        cur["ins"].append(r)
seen = set()
for k in kernels:
    h = k["hdr"]
    if not h or not k["ins"]:
        continue
    key = (k["name"].replace(" (repeat)", ""), len(k["ins"]), k["ins"][0][0])
    if key in seen:                 # the CSV repeats every kernel's table
        continue
    seen.add(key)
    ci = {n: i for i, n in enumerate(h)}
    if "# Samples" not in ci:
        continue
    tot = sum(int(r[ci["# Samples"]] or 0) for r in k["ins"])
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = sorted(((s, sum(int(r[ci[s]] or 0) for r in k["ins"])) for s in stalls), key=lambda x: -x[1])
    print(f"### `{k['name'][:120]}`\n")
    print(f"{len(k['ins'])} SASS instructions, {tot} warp-stall samples; by reason: "
          + ", ".join(f"{s[6:]} {100.0 * v / max(tot, 1):.1f}%" for s, v in agg[:8]) + "\n")
    print("| # | instruction | samples | share | executed (warp) | top stall reasons |")
    print("|---|---|---|---|---|---|")
    order = sorted(range(len(k["ins"])), key=lambda i: -int(k["ins"][i][ci["# Samples"]] or 0))[:top_n]
    for i in sorted(order):
        r = k["ins"][i]
        st = sorted(((s[6:], int(r[ci[s]] or 0)) for s in stalls), key=lambda x: -x[1])[:2]
        n = int(r[ci["# Samples"]] or 0)
        print(f"| {i} | `{r[1].strip()[:70]}` | {n} | {100.0 * n / max(tot, 1):.1f}% | {r[ci['Instructions Executed']]} | "
              + ", ".join(f"{a} {b}" for a, b in st) + " |")
    print()
