"""Condense an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into per-kernel totals -> profiles/*.csv

    python tools/launch_summary.py gpurun_out/launches.csv profiles/launches_summary.csv [forwards] ["comment"]

`forwards` = how many identical forwards the command ran; the LAST one is summarised (warm caches, no set-up kernels).
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import collections
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    forwards = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    comment = sys.argv[4] if len(sys.argv) > 4 else ""
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    ls = [(r[ki], float(r[vi].replace(",", ""))) for r in data if len(r) > vi and r[mi] == "gpu__time_duration.sum"]
    # forwards are separated by the marker kernel tools/model_once.py launches (a float64 fill); keep the last one
    marks = [i for i, (k, _) in enumerate(ls) if "FillFunctor<double>" in k]
    if marks:
        ls = ls[marks[-1] + 1:]
    elif forwards > 1:
        ls = ls[len(ls) - len(ls) // forwards:]
    agg = collections.OrderedDict()
    for k, v in ls:
        k = re.sub(r"\(.*", "", k)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        if comment:
            f.write("# %s\n" % comment)
        f.write("# kernel,launches,total_us,share_pct   (one forward; ncu per-launch times are cold-cache and serialised: compare shares)\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.2f\n" % (k.replace(",", ";"), c, v / 1e3, 100 * v / tot))
        f.write("TOTAL,%d,%.1f,100.00\n" % (sum(v[0] for v in agg.values()), tot / 1e3))


if __name__ == "__main__":
    main()
