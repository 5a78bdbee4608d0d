"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        rows.append((name, ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0.0, 0])
    for n, ns in rows:
        agg[n][0] += ns
        agg[n][1] += 1
    lines = ["%d launches, total device time %.3f ms (ncu: serialised, cold-cache; compare SHARES)" % (len(rows), tot / 1e6)]
    for n, (ns, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        lines.append("%8.3f ms %5.1f%%  x%-4d avg %8.1f us  %s" % (ns / 1e6, 100 * ns / tot, c, ns / c / 1e3, n[:110]))
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
