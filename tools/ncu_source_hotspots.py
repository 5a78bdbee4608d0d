"""Per-instruction hot spots of one kernel from `ncu -i <rep> --page source --csv --launch-skip K --launch-count 1`:
stall-reason totals, the most-sampled SASS instructions, the executed-instruction histogram (which loop a row belongs to shows
in its execution count) and the opcode mix of the hottest loop body. Usage: python tools/ncu_source_hotspots.py src.csv [top]"""
import csv
import re
import sys
from collections import Counter


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, body, name = None, [], ""
    for r in rows:
        if r and r[0] == "Kernel Name":
            if hdr is not None:
                break
            name = r[1][:90]
            continue
        if hdr is None and r and r[0] == "Address":
            hdr = r
            continue
        if hdr is not None and r:
            body.append(r)
    return name, {n: i for i, n in enumerate(hdr)}, body


def main(path, top=16):
    name, h, body = load(path)
    si, ie, src = h["# Samples"], h["Instructions Executed"], h["Source"]
    stall_cols = [(n, i) for n, i in h.items() if n.startswith("stall_") and "Not Issued" not in n]
    tot_s = sum(int(r[si] or 0) for r in body)
    tot_i = sum(int(r[ie] or 0) for r in body)
    print("kernel %s\n%d SASS rows, %d warp-level instructions executed, %d stall samples" % (name, len(body), tot_i, tot_s))
    agg = {n: sum(int(r[i] or 0) for r in body) for n, i in stall_cols}
    print("stall samples by reason: " + ", ".join("%s %d" % (n[6:], v) for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
    print("-- most-sampled instructions")
    for idx, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][si] or 0))[:top]:
        st = sorted(((int(r[i] or 0), n[6:]) for n, i in stall_cols), reverse=True)[:2]
        print("  %5d samples %4.1f%%  row %5d  %-62s %s" % (int(r[si] or 0), 100.0 * int(r[si] or 0) / max(tot_s, 1), idx, r[src][:62], st))
    print("-- executed-instruction histogram (execution count x rows = share of all executed instructions)")
    c = Counter(int(r[ie] or 0) for r in body)
    hot = None
    for k, v in sorted(c.items(), key=lambda kv: -kv[0] * kv[1])[:8]:
        if hot is None:
            hot = k
        print("  executed %9d times x %4d rows = %5.1f%%" % (k, v, 100.0 * k * v / max(tot_i, 1)))
    ops = Counter()
    for r in body:
        if int(r[ie] or 0) == hot:
            ops[re.sub(r"^@!?U?P\d\s+", "", r[src].strip()).split()[0].split(".")[0]] += 1
    print("-- opcode mix of the hottest loop body (%d rows): %s" % (sum(ops.values()), ", ".join("%s %d" % kv for kv in ops.most_common(18))))
    spin = [(i, r) for i, r in enumerate(body) if "TRYWAIT" in r[src] and int(r[ie] or 0) > hot]
    for i, r in spin:
        print("-- spin loop at row %d: %s executed %s times (single thread per warp)" % (i, r[src].strip()[:60], r[ie]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)
