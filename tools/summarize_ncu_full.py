"""Condense an `ncu --page raw --csv` export: one line per launch with the metrics the roofline needs."""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active", "gmma%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, units, body = r, rows[i + 1], rows[i + 2:]
            break
    if hdr is None:
        print("no header found")
        return
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel | " + " | ".join(k for _, k in KEYS))
    for r in body:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0][-40:]
        vals = []
        for full, short in KEYS:
            if full in idx:
                vals.append("%s=%s%s" % (short, r[idx[full]], units[idx[full]] if units[idx[full]] not in ("", "%") else ""))
        stalls = []
        for h, i in idx.items():           # warp-state sampling: warps stalled per issue-active cycle, by reason
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "selected" not in h:
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            vals.append("stalls: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:4]))
        print(name + " | " + " | ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
