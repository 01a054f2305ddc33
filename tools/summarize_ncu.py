#!/usr/bin/env python
"""Turns .ncu-rep captures (gpurun_out/) into small tracked summaries under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dsmem")]


def summarize(rep, labels=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = ["| # | kernel | " + " | ".join(k for _, k in KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    ki = hdr.index("Kernel Name")
    for n, r in enumerate(rows[2:]):
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        if labels and n < len(labels):
            name += " — " + labels[n]
        vals = []
        for key, _ in KEYS:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = "%.3g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                vals.append("%s %s" % (v, units[i]) if units[i] not in ("", "%") else v)
            else:
                vals.append("-")
        lines.append("| %d | %s | %s |" % (n, name, " | ".join(vals)))
    return "\n".join(lines)


if __name__ == "__main__":
    rep = sys.argv[1]
    labels = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    print(summarize(rep, labels))
