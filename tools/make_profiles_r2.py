#!/usr/bin/env python
"""gpurun_out/ (scratch) -> profiles/ (tracked), round 2: launch list of one detect step (r02_launches.csv) and the raw-page
exports of the ncu --set full captures (prof_r2_misc.raw.csv: ProposalLayer / row-per-warp ROIAlign / DetectionLayer / unmold;
prof_r2_train.raw.csv: train-mode kernels) as small tables."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def read_long(path):
    """ncu --csv --log-file long format -> [{"name":..., "grid":..., "block":..., metric: value}] in launch order"""
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0] != "ID" and r[0].isdigit()]
    out = {}
    for r in rows:
        d = out.setdefault(int(r[0]), {"name": r[4].replace("<unnamed>::", "").replace("void ", "").split("(")[0], "grid": r[8], "block": r[7]})
        d[r[12]] = float(r[14].replace(",", ""))
    return [out[k] for k in sorted(out)]


def short(name):
    return name.replace("CUtensorMap_st", "tmap")


KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dsmem")]


def table_from_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    lines = ["| # | kernel | " + " | ".join(k for _, k in KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for n, r in enumerate(rows[2:]):
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
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


def main():
    ll = read_long(os.path.join(OUT, "r02_launches.csv"))
    tot = sum(d["gpu__time_duration.sum"] for d in ll)
    with open(os.path.join(PROF, "r02_launches_one_step.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: every launch of one detect_maps step (B=64, S=256), round 2\n")
        f.write("# cold-cache, serialised times: compare shares, not absolutes. total %.3f ms over %d launches\n" % (tot / 1e6, len(ll)))
        f.write("idx,kernel,grid,block,time_us,share_pct\n")
        for i, d in enumerate(ll):
            f.write("%d,%s,\"%s\",\"%s\",%.2f,%.2f\n" % (i, short(d["name"]), d["grid"], d["block"], d["gpu__time_duration.sum"] / 1e3,
                                                          100 * d["gpu__time_duration.sum"] / tot))
    fam = {}
    for d in ll:
        k = d["name"].split("<")[0]
        fam[k] = fam.get(k, 0.0) + d["gpu__time_duration.sum"]
    shares = {k: "%.1f%%" % (100 * v / tot) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}
    print("kernel shares of the step under ncu:", shares)
    with open(os.path.join(PROF, "r02_ncu_summary.md"), "w") as f:
        f.write("# ncu summaries, round 2 (one B200, `tools/gpu_runs/prof_r2.sh`; captures under a profiler are never bench values)\n\n")
        f.write("Kernel-family shares of one detect step (launch list `r02_launches_one_step.csv`, %d launches, %.2f ms serialised): %s\n\n"
                % (len(ll), tot / 1e6, ", ".join("%s %s" % kv for kv in shares.items())))
        for title, name in (("Detect path, kernels that changed in round 2 (`ncu --set full`, second step)", "prof_r2_misc.raw.csv"),
                            ("Train mode (`tools/train_profile.py`, one eager step after three warm-up steps)", "prof_r2_train.raw.csv")):
            path = os.path.join(OUT, name)
            if os.path.exists(path):
                f.write("## %s\n\n%s\n\n" % (title, table_from_raw(path)))
    print("wrote profiles/r02_launches_one_step.csv, profiles/r02_ncu_summary.md")


if __name__ == "__main__":
    main()
