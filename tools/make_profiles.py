#!/usr/bin/env python
"""gpurun_out/ (scratch) -> profiles/ (tracked): launch list of one detect step, light metrics of every GEMM launch,
ncu --set full summaries, and profiles/roofline_traffic.json (DRAM bytes per GEMM launch, read by bench.py)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))
from summarize_ncu import summarize  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def read_long(path):
    """ncu --csv --log-file long format -> {id: {"name":..., "grid":..., "block":..., metric: value}} in launch order"""
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0] != "ID" and r[0].isdigit()]
    out = {}
    for r in rows:
        d = out.setdefault(int(r[0]), {"name": r[4].replace("<unnamed>::", "").replace("void ", "").split("(")[0], "grid": r[8], "block": r[7]})
        d[r[12]] = float(r[14].replace(",", ""))
    return [out[k] for k in sorted(out)]


def short(name):
    return name.replace("CUtensorMap_st", "tmap")


# ---- launch list of one step ------------------------------------------------------------------
ll = read_long(os.path.join(OUT, "launches.csv"))
tot = sum(d["gpu__time_duration.sum"] for d in ll)
with open(os.path.join(PROF, tag + "_launches_one_step.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: every launch of one detect_maps step (B=64, S=256)\n")
    f.write("# cold-cache, serialised times: compare shares, not absolutes. total %.3f ms over %d launches\n" % (tot / 1e6, len(ll)))
    f.write("idx,kernel,grid,block,time_us,share_pct\n")
    for i, d in enumerate(ll):
        f.write("%d,%s,\"%s\",\"%s\",%.2f,%.2f\n" % (i, short(d["name"]), d["grid"], d["block"], d["gpu__time_duration.sum"] / 1e3,
                                                      100 * d["gpu__time_duration.sum"] / tot))
fam = {}
for d in ll:
    k = d["name"].split("<")[0]
    fam[k] = fam.get(k, 0.0) + d["gpu__time_duration.sum"]
print("kernel shares of the step under ncu:", {k: "%.1f%%" % (100 * v / tot) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])})

# ---- every GEMM launch: time, DRAM bytes, tensor pipe ---------------------------------------------
gl = read_long(os.path.join(OUT, "gemm_all_light.csv"))
labels = []
lt = os.path.join(OUT, "layer_table.txt")
if os.path.exists(lt):
    labels = [l.split()[0] for l in open(lt) if " conv_gemm " in l]
rd = sum(d["dram__bytes_read.sum"] for d in gl)
wr = sum(d["dram__bytes_write.sum"] for d in gl)
with open(os.path.join(PROF, tag + "_gemm_all_launches.csv"), "w") as f:
    f.write("# every conv_gemm launch of one step: ncu light metric pass (time, DRAM bytes, tensor pipe, SM throughput)\n")
    f.write("idx,layer,kernel,grid,time_us,dram_read_MB,dram_write_MB,tensor_pipe_pct,sm_pct\n")
    for i, d in enumerate(gl):
        f.write("%d,%s,%s,\"%s\",%.2f,%.2f,%.2f,%.1f,%.1f\n" % (
            i, labels[i] if i < len(labels) else "", short(d["name"]), d["grid"], d["gpu__time_duration.sum"] / 1e3,
            d["dram__bytes_read.sum"] / 1e6, d["dram__bytes_write.sum"] / 1e6,
            d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"], d["sm__throughput.avg.pct_of_peak_sustained_elapsed"]))
traffic = {"conv_gemm_dram_bytes_per_launch": (rd + wr) / len(gl), "launches": len(gl), "dram_read_bytes_per_step": rd,
           "dram_write_bytes_per_step": wr,
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over all %d conv_gemm launches of one step (B=64, S=256), %s" % (len(gl), tag)}
json.dump(traffic, open(os.path.join(PROF, "roofline_traffic.json"), "w"), indent=1)
tw = sum(d["gpu__time_duration.sum"] * d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] for d in gl) / sum(d["gpu__time_duration.sum"] for d in gl)
print("GEMM launches: %d, DRAM read %.2f GB write %.2f GB per step, time-weighted tensor pipe %.1f%%" % (len(gl), rd / 1e9, wr / 1e9, tw))

# ---- full-set summaries -----------------------------------------------------------------------
md = ["# %s — ncu summaries (B200, B=64, S=256, synthetic maps, seeded random weights)\n" % tag,
      "Captured with `ncu --set full --clock-control none` under gpurun after a plain run of the same command "
      "(`tools/gpu_runs/prof_r1.sh`); tables made by `tools/make_profiles.py` from the .ncu-rep files (kept in gpurun_out/, not tracked). "
      "Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n"]
sets = [("prof_gemm_heads", "Class + mask head GEMMs", "FC1 12544->1024 (M=64000),FC2 1024->1024,class/bbox head N=20,mask conv1 3x3 (im2col TMA),mask conv2,mask conv3,mask conv4,deconv 2x2 + ReLU + 1x1 logits + sigmoid (fused)"),
        ("prof_gemm_res4b", "res4b bottleneck (M = 16384 rows)", "2a 1x1 1024->256,2b 3x3 256->256,2c 1x1 256->1024 + residual (TMA epilogue)"),
        ("prof_misc3", "Non-GEMM kernels of one step", None)]
for rep, title, labs in sets:
    path = os.path.join(OUT, rep + ".ncu-rep")
    if os.path.exists(path):
        md.append("## %s\n" % title)
        md.append(summarize(path, labs.split(",") if labs else None) + "\n")
open(os.path.join(PROF, tag + "_ncu_summary.md"), "w").write("\n".join(md))
print("wrote profiles/")
