#!/usr/bin/env python
"""ncu CSV of tools/ncu_metrics_target.py -> profiles/roofline_traffic.json (per-size DRAM bytes per launch of the final
kernels, FFT_multiple FMA-pipe utilisation).  Runs here, no GPU.   python tools/ncu_metrics_parse.py in.csv out.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
idx = {k: i for i, k in enumerate(h)}
launches = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    launches.setdefault(int(r[idx["ID"]]), {"kernel": r[idx["Kernel Name"]]})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
L = [launches[k] for k in sorted(launches)]
sizes = [32, 64, 128, 256, 512, 1024, 2048, 4096]
keys = [f"{n}{'r' if r else 'n'}" for n in sizes for r in (1, 0)]
out = {"source": sys.argv[1], "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none, one launch per configuration on the 4 GiB batch",
       "per_size": {}, "multiple": {}, "real": {}}


def unit_bytes(d, name):
    return d.get(name, 0.0)


ext, mul, real = L[:16], L[16:32], L[32:40]
for k, d in zip(keys, ext):
    out["per_size"][k] = {"dram_bytes": unit_bytes(d, "dram__bytes_read.sum") + unit_bytes(d, "dram__bytes_write.sum"),
                          "read": unit_bytes(d, "dram__bytes_read.sum"), "write": unit_bytes(d, "dram__bytes_write.sum"),
                          "ncu_us": d.get("gpu__time_duration.sum", 0) / 1e3, "smem_pipe_pct": d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                          "issue_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "kernel": d["kernel"][:120]}
for k, d in zip(keys, mul):
    out["multiple"][k] = {"fma_pipe_pct": d.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                          "smem_pipe_pct": d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                          "issue_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "ncu_us": d.get("gpu__time_duration.sum", 0) / 1e3}
for k, d in zip([f"{n}_{m}" for n in (512, 1024, 2048, 4096) for m in ("r2c", "c2r")], real):
    out["real"][k] = {"dram_bytes": unit_bytes(d, "dram__bytes_read.sum") + unit_bytes(d, "dram__bytes_write.sum"), "ncu_us": d.get("gpu__time_duration.sum", 0) / 1e3,
                      "smem_pipe_pct": d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                      "issue_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active")}
vals = [v["dram_bytes"] for v in out["per_size"].values() if v["dram_bytes"] > 0]
out["dram_bytes_per_launch"] = sum(vals) / len(vals) if vals else None
out["algorithmic_bytes_per_launch"] = 8589934592
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: round(v["dram_bytes"] / 8589934592, 4) for k, v in out["per_size"].items()}))
print(json.dumps({k: v["fma_pipe_pct"] for k, v in out["multiple"].items()}))
