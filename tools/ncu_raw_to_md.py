#!/usr/bin/env python
"""tools/ncu_raw_to_md.py out.md title raw.csv [raw.csv ...] -- compact per-kernel table from `ncu -i rep --page raw --csv`
exports (several kernels per report are fine): duration, grid, block, registers, DRAM bytes, issue / shared-pipe utilisation,
occupancy, top warp-stall reasons.  Runs here, no GPU."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "duration us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("dram__bytes.sum.per_second", "DRAM B/s"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("smsp__inst_executed.sum", "warp instructions")]
out, title, files = sys.argv[1], sys.argv[2], sys.argv[3:]
md = [f"# {title}", "", "`ncu --set full --clock-control none`, one launch per kernel on a 1 GiB batch; cold-cache, serialised replays: compare shares, not absolutes.", ""]
for path in files:
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    idx = {k: i for i, k in enumerate(h)}
    stall = [k for k in h if "smsp__pcsamp_warps_issue_stalled" in k and "not_issued" not in k]
    md += [f"## `{path.split('/')[-1]}`", "", "| kernel | " + " | ".join(n for _, n in WANT) + " | top stalls |", "|---|" + "---|" * (len(WANT) + 1)]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("void ", "").split("(")[0].replace("smfft::big::<unnamed>::", "")
        vals = []
        for k, _ in WANT:
            v = r[idx[k]] if k in idx else ""
            u = units[idx[k]] if k in idx else ""
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {u}".strip())
        tot = sum(float(r[idx[k]] or 0) for k in stall) or 1.0
        top = sorted(((float(r[idx[k]] or 0) / tot, k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k in stall), reverse=True)[:5]
        md.append(f"| `{name}` | " + " | ".join(vals) + " | " + ", ".join(f"{b} {a * 100:.0f}%" for a, b in top) + " |")
    md.append("")
open(out, "w").write("\n".join(md))
