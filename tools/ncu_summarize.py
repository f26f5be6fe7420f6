#!/usr/bin/env python
"""Turn ncu reports (gpurun_out/*.ncu-rep) into a markdown summary for profiles/ (run here, no GPU needed).

    python tools/ncu_summarize.py profiles/r01_ncu_summary.md gpurun_out/prof_*.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__bytes.sum.per_second", "DRAM throughput"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "L1/shared data pipe (LSU wavefronts)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "  of which shared memory"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(out_path, reps):
    md = ["# ncu summaries (`ncu --set full --clock-control none --import-source on`, one launch each)", "",
          "Per-launch values from cold-cache, serialised replays: compare shares, not absolutes (B200_PROFILING.md).", ""]
    for rep in reps:
        rows = ncu_csv(rep, "raw")
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(vals, units)))
            name = d.get("Kernel Name", ("?", ""))[0]
            md += [f"## `{rep.split('/')[-1]}`", "", f"kernel: `{name[:160]}`", "", "| metric | value |", "|---|---|"]
            for k, label in KEYS:
                if k in d:
                    md.append(f"| {label} (`{k}`) | {d[k][0]} {d[k][1]} |")
            md.append("")
        src = ncu_csv(rep, "source")
        if len(src) > 2:
            h = src[1]
            idx = {x: i for i, x in enumerate(h)}
            agg = collections.Counter()
            stall = collections.Counter()
            ops = collections.Counter()
            for r in src[2:]:
                if len(r) < len(h):
                    continue
                toks = r[idx["Source"]].split()
                if not toks:
                    continue
                op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
                ops[op.split(".")[0]] += 1
                wf, ideal = r[idx.get("L1 Wavefronts Shared", 0)], r[idx.get("L1 Wavefronts Shared Ideal", 0)]
                if wf not in ("", "0") and op.startswith(("LDS", "STS")):
                    agg[(op, "n")] += 1
                    agg[(op, "wf")] += int(wf)
                    agg[(op, "ideal")] += int(ideal)
                for k in h:
                    if k.startswith("stall_") and "Not Issued" not in k and r[idx[k]]:
                        stall[k] += int(r[idx[k]])
            md += ["shared-memory accesses (per SASS instruction class, summed over the launch):", "",
                   "| SASS | static count | wavefronts | ideal | ratio |", "|---|---|---|---|---|"]
            for op in sorted({k[0] for k in agg}):
                md.append(f"| {op} | {agg[(op, 'n')]} | {agg[(op, 'wf')]} | {agg[(op, 'ideal')]} | {agg[(op, 'wf')] / max(1, agg[(op, 'ideal')]):.3f} |")
            tot = sum(stall.values()) or 1
            md += ["", "warp-state samples: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in stall.most_common(8)), "",
                   "static SASS mix: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)), ""]
    open(out_path, "w").write("\n".join(md) + "\n")
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
