#!/usr/bin/env python
"""Summarises an `ncu --set full --import-source on` report of one kernel launch into a small JSON for profiles/:
the headline metrics (raw page), and from the source page (SASS) the share of executed warp instructions and of
stall samples per barrier-delimited region of the kernel (for the batch forward kernel these are the re-skew passes
and the eight sweeps) plus the instructions that collect the most stall samples.

  python benchmarks/ncu_summary.py gpurun_out/r2_v3b.ncu-rep profiles/r02_ncu_summary_v3b.json "free text: what was run"
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_config_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    what = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = ncu_csv(rep, "raw")
    hdr, units, row = raw[0], raw[1], raw[2]
    metrics, stalls = {}, {}
    for h, u, v in zip(hdr, units, row):
        if h in KEYS:
            metrics[h] = f"{v} {u}".strip()
        if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio"):
            stalls[h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]] = float(v)
    kernel = row[hdr.index("Kernel Name")]
    src = ncu_csv(rep, "source", ["--print-source", "sass"])
    h2 = src[1]
    ix = {h: i for i, h in enumerate(h2)}
    data = src[2:]
    tot_s = sum(int(r[ix["# Samples"]]) for r in data) or 1
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
    regions, cur = [], {"first": 0, "samples": 0, "inst": 0}
    for k, r in enumerate(data):
        cur["samples"] += int(r[ix["# Samples"]])
        cur["inst"] += int(r[ix["Instructions Executed"]])
        if "BAR.SYNC" in r[ix["Source"]]:
            cur["last"] = k
            regions.append(cur)
            cur = {"first": k + 1, "samples": 0, "inst": 0}
    cur["last"] = len(data) - 1
    regions.append(cur)
    regions = [{"sass_lines": [g["first"], g["last"]], "warp_inst_G": round(g["inst"] / 1e9, 3),
                "inst_share_pct": round(100.0 * g["inst"] / tot_i, 2), "stall_sample_share_pct": round(100.0 * g["samples"] / tot_s, 2)}
               for g in regions if g["inst"] > 0.002 * tot_i or g["samples"] > 0.002 * tot_s]
    scols = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:12]
    hot = []
    for r in top:
        best = max(scols, key=lambda h: int(r[ix[h]]))
        hot.append({"sass": " ".join(r[ix["Source"]].split()), "sample_share_pct": round(100.0 * int(r[ix["# Samples"]]) / tot_s, 2),
                    "main_stall": best[6:], "executed_M": round(int(r[ix["Instructions Executed"]]) / 1e6, 1)})
    out = {"what": what, "report": rep, "kernel": kernel, "metrics": metrics, "stalls_per_issue": stalls,
           "warp_instructions_total_G": round(tot_i / 1e9, 2), "regions_between_barriers": regions, "hottest_instructions": hot}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("kernel", "warp_instructions_total_G")}, indent=1))


if __name__ == "__main__":
    main()
