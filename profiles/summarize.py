#!/usr/bin/env python
"""Turn an ncu report (gpurun_out/*.ncu-rep, scratch) into the small text summary kept under profiles/.

    python profiles/summarize.py gpurun_out/prof_r1b.ncu-rep profiles/r01_full_b.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n== {d['Kernel Name']}  (launch id {d.get('ID', '?')})\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:70s} {d[k]:>18s} {units[hdr.index(k)]}\n")
            st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[h])
                  for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")}
            f.write("warp stall reasons per issue (top 8): "
                    + ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
