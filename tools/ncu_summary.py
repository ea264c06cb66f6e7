#!/usr/bin/env python3
"""One-screen summary of an ncu report: duration, DRAM traffic, pipe utilisation, issue rate,
registers / shared memory and the stall-reason split.   python tools/ncu_summary.py <report.ncu-rep>"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]

def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = {a: (b, c) for a, b, c in zip(h, u, v)}
        print("kernel:", d.get("Kernel Name", ("", ""))[1][:110])
        for k in KEYS:
            for name in d:
                if name.endswith(k):
                    print("  %-72s %-10s %s" % (name, d[name][0], d[name][1]))
        stalls = sorted(((float(d[n][1] or 0), n.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""))
                         for n in d if "average_warps_issue_stalled" in n and n.endswith("per_issue_active.ratio")), reverse=True)
        print("  stall reasons (warps per issue-active cycle):", ", ".join("%s %.2f" % (b, a) for a, b in stalls[:8]))

if __name__ == "__main__":
    main()
