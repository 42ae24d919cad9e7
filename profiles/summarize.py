#!/usr/bin/env python
"""
Summarise an Nsight Compute report (.ncu-rep) into a small text file that can be committed (the
.ncu-rep itself stays in gpurun_out/, which is scratch).

    python profiles/summarize.py gpurun_out/prof_far_r2.ncu-rep "<ncu command line>" [units] [unit name] > profiles/far_sweep_r2.md

`units` = number of work units of the launch ((piece, point) pairs for the sweep, nodes + leaves
for the walk) for the per-unit instruction counts.
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_active.avg",
    "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of `%s`" % rep.split("/")[-1])
    print()
    print("command: `%s`" % cmd)
    print()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("## %s" % d.get("Kernel Name", "?"))
        print()
        print("| metric | value | unit |")
        print("|---|---|---|")
        for k in KEYS:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        pairs = float(sys.argv[3]) if len(sys.argv) > 3 else None
        unit = sys.argv[4] if len(sys.argv) > 4 else "(piece, point) pair"
        try:
            cyc = float(d["smsp__cycles_elapsed.avg"]) if "smsp__cycles_elapsed.avg" in d \
                else float(d["sm__cycles_elapsed.avg"])
            ops = {o: float(d["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % o]) * cyc
                   for o in ("dadd", "dmul", "dfma")}
            print()
            print("derived: executed FP64 arithmetic thread-instructions: DADD %.4g, DMUL %.4g, "
                  "DFMA %.4g; hardware flops (FMA = 2) %.4g" %
                  (ops["dadd"], ops["dmul"], ops["dfma"],
                   ops["dadd"] + ops["dmul"] + 2 * ops["dfma"]))
            if pairs:
                print()
                print("per %s (%.4g in this launch): DADD %.1f, DMUL %.1f, "
                      "DFMA %.1f = %.1f FP64 arithmetic instructions, %.1f hardware flops; "
                      "all warp-instructions x32 / unit = %.1f" %
                      (unit, pairs, ops["dadd"] / pairs, ops["dmul"] / pairs, ops["dfma"] / pairs,
                       sum(ops.values()) / pairs,
                       (ops["dadd"] + ops["dmul"] + 2 * ops["dfma"]) / pairs,
                       float(d["smsp__inst_executed.sum"]) * 32 / pairs))
        except Exception as exc:       # keep the summary usable if a metric is missing
            print()
            print("(derived figures unavailable: %s)" % exc)
        # stall reasons
        stalls = [(h, d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled")
                  and h.endswith("_per_issue_active.ratio")]
        stalls = sorted(stalls, key=lambda kv: -float(kv[1] or 0))[:8]
        if stalls:
            print()
            print("top warp stall reasons (warps per issue-active cycle):")
            print()
            for h, v in stalls:
                print("* %s = %s" % (h.replace("smsp__average_warps_issue_stalled_", "")
                                     .replace("_per_issue_active.ratio", ""), v))
        print()


if __name__ == "__main__":
    main()
