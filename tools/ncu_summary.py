#!/usr/bin/env python3
"""Summarises an .ncu-rep (one kernel, --set full) into profiles/: key metrics + per-launch DRAM traffic.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_<kernel>_rNN"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines, traffic = [], []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "?")
        lines.append(f"== {name} (launch id {d.get('ID', '?')})")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:90s} {d[k]:>16s} {units[hdr.index(k)]}")

        def num(k):
            v = float(d[k].replace(",", ""))
            u = units[hdr.index(k)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic.append(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
    open(out + "_summary.txt", "w").write("\n".join(lines) + "\n")
    json.dump({"source": rep, "dram_bytes_per_launch": sum(traffic) / len(traffic), "launches": len(traffic)},
              open(out + "_traffic.json", "w"))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
