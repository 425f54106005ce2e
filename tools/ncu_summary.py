#!/usr/bin/env python
"""Key numbers of an `ncu --set full` report (read on the CPU box): python tools/ncu_summary.py rep.ncu-rep [title]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_registers", "occ limit regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global ld requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global ld sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "global red requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "global red sectors"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instr"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency / inst"),
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([ln for ln in out.splitlines() if not ln.startswith("==")]))
    head, units = rows[0], rows[1]
    print(title)
    for r in rows[2:]:
        rec = dict(zip(head, r))
        un = dict(zip(head, units))
        print("kernel:", rec.get("Kernel Name", "?")[:110])
        seen = set()
        for k, label in KEYS:
            if k in rec and rec[k] != "" and k not in seen:
                seen.add(k)
                print("  %-28s %-18s %s   [%s]" % (label, rec[k], un.get(k, ""), k))
        try:
            tr = float(rec["dram__bytes_read.sum"].replace(",", "")) + float(rec["dram__bytes_write.sum"].replace(",", ""))
            print("  %-28s %.4g %s (read + write, this launch)" % ("dram traffic", tr, un.get("dram__bytes_read.sum", "")))
        except Exception:
            pass


if __name__ == "__main__":
    main()
