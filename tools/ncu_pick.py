#!/usr/bin/env python
"""Prints the handful of metrics profiles/r01_summary.md quotes from an `ncu --set full` report.
    ncu -i gpurun_out/ris_r01f.ncu-rep --page raw --csv > /tmp/ris.csv && python tools/ncu_pick.py /tmp/ris.csv
(reports are scratch under gpurun_out/; the numbers are copied into profiles/ by hand)"""
import csv
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    header, units = rows[0], rows[1]
    for launch in rows[2:]:
        for i, h in enumerate(header):
            if h in WANT:
                print(f"{h} | {units[i]} | {launch[i]}")
        print("---")


if __name__ == "__main__":
    main()
