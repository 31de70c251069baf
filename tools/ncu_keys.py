#!/usr/bin/env python
"""Print selected counters of every launch in an .ncu-rep (ncu -i ... --page raw --csv).
   python tools/ncu_keys.py gpurun_out/prof_x.ncu-rep [extra-metric-substring ...]"""
import csv, subprocess, sys
WANT = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("-" * 60)
    for w in WANT:
        if w in idx:
            print(f"{w:85s} {r[idx[w]][:100]} {units[idx[w]]}")
    for e in extra:
        for h in hdr:
            if e in h and h not in WANT:
                print(f"{h:85s} {r[idx[h]][:60]} {units[idx[h]]}")
