#!/usr/bin/env python
"""Reduce `ncu -i X.ncu-rep --page raw --csv` to the columns the roofline discussion uses.
usage: ncu -i rep --page raw --csv | python tools/ncu_extract.py > profiles/<name>.csv"""
import csv
import sys

KEEP = ("ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "gpc__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct")

rows = list(csv.reader(sys.stdin))
hdr = rows[0]
idx = [i for i, h in enumerate(hdr) if h in KEEP]
w = csv.writer(sys.stdout)
for r in rows:
    if len(r) == len(hdr):
        w.writerow([r[i] if i == hdr.index("Kernel Name") and False else (r[i].split("(")[0] if hdr[i] == "Kernel Name" else r[i]) for i in idx])
