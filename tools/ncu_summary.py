#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report, one column per launch (what profiles/*_summary.txt hold).
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>_summary.txt
"""
import csv, io, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H, units, data = rows[0], rows[1], rows[2:]
ki = H.index("Kernel Name")
print(f"# {rep}: {len(data)} launch(es)")
for i, r in enumerate(data):
    print(f"# [{i}] {r[ki][:150]}")
for w in WANT:
    hits = [i for i, h in enumerate(H) if h == w or h.endswith("." + w)]
    if not hits:
        continue
    i = hits[0]
    print(f"{w:95s} {units[i]:16s} " + "  ".join(r[i] for r in data))
