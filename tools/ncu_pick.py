"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` output. Usage: ncu_pick.py raw.csv [pattern ...]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or [
    r"^gpu__time_duration.sum", r"^dram__bytes_(read|write).sum$", r"^dram__throughput.avg.pct", r"^lts__t_sector_hit_rate.pct",
    r"^l1tex__t_sector_hit_rate.pct", r"^lts__t_bytes.sum$", r"^lts__throughput.avg.pct", r"^l1tex__throughput.avg.pct",
    r"^sm__throughput.avg.pct", r"^sm__warps_active.avg.pct", r"^launch__registers_per_thread", r"^launch__occupancy_limit",
    r"^launch__grid_size", r"^launch__block_size", r"^sm__inst_executed_pipe_fma\.avg.pct", r"^smsp__inst_executed.sum$",
    r"^smsp__issue_active.avg.pct", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum$", r"^l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum$",
    r"^lts__t_sectors_srcunit_tex_op_read.sum$", r"^lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum$", r"^l1tex__m_xbar2l1tex_read_bytes.sum$",
    r"stalled_long_scoreboard_per_warp_active", r"stalled_lg_throttle_per_warp_active", r"stalled_.*_per_warp_active.pct",
    r"^sm__cycles_elapsed.avg$", r"^lts__cycles_elapsed.avg$", r"^dram__cycles_elapsed.avg$", r"sm__cycles_elapsed.avg.per_second", r"lts__cycles_elapsed.avg.per_second",
]
for r in rows[2:]:
    kn = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    print("==", kn[:80])
    for i, h in enumerate(hdr):
        if any(re.search(p, h) for p in pats):
            print(f"  {h} [{units[i]}] = {r[i]}")
    break
