"""Print selected metrics from an `ncu --page raw --csv` export. usage: raw_metrics.py <raw.csv> [regex]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
    r"gpu__time_duration.sum|sm__throughput.avg.pct|smsp__inst_executed.sum$|sm__inst_executed_pipe_(alu|fma|lsu|xu|fp64|uniform)[^.]*\.sum$|"
    r"smsp__issue_active.avg.pct|sm__warps_active.avg.pct_of_peak|l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|"
    r"lts__throughput.avg.pct|l1tex__throughput.avg.pct|dram__bytes_(read|write).sum$|lts__t_sectors.sum$|l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum$|"
    r"l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum$|launch__registers_per_thread|launch__occupancy_limit|sm__maximum_warps|achieved_occupancy|"
    r"smsp__average_warps?_issue_stalled_.*_per_issue_active|smsp__pcsamp_warps_issue_stalled_[a-z_]+$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
    r"smsp__inst_executed_op_(global|shared|local)_(ld|st).sum$|local_load|local_store|derived__smsp__inst_executed_op_branch")
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    print('==', name[:90])
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            print(f"  {h:80s} {v} {u}")
