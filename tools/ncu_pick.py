#!/usr/bin/env python3
"""Print a few headline metrics per launch from an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
extra = [a for a in sys.argv[2:]]
cols = [c for c in want if c in h] + [c for c in h if any(e in c for e in extra)]
for r in rows[2:]:
    print("---")
    for c in cols:
        i = h.index(c)
        print(f"  {c:80s} {r[i]:>16s} {u[i]}")
