#!/usr/bin/env python
"""Reduce `ncu --page raw --csv` (stdin) to the columns the roofline discussion uses (stdout, CSV)."""
import csv
import re
import sys

KEEP = re.compile(r"^(ID|Kernel Name|Grid Size|Block Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|dram__bytes_(read|write)\.sum\.per_second|"
                  r"dram__cycles_active|gpu__dram_throughput|sm__pipe_tensor.*cycles_active.*pct|sm__inst_executed_pipe_tensor.*pct|sm__warps_active.*pct|"
                  r"launch__registers_per_thread|launch__occupancy_limit|launch__shared_mem_per_block|lts__t_sector_hit_rate\.pct|sm__throughput.*pct|"
                  r"lts__throughput.*pct|l1tex__throughput.*pct|smsp__issue_active.*pct|sm__cycles_elapsed\.avg$|smsp__warp_issue_stalled.*pct)")
rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    sys.exit(0)
hdr = rows[0]
idx = [i for i, h in enumerate(hdr) if KEEP.search(h)]
w = csv.writer(sys.stdout)
for r in rows:
    if len(r) >= len(hdr):
        w.writerow([r[i] for i in idx])
