"""Turn `ncu -i X.ncu-rep --page raw --csv` (stdin) into the two-column text summary kept under profiles/.
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python scripts/ncu_raw_summary.py > profiles/NAME.ncu_raw.txt
Keeps identification columns plus the metric families the roofline discussion uses."""
import csv
import re
import sys

KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|Device|CC|gpu__time_duration|gpu__dram_throughput|dram__bytes|dram__cycles_active|"
                  r"dram__throughput|sm__pipe_tensor|sm__inst_executed_pipe_tensor|sm__warps_active|sm__throughput|sm__cycles_active|"
                  r"launch__|lts__t_sector_hit_rate|lts__t_bytes|lts__throughput|l1tex__m_xbar2l1tex_read_bytes|l1tex__throughput|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared|smsp__cycles_active|smsp__inst_executed\.sum|smsp__warp_issue_stalled.*_per_warp_active|"
                  r"sm__sass_inst_executed_op_shared|smsp__pcsamp_warps_issue_stalled|sm__pipe_fp64_cycles_active|sm__inst_executed_pipe_fp64|"
                  r"sm__inst_executed_pipe_lsu|l1tex__data_pipe_lsu_wavefronts|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|smsp__issue_active)")
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if KEEP.match(h):
            print(f"{(h + ('  [' + u + ']' if u else '')):110s}{v}")
    print()
