"""Summarise a .ncu-rep (one `ncu --set full` capture) into the handful of metrics DESIGN.md / profiles/README.md quote:
  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_ncu_x.csv"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(r"^(Kernel Name|Grid Size|Block Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|"
                  r"dram__bytes_(read|write)\.sum\.per_second|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__pipe_tensor_cycles_active.*pct_of_peak_sustained_(active|elapsed)|"
                  r".*sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__inst_executed_pipe_(fma|fmaheavy|alu|lsu|tensor.*hmma)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum|"
                  r"launch__registers_per_thread|launch__shared_mem_per_block_dynamic|launch__occupancy_limit.*|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|mio_throttle|barrier|math_pipe_throttle|lg_throttle|wait)_per_issue_active\.ratio)$")

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
for i, name in enumerate(head):
    if KEEP.match(name):
        w.writerow([name, units[i]] + [r[i] for r in rows[2:]])
