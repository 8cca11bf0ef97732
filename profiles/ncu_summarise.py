#!/usr/bin/env python3
"""Write the text summary kept under profiles/ for one `ncu --set full` capture: the key raw metrics of the first launch of
the named kernel plus the per-phase / per-opcode SASS breakdown (ncu_sass_phases.py).
usage: ncu_summarise.py capture.ncu-rep kernel-substring "header line" > summary.txt"""
import csv, io, os, subprocess, sys, tempfile
rep, kern, header = sys.argv[1], sys.argv[2], sys.argv[3]
KEYS = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__occupancy_limit", "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_alu.sum.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__throughput.avg.pct",
        "sm__warps_active.avg.pct", "smsp__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "smsp__average_warp", "smsp__pcsamp_warps_issue_stalled")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
names, units = rows[0], rows[1]
ki = names.index("Kernel Name")
row = next(r for r in rows[2:] if kern in r[ki])
print(f"# {header}")
print(f"# kernel: {row[ki]}   (ncu -i {os.path.basename(rep)} --page raw --csv; ncu --set full --clock-control none under gpurun, B200)")
for n, u, v in sorted(zip(names, units, row)):
    if n.startswith(KEYS) and v not in ("", "n/a"):
        print(f"{n} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern.split("<")[0].split("(")[0]], capture_output=True, text=True).stdout
with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
    f.write(src)
print("\n# per-phase / per-opcode breakdown (profiles/ncu_sass_phases.py on --page source --print-source sass)")
print(subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_sass_phases.py"), f.name], capture_output=True, text=True).stdout)
os.unlink(f.name)
