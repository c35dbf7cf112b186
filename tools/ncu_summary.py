"""Developer tool: the judged subset of an `ncu --set full` report as text (profiles/*.txt are written with it).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "command line that was profiled" > profiles/r2_x_ncu.txt
"""
import csv
import subprocess
import sys

rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit",
        "smsp__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_ldgsts.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled", "sm__cycles_elapsed.avg ",
        "dram__cycles_active.avg.pct", "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct"]
print(f"# ncu --set full --clock-control none (one launch), report {rep.split('/')[-1]}")
if cmd:
    print(f"# profiled command: {cmd}")
for r in rows[2:]:
    print()
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(w) for w in WANT) and not h.endswith((".peak_sustained", ".per_second")) and ".max." not in h and ".min." not in h:
            print(f"{h:100s} {u:14s} {v}")
