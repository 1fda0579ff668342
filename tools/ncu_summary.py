"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into a small text table for profiles/.

    python tools/ncu_summary.py "title / command that was profiled" report1.ncu-rep [report2.ncu-rep ...] > profiles/rNN_xxx_ncu.txt

One block per profiled kernel launch: duration, DRAM bytes, lanes active per instruction, issue-slot and pipe
utilisation, occupancy and the largest warp-stall reasons.  Reads the raw page with `ncu -i ... --page raw --csv`."""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("inst_executed", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads active per instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy % of 64 warps"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
]
STALL = "smsp__average_warps_issue_stalled_"

print(f"# {sys.argv[1]}")
print("# ncu --set full --clock-control none --import-source on (one launch each; caches flushed per replay pass: cold-cache numbers)")
for path in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        name = v[h.index("Kernel Name")].split("(")[0].replace("void ", "")
        print(f"\n== {name}   [{path.split('/')[-1]}]")
        for key, label in METRICS:
            if key in h:
                i = h.index(key)
                print(f"  {label:46s} {v[i]} {u[i]}")
        stalls = []
        for i, k in enumerate(h):
            if k.startswith(STALL) and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(v[i]), k[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  warp stalls per issued instruction, top 5       " + ", ".join(f"{n} {x:.2f}" for x, n in stalls[:5]))
