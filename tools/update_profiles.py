"""Turn what tools/refresh_profiles.sh left in gpurun_out/ into the committed files under profiles/.

    python tools/update_profiles.py [round_tag]        (default r01)
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

for src, dst in (("bench_default", "bench_default_cfg3"), ("bench_cfg5", "bench_cfg5"), ("bench_cfg2", "bench_cfg2"),
                 ("bench_reference", "bench_reference_arm")):
    shutil.copy(os.path.join(G, src + ".json"), os.path.join(P, f"{tag}_{dst}.json"))

for cfg in ("cfg3", "cfg5"):
    cmd = f"python bench.py --workload {cfg} --steps 2 --warmup 1 --no-cpu-baseline --no-e2e (final code)"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarise_launches.py"),
                          os.path.join(G, f"launches_{cfg}.csv"), cmd], capture_output=True, text=True, check=True).stdout
    open(os.path.join(P, f"{tag}_launches_{cfg}.csv"), "w").write(out)

# the four broad-phase kernels of the timed step (second group of the capture; the cell-table scans have 1024 CTAs)
raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_broad_r01b.ncu-rep"), "--page", "raw", "--csv"], capture_output=True,
                     text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
col = hdr.index
sel = rows[2 + 8:2 + 12]
lines = ["# ncu --set full --clock-control none --import-source on -k regex:keys_kernel|tile_sums_kernel|scan_kernel|scatter_kernel -c 12",
         "#   python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e  (final code, N = 1,048,576 random cloud)",
         "# the four kernels of the timed step's broad phase (cell-table scans: 1024 CTAs); cold caches per kernel under ncu",
         "kernel,grid,duration_us,dram_read_MB,dram_write_MB,dram_read_pct_of_peak,l2_hit_pct"]
tot = 0.0
for r in sel:
    name = r[col("Kernel Name")].split("(")[0].replace("void ", "")
    rd, wr = float(r[col("dram__bytes_read.sum")]), float(r[col("dram__bytes_write.sum")])
    tot += rd + wr
    lines.append(f"{name},\"{r[col('Grid Size')]}\",{float(r[col('gpu__time_duration.sum')]):.2f},{rd:.2f},{wr:.2f},"
                 f"{float(r[col('dram__bytes_read.sum.pct_of_peak_sustained_elapsed')]):.2f},{float(r[col('lts__t_sector_hit_rate.pct')]):.2f}")
lines.append(f"# total DRAM traffic of the four kernels: {tot:.1f} MB = {tot * 1e6 / 1048576:.1f} B/body (algorithmic budget 140 B/body);")
lines.append("# dirty lines written by one kernel are mostly still in L2 when the capture of that kernel ends, so writes are under-counted")
open(os.path.join(P, f"{tag}_broadphase_kernels_cfg3_ncu.csv"), "w").write("\n".join(lines) + "\n")
json.dump({"cfg3:1048576": int(tot * 1e6),
           "_source": f"sum of dram__bytes_read.sum + dram__bytes_write.sum over keys/tile_sums/scan/scatter of one step, profiles/{tag}_broadphase_kernels_cfg3_ncu.csv"},
          open(os.path.join(P, "broadphase_traffic.json"), "w"), indent=1)
d = json.load(open(os.path.join(P, f"{tag}_bench_default_cfg3.json")))
print("cfg3:", d["value"], "int/s, gravity frac", round(d["roofline"]["frac"], 4), "broad frac", round(d["roofline_broadphase"]["frac"], 4),
      d["collision_ms"])
d5 = json.load(open(os.path.join(P, f"{tag}_bench_cfg5.json")))
print("cfg5:", d5["ms_per_step"], "ms/step", d5["collision_ms"], "broad frac", round(d5["roofline"]["frac"], 4))
