"""Turn what tools/refresh_profiles.sh left in gpurun_out/ into the committed files under profiles/.

    python tools/update_profiles.py [round_tag]        (default r02)
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
N = 1048576


def have(name):
    return os.path.exists(os.path.join(G, name))


def run(args, **kw):
    return subprocess.run(args, capture_output=True, text=True, check=True, **kw).stdout


# ---- bench lines and tool outputs, as they are
for src, dst in (("bench_default.json", "bench_default_cfg3.json"), ("bench_reference.json", "bench_reference_arm.json"),
                 ("small_world_latency.txt", "small_world_latency.txt"), ("hub_scene.txt", "hub_scene.txt"),
                 ("dense_longrun.txt", "dense_longrun_cfg5.txt")):
    if have(src):
        shutil.copy(os.path.join(G, src), os.path.join(P, f"{tag}_{dst}"))
with open(os.path.join(P, f"{tag}_collide_bench.txt"), "w") as f:
    f.write("# python tools/collide_bench.py cfg3|cfg5: collision pipeline only, per-phase device times from the in-graph globaltimer stamps, L2 flushed before every step\n")
    for c in ("cfg3", "cfg5"):
        if have(f"collide_bench_{c}.txt"):
            f.write(open(os.path.join(G, f"collide_bench_{c}.txt")).read().strip().splitlines()[-1] + "\n")

# ---- launch lists
for name, cmd in (("collide_cfg3", "python tools/collide_bench.py cfg3 1048576 4"), ("collide_cfg5", "python tools/collide_bench.py cfg5 1048576 4"),
                  ("bench_cfg3", "python bench.py --workload cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e")):
    if have(f"launches_{name}.csv"):
        out = run([sys.executable, os.path.join(ROOT, "tools", "summarise_launches.py"), os.path.join(G, f"launches_{name}.csv"), cmd])
        open(os.path.join(P, f"{tag}_launches_{name}.csv"), "w").write(out)

# ---- ncu full captures -> text summaries
groups = {
    "collision_kernels_ncu.txt": (["prof_narrow_cfg3", "prof_narrow_cfg5", "prof_sparse_cfg3", "prof_prep_cfg3", "prof_prep_cfg5", "prof_flow_cfg5", "prof_tiny"],
                                  "narrow_kernel / sparse_step_kernel / prep_kernel (cfg3: with TR_NO_SPARSE=1) / response_flow_kernel at N = 1,048,576 (python tools/collide_bench.py cfg3|cfg5 1048576 6, "
                                  "launch 9 of the kernel) and tiny_step_kernel (python tools/small_world_latency.py, default scene)"),
    "broadphase_kernels_cfg3_ncu.txt": (["prof_broad_cfg3"], "the four broad-phase kernels of one timed step, N = 1,048,576 random cloud (python tools/collide_bench.py cfg3 1048576 6)"),
    "gravity_kernel_cfg3_ncu.txt": (["prof_gravity_cfg3"], "gravity_kernel, N = 1,048,576 (python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary, 2nd launch)"),
}
for fname, (reps, title) in groups.items():
    paths = [os.path.join(G, r + ".ncu-rep") for r in reps if have(r + ".ncu-rep")]
    if paths:
        out = run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), title] + paths)
        open(os.path.join(P, f"{tag}_{fname}"), "w").write(out)


def raw_rows(rep):
    rows = list(csv.reader(run(["ncu", "-i", os.path.join(G, rep + ".ncu-rep"), "--page", "raw", "--csv"]).splitlines()))
    h, u = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    out = []
    for v in rows[2:]:
        def val(key):
            i = h.index(key)
            return float(v[i]) * scale.get(u[i], 1.0)
        out.append({"name": v[h.index("Kernel Name")].split("(")[0].replace("void ", ""), "us": val("gpu__time_duration.sum"),
                    "dram": val("dram__bytes_read.sum") + val("dram__bytes_write.sum")})
    return out


# ---- the numbers bench.py reads back (traffic / ncu stopwatch), with their provenance
if have("prof_broad_cfg3.ncu-rep"):
    k = raw_rows("prof_broad_cfg3")
    # durations: the timing-only launch list (the profiling recipe's `--metrics gpu__time_duration.sum` pass, averaged over its
    # launches); DRAM bytes: the --set full capture of one step
    avg = {}
    lst = os.path.join(P, f"{tag}_launches_collide_cfg3.csv")
    if os.path.exists(lst):
        for ln in open(lst):
            f = ln.strip().split(",")
            if len(f) == 5 and f[0].startswith("tr::"):
                avg[f[0][4:]] = float(f[3])
    names = ["keys_kernel", "tile_sums_kernel", "scan_kernel", "scatter_kernel"]
    us = [avg.get(nm, next(x["us"] for x in k if x["name"].endswith(nm))) for nm in names]
    json.dump({f"cfg3:{N}": {"kernel_us_sum": round(sum(us), 2), "dram_bytes": int(sum(x["dram"] for x in k)),
                             "kernels": [[nm, round(u, 2)] for nm, u in zip(names, us)],
                             "source": f"durations: profiles/{tag}_launches_collide_cfg3.csv (ncu --metrics gpu__time_duration.sum, average per launch, cold caches and "
                                       f"serialised); DRAM bytes: profiles/{tag}_broadphase_kernels_cfg3_ncu.txt (ncu --set full, the four kernels of one step)"}},
              open(os.path.join(P, "broadphase_ncu.json"), "w"), indent=1)
nar = {}
for c in ("cfg3", "cfg5"):
    if have(f"prof_narrow_{c}.ncu-rep"):
        k = raw_rows(f"prof_narrow_{c}")[0]
        nar[f"{c}:{N}"] = {"dram_bytes": int(k["dram"]), "kernel_us": k["us"], "source": f"profiles/{tag}_collision_kernels_ncu.txt"}
if nar:
    json.dump(nar, open(os.path.join(P, "narrowphase_ncu.json"), "w"), indent=1)
if have("prof_gravity_cfg3.ncu-rep"):
    gt = json.load(open(os.path.join(P, "gravity_traffic.json")))
    gt[str(N)] = int(raw_rows("prof_gravity_cfg3")[0]["dram"])
    gt["_source"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per gravity_kernel launch: N = 1,048,576 from profiles/{tag}_gravity_kernel_cfg3_ncu.txt, "
                     "N = 65,536 from profiles/r01_gravity_kernel_cfg2_ncu.csv")
    json.dump(gt, open(os.path.join(P, "gravity_traffic.json"), "w"), indent=1)

# ---- SASS evidence (no GPU needed)
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dump_sass.py"), tag], check=True)

if have("bench_default.json"):
    d = json.load(open(os.path.join(G, "bench_default.json")))
    print("cfg3:", d["value"], "int/s, gravity frac", round(d["roofline"]["frac"], 4), "broad frac", round(d["roofline_broadphase"]["frac"], 4), d["collision_ms"])
    print("e2e:", d["e2e"]["ms_per_step"], "ms; full AoS", d["e2e_full_aos"]["ms_per_step"])
    for k, v in (d.get("secondary") or {}).items():
        if isinstance(v, dict) and "ms_per_step" in v:
            print(k, v["ms_per_step"], "ms/step", v.get("collision_ms"), "e2e", (v.get("e2e") or {}).get("ms_per_step"))
