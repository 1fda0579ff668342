#!/usr/bin/env python
"""bench.py - throughput of the TraceIt physics step on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5]
    python bench.py --impl reference ...      # the reference's own CPU step, same JSON shape

One "step" = one pass of the hot path over the synthetic cloud the workload names.
Rank 0 prints ONE JSON line.  `value` is ordered body-pair interactions per second over the
whole job (N^2 per step; the self pair is included and contributes zero) with
inputs resident in HBM; `e2e` is the same metric through the C-ABI host-buffer call
(tr_step_host: host AoS -> device, step, device -> host AoS every step).

Workloads (BASELINE.json configs):
  cfg2  64K-body cloud: all-pairs gravity + drag + integration, collisions off
  cfg3  1M-body cloud: gravity + drag + integration + grid broad-phase sphere collisions  [default, 1 GPU]
  cfg4  4M-body cloud: sharded gravity, NCCL all-gather of float4(x,y,z,m) per step       [default, N>1 GPUs]
  cfg5  1M packed spheres: collisions only (broad phase + narrow phase + ordered response)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LANE_OPS_PER_INTERACTION = 12      # 3 FADD + (1 FMUL + 2 FFMA) + 3 FMUL + 3 FFMA  
BROADPHASE_BYTES_PER_BODY = 140    # keygen 24 + sort 68 + ranges 12 + gather 36
CPU_SAMPLE_N = 16384               # bodies in the bounded CPU sample
GRAVITY_G = 1e-3                   # well-conditioned regime (no hard encounters over the run)

WORKLOADS = {
    "cfg2": dict(n=65536, gravity=True, collisions=False, desc="64K-body random cloud: all-pairs gravity + drag + integration, fp32, collisions off"),
    "cfg3": dict(n=1 << 20, gravity=True, collisions=True, desc="1M-body cloud: gravity + drag + grid broad-phase sphere collisions"),
    "cfg4": dict(n=1 << 22, gravity=True, collisions=False, desc="4M-body cloud: sharded gravity with NCCL position all-gather over NVLink"),
    "cfg5": dict(n=1 << 20, gravity=False, collisions=True, desc="Dense-collision stress: 1M small spheres in a packed box"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=None)
    ap.add_argument("--bodies", type=int, default=None, help="override the workload's body count")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush between timed steps")
    return ap.parse_args()


def make_scene(name: str, n: int) -> np.ndarray:
    from traceit_b200 import scenes
    if name == "cfg2":
        return scenes.random_cloud(n, scenes.SEED_CFG2, spheres=True)
    if name == "cfg3":
        return scenes.random_cloud(n, scenes.SEED_CFG3, spheres=True)
    if name == "cfg4":
        return scenes.random_cloud(n, scenes.SEED_CFG4, spheres=True)
    return scenes.cfg5_dense(n)


# ------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], None, [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load": samples at or above the median power draw
        if sm:
            med_p = statistics.median(pw)
            loaded = [c for c, p in zip(sm, pw) if p >= med_p] or sm
            sm_med = statistics.median(loaded)
        else:
            sm_med = None
        return {"sm_mhz": sm_med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------
def cpu_reference_step_rate(name: str, steps: int, warmup: int, n: int = CPU_SAMPLE_N):
    """The reference's own CPU step on a bounded sample of the workload's cloud.

    Uses oracle/_ref (the unmodified reference TU, 1 thread: the reference is single-threaded)
    when it was built, else the restated C port.  A reference step visits every unordered pair
    once (src/main.cpp:674-675) = 2 ordered interactions."""
    from oracle import pyoracle as po
    b = make_scene(name, n)
    cfg = WORKLOADS[name]
    if po.ref_available():
        kind = "reference"
        w = po.RefWorld(1.2)
        w.add(b)
        stepper = lambda k: w.step(k)  # noqa: E731
        cores = 1
        what = "verbatim reference updatePhysics (oracle/_ref, -O2 -ffp-contract=off): drag + uniform gravity + integration + O(N^2) collision sweep"
    else:
        kind = "port"
        p = po.default_params()
        stepper = lambda k: po.step(b, p, k)  # noqa: E731
        cores = 1
        what = "restated C oracle (oracle/restated.c): same step, 1 thread"
    stepper(max(0, warmup))
    t0 = time.perf_counter()
    stepper(steps)
    dt = time.perf_counter() - t0
    pairs = n * (n - 1) / 2 * steps
    out = {
        "value": 2 * pairs / dt, "unit": "interactions/s", "cores": cores, "kind": kind,
        "sample": f"{steps} steps of the {name} cloud cut to N={n} ({what}); {pairs / dt:.3e} unordered pair visits/s x2",
        "pair_visits_per_s": pairs / dt, "ms_per_step": dt / steps * 1e3, "n_bodies": n,
    }
    # the reference as shipped: its CMakeLists.txt sets no build type, i.e. -O0 (same bits, ~5x slower)
    if po.ref_available(o0=True):
        n0 = 8192
        w0 = po.RefWorld(1.2, o0=True)
        w0.add(make_scene(name, n0))
        w0.step(1)
        t0 = time.perf_counter()
        w0.step(2)
        dt0 = time.perf_counter() - t0
        out["as_shipped_O0"] = {"value": n0 * (n0 - 1) * 2 / dt0, "unit": "interactions/s", "cores": 1, "kind": "reference",
                                "sample": f"2 steps at N={n0}, verbatim reference built -O0 as its CMakeLists.txt does"}
        w0.close()
    # O(N^2) sweep: extrapolated (NOT measured) time of one reference step at the workload's size
    out["extrapolated_s_per_step_at_workload_n"] = cfg["n"] * (cfg["n"] - 1) / 2 / (pairs / dt)
    if cfg["gravity"]:
        # second line: the restated pair-gravity term (dead code in the reference) on all host cores
        threads = os.cpu_count() or 1
        m = n
        t0 = time.perf_counter()
        po.pair_gravity_f32(b, GRAVITY_G)
        dt2 = time.perf_counter() - t0
        out["port_pair_gravity"] = {"value": m * (m - 1) / dt2, "unit": "interactions/s", "cores": threads, "kind": "port",
                                    "sample": f"one restated pair-gravity pass (OpenMP over targets) at N={m}"}
    return out


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    name = args.workload or ("cfg3" if args.gpus == 1 else "cfg4")
    r = cpu_reference_step_rate(name, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "body-pair interactions/sec", "value": r["value"], "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "n_bodies_sample": r["n_bodies"], "note": "CPU reference on a bounded sample; pair rate is size independent (O(N^2) sweep)"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout: anything a library prints there (NCCL's version
    banner under NCCL_DEBUG=VERSION, for one) is sent to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        protect_stdout()
        run_reference_arm(args, rank)
        return

    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch ourselves one rank per GPU (the driver does this itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    protect_stdout()
    import torch
    import torch.distributed as dist

    import traceit_b200 as tr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible and there is no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload or ("cfg3" if world_size == 1 else "cfg4")
    cfg = WORKLOADS[name]
    n = args.bodies or cfg["n"]
    bodies = make_scene(name, n)
    n = len(bodies)

    flags = (tr.PAIR_GRAVITY if cfg["gravity"] else 0) | (tr.COLLISIONS if cfg["collisions"] else 0)
    params = tr.default_params(flags=flags, G=np.float32(GRAVITY_G), g=(0, 0, 0))

    if world_size > 1:
        uid =torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(tr.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        world = tr.World(params, device=local_rank, rank=rank, nranks=world_size, unique_id=bytes(uid.cpu().numpy().tobytes()))
    else:
        world = tr.World(params, device=local_rank)

    # a dedicated torch stream: the library launches on it and torch.cuda.Event records on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    world.set_stream(stream.cuda_stream)
    world.add(bodies)
    own_first, own_count = world.owned_range()
    world.step(0)
    world.enable_timing(True)

    flush_buf = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(max(args.warmup, 0)):
        world.step(1, sync=False)
    world.sync()
    world.reset_stats()

    # ---- timed region: K steps, each bracketed by CUDA events on the launching stream; an L2
    # flush (256 MB memset, untimed) separates consecutive steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        if flush_buf is not None:
            flush_buf.zero_()
        ev[k][0].record(stream)
        world.step(1, sync=False)
        ev[k][1].record(stream)
    world.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    st = world.stats()
    launches = int(st.kernel_launches)

    tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms_max = float(tm.item())
    ms_per_step = total_ms_max / args.steps
    interactions_per_step = float(n) * float(n)     # ordered pairs, self pair included
    value = interactions_per_step * args.steps / (total_ms_max * 1e-3)

    # ---- per-kernel numbers (library-side CUDA events around each launch, same timed region)
    grav_ms = st.gravity_ms / max(st.gravity_launches, 1)
    broad_ms = st.broadphase_ms / max(st.broadphase_launches, 1)
    roofline = None
    extra = {}
    if cfg["gravity"] and st.gravity_launches:
        peak_ops, peak_clk = tr.measure_fma_peak(local_rank)
        achieved = float(own_count) * float(n) * LANE_OPS_PER_INTERACTION / (grav_ms * 1e-3)
        traffic = None
        prof = os.path.join(ROOT, "profiles", "gravity_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get(str(n))
            except Exception:
                traffic = None
        roofline = {
            "bound": "fp32_fma", "kernel": "gravity_kernel", "achieved": achieved / 1e12, "peak": peak_ops / 1e12,
            "unit": "Tlaneop/s", "frac": achieved / peak_ops, "traffic": traffic,
            "peak_source": f"measured live: FFMA-only kernel, {peak_clk:.0f} MHz (MEASURED_PEAKS.json has no FP32 row)",
            "algorithmic": f"{LANE_OPS_PER_INTERACTION} FP32-pipe lane-ops x {own_count} targets x {n} sources per launch",
            "kernel_ms": grav_ms, "share_of_step": grav_ms / ms_per_step,
            "interactions_per_s_kernel": float(own_count) * float(n) / (grav_ms * 1e-3),
        }
    if cfg["collisions"] and st.broadphase_launches:
        hbm = 6538.0
        src = "fallback 6538 (file absent)"
        try:
            hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        ach = BROADPHASE_BYTES_PER_BODY * n / (broad_ms * 1e-3) / 1e9
        bp_traffic = None
        prof = os.path.join(ROOT, "profiles", "broadphase_traffic.json")
        if os.path.exists(prof):
            try:
                bp_traffic = json.load(open(prof)).get(f"{name}:{n}")   # ncu capture of this workload only
            except Exception:
                bp_traffic = None
        bp = {"bound": "hbm", "kernel": "broad phase K2-K3b (keys + histogram, cell-table scan, scatter of 32-byte body records into cell order)", "achieved": ach,
              "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": bp_traffic, "peak_source": src,
              "algorithmic": f"{BROADPHASE_BYTES_PER_BODY} B/body x {n} bodies", "kernel_ms": broad_ms}
        if roofline is None:
            roofline = bp
        else:
            extra["roofline_broadphase"] = bp
        extra["collision_ms"] = {"broad": broad_ms, "narrow": st.narrowphase_ms / max(st.broadphase_launches, 1),
                                 "response": st.response_ms / max(st.broadphase_launches, 1),
                                 "contacts_last_step": int(st.last_contacts), "response_rounds_last_step": int(st.last_response_rounds)}
    if world_size > 1:
        extra["allgather_ms"] = st.allgather_ms / max(args.steps, 1)

    # ---- e2e: the C-ABI host-buffer call, host<->device copies inside the timed region
    e2e = None
    if not args.no_e2e:
        host = world.read_bodies(own_first, own_count)
        world.step_host(host, 1)          # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            world.step_host(host, 1)
        barrier()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": interactions_per_step * args.steps / dt, "unit": "interactions/s",
               "h2d_bytes_per_step": int(own_count) * 92, "d2h_bytes_per_step": int(own_count) * 92,
               "ms_per_step": dt / args.steps * 1e3, "api": "tr_step_host (AoS tr_body_desc in/out per step)"}

    # ---- CPU baseline beside it (rank 0, 1 GPU only)
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_step_rate(name, steps=12, warmup=1)
        except Exception as e:  # the baseline is a reported figure, never a reason to lose the GPU number
            cpu = {"value": None, "unit": "interactions/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}

    world.close()

    # ---- secondary line (1 GPU, default workload only): BASELINE.json configs[1], the 64K gravity-only cloud
    secondary = None
    if world_size == 1 and args.workload is None and args.bodies is None:
        try:
            scfg = WORKLOADS["cfg2"]
            sb = make_scene("cfg2", scfg["n"])
            sp = tr.default_params(flags=tr.PAIR_GRAVITY, G=np.float32(GRAVITY_G), g=(0, 0, 0))
            with tr.World(sp, device=local_rank) as w2:
                w2.set_stream(stream.cuda_stream)
                w2.add(sb)
                w2.step(5)
                ssteps = 40
                sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ssteps)]
                for k in range(ssteps):
                    if flush_buf is not None:
                        flush_buf.zero_()
                    sev[k][0].record(stream)
                    w2.step(1, sync=False)
                    sev[k][1].record(stream)
                w2.sync()
                sms = sum(a.elapsed_time(b) for a, b in sev) / ssteps
            sn = float(scfg["n"])
            peak_ops2, _ = tr.measure_fma_peak(local_rank)
            secondary = {"workload": f"cfg2: {scfg['desc']}", "n_bodies": scfg["n"], "steps": ssteps, "ms_per_step": sms,
                         "value": sn * sn / (sms * 1e-3), "unit": "interactions/s", "steps_per_s": 1e3 / sms,
                         "roofline_frac": sn * sn * LANE_OPS_PER_INTERACTION / (sms * 1e-3) / peak_ops2}
        except Exception as e:   # never lose the headline over the secondary line
            secondary = {"error": str(e)}

    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    line = {
        "metric": "body-pair interactions/sec", "value": value, "unit": "interactions/s", "n_gpus": world_size,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "steps_per_s": 1e3 / ms_per_step,
        "higher_is_better": True, "scaling": "strong" if world_size > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: {cfg['desc']}", "n_bodies": n, "G": GRAVITY_G, "g": [0, 0, 0],
                   "parallelism": f"targets sharded over {world_size} GPU(s), sources replicated",
                   "l2": "flushed between timed steps (256 MB memset, outside the per-step events)" if flush_buf is not None else "not flushed",
                   "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks",
                   "interaction": ("ordered (target, source) gravity evaluations, N^2 per step" if cfg["gravity"] else
                                   "ordered pair visits the reference's O(N^2) sweep would make, N^2 per step "
                                   "(the grid broad phase prunes them; see collision_ms / steps_per_s for the physical cost)")},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3, "secondary": secondary,
    }
    line.update(extra)
    emit(line)


if __name__ == "__main__":
    main()
