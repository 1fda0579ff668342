#!/usr/bin/env python
"""bench.py - throughput of the TraceIt physics step on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5]
    python bench.py --impl reference ...      # the reference's own CPU step, same JSON shape

One "step" = one pass of the hot path over the synthetic cloud the workload names.
Rank 0 prints ONE JSON line.  `value` is ordered body-pair interactions per second over the
whole job (N^2 per step; the self pair is included and contributes zero) with
inputs resident in HBM; `e2e` is the same metric through the C ABI with HOST buffers and the copies
inside the timed region: tr_step_frame (the step's per-frame input, the external forces, host -> device;
step; what the step changes - pos, vel, collider centre - device -> host), and `e2e_full_aos` through
tr_step_host (the whole 92-byte descriptor of every body both ways every step).

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
NARROW_BYTES_PER_BODY = 16         # "R16 own + 8 B per emitted contact"
NARROW_BYTES_PER_CONTACT = 8
RESPONSE_BYTES_PER_CONTACT = 136   # pair 8 + pos/vel of two bodies read (64) and written (64)
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
    ap.add_argument("--no-secondary", action="store_true", help="only the main workload (skip cfg2/cfg5/T1/parity lines)")
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
# CPU legs (the only places bench.py executes oracle/ as something that is TIMED; the sharded
# parity check below uses it as a checker)
# ------------------------------------------------------------------------------------------
def _omp_threads(k: int | None):
    """Set the OpenMP team size of the restated oracle's parallel loops (None = all cores)."""
    import ctypes
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(k or (os.cpu_count() or 1)))
    except OSError:
        pass


def _ref_rate(name: str, n: int, steps: int, warmup: int, o0: bool = False):
    """Verbatim reference updatePhysics (oracle/_ref) on the workload's cloud generator at n bodies."""
    from oracle import pyoracle as po
    w = po.RefWorld(1.2, o0=o0)
    w.add(make_scene(name, n))
    w.step(max(0, warmup))
    t0 = time.perf_counter()
    w.step(steps)
    dt = time.perf_counter() - t0
    w.close()
    return dt


def cpu_reference_step_rate(name: str, steps: int, warmup: int, n: int = CPU_SAMPLE_N, extended: bool = False):
    """The reference's own CPU step on a bounded sample of the workload's cloud.

    Uses oracle/_ref (the unmodified reference TU, 1 thread: the reference is single-threaded)
    when it was built, else the restated C port.  A reference step visits every unordered pair
    once (src/main.cpp:674-675) = 2 ordered interactions.  extended=True adds the lines SURVEY.md 8d
    asks for beside the headline sample: N = 65,536, the as-shipped -O0 build, and the restated
    pair-gravity pass (what the GPU arm's interactions are) on ONE core and on all cores."""
    from oracle import pyoracle as po
    b = make_scene(name, n)
    cfg = WORKLOADS[name]
    if po.ref_available():
        kind = "reference"
        dt = _ref_rate(name, n, steps, warmup)
        cores = 1
        what = "verbatim reference updatePhysics (oracle/_ref, -O2 -ffp-contract=off): drag + uniform gravity + integration + O(N^2) collision sweep"
    else:
        kind = "port"
        p = po.default_params()
        bb = b.copy()
        po.step(bb, p, max(0, warmup))
        t0 = time.perf_counter()
        po.step(bb, p, steps)
        dt = time.perf_counter() - t0
        cores = 1
        what = "restated C oracle (oracle/restated.c): same step, 1 thread"
    pairs = n * (n - 1) / 2 * steps
    out = {
        "value": 2 * pairs / dt, "unit": "interactions/s", "cores": cores, "kind": kind,
        "sample": f"{steps} steps of the {name} cloud cut to N={n} ({what}); {pairs / dt:.3e} unordered pair visits/s x2",
        "pair_visits_per_s": pairs / dt, "ms_per_step": dt / steps * 1e3, "n_bodies": n,
    }
    # O(N^2) sweep: extrapolated (NOT measured) time of one reference step at the workload's size
    out["extrapolated_s_per_step_at_workload_n"] = cfg["n"] * (cfg["n"] - 1) / 2 / (pairs / dt)
    if not extended:
        return out
    if po.ref_available():
        # the second size SURVEY.md 8d names
        n2, s2 = 65536, 2
        dt2 = _ref_rate(name, n2, s2, 0)
        out["sample_65536"] = {"value": n2 * (n2 - 1) * s2 / dt2, "unit": "interactions/s", "cores": 1, "kind": "reference",
                               "ms_per_step": dt2 / s2 * 1e3, "sample": f"{s2} steps at N={n2}, same build, no warm-up"}
    # the reference as shipped: its CMakeLists.txt sets no build type, i.e. -O0 (same bits, ~3.4x slower)
    if po.ref_available(o0=True):
        n0 = 8192
        dt0 = _ref_rate(name, n0, 2, 1, o0=True)
        out["as_shipped_O0"] = {"value": n0 * (n0 - 1) * 2 / dt0, "unit": "interactions/s", "cores": 1, "kind": "reference",
                                "sample": f"2 steps at N={n0}, verbatim reference built -O0 as its CMakeLists.txt does"}
    if cfg["gravity"]:
        # like for like with what the GPU arm counts: the restated pair-gravity term (dead code in the
        # reference, src/main.cpp:679-693), one pass over N targets x N sources
        m = n
        threads = os.cpu_count() or 1
        _omp_threads(1)
        t0 = time.perf_counter()
        po.pair_gravity_f32(b, GRAVITY_G)
        dt1 = time.perf_counter() - t0
        _omp_threads(None)
        t0 = time.perf_counter()
        po.pair_gravity_f32(b, GRAVITY_G)
        dt2 = time.perf_counter() - t0
        out["port_pair_gravity_1core"] = {"value": m * (m - 1) / dt1, "unit": "interactions/s", "cores": 1, "kind": "port",
                                          "sample": f"one restated pair-gravity pass at N={m}, 1 thread"}
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
        "config": {"workload": name, "n_bodies_sample": r["n_bodies"],
                   "note": "CPU reference on a bounded sample; pair rate is size independent (O(N^2) sweep). Its interaction is a "
                           "collision pair visit (x2 for ordered pairs): pair gravity is dead code in the reference, so the "
                           "like-for-like CPU figure for the GPU arm's gravity interactions is cpu_baseline.port_pair_gravity* "
                           "in the GPU arm's line"},
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


def _profile_json(fname: str, key: str):
    path = os.path.join(ROOT, "profiles", fname)
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(key)
        except Exception:
            return None
    return None


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6538.0, "fallback 6538 (file absent)"


# ------------------------------------------------------------------------------------------
# one workload on the GPU(s)
# ------------------------------------------------------------------------------------------
class Ctx:
    """torch / NCCL plumbing shared by the workloads of one bench.py process."""

    def __init__(self, torch, dist, tr, rank, world_size, local_rank, no_flush):
        self.torch, self.dist, self.tr = torch, dist, tr
        self.rank, self.world_size, self.local_rank = rank, world_size, local_rank
        self.dev = torch.device("cuda", local_rank)
        # a dedicated torch stream: the library launches on it and torch.cuda.Event records on it
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.flush_buf = None if no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world_size > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world_size > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def make_world(self, params, sharded: bool):
        tr, torch = self.tr, self.torch
        if sharded and self.world_size > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device=self.dev)
            if self.rank == 0:
                uid = torch.frombuffer(bytearray(tr.comm_unique_id()), dtype=torch.uint8).to(self.dev)
            self.dist.broadcast(uid, 0)
            w = tr.World(params, device=self.local_rank, rank=self.rank, nranks=self.world_size,
                         unique_id=bytes(uid.cpu().numpy().tobytes()))
        else:
            w = tr.World(params, device=self.local_rank)
        w.set_stream(self.stream.cuda_stream)
        return w


def measure(ctx: Ctx, name: str, n: int | None, steps: int, warmup: int, *, sharded: bool = True, want_e2e: bool = True,
            sample_clocks: bool = False, collisions: bool | None = None) -> dict:
    """W warm-up steps, then K steps each bracketed by CUDA events on the launching stream (L2 flushed
    between steps, outside the events); max over ranks.  Returns the workload's result object."""
    torch, tr = ctx.torch, ctx.tr
    cfg = dict(WORKLOADS[name])
    if collisions is not None:
        cfg["collisions"] = collisions
    nranks = ctx.world_size if sharded else 1
    bodies = make_scene(name, n or cfg["n"])
    n = len(bodies)
    flags = (tr.PAIR_GRAVITY if cfg["gravity"] else 0) | (tr.COLLISIONS if cfg["collisions"] else 0)
    params = tr.default_params(flags=flags, G=np.float32(GRAVITY_G), g=(0, 0, 0))
    world = ctx.make_world(params, sharded)
    world.add(bodies)
    own_first, own_count = world.owned_range()
    world.step(0)
    world.enable_timing(True)
    for _ in range(max(warmup, 0)):
        world.step(1, sync=False)
    world.sync()
    world.reset_stats()

    sampler = ClockSampler(ctx.local_rank) if (sample_clocks and ctx.rank == 0) else None
    if sharded:
        ctx.barrier()
    else:
        torch.cuda.synchronize()
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t_wall0 = time.perf_counter()
    for k in range(steps):
        if ctx.flush_buf is not None:
            ctx.flush_buf.zero_()
        ev[k][0].record(ctx.stream)
        world.step(1, sync=False)
        ev[k][1].record(ctx.stream)
    world.sync()
    if sharded:
        ctx.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    st = world.stats()
    total_ms_max = ctx.max_over_ranks(total_ms) if sharded else total_ms
    ms_per_step = total_ms_max / steps
    pair_visits = float(n) * float(n)          # ordered pairs, self pair included
    res = {
        "workload": f"{name}: {cfg['desc']}" + ("" if collisions is None else (" + replicated grid collisions" if collisions else "")),
        "n_bodies": n, "n_gpus": nranks, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "steps_per_s": 1e3 / ms_per_step, "body_steps_per_s": n * 1e3 / ms_per_step,
        "gpu_launches": int(st.kernel_launches),
        "wall_ms_per_step_incl_flush": t_wall / steps * 1e3,
    }
    if cfg["gravity"]:
        res["value"], res["unit"] = pair_visits * steps / (total_ms_max * 1e-3), "interactions/s"
        res["interaction"] = "ordered (target, source) gravity evaluations, N^2 per step"
    else:
        # no all-pairs work is executed here: the grid prunes the reference's O(N^2) sweep
        res["value"], res["unit"] = pair_visits * steps / (total_ms_max * 1e-3), "equivalent pair visits/s"
        res["interaction"] = ("EQUIVALENT pair visits: what the reference's O(N^2) sweep would visit per step; the GPU does O(N) "
                              "grid work, so compare steps_per_s / body_steps_per_s, not this, with the gravity workloads")
    if clocks is not None:
        res["clocks"] = clocks

    nsteps_t = max(st.broadphase_launches, 1)
    grav_ms = st.gravity_ms / max(st.gravity_launches, 1)
    if cfg["gravity"] and st.gravity_launches:
        peak_ops, peak_clk = tr.measure_fma_peak(ctx.local_rank)
        achieved = float(own_count) * float(n) * LANE_OPS_PER_INTERACTION / (grav_ms * 1e-3)
        res["roofline"] = {
            "bound": "fp32_fma", "kernel": "gravity_kernel", "achieved": achieved / 1e12, "peak": peak_ops / 1e12,
            "unit": "Tlaneop/s", "frac": achieved / peak_ops, "traffic": _profile_json("gravity_traffic.json", str(n)),
            "peak_source": f"measured live: FFMA-only kernel, {peak_clk:.0f} MHz (MEASURED_PEAKS.json has no FP32 row)",
            "frac_of_nominal": achieved / (148 * 128 * 1.965e9),
            "algorithmic": f"{LANE_OPS_PER_INTERACTION} FP32-pipe lane-ops x {own_count} targets x {n} sources per launch",
            "kernel_ms": grav_ms, "share_of_step": grav_ms / ms_per_step,
            "interactions_per_s_kernel": float(own_count) * float(n) / (grav_ms * 1e-3),
        }
    if cfg["collisions"] and st.broadphase_launches:
        hbm, src = hbm_peak()
        broad_ms = st.broadphase_ms / nsteps_t
        narrow_ms = st.narrowphase_ms / nsteps_t
        resp_ms = st.response_ms / nsteps_t
        contacts = int(st.last_contacts)
        ach = BROADPHASE_BYTES_PER_BODY * n / (broad_ms * 1e-3) / 1e9
        ncu = _profile_json("broadphase_ncu.json", f"{name}:{n}") or {}
        bp = {"bound": "hbm", "kernel": "broad phase K2-K3b (keys + histogram, cell-table scan, scatter of 32-byte body records into cell order)",
              "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": ncu.get("dram_bytes"), "peak_source": src,
              "algorithmic": f"{BROADPHASE_BYTES_PER_BODY} B/body x {n} bodies (SURVEY.md 8d: the budget of a 4-pass radix sort + gather)",
              "kernel_ms": broad_ms,
              "stopwatch": "%globaltimer stamps inside the step's CUDA graph: first CTA of K2 -> first CTA of K5 (launch gaps between the four kernels included)",
              "note": "TIME AGAINST A BYTE BUDGET, not achieved DRAM bandwidth: the shipped one-digit counting sort moves fewer "
                      "bytes than the budget it is scored against (see measured_bytes_per_body)"}
        if ncu:
            bp["by_ncu_kernel_sum"] = {"kernel_us_sum": ncu.get("kernel_us_sum"), "frac": (BROADPHASE_BYTES_PER_BODY * n / (ncu["kernel_us_sum"] * 1e-6) / 1e9 / hbm) if ncu.get("kernel_us_sum") else None,
                                       "measured_bytes_per_body": (ncu["dram_bytes"] / n) if ncu.get("dram_bytes") else None,
                                       "source": ncu.get("source")}
        res["roofline_broadphase"] = bp
        nb = (NARROW_BYTES_PER_BODY * n + NARROW_BYTES_PER_CONTACT * contacts)
        res["roofline_narrowphase"] = {
            "bound": "hbm", "kernel": "narrow_kernel (+ side / far lists)", "achieved": nb / (narrow_ms * 1e-3) / 1e9 if narrow_ms > 0 else None,
            "peak": hbm, "unit": "GB/s", "frac": (nb / (narrow_ms * 1e-3) / 1e9 / hbm) if narrow_ms > 0 else None,
            "traffic": (_profile_json("narrowphase_ncu.json", f"{name}:{n}") or {}).get("dram_bytes"),
            "algorithmic": f"{NARROW_BYTES_PER_BODY} B/body x {n} + {NARROW_BYTES_PER_CONTACT} B x {contacts} contacts (SURVEY.md 8d; neighbour reads are cache-served)",
            "kernel_ms": narrow_ms, "peak_source": src}
        rb = RESPONSE_BYTES_PER_CONTACT * contacts
        res["roofline_response"] = {
            "bound": "hbm", "kernel": ("K6s sparse_step_kernel: ordering + ordered replay + step record in one 16-CTA cluster (distributed shared memory)"
                                       if int(st.sparse_steps) > 0 else "K6: contact ordering + dependency build + ordered dataflow replay"), "achieved": rb / (resp_ms * 1e-3) / 1e9 if resp_ms > 0 else None,
            "peak": hbm, "unit": "GB/s", "frac": (rb / (resp_ms * 1e-3) / 1e9 / hbm) if resp_ms > 0 else None, "traffic": None,
            "algorithmic": f"{RESPONSE_BYTES_PER_CONTACT} B/contact x {contacts} (pair 8 + two bodies' pos+vel read 64 and written 64)",
            "kernel_ms": resp_ms, "peak_source": src,
            "note": "latency bound by construction: the reference's Gauss-Seidel order is a dependency chain "
                    f"({int(st.last_response_rounds)} hand-offs on the longest worker / chain through one body in the last step), so the byte fraction is small"}
        res["collision_ms"] = {"broad": broad_ms, "narrow": narrow_ms, "response": resp_ms, "total": broad_ms + narrow_ms + resp_ms,
                               "contacts_last_step": contacts, "response_rounds_last_step": int(st.last_response_rounds),
                               "sparse_steps": int(st.sparse_steps), "sparse_bailouts": int(st.sparse_bailouts)}
        if "roofline" not in res:
            res["roofline"] = bp
    if nranks > 1:
        res["allgather_ms"] = st.allgather_ms / max(steps, 1)

    # ---- e2e: through the C ABI with HOST buffers, host<->device copies inside the timed region, wall clock.
    #   e2e           tr_step_frame: the per-frame call of a device-resident world - the step's per-frame INPUT (external
    #                 forces, 12 B/body) goes up, what the step CHANGES (pos, vel, collider centre, 36 B/body) comes down
    #   e2e_full_aos  tr_step_host: the whole 92-byte descriptor of every body up AND down every step (round 1's e2e)
    if want_e2e:
        def timed(call):
            # every e2e figure starts from the workload's initial state, like the device-timed loop did (the packed
            # lattice of cfg5 is not stationary: the reference's inverted approach test blows it apart after ~30 steps)
            world.write_bodies(bodies)
            call()
            call()                        # (the second call page-locks the caller's buffers in place)
            if sharded:
                ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                call()
            if sharded:
                ctx.barrier()
            dt = time.perf_counter() - t0
            return ctx.max_over_ranks(dt) if sharded else dt

        frame = np.zeros(own_count, dtype=tr.FRAME_DTYPE)
        forces = np.zeros((own_count, 3), dtype=np.float32)
        dt = timed(lambda: world.step_frame(frame, forces, 1))
        res["e2e"] = {"value": pair_visits * steps / dt, "unit": res["unit"],
                      "h2d_bytes_per_step": int(own_count) * 12, "d2h_bytes_per_step": int(own_count) * 36,
                      "ms_per_step": dt / steps * 1e3, "steps_per_s": steps / dt,
                      "api": "tr_step_frame (per-body external force in, pos/vel/centre out, every step)"}
        host = bodies[own_first:own_first + own_count].copy()
        dt = timed(lambda: world.step_host(host, 1))
        res["e2e_full_aos"] = {"value": pair_visits * steps / dt, "unit": res["unit"],
                               "h2d_bytes_per_step": int(own_count) * 92, "d2h_bytes_per_step": int(own_count) * 92,
                               "ms_per_step": dt / steps * 1e3, "steps_per_s": steps / dt,
                               "api": "tr_step_host (AoS tr_body_desc in/out per step)"}
    world.close()
    return res


def shim_frame_time(n: int, frames: int = 6, pack: bool = True):
    """ms per frame through the REAL drop-in path: traceit_headless drives the C++ shim
    traceit_b200::updatePhysics(world&, ofstream&) on reference-shaped 232-byte objects
    (pack -> tr_step_host -> unpack on the host, every frame).  Reference call site: src/main.cpp:1589."""
    exe = os.path.join(ROOT, "traceit_b200", "host", "traceit_headless")
    if not os.path.exists(exe):
        return {"error": "traceit_headless not built"}
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([exe, "--cloud", str(n), "--steps", str(frames), "--no-file", "--time"] + (["--pack"] if pack else []),
                           capture_output=True, text=True, cwd=d, timeout=600)
    if r.returncode != 0:
        return {"error": (r.stderr or r.stdout)[-300:]}
    for ln in r.stdout.splitlines():
        if ln.startswith("TIME "):
            try:
                return json.loads(ln[5:])
            except Exception:
                pass
    return {"error": "no TIME line"}


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
    ctx = Ctx(torch, dist, tr, rank, world_size, local_rank, args.no_flush)

    name = args.workload or ("cfg3" if world_size == 1 else "cfg4")
    default_run = args.workload is None and args.bodies is None
    main_res = measure(ctx, name, args.bodies, args.steps, args.warmup, sharded=True, want_e2e=not args.no_e2e, sample_clocks=True)

    extra = {}
    secondary = None
    if default_run and not args.no_secondary:
        secondary = {}

        def guarded(key, fn):
            try:                       # never lose the headline over a secondary line
                secondary[key] = fn()
            except Exception as e:
                secondary[key] = {"error": f"{type(e).__name__}: {e}"}

        if world_size == 1:
            # every other BASELINE.json configuration that fits one GPU, each as a full object
            guarded("cfg2", lambda: measure(ctx, "cfg2", None, 40, 5, sharded=False, want_e2e=True))
            guarded("cfg5", lambda: measure(ctx, "cfg5", None, 20, 5, sharded=False, want_e2e=True))
            # T1 of SURVEY.md 8d config 4: the 4M cloud on ONE GPU (the denominator of the same-N efficiency)
            guarded("cfg4_t1", lambda: measure(ctx, "cfg4", None, 2, 1, sharded=False, want_e2e=False))
            if rank == 0:
                guarded("e2e_shim", lambda: {"cfg5_1m": shim_frame_time(1 << 20), "sparse_cloud_4k": shim_frame_time(4096, 200, pack=False),
                                             "default_scene": shim_frame_time(0, 500)})
        else:
            # T1 at the SAME N on one GPU (rank 0, unsharded world; the other ranks wait at the barrier)
            if rank == 0:
                guarded("cfg4_t1", lambda: measure(ctx, "cfg4", None, 2, 1, sharded=False, want_e2e=False))
            ctx.barrier()
            t1 = ctx.max_over_ranks(secondary.get("cfg4_t1", {}).get("ms_per_step", 0.0) if rank == 0 else 0.0)
            if t1 > 0 and name == "cfg4":
                extra["efficiency_same_n"] = {"t1_ms": t1, "tr_ms": main_res["ms_per_step"], "ranks": world_size,
                                              "value": t1 / (world_size * main_res["ms_per_step"]),
                                              "definition": "T1 / (R * T_R), both at N = 4,194,304 (SURVEY.md 8d config 4); T1 measured in this run on rank 0's GPU"}
            # the 4M cloud with the REPLICATED collision pipeline on every rank
            guarded("cfg4c", lambda: measure(ctx, "cfg4", None, 3, 1, sharded=True, want_e2e=False, collisions=True))
            c4 = secondary.get("cfg4c", {})
            if "collision_ms" in c4:
                c4["replicated_share_of_step"] = c4["collision_ms"]["total"] / c4["ms_per_step"]
            # parity the driver can see: sharded == unsharded bit for bit, and within tolerance of the CPU oracle
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import sharded_check
            par = {}
            for mode, nn, ss in (("gravity", 20000, 4), ("collisions", 12001, 5)):
                try:
                    par[mode] = sharded_check.run_check(torch, dist, dev, rank, world_size, local_rank,
                                                        "collide" if mode == "collisions" else "gravity", nn, ss)
                except Exception as e:
                    par[mode] = {"ok": False, "error": f"{type(e).__name__}: {e}"}
            extra["sharded_parity"] = par

    # ---- CPU baseline beside it (rank 0, 1 GPU only)
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_step_rate(name, steps=12, warmup=1, extended=True)
        except Exception as e:  # the baseline is a reported figure, never a reason to lose the GPU number
            cpu = {"value": None, "unit": "interactions/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}

    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    cfg = WORKLOADS[name]
    line = {
        "metric": "body-pair interactions/sec" if cfg["gravity"] else "equivalent body-pair visits/sec (see config.interaction) and physics steps/sec",
        "value": main_res["value"], "unit": main_res["unit"], "n_gpus": world_size,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "steps_per_s": main_res["steps_per_s"],
        "higher_is_better": True, "scaling": "strong" if world_size > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": main_res["workload"], "n_bodies": main_res["n_bodies"], "G": GRAVITY_G, "g": [0, 0, 0],
                   "parallelism": f"targets sharded over {world_size} GPU(s), sources replicated",
                   "l2": "flushed between timed steps (256 MB memset, outside the per-step events)" if ctx.flush_buf is not None else "not flushed",
                   "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks",
                   "interaction": main_res["interaction"]},
        "clocks": main_res.get("clocks"), "e2e": main_res.get("e2e"), "e2e_full_aos": main_res.get("e2e_full_aos"),
        "gpu_launches": main_res["gpu_launches"],
        "roofline": main_res.get("roofline"), "cpu_baseline": cpu,
        "wall_ms_per_step_incl_flush": main_res["wall_ms_per_step_incl_flush"], "secondary": secondary,
    }
    for k in ("roofline_broadphase", "roofline_narrowphase", "roofline_response", "collision_ms", "allgather_ms"):
        if k in main_res and not (k == "roofline_broadphase" and main_res.get("roofline") is main_res.get(k)):
            line[k] = main_res[k]
    if cpu and cfg["gravity"] and cpu.get("port_pair_gravity_1core"):
        e2e_v = (main_res.get("e2e") or {}).get("value")
        line["speedup_like_for_like"] = {
            "what": "GPU e2e gravity interactions/s vs the restated pair-gravity pass on the host (the reference's own step has no "
                    "live gravity: vs_reference compares against its O(N^2) collision sweep instead)",
            "vs_1_core": e2e_v / cpu["port_pair_gravity_1core"]["value"] if e2e_v else None,
            "vs_all_cores": e2e_v / cpu["port_pair_gravity"]["value"] if e2e_v else None,
            "cores": cpu["port_pair_gravity"]["cores"]}
    line.update(extra)
    emit(line)


if __name__ == "__main__":
    main()
