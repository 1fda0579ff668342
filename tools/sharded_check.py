"""Sharded-world parity: R ranks over NCCL vs (a) an unsharded world on the same GPU, bit for bit, and
(b) the CPU oracle (restated C step) on the same initial conditions.

Two uses:
  * torchrun worker (tests/test_gpu_sharded.py):
        python -m torch.distributed.run --nproc-per-node 2 tools/sharded_check.py N STEPS [gravity|collide|grow|growc]
  * imported by bench.py at WORLD_SIZE > 1: `run_check(...)` after the timed region, so that every
    SCALE line carries a `sharded_parity` object the driver can see.

The oracle is used here only as the checker (rank 0 runs it, the result is broadcast).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOL = 1e-4   # max |error| / (cloud extent | rms speed) vs the fp32 restated oracle


def _scene(mode, n):
    import traceit_b200 as tr
    from traceit_b200 import scenes
    if mode == "grow":
        # collisions OFF (only positions are exchanged per step) and a pending external force on every body:
        # a new owner that continued from upload-time velocity / force / centre would be visibly wrong
        b = scenes.random_cloud(n, 7, spheres=True)
        b["force"] = np.random.RandomState(5).uniform(-50, 50, size=(n, 3)).astype(np.float32)
        b["stationary"][::29] = 1
        p = tr.default_params(flags=tr.PAIR_GRAVITY, G=np.float32(1e-2))
    elif mode in ("collide", "growc"):
        # dense cloud: gravity + drag + integration + sphere collisions, replicated collision pipeline
        b = scenes.random_cloud(n, 7, spheres=True, density_l=0.6)
        b["stationary"][::37] = 1
        b["isStatic"][::11] = 1
        p = tr.default_params(flags=tr.PAIR_GRAVITY | tr.COLLISIONS, G=np.float32(1e-3))
    else:
        b = scenes.random_cloud(n, 7, spheres=True)
        p = tr.default_params(flags=tr.PAIR_GRAVITY, G=np.float32(1e-2), g=(0, 0, 0))
    return b, p


def _oracle(mode, b, p, steps):
    """Restated CPU step from the same initial conditions: final state + the last step's contact set."""
    from oracle import pyoracle as po
    want = b.copy()
    pg = po.default_params(flags=po.PAIR_GRAVITY, G=np.float32(p.G), g=tuple(p.g), dt=np.float32(p.dt),
                           atm_density=np.float32(p.atm_density))
    contacts = np.zeros((0, 2), dtype=np.int32)
    for _ in range(steps):
        po.step(want, pg, 1)                      # pair gravity -> drag + integrate
        if mode in ("collide", "growc"):
            contacts = po.contacts(want)          # brute force, parallel over i
            po.resolve_list(want, contacts)       # == the sequential O(N^2) sweep, bit for bit
    return want, contacts


def run_check(torch, dist, dev, rank, world, local, mode="gravity", n=20000, steps=4, with_oracle=True):
    """Returns a dict (identical on every rank)."""
    import traceit_b200 as tr
    from traceit_b200 import sharding

    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.frombuffer(bytearray(tr.comm_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(uid, 0)
    b, p = _scene(mode, n)
    growing = mode in ("grow", "growc")
    has_contacts = mode in ("collide", "growc")
    grow_at = steps // 2 if growing else None
    n0 = n - n // 3 if growing else n                 # "grow": the last third is added mid-run
    with tr.World(p, device=local, rank=rank, nranks=world, unique_id=uid.cpu().numpy().tobytes()) as w:
        w.add(b[:n0])
        if grow_at is None:
            lo, cnt = w.owned_range()
            assert (lo, cnt) == sharding.owned_range(n, rank, world)
            w.step(steps)
        else:
            w.step(grow_at)
            w.add(b[n0:])                              # ownership of most bodies moves to another rank here
            w.step(steps - grow_at)
            lo, cnt = w.owned_range()
        mine = w.read_bodies(lo, cnt)
        pos_all, _ = w.read_state(0, n, vel=False)
        contacts = w.contacts() if has_contacts else None
        w.step_host(mine.copy(), 1)                    # host-buffer path on the shard
    with tr.World(p, device=local) as w1:
        w1.add(b[:n0])
        if grow_at is None:
            w1.step(steps)
        else:
            w1.step(grow_at)
            w1.add(b[n0:])
            w1.step(steps - grow_at)
        ref = w1.read_bodies()
        ref_contacts = w1.contacts() if has_contacts else None
    same = True
    if has_contacts:
        same &= bool(np.array_equal(contacts, ref_contacts) and len(contacts) > 0)
    for f in ("pos", "vel", "center"):
        same &= bool(np.array_equal(mine[f].view(np.uint32), ref[f][lo:lo + cnt].view(np.uint32)))
    same &= bool(np.array_equal(pos_all.view(np.uint32), ref["pos"].view(np.uint32)))

    out = {"n_bodies": n, "steps": steps, "ranks": world}
    err = torch.zeros(3, dtype=torch.float64, device=dev)   # [pos err, vel err, contact-set mismatch]
    if with_oracle and not growing:
        # rank 0 runs the CPU oracle once; every rank compares its OWNED block against it
        want_pv = torch.zeros((n, 6), dtype=torch.float32, device=dev)
        ncon = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            want, wc = _oracle(mode, b, p, steps)
            want_pv = torch.from_numpy(np.concatenate([want["pos"], want["vel"]], axis=1)).to(dev)
            ncon[0] = len(wc)
        dist.broadcast(want_pv, 0)
        dist.broadcast(ncon, 0)
        wc_t = torch.zeros((int(ncon.item()), 2), dtype=torch.int32, device=dev)
        if rank == 0 and len(wc):
            wc_t = torch.from_numpy(wc).to(dev)
        if int(ncon.item()):
            dist.broadcast(wc_t, 0)
        wp = want_pv.cpu().numpy()
        ok = np.isfinite(wp).all(axis=1)
        extent = float(np.abs(wp[ok, :3]).max())
        vrms = float(np.sqrt((wp[ok, 3:].astype(np.float64) ** 2).sum(axis=1).mean()))
        sl = slice(lo, lo + cnt)
        okb = ok[sl]
        err[0] = float(np.abs(mine["pos"][okb].astype(np.float64) - wp[sl, :3][okb]).max() / extent) if okb.any() else 0.0
        err[1] = float(np.abs(mine["vel"][okb].astype(np.float64) - wp[sl, 3:][okb]).max() / vrms) if okb.any() else 0.0
        if has_contacts:
            err[2] = 0.0 if np.array_equal(contacts, wc_t.cpu().numpy()) else 1.0
    t = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    out["vs_unsharded_same_gpu"] = "bit-identical" if t.item() == 1.0 else "MISMATCH"
    if with_oracle and not growing:
        out["vs_cpu_oracle"] = {"max_pos_err_over_extent": float(err[0].item()), "max_vel_err_over_vrms": float(err[1].item()),
                                "tolerance": TOL, "oracle": "oracle/restated.c fp32 step, owned block of every rank"}
        if has_contacts:
            out["vs_cpu_oracle"]["last_step_contact_set"] = "equal" if err[2].item() == 0.0 else "DIFFERS"
            out["contacts_last_step"] = int(len(contacts))
        out["ok"] = bool(t.item() == 1.0 and err[0].item() <= TOL and err[1].item() <= TOL and err[2].item() == 0.0)
    else:
        out["ok"] = bool(t.item() == 1.0)
    return out


def main():
    import torch
    import torch.distributed as dist
    n, steps = int(sys.argv[1]), int(sys.argv[2])
    mode = sys.argv[3] if len(sys.argv) > 3 else "gravity"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    r = run_check(torch, dist, dev, rank, world, local, mode, n, steps)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(r)
        print("SHARDED_OK" if r["ok"] else "SHARDED_MISMATCH")
    sys.exit(0 if r["ok"] else 1)


if __name__ == "__main__":
    main()
