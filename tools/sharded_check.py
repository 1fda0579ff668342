"""torchrun worker: sharded world over NCCL vs an unsharded world on the same GPU, bitwise."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import traceit_b200 as tr  # noqa: E402
from traceit_b200 import scenes, sharding  # noqa: E402

n, steps = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "gravity"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid = torch.frombuffer(bytearray(tr.comm_unique_id()), dtype=torch.uint8).to(dev)
dist.broadcast(uid, 0)
if mode == "collide":
    # dense cloud: gravity + drag + integration + sphere collisions, replicated collision pipeline
    b = scenes.random_cloud(n, 7, spheres=True, density_l=0.6)
    b["stationary"][::37] = 1
    b["isStatic"][::11] = 1
    p = tr.default_params(flags=tr.PAIR_GRAVITY | tr.COLLISIONS, G=np.float32(1e-3))
else:
    b = scenes.random_cloud(n, 7, spheres=True)
    p = tr.default_params(flags=tr.PAIR_GRAVITY, G=np.float32(1e-2), g=(0, 0, 0))
with tr.World(p, device=local, rank=rank, nranks=world, unique_id=uid.cpu().numpy().tobytes()) as w:
    w.add(b)
    lo, cnt = w.owned_range()
    assert (lo, cnt) == sharding.owned_range(n, rank, world)
    w.step(steps)
    mine = w.read_bodies(lo, cnt)
    pos_all, _ = w.read_state(0, n, vel=False)
    contacts = w.contacts() if mode == "collide" else None
    # host-buffer path on the shard
    w.step_host(mine.copy(), 1)
with tr.World(p, device=local) as w1:
    w1.add(b)
    w1.step(steps)
    ref = w1.read_bodies()
    ref_contacts = w1.contacts() if mode == "collide" else None
ok = True
if mode == "collide":
    ok &= np.array_equal(contacts, ref_contacts) and len(contacts) > 0
for f in ("pos", "vel", "center"):
    ok &= np.array_equal(mine[f].view(np.uint32), ref[f][lo:lo + cnt].view(np.uint32))
ok &= np.array_equal(pos_all.view(np.uint32), ref["pos"].view(np.uint32))
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("SHARDED_OK" if int(t.item()) == 1 else "SHARDED_MISMATCH")
sys.exit(0 if int(t.item()) == 1 else 1)
