"""CPU tests of the multi-GPU host logic: shard arithmetic, and a world_size-2 gloo run of the
sharded step protocol (own targets -> exchange positions -> next step) with the CPU oracle
standing in for the kernels - it must equal the unsharded oracle bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

from traceit_b200 import scenes, sharding


def test_owned_ranges_partition_everything():
    for n in (0, 1, 7, 1000, 4099, 1 << 20):
        for r in (1, 2, 3, 4, 8):
            rs = sharding.all_ranges(n, r)
            assert sum(c for _, c in rs) == n
            pos = 0
            for lo, c in rs:
                assert lo == min(pos, n)
                pos += c
            assert sharding.uses_allgather(n, r) == (n % r == 0 and n > 0 or n == 0)
    assert sharding.owned_range(10, 3, 4) == (9, 1)
    assert sharding.owned_range(4, 3, 8) == (3, 1)
    assert sharding.owned_range(4, 5, 8) == (4, 0)
    with pytest.raises(ValueError):
        sharding.owned_range(10, 4, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist

    from oracle import pyoracle as po
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = scenes.random_cloud(n, 99, spheres=True)
    p = po.default_params(flags=po.PAIR_GRAVITY, G=np.float32(1e-2), g=(0, 0, 0))
    lo, cnt = sharding.owned_range(n, rank, world)
    per = (n + world - 1) // world
    for _ in range(steps):
        # each rank: forces and integration for its own targets only (all sources visible)
        f = po.pair_gravity_f32(b, p.G)
        mine = b[lo:lo + cnt].copy()
        po.integrate(mine, p, f[lo:lo + cnt])
        b[lo:lo + cnt] = mine
        # exchange: padded all-gather of the owned blocks' positions (pos is what sources need)
        send = np.zeros((per, 3), dtype=np.float32)
        send[:cnt] = b["pos"][lo:lo + cnt]
        recv = [torch.zeros(per, 3) for _ in range(world)]
        dist.all_gather(recv, torch.from_numpy(send))
        for r, (rlo, rc) in enumerate(sharding.all_ranges(n, world)):
            b["pos"][rlo:rlo + rc] = recv[r].numpy()[:rc]
    np.save(os.path.join(out_dir, f"pos_{rank}.npy"), b["pos"])
    np.save(os.path.join(out_dir, f"vel_{rank}.npy"), b["vel"][lo:lo + cnt])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [600, 601])
def test_sharded_protocol_world_size_2_gloo(tmp_path, oracle, n):
    import torch.multiprocessing as mp
    steps = 3
    port = _free_port()
    mp.start_processes(_worker, args=(2, port, n, steps, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    ref = scenes.random_cloud(n, 99, spheres=True)
    oracle.step(ref, oracle.default_params(flags=oracle.PAIR_GRAVITY, G=np.float32(1e-2), g=(0, 0, 0)), steps)
    for r in range(2):
        pos = np.load(tmp_path / f"pos_{r}.npy")
        assert np.array_equal(pos.view(np.uint32), ref["pos"].view(np.uint32)), f"rank {r} positions"
        lo, cnt = sharding.owned_range(n, r, 2)
        vel = np.load(tmp_path / f"vel_{r}.npy")
        assert np.array_equal(vel.view(np.uint32), ref["vel"][lo:lo + cnt].view(np.uint32))
