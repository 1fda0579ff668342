"""Collision pipeline only (no pair gravity) on the cfg3 cloud / cfg5 pack: per-phase device times.

    python tools/collide_bench.py [cfg3|cfg5] [N] [steps]

L2 is flushed (256 MB memset) before every step, as in bench.py.  With TR_BP_PROFILE=1 the library
prints the broad phase kernel by kernel (plain launches instead of the CUDA graph)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import traceit_b200 as tr
from traceit_b200 import scenes

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
b = scenes.cfg5_dense(n) if name == "cfg5" else scenes.cfg3_cloud(n)
p = tr.default_params(flags=tr.COLLISIONS, g=(0, 0, 0), response_mode=int(os.environ.get("TR_RESPONSE_MODE", "0")))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with tr.World(p) as w:
    w.set_stream(stream.cuda_stream)
    w.add(b)
    w.enable_timing(True)
    for _ in range(3):
        w.step(1)
    w.reset_stats()
    for _ in range(steps):
        flush.zero_()
        w.step(1)
    st = w.stats()
    print(f"{name} n={len(b)} steps={steps}: broad {st.broadphase_ms / steps * 1e3:.1f} us "
          f"({140.0 * len(b) / (st.broadphase_ms / steps * 1e-3) / 1e9:.0f} GB/s algorithmic), "
          f"narrow {st.narrowphase_ms / steps * 1e3:.1f} us, response {st.response_ms / steps * 1e3:.1f} us, "
          f"contacts {st.last_contacts}, side {st.side_list_size}, rounds {st.last_response_rounds}")
