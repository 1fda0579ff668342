"""Host-side shard arithmetic for sharded worlds (one process per GPU).

Mirrors set_owned_range() in traceit_b200/csrc/api.cu: contiguous blocks of ceil(N/R)
targets in creation order, so body ids stay stable and the ragged tail lands on the last
non-empty rank.  Every rank holds all N source positions; only targets are partitioned.
"""
from __future__ import annotations


def owned_range(n: int, rank: int, nranks: int) -> tuple[int, int]:
    """(first, count) of the targets rank `rank` integrates."""
    if nranks < 1 or not (0 <= rank < nranks):
        raise ValueError("bad rank / nranks")
    per = (n + nranks - 1) // nranks
    lo = min(per * rank, n)
    hi = min(lo + per, n)
    return lo, hi - lo


def all_ranges(n: int, nranks: int) -> list[tuple[int, int]]:
    return [owned_range(n, r, nranks) for r in range(nranks)]


def uses_allgather(n: int, nranks: int) -> bool:
    """True when every block has the same size (one ncclAllGather); otherwise the exchange is
    a group of per-rank broadcasts (nccl_allgather_pos in nccl_dyn.cu)."""
    per = (n + nranks - 1) // nranks
    return per * nranks == n
