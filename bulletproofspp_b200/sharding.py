"""Multi-GPU partitioning helpers (host side; one process per GPU).

* Batches of independent proofs shard with no data-path collective: rank r proves/verifies the
  proofs [proof_offset(r, batch_per_gpu), ...) -- weak scaling.
* One large argument shards its vectors contiguously (the fold pairs ADJACENT elements,
  src/Bulletproof.hs:77-90, so a contiguous shard folds locally); each round's commitment is the
  sum of per-rank partial MSMs.  EC addition is not an NCCL reduction op: gather, then add.
"""


def shard_range(n, rank, world):
    """contiguous [lo, hi) of n items for `rank`"""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def proof_offset(rank, batch_per_gpu):
    return rank * batch_per_gpu


def _point_to_tensor(p):
    import torch
    b = bytes(64) if p is None else int(p[0]).to_bytes(32, "little") + int(p[1]).to_bytes(32, "little")
    return torch.frombuffer(bytearray(b), dtype=torch.uint8).clone()


def _tensor_to_point(t):
    b = bytes(t.cpu().tolist())
    x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:], "little")
    return None if x == 0 and y == 0 else (x, y)


def combine_partials(partial, group, dist, device=None, add=None):
    """all-gather one partial point per rank (64 bytes each) and add them locally.  `group` supplies
    `.add`; on GPUs the caller passes `add` bound to the device group law."""
    import torch
    t = _point_to_tensor(partial)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    acc = None
    f = add or group.add
    for o in outs:
        acc = f(acc, _tensor_to_point(o))
    return acc
