"""Sharded synthetic norm argument under torchrun (one rank per GPU): N = 2^e norm elements split
contiguously over the ranks; per round 128 bytes per rank are all-gathered (NCCL) and added.
    python -m torch.distributed.run --nproc-per-node W tools/sweep_sharded.py e
Prints one JSON line on rank 0 (prove time, max over ranks)."""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import torch.distributed as dist
import bulletproofspp_b200 as bp
from bulletproofspp_b200 import lib as L
from bulletproofspp_b200.sharded import Shard, prove_sharded
from oracle.field import R
import sweep

e = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = bp.Context(local)
N, M, k = 1 << e, 6, e - 2
Ls = N // world
# every rank derives the same generators / witness deterministically and keeps its slice
pts = sweep.generators(ctx, 1 + N + M)
g = L.bytes_to_point(pts[:64])
Gs = L.bytes_to_points(pts[64 * (1 + rank * Ls):64 * (1 + (rank + 1) * Ls)])
Hs = L.bytes_to_points(pts[64 * (1 + N):])
q = L.le_to_int(sweep.scalars("q%d" % e, 1))
w = L.bytes_to_ints(sweep.scalars("w%d" % e, N)[32 * rank * Ls:32 * (rank + 1) * Ls])
l, c = L.bytes_to_ints(sweep.scalars("l%d" % e, M)), L.bytes_to_ints(sweep.scalars("c%d" % e, M))


_buf = torch.empty(128, dtype=torch.uint8, device="cuda")
_all = torch.empty(world * 128, dtype=torch.uint8, device="cuda")


def gather(vals):
    """per-round partial commitments: one 128-byte CUDA tensor per rank through ncclAllGather;
    the (rare) state hand-offs go through all_gather_object"""
    if world == 1:
        return vals
    v = vals[0]
    if isinstance(v, tuple) and len(v) == 2 and len(v[0]) == 1 and (v[0][0] is None or isinstance(v[0][0], tuple)):
        payload = L.point_to_bytes(v[0][0]) + L.point_to_bytes(v[1][0])
        _buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
        dist.all_gather_into_tensor(_all, _buf)
        raw = bytes(_all.cpu().numpy())
        return [([L.bytes_to_point(raw[128 * r:128 * r + 64])], [L.bytes_to_point(raw[128 * r + 64:128 * r + 128])])
                for r in range(world)]
    out = [None] * world
    dist.all_gather_object(out, v)
    return out


def oracle(X, Rr):
    return int.from_bytes(hashlib.sha256(L.point_to_bytes(X) + L.point_to_bytes(Rr)).digest(), "big") % R


sh = Shard(ctx, rank, world, N, g, Gs, Hs, q, 12345, w, l, c)
if world > 1:
    dist.barrier()
ctx.sync()
t0 = time.time()
round_t = []
_orc = oracle


def oracle(X, Rr):
    round_t.append(time.time())
    return _orc(X, Rr)


resp, s_fin, fw, fl = prove_sharded([sh], gather, k, oracle, q, M)
ctx.sync()
dt = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    chk = hashlib.sha256(repr((resp, s_fin, fw, fl)).encode()).hexdigest()[:16]
    print(json.dumps({"workload": "sharded norm argument prove", "e": e, "N": N, "M": M, "rounds": k, "n_gpus": world,
                      "prove_s": round(dt.item(), 4), "proof_checksum": chk, "final": [len(fw), len(fl)],
                      "round_ms": [round((b - a) * 1e3, 2) for a, b in zip([t0] + round_t[:-1], round_t)]}), flush=True)
if world > 1:
    dist.destroy_process_group()
