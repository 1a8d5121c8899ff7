"""One synthetic norm argument sharded over the GPUs of a box: bppp_nl_prove_sharded with NCCL inside the
library.  One rank per GPU under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port 29533 \
        tools/sweep_sharded.py [e ...]

Every rank derives the same generators and witness, keeps its contiguous slice, proves, and checks the proof against the
unsharded device proof bit for bit.  Rank 0 prints one JSON line per size (times: max over ranks, CUDA events)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bulletproofspp_b200 as bp
from bulletproofspp_b200 import sweep


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [16, 20]
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    ctx = bp.Context(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for e in sizes:
        res = sweep.run_sharded(ctx, e, rank, world, dist if world > 1 else None)
        if rank == 0:
            print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
