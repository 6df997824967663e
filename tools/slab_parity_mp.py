"""N slab processes (one per GPU) against one context, through the real transport.  Launch with torch.distributed.run:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/slab_parity_mp.py --grid 64 --steps 6 --exchange peer

Rank 0 prints one JSON line (fluidsolvergpu_b200.slab.parity_against_single for the symmetric and the gather kernel)."""
import argparse
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=64)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--drift", type=float, default=25.0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import fluidsolvergpu_b200 as fsg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world, "exchange": args.exchange}
    for mode, name in ((0, "symmetric_kernel"), (1, "gather_kernel")):
        rec = fsg.slab.parity_against_single(args.grid, rank, world, local, exchange=args.exchange, steps=args.steps, pair_mode=mode,
                                             drift=args.drift)
        if rank == 0:
            out[name] = rec
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
