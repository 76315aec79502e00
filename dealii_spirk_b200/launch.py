"""Launcher: `python -m dealii_spirk_b200.launch [--dim 3] input_0.json [input_1.json ...]`

Single process: the reference's command line (main.cc:3608-3791).  Under torchrun
(`python -m torch.distributed.run --nproc-per-node N ... -m dealii_spirk_b200.launch ...`) every rank
drives one GPU; the NCCL unique id of the C++ layer's communicator is exchanged with
torch.distributed and the ranks form the stage ("row") communicator of the reference's virtual
topology (main.cc:3660-3698).  With one GPU per stage the column (space) communicator has size 1.
"""
import argparse
import ctypes as C
import json
import os
import sys


def lex_to_pair(rank, size1, size2, do_row_major):
    """(row index, column index) of a rank in the size1 x size2 process grid (reference main.cc:281-293)."""
    if rank >= size1 * size2:
        raise ValueError("Invalid rank.")
    return (rank % size1, rank // size1) if do_row_major else (rank // size2, rank % size2)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("inputs", nargs="+")
    a = ap.parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import dealii_spirk_b200 as pkg
    from dealii_spirk_b200 import hostapi
    dev = pkg.device_lib()
    host = hostapi.HostLib(pkg.HOST_LIB_PATH, pkg.TABLES_PATH)
    nccl_id = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        buf = C.create_string_buffer(128)
        if rank == 0:
            dev.call("spirk_comm_unique_id", buf)
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        nccl_id = bytes(t.cpu().numpy().tobytes())
    if rank == 0:
        print(f"Running in {a.dim}D on backend {host.backend()} with {world} rank(s)")
    for path in a.inputs:
        params = json.load(open(path))
        if rank == 0:
            print(f"\nProcessing {path}")
        with hostapi.Run(host, params, dim=a.dim, device=local_rank, nccl_id=nccl_id, rank=rank, world=world, verbose=True) as run:
            run.run()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
