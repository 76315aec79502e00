"""Worker of tests/test_gpu_multirank.py: one rank (= one GPU) of a stage-parallel run of the PRODUCT stack
(libspirk_b200.so + libspirk_host.so) under torchrun: NCCL communicator of the C++ layer, peer-mapped exchange buffers,
fused mixing kernels, all-reduced Krylov scalars (reference main.cc:1229-1760, 2382-2934)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import host_checks as hc  # noqa: E402
import dealii_spirk_b200 as pkg  # noqa: E402
from dealii_spirk_b200 import hostapi  # noqa: E402


def main():
    scheme, dim, k, r, q, tol, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), \
        float(sys.argv[6]), sys.argv[7]
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = pkg.device_lib()
    host = hostapi.HostLib(pkg.HOST_LIB_PATH, hc.TABLES)
    assert host.backend() == "cuda-sm_100a"
    buf = C.create_string_buffer(128)
    if rank == 0:
        dev.call("spirk_comm_unique_id", buf)
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    nccl_id = bytes(t.cpu().numpy().tobytes())
    with hostapi.Run(host, hc.params(scheme, k, r, q, tol=tol), dim=dim, device=local_rank, nccl_id=nccl_id, rank=rank,
                     world=world) as run:
        run.setup()
        while not run.finished():
            run.step()
        run.finish()
        # the solution of the whole mesh: with a spatial partition the ranks of stage 0 hold the z-slabs, in rank order
        u = None
        if r <= 4:
            mine = torch.from_numpy(run.solution()).cuda()
            first = int(run.scalar("first_owned"))
            info = torch.tensor([first, mine.numel()], dtype=torch.int64, device="cuda")
            infos = [torch.zeros_like(info) for _ in range(world)]
            dist.all_gather(infos, info)
            nmax = max(int(i[1]) for i in infos)
            pad = torch.zeros(nmax, dtype=torch.float64, device="cuda")
            pad[:mine.numel()] = mine
            parts = [torch.zeros_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad)
            full = np.zeros(int(run.scalar("n_dofs")))
            for i, p in zip(infos, parts):
                full[int(i[0]):int(i[0]) + int(i[1])] = p[:int(i[1])].cpu().numpy()
            u = full.tolist()
        res = {"u": u, "outer": run.array("outer_iterations").tolist(),
               "error_L2": run.array("error_L2").tolist(), "error_Linf": run.array("error_Linf").tolist(),
               "norm": run.array("solution_l2").tolist(), "launches": run.scalar("launch_count")}
    if rank == 0:
        json.dump(res, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
