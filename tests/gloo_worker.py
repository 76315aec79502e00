"""Worker of tests/test_multirank_cpu.py: one rank of a world_size-2 gloo job that runs the C++ host
layer's stage-parallel integrators (IRKStageParallel / ComplexSPIRK, one stage or conjugate pair per
rank) on the CPU double, with the collectives carried by torch.distributed (gloo)."""
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
from dealii_spirk_b200 import hostapi  # noqa: E402


def main():
    scheme, dim, k, r, q, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cpu = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libspirk_cpu.so"), mode=C.RTLD_GLOBAL)
    AR = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_longlong)
    AG = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_longlong)

    def allreduce(buf, n):
        a = np.ctypeslib.as_array(buf, shape=(n,))
        t = torch.from_numpy(a)
        dist.all_reduce(t)

    def allgather(recv, send, n):
        s = torch.from_numpy(np.ctypeslib.as_array(send, shape=(n,)).copy())
        outs = [torch.empty(n, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(outs, s)
        np.ctypeslib.as_array(recv, shape=(n * world,))[:] = torch.cat(outs).numpy()

    cb = (AR(allreduce), AG(allgather))
    cpu.spirk_cpu_set_comm(rank, world, cb[0], cb[1])
    host = hostapi.HostLib(os.path.join(ROOT, "oracle", "_build", "libspirk_host_cpu.so"), hc.TABLES)
    with hostapi.Run(host, hc.params(scheme, k, r, q), dim=dim, nccl_id=b"\0" * 128, rank=rank, world=world) as run:
        run.setup()
        while not run.finished():
            run.step()
        res = {"u": run.solution().tolist(), "outer": run.array("outer_iterations").tolist(),
               "error_L2": run.array("error_L2").tolist()}
    if rank == 0:
        json.dump(res, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
