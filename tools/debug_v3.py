#!/usr/bin/env python3
"""Debug helper: where does the fast path differ from the general kernel?"""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dealii_spirk_b200 as pkg
from dealii_spirk_b200 import capi
import abi_checks as ac

dev = pkg.device_lib()
for r, nb in [(3, 1), (3, 4), (4, 1)]:
    lvl, olv = ac.make_level(3, 4, r)
    u = ac.block_input(olv, nb, seed=1)
    op = capi.real_op([16.0, 3.1, 2.9, 5.6][:nb], [0.1])
    with capi.Context(dev) as ctx:
        src, dst = ctx.upload(u), ctx.alloc(u.size)
        outs = {}
        for v in (1, 0):
            ctx.call("spirk_ctx_set_option", b"apply_variant", v)
            ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), dst, src, olv.N)
            outs[v] = ctx.download(dst, u.shape)
    d = np.abs(outs[0] - outs[1])
    print(f"r={r} nb={nb}: max abs diff {d.max():.3e} (ref max {np.abs(outs[1]).max():.3e}), n bad {np.sum(d > 1e-10)}")
    bad = np.argwhere(d > 1e-10)
    if len(bad):
        print(" first bad (b,z,y,x):", bad[:12].tolist())
        for ax, name in enumerate("bzyx"):
            vals, cnt = np.unique(bad[:, ax], return_counts=True)
            print(f"  {name}: ", dict(zip(vals.tolist()[:40], cnt.tolist()[:40])))
