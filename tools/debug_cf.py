#!/usr/bin/env python3
"""Debug helper: spirk_op_cheb_first fast path (variant 0) against the general path (variant 1)."""
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
for r, nb in [(4, 2), (5, 1), (5, 4), (6, 1)]:
    lvl, olv = ac.make_level(3, 4, r)
    b = ac.block_input(olv, nb, seed=6)
    op = capi.real_op([16.0, 3.1, 2.9, 5.6][:nb], [0.1])
    f0, _0 = capi.darr([0.7, 0.6, 0.65, 0.75][:nb])
    f1, _1 = capi.darr([0.3, 0.2, 0.25, 0.35][:nb])
    f2, _2 = capi.darr([1.1, 0.9, 1.0, 1.2][:nb])
    outs = {}
    with capi.Context(dev) as ctx:
        db = ctx.upload(b)
        x1, x2 = ctx.alloc(b.size), ctx.alloc(b.size)
        for v in (1, 0):
            ctx.call("spirk_ctx_set_option", b"apply_variant", v)
            ctx.call("spirk_op_cheb_first", C.byref(lvl), C.byref(op), x1, x2, db, olv.N, f0, f1, f2)
            outs[v] = (ctx.download(x1, b.shape), ctx.download(x2, b.shape))
    for which in (0, 1):
        d = np.abs(outs[0][which] - outs[1][which])
        ref = np.abs(outs[1][which]).max()
        bad = np.argwhere(d > 1e-10 * ref)
        print(f"r={r} nb={nb} x{which + 1}: max abs diff {d.max():.3e} (ref {ref:.3e}), n bad {len(bad)}")
        if len(bad):
            print("  first bad (b,z,y,x):", bad[:8].tolist())
            for ax, name in enumerate("bzyx"):
                vals, cnt = np.unique(bad[:, ax], return_counts=True)
                print(f"   {name}: ", dict(zip(vals.tolist()[:24], cnt.tolist()[:24])))
