python -m pytest tests -m gpu -x -q 2>&1 | tail -6
K="--kernels apply cheb_step_own_dinv residual"
B="python tools/bench_vmult.py --variants 0 --reps 20 $K"
$B --refine 6 --tag dyn_auto --nb 1
$B --refine 6 --tag dyn_auto --nb 2
$B --refine 6 --tag dyn4 --nb 2 --opt v3_chunk=4
$B --refine 6 --tag dyn_npt2 --nb 2 --opt v3_npt=2
$B --refine 6 --tag dyn_npt2 --nb 1 --opt v3_npt=2
$B --refine 5 --tag dyn_auto --nb 2
$B --refine 4 --tag dyn_auto --nb 2
$B --refine 7 --tag dyn_auto --nb 1 --reps 10
python - <<'PY'
import sys, time, ctypes as C, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'oracle'); sys.path.insert(0,'.')
import dealii_spirk_b200 as pkg
from dealii_spirk_b200 import capi
dev = pkg.device_lib()
for r in (5,6):
    lvl = capi.Level(3,4,2**r,0); lc = capi.Level(3,4,2**(r-1),0)
    nb=2
    with capi.Context(dev) as ctx:
        f = ctx.alloc(nb*lvl.n_dofs); c = ctx.alloc(nb*lc.n_dofs); d = ctx.alloc(nb*lvl.n_dofs); w = ctx.alloc(nb*lvl.n_dofs)
        ctx.call("spirk_vec_set", f, nb*lvl.n_dofs, 0.5)
        for variant in (1,0):
            ctx.call("spirk_ctx_set_option", b"transfer_variant", variant)
            for name, fn in (("restrict", lambda: ctx.call("spirk_mg_restrict", C.byref(lvl), nb, c, lc.n_dofs, f, lvl.n_dofs)),
                             ("prolongate", lambda: ctx.call("spirk_mg_prolongate_add", C.byref(lvl), nb, f, lvl.n_dofs, c, lc.n_dofs))):
                for _ in range(3): fn()
                ctx.call("spirk_ctx_timer_begin")
                for _ in range(10): fn()
                ms = ctx.scalar_call("spirk_ctx_timer_end")/10
                print(f"transfer r={r} nb={nb} variant={variant} {name}: {ms*1e3:.1f} us", flush=True)
        pl,_1 = capi.darr([0.1]*nb); pm,_2 = capi.darr([1.0]*nb)
        fn = lambda: ctx.call("spirk_op_apply_km", C.byref(lvl), nb, d, f, w, lvl.n_dofs, pl, pm)
        for _ in range(3): fn()
        ctx.call("spirk_ctx_timer_begin")
        for _ in range(10): fn()
        ms = ctx.scalar_call("spirk_ctx_timer_end")/10
        print(f"apply_km r={r} nb={nb}: {ms*1e3:.1f} us = {nb*lvl.n_dofs/ms*1e-6:.1f} GDoF/s, {24*nb*lvl.n_dofs/ms*1e-6:.0f} GB/s", flush=True)
PY
python bench.py --steps 6 --warmup 3 --no-cpu-baseline
