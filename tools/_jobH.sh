python -m pytest tests/test_gpu_abi.py -x -q -k "transfer or km" 2>&1 | tail -3
python - <<'PY'
import sys, time, ctypes as C, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'oracle'); sys.path.insert(0,'.')
import dealii_spirk_b200 as pkg
from dealii_spirk_b200 import capi
dev = pkg.device_lib()
for r in (4,5,6):
    lvl = capi.Level(3,4,2**r,0); lc = capi.Level(3,4,2**(r-1),0)
    nb=2
    with capi.Context(dev) as ctx:
        f = ctx.alloc(nb*lvl.n_dofs); c = ctx.alloc(nb*lc.n_dofs)
        ctx.call("spirk_vec_set", f, nb*lvl.n_dofs, 0.5)
        for name, fn in (("restrict", lambda: ctx.call("spirk_mg_restrict", C.byref(lvl), nb, c, lc.n_dofs, f, lvl.n_dofs)),
                         ("prolongate", lambda: ctx.call("spirk_mg_prolongate_add", C.byref(lvl), nb, f, lvl.n_dofs, c, lc.n_dofs))):
            for _ in range(3): fn()
            ctx.call("spirk_ctx_timer_begin")
            for _ in range(10): fn()
            ms = ctx.scalar_call("spirk_ctx_timer_end")/10
            print(f"transfer r={r} nb={nb} {name}: {ms*1e3:.1f} us", flush=True)
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-scaling-reference | tail -c 900
