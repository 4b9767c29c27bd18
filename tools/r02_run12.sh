#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --tb=short 2>&1 | tail -30
exit 0
python - <<'PY'
# e2e of one 1024^3 volume through jwc_fwt3d: one device vs the device group
import time, ctypes as C, numpy as np, torch
import jwave_b200 as jw
from jwave_b200.transforms import CudaContext
n, lv = 1024, 10
nd = torch.cuda.device_count()
for devs in ([0], list(range(nd))):
    ctx = CudaContext(devs if len(devs) > 1 else 0)
    t = jw.CudaFastWaveletTransform(jw.WaveletBuilder.create("Coiflet5"), context=ctx)
    hx = torch.randn(n, n, n, dtype=torch.float64).pin_memory()
    hc = torch.empty_like(hx).pin_memory(); hb = torch.empty_like(hx).pin_memory()
    L = ctx._lib
    def step():
        ctx.check(L.jwc_fwt3d(ctx.handle, t._wid, 0, hx.data_ptr(), hc.data_ptr(), n, n, n, lv, lv, lv), "f")
        ctx.check(L.jwc_fwt3d(ctx.handle, t._wid, 1, hc.data_ptr(), hb.data_ptr(), n, n, n, lv, lv, lv), "r")
    step()
    t0 = time.perf_counter(); step(); step(); dt = (time.perf_counter() - t0) / 2
    print(f"jwc_fwt3d e2e 1024^3 on {len(devs)} device(s): {2 * n**3 / dt * 1e-9:.2f} GS/s, {dt * 1e3:.0f} ms per forward+reverse, round trip err {float((hb - hx).abs().max()):.2e}", flush=True)
    del hx, hc, hb
    ctx.close()
PY
