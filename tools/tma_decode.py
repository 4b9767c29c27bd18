"""Decode which (row, column) the TMA-staged strided kernel actually reads: Haar level 1 on
x[s][c] = 16 s + c gives a, d from which both source samples can be recovered exactly."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jwave_b200 as jw
from jwave_b200 import _lib
from jwave_b200.device import DeviceTransforms
n, inner = 512, 8
x = (16.0 * np.arange(n)[:, None] + np.arange(inner)[None, :])
dev = DeviceTransforms(jw.WaveletBuilder.create("Haar"))
y = dev.axis(_lib.FWT, _lib.FORWARD, torch.from_numpy(x).cuda().reshape(1, n, inner), 1, n, inner, 1).cpu().numpy().reshape(n, inner)
a, d = y[: n // 2] * math.sqrt(2), y[n // 2:] * math.sqrt(2)
x1, x2 = (a + d) / 2, (a - d) / 2   # samples read as x[2i] and x[2i+1]
bad = 0
for i in range(n // 2):
    for c in range(inner):
        for k, v in enumerate((x1[i, c], x2[i, c])):
            s_read, c_read = int(round(v)) // 16, int(round(v)) % 16
            if (s_read, c_read) != (2 * i + k, c):
                bad += 1
                if bad <= 40:
                    print(f"want (s={2*i+k:3d}, c={c}) got (s={s_read:3d}, c={c_read})")
print("mismatches", bad)
