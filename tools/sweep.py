#!/usr/bin/env python3
"""Sweep JWC_TUNE launch shapes with bench.py (device-resident part only) and print one line each."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
tunes = sys.argv[2:] or [""]
for t in tunes:
    env = dict(os.environ, JWC_TUNE=t)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "10", "--warmup", "3",
           "--no-e2e", "--no-cpu"]
    if os.environ.get("SWEEP_BATCH"):
        cmd += ["--batch", os.environ["SWEEP_BATCH"]]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"{t or '(default)':45s} fwd {d['forward_gsps']:7.1f} GS/s ({d['roofline']['forward_frac']:.3f})  "
              f"rev {d['reverse_gsps']:7.1f} GS/s ({d['roofline']['reverse_frac']:.3f})  rt_err {d['roundtrip_max_abs_err']:.2e}",
              flush=True)
    except Exception as e:
        print(t, "FAILED", e, out.stderr[-500:], flush=True)
