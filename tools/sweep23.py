#!/usr/bin/env python3
"""Sweep JWC_TUNE for the 2-D / 3-D workloads (c4 with 4 images, c5)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload = sys.argv[1]
extra = ["--batch", "4"] if workload == "c4" else []
for t in sys.argv[2:] or [""]:
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "3", "--warmup", "3"] + extra,
                         env=dict(os.environ, JWC_TUNE=t), capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"{workload} {t or '(default)':45s} fwd {d['forward_gsps']:6.1f} ({d['roofline']['forward_frac']:.3f}) rev {d['reverse_gsps']:6.1f} ({d['roofline']['reverse_frac']:.3f}) launches {d['gpu_launches']}", flush=True)
    except Exception as e:
        print(workload, t, "FAILED", e, out.stderr[-400:], flush=True)
