#!/usr/bin/env python3
"""bench.py - the benchmark of the JWave wavelet hot path on B200: every BASELINE.json config in one line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|c2|c3|c4|c5] [--impl reference]

Workloads (BASELINE.json `configs`; C1, one 2^16 Haar signal, is the reference's CPU-runnable case and a parity test):
  c2  Daubechies4 FWT 1-D, full depth (14 levels), 65,536 signals x 2^14 fp64 PER GPU        <- the headline line
  c3  Symlet8 WPT 1-D, 6 levels, 4,096 signals x 2^16 fp64 PER GPU
  c4  Daubechies20 FWT 2-D, full depth, 64 images of 8192 x 8192 fp64 PER GPU
  c5  Coiflet5 FWT 3-D, full depth, ONE 1024^3 fp64 volume: on one GPU at N = 1, slab-decomposed over all N
      GPUs at N > 1 (the only workload with an exchange step), with full-size parity against the oracle
The default run measures all four: the JSON line's top level is c2 (value, roofline, cpu_baseline, e2e, ...) and
`workloads` holds the same record for c3, c4 and c5.

One *step* = forward transform of the whole batch followed by the reverse transform of the coefficients (the
reference's own timing unit: ParallelWPTPerformanceTest.java:270-295).  `value` = samples transformed per second,
counting both directions, inputs resident in HBM, CUDA-event time, max over ranks.  `e2e` = the same step through
the host-buffer C ABI (jwc_fwt1d / jwc_wpt1d / jwc_fwt2d / jwc_fwt3d) with pinned HOST arrays, H2D and D2H inside
the timed region, and the PCIe ceiling measured beside it (concurrent pinned copies on all ranks at once).

N > 1: one process per GPU (torchrun).  c2-c4 shard independent signals / images by rank with NO collective on
the data path: weak scaling (per-GPU batch fixed) for `value` AND for `e2e`.  c5 is one volume over all GPUs
(strong scaling; stated in its record).

--impl reference times the CPU restatement of JWave (oracle/, all host threads) on the stated c2 batch (and bounded
samples of c3-c5).  Together with cpu_sample() it is the only code here that executes oracle/ for timing; the parity
checks of c5 at N > 1 use it as the checker.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (wavelet class, kind, n, level, items per GPU, description)
    "c2": ("Daubechies4", "fwt", 1 << 14, 14, 65536,
           "Daubechies4 FWT 1-D full depth (14 levels), 65536 signals x 2^14 fp64 per GPU"),
    "c3": ("Symlet8", "wpt", 1 << 16, 6, 4096,
           "Symlet8 WPT 1-D 6 levels, 4096 signals x 2^16 fp64 per GPU"),
    "c4": ("Daubechies20", "fwt2d", 8192, 13, 64,
           "Daubechies20 FWT 2-D full depth (13 + 13 levels), 64 images of 8192 x 8192 fp64 per GPU"),
    "c5": ("Coiflet5", "fwt3d", 1024, 10, 1,
           "Coiflet5 FWT 3-D full depth (10 + 10 + 10 levels) on one 1024^3 fp64 volume"),
}
METRIC = "Daub4 FWT / Sym8 WPT GSamples/s at 1-8 B200, % HBM roofline, vs JWave CPU"
FP64_PEAK_TFLOPS = 36.7  # measured here with tools/microbench.cu (DFMA), see DESIGN.md
DIMS = {"fwt": 1, "wpt": 1, "fwt2d": 2, "fwt3d": 3}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md):
    NVML every 5 ms when pynvml is importable, else one nvidia-smi query per 100 ms."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, uuid=None):
        self.index = index
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (f"GPU-{uuid}", str(uuid)):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        try:
                            h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                            break
                        except Exception:
                            h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml, self._h = pynvml, h
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n, h = self._nvml, self._h
        self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)))
        try:
            self.power.append(n.nvmlDeviceGetPowerUsage(h) / 1000.0)
        except Exception:
            pass
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for name, bit in self.BITS:
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [t.strip() for t in out.strip().split(",")]
        if len(f) >= 7:
            if f[0].replace(".", "").isdigit():
                self.sm.append(float(f[0]))
            if f[1].replace(".", "").isdigit():
                self.mx.append(float(f[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(float(os.environ.get("JWB_CLOCK_PERIOD", 0.005)) if self._nvml is not None else 0.1)

    def start(self):
        self._t.start()

    def stop(self):
        self._stop.set()
        self._t.join(timeout=10)
        out = {"sm_mhz": statistics.median(self.sm) if self.sm else None,
               "sm_min_mhz": min(self.sm) if self.sm else None,
               "sm_max_mhz": max(self.mx) if self.mx else None,
               "samples": len(self.sm), "reasons": sorted(self.reasons),
               "source": "nvml" if self._nvml is not None else "nvidia-smi"}
        if self.power:
            out["power_w_median"] = statistics.median(self.power)
        return out


# ---- CPU side: the oracle port of JWave's own parallel drivers ----------------------------------------------

def oracle_step(co, cls, kind, x, level, threads, pwpt=False):
    """forward + reverse on the host cores with the reference's own work decomposition:
      fwt / wpt : independent signals, one per task on a fixed pool (ParallelizationOpportunityTest.java:79-110);
                  pwpt=True: ParallelWaveletPacketTransform (packets of ONE signal in parallel, signals looped)
      fwt2d     : ParallelTransform.forward/reverse(double[][]) per image (ParallelTransform.java:70-134)
      fwt3d     : ParallelTransform.forward/reverse(double[][][]) (ParallelTransform.java:137-213)"""
    if kind == "fwt":
        c = co.batch_1d(co.FWT, co.FORWARD, cls, x, level, threads)
        return co.batch_1d(co.FWT, co.REVERSE, cls, c, level, threads)
    if kind == "wpt":
        if pwpt:
            c = co.parallel_wpt(co.FORWARD, cls, x, level, threads)
            return co.parallel_wpt(co.REVERSE, cls, c, level, threads)
        c = co.batch_1d(co.WPT, co.FORWARD, cls, x, level, threads)
        return co.batch_1d(co.WPT, co.REVERSE, cls, c, level, threads)
    if kind == "fwt2d":
        lv = x.shape[-1].bit_length() - 1
        c = co.parallel_2d(co.FWT, co.FORWARD, cls, x, lv, lv, threads)
        return co.parallel_2d(co.FWT, co.REVERSE, cls, c, lv, lv, threads)
    lv = [s.bit_length() - 1 for s in x.shape]
    c = co.parallel_3d(co.FWT, co.FORWARD, cls, x, lv[0], lv[1], lv[2], threads)
    return co.parallel_3d(co.FWT, co.REVERSE, cls, c, lv[0], lv[1], lv[2], threads)


CPU_WHAT = {
    "fwt": "FastWaveletTransform per signal on a fixed thread pool",
    "wpt": "WaveletPacketTransform per signal on a fixed thread pool",
    "fwt2d": "ParallelTransform(FastWaveletTransform) 2-D: rows, then columns, as pool tasks; images looped",
    "fwt3d": "ParallelTransform(FastWaveletTransform) 3-D: slices as pool tasks, then the outer axis over blocks of j",
}


def cpu_sample(cls, kind, n, level, target_s=8.0, fixed_signals=None):
    """Time the oracle on a bounded sample of the workload on all host threads (about target_s of CPU wall time;
    1-D: the number of signals is grown until one step takes that long; 2-D: one full-size image; 3-D: a 512^3
    volume - an eighth of the samples, same filter, same drivers)."""
    import numpy as np
    from oracle import c_oracle as co
    threads = co.max_threads()
    rng = np.random.default_rng(42)
    if kind in ("fwt2d", "fwt3d"):
        shape = (1, n, n) if kind == "fwt2d" else (min(n, 512),) * 3
        x = rng.standard_normal(shape)
        small = rng.standard_normal((1, 256, 256) if kind == "fwt2d" else (64, 64, 64))
        oracle_step(co, cls, kind, small, level, threads)  # warm the pool
        t0 = time.perf_counter()
        oracle_step(co, cls, kind, x, level, threads)
        dt = time.perf_counter() - t0
        return {"value": 2.0 * x.size / dt * 1e-9, "unit": "GSamples/s", "cores": threads, "kind": "port",
                "sample": ("one 8192 x 8192 image" if kind == "fwt2d" else f"one {shape[0]}^3 volume")
                          + f" (forward+reverse), {dt:.2f} s wall",
                "what": CPU_WHAT[kind] + " - C restatement of the JWave CPU path, -O2 -ffp-contract=off"}
    signals = max(threads, 8)
    oracle_step(co, cls, kind, rng.standard_normal((signals, n)), level, threads)  # warm the pool
    if fixed_signals:
        signals = fixed_signals
    else:
        while True:  # grow the probe until it is long enough to extrapolate from
            probe = rng.standard_normal((signals, n))
            t0 = time.perf_counter()
            oracle_step(co, cls, kind, probe, level, threads)
            dt = max(time.perf_counter() - t0, 1e-6)
            if dt >= 0.1 * target_s or signals >= (1 << 15):
                break
            signals *= 4
        signals = int(max(threads, min(signals * target_s / dt, 1 << 16)))
    x = rng.standard_normal((signals, n))
    t0 = time.perf_counter()
    oracle_step(co, cls, kind, x, level, threads)
    dt = time.perf_counter() - t0
    out = {"value": 2.0 * signals * n / dt * 1e-9, "unit": "GSamples/s", "cores": threads, "kind": "port",
           "sample": f"{signals} signals x {n} (forward+reverse), {dt:.2f} s wall",
           "what": CPU_WHAT[kind] + " - C restatement of the JWave CPU path, -O2 -ffp-contract=off"}
    if kind == "wpt":  # also the reference's own within-signal decomposition, on a small sample
        few = rng.standard_normal((max(2, min(signals, 64)), n))
        oracle_step(co, cls, kind, few[:2], level, threads, pwpt=True)
        t0 = time.perf_counter()
        oracle_step(co, cls, kind, few, level, threads, pwpt=True)
        out["parallel_wpt_value"] = 2.0 * few.shape[0] * n / (time.perf_counter() - t0) * 1e-9
        out["parallel_wpt_what"] = "ParallelWaveletPacketTransform decomposition (packets of one signal in parallel, signals looped)"
    return out


def run_reference(args):
    """--impl reference: the CPU path only (no GPU, none of our kernels).  Each step transforms the STATED c2 batch
    (65536 signals x 2^14, forward + reverse) on all host threads; c3-c5 are timed on bounded samples."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import c_oracle as co
    name = "c2" if args.workload == "all" else args.workload
    cls, kind, n, level, batch, desc = WORKLOADS[name]
    threads = co.max_threads()
    if args.batch:
        batch = args.batch
    if DIMS[kind] == 1:
        # the stated batch, in blocks of 8192 signals (1 GiB) so the host arrays stay small; same work per step
        block = min(batch, 8192)
        x = np.random.default_rng(42).standard_normal((block, n))
        reps = max(1, batch // block)

        def step():
            for _ in range(reps):
                oracle_step(co, cls, kind, x, level, threads)
        samples = reps * block * n
        sample = f"{reps * block} signals x {n} per step (the stated batch), in blocks of {block}"
    else:
        base = cpu_sample(cls, kind, n, level)
        x = np.random.default_rng(42).standard_normal((1, n, n) if kind == "fwt2d" else (min(n, 512),) * 3)

        def step():
            oracle_step(co, cls, kind, x, level, threads)
        samples = x.size
        sample = base["sample"].split(" (")[0] + " per step"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = 2.0 * samples * args.steps / dt * 1e-9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GSamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "step": "forward + reverse of the whole batch", "items_per_gpu": batch, "n": n,
                   "level": level, "wavelet": cls, "sample_per_step": sample, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": "GSamples/s", "cores": threads, "kind": "port",
                         "sample": f"{sample}, {args.steps} steps",
                         "what": CPU_WHAT[kind] + " - C restatement of the JWave CPU path, -O2 -ffp-contract=off"},
        "e2e": {"value": value, "unit": "GSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.workload == "all" and not args.no_cpu:
        line["workloads"] = {w: {"cpu_baseline": cpu_sample(*WORKLOADS[w][:4])} for w in ("c3", "c4", "c5")}
    print(json.dumps(line))
    return 0


# ---- GPU side -------------------------------------------------------------------------------------------------

def host_fit(items, bytes_per_item, world, arrays=3, frac=0.4):
    """Largest power-of-two fraction of `items` whose `arrays` pinned host copies, on all `world` ranks of this
    box, stay below `frac` of the available host memory."""
    avail = 64 << 30
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
    except Exception:
        pass
    while items > 1 and arrays * items * bytes_per_item * world > frac * avail:
        items //= 2
    return max(items, 1)


def pcie_probe(torch, dist, world, nbytes=1 << 30, reps=4):
    """Concurrent pinned H2D + D2H on every rank at once: the per-GPU bus ceiling the e2e pipeline works under."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    both()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        both()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    del h_in, h_out, d_in, d_out
    return nbytes * reps / float(dt[0]) * 1e-9  # GB/s per direction per GPU, both directions busy


class SlabRunner:
    """forward / reverse of this rank's slab through one of the slab classes; layout "j" keeps the coefficients in
    the j-slab layout between the two (one re-cut per direction), layout "i" returns them as i-slabs (two)."""

    def __init__(self, slab, n, level, layout):
        self.slab, self.n, self.level = slab, n, level
        self.layout = layout if hasattr(slab, "forward_t") and getattr(slab, "exchange", "") == "copies" else "i"

    def forward(self, x, out=None):
        if self.layout == "j":
            return self.slab.forward_t(x, self.n, self.level, self.level, self.level)
        return self.slab.forward(x, self.n, self.level, self.level, self.level, out=out)

    def reverse(self, c, out=None):
        if self.layout == "j":
            return self.slab.reverse_t(c, self.n, self.level, self.level, self.level)
        return self.slab.reverse(c, self.n, self.level, self.level, self.level, out=out)

    def dense_coef(self, c):
        """(dense slab of coefficients, the axis the ranks' slabs are concatenated along)"""
        return (self.slab.t_to_dense(c), 1) if self.layout == "j" else (c, 0)


def slab_parity(torch, dist, np, make_runner, cls, n, level, rank, world):
    """c5 at N > 1: parity of the slab-decomposed path against the oracle, visible in the bench line because the
    driver's test box has one GPU.  (a) a small random volume through the same slab code, gathered and compared
    with the oracle's 3-D transform; (b) the FULL 1024^3 volume with a separable probe x = u (x) v (x) w: the
    3-D transform of a rank-one volume is the outer product of the three 1-D transforms (linearity; the axis
    passes act on different indices), so every rank checks its whole slab against three oracle vectors."""
    from oracle import c_oracle as co
    out = {}
    ns = 128 if world <= 4 else 256  # P/W and Q/W must be powers of two >= 16 for the peer-mapped form
    ls = ns.bit_length() - 1
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    full = torch.randn(ns, ns, ns, dtype=torch.float64, device="cuda", generator=g)  # same on every rank
    run = make_runner(ns, ls)
    mine = full[rank * (ns // world):(rank + 1) * (ns // world)].contiguous()
    fc = run.forward(mine)
    f, axis = run.dense_coef(fc)
    f = f.clone()
    r = run.reverse(fc).clone()
    parts_f = [torch.empty_like(f) for _ in range(world)]
    parts_r = [torch.empty_like(r) for _ in range(world)]
    dist.all_gather(parts_f, f)
    dist.all_gather(parts_r, r)
    if rank == 0:
        xf = full.cpu().numpy()
        want_f = co.transform_3d(co.FWT, co.FORWARD, cls, xf, ls, ls, ls)
        got_f = torch.cat(parts_f, dim=axis).cpu().numpy()
        want_r = co.transform_3d(co.FWT, co.REVERSE, cls, got_f, ls, ls, ls)
        out["small_volume"] = f"{ns}^3"
        out["small_forward_max_err"] = float(np.abs(got_f - want_f).max())
        out["small_reverse_max_err"] = float(np.abs(torch.cat(parts_r).cpu().numpy() - want_r).max())
        out["small_tol"] = 1e-12 * float(max(np.abs(xf).max(), np.abs(got_f).max()))
    del run, full, mine, f, fc, r, parts_f, parts_r
    # (b) separable probe at full size
    rng = np.random.default_rng(77)
    u, v, w = (rng.standard_normal(n) for _ in range(3))
    fu, fv, fw = (co.transform_1d(co.FWT, co.FORWARD, cls, t, level) for t in (u, v, w))
    tu, tv, tw, tfu, tfv, tfw = (torch.from_numpy(t).cuda() for t in (u, v, w, fu, fv, fw))
    p = n // world
    sl = slice(rank * p, (rank + 1) * p)
    x = ((tu[sl, None, None] * tv[None, :, None]) * tw[None, None, :]).contiguous()
    run = make_runner(n, level)
    fc = run.forward(x)
    f, axis = run.dense_coef(fc)
    if axis == 0:
        want = (tfu[sl, None, None] * tfv[None, :, None]) * tfw[None, None, :]
    else:
        want = (tfu[:, None, None] * tfv[None, sl, None]) * tfw[None, None, :]
    e_f = (f - want).abs().max()
    del want, f
    back = run.reverse(fc)
    e_r = (back - x).abs().max()
    t = torch.stack([e_f, e_r, x.abs().max()])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["full_probe"] = f"{n}^3 rank-one volume u(x)v(x)w, every slab entry against the oracle's three 1-D transforms"
    out["full_forward_max_err"] = float(t[0])
    out["full_roundtrip_max_err"] = float(t[1])
    out["full_tol"] = 1e-12 * float(t[2])
    out["full_roundtrip_note"] = "Coiflet5's tap table reconstructs to ~5e-8 on the reference CPU too (SURVEY F8)"
    out["coefficient_layout"] = run.layout + "-slabs"
    del run, x, fc, back
    torch.cuda.empty_cache()
    return out


def run_workload(name, args, env):
    """Device-resident timing, per-kernel roofline, e2e through the C ABI and the CPU sample of one workload."""
    torch, dist, np, jw, _lib = env["torch"], env["dist"], env["np"], env["jw"], env["_lib"]
    from jwave_b200.device import DeviceTransforms
    rank, world, local = env["rank"], env["world"], env["local"]
    cls, kind, n, level, batch, desc = WORKLOADS[name]
    if args.batch:
        batch = args.batch
    dims = DIMS[kind]
    K = _lib.WPT if kind == "wpt" else _lib.FWT
    wavelet = jw.WaveletBuilder.create(cls)
    L = wavelet.getMotherWavelength()
    dev = DeviceTransforms(wavelet, local)
    steps = args.steps if dims == 1 else max(3, min(args.steps, 5))  # the 2-D / 3-D steps are 50-150 ms each
    warmup = max(args.warmup, 3) if dims == 1 else 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    slab, slab_mode, parity = None, None, None
    shape = {1: (batch, n), 2: (batch, n, n), 3: (n, n, n)}[dims]
    if dims == 3 and world > 1:
        # config 5 proper: ONE volume, slab-decomposed along i, re-cut along j for the i pass
        from jwave_b200.distributed import PeerSlabVolumeTransform, SlabVolumeTransform, device_axis_fn

        def make_slab(edge):
            if args.slab in ("peer", "copies"):
                try:
                    return PeerSlabVolumeTransform(dev, edge, edge, edge,
                                                   exchange="stores" if args.slab == "peer" else "copies")
                except Exception as e:  # no symmetric memory / shape not covered: NCCL path
                    if rank == 0:
                        print(f"[bench] peer-mapped slabs unavailable ({e}); using the all-to-all path", file=sys.stderr)
            return SlabVolumeTransform(device_axis_fn(dev), kind=K)

        def make_runner(edge, lv):
            return SlabRunner(make_slab(edge), edge, lv, args.slab_layout)
        parity = slab_parity(torch, dist, np, make_runner, cls, n, level, rank, world)
        slab = make_runner(n, level)
        slab_mode = {"PeerSlabVolumeTransform": "peer " + getattr(slab.slab, "exchange", ""),
                     "SlabVolumeTransform": "all-to-all"}[type(slab.slab).__name__] + f", coefficients in {slab.layout}-slabs"
        shape = (n // world, n, n)

    gen = torch.Generator(device="cuda")
    gen.manual_seed(42 + rank)
    if dims == 2:
        # 64 images = 32 GiB per array: the reverse writes back into `x` (x -> coef -> x), three full-size arrays
        # (x, coef, the library's axis scratch) instead of four; image 0 is kept aside for the round-trip check
        x = torch.empty(*shape, dtype=torch.float64, device="cuda")
        for i in range(shape[0]):
            x[i].normal_(generator=gen)
        x0 = x[0].clone()
        coef = torch.empty_like(x)
        back = x
    else:
        x = torch.randn(*shape, dtype=torch.float64, device="cuda", generator=gen)
        coef = torch.empty_like(x)
        back = torch.empty_like(x)

    def run(direction, src, dst):
        if slab is not None:
            out = dst if slab_mode.startswith("all-to-all") else None  # the peer paths return their own buffers
            return slab.forward(src, out=out) if direction == _lib.FORWARD else slab.reverse(src, out=out)
        elif dims == 1:
            dev.transform1d(K, direction, src, level, out=dst)
        elif dims == 2:
            dev.transform2d(K, direction, src, level, level, out=dst)
        else:
            dev.transform3d(K, direction, src, level, level, level, out=dst)
        return dst

    res = None
    for _ in range(warmup):
        res = run(_lib.REVERSE, run(_lib.FORWARD, x, coef), back)
    barrier()
    rt_err = float((res[0] - x0).abs().max()) if dims == 2 else float((res - x).abs().max())

    # ---- timed region: `steps` steps, device resident ------------------------------------------------
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None)) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fwd_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches0 = dev.launch_count()
    if sampler:
        sampler.start()
    dev.ctx.profile(True)  # event pair around every kernel launch, read after the timed region
    barrier()
    ev[0].record()
    for i in range(steps):
        fwd_ev[i][0].record()
        c = run(_lib.FORWARD, x, coef)
        fwd_ev[i][1].record()
        run(_lib.REVERSE, c, back)
    ev[1].record()
    barrier()
    clocks = sampler.stop() if sampler else None
    prof = dev.ctx.profile_report()
    dev.ctx.profile(False)
    launches = dev.launch_count() - launches0
    ms = ev[0].elapsed_time(ev[1])
    fwd_ms = statistics.mean(a.elapsed_time(b) for a, b in fwd_ev)
    t = torch.tensor([ms, fwd_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, fwd_ms = float(t[0]), float(t[1])
    ms_per_step = ms / steps
    rev_ms = ms_per_step - fwd_ms
    samples = x.numel()  # per GPU
    value = 2.0 * samples * world / (ms_per_step * 1e-3) * 1e-9

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    # Per-kernel CUDA-event times from the timed region (jwc_profile_*).  Algorithmic work of one launch
    # (SURVEY.md section 8d): 16 B per sample it transforms (one read + one write), and direct-form flops
    # 4 L (1 - 2^-m) per sample for m fused FWT levels, 2 L m for m WPT levels.
    hbm_peak, peak_src = peaks()
    bytes_per_sample = 16.0 * dims
    flops_per_sample = 2.0 * L * level if kind == "wpt" else dims * (2.0 * L * 2.0 * (1.0 - 0.5 ** level))
    prof_total = sum(r[2] for r in prof) or 1.0
    kernels = []
    for kname, count, total_ms, units, lv in prof:
        tk = total_ms / count * 1e-3
        k_flops = units * (2.0 * L * lv if "wpt" in kname else 4.0 * L * (1.0 - 0.5 ** lv))
        kernels.append({"kernel": f"{kname}<{L}>", "launches": count, "avg_ms": total_ms / count,
                        "share": total_ms / prof_total, "samples_per_launch": units, "levels": lv,
                        "hbm_frac": 16.0 * units / tk * 1e-9 / hbm_peak, "fp64_frac": k_flops / tk * 1e-12 / FP64_PEAK_TFLOPS})
    kernels.sort(key=lambda k: -k["share"])
    dom = kernels[0]
    t_dom = dom["avg_ms"] * 1e-3
    if dom["hbm_frac"] >= dom["fp64_frac"]:
        achieved = 16.0 * dom["samples_per_launch"] / t_dom * 1e-9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": peak_src}
    else:
        achieved = dom["fp64_frac"] * FP64_PEAK_TFLOPS
        roof = {"bound": "fp64", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": dom["fp64_frac"], "peak_source": "measured DFMA peak (tools/microbench.cu, profiles/r01_microbench.txt)"}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from ncu --set full captures
    if os.path.exists(tpath):
        # per launch at the bench batch, or (c4 / c5: captured on a smaller batch) DRAM bytes per sample x samples per launch
        tj = json.load(open(tpath))
        traffic = tj.get(f"{name}:{dom['kernel']}")
        if traffic is None and tj.get(f"{name}:{dom['kernel']}:per_sample") is not None:
            traffic = tj[f"{name}:{dom['kernel']}:per_sample"] * dom["samples_per_launch"]
    t_hbm = samples * bytes_per_sample / (hbm_peak * 1e9)
    t_fp64 = samples * flops_per_sample / (FP64_PEAK_TFLOPS * 1e12)
    t_roof = max(t_hbm, t_fp64)
    roof.update({"traffic": traffic, "kernel": dom["kernel"], "kernel_share_of_step": dom["share"],
                 "kernel_avg_ms": dom["avg_ms"], "kernels": kernels,
                 "forward_ms": fwd_ms, "reverse_ms": rev_ms,
                 "algorithmic_bytes_per_sample": bytes_per_sample, "algorithmic_flops_per_sample": flops_per_sample,
                 "direction_roofline": "hbm" if t_hbm >= t_fp64 else "fp64"})
    if slab is None:  # per-GPU roofline fractions of a whole direction (a slab step also holds the exchanges)
        roof.update({"forward_frac": t_roof / (fwd_ms * 1e-3), "reverse_frac": t_roof / (rev_ms * 1e-3)})
    slab_info = None
    if slab is not None:
        xb = x.numel() * 8 * (world - 1) // world
        nx = 1 if slab.layout == "j" else 2
        slab_info = {"mode": slab_mode, "exchanges_per_direction": nx, "bytes_sent_per_gpu_per_exchange": xb,
                     "local_compute_roofline_ms_per_step": 2.0 * t_roof * 1e3,
                     "note": "value is strong-scaled: one 1024^3 volume over all GPUs"}
        if hasattr(slab.slab, "measure"):
            mm = slab.slab.measure(x, level, coef=slab.layout)  # one forward call: as run / local passes only / copies only
            if mm:
                tm = torch.tensor([mm["full_ms"], mm["compute_only_ms"], mm["copies_only_ms"]], dtype=torch.float64, device="cuda")
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                slab_info.update({"forward_ms_as_run": float(tm[0]), "forward_ms_local_passes_only": float(tm[1]),
                                  "forward_ms_copies_only": float(tm[2]), "chunks_per_exchange": slab.slab.chunks,
                                  "exposed_exchange_ms_per_direction": float(tm[0] - tm[1]),
                                  "nvlink_gbs_sent_per_gpu_copies_alone": nx * xb / (float(tm[2]) * 1e-3) * 1e-9})

    group = None
    if slab is not None and not args.no_e2e:
        # The same volume through the C ABI's own multi-device form: ONE context over all GPUs of the box
        # (jwc_create_multi), one process, host buffers in and out - run by rank 0 while the other ranks wait.
        # This is the e2e of c5 at N > 1 and the driver-visible check of the device group (the test box has 1 GPU).
        del coef, back
        coef = back = None
        torch.cuda.empty_cache()
        barrier()
        if rank == 0:
            try:
                from jwave_b200.transforms import CudaContext
                from oracle import c_oracle as co
                gctx = CudaContext(list(range(world)))
                gt = jw.CudaFastWaveletTransform(wavelet, context=gctx)
                rng = np.random.default_rng(77)
                u, v, w = (rng.standard_normal(n) for _ in range(3))
                fu, fv, fw = (torch.from_numpy(co.transform_1d(co.FWT, co.FORWARD, cls, t_, level)) for t_ in (u, v, w))
                hx = torch.empty(n, n, n, dtype=torch.float64).pin_memory()
                torch.mul(torch.from_numpy(u)[:, None, None] * torch.from_numpy(v)[None, :, None], torch.from_numpy(w)[None, None, :], out=hx)
                hc = torch.empty_like(hx).pin_memory()
                hb = torch.empty_like(hx).pin_memory()
                gl, gh = gctx._lib, gctx.handle

                def gstep():
                    gctx.check(gl.jwc_fwt3d(gh, gt._wid, _lib.FORWARD, hx.data_ptr(), hc.data_ptr(), n, n, n, level, level, level), "group forward")
                    gctx.check(gl.jwc_fwt3d(gh, gt._wid, _lib.REVERSE, hc.data_ptr(), hb.data_ptr(), n, n, n, level, level, level), "group reverse")
                gstep()
                t0 = time.perf_counter()
                gstep()
                gstep()
                gdt = (time.perf_counter() - t0) / 2
                # parity of the full volume: rank-one probe against the oracle's three 1-D transforms, plane by plane
                err = 0.0
                for i0 in range(0, n, 128):
                    want = (fu[i0:i0 + 128, None, None] * fv[None, :, None]) * fw[None, None, :]
                    err = max(err, float((hc[i0:i0 + 128] - want).abs().max()))
                group = {"value": 2.0 * n ** 3 / gdt * 1e-9, "unit": "GSamples/s", "devices": world,
                         "ms_per_step": gdt * 1e3, "h2d_bytes_per_step": 2 * 8 * n ** 3, "d2h_bytes_per_step": 2 * 8 * n ** 3,
                         "api": "jwc_create_multi + jwc_fwt3d (one process, one context, host buffers, pinned)",
                         "forward_max_err_vs_oracle": err, "tol": 1e-12 * float(hx.abs().max()),
                         "roundtrip_max_abs_err": float((hb - hx).abs().max())}
                del hx, hc, hb
                gctx.close()
            except Exception as e:
                group = {"error": f"{type(e).__name__}: {e}"}
        # the other ranks wait on the HOST (a key in the rendezvous store): an NCCL barrier would park a spinning
        # kernel on the GPUs rank 0 is using
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set(f"group_done_{name}", "1")
        else:
            store.wait([f"group_done_{name}"])
        barrier()
    else:
        del coef, back
    if dims == 2:
        del x0
    # ---- e2e: the same step through the host-buffer C ABI, weak-scaled like `value` ----------------------
    e2e = None
    if not args.no_e2e and slab is None:
        item = {1: n, 2: n * n, 3: n * n * n}[dims]
        want = {1: batch, 2: min(batch, 8), 3: 1}[dims]  # 2-D: 8 images (4 GiB per host array) stand for the 64
        eb = host_fit(want, item * 8, env["ranks_per_node"])
        lib, hnd, wid = dev.ctx._lib, dev.ctx.handle, dev.wid
        hx = torch.empty(eb * item, dtype=torch.float64).pin_memory()
        hc = torch.empty_like(hx).pin_memory()
        hb = torch.empty_like(hx).pin_memory()
        hx.copy_(x.reshape(-1)[:eb * item])

        def call(direction, src, dst):
            if dims == 1:
                fn = lib.jwc_fwt1d if kind == "fwt" else lib.jwc_wpt1d
                st = fn(hnd, wid, direction, src.data_ptr(), dst.data_ptr(), eb, n, level)
            elif dims == 2:
                st = lib.jwc_fwt2d(hnd, wid, direction, src.data_ptr(), dst.data_ptr(), eb, n, n, level, level)
            else:
                st = lib.jwc_fwt3d(hnd, wid, direction, src.data_ptr(), dst.data_ptr(), n, n, n, level, level, level)
            dev.ctx.check(st, "e2e")

        def e2e_step():
            call(_lib.FORWARD, hx, hc)
            call(_lib.REVERSE, hc, hb)

        e2e_step()
        e2e_steps = max(2, min(args.steps, 3 if dims > 1 else 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nbytes = eb * item * 8
        step_s = float(dt[0]) / e2e_steps
        e2e = {"value": 2.0 * eb * item * world / step_s * 1e-9, "unit": "GSamples/s",
               "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": 2 * nbytes,
               "items_per_gpu": eb, "stated_items_per_gpu": batch, "steps": e2e_steps, "scaling": "weak",
               "roundtrip_max_abs_err": float((hb - hx).abs().max()),
               "api": {1: "jwc_fwt1d / jwc_wpt1d", 2: "jwc_fwt2d", 3: "jwc_fwt3d"}[dims] +
                      " (host buffers, pinned; chunked H2D / compute / D2H pipeline)"}
        del hx, hc, hb
        if env.get("pcie_gbs") is None:
            env["pcie_gbs"] = pcie_probe(torch, dist, world)
        # both directions of the bus carry 2 * nbytes per step; the probe keeps both busy at once
        e2e["pcie"] = {"gbs_per_direction_per_gpu_concurrent": env["pcie_gbs"],
                       "achieved_gbs_per_direction_per_gpu": 2 * nbytes / step_s * 1e-9,
                       "how": "1 GiB pinned H2D and D2H copies in flight together on every rank at once"}
        e2e["frac_of_pcie_ceiling"] = e2e["pcie"]["achieved_gbs_per_direction_per_gpu"] / env["pcie_gbs"]
    del x
    dev.close()
    torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_sample(cls, kind, n, level)

    rec = {
        "value": value, "unit": "GSamples/s", "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
        "scaling": "strong" if slab is not None else "weak",
        "config": {"workload": desc + (f"; slab-decomposed over {world} GPUs ({slab_mode})" if slab is not None else ""),
                   "step": "forward + reverse of the whole batch",
                   "items_per_gpu": batch, "shape": list(shape), "n": n, "level": level, "wavelet": cls, "taps": L,
                   "parallelism": (f"one volume in i-slabs over {world} GPUs, two re-cuts per direction" if slab is not None
                                   else f"signals sharded over {world} GPU(s), no collective"),
                   "l2": f"inputs ({samples * 8 / 2**30:.1f} GiB per array) exceed the 126 MB L2; no flush needed"},
        "forward_gsps": samples * world / (fwd_ms * 1e-3) * 1e-9,
        "reverse_gsps": samples * world / (rev_ms * 1e-3) * 1e-9,
        "roundtrip_max_abs_err": rt_err,
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    if slab_info:
        rec["slab"] = slab_info
    if group:
        rec["e2e"] = group  # c5 at N > 1: the device group IS the end-to-end path of one volume
    if parity:
        rec["slab_parity"] = parity
        rec["slab_parity_max_err"] = max(parity.get("small_forward_max_err", 0.0), parity.get("small_reverse_max_err", 0.0),
                                         parity["full_forward_max_err"])
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=["all"] + sorted(WORKLOADS), default="all")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=0, help="override items per GPU (debug only)")
    ap.add_argument("--slab", choices=["peer", "copies", "all2all"], default="copies",
                    help="c5 on N > 1 GPUs: peer-mapped slabs filled by chunked device copies (default) or by the "
                         "axis kernels' own stores (peer), or NCCL all-to-all")
    ap.add_argument("--slab-layout", choices=["j", "i"], default="j",
                    help="c5 on N > 1 GPUs (copies): leave the coefficients in j-slabs between forward and reverse (one "
                         "re-cut per direction, SURVEY.md 8e) or return them as i-slabs (two)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import jwave_b200 as jw
    from jwave_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    env = {"torch": torch, "dist": dist, "np": np, "jw": jw, "_lib": _lib, "rank": rank, "world": world, "local": local,
           "ranks_per_node": int(os.environ.get("LOCAL_WORLD_SIZE", world)), "pcie_gbs": None}

    names = ["c2", "c3", "c4", "c5"] if args.workload == "all" else [args.workload]
    recs = {}
    for nm in names:
        try:
            recs[nm] = run_workload(nm, args, env)
        except Exception as e:  # a failing side workload must not take the headline down with it
            if nm == names[0]:
                raise
            recs[nm] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
    if rank == 0:
        head = recs[names[0]]
        line = {"metric": METRIC, "value": head["value"], "unit": "GSamples/s", "n_gpus": world,
                "steps": head["steps"], "warmup": head["warmup"], "ms_per_step": head["ms_per_step"],
                "higher_is_better": True, "scaling": head["scaling"], "vs_baseline": None, "dtype": "f64",
                "data": "synthetic"}
        line.update({k: v for k, v in head.items() if k not in line})
        if len(names) > 1:
            line["workloads"] = {nm: recs[nm] for nm in names[1:]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
