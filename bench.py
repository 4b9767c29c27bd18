#!/usr/bin/env python3
"""bench.py - the headline benchmark of the JWave wavelet hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3] [--impl reference]

Workloads (BASELINE.json `configs`):
  c2 (default) Daubechies4 FWT 1-D, full depth (14 levels), 65,536 signals x 2^14 fp64 PER GPU
  c3           Symlet8 WPT 1-D, 6 levels, 4,096 signals x 2^16 fp64 PER GPU

One *step* = forward transform of the whole batch followed by the reverse transform of the
coefficients (the reference's own timing unit: ParallelWPTPerformanceTest.java:270-295).
`value` = samples transformed per second, counting both directions (2 x batch x n per step),
inputs resident in HBM.  `e2e` = the same step through the host-buffer C ABI (jwc_fwt1d /
jwc_wpt1d) with pinned HOST arrays, H2D and D2H inside the timed region.

N > 1: one process per GPU (torchrun), independent signals sharded by rank, no collective on the
data path (weak scaling: the per-GPU batch is fixed).  Times are CUDA-event times, max over ranks.

--impl reference times the CPU restatement of JWave (oracle/, all host threads) on a bounded
sample of the same workload.  It is the only code path here that executes oracle/ for timing.
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
    # 2-D / 3-D configs (device-resident timing only; `n` is the edge length, level = log2 n per axis)
    "c4": ("Daubechies20", "fwt2d", 8192, 13, 16,
           "Daubechies20 FWT 2-D full depth, 8192 x 8192 fp64 images, 16 per GPU per step (BASELINE batch: 64)"),
    "c5": ("Coiflet5", "fwt3d", 1024, 10, 1,
           "Coiflet5 FWT 3-D full depth on one 1024^3 fp64 volume per GPU (single-GPU form of config 5)"),
}
METRIC = "Daub4 FWT / Sym8 WPT GSamples/s at 1-8 B200, % HBM roofline, vs JWave CPU"
FP64_PEAK_TFLOPS = 36.7  # measured here with tools/microbench.cu (DFMA), see DESIGN.md


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
            self._stop.wait(0.005 if self._nvml is not None else 0.1)

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


def oracle_step(co, cls, kind, x, level, threads, pwpt=False):
    """forward + reverse of a [batch][n] sample on the host cores.  Independent signals run one
    per task on a fixed pool (the pattern of ParallelizationOpportunityTest.java:79-110) - for the
    WPT that is what a batch user of the reference gets the most out of.  pwpt=True times the
    reference's own ParallelWaveletPacketTransform decomposition instead (packets of ONE signal in
    parallel, signals looped; README: "1.2-1.3x")."""
    if kind == "fwt":
        c = co.batch_1d(co.FWT, co.FORWARD, cls, x, level, threads)
        return co.batch_1d(co.FWT, co.REVERSE, cls, c, level, threads)
    if pwpt:
        c = co.parallel_wpt(co.FORWARD, cls, x, level, threads)
        return co.parallel_wpt(co.REVERSE, cls, c, level, threads)
    c = co.batch_1d(co.WPT, co.FORWARD, cls, x, level, threads)
    return co.batch_1d(co.WPT, co.REVERSE, cls, c, level, threads)


def cpu_sample(cls, kind, n, level, target_s=2.0, max_signals=None):
    """Time the oracle on a bounded sample sized for ~target_s of wall time on all host threads."""
    import numpy as np
    from oracle import c_oracle as co
    threads = co.max_threads()
    rng = np.random.default_rng(42)
    signals = max(threads, 8)
    oracle_step(co, cls, kind, rng.standard_normal((signals, n)), level, threads)  # warm the pool
    while True:  # grow the probe until it is long enough to extrapolate from
        probe = rng.standard_normal((signals, n))
        t0 = time.perf_counter()
        oracle_step(co, cls, kind, probe, level, threads)
        dt = max(time.perf_counter() - t0, 1e-6)
        if dt >= 0.25 * target_s or signals >= (1 << 15):
            break
        signals *= 4
    signals = int(max(threads, min(signals * target_s / dt, 1 << 16)))
    if max_signals:
        signals = min(signals, max_signals)
    for _ in range(3):
        x = rng.standard_normal((signals, n))
        t0 = time.perf_counter()
        oracle_step(co, cls, kind, x, level, threads)
        dt = time.perf_counter() - t0
        if dt >= 0.5 * target_s or signals >= (max_signals or (1 << 16)):
            break
        signals = int(min(signals * target_s / dt, max_signals or (1 << 16)))
    out = {"value": 2.0 * signals * n / dt * 1e-9, "unit": "GSamples/s", "cores": threads, "kind": "port",
           "sample": f"{signals} signals x {n} (forward+reverse), {dt:.2f} s wall",
           "what": ("FastWaveletTransform" if kind == "fwt" else "WaveletPacketTransform") +
                   " per signal on a fixed thread pool - C restatement of the JWave CPU path, -O2 -ffp-contract=off"}
    if kind == "wpt":  # also the reference's own within-signal decomposition, on a small sample
        few = rng.standard_normal((max(2, min(signals, 64)), n))
        oracle_step(co, cls, kind, few[:2], level, threads, pwpt=True)
        t0 = time.perf_counter()
        oracle_step(co, cls, kind, few, level, threads, pwpt=True)
        out["parallel_wpt_value"] = 2.0 * few.shape[0] * n / (time.perf_counter() - t0) * 1e-9
        out["parallel_wpt_what"] = "ParallelWaveletPacketTransform decomposition (packets of one signal in parallel, signals looped)"
    return out


def run_reference(args):
    """--impl reference: the CPU path only (no GPU, none of our kernels)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import c_oracle as co
    cls, kind, n, level, batch, desc = WORKLOADS[args.workload]
    threads = co.max_threads()
    base = cpu_sample(cls, kind, n, level, target_s=1.0)
    signals = int(base["sample"].split()[0])
    x = np.random.default_rng(42).standard_normal((signals, n))
    for _ in range(args.warmup):
        oracle_step(co, cls, kind, x, level, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(co, cls, kind, x, level, threads)
    dt = time.perf_counter() - t0
    value = 2.0 * signals * n * args.steps / dt * 1e-9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GSamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "sample_per_step": f"{signals} signals x {n}", "host_threads": threads},
        "cpu_baseline": dict(base, value=value, sample=f"{signals} signals x {n} per step, {args.steps} steps"),
        "e2e": {"value": value, "unit": "GSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=0, help="override signals per GPU (debug only)")
    ap.add_argument("--slab", choices=["peer", "copies", "all2all"], default="copies",
                    help="c5 on N > 1 GPUs: peer-mapped slabs filled by strided device copies (default) or by the "
                         "axis kernels' own stores (peer), or NCCL all-to-all")
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
    from jwave_b200.device import DeviceTransforms

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cls, kind, n, level, batch, desc = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    dims = {"fwt": 1, "wpt": 1, "fwt2d": 2, "fwt3d": 3}[kind]
    K = _lib.WPT if kind == "wpt" else _lib.FWT
    wavelet = jw.WaveletBuilder.create(cls)
    L = wavelet.getMotherWavelength()
    dev = DeviceTransforms(wavelet, local)

    gen = torch.Generator(device="cuda")
    gen.manual_seed(42 + rank)
    slab = None
    shape = {1: (batch, n), 2: (batch, n, n), 3: (n, n, n)}[dims]
    if dims == 3 and world > 1:
        # config 5 proper: ONE volume, slab-decomposed along i, all-to-all for the i pass
        from jwave_b200.distributed import PeerSlabVolumeTransform, SlabVolumeTransform, device_axis_fn
        slab_mode = "all-to-all"
        if args.slab in ("peer", "copies") and K == _lib.FWT:
            try:  # peer-mapped slabs: exchanges folded into the kernels' stores, or strided device copies
                slab = PeerSlabVolumeTransform(dev, n, n, n, exchange="stores" if args.slab == "peer" else "copies")
                slab_mode = "peer stores" if args.slab == "peer" else "peer copies"
            except Exception as e:  # no symmetric memory / shape not covered: NCCL path
                print(f"[bench] peer-mapped slabs unavailable ({e}); using the all-to-all path", file=sys.stderr)
        if slab is None:
            slab = SlabVolumeTransform(device_axis_fn(dev), kind=K)
        shape = (n // world, n, n)
    x = torch.randn(*shape, dtype=torch.float64, device="cuda", generator=gen)
    coef = torch.empty_like(x)
    back = torch.empty_like(x)

    def run(direction, src, dst):
        if slab is not None:
            out = dst if slab_mode == "all-to-all" else None  # the peer paths return their own symmetric buffer
            if direction == _lib.FORWARD:
                return slab.forward(src, n, level, level, level, out=out)
            return slab.reverse(src, n, level, level, level, out=out)
        elif dims == 1:
            dev.transform1d(K, direction, src, level, out=dst)
        elif dims == 2:
            dev.transform2d(K, direction, src, level, level, out=dst)
        else:
            dev.transform3d(K, direction, src, level, level, level, out=dst)
        return dst

    def step():
        return run(_lib.REVERSE, run(_lib.FORWARD, x, coef), back)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        back = step()
    barrier()
    rt_err = float((back - x).abs().max())  # sanity: the timed work really is a transform pair

    # ---- timed region: K steps, device resident ------------------------------------------------
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None)) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fwd_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = dev.launch_count()
    if sampler:
        sampler.start()
    dev.ctx.profile(True)  # event pair around every kernel launch, read after the timed region
    barrier()
    ev[0].record()
    for i in range(args.steps):
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
    ms_per_step = ms / args.steps
    rev_ms = ms_per_step - fwd_ms
    samples = x.numel()  # per GPU
    value = 2.0 * samples * world / (ms_per_step * 1e-3) * 1e-9

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    # Per-kernel CUDA-event times from the timed region (jwc_profile_*).  Algorithmic work of one
    # launch: 16 B per sample it transforms (one read + one write), and
    # direct-form flops 4 L (1 - 2^-m) per sample for m fused FWT levels, 2 L m for m WPT levels.
    hbm_peak, peak_src = peaks()
    bytes_per_sample = 16.0 * dims
    flops_per_sample = 2.0 * L * level if kind == "wpt" else dims * (2.0 * L * 2.0 * (1.0 - 0.5 ** level))
    prof_total = sum(r[2] for r in prof) or 1.0
    kernels = []
    for name, count, total_ms, units, lv in prof:
        t = total_ms / count * 1e-3
        k_bytes = 16.0 * units
        k_flops = units * (2.0 * L * lv if "wpt" in name else 4.0 * L * (1.0 - 0.5 ** lv))
        kernels.append({"kernel": f"{name}<{L}>", "launches": count, "avg_ms": total_ms / count,
                        "share": total_ms / prof_total, "samples_per_launch": units, "levels": lv,
                        "hbm_frac": k_bytes / t * 1e-9 / hbm_peak, "fp64_frac": k_flops / t * 1e-12 / FP64_PEAK_TFLOPS})
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
        traffic = json.load(open(tpath)).get(f"{args.workload}:{dom['kernel']}")
    t_hbm = samples * bytes_per_sample / (hbm_peak * 1e9)
    t_fp64 = samples * flops_per_sample / (FP64_PEAK_TFLOPS * 1e12)
    t_roof = max(t_hbm, t_fp64)
    roof.update({"traffic": traffic, "kernel": dom["kernel"], "kernel_share_of_step": dom["share"],
                 "kernel_avg_ms": dom["avg_ms"], "kernels": kernels,
                 "forward_ms": fwd_ms, "reverse_ms": rev_ms,
                 "algorithmic_bytes_per_sample": bytes_per_sample, "algorithmic_flops_per_sample": flops_per_sample,
                 "direction_roofline": "hbm" if t_hbm >= t_fp64 else "fp64",
                 "forward_frac": t_roof / (fwd_ms * 1e-3), "reverse_frac": t_roof / (rev_ms * 1e-3)})

    # ---- e2e: the same step through the host-buffer C ABI ------------------------------------
    e2e = None
    if not args.no_e2e and dims == 1:
        eb = batch if world == 1 else max(batch // world, 1)
        host = jw.CudaFastWaveletTransform(wavelet, context=dev.ctx) if kind == "fwt" else \
            jw.CudaWaveletPacketTransform(wavelet, context=dev.ctx)
        hx = torch.empty(eb, n, dtype=torch.float64).pin_memory()
        hc = torch.empty_like(hx).pin_memory()
        hb = torch.empty_like(hx).pin_memory()
        hx.copy_(x[:eb])
        f1d = dev.ctx._lib.jwc_fwt1d if kind == "fwt" else dev.ctx._lib.jwc_wpt1d

        def e2e_step():
            st = f1d(dev.ctx.handle, dev.wid, _lib.FORWARD, hx.data_ptr(), hc.data_ptr(), eb, n, level)
            dev.ctx.check(st, "e2e forward")
            st = f1d(dev.ctx.handle, dev.wid, _lib.REVERSE, hc.data_ptr(), hb.data_ptr(), eb, n, level)
            dev.ctx.check(st, "e2e reverse")

        e2e_step()
        e2e_steps = max(2, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nbytes = eb * n * 8
        e2e = {"value": 2.0 * eb * n * world * e2e_steps / float(dt[0]) * 1e-9, "unit": "GSamples/s",
               "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": 2 * nbytes,
               "signals_per_gpu": eb, "steps": e2e_steps,
               "roundtrip_max_abs_err": float((hb - hx).abs().max()),
               "api": "jwc_fwt1d / jwc_wpt1d (host buffers, pinned; chunked H2D/compute/D2H pipeline)"}
        del hx, hc, hb

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and dims == 1:
        cpu = cpu_sample(cls, kind, n, level)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GSamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if slab is not None else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc + (f"; one volume slab-decomposed over {world} GPUs, "
                                           + {"peer stores": "exchanges folded into the axis kernels' peer stores",
                                              "peer copies": "blocks copied straight into the peers' slabs (no pack / unpack / NCCL)",
                                              "all-to-all": "2 all-to-all per direction"}[slab_mode]
                                           if slab is not None else ""), "step": "forward + reverse of the whole batch",
                       "items_per_gpu": batch, "shape": list(shape), "n": n, "level": level, "wavelet": cls, "taps": L,
                       "parallelism": f"signals sharded over {world} GPU(s), no collective",
                       "l2": f"inputs ({samples * 8 / 2**30:.1f} GiB per array) exceed the 126 MB L2; no flush needed"},
            "forward_gsps": samples * world / (fwd_ms * 1e-3) * 1e-9,
            "reverse_gsps": samples * world / (rev_ms * 1e-3) * 1e-9,
            "roundtrip_max_abs_err": rt_err,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
