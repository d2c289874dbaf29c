#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 resampling path.

Workload (BASELINE.json configs[1]): a batch of 1024 clips x 1 s stereo float32,
48000 -> 44100 Hz, quality VHQ, synthetic white noise.  One "step" = one pass of
the hot path over the whole batch.  Metric: Msamples/s, counting INPUT samples
(frames x channels) per second.

  value     device-resident inputs/outputs, CUDA events around K steps
  e2e       the same batch through the C ABI with HOST (pinned) buffers: H2D and
            D2H copies inside the timed region
  roofline  algorithmic bytes (input + output samples) / measured
            duration of the fused kernel (one launch per step), against the
            measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference (oracle/_ref/libsoxr_ref.so, built from
            /root/reference by oracle/Makefile) on all host cores, one soxr_create
            per clip like soxr.resample, over a bounded sample of the same clips

`--impl reference` times only that CPU arm (rank 0) and prints the same line shape.
Multi-GPU (`torchrun ... bench.py --gpus N`): every rank resamples its own 1024
clips (weak scaling, no data-path collective); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))

IN_RATE, OUT_RATE, QUALITY = 48000, 44100, 'VHQ'
N_CLIPS, FRAMES, CH = 1024, 48000, 2
METRIC = "Msamples/s 48k->44.1k VHQ f32 (input samples/s)"
WORKLOAD = "batch=1024 clips x 1s stereo float32 48000->44100 VHQ"
KERNEL_DESC = ("k_fdr<2560->2352> (single-pass frequency-domain resampler, 1 launch per step) "
               "+ k_fdr_head (direct form, first 5 output frames of each clip)")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _lib_sha():
    import hashlib
    try:
        with open(os.path.join(ROOT, "python-soxr_b200", "soxr", "libb200soxr.so"), "rb") as f:
            return hashlib.sha256(f.read()).hexdigest()[:16]
    except OSError:
        return None


def _traffic():
    """(dram bytes per launch of the dominant kernel, provenance).  ncu cannot run inside the timed region, so the
    figure comes from the committed `ncu --set full` capture (profiles/latest_traffic.json); the provenance says
    which build of the kernel source it was taken from, and whether that is the source this run was built from."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "latest_traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None, "no capture committed"
    h = hashlib.sha256()
    for name in ("fdr_stage.cuh", "fft_regs.cuh", "fdr_stage.cu"):
        try:
            with open(os.path.join(ROOT, "python-soxr_b200", "csrc", name), "rb") as f:
                h.update(f.read())
        except OSError:
            pass
    cur = h.hexdigest()[:16]
    same = t.get("kernel_source_sha") == cur
    src = "ncu --set full capture %s (kernel source sha %s; this run's source %s: %s)" % (
        t.get("capture", "?"), t.get("kernel_source_sha", "?"), cur, "same" if same else "DIFFERENT - re-capture")
    return t.get("dram_bytes_per_launch"), src


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # median of the upper half = clocks under load (idle samples before/after the region are lower)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(n_clips_sample, passes, threads, warmup_passes=1):
    """The unmodified reference on the host cores: soxr.resample semantics (one
    soxr_create + div_len loop + flush per clip, soxr_ext.cpp:224-260), clips spread
    over a thread pool with the GIL released by ctypes (tests/test_resample.py:184-188)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_soxr as R
    from concurrent.futures import ThreadPoolExecutor
    if not R.available():
        return None
    rng = np.random.default_rng(1)
    clips = [rng.standard_normal((FRAMES, CH)).astype(np.float32) for _ in range(min(n_clips_sample, 64))]

    def one(i):
        return R.divide_proc(IN_RATE, OUT_RATE, clips[i % len(clips)], QUALITY).shape[0]

    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(warmup_passes):                      # warm-up (thread start, malloc arenas, page cache)
            list(ex.map(one, range(n_clips_sample)))
        times = []
        for _ in range(passes):
            t0 = time.perf_counter()
            list(ex.map(one, range(n_clips_sample)))
            times.append(time.perf_counter() - t0)
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary (strong-scaling / gather / pageable) measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    in_samples = N_CLIPS * FRAMES * CH

    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = 256
        times = cpu_reference_arm(sample, max(1, args.steps), cores, warmup_passes=min(args.warmup, 3))
        if times is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsoxr_ref.so not built"}))
            return 0
        t = sum(times) / len(times)
        v = sample * FRAMES * CH / t / 1e6
        desc = "%d of the 1024 clips per step, %d threads, one soxr_create per clip" % (sample, cores)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (reference VHQ engine cr64s), f32 I/O",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": desc},
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "reference", "sample": desc},
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    import numpy as np
    import torch
    import soxr
    from soxr import soxr_ext
    if not torch.cuda.is_available() or not soxr_ext.device_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        # one process per GPU: run (and allocate the pinned host buffers of the e2e leg) on the CPUs next to this
        # rank's GPU, as a multi-GPU launcher would; harmless where the topology is flat or the call is refused
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa = "rank bound to its GPU's CPU set (%d cpus)" % len(os.sched_getaffinity(0))
        except Exception as exc:       # noqa: BLE001
            numa = "not bound (%s)" % type(exc).__name__
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    out_frames = soxr_ext.output_frames(IN_RATE, OUT_RATE, FRAMES)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1 + rank)
    x = torch.randn((N_CLIPS, FRAMES, CH), dtype=torch.float32, device=dev, generator=gen)
    y = torch.empty((N_CLIPS, out_frames, CH), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        soxr.resample_batch(x, IN_RATE, OUT_RATE, QUALITY, out=y, stream=stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the clock sampler (nvidia-smi -lms 100) is started and given time to come up BEFORE the warm-up steps, so that
    # nothing but a barrier separates the warm-up from the timed steps (an idle quarter second there made the first
    # timed steps slower than the rest: kernel_ms_avg 3 % above the median)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = soxr_ext.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_beg.record(stream)
    for a, b in evs:
        a.record(stream)
        step()
        b.record(stream)
    e_end.record(stream)
    barrier()
    total_ms = e_beg.elapsed_time(e_end)
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    launches = soxr_ext.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = world * in_samples * args.steps / (total_ms_max * 1e-3) / 1e6

    # the optional final gather of the output shards (NCCL all_gather over NVLink), timed on its own:
    # `value` is the sharded throughput (no data-path collective)
    gather_ms = None
    if dist is not None:
        from soxr import distributed as sdist
        full = sdist.gather_shards(y, world * N_CLIPS)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(3):
            full = sdist.gather_shards(y, world * N_CLIPS)
        g1.record(stream)
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather_ms = float(tg.item())
        del full

    # ---- secondary measurements (BASELINE.json configs[1] and configs[3] as STRONG scaling: total work fixed) ----
    def timed(fn, steps, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        barrier()
        tt = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    extra = {}
    if not args.no_extra:
        n_steps2 = max(3, min(args.steps, 10))
        # cfg2, 1024 clips in TOTAL: every rank takes 1024 / N of them
        n2 = N_CLIPS // world
        ms = timed(lambda: soxr.resample_batch(x[:n2], IN_RATE, OUT_RATE, QUALITY, out=y[:n2], stream=stream.cuda_stream), n_steps2)
        extra["strong_cfg2_1024_clips_total"] = {"clips_per_gpu": n2, "ms_per_step": ms,
                                                 "value_msamples_s": N_CLIPS * FRAMES * CH / (ms * 1e-3) / 1e6}
        # cfg4: int16 96000 -> 48000 HQ, 4096 mono clips in TOTAL, sharded over the ranks (512 per GPU at N = 8)
        n4 = 4096 // world
        x4 = (torch.randn((n4, 96000), device=dev, generator=gen) * 5000).round().to(torch.int16)
        y4 = torch.empty((n4, 48000), dtype=torch.int16, device=dev)
        ms4 = timed(lambda: soxr.resample_batch(x4, 96000, 48000, 'HQ', out=y4, stream=stream.cuda_stream), n_steps2)
        alg4 = (x4.numel() + y4.numel()) * 2
        extra["strong_cfg4_4096_int16_clips_total"] = {"clips_per_gpu": n4, "ms_per_step": ms4,
                                                       "value_msamples_s": 4096 * 96000 / (ms4 * 1e-3) / 1e6,
                                                       "achieved_gbs_per_gpu": alg4 / (ms4 * 1e-3) / 1e9}
        del x4, y4
        if dist is not None:
            # resample + gather of the outputs on every rank, pipelined over sub-batches
            # (soxr.distributed.resample_batch_gathered: pieces pushed into every peer's result over NVLink peer
            # mappings by the copy engines while the next piece is computed; NCCL send/receive where there is no peer access)
            from soxr import distributed as sdist
            msg = timed(lambda: sdist.resample_batch_gathered(x, IN_RATE, OUT_RATE, QUALITY, sub_batches=2), 5, warm=2)
            recv = (world - 1) * N_CLIPS * out_frames * CH * 4
            extra["weak_with_overlapped_all_gather"] = {
                "ms_per_step": msg, "value_msamples_s": world * in_samples / (msg * 1e-3) / 1e6,
                "bytes_received_per_gpu": recv, "link_gbs_per_gpu": recv / (msg * 1e-3) / 1e9,
                "exchange": "symmetric-memory peer pushes (copy engines)" if sdist._SYMM_CACHE and any(v is not None for v in sdist._SYMM_CACHE.values()) else "NCCL send/recv",
                "note": "every rank receives (N - 1) output shards: the step is bound by the link, not by the kernels; "
                        "the compute is hidden behind the exchange"}

    # sanity: the result is finite and has signal power (white noise through a unity-gain LPF)
    chk = float(y[:4].double().pow(2).mean().item())
    if not (0.5 < chk < 1.5):
        raise SystemExit("bench.py: implausible output power %r" % chk)

    # ---- end to end through the C ABI with host buffers --------------------------------
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((N_CLIPS, FRAMES, CH), dtype=torch.float32).pin_memory()
        yh = torch.empty((N_CLIPS, out_frames, CH), dtype=torch.float32).pin_memory()
        xh.copy_(x)
        torch.cuda.synchronize(dev)
        xn, yn = xh.numpy(), yh.numpy()
        n_e2e = max(3, min(args.steps, 10))
        for _ in range(2):
            soxr.resample_batch(xn, IN_RATE, OUT_RATE, QUALITY, out=yn)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            soxr.resample_batch(xn, IN_RATE, OUT_RATE, QUALITY, out=yn)   # synchronous: returns with yn filled
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        if not np.allclose(yn[3, :64], y[3, :64].cpu().numpy(), atol=1e-5):
            raise SystemExit("bench.py: host path disagrees with device path")
        e2e = {"value": world * in_samples * n_e2e / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(xh.numel() * 4), "d2h_bytes_per_step": int(yh.numel() * 4),
               "steps": n_e2e, "ms_per_step": dt / n_e2e * 1e3,
               "api": "soxr.resample_batch(numpy) -> b200soxr_resample_batch, pinned host buffers"}
        if not args.no_extra:
            # the same call with ordinary (pageable) ndarrays: the library stages them through pinned slots itself
            xp = np.empty(xn.shape, np.float32)
            xp[...] = xn
            yp = np.empty(yn.shape, np.float32)
            for _ in range(2):
                soxr.resample_batch(xp, IN_RATE, OUT_RATE, QUALITY, out=yp)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                soxr.resample_batch(xp, IN_RATE, OUT_RATE, QUALITY, out=yp)
            dtp = time.perf_counter() - t0
            tp = torch.tensor([dtp], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            dtp = float(tp.item())
            if not np.array_equal(yp[5, :256], yn[5, :256]):
                raise SystemExit("bench.py: pageable host path disagrees with the pinned one")
            e2e["pageable"] = {"value": world * in_samples * n_e2e / dtp / 1e6, "ms_per_step": dtp / n_e2e * 1e3,
                               "slowdown_vs_pinned": dtp / dt,
                               "api": "the same call with ordinary (pageable) numpy arrays"}
            del xp, yp
        del xh, yh

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = _peaks()
    traffic, traffic_src = _traffic()
    alg_bytes = (N_CLIPS * FRAMES * CH + N_CLIPS * out_frames * CH) * 4
    k_ms = sorted(kernel_ms)
    k_avg = sum(kernel_ms) / len(kernel_ms)
    achieved = alg_bytes / (k_avg * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_gpu": N_CLIPS, "frames": FRAMES, "channels": CH,
                   "in_rate": IN_RATE, "out_rate": OUT_RATE, "quality": QUALITY,
                   "l2": "inputs (393 MB) + outputs (361 MB) per step exceed the 126 MB L2; no explicit flush",
                   "output_msamples_per_s": value * out_frames / FRAMES,
                   "sharding": "clips split contiguously over ranks, no data-path collective",
                   "final_all_gather_ms": gather_ms, "cpu_affinity": numa, "secondary": extra or None},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": KERNEL_DESC, "library_sha256_16": _lib_sha(),
                     "kernel_ms_avg": k_avg, "kernel_ms_min": k_ms[0], "kernel_ms_median": k_ms[len(k_ms) // 2],
                     "kernel_ms_each": [round(v, 4) for v in kernel_ms],
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "note": "kernel_ms_* = CUDA-event time of one whole step (all its launches) on the launching stream; "
                             "shared-memory/FMA bound, not HBM bound: see DESIGN.md"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        sample = 256
        times = cpu_reference_arm(sample, 3, cores)
        if times:
            tb = min(times)
            line["cpu_baseline"] = {"value": sample * FRAMES * CH / tb / 1e6, "unit": "Msamples/s", "cores": cores,
                                    "kind": "reference",
                                    "sample": "%d of the 1024 clips, best of 3 passes, %d threads, one soxr_create per clip"
                                              % (sample, cores)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
