#!/usr/bin/env python
"""bench.py -- the headline benchmark of the chaos-ultra B200 render backend.

A *step* is one quality frame of the hot path: iteration kernel (adaptive supersampling) + compose.
Workloads are the configurations of BASELINE.json; the default is configs[1]:
mandelbrot 3840x2160, maxIter 10000, adaptive supersampling (maxSS 8), FP64, full-set viewport.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

ONE JSON line on stdout (rank 0).  Keys beyond the base contract:
  value      pixel-iterations/s, whole job, records + RGBA staying in HBM (DEVICE output mode)
  e2e        the same metric through the reference-facing C ABI with HOST output: every step ends with the
             composed RGBA frame in pinned host memory
  roofline   dominant kernel (fractalRenderMain*): FP pipe utilisation against the FMA issue peak measured in
             this run with bench_kernels/peak.cubin (MEASURED_PEAKS.json has no FP64/FP32 figure)
  cpu_baseline  the oracle port of the same sampling algorithm on the host cores, bounded sample
--impl reference runs the reference's OWN kernels (oracle/_ref: src/main/cuda compiled by nvcc 12.9 for
sm_100a, launched like the Java host does) on the same device; the reference has no CPU implementation.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

A, FOV, REUSE, ZOOMING, ZOOM_IN = 1, 4, 8, 16, 32

# roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel(s), per frame, from one
# `ncu --set full` capture of the same command (never measured in a bench run); workload -> (bytes, capture)
NCU_TRAFFIC = {
    "c2": (int((8.36 + 92.24 + 11.22 + 0.003 + 10.89 + 0.76) * 1e6), "profiles/r01g_passes_c2.txt: chaosPassA + chaosPassB + chaosPassC Double"),
    "c3": (int((127.08 + 85.31 + 132.73 + 9.60) * 1e6), "profiles/r01b_fast_frame_c3.txt: chaosReusePassFloat + compose "
           "(part of the 133 MB of records written and of the 33 MB frame stays in the 126 MB L2)"),
}


def seg(cx, cy, zoom, W, H):
    relW = 1.0 / float(H) * W
    return [cx - relW * zoom / 2, cy - zoom / 2, cx + relW * zoom / 2, cy + zoom / 2]


# name -> workload
WORKLOADS = {
    "c1": dict(fractal="mandelbrot", W=1024, H=1024, center=(-0.5, 0.0), zoom=2.0, maxIter=500, maxSS=1.0, flags=0, double=True,
               desc="mandelbrot 1024x1024 maxIter 500 FP64 1 sample, full set"),
    "c2": dict(fractal="mandelbrot", W=3840, H=2160, center=(-0.5, 0.0), zoom=2.0, maxIter=10000, maxSS=8.0, flags=A, double=True,
               desc="mandelbrot 3840x2160 maxIter 10000 adaptive SS (maxSS 8) FP64, centre (-0.5,0) h=2"),
    "c2ex2": dict(fractal="mandelbrot", W=3840, H=2160, center=(-0.235125, 0.827215), zoom=4.0e-5, maxIter=10000, maxSS=8.0,
                  flags=A, double=True, desc="mandelbrot 3840x2160 maxIter 10000 adaptive SS (maxSS 8) FP64, 'M ex 2'"),
    "c2f32": dict(fractal="mandelbrot", W=3840, H=2160, center=(-0.5, 0.0), zoom=2.0, maxIter=10000, maxSS=8.0, flags=A,
                  double=False, desc="mandelbrot 3840x2160 maxIter 10000 adaptive SS (maxSS 8) FP32, full set"),
    "c3": dict(kind="zoom", fractal="mandelbrot", W=3840, H=2160, center=(-0.748, 0.1), zoom=2.0, maxIter=1600, maxSS=2.0,
               flags=A | FOV | REUSE | ZOOMING | ZOOM_IN, double=None, focus=(1920, 1080),
               desc="zoom sequence 3840x2160: frame 0 quality, then fast frames (sample reuse + foveation + compose), "
                    "zoomAt(centre, in) per frame, mandelbrot maxIter 1600 maxSS 2, precision by the reference's rule"),
    "c4": dict(fractal="mandelbrot", W=8192, H=8192, center=(-0.551042868375875, 0.62714332109057), zoom=8.00592947491907e-9,
               maxIter=200000, maxSS=1.0, flags=0, double=True, desc="mandelbrot 8192x8192 maxIter 200000 FP64 1 sample, 'M ex 5'"),
    "c5": dict(fractal="julia", W=3840, H=2160, center=(0.0, 0.0), zoom=4.0, maxIter=900, maxSS=8.0, flags=A, double=True,
               julia_c=(-0.4, 0.6), desc="julia c=(-0.4,0.6) 3840x2160 maxIter 900 adaptive SS (maxSS 8) FP64"),
}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md, clocks line): NVML polled
    every 5 ms from a thread (nvidia-smi takes longer to start than a short timed region lasts); falls back to
    `nvidia-smi -lms` if pynvml is unavailable."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.lines = []
        self.samples = []      # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _poll(self):
        n, h = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append((n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), n.nvmlDeviceGetCurrentClocksEventReasons(h)))
            except Exception:
                try:
                    self.samples.append((n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), n.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
                except Exception:
                    break
            self._stop.wait(0.005)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.device]) if vis and vis.split(",")[self.device].strip().isdigit() else self.device
            h = n.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            self._nvml = (n, h)
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=1)
            n = self._nvml[0]
            names = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                     ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))
            reasons = set()
            for _, mask in self.samples:
                for name, attr in names:
                    if mask & getattr(n, attr, 0):
                        reasons.add(name)
            sm = sorted(float(c) for c, _ in self.samples)
            if not sm:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml"}
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvidia-smi"}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measure_fma_peak(device: int, double: bool):
    """Peak FMA lane-operations/s of the FP64 (or FP32) pipe, measured with bench_kernels/peak.cubin through
    cuda.bindings.driver; best of 5 launches after a warm-up."""
    import numpy as np
    from cuda.bindings import driver as cu

    def ok(res):
        if res[0] != cu.CUresult.CUDA_SUCCESS:
            raise RuntimeError("cuda driver error %s" % (res[0],))
        return res[1] if len(res) == 2 else res[1:]

    ok(cu.cuInit(0))
    dev = ok(cu.cuDeviceGet(device))
    ctx = ok(cu.cuDevicePrimaryCtxRetain(dev))
    ok(cu.cuCtxPushCurrent(ctx))
    try:
        sms = ok(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev))
        mod = ok(cu.cuModuleLoad(str(ROOT / "bench_kernels" / "peak.cubin").encode()))
        fn = ok(cu.cuModuleGetFunction(mod, b"peak_fp64" if double else b"peak_fp32"))
        blocks, threads, trips = sms * 8, 256, 1 << 16
        out = ok(cu.cuMemAlloc(blocks * threads * 8))
        e0, e1 = ok(cu.cuEventCreate(0)), ok(cu.cuEventCreate(0))
        seed = np.array([1.0], dtype=np.float64 if double else np.float32)
        args = (np.array([int(out)], dtype=np.uint64), np.array([trips], dtype=np.uint32), seed)
        argp = np.array([a.ctypes.data for a in args], dtype=np.uint64)
        best = 1e30
        for it in range(6):
            ok(cu.cuEventRecord(e0, 0))
            ok(cu.cuLaunchKernel(fn, blocks, 1, 1, threads, 1, 1, 0, 0, argp.ctypes.data, 0))
            ok(cu.cuEventRecord(e1, 0))
            ok(cu.cuEventSynchronize(e1))
            ms = ok(cu.cuEventElapsedTime(e0, e1))
            if it:
                best = min(best, ms)
        ok(cu.cuMemFree(out))
        ok(cu.cuModuleUnload(mod))
        return blocks * threads * trips * 8 / (best * 1e-3)   # FMA lane-ops per second
    finally:
        cu.cuCtxPopCurrent()
        cu.cuDevicePrimaryCtxRelease(dev)


def quiet_nccl():
    """the image exports NCCL_DEBUG=VERSION, which makes NCCL print a banner on stdout next to the JSON line"""
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_model(cu, wl):
    m = cu.RenderingModel(canvasWidth=wl["W"], canvasHeight=wl["H"])
    m.setPlaneSegmentFromCenter(wl["center"][0], wl["center"][1], wl["zoom"])
    m.maxIterations = wl["maxIter"]
    m.maxSuperSampling = wl["maxSS"]
    m.useAdaptiveSuperSampling = bool(wl["flags"] & A)
    m.useFoveatedRendering = False
    m.useSampleReuse = False
    m.forcePrecision = 2 if wl["double"] else 1
    return m


def run_ours(args, wl, rank, world, local):
    import numpy as np
    import torch
    cu = importlib.import_module("chaos-ultra_b200")

    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.engine is not None:
        os.environ["CHAOS_ENGINE"] = str(args.engine)
    prov = cu.CudaFractalRendererProvider(kernels_dir=os.environ.get("CHAOS_KERNELS_DIR"), device=local)   # (diagnostic builds)
    r = prov.getRenderer(wl["fractal"], False)
    if wl["fractal"] == "julia":
        r.setFractalCustomParams("%r;%r" % tuple(wl["julia_c"]))
    W, H = wl["W"], wl["H"]
    band_rows = 32
    model = make_model(cu, wl)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    class DevBuf:   # wraps the renderer's device RGBA frame for torch (NCCL gather of row bands)
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (H, W), "typestr": "<i4", "data": (ptr, False), "version": 2}

    part = importlib.import_module("chaos-ultra_b200.partition")

    def gather_to_rank0(frame):
        """the one exchange step of the multi-GPU path: composed RGBA row bands -> rank 0, NCCL send/recv over NVLink"""
        part.gather_bands(frame, rank, world, band_rows, dist)

    def timed_loop(mode, steps, warmup, sampler=None):
        """W untimed + K timed steps.  Device time: one GPU -- the library's CUDA events around each frame's kernels
        (stats.frame_ms, recorded on the stream the kernels are launched on), summed; several GPUs -- torch CUDA events
        bracketing the K steps on the stream that carries the exchange step (every step ends there).  The host clock
        between the two barriers is kept as well.  All reduced with MAX over ranks."""
        if r.getState() == cu.STATE_READY_TO_RENDER:
            r.freeRenderingResources()
        r.initializeRendering(W, H, None, mode)
        r.setPartition(rank, world, band_rows)
        if world == 1 and os.environ.get("CHAOS_EMULATE_PART"):   # diagnostics: what ONE rank of an N-GPU run does
            pi, pn = map(int, os.environ["CHAOS_EMULATE_PART"].split(":"))
            r.setPartition(pi, pn, band_rows)
        frame = shared = token = None
        if world > 1:
            frame = torch.as_tensor(DevBuf(r.outputRGBADevicePointer()), device="cuda") if mode == cu.OUTPUT_DEVICE else None
            # compose straight into rank 0's frame over NVLink (CUDA IPC); falls back to the NCCL band gather
            shared = part.share_frame(r, rank, world, dist) if (frame is not None and not args.nccl_gather) else None
            token = torch.zeros(1, device="cuda", dtype=torch.int32)
        host_frame = torch.empty((H, W), dtype=torch.int32).pin_memory() if (world > 1 and args.e2e_host_copy) else None
        iters = launches = skipped = 0
        rms = cms = fms = 0.0
        e_start = e_end = None
        for it in range(warmup + steps):
            if it == warmup:
                barrier()
                if sampler is not None:
                    sampler.start()
                t0 = time.perf_counter()
                iters = launches = skipped = 0
                rms = cms = fms = 0.0
                if world > 1:
                    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e_start.record()
            r.renderQuality(model)
            st = r.stats()
            iters += st.pixel_iterations
            skipped += st.skipped_iterations
            launches += st.kernel_launches
            rms += st.render_ms
            cms += st.compose_ms
            fms += st.frame_ms
            if world > 1 and frame is not None:
                if shared is not None:
                    dist.all_reduce(token)                  # the exchange step: every rank's bands have landed in rank 0's frame
                else:
                    gather_to_rank0(frame)
                if host_frame is not None:
                    if rank == 0:
                        host_frame.copy_(frame, non_blocking=False)
                    dist.all_reduce(token)                  # nobody composes the next frame into rank 0's before it is out
        if e_end is not None:
            e_end.record()
        barrier()
        dt = time.perf_counter() - t0
        clocks = sampler.stop() if sampler is not None else None
        dev_ms = e_start.elapsed_time(e_end) if e_end is not None else fms
        if world > 1:
            t = torch.tensor([dt, rms, cms, dev_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, rms, cms, dev_ms = t.tolist()
            c = torch.tensor([iters, launches, skipped], device="cuda", dtype=torch.int64)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            iters, launches, skipped = c.tolist()
        exchange = "none" if world == 1 else ("compose writes into rank 0's frame over NVLink (CUDA IPC) + completion all-reduce" if shared is not None
                                              else "NCCL send/recv of row bands to rank 0")
        if shared is not None:
            torch.cuda.synchronize()
            shared.close()
        return dict(seconds=dt, device_seconds=dev_ms * 1e-3, iters=iters, skipped=skipped, launches=launches, render_ms=rms,
                    compose_ms=cms, exchange_ms=max(0.0, dev_ms - fms) if world > 1 else 0.0, clocks=clocks, exchange=exchange)

    sampler = ClockSampler(local) if rank == 0 else None
    dev = timed_loop(cu.OUTPUT_DEVICE, args.steps, args.warmup, sampler)
    # e2e: the public C-ABI call with HOST output (single GPU: compose writes the pinned frame; multi GPU: bands are
    # gathered on the device, then rank 0 copies the frame to pinned host memory)
    args.e2e_host_copy = True
    e2e_mode = cu.OUTPUT_HOST if world == 1 else cu.OUTPUT_DEVICE
    e2e = timed_loop(e2e_mode, args.steps, args.warmup)
    if rank == 0 and world == 1:
        frame_host = r.outputRGBA()
        checksum = int(np.bitwise_xor.reduce(frame_host.ravel()))
    else:
        checksum = None
    # the same frame with every trip executed and tested (CHAOS_SHORTCUTS=0): what the iteration kernel does at full work
    full = None
    if world == 1 and not args.no_full_trips and os.environ.get("CHAOS_SHORTCUTS") is None:
        os.environ["CHAOS_SHORTCUTS"] = "0"
        try:
            r.close()
            r = prov.getRenderer(wl["fractal"], True)
            if wl["fractal"] == "julia":
                r.setFractalCustomParams("%r;%r" % tuple(wl["julia_c"]))
            full = timed_loop(cu.OUTPUT_DEVICE, max(2, min(args.steps, 5)), 3)
            full["steps"] = max(2, min(args.steps, 5))
        finally:
            os.environ.pop("CHAOS_SHORTCUTS", None)

    out = None
    if rank == 0:
        # value: pixel-iterations as SURVEY.md 8d defines them -- the trip counts of the reference's loop for this frame
        # (exact integer from the device counter, equal to the oracle's by the parity tests) -- per second of device time
        value = dev["iters"] / dev["device_seconds"]
        e2e_value = e2e["iters"] / e2e["seconds"]
        px = W * H
        executed = dev["iters"] - dev["skipped"]
        out = {
            "metric": "pixel-iterations/s", "value": value, "unit": "pixel-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev["device_seconds"] * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if wl["double"] else "f32",
            "data": "synthetic", "frames_per_s": args.steps / dev["device_seconds"],
            "wall_ms_per_step": dev["seconds"] * 1e3 / args.steps,
            "timing": "ms_per_step = CUDA events on the stream the kernels are launched on (first render kernel .. compose end"
                      + "), summed over the K steps" + ("" if world == 1 else "; several GPUs: torch CUDA events bracketing the K steps on the stream of the exchange step") + ", MAX over ranks; "
                      "wall_ms_per_step = host clock between the two barrier+synchronize brackets",
            "config": {"workload": args.workload + ": " + wl["desc"], "width": W, "height": H, "max_iterations": wl["maxIter"],
                       "max_super_sampling": wl["maxSS"], "adaptive_ss": bool(wl["flags"] & A),
                       "parallelism": "1 GPU" if world == 1 else "row bands of %d px dealt round-robin over %d GPUs; %s" % (band_rows, world, dev["exchange"]),
                       "pixel_iterations_per_step": dev["iters"] // args.steps,
                       "executed_pixel_iterations_per_step": executed // args.steps,
                       "work_accounting": "pixel_iterations = trips of the reference's loop for this frame (inside points count maxIterations). "
                                          "Orbits whose state recurs bit for bit are PROVEN never to escape and stop early with the same result; "
                                          "executed_pixel_iterations is what the FP pipe actually iterated. CHAOS_SHORTCUTS=0 executes every trip (see full_trips).",
                       "l2": "no input is re-read between steps (inputs are viewport scalars); the %d MB record buffer written per step exceeds the 126 MB L2" % (px * 16 // 1000000),
                       "engine": os.environ.get("CHAOS_ENGINE", "default"), "shortcuts": os.environ.get("CHAOS_SHORTCUTS", "default (3)")},
            "e2e": {"value": e2e_value, "unit": "pixel-iterations/s", "h2d_bytes_per_step": 512, "d2h_bytes_per_step": px * 4 + 32,
                    "ms_per_step": e2e["seconds"] * 1e3 / args.steps, "frames_per_s": args.steps / e2e["seconds"],
                    "note": "chaos_render_quality through the C ABI, host clock around K synchronous calls; inputs are the chaos_params viewport struct "
                            "(kernel parameter space), result = composed RGBA8 frame in pinned host memory" + ("" if world == 1 else " on rank 0 after the NCCL gather"),
                    "rgba_xor_checksum": checksum},
            "gpu_launches": dev["launches"],
            "clocks": dev["clocks"],
            "device_ms_per_step": {"render_kernel": dev["render_ms"] / args.steps, "compose_kernel": dev["compose_ms"] / args.steps,
                                   "exchange_and_skew": dev["exchange_ms"] / args.steps},
        }
        # roofline of the dominant kernel: the FP pipe.  peak = FMA lane-ops/s measured just now on this device.
        # achieved counts FP instructions the kernel ISSUED at the least: 5 per executed trip (the untested scaled form;
        # tested trips issue 6, unscaled ones up to 7), so frac is a lower bound of the pipe's utilisation and cannot
        # exceed 1.  The reference form of the same work (7 instructions per reference trip) is given beside it.
        try:
            peak = measure_fma_peak(local, wl["double"])
            kernel_s = dev["render_ms"] * 1e-3
            min_ops = 5 if wl["double"] else 6
            achieved = executed / world * min_ops / kernel_s
            out["roofline"] = {"bound": "fp64" if wl["double"] else "fp32", "achieved": achieved / 1e9, "peak": peak / 1e9,
                               "unit": "G FP-lane-ops/s", "frac": achieved / peak,
                               "traffic": NCU_TRAFFIC.get(args.workload, (None, None))[0] if world == 1 else None,
                               "traffic_source": NCU_TRAFFIC.get(args.workload, (None, None))[1] if world == 1 else None,
                               "kernel": ("fractalRenderMain%s" if round(wl["maxSS"]) <= 1 else "chaosPassA%s + chaosPassB%s + chaosPassC%s (the iteration kernels of a multi-sample frame)").replace("%s", "Double" if wl["double"] else "Float"),
                               "peak_source": "measured in this run: bench_kernels/peak.cubin, independent FMA chains, best of 5 (not in MEASURED_PEAKS.json)",
                               "achieved_definition": "%d FP instructions x %d executed pixel-iterations per launch sequence / render-kernel time (lower bound of issued instructions)"
                                                      % (min_ops, executed // args.steps // world),
                               "reference_form": {"ops": "7 FP instructions x %d reference pixel-iterations" % (dev["iters"] // args.steps // world),
                                                  "G_ops_per_s": dev["iters"] / world * 7 / kernel_s / 1e9,
                                                  "ratio_to_peak": dev["iters"] / world * 7 / kernel_s / peak,
                                                  "note": "work the reference's loop needs for this frame per second of this kernel; exceeds 1 when trips are proven instead of executed"},
                               "note": "tensor cores and HBM do not bound this kernel (16 B stored per pixel)"}
            if full is not None:
                fk = full["render_ms"] * 1e-3
                out["roofline"]["full_trips"] = {
                    "ms_per_step": full["device_seconds"] * 1e3 / full["steps"], "render_kernel_ms": full["render_ms"] / full["steps"],
                    "pixel_iterations_per_s": full["iters"] / full["device_seconds"],
                    "frac": full["iters"] * (6 if wl["double"] else 7) / fk / peak,
                    "note": "same frame, CHAOS_SHORTCUTS=0: every trip executed with its test (%d FP instructions per trip issued); "
                            "this is the pipe utilisation of the iteration kernel at the reference's full work" % (6 if wl["double"] else 7)}
        except Exception as e:  # measurement helper failed: say so rather than invent a peak
            out["roofline"] = {"bound": "fp64" if wl["double"] else "fp32", "achieved": None, "peak": None, "unit": "G FP-lane-ops/s",
                               "frac": None, "traffic": None, "error": repr(e)}
        # host baseline: the oracle port on the host cores, bounded sample of the same workload (N=1 only)
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            cores = os.cpu_count() or 1
            stride = args.cpu_row_stride
            it_c, smp_c, sec = oracle.render_main_rows_threaded(wl["fractal"], W, H, model.planeSegment, wl["maxIter"], wl["maxSS"],
                                                               wl["flags"], wl["double"], row_stride=stride, threads=cores,
                                                               julia_c=wl.get("julia_c", (0.0, 0.0)))
            out["cpu_baseline"] = {"value": it_c / sec, "unit": "pixel-iterations/s", "cores": cores, "kind": "port",
                                   "sample": "every %dth vote-tile row of the same frame (%d pixel-iterations, %.1f s), oracle/chaos_oracle.c on %d threads"
                                             % (stride, it_c, sec, cores)}
    r.close()
    prov.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_reference(args, wl, rank, world, local):
    """The reference's own kernels (oracle/_ref/<fractal>.<kind>.cubin) on the GPU with the Java host's frame sequence."""
    if rank != 0:
        return None
    import oracle
    cu = importlib.import_module("chaos-ultra_b200")
    W, H = wl["W"], wl["H"]
    model = make_model(cu, wl)
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    base = {"metric": "pixel-iterations/s", "unit": "pixel-iterations/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if wl["double"] else "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": args.workload + ": " + wl["desc"], "width": W, "height": H, "max_iterations": wl["maxIter"],
                       "max_super_sampling": wl["maxSS"], "adaptive_ss": bool(wl["flags"] & A)}}
    ref_ok = have_gpu and oracle.REFRUN_LIB.exists() and (oracle.REF_DIR / ("%s.%s.cubin" % (wl["fractal"], args.ref_kind))).exists()
    if ref_ok:
        # work count of this workload: exact integer from this backend's device counter (bit-exact with the reference
        # by the parity tests); taken once, outside the timed region.  The timed path below is reference code only.
        prov = cu.CudaFractalRendererProvider(device=0)
        r = prov.getRenderer(wl["fractal"], False)
        if wl["fractal"] == "julia":
            r.setFractalCustomParams("%r;%r" % tuple(wl["julia_c"]))
        r.initializeRendering(W, H, None, cu.OUTPUT_DEVICE)
        r.renderQuality(model)
        iters_per_step = r.stats().pixel_iterations
        r.close(); prov.close()
        sampler = ClockSampler(0)
        with oracle.RefRun(wl["fractal"], args.ref_kind) as rr:
            if wl["fractal"] == "julia":
                import numpy as np
                rr.write_constant("julia_c", np.array(wl["julia_c"], dtype=np.float64).tobytes())
            sampler.start()
            wall_ms, main_ms, comp_ms, _ = rr.frames(W, H, model.planeSegment, wl["maxIter"], wl["maxSS"], wl["flags"],
                                                     oracle.default_palette(), wl["double"], args.warmup, args.steps, True)
            clocks = sampler.stop()
            wall_dev, main_dev, comp_dev, _ = rr.frames(W, H, model.planeSegment, wl["maxIter"], wl["maxSS"], wl["flags"],
                                                        oracle.default_palette(), wl["double"], 1, args.steps, False)
        v_e2e = iters_per_step * args.steps / (wall_ms * 1e-3)
        v_dev = iters_per_step * args.steps / (wall_dev * 1e-3)
        base.update({"value": v_dev, "ms_per_step": wall_dev / args.steps, "frames_per_s": args.steps / (wall_dev * 1e-3),
                     "e2e": {"value": v_e2e, "unit": "pixel-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": W * H * 4,
                             "ms_per_step": wall_ms / args.steps},
                     "gpu_launches": 2 * args.steps, "clocks": clocks,
                     "device_ms_per_step": {"render_kernel": main_dev / args.steps, "compose_kernel": comp_dev / args.steps},
                     "cpu_baseline": {"value": v_e2e, "unit": "pixel-iterations/s", "cores": 0, "kind": "reference",
                                      "sample": "NOT a CPU run: the reference ships no CPU path; these are its own CUDA kernels "
                                                "(oracle/_ref/%s.%s.cubin, reference sources compiled by nvcc 12.9 for sm_100a) on the same B200, "
                                                "block 32x32 / grid ceil(W/32) x ceil(H/32) / sync after every launch as in CudaFractalRenderer.java; "
                                                "whole frame, %d steps" % (wl["fractal"], args.ref_kind, args.steps)},
                     "work_count_from": "device counter of this backend for the same frame (parity-verified), outside the timed region"})
        return base
    # no GPU or no prebuilt reference modules: fall back to the oracle port on the host cores
    cores = os.cpu_count() or 1
    tot_it, tot_s = 0, 0.0
    for _ in range(args.warmup + args.steps):
        it_c, _, sec = oracle.render_main_rows_threaded(wl["fractal"], W, H, model.planeSegment, wl["maxIter"], wl["maxSS"], wl["flags"],
                                                        wl["double"], row_stride=args.cpu_row_stride * 4, threads=cores,
                                                        julia_c=wl.get("julia_c", (0.0, 0.0)))
        tot_it, tot_s = tot_it + it_c, tot_s + sec
    v = tot_it / tot_s
    base.update({"value": v, "ms_per_step": tot_s * 1e3 / (args.warmup + args.steps),
                 "e2e": {"value": v, "unit": "pixel-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0,
                 "cpu_baseline": {"value": v, "unit": "pixel-iterations/s", "cores": cores, "kind": "port",
                                  "sample": "every %dth vote-tile row of the frame per step, oracle port on %d threads (reference kernels unavailable: no GPU or oracle/_ref missing)"
                                            % (args.cpu_row_stride * 4, cores)}})
    return base


# ---------------------------------------------------------------------------------------------------------
# config c3: real-time zoom sequence (fast frames: reuse/reprojection + foveated resampling + compose)
# ---------------------------------------------------------------------------------------------------------
HBM_BYTES_PER_PIXEL_FAST_FRAME = 52   # reuse pass 16 R + 16 W, compose 16 R + 4 W


def hbm_peak_gbs():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth on this pool)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s from B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def zoom_segments(cu, wl, n):
    m = cu.RenderingModel(canvasWidth=wl["W"], canvasHeight=wl["H"])
    m.setPlaneSegmentFromCenter(wl["center"][0], wl["center"][1], wl["zoom"])
    segs = [list(m.planeSegment)]
    for _ in range(n - 1):
        m.zoomAt(wl["focus"], True)
        segs.append(list(m.planeSegment))
    return segs


def zoom_model(cu, wl, segment):
    m = cu.RenderingModel(canvasWidth=wl["W"], canvasHeight=wl["H"])
    m.planeSegment = list(segment)
    m.maxIterations = wl["maxIter"]
    m.maxSuperSampling = wl["maxSS"]
    fl = wl["flags"]
    m.useAdaptiveSuperSampling = bool(fl & A)
    m.useFoveatedRendering = bool(fl & FOV)
    m.useSampleReuse = bool(fl & REUSE)
    m.zooming = bool(fl & ZOOMING)
    m.zoomingIn = bool(fl & ZOOM_IN)
    m.mouseFocus = tuple(wl["focus"])
    return m


def run_zoom_ours(args, wl, rank, world, local):
    """every rank renders the whole sequence (the reuse pass needs the previous frame around every pixel; the
    sequence is not partitioned in this round): N > 1 = independent replicas, scaling weak"""
    import torch
    cu = importlib.import_module("chaos-ultra_b200")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = wl["W"], wl["H"]
    prov = cu.CudaFractalRendererProvider(device=local)
    r = prov.getRenderer(wl["fractal"], False)
    segs = zoom_segments(cu, wl, 1 + args.warmup + args.steps)

    def loop(mode, sampler=None):
        if r.getState() == cu.STATE_READY_TO_RENDER:
            r.freeRenderingResources()
        r.initializeRendering(W, H, None, mode)
        m = zoom_model(cu, wl, segs[0])
        m.maxSuperSampling = max(1.0, wl["maxSS"])
        r.renderQuality(m)                                  # frame 0
        acc = dict(iters=0, launches=0, render_ms=0.0, compose_ms=0.0, reuse_ms=0.0, frame_ms=0.0, precisions=set())
        for f in range(1, 1 + args.warmup + args.steps):
            if f == 1 + args.warmup:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                if sampler is not None:
                    sampler.start()
                t0 = time.perf_counter()
                acc = dict(iters=0, launches=0, render_ms=0.0, compose_ms=0.0, reuse_ms=0.0, frame_ms=0.0, precisions=set())
            m = zoom_model(cu, wl, segs[f])
            r.renderFast(m)
            st = r.stats()
            acc["iters"] += st.pixel_iterations
            acc["launches"] += st.kernel_launches
            acc["render_ms"] += st.render_ms
            acc["compose_ms"] += st.compose_ms
            acc["reuse_ms"] += st.reuse_ms
            acc["frame_ms"] += st.frame_ms
            acc["precisions"].add(m.floatingPointPrecision)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        acc["seconds"] = time.perf_counter() - t0
        acc["clocks"] = sampler.stop() if sampler is not None else None
        if world > 1:
            t = torch.tensor([acc["seconds"], acc["frame_ms"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            acc["seconds"], acc["frame_ms"] = t.tolist()
        acc["device_seconds"] = acc["frame_ms"] * 1e-3
        return acc

    dev = loop(cu.OUTPUT_DEVICE, ClockSampler(local) if rank == 0 else None)
    e2e = loop(cu.OUTPUT_HOST)
    out = None
    if rank == 0:
        px = W * H
        peak, peak_src = hbm_peak_gbs()
        mem_s = (dev["reuse_ms"] + dev["compose_ms"]) * 1e-3 / args.steps   # the two memory passes; the sampling pass is compute
        # With CHAOS_FUSE_FAST=2 the reuse pass colours the pixels it finishes itself also when the frame stays in device
        # memory (by default it does so only for host output): those records are not read back, so the bytes the build
        # MOVES are 16 R + 16 W + 4 W = 36 per pixel, not the reference's 52 -- counted as such.
        fused = os.environ.get("CHAOS_FUSE_FAST") == "2"
        bytes_px = 36 if fused else HBM_BYTES_PER_PIXEL_FAST_FRAME
        achieved = bytes_px * px / mem_s / 1e9
        out = {
            "metric": "4K frames/s (zoom sequence, fast frames)", "value": world * args.steps / dev["device_seconds"], "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev["device_seconds"] * 1e3 / args.steps,
            "wall_ms_per_step": dev["seconds"] * 1e3 / args.steps,
            "timing": "ms_per_step = CUDA events on the stream the kernels are launched on (reuse pass start .. compose end), summed over "
                      "the K frames, MAX over ranks; wall_ms_per_step = host clock between the two synchronize brackets",
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if dev["precisions"] == {0} else ("f64" if 0 not in dev["precisions"] else "f32+f64"), "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "width": W, "height": H, "max_iterations": wl["maxIter"],
                       "max_super_sampling": wl["maxSS"], "parallelism": "1 GPU" if world == 1 else "%d independent replicas" % world,
                       "pixel_iterations_per_step": dev["iters"] // args.steps,
                       "l2": "each frame reads the previous frame's 133 MB record buffer and writes another 133 MB one (> 126 MB L2)"},
            "e2e": {"value": world * args.steps / e2e["seconds"], "unit": "frames/s", "h2d_bytes_per_step": 512,
                    "d2h_bytes_per_step": px * 4 + 32, "ms_per_step": e2e["seconds"] * 1e3 / args.steps,
                    "note": "chaos_render_fast through the C ABI, composed RGBA8 frame written to pinned host memory every frame"},
            "gpu_launches": dev["launches"], "clocks": dev["clocks"],
            "device_ms_per_step": {"reuse_pass": dev["reuse_ms"] / args.steps, "sample_pass": (dev["render_ms"] - dev["reuse_ms"]) / args.steps,
                                   "compose_kernel": dev["compose_ms"] / args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC.get(args.workload, (None, None))[0], "traffic_source": NCU_TRAFFIC.get(args.workload, (None, None))[1],
                         "kernel": "chaosReusePass* + compose (the two memory passes of a fast frame)", "peak_source": peak_src,
                         "algorithmic_bytes": ("%d B/pixel x %d pixels per frame (" % (bytes_px, px)) +
                                              ("reuse pass 16 R + 16 W + 4 W: it colours its own pixels; the reference's two passes move 52)" if fused
                                               else "reuse 16 R + 16 W, compose 16 R + 4 W)"),
                         "note": "the foveal disc and the pixels without history are resampled by a separate compute-bound launch (sample_pass), not part of this figure"},
        }
    r.close()
    prov.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_zoom_reference(args, wl, rank, world, local):
    if rank != 0:
        return None
    import oracle
    cu = importlib.import_module("chaos-ultra_b200")
    W, H = wl["W"], wl["H"]
    n = 1 + args.warmup + args.steps
    segs = zoom_segments(cu, wl, n)
    doubles = [oracle.choose_precision(sg, W, H) != 0 for sg in segs]
    pal = oracle.default_palette()
    with oracle.RefRun(wl["fractal"], args.ref_kind) as rr:
        sampler = ClockSampler(0)
        sampler.start()
        wall, adv, comp, _ = rr.zoom(W, H, segs, doubles, wl["maxIter"], wl["maxSS"], wl["flags"], wl["focus"], pal, args.warmup, args.steps, True)
        clocks = sampler.stop()
        wall_d, adv_d, comp_d, _ = rr.zoom(W, H, segs, doubles, wl["maxIter"], wl["maxSS"], wl["flags"], wl["focus"], pal, args.warmup, args.steps, False)
    peak, peak_src = hbm_peak_gbs()
    achieved = HBM_BYTES_PER_PIXEL_FAST_FRAME * W * H / ((adv_d + comp_d) * 1e-3 / args.steps) / 1e9
    return {"metric": "4K frames/s (zoom sequence, fast frames)", "value": args.steps / (wall_d * 1e-3), "unit": "frames/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_d / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if all(doubles) else ("f32" if not any(doubles) else "f32+f64"), "data": "synthetic",
            "impl": "reference", "config": {"workload": args.workload + ": " + wl["desc"], "width": W, "height": H},
            "e2e": {"value": args.steps / (wall * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": W * H * 4,
                    "ms_per_step": wall / args.steps},
            "gpu_launches": 2 * args.steps, "clocks": clocks,
            "device_ms_per_step": {"reuse_kernel": adv_d / args.steps, "compose_kernel": comp_d / args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src},
            "cpu_baseline": {"value": args.steps / (wall * 1e-3), "unit": "frames/s", "cores": 0, "kind": "reference",
                             "sample": "NOT a CPU run: the reference's own CUDA kernels (oracle/_ref/%s.%s.cubin) on the same B200 with the "
                                       "Java host's fast-frame sequence (CudaFractalRenderer.renderFast); the reference ships no CPU path"
                                       % (wl["fractal"], args.ref_kind)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 20 frames; c3: 119 fast frames (the 120-frame zoom sequence)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--engine", type=int, default=None)
    ap.add_argument("--ref-kind", default="src", choices=["src", "ptx92"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-gather", action="store_true", help="multi-GPU: gather the bands with NCCL send/recv instead of composing into rank 0's frame")
    ap.add_argument("--no-full-trips", action="store_true", help="skip the CHAOS_SHORTCUTS=0 comparison run")
    ap.add_argument("--cpu-row-stride", type=int, default=2, help="host baseline: every n-th vote-tile row of the frame (bounds the CPU time)")
    args = ap.parse_args()
    args.e2e_host_copy = False
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    quiet_nccl()
    rank, world, local = dist_env()
    wl = WORKLOADS[args.workload]
    zoom = wl.get("kind") == "zoom"
    if args.steps is None:
        args.steps = 119 if zoom else 20
    if args.impl == "reference":
        out = (run_zoom_reference if zoom else run_reference)(args, wl, rank, world, local)
    else:
        out = (run_zoom_ours if zoom else run_ours)(args, wl, rank, world, local)
    if rank == 0 and out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
