#!/usr/bin/env python
"""bench.py -- the headline benchmark of the chaos-ultra B200 render backend.

A *step* is one frame of the hot path through the reference-facing C ABI.  The headline workload is configs[1] of
BASELINE.json: mandelbrot 3840x2160, maxIter 10000, adaptive supersampling (maxSS 8), FP64, full-set view.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference] [--no-extras]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

ONE JSON line on stdout (rank 0).  Keys beyond the base contract:
  value        pixel-iterations/s (the reference loop's trip count of the frame per second), whole job, frame in HBM
  executed_per_s   the trips the FP pipe really iterated per second (the rest is PROVEN, see config.work_accounting)
  e2e          the same metric with the composed RGBA frame of every step in HOST memory: one GPU -- compose writes the
               library's pinned frame; N GPUs -- every rank's compose writes its bands over its own PCIe link into one
               frame in host shared memory; e2e.rgba_crc32 is ASSERTED against tests/golden/workloads.json
  roofline     the iteration kernels against the FP64/FP32 FMA issue peak measured in this run (bench_kernels/peak.cubin)
  extra        the other BASELINE.json configurations in the same record (c4 deep zoom at every N, c3/c1/c5 at N = 1)
  cpu_baseline the oracle port on the host cores, bounded sample; baselines.reference_ptx92 the reference's shipped
               CUDA 9.2 PTX on this GPU (N = 1)
--impl reference runs the reference's OWN kernels (oracle/_ref: src/main/cuda compiled by nvcc 12.9 for sm_100a, and
the shipped PTX) with the Java host's launch sequence; that process never loads this backend's library.
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
import zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

A, FOV, REUSE, ZOOMING, ZOOM_IN = 1, 4, 8, 16, 32

# roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel(s), per frame, from one
# `ncu --set full` capture of the same command (never measured in a bench run); workload -> (bytes, capture)
NCU_TRAFFIC = {
    "c2": (int((0.17 + 171.61 + 140.50 + 29.68 + 19.51 + 0.05 + 0.06 + 0.0 + 7.59 + 23.50 + 76.16 + 27.01 + 58.58 + 0.27) * 1e6),
           "profiles/r03_passes_c2.txt: chaosProbe + chaosLong + chaosFinish of passes A and C + chaosPassB, Double (133 MB of that are the records "
           "themselves; the rest is the long list -- 32 B per surviving orbit, written by the probe and read by the long kernel -- the finish list "
           "and the export arrays; HBM is 2 % busy)"),
    "c3": (int((127.10 + 85.13 + 132.73 + 6.94) * 1e6), "profiles/r02_fast_frame_c3.txt: chaosReusePassFloat + compose "
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


def load_constants():
    """tests/golden/workloads.json: exact work count and frame checksum of every workload (make_workload_constants.py)"""
    try:
        return json.loads((ROOT / "tests" / "golden" / "workloads.json").read_text())
    except Exception:
        return {}


def frame_crc(frame):
    import numpy as np
    return zlib.crc32(np.ascontiguousarray(frame).tobytes()) & 0xFFFFFFFF


def workload_config(name, wl):
    """the `config` object: identical in both arms (the driver compares them)"""
    return {"workload": name + ": " + wl["desc"], "width": wl["W"], "height": wl["H"], "max_iterations": wl["maxIter"],
            "max_super_sampling": wl["maxSS"], "adaptive_ss": bool(wl["flags"] & A),
            "l2": "no input is re-read between steps (inputs are viewport scalars); the %d MB of records written per step exceed the 126 MB L2"
                  % (wl["W"] * wl["H"] * 16 // 1000000) if wl["W"] * wl["H"] * 16 > 126e6 else
                  "inputs are viewport scalars, nothing is re-read between steps; every step rewrites all %d MB of records" % (wl["W"] * wl["H"] * 16 // 1000000)}


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


BAND_ROWS = int(os.environ.get("CHAOS_BENCH_BAND_ROWS", "32"))   # multi-GPU quality frames: rows per band (diagnostics: the environment overrides)


class Job:
    """one rank of the run: device, process group, provider, host shared memory of the job"""

    def __init__(self, rank, world, local):
        import torch
        self.torch = torch
        self.rank, self.world, self.local = rank, world, local
        torch.cuda.set_device(local)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            self.dist = dist
        self.cu = importlib.import_module("chaos-ultra_b200")
        self.part = importlib.import_module("chaos-ultra_b200.partition")
        self.prov = self.cu.CudaFractalRendererProvider(kernels_dir=os.environ.get("CHAOS_KERNELS_DIR"), device=local)   # (diagnostic builds)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def gather(self, value):
        """one float per rank -> the list of all ranks' values"""
        if self.dist is None:
            return [float(value)]
        t = self.torch.zeros(self.world, device="cuda", dtype=self.torch.float64)
        t[self.rank] = float(value)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def reduce(self, values, op="max"):
        if self.dist is None:
            return list(values)
        t = self.torch.tensor(list(values), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return t.tolist()

    def renderer(self, wl):
        r = self.prov.getRenderer(wl["fractal"], False)
        if wl["fractal"] == "julia":
            r.setFractalCustomParams("%r;%r" % tuple(wl["julia_c"]))
        return r

    def close(self):
        self.prov.close()
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_quality_loop(job, wl, to_host, steps, warmup, sampler=None):
    """W untimed + K timed quality frames.  One GPU: the renderer's own frame (device memory, or the library's pinned host
    frame when to_host).  N GPUs: every rank renders and composes its row bands straight into ONE frame -- rank 0's device
    frame (CUDA IPC: the bands cross NVLink as the compose kernel's stores) or, when to_host, a frame in host shared
    memory that every GPU writes over its own PCIe link -- and every render call ends at the library's frame barrier.
    Device time of a step = the slowest rank's frame (CUDA events on the launching stream, first kernel .. compose end);
    wall time = host clock between the barrier + synchronize brackets, MAX over ranks."""
    cu, world, rank = job.cu, job.world, job.rank
    W, H = wl["W"], wl["H"]
    r = job.renderer(wl)
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    model = make_model(cu, wl)
    shm = None
    if world == 1:
        r.initializeRendering(W, H, None, cu.OUTPUT_HOST if to_host else cu.OUTPUT_DEVICE)
        if os.environ.get("CHAOS_EMULATE_PART"):   # diagnostics: what ONE rank of an N-GPU run does
            pi, pn = map(int, os.environ["CHAOS_EMULATE_PART"].split(":"))
            r.setPartition(pi, pn, BAND_ROWS)
        exchange = "none"
    else:
        r.initializeRendering(W, H, None, cu.OUTPUT_DEVICE)
        r.setPartition(rank, world, BAND_ROWS)
        shm = job.part.JobSharedMemory(rank, world, H, W, job.dist)
        if to_host:
            shm.attach(r, host_target=True, barrier=True)
            exchange = "every rank's compose kernel writes its bands into one frame in host shared memory over its own PCIe link; frame barrier in host shared memory"
        else:
            if not job.part.share_frame_native(r, rank, world, job.dist):
                raise RuntimeError("CUDA IPC refused: cannot map rank 0's frame")
            shm.attach(r, host_target=False, barrier=True)
            exchange = "every rank's compose kernel writes its bands into rank 0's device frame over NVLink (CUDA IPC); frame barrier in host shared memory"
        # every rank's record buffers and tile cursor mapped everywhere: one-sample frames redistribute their tiles dynamically
        job.part.share_records(r, rank, world, job.dist)
        if round(wl["maxSS"]) <= 1:
            exchange += "; tiles of ranks that are still busy are taken over by ranks that are done (cursors and records over NVLink)"
    per_step = []
    acc = dict(iters=0, skipped=0, launches=0, render_ms=0.0, compose_ms=0.0, foreign=0)
    t0 = 0.0
    for it in range(warmup + steps):
        if it == warmup:
            job.barrier()
            if sampler is not None:
                sampler.start()
            t0 = time.perf_counter()
        r.renderQuality(model)
        if it >= warmup:
            st = r.stats()
            per_step.append(st.frame_ms)
            acc["iters"] += st.pixel_iterations; acc["skipped"] += st.skipped_iterations; acc["launches"] += st.kernel_launches
            acc["render_ms"] += st.render_ms; acc["compose_ms"] += st.compose_ms; acc["foreign"] += st.foreign_orbits
    job.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler is not None else None
    frame = None
    if rank == 0:      # the frame of the last step, for the checksum (outside the timed region)
        frame = (shm.frame.copy() if (shm is not None and to_host) else r.outputRGBA().copy())
    job.barrier()
    slowest = job.reduce(per_step, "max")                    # per step: the slowest rank's frame
    per_rank = job.gather(sum(per_step) / max(1, len(per_step)))
    wall, render_ms, compose_ms = job.reduce([wall, acc["render_ms"], acc["compose_ms"]], "max")
    iters, skipped, launches, foreign = [int(v) for v in job.reduce([acc["iters"], acc["skipped"], acc["launches"], acc["foreign"]], "sum")]
    if shm is not None:
        r.setFrameBarrier(0, 0)
        r.setOutputTarget(0)
        shm.close(job.dist)
    r.freeRenderingResources()
    return dict(device_seconds=sum(slowest) * 1e-3, seconds=wall, iters=iters, skipped=skipped, launches=launches, render_ms=render_ms,
                compose_ms=compose_ms, clocks=clocks, exchange=exchange, frame=frame, steps=steps, foreign=foreign, per_rank_ms=per_rank)


def check_frame(name, frame, constants, what):
    """the frame that was timed must be THE frame of this workload (tests/golden/workloads.json, cross-checked against the
    reference kernels by tests/test_bench_constants_gpu.py); a wrong frame voids the number, so it is an error"""
    crc = frame_crc(frame)
    want = constants.get(name, {}).get("rgba_crc32")
    if os.environ.get("CHAOS_EMULATE_PART"):       # diagnostics: one rank's bands only, the rest of the frame is not rendered
        want = None
    if want is not None and crc != want:
        raise SystemExit("bench: %s frame of %s has crc32 %08x, expected %08x -- the timed frame is wrong" % (what, name, crc, want))
    return {"rgba_crc32": crc, "rgba_crc32_expected": want, "rgba_crc32_ok": None if want is None else crc == want}


def quality_block(job, name, wl, steps, warmup, constants, sampler=None):
    """device-resident and end-to-end figures of one quality-frame workload, checked against the committed constants"""
    dev = timed_quality_loop(job, wl, False, steps, warmup, sampler)
    e2e = timed_quality_loop(job, wl, True, steps, warmup)
    out = None
    if job.rank == 0:
        px = wl["W"] * wl["H"]
        want_iters = constants.get(name, {}).get("pixel_iterations")
        if want_iters is not None and dev["iters"] != want_iters * steps and not os.environ.get("CHAOS_EMULATE_PART"):
            raise SystemExit("bench: %s counted %d pixel-iterations per step, expected %d" % (name, dev["iters"] // steps, want_iters))
        out = {"ms_per_step": dev["device_seconds"] * 1e3 / steps, "value": dev["iters"] / dev["device_seconds"], "unit": "pixel-iterations/s",
               "executed_per_s": (dev["iters"] - dev["skipped"]) / dev["device_seconds"], "frames_per_s": steps / dev["device_seconds"],
               "wall_ms_per_step": dev["seconds"] * 1e3 / steps, "steps": steps, "warmup": warmup,
               "pixel_iterations_per_step": dev["iters"] // steps, "executed_pixel_iterations_per_step": (dev["iters"] - dev["skipped"]) // steps,
               "device_ms_per_step": {"render_kernels": dev["render_ms"] / steps, "compose_kernel": dev["compose_ms"] / steps},
               "frame_check": check_frame(name, dev["frame"], constants, "device"),
               "e2e": dict({"value": e2e["iters"] / e2e["seconds"], "unit": "pixel-iterations/s", "ms_per_step": e2e["seconds"] * 1e3 / steps,
                            "frames_per_s": steps / e2e["seconds"], "h2d_bytes_per_step": 512 * job.world, "d2h_bytes_per_step": px * 4 + 32 * job.world},
                           **check_frame(name, e2e["frame"], constants, "end-to-end")),
               "orbits_redistributed_per_step": dev["foreign"] // steps,
               "mean_frame_ms_per_rank": [round(v, 4) for v in dev["per_rank_ms"]],
               "gpu_launches": dev["launches"], "parallelism": "1 GPU" if job.world == 1 else
               "row bands of %d px dealt round-robin over %d GPUs; %s" % (BAND_ROWS, job.world, dev["exchange"]),
               "e2e_exchange": e2e["exchange"]}
    return out, dev


def run_ours(args, wl, rank, world, local):
    job = Job(rank, world, local)
    cu = job.cu
    if args.engine is not None:
        os.environ["CHAOS_ENGINE"] = str(args.engine)
    constants = load_constants()
    W, H = wl["W"], wl["H"]
    sampler = ClockSampler(local) if rank == 0 else None
    head, dev = quality_block(job, args.workload, wl, args.steps, args.warmup, constants, sampler)
    # the same frame with every trip executed and tested (CHAOS_SHORTCUTS=0): what the iteration kernels do at full work
    full = None
    if world == 1 and not args.no_full_trips and os.environ.get("CHAOS_SHORTCUTS") is None:
        os.environ["CHAOS_SHORTCUTS"] = "0"
        try:
            job.prov.getRenderer(wl["fractal"], True)        # the knob is read when a renderer is opened
            full = timed_quality_loop(job, wl, False, max(2, min(args.steps, 5)), 3)
        finally:
            os.environ.pop("CHAOS_SHORTCUTS", None)
            job.prov.getRenderer(wl["fractal"], True)
    # the other configurations of BASELINE.json in the same record
    extra = {}
    if args.extras and args.workload == "c2":
        blk, _ = quality_block(job, "c4", WORKLOADS["c4"], 5, 3, constants)
        if rank == 0:
            extra["c4"] = dict(blk, workload="c4: " + WORKLOADS["c4"]["desc"])
        blk = zoom_block(job, "c3", WORKLOADS["c3"], 116, 3, constants)     # frame 0 + 3 warm-up + 116 timed fast frames = the 120-frame sequence
        if rank == 0:
            extra["c3"] = blk
        if world == 1:
            extra["c3_closed_loop"] = closed_loop_block(job, WORKLOADS["c3"])
            for name in ("c1", "c5"):
                blk, _ = quality_block(job, name, WORKLOADS[name], 20, 3, constants)
                extra[name] = dict(blk, workload=name + ": " + WORKLOADS[name]["desc"])
    out = None
    if rank == 0:
        executed = dev["iters"] - dev["skipped"]
        out = {
            "metric": "pixel-iterations/s", "value": head["value"], "unit": "pixel-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if wl["double"] else "f32", "data": "synthetic", "frames_per_s": head["frames_per_s"],
            "executed_per_s": head["executed_per_s"], "wall_ms_per_step": head["wall_ms_per_step"],
            "timing": "ms_per_step = CUDA events on the stream the kernels are launched on (first render kernel .. compose end) per step, "
                      "the slowest rank's frame of every step, summed over the K steps; wall_ms_per_step = host clock between the two "
                      "barrier + synchronize brackets, MAX over ranks (several GPUs: it includes the frame barrier of every step)",
            "config": workload_config(args.workload, wl),
            "details": {"parallelism": head["parallelism"], "pixel_iterations_per_step": head["pixel_iterations_per_step"],
                        "executed_pixel_iterations_per_step": head["executed_pixel_iterations_per_step"],
                        "work_accounting": "pixel_iterations = trips of the reference's loop for this frame (inside points count maxIterations). "
                                           "Orbits whose state recurs bit for bit are PROVEN never to escape and stop early with the same result; "
                                           "executed_pixel_iterations is what the FP pipe actually iterated. CHAOS_SHORTCUTS=0 executes every trip (see roofline.full_trips).",
                        "engine": os.environ.get("CHAOS_ENGINE", "default (2: orbit streams)"), "shortcuts": os.environ.get("CHAOS_SHORTCUTS", "default (3)"),
                        "frame_check": head["frame_check"], "e2e_exchange": head["e2e_exchange"]},
            "e2e": dict(head["e2e"], note="chaos_render_quality through the C ABI, host clock around K synchronous calls; inputs are the chaos_params viewport struct "
                                          "(kernel parameter space), result = the composed RGBA8 frame in host memory, checksum asserted"),
            "gpu_launches": head["gpu_launches"], "clocks": dev["clocks"], "device_ms_per_step": head["device_ms_per_step"],
        }
        if extra:
            out["extra"] = extra
        # roofline of the iteration kernels: the FP pipe.  peak = FMA lane-ops/s measured just now on this device.  achieved counts
        # the FP instructions the kernels issue AT THE LEAST: 5 per executed trip (the untested scaled form; tested trips issue 6,
        # unscaled ones up to 7), so frac is a lower bound of the pipe's utilisation and cannot exceed 1.
        try:
            peak = measure_fma_peak(local, wl["double"])
            kernel_s = dev["render_ms"] * 1e-3
            min_ops = 5 if wl["double"] else 6
            achieved = executed / world * min_ops / kernel_s
            multi = round(wl["maxSS"]) > 1
            out["roofline"] = {"bound": "fp64" if wl["double"] else "fp32", "achieved": achieved / 1e9, "peak": peak / 1e9,
                               "unit": "G FP-lane-ops/s", "frac": achieved / peak,
                               "traffic": NCU_TRAFFIC.get(args.workload, (None, None))[0] if world == 1 else None,
                               "traffic_source": NCU_TRAFFIC.get(args.workload, (None, None))[1] if world == 1 else None,
                               "kernel": ("chaosProbe%s + chaosLong%s + chaosFinish%s" + (" (passes A and C of a multi-sample frame) + chaosPassB%s" if multi else ""))
                                         .replace("%s", "Double" if wl["double"] else "Float"),
                               "peak_source": "measured in this run: bench_kernels/peak.cubin, independent FMA chains, best of 5 (MEASURED_PEAKS.json has no FP64/FP32 figure)",
                               "achieved_definition": "%d FP instructions x %d executed pixel-iterations per step / time of the render kernels of a step (CUDA events, first render kernel .. last), max over ranks"
                                                      % (min_ops, executed // args.steps // world),
                               "reference_form": {"ops": "7 FP instructions x %d reference pixel-iterations" % (dev["iters"] // args.steps // world),
                                                  "G_ops_per_s": dev["iters"] / world * 7 / kernel_s / 1e9,
                                                  "ratio_to_peak": dev["iters"] / world * 7 / kernel_s / peak,
                                                  "note": "work the reference's loop needs for this frame per second of these kernels; exceeds 1 because trips are proven instead of executed"},
                               "note": "tensor cores and HBM do not bound these kernels (16 B stored per pixel)"}
            if full is not None:
                fk = full["render_ms"] * 1e-3
                out["roofline"]["full_trips"] = {
                    "ms_per_step": full["device_seconds"] * 1e3 / full["steps"], "render_kernels_ms": full["render_ms"] / full["steps"],
                    "pixel_iterations_per_s": full["iters"] / full["device_seconds"],
                    "frac": full["iters"] * (6 if wl["double"] else 7) / fk / peak,
                    "note": "same frame, CHAOS_SHORTCUTS=0: every trip executed with its test (%d FP instructions per trip issued); "
                            "this is the pipe utilisation of the iteration kernels at the reference's full work" % (6 if wl["double"] else 7)}
        except Exception as e:  # measurement helper failed: say so rather than invent a peak
            out["roofline"] = {"bound": "fp64" if wl["double"] else "fp32", "achieved": None, "peak": None, "unit": "G FP-lane-ops/s",
                               "frac": None, "traffic": None, "error": repr(e)}
    # the two baselines north_star asks for, same run, N = 1 only (both execute oracle/: the cpu_baseline leg)
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        import oracle
        cores = os.cpu_count() or 1
        stride = args.cpu_row_stride
        model = make_model(cu, wl)
        it_c, smp_c, sec = oracle.render_main_rows_threaded(wl["fractal"], W, H, model.planeSegment, wl["maxIter"], wl["maxSS"],
                                                           wl["flags"], wl["double"], row_stride=stride, threads=cores,
                                                           julia_c=wl.get("julia_c", (0.0, 0.0)))
        out["cpu_baseline"] = {"value": it_c / sec, "unit": "pixel-iterations/s", "cores": cores, "kind": "port",
                               "sample": "every %dth vote-tile row of the same frame (%d pixel-iterations, %.1f s), oracle/chaos_oracle.c on %d threads"
                                         % (stride, it_c, sec, cores)}
        try:
            job.prov.close()          # the reference kernels get the device to themselves
            out["baselines"] = {"reference_ptx92": reference_figures(args.workload, wl, "ptx92", 1, max(3, min(args.steps, 5)), constants),
                                "note": "the reference's shipped CUDA 9.2 PTX (src/main/cuda/fractals/*.ptx, assembled for sm_100a) on this GPU with the Java host's "
                                        "launch sequence, timed after this backend's run, outside every timed region; the --impl reference arm times the nvcc 12.9 build of the same sources"}
        except Exception as e:
            out["baselines"] = {"reference_ptx92": None, "error": repr(e)}
    job.close()
    return out


def reference_figures(name, wl, kind, warmup, steps, constants):
    """the reference's own kernels (oracle/_ref/<fractal>.<kind>.cubin) with the Java host's frame sequence; work count from
    the committed constants, so nothing of this backend runs in (or is loaded into) the timing process"""
    import numpy as np
    import oracle
    W, H = wl["W"], wl["H"]
    image = seg(wl["center"][0], wl["center"][1], wl["zoom"], W, H)
    iters_per_step = constants.get(name, {}).get("pixel_iterations")
    with oracle.RefRun(wl["fractal"], kind) as rr:
        if wl["fractal"] == "julia":
            rr.write_constant("julia_c", np.array(wl["julia_c"], dtype=np.float64).tobytes())
        pal = oracle.default_palette()
        wall_ms, main_ms, comp_ms, rgba = rr.frames(W, H, image, wl["maxIter"], wl["maxSS"], wl["flags"], pal, wl["double"], warmup, steps, True)
        wall_dev, main_dev, comp_dev, _ = rr.frames(W, H, image, wl["maxIter"], wl["maxSS"], wl["flags"], pal, wl["double"], 1, steps, False)
    crc = frame_crc(rgba)
    want = constants.get(name, {}).get("rgba_crc32")
    return {"kind": kind, "ms_per_step": wall_dev / steps, "frames_per_s": steps / (wall_dev * 1e-3),
            "value": None if iters_per_step is None else iters_per_step * steps / (wall_dev * 1e-3), "unit": "pixel-iterations/s",
            "e2e": {"value": None if iters_per_step is None else iters_per_step * steps / (wall_ms * 1e-3), "unit": "pixel-iterations/s",
                    "ms_per_step": wall_ms / steps, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": W * H * 4},
            "device_ms_per_step": {"render_kernel": main_dev / steps, "compose_kernel": comp_dev / steps}, "steps": steps,
            "rgba_crc32": crc, "same_frame_as_this_backend": None if want is None else (crc == want if kind == "src" else "ptx92 differs from the src build in c.y's rounding (SURVEY.md 8c)")}


def run_reference(args, wl, rank, world, local):
    """The reference arm: the reference's own kernels on the GPU, launched like the Java host launches them (the reference has
    no CPU implementation of the path).  Rank 0 alone; this backend's library is never loaded here."""
    if rank != 0:
        return None
    import oracle
    constants = load_constants()
    base = {"metric": "pixel-iterations/s", "unit": "pixel-iterations/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if wl["double"] else "f32",
            "data": "synthetic", "impl": "reference", "config": workload_config(args.workload, wl)}
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu and oracle.REFRUN_LIB.exists() and (oracle.REF_DIR / ("%s.%s.cubin" % (wl["fractal"], args.ref_kind))).exists():
        sampler = ClockSampler(0)
        sampler.start()
        fig = reference_figures(args.workload, wl, args.ref_kind, args.warmup, args.steps, constants)
        clocks = sampler.stop()
        base.update({"value": fig["value"], "ms_per_step": fig["ms_per_step"], "frames_per_s": fig["frames_per_s"], "e2e": fig["e2e"],
                     "gpu_launches": 2 * args.steps, "clocks": clocks, "device_ms_per_step": fig["device_ms_per_step"],
                     "rgba_crc32": fig["rgba_crc32"], "same_frame_as_this_backend": fig["same_frame_as_this_backend"],
                     "cpu_baseline": {"value": fig["e2e"]["value"], "unit": "pixel-iterations/s", "cores": 0, "kind": "reference",
                                      "sample": "NOT a CPU run: the reference ships no CPU path; these are its own CUDA kernels "
                                                "(oracle/_ref/%s.%s.cubin, reference sources compiled by nvcc 12.9 for sm_100a) on the same B200, "
                                                "block 32x32 / grid ceil(W/32) x ceil(H/32) / sync after every launch as in CudaFractalRenderer.java; "
                                                "whole frame, %d steps" % (wl["fractal"], args.ref_kind, args.steps)},
                     "work_count_from": "tests/golden/workloads.json (exact trip count of this frame, equal in both implementations by the parity tests)"})
        other = "ptx92" if args.ref_kind == "src" else "src"
        try:
            base["reference_" + other] = reference_figures(args.workload, wl, other, 1, max(3, min(args.steps, 5)), constants)
        except Exception as e:
            base["reference_" + other] = {"error": repr(e)}
        if args.extras and args.workload == "c2":
            base["extra"] = {}
            for name in ("c4", "c1", "c5"):
                try:
                    base["extra"][name] = reference_figures(name, WORKLOADS[name], args.ref_kind, 1, 3 if name == "c4" else 10, constants)
                except Exception as e:
                    base["extra"][name] = {"error": repr(e)}
        return base
    # no GPU or no prebuilt reference modules: the oracle port on the host cores
    cores = os.cpu_count() or 1
    image = seg(wl["center"][0], wl["center"][1], wl["zoom"], wl["W"], wl["H"])
    tot_it, tot_s = 0, 0.0
    for _ in range(args.warmup + args.steps):
        it_c, _, sec = oracle.render_main_rows_threaded(wl["fractal"], wl["W"], wl["H"], image, wl["maxIter"], wl["maxSS"], wl["flags"],
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


def zoom_loop(job, wl, to_host, steps, warmup, sampler=None):
    """frame 0 quality, W untimed + K timed fast frames (the sequence restarts at frame 0 for every loop).  N GPUs: one slab
    of rows per rank; every rank keeps the records of its slab, previous-frame taps into another slab are peer loads from
    the owner's buffer (CUDA IPC, NVLink), every rank composes its slab into the one frame (host shared memory over its own
    PCIe link, or rank 0's device frame), and the library's frame barrier keeps the ranks in step."""
    cu, world, rank = job.cu, job.world, job.rank
    W, H = wl["W"], wl["H"]
    r = job.renderer(wl)
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    shm = None
    if world == 1:
        r.initializeRendering(W, H, None, cu.OUTPUT_HOST if to_host else cu.OUTPUT_DEVICE)
    else:
        r.initializeRendering(W, H, None, cu.OUTPUT_DEVICE)
        r.setPartition(rank, world, job.part.slab_rows(H, world))
        shm = job.part.JobSharedMemory(rank, world, H, W, job.dist)
        if to_host:
            shm.attach(r, host_target=True, barrier=True)
        else:
            if not job.part.share_frame_native(r, rank, world, job.dist):
                raise RuntimeError("CUDA IPC refused: cannot map rank 0's frame")
            shm.attach(r, host_target=False, barrier=True)
        job.part.share_records(r, rank, world, job.dist)
    segs = zoom_segments(cu, wl, 1 + warmup + steps)
    m = zoom_model(cu, wl, segs[0])
    m.maxSuperSampling = max(1.0, wl["maxSS"])
    r.renderQuality(m)
    acc = dict(iters=0, launches=0, render_ms=0.0, compose_ms=0.0, reuse_ms=0.0, precisions=set())
    per_step = []
    t0 = 0.0
    for f in range(1, 1 + warmup + steps):
        if f == 1 + warmup:
            job.barrier()
            if sampler is not None:
                sampler.start()
            t0 = time.perf_counter()
        m = zoom_model(cu, wl, segs[f])
        r.renderFast(m)
        if f >= 1 + warmup:
            st = r.stats()
            per_step.append(st.frame_ms)
            acc["iters"] += st.pixel_iterations; acc["launches"] += st.kernel_launches
            acc["render_ms"] += st.render_ms; acc["compose_ms"] += st.compose_ms; acc["reuse_ms"] += st.reuse_ms
            acc["precisions"].add(m.floatingPointPrecision)
    job.barrier()
    acc["seconds"] = time.perf_counter() - t0
    acc["clocks"] = sampler.stop() if sampler is not None else None
    acc["frame"] = None
    if rank == 0:
        acc["frame"] = shm.frame.copy() if (shm is not None and to_host) else r.outputRGBA().copy()
    job.barrier()
    acc["last_frame_index"] = warmup + steps
    acc["device_seconds"] = sum(job.reduce(per_step, "max")) * 1e-3          # per frame: the slowest rank
    acc["seconds"], acc["render_ms"], acc["compose_ms"], acc["reuse_ms"] = job.reduce([acc["seconds"], acc["render_ms"], acc["compose_ms"], acc["reuse_ms"]], "max")
    acc["iters"], acc["launches"] = [int(v) for v in job.reduce([acc["iters"], acc["launches"]], "sum")]
    if shm is not None:
        r.setFrameBarrier(0, 0)
        r.setOutputTarget(0)
        shm.close(job.dist)
    r.freeRenderingResources()
    return acc


def zoom_block(job, name, wl, steps, warmup, constants, sampler=None):
    dev = zoom_loop(job, wl, False, steps, warmup, sampler)
    e2e = zoom_loop(job, wl, True, steps, warmup)
    if job.rank != 0:
        return None
    px = wl["W"] * wl["H"]
    peak, peak_src = hbm_peak_gbs()
    mem_s = (dev["reuse_ms"] + dev["compose_ms"]) * 1e-3 / steps   # the two memory passes (slowest rank); the sampling pass is compute
    # With CHAOS_FUSE_FAST=2 the reuse pass colours the pixels it finishes itself also when the frame stays in device memory (by
    # default only for host output): those records are not read back, so the bytes the build MOVES are 16 R + 16 W + 4 W = 36 per
    # pixel, not the reference's 52 -- counted as such.
    fused = os.environ.get("CHAOS_FUSE_FAST") == "2"
    bytes_px = 36 if fused else HBM_BYTES_PER_PIXEL_FAST_FRAME
    achieved = bytes_px * px / job.world / mem_s / 1e9              # per GPU: every rank moves its slab
    crcs = constants.get(name, {}).get("frame_crc32") or []

    def check(acc, what):
        crc, idx = frame_crc(acc["frame"]), acc["last_frame_index"]
        want = crcs[idx] if idx < len(crcs) else None
        if want is not None and crc != want:
            raise SystemExit("bench: %s frame %d of %s has crc32 %08x, expected %08x -- the timed sequence is wrong" % (what, idx, name, crc, want))
        return {"frame": idx, "rgba_crc32": crc, "rgba_crc32_expected": want, "rgba_crc32_ok": None if want is None else crc == want}

    return {"workload": name + ": " + wl["desc"], "metric": "4K frames/s (zoom sequence, fast frames)", "value": steps / dev["device_seconds"], "unit": "frames/s",
            "ms_per_step": dev["device_seconds"] * 1e3 / steps, "wall_ms_per_step": dev["seconds"] * 1e3 / steps, "steps": steps, "warmup": warmup,
            "dtype": "f32" if dev["precisions"] == {0} else ("f64" if 0 not in dev["precisions"] else "f32+f64"),
            "pixel_iterations_per_step": dev["iters"] // steps, "gpu_launches": dev["launches"], "clocks": dev["clocks"],
            "parallelism": "1 GPU" if job.world == 1 else "one slab of %d rows per GPU; taps into other slabs are peer loads over NVLink; every GPU composes its slab into the one frame" % job.part.slab_rows(wl["H"], job.world),
            "frame_check": check(dev, "device"),
            "e2e": dict({"value": steps / e2e["seconds"], "unit": "frames/s", "ms_per_step": e2e["seconds"] * 1e3 / steps, "h2d_bytes_per_step": 512 * job.world,
                         "d2h_bytes_per_step": px * 4 + 32 * job.world,
                         "note": "chaos_render_fast through the C ABI, composed RGBA8 frame in host memory every frame" + ("" if job.world == 1 else
                                 " (host shared memory, every GPU writes its slab over its own PCIe link)")}, **check(e2e, "end-to-end")),
            "device_ms_per_step": {"reuse_pass": dev["reuse_ms"] / steps, "sample_pass": (dev["render_ms"] - dev["reuse_ms"]) / steps,
                                   "compose_kernel": dev["compose_ms"] / steps},
            "l2": "each frame reads the previous frame's 133 MB record buffer and writes another 133 MB one (> 126 MB L2)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC.get(name, (None, None))[0], "traffic_source": NCU_TRAFFIC.get(name, (None, None))[1],
                         "kernel": "chaosReusePass* + compose (the two memory passes of a fast frame)", "peak_source": peak_src,
                         "algorithmic_bytes": ("%d B/pixel x %d pixels per frame (" % (bytes_px, px)) +
                                              ("reuse pass 16 R + 16 W + 4 W: it colours its own pixels; the reference's two passes move 52)" if fused
                                               else "reuse 16 R + 16 W, compose 16 R + 4 W)"),
                         "note": "the foveal disc and the pixels without history are resampled by a separate compute-bound launch (sample_pass), not part of this figure"}}


def closed_loop_block(job, wl, ticks=120):
    """The reference's closed loop around the real renderer: the native frame driver (chaos_driver_*) holds
    the mouse button down at the focus for `ticks` animator ticks -- every tick zooms, retargets maxSuperSampling so that a
    frame takes 15 ms of DEVICE time (GLRenderer.java:200-245) and renders a fast frame -- then releases it and refines
    progressively until the machine is back in Waiting.  Reported, not a throughput claim: what the controller does on a B200."""
    cu = job.cu
    drv = importlib.import_module("chaos-ultra_b200.driver")
    W, H = wl["W"], wl["H"]
    r = job.renderer(wl)
    if r.getState() == cu.STATE_READY_TO_RENDER:
        r.freeRenderingResources()
    r.initializeRendering(W, H, None, cu.OUTPUT_HOST)
    m = zoom_model(cu, wl, zoom_segments(cu, wl, 1)[0])
    m.zooming = m.zoomingIn = False
    d = drv.FrameDriver(r, m)
    t0 = time.perf_counter()
    n = d.run_zoom_session(wl["focus"], True, ticks)
    wall = time.perf_counter() - t0
    fast = [e for e in d.log if e[1] == "fast"]
    quality = [e for e in d.log if e[1] == "quality"]
    t_fast = sum(e[3] for e in fast)
    out = {"workload": "c3 view, closed loop: %d ticks with the button down at %s, then progressive refinement" % (ticks, (wl["focus"],)),
           "controller": "native (csrc/chaos_driver.cpp), device clock, target 15 ms per zooming frame, 30 << level ms while refining",
           "frames_rendered": n, "fast_frames": len(fast), "quality_frames": len(quality), "wall_s": wall,
           "zooming_device_ms_per_frame": t_fast / max(1, len(fast)), "zooming_fps_device": 1e3 * len(fast) / max(t_fast, 1e-9),
           "max_super_sampling_while_zooming": [round(e[2], 2) for e in fast[:6]] + ["..."] + [round(e[2], 2) for e in fast[-3:]],
           "refinement": [{"max_super_sampling": round(e[2], 2), "device_ms": round(e[3], 3)} for e in quality],
           "note": "a 4K fast frame costs well under a millisecond on a B200, so the controller drives the sample budget to its cap (64) "
                   "within a few frames: the 15 ms target of the reference (GLRenderer.java:190) is never the binding constraint here"}
    d.close()
    r.freeRenderingResources()
    return out


def run_zoom_ours(args, wl, rank, world, local):
    """headline = the zoom sequence (--workload c3); several GPUs render ONE sequence, a slab of every frame each"""
    job = Job(rank, world, local)
    constants = load_constants()
    blk = zoom_block(job, args.workload, wl, args.steps, args.warmup, constants, ClockSampler(local) if rank == 0 else None)
    out = None
    if rank == 0:
        out = {"metric": blk["metric"], "value": blk["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": blk["ms_per_step"], "wall_ms_per_step": blk["wall_ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": blk["dtype"], "data": "synthetic",
               "timing": "ms_per_step = CUDA events on the stream the kernels are launched on (reuse pass start .. compose end) per frame, the slowest rank's, summed over the K frames",
               "config": dict(workload_config(args.workload, wl), l2=blk["l2"]),
               "details": {"parallelism": blk["parallelism"], "pixel_iterations_per_step": blk["pixel_iterations_per_step"],
                           "frame_check": blk["frame_check"]},
               "e2e": blk["e2e"],
               "gpu_launches": blk["gpu_launches"], "clocks": blk["clocks"], "device_ms_per_step": blk["device_ms_per_step"], "roofline": blk["roofline"]}
    job.close()
    return out


def run_zoom_reference(args, wl, rank, world, local):
    if rank != 0:
        return None
    import oracle
    W, H = wl["W"], wl["H"]
    n = 1 + args.warmup + args.steps
    segs = [seg(wl["center"][0], wl["center"][1], wl["zoom"], W, H)]
    for _ in range(n - 1):
        segs.append(oracle.zoom_at(segs[-1], W, H, wl["focus"], True))
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
            "impl": "reference", "config": dict(workload_config(args.workload, wl), l2="each frame reads the previous frame's 133 MB record buffer and writes another 133 MB one (> 126 MB L2)"),
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
    ap.add_argument("--steps", type=int, default=None, help="default: 20 frames; c3: 116 fast frames (frame 0 + 3 warm-up + 116 timed = the 120-frame zoom sequence)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--engine", type=int, default=None)
    ap.add_argument("--ref-kind", default="src", choices=["src", "ptx92"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="only the headline workload (default: c4 at every N; c1, c5, c3 at N = 1 as well)")
    ap.add_argument("--no-full-trips", action="store_true", help="skip the CHAOS_SHORTCUTS=0 comparison run")
    ap.add_argument("--cpu-row-stride", type=int, default=2, help="host baseline: every n-th vote-tile row of the frame (bounds the CPU time)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    quiet_nccl()
    rank, world, local = dist_env()
    wl = WORKLOADS[args.workload]
    zoom = wl.get("kind") == "zoom"
    if args.steps is None:
        args.steps = 116 if zoom else 20
    if args.impl == "reference":
        out = (run_zoom_reference if zoom else run_reference)(args, wl, rank, world, local)
    else:
        out = (run_zoom_ours if zoom else run_ours)(args, wl, rank, world, local)
    if rank == 0 and out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
