"""Measures the FP pipe peaks used as roofline denominators (bench_kernels/peak.cubin)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cuda.bindings import driver as cu


def ok(res):
    if res[0] != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(res[0]))
    return res[1] if len(res) == 2 else res[1:]


def run(name, ops_per_trip, dtype, blocks_per_sm, trips):
    ok(cu.cuInit(0))
    dev = ok(cu.cuDeviceGet(0))
    ctx = ok(cu.cuDevicePrimaryCtxRetain(dev))
    ok(cu.cuCtxPushCurrent(ctx))
    sms = ok(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev))
    mod = ok(cu.cuModuleLoad(str(ROOT / "bench_kernels" / "peak.cubin").encode()))
    fn = ok(cu.cuModuleGetFunction(mod, name.encode()))
    blocks, threads = sms * blocks_per_sm, 256
    out = ok(cu.cuMemAlloc(blocks * threads * 8))
    e0, e1 = ok(cu.cuEventCreate(0)), ok(cu.cuEventCreate(0))
    args = (np.array([int(out)], dtype=np.uint64), np.array([trips], dtype=np.uint32), np.array([1.0], dtype=dtype))
    argp = np.array([a.ctypes.data for a in args], dtype=np.uint64)
    best = 1e30
    for it in range(5):
        ok(cu.cuEventRecord(e0, 0))
        ok(cu.cuLaunchKernel(fn, blocks, 1, 1, threads, 1, 1, 0, 0, argp.ctypes.data, 0))
        ok(cu.cuEventRecord(e1, 0))
        ok(cu.cuEventSynchronize(e1))
        ms = ok(cu.cuEventElapsedTime(e0, e1))
        if it:
            best = min(best, ms)
    ok(cu.cuMemFree(out))
    rate = blocks * threads * trips * ops_per_trip / (best * 1e-3)
    print("%-16s blocks/SM %d  %8.3f ms  %8.1f G lane-ops/s" % (name, blocks_per_sm, best, rate / 1e9), flush=True)
    cu.cuCtxPopCurrent()
    return rate


if __name__ == "__main__":
    for bps in (1, 2, 4, 8):
        run("peak_fp64", 8, np.float64, bps, 1 << 15)
    for bps in (1, 2, 3, 4, 6, 8):
        run("peak_mandel_mix", 24, np.float64, bps, 1 << 14)
    run("peak_fp32", 8, np.float32, 8, 1 << 16)
