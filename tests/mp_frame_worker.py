"""Worker of tests/test_multi_gpu.py: one rank of an N-process render, launched by torch.distributed.run.

Every rank renders its bands of the same frame; the frame is assembled by the compose kernels (host shared memory or
rank 0's device frame) under the library's frame barrier, and rank 0 compares it with the unpartitioned frame it rendered
first.  Exit code 0 = every check passed on every rank.  With fewer devices than ranks all ranks share device 0."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    import cases
    import helpers
    cu = importlib.import_module("chaos-ultra_b200")
    part = importlib.import_module("chaos-ultra_b200.partition")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ["LOCAL_RANK"]) if torch.cuda.device_count() >= world else 0
    dist.init_process_group("gloo")
    mode = sys.argv[1]
    if mode == "steal":
        steal_tiles(cu, part, dist, rank, world, dev)
        dist.destroy_process_group()
        return
    if mode.startswith("zoom"):
        zoom_sequence(cu, part, dist, rank, world, dev, mode)
        dist.destroy_process_group()
        return
    case = dict(cases.MAIN_CASES[2], W=1531, H=1077, image=cases.seg(-0.5, 0.0, 2.0, 1531, 1077), maxIter=2100)   # ragged, several samples
    W, H, band = case["W"], case["H"], 32
    with cu.CudaFractalRendererProvider(device=dev) as prov:
        r = prov.getRenderer("mandelbrot", False)
        r.initializeRendering(W, H, None, cu.OUTPUT_DEVICE)
        m = helpers.model_for(cu, case)
        r.renderQuality(m)
        whole, whole_it = r.outputRGBA().copy(), r.stats().pixel_iterations
        job = part.JobSharedMemory(rank, world, H, W, dist)
        r.setPartition(rank, world, band)
        if mode == "host":
            job.attach(r, host_target=True, barrier=True)
        else:
            assert part.share_frame_native(r, rank, world, dist)
            job.attach(r, host_target=False, barrier=True)
        its = []
        for frame in range(3):          # several frames: the barrier counts announcements, not a flag
            r.renderQuality(helpers.model_for(cu, case))
            its.append(r.stats().pixel_iterations)
            if rank == 0:               # on return every rank's bands are in the frame
                got = job.frame if mode == "host" else r.outputRGBA()
                assert np.array_equal(got, whole), "frame %d assembled from %d ranks differs from the whole frame" % (frame, world)
            dist.barrier()              # nobody composes the next frame over the one rank 0 is looking at
        t = torch.tensor([its[-1]], dtype=torch.int64)
        dist.all_reduce(t)
        assert int(t.item()) == whole_it, "pixel-iterations of the parts %d != whole frame %d" % (int(t.item()), whole_it)
        # a fast frame of a partitioned renderer renders the own bands afresh (include/chaos_ultra.h)
        r.renderFast(helpers.model_for(cu, case))
        assert r.stats().reuse_ms == 0
        dist.barrier()
        r.setFrameBarrier(0, 0)
        if mode == "host":
            r.setHostTarget(0, 0)
        else:
            r.setOutputTarget(0)
        job.close(dist)
    dist.destroy_process_group()


def steal_tiles(cu, part, dist, rank, world, dev):
    """one-sample frame (the deep-zoom configuration) with a deliberately lopsided deal: contiguous slabs of a view whose
    expensive part lies in the first slab.  The ranks that finish early must take tiles from the others' cursors (records
    stored into the owner's buffer); the assembled frame, every rank's own records and the total work must not change."""
    import cases
    import helpers
    import torch
    W, H = 1024, 768
    case = dict(name="mpsteal", fractal="mandelbrot", W=W, H=H, image=cases.seg(-0.6, 0.55, 1.2, W, H), maxIter=20000, maxSS=1.0, flags=0,
                double=True, julia_c=(0.0, 0.0), amplifier=10)       # the set fills the upper part of the view, the lower part escapes at once
    with cu.CudaFractalRendererProvider(device=dev) as prov:
        r = prov.getRenderer("mandelbrot", False)
        r.initializeRendering(W, H, None, cu.OUTPUT_DEVICE)
        r.renderQuality(helpers.model_for(cu, case))
        whole_rec, whole_rgba, whole_it = r.downloadRecords(), r.outputRGBA().copy(), r.stats().pixel_iterations
        slab = part.slab_rows(H, world)
        r.setPartition(rank, world, slab)
        job = part.JobSharedMemory(rank, world, H, W, dist)
        job.attach(r, host_target=True, barrier=True)
        part.share_records(r, rank, world, dist)
        r0, r1 = rank * slab, min(H, (rank + 1) * slab)
        stolen = 0
        for frame in range(4):
            r.renderQuality(helpers.model_for(cu, case))
            st = r.stats()
            helpers.assert_records_equal(r.downloadRecords()[r0:r1], whole_rec[r0:r1], "rank %d frame %d: own rows (some written by other ranks)" % (rank, frame))
            t = torch.tensor([st.pixel_iterations, st.foreign_orbits], dtype=torch.int64)
            dist.all_reduce(t)
            assert int(t[0]) == whole_it, "frame %d: work of the ranks %d != whole frame %d" % (frame, int(t[0]), whole_it)
            stolen += int(t[1])
            if rank == 0:
                assert np.array_equal(job.frame, whole_rgba), "frame %d differs" % frame
            dist.barrier()
        # ranks that share ONE device are time-sliced, their kernels rarely overlap and there is nobody to steal from; on
        # devices of their own the lopsided deal must move tiles
        if torch.cuda.device_count() >= world:
            assert stolen > 0, "no tile changed ranks although the deal was lopsided"
        if rank == 0:
            print("tile stealing: %d orbits changed ranks over 4 frames (%d ranks on %d devices)" % (stolen, world, min(world, torch.cuda.device_count())))
        r.setFrameBarrier(0, 0)
        r.setHostTarget(0, 0)
        job.close(dist)


def zoom_sequence(cu, part, dist, rank, world, dev, mode):
    """multi-GPU fast frames: one slab of rows per rank, previous-frame taps into other slabs are peer loads; every frame's
    records (own slab) and RGBA (whole frame, rank 0) must equal those of the same sequence rendered whole by one renderer"""
    import cases
    import helpers
    W, H, focus = 643, 362, (200, 301)                     # ragged; the focus near the bottom: origins cross slab edges in both directions
    flags = cases.A | cases.FOV | cases.REUSE | cases.ZOOMING | (cases.ZOOM_IN if mode != "zoom_out" else 0)
    base = dict(name="mpzoom", fractal="mandelbrot", W=W, H=H, maxIter=600, maxSS=2.0, flags=flags, double=(mode == "zoom_f64"),
                julia_c=(0.0, 0.0), amplifier=10, focus=focus)
    segs = [cases.seg(-0.748, 0.1, 2.0 if not base["double"] else 1.0e-4, W, H)]
    for _ in range(7):
        segs.append(cases.zoom_at(segs[-1], W, H, focus, mode != "zoom_out"))
    with cu.CudaFractalRendererProvider(device=dev) as prov:
        r = prov.getRenderer("mandelbrot", False)
        r.initializeRendering(W, H, None, cu.OUTPUT_DEVICE)
        whole = []
        r.renderQuality(helpers.model_for(cu, dict(base, image=segs[0])))
        whole.append((r.downloadRecords(), r.outputRGBA().copy(), r.stats().pixel_iterations))
        for sg in segs[1:]:
            r.renderFast(helpers.model_for(cu, dict(base, image=sg)))
            whole.append((r.downloadRecords(), r.outputRGBA().copy(), r.stats().pixel_iterations))
        assert whole[3][0]["isReused"].mean() > 0.8
        # the same sequence on `world` ranks
        r.freeRenderingResources()
        r.initializeRendering(W, H, None, cu.OUTPUT_DEVICE)
        slab = part.slab_rows(H, world)
        r.setPartition(rank, world, slab)
        job = part.JobSharedMemory(rank, world, H, W, dist)
        job.attach(r, host_target=True, barrier=True)
        part.share_records(r, rank, world, dist)
        r0, r1 = rank * slab, min(H, (rank + 1) * slab)
        import torch
        for f, sg in enumerate(segs):
            m = helpers.model_for(cu, dict(base, image=sg))
            (r.renderQuality if f == 0 else r.renderFast)(m)
            st = r.stats()
            if f:
                assert st.reuse_ms > 0, "frame %d was not reprojected" % f
            got = r.downloadRecords()
            helpers.assert_records_equal(got[r0:r1], whole[f][0][r0:r1], "rank %d frame %d rows %d..%d" % (rank, f, r0, r1))
            t = torch.tensor([st.pixel_iterations], dtype=torch.int64)
            dist.all_reduce(t)
            assert int(t.item()) == whole[f][2], "frame %d: pixel-iterations of the slabs %d != whole %d" % (f, int(t.item()), whole[f][2])
            if rank == 0:
                assert np.array_equal(job.frame, whole[f][1]), "frame %d assembled from %d slabs differs" % (f, world)
            dist.barrier()
        r.setFrameBarrier(0, 0)
        r.setHostTarget(0, 0)
        job.close(dist)


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        print("RANK %s FAILED:\n%s" % (os.environ.get("RANK"), traceback.format_exc()), flush=True)
        raise
