"""Multi-GPU frame partition: row bands dealt round-robin over the ranks, gathered to rank 0.

New functionality (the reference is hard-wired to device 0, cudarenderer/CudaHelpers.java:36-41).  Every pixel of a
quality frame is independent and a vote tile is 8x4 pixels, so the frame is cut into bands of ``band_rows`` pixel
rows (a multiple of 4: vote tiles never straddle two GPUs), band ``b`` belongs to rank ``b % world``.  Each rank
renders and composes only its bands (``chaos_set_partition``); the one exchange step of the path is the gather of
the composed RGBA bands into rank 0's frame.  Two ways:

* ``share_frame`` (default on GPUs): rank 0's device frame is mapped into every other rank's process with CUDA IPC and
  set as their compose target (``chaos_set_output_target``), so the bands cross NVLink as the compose kernel's own
  128-bit stores -- compose and gather are one kernel -- and the exchange step shrinks to a completion barrier;
* ``gather_bands``: a band is a contiguous ``rows x width x 4`` byte range, so it is one send/recv pair per band,
  batched.  Runs over NCCL (device tensors) or gloo (CPU tensors, the tests).
"""
from __future__ import annotations

from typing import List, Tuple


def n_bands(height: int, band_rows: int) -> int:
    return (height + band_rows - 1) // band_rows


def band_rows_range(band: int, height: int, band_rows: int) -> Tuple[int, int]:
    return band * band_rows, min(height, (band + 1) * band_rows)


def band_owner(band: int, world: int) -> int:
    return band % world


def rows_owned(rank: int, world: int, height: int, band_rows: int) -> List[Tuple[int, int]]:
    """[(row0, row1)] of the bands rank renders, top to bottom."""
    if band_rows <= 0 or band_rows % 4:
        raise ValueError("band_rows must be a positive multiple of 4 (vote tiles are 4 rows high)")
    return [band_rows_range(b, height, band_rows) for b in range(rank, n_bands(height, band_rows), world)]


def tiles_owned(rank: int, world: int, width: int, height: int, band_rows: int) -> int:
    """number of 8x4 vote tiles in the rank's bands (what chaos_render_args.n_tiles holds)"""
    tiles_x = (width + 7) // 8
    return sum(((r1 - r0) + 3) // 4 for r0, r1 in rows_owned(rank, world, height, band_rows)) * tiles_x


def gather_plan(rank: int, world: int, height: int, band_rows: int) -> List[Tuple[str, int, int, int]]:
    """[(op, row0, row1, peer)]: rank 0 receives every foreign band from its owner, every other rank sends its
    bands to rank 0.  Order is by band index on both sides, so the batched sends and receives pair up."""
    plan = []
    for b in range(n_bands(height, band_rows)):
        owner = band_owner(b, world)
        if owner == 0:
            continue
        r0, r1 = band_rows_range(b, height, band_rows)
        if rank == 0:
            plan.append(("recv", r0, r1, owner))
        elif rank == owner:
            plan.append(("send", r0, r1, 0))
    return plan


def gather_bands(frame, rank: int, world: int, band_rows: int, dist) -> int:
    """Run the plan on ``frame`` (H x W tensor of packed RGBA, on the device for NCCL, on the CPU for gloo).
    Returns the number of bytes this rank sent or received."""
    if world == 1:
        return 0
    ops, nbytes = [], 0
    for op, r0, r1, peer in gather_plan(rank, world, frame.shape[0], band_rows):
        rows = frame[r0:r1]
        ops.append(dist.P2POp(dist.irecv if op == "recv" else dist.isend, rows, peer))
        nbytes += rows.numel() * rows.element_size()
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return nbytes


def share_frame(renderer, rank: int, world: int, dist):
    """Map rank 0's device frame into the other ranks (CUDA IPC) and make it their compose target.
    Returns an object that must stay alive while the target is in use (close() unmaps), or None if this cannot be
    done here (no cuda.bindings, IPC refused) -- callers then fall back to gather_bands.  Collective: every rank calls it."""
    if world == 1:
        return None
    try:
        from cuda.bindings import driver as cu
    except Exception:
        cu = None
    payload = [None]
    if rank == 0 and cu is not None:
        res, handle = cu.cuIpcGetMemHandle(renderer.outputRGBADevicePointer())
        if res == cu.CUresult.CUDA_SUCCESS:
            payload = [bytes(handle.reserved)]
    dist.broadcast_object_list(payload, src=0)
    ok, mapped = 1, None
    if rank != 0:
        ok = 0
        if cu is not None and payload[0] is not None:
            handle = cu.CUipcMemHandle()
            handle.reserved = payload[0]
            res, ptr = cu.cuIpcOpenMemHandle(handle, cu.CUipcMem_flags.CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS)
            if res == cu.CUresult.CUDA_SUCCESS:
                mapped, ok = int(ptr), 1
    elif payload[0] is None:
        ok = 0
    import torch
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:                     # somebody could not map it: nobody uses it
        if mapped is not None:
            cu.cuIpcCloseMemHandle(mapped)
        return None
    if mapped is not None:
        renderer.setOutputTarget(mapped)
    return _SharedFrame(renderer, mapped, cu)


class _SharedFrame:
    def __init__(self, renderer, mapped, cu):
        self.renderer, self.mapped, self.cu = renderer, mapped, cu

    def close(self):
        if self.mapped is not None:
            try:
                self.renderer.setOutputTarget(0)
            except Exception:
                pass
            self.cu.cuIpcCloseMemHandle(self.mapped)
            self.mapped = None
