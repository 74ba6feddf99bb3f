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


class JobSharedMemory:
    """Host shared memory of an N-process job: [ 64-byte barrier block | pad to a page | RGBA frame ].

    Rank 0 creates the segment, the others attach by name (``dist.broadcast_object_list``).  The frame part is what
    ``chaos_set_host_target`` takes: every rank's compose kernel writes its bands there over its own PCIe link, so the
    whole frame of an N-GPU render arrives in host memory N links wide and nothing is gathered on a device first.  The
    barrier block is what ``chaos_set_frame_barrier`` takes.  Collective: every rank constructs it."""

    PAGE = 4096

    def __init__(self, rank: int, world: int, height: int, width: int, dist=None):
        from multiprocessing import resource_tracker, shared_memory
        import numpy as np
        self.rank, self.world = rank, world
        self.frame_bytes = height * width * 4
        size = self.PAGE + ((self.frame_bytes + self.PAGE - 1) // self.PAGE) * self.PAGE
        name = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=size)
            self.shm.buf[:self.PAGE] = bytes(self.PAGE)
            name[0] = self.shm.name
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:    # the creator unlinks it; an attaching process must not (Python < 3.13 registers it regardless)
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        base = np.frombuffer(self.shm.buf, dtype=np.uint8)
        self.barrier_address = base.ctypes.data
        self.frame_address = base.ctypes.data + self.PAGE
        self.frame = base[self.PAGE:self.PAGE + self.frame_bytes].view(np.uint32).reshape(height, width)
        self._base = base

    def attach(self, renderer, host_target: bool = True, barrier: bool = True):
        if host_target:
            renderer.setHostTarget(self.frame_address, self.frame_bytes)
        if barrier:
            renderer.setFrameBarrier(self.barrier_address, self.world)

    def close(self, dist=None):
        self.frame = self._base = None
        if dist is not None and self.world > 1:
            dist.barrier()
        try:
            self.shm.close()
        except BufferError:
            pass
        if self.rank == 0:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass


def share_frame_native(renderer, rank: int, world: int, dist) -> bool:
    """Rank 0 exports its device frame (chaos_ipc_export_frame), the others open it as their compose target
    (chaos_ipc_open_frame).  Collective.  False if some rank could not map it (nobody then uses it)."""
    import torch
    payload = [None]
    if rank == 0:
        try:
            payload = [renderer.exportFrameHandle()]
        except Exception:
            payload = [None]
    dist.broadcast_object_list(payload, src=0)
    ok = 1 if payload[0] is not None else 0
    if rank != 0 and ok:
        try:
            renderer.openFrameHandle(payload[0])
        except Exception:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        if rank != 0:
            renderer.setOutputTarget(0)
        return False
    return True


def slab_rows(height: int, world: int) -> int:
    """rows of one rank's slab: the frame cut into `world` contiguous slabs whose height is a multiple of the 4-row vote tile"""
    return ((height + world - 1) // world + 3) // 4 * 4


def share_records(renderer, rank: int, world: int, dist) -> None:
    """Multi-GPU fast frames: every rank exports its two record buffers and opens every other rank's pair
    (chaos_ipc_export_records / chaos_ipc_open_records).  Collective; raises where a mapping is refused."""
    mine = renderer.exportRecordHandles()
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    for q, handles in enumerate(everyone):
        if q != rank:
            renderer.openRecordHandles(q, handles)


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
