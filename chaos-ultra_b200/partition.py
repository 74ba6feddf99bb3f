"""Multi-GPU frame partition: row bands dealt round-robin over the ranks, gathered to rank 0.

New functionality (the reference is hard-wired to device 0, cudarenderer/CudaHelpers.java:36-41).  Every pixel of a
quality frame is independent and a vote tile is 8x4 pixels, so the frame is cut into bands of ``band_rows`` pixel
rows (a multiple of 4: vote tiles never straddle two GPUs), band ``b`` belongs to rank ``b % world``.  Each rank
renders and composes only its bands (``chaos_set_partition``); the one exchange step of the path is the gather of
the composed RGBA bands into rank 0's frame -- a band is a contiguous ``rows x width x 4`` byte range, so it is one
send/recv pair per band, batched.  The same plan runs over NCCL (device tensors, NVLink) in bench.py and over gloo
(CPU tensors) in the tests.
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
