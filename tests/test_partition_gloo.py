"""Multi-rank host logic on CPU: the row-band partition covers the frame exactly once, agrees with what the C
library computes for chaos_render_args.n_tiles, and the gather plan, run over gloo with world_size 2 and 3, assembles
the full frame on rank 0 (the same code path bench.py runs over NCCL)."""
import importlib
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
part = importlib.import_module("chaos-ultra_b200.partition")


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("H,band", [(2160, 32), (8192, 64), (117, 4), (35, 8), (1, 4), (1077, 36)])
def test_bands_cover_every_row_once(world, H, band):
    seen = [0] * H
    for rank in range(world):
        for r0, r1 in part.rows_owned(rank, world, H, band):
            assert r0 % 4 == 0 and r0 < r1 <= H
            for y in range(r0, r1):
                seen[y] += 1
    assert seen == [1] * H
    W = 203
    total = sum(part.tiles_owned(r, world, W, H, band) for r in range(world))
    assert total == ((W + 7) // 8) * ((H + 3) // 4)


def test_band_rows_must_be_multiple_of_four():
    with pytest.raises(ValueError):
        part.rows_owned(0, 2, 100, 6)


def test_plan_pairs_up():
    for world in (2, 4, 8):
        H, band = 2160, 32
        recvs = [(r0, r1, peer) for op, r0, r1, peer in part.gather_plan(0, world, H, band)]
        sends = []
        for rank in range(1, world):
            sends += [(r0, r1, rank) for op, r0, r1, peer in part.gather_plan(rank, world, H, band) if op == "send" and peer == 0]
        assert sorted(recvs) == sorted(sends)
        assert all(op == "recv" for op, *_ in part.gather_plan(0, world, H, band))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, band, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frame = torch.zeros((H, W), dtype=torch.int32)
        for r0, r1 in part.rows_owned(rank, world, H, band):
            # what a rank's compose pass leaves: its own bands filled, the rest untouched
            ys = torch.arange(r0, r1, dtype=torch.int32).unsqueeze(1)
            xs = torch.arange(W, dtype=torch.int32).unsqueeze(0)
            frame[r0:r1] = (rank + 1) * 1000000 + ys * 1000 + xs
        nbytes = part.gather_bands(frame, rank, world, band, dist)
        # max-over-ranks timing reduction used by bench.py, on the gloo backend
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = True
        if rank == 0:
            ys = torch.arange(H, dtype=torch.int32).unsqueeze(1)
            xs = torch.arange(W, dtype=torch.int32).unsqueeze(0)
            owner = (ys // band) % world
            want = (owner + 1) * 1000000 + ys * 1000 + xs
            ok = bool((frame == want).all())
        q.put((rank, ok, nbytes, t.item()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W,band", [(2, 117, 64, 4), (2, 216, 96, 32), (3, 100, 40, 8)])
def test_gather_over_gloo(world, H, W, band):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, band, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert all(t == float(world) for *_, t in res)
    foreign = sum((r1 - r0) * W * 4 for rank in range(1, world) for r0, r1 in part.rows_owned(rank, world, H, band))
    assert res[0][2] == foreign                      # rank 0 received every foreign band exactly once
    assert sum(r[2] for r in res[1:]) == foreign


def test_a_ranks_bands_subdivide_into_strands_and_parts():
    """chaos_abi.cpp runs a rank's frame as G strands / K parts: sub-partition k of rank q is partition q + world * k of
    world * K.  Those must be exactly rank q's bands, each once (the kernels and the compose filter only know
    `band % part_count == part_index`)."""
    part = importlib.import_module("chaos-ultra_b200.partition")
    for world in (1, 2, 3, 8):
        for sub in (2, 3, 4, 8):
            for height, band_rows in ((2160, 32), (2161, 32), (8192, 64), (130, 4)):
                for rank in range(world):
                    mine = part.rows_owned(rank, world, height, band_rows)
                    pieces = []
                    for k in range(sub):
                        pieces += part.rows_owned(rank + world * k, world * sub, height, band_rows)
                    assert sorted(pieces) == sorted(mine)
                    assert sum(part.tiles_owned(rank + world * k, world * sub, 3840, height, band_rows) for k in range(sub)) == \
                        part.tiles_owned(rank, world, 3840, height, band_rows)
