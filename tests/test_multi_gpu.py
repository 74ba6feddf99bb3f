"""Multi-process rendering of ONE frame (one process per rank, as bench.py --gpus N runs it): partitioned bands composed
into a frame in host shared memory (every GPU over its own PCIe link) or into rank 0's device frame (CUDA IPC, NVLink),
under the library's frame barrier; rank 0 checks the assembled frame against the unpartitioned one.  With one visible
device both ranks share it (CUDA IPC and host registration work between processes on one device)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _failure(res):
    """what the worker said (it prints its own traceback), not torchrun's summary of exit codes"""
    said = [ln for ln in res.stdout.splitlines() if ln.strip()]
    i = next((k for k, ln in enumerate(said) if "FAILED:" in ln), max(0, len(said) - 30))
    return "\n".join(said[i:i + 40]) + "\n" + res.stderr[-600:]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["host", "ipc"])
@pytest.mark.parametrize("world", [2, 3])
def test_frame_assembled_by_several_processes_equals_the_whole_frame(mode, world):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "mp_frame_worker.py"), mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert res.returncode == 0, _failure(res)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["zoom", "zoom_f64", "zoom_out"])
@pytest.mark.parametrize("world", [2, 3])
def test_zoom_sequence_on_several_ranks_equals_the_whole_sequence(mode, world):
    """multi-GPU fast frames: one slab per rank, taps into other slabs are peer loads from the owner's record buffer"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "mp_frame_worker.py"), mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert res.returncode == 0, _failure(res)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_tiles_are_stolen_across_ranks_and_nothing_changes(world):
    """dynamic redistribution of a one-sample frame: ranks that run out of tiles take them from the other ranks' cursors"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "mp_frame_worker.py"), "steal"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert res.returncode == 0, _failure(res)
    assert "tile stealing:" in res.stdout
