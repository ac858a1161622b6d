"""Inter-process interconnect parity (-m gpu): one process per GPU-segment.

test_ranks_share_one_device runs on ANY box, the single-GPU one included: the interconnect is the peer-memory windows with
device-side signalling, bootstrapped through a gloo all-gather (cbgpu_motion_create_boot, no NCCL communicator), and the
two or three rank processes share device 0 - CUDA IPC maps one process's window into the other exactly as between two
GPUs, and the signalling kernels of the processes take turns on the device.  test_two_ranks_golden needs >= 2 devices and
adds the NCCL transport (staged path, exact-size redo after an overflowed window)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    from cloudberry_b200 import capi
    return capi.gpu().cbgpu_device_count()


@pytest.mark.parametrize("motion", ["p2p", "nccl"])
@pytest.mark.parametrize("replicated", ["1", "0"])
def test_two_ranks_golden(replicated, motion):
    """motion=p2p: Redistribute fused into the sender slice's kernel over peer memory (falls back to the
    staged path by itself where CUDA IPC / P2P is unavailable); motion=nccl: the staged path forced."""
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    env = dict(os.environ, CB_REPLICATED=replicated, CBGPU_MOTION=motion)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "multirank_worker.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MULTIRANK PASS" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("replicated", ["1", "0"])
def test_ranks_share_one_device(replicated, world):
    """golden Q1 / Q3 / Q5 through Redistribute / Gather Motions, load-time DISTRIBUTED BY, an asymmetric NULL map, a
    90 % skewed Motion and a segment failing mid-query, between PROCESSES over the peer-memory windows"""
    env = dict(os.environ, CB_REPLICATED=replicated, CB_BOOT="gloo", CBGPU_MOTION_WINDOW_MB="512", CBGPU_MOTION_TIMEOUT_MS="60000")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29621 + world), os.path.join(ROOT, "tests", "multirank_worker.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "MULTIRANK PASS" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def test_window_overflow_is_redone_staged():
    """>= 2 GPUs: a window too small for the skewed Motion -> CBGPU_DX_OVERFLOW seen by every rank -> the same Motion
    redone over NCCL with exactly sized buffers (instrument.motion_repartitions), same rows"""
    if _ngpus() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, CB_REPLICATED="0", CBGPU_MOTION="p2p", CBGPU_MOTION_WINDOW_MB="160", CB_SKEW_ROWS="6000000")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tests", "multirank_worker.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "MULTIRANK PASS" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
