"""Multi-GPU parity (-m gpu, needs >= 2 devices): one process per GPU-segment over the NCCL interconnect."""
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
