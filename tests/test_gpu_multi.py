"""GPU (-m gpu), needs >= 2 devices: the product's multi-GPU path -- lanes sharded over ranks by LaneSharder, results
gathered with NCCL from the device buffers tvs_fit wrote -- launched as the driver launches bench.py (torchrun, one
process per GPU).  Skips on a single-GPU box; tests/test_sharding_gloo.py covers the same host logic on the CPU.
Reference flow replaced: traversome.py:501-617 (serial bootstrap loop)."""
import os
import socket
import subprocess
import sys

import pytest

from traversome_b200 import _cabi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_nccl_sharded_flow_and_gather(lib_built):
    n = _cabi.device_count()
    if n < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    if n < 2:
        pytest.skip("needs >= 2 GPUs (NCCL all_gather across ranks)")
    world = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = (out.stdout + out.stderr)[-4000:]
    assert out.returncode == 0, tail
    assert out.stdout.count("support table identical") == world, tail
    assert out.stdout.count("own slice bitwise True") == world, tail
    assert out.stdout.count("equal to one at a time True, from host weights True") == world, tail
