"""The multi-GPU paths on REAL peer GPUs (the mailbox protocol had only a one-GPU stand-in test):
torchrun with one process per GPU, NCCL for the plumbing, axb_reduce_sharded over NVLink peer memory. Needs >= 2 visible
GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`); on a one-GPU box the test reports itself
skipped — the single-GPU emulation in test_reduce_gpu.py and the gloo tests still cover the protocol and the host logic."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_reductions_on_peer_gpus(gpu, world):
    import ctypes
    from axiom_b200 import capi
    n = ctypes.c_int32(0)
    capi.lib().axb_device_count(ctypes.byref(n))
    if n.value < world:
        pytest.skip(f"{n.value} GPU(s) visible, this case needs {world}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-3000:] + r.stderr[-3000:]
    j = json.loads(lines[-1])
    assert j["world"] == world and not j["failures"], j["failures"]
    assert j["cases"] >= 11
