"""Collected multi-GPU tests (-m gpu): skipped on a box with fewer than 2 GPUs.  Each launches one process per GPU with torchrun on
127.0.0.1 and runs a check script of this directory: gene shards + the peer-memory all-reduce inside the kernels must reproduce
the single-GPU chain, all ranks bit-identical (tests/multi_gpu_chain_check.py)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(script, nproc, port, timeout=600):
    env = dict(os.environ)
    env.pop("ZZ_CHAIN_GENE_SUMS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_sharded_device_resident_chain_equals_single_gpu():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on one node")
    r = _torchrun("multi_gpu_chain_check.py", 2, 29500 + os.getpid() % 400)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-GPU device-resident chain OK" in r.stdout and "host mirror, schedule=device on 2 shards OK" in r.stdout
