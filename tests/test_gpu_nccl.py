"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): the real NCCL path of row e.  Two ranks,
uneven shards, in-library all-gathers -- results must equal the single-GPU path bit for bit."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_pipeline_equals_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_parity_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("NCCL_PARITY ")]
    assert line, r.stdout[-2000:]
    res = json.loads(line[-1][len("NCCL_PARITY "):])
    assert all(x["pipelined_equal"] for x in res)
    assert all(v == 0 for v in res[0]["diffs"].values()), res[0]["diffs"]
    assert res[0]["planted"] == 1.0
