"""GPU, 2 ranks (run with ``gpurun --gpus 2``; skipped on a single-GPU box): configs[4] as written -- views
block-sharded over the ranks, NCCL all_gather of the per-view records -- must give records BIT-IDENTICAL to the
single-GPU run (the SHA-256 of the gathered records, wall-clock fields zeroed), including a view count that does
not divide by the world size."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(nproc, views):
    args = ["bench.py", "--workload", "sweep64", "--views", str(views), "--height", "48", "--width", "64", "--members", "3",
            "--gpus", str(nproc)]
    if nproc == 1:
        cmd = [sys.executable] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
               "--master-addr", "127.0.0.1", "--master-port", str(_port())] + args
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("views", [8, 7])
def test_sharded_sweep_is_bit_identical_to_single_gpu(built_library, views):
    one, two = _run(1, views), _run(2, views)
    assert one["sweep"]["views_aggregated"] == views and two["sweep"]["views_aggregated"] == views
    assert two["n_gpus"] == 2 and two["scaling"] == "strong"
    assert one["sweep"]["records_sha256"] == two["sweep"]["records_sha256"]
    assert one["sweep"]["check"] == two["sweep"]["check"]
