"""N>1 on real GPUs: one operand split over 2 GPUs, NCCL combines.  The CPU twin
of this test is tests/test_parallel.py (gloo, test double)."""
import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


def test_split_operand_ops_over_two_gpus(dev):
    if dev.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    worker = Path(__file__).resolve().parent / "dist_gpu_worker.py"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(worker)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0]
    out = json.loads(line[len("RESULT "):])
    assert out["world"] == 2 and out["launches"] > 0
    for key in ("sum_ok", "mean_ok", "max_ok", "min_ok", "argmax_ok", "argmin_ok", "cumsum_ok", "sort_ok",
                "sort_total_ok", "matmul_ok"):
        assert out[key] is True, (key, out)
    assert min(out["sort_sizes"]) > 20_000_001 // 8          # sample sort keeps the buckets balanced
