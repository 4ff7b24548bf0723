"""N>1 path on CPU: world_size-2 gloo processes exercise request sharding, result gathering
and the split-operand combines of bee_code_interpreter_b200/parallel.py."""
import json
import socket
import subprocess
import sys
from pathlib import Path

from bee_code_interpreter_b200 import parallel as par


def test_shard_and_chunk_bounds_cover_everything():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in par.shard(1024 + 5, r, world))
        assert seen == list(range(1029))
        assert all(par.owner(i, world) == i % world for i in range(40))
        bounds = [par.chunk_bounds(1_000_003, r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == 1_000_003
        assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in bounds]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_run(built):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    worker = Path(__file__).resolve().parent / "dist_worker.py"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(worker)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0]
    out = json.loads(line[len("RESULT "):])
    assert out["owners"] == [i % 2 for i in range(11)]
    for key in ("round_robin_ok", "sum_ok", "mean_ok", "max_ok", "min_ok", "argmax_ok", "argmin_ok", "cumsum_ok",
                "nan_ok", "int_ok", "sort_ok", "sort_balanced", "sort_dups_ok", "sort_empty_rank_ok", "matmul_ok"):
        assert out[key] is True, (key, out)
