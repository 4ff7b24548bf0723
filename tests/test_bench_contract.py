"""bench.py's output contract on the arm that runs without a GPU (`--impl reference`): exactly one
line on stdout, valid JSON, the keys the driver reads; and, under a 2-rank torchrun, rank 0 alone
prints it."""
import json
import socket
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _check(stdout, n_gpus):
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["value"] > 0 and d["higher_is_better"] is True
    assert d["unit"] == "requests/s" and d["dtype"] == "f64" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "requests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_prints_one_json_line(built):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--n", "200000", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    _check(r.stdout, 1)


def test_reference_arm_under_torchrun_only_rank0_prints(built):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--array-elems", "200000",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    _check(r.stdout, 2)


def test_our_arm_control_flow_on_the_cpu_test_double(built):
    """bench.py's own arm with tests/fake_device.py standing in for the GPU binding (BCI_BENCH_DRY: tiny sizes,
    numbers meaningless): one JSON line with the headline keys, `roofline`, `cpu_baseline`, `e2e` through execute(),
    and the `configs` block — C2..C5 each with value, roofline, parity, e2e and cpu_baseline."""
    import os

    env = dict(os.environ, BCI_BENCH_DRY="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1", "--n", "20000", "--requests", "8",
                        "--wave", "4", "--e2e-requests", "8", "--cpu-budget", "1", "--cfg-cpu-budget", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert (KEYS - {"impl"}) | {"roofline", "clocks", "configs"} <= set(d)
    assert d["value"] > 0 and d["gpu_launches"] >= 1 and d["e2e"]["value"] > 0 and d["e2e"]["all_requests_on_gpu"] and d["e2e"]["parity"]
    assert d["roofline"]["kernel"].startswith("reduce_rand_kernel") and {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["configs"]) == {"c2", "c3", "c4", "c5"}
    for name, c in d["configs"].items():
        assert "error" not in c, (name, c)
        assert c["value"] > 0 and c["parity"]["ok"], (name, c["parity"])
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(c["roofline"])
        assert c["e2e"]["value"] > 0 and c["e2e"]["parity"], (name, c["e2e"])
        assert c["cpu_baseline"]["value"] > 0 and c["cpu_baseline"]["kind"] == "port"


def test_our_arm_under_torchrun_on_the_cpu_test_double(built):
    """The multi-rank control flow of bench.py (round-robin sharding, max-over-ranks timing, strong-scaling datum,
    per-rank e2e harness, configs block) with gloo and the test double: rank 0 alone prints the line."""
    import os

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, BCI_BENCH_DRY="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(ROOT / "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
                        "--array-elems", "20000", "--requests", "8", "--wave", "4", "--e2e-requests", "8", "--cpu-budget", "1",
                        "--cfg-cpu-budget", "1"], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["config"]["global_batch"] == 16
    assert d["strong_scaling"]["scaling"] == "strong" and d["strong_scaling"]["global_batch"] == 1024 and d["strong_scaling"]["value"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["requests_per_gpu"] == 8
    for name, c in d["configs"].items():
        assert "error" not in c and c["n_gpus"] == 2 and c["value"] > 0 and c["e2e"]["value"] > 0, (name, c)
