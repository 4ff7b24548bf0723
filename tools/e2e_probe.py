#!/usr/bin/env python
"""Where does an end-to-end request spend its time?  Run on the GPU box.
  A. broker + GPU only: T threads, each its own broker connection, rand(1e8) + fused sum(square)
  B. executor spawn path only: `pass` scripts (no GPU work)
  C. the metric script through the executor, script-reported compute time vs request latency
"""
from __future__ import annotations

import asyncio
import json
import statistics
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from bee_code_interpreter_b200 import _lib as L  # noqa: E402
from bee_code_interpreter_b200 import workloads as W  # noqa: E402
from bee_code_interpreter_b200.broker import BrokerProcess  # noqa: E402
from bee_code_interpreter_b200.broker_client import RemoteBackend  # noqa: E402


def probe_broker(sock, threads, per_thread, n):
    lat = []
    lock = threading.Lock()

    def work():
        for _ in range(per_thread):
            t = time.perf_counter()
            rb = RemoteBackend(sock)
            p = rb.rand(n, 1234)
            rb.reduce(L.RED_SUM, np.float64, [p], [n], pre=L.PRE_SQUARE)
            rb.free(p)
            rb.close()
            with lock:
                lat.append(time.perf_counter() - t)

    ts = [threading.Thread(target=work) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return {"threads": threads, "req_per_s": threads * per_thread / dt, "p50_ms": 1e3 * statistics.median(lat)}


async def run_many(ex, src, k):
    lat = []

    async def one():
        t = time.perf_counter()
        r = await ex.execute(src)
        lat.append(time.perf_counter() - t)
        return r

    t0 = time.perf_counter()
    res = await asyncio.gather(*(one() for _ in range(k)))
    return res, time.perf_counter() - t0, lat


def main():
    from bee_code_interpreter_b200.executor import LocalGpuExecutor
    import types

    out = {}
    broker = BrokerProcess(device=0)
    n = 10**8
    try:
        probe_broker(broker.socket_path, 8, 4, n)  # warm-up
        out["A_broker_gpu_only"] = [probe_broker(broker.socket_path, t, 16, n) for t in (16, 64, 128)]
        shared = types.SimpleNamespace(socket_path=broker.socket_path, close=lambda: None)
        for zy, infl in ((4, 32), (8, 64), (16, 128)):
            ex = LocalGpuExecutor(broker=shared, max_inflight=infl, timeout=300, zygotes=zy)
            asyncio.run(run_many(ex, "pass", 64))
            _, dt, _ = asyncio.run(run_many(ex, "pass", 1024))
            out.setdefault("B_spawn_only", []).append({"zygotes": zy, "inflight": infl, "req_per_s": 1024 / dt})
            src = W.benchmark_numpy_source(n)
            asyncio.run(run_many(ex, src, 2 * infl))
            res, dt, lat = asyncio.run(run_many(ex, src, 1024))
            comp = [float(r.stdout.splitlines()[1].split()[2]) for r in res if r.exit_code == 0]
            out.setdefault("C_metric_script", []).append({
                "zygotes": zy, "inflight": infl, "req_per_s": 1024 / dt,
                "script_compute_p50_ms": 1e3 * statistics.median(comp), "script_compute_max_ms": 1e3 * max(comp),
                "ok": len(comp)})
            ex.close()
        out["broker_stats"] = RemoteBackend(broker.socket_path).ping()["stats"]
    finally:
        broker.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
