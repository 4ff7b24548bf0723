#!/usr/bin/env python
"""Where a request's wall time goes in ONE service process: submit -> script start (zygote queue
+ fork + child set-up), script start -> end (two broker round trips), script end -> result
(exit, capture files, changed-file scan, clean-up).  The script prints its own timestamps.
    python tools/e2e_trace.py [--zygotes 4] [--inflight 32] [--requests 2048] [--n 100000000]"""
import argparse
import asyncio
import json
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

SRC = """import time
t0 = time.time()
import numpy as np
x = np.random.rand({n})
t1 = time.time()
r = np.sum(np.square(x))
t2 = time.time()
print("T", repr(t0), repr(t1), repr(t2), float(r))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zygotes", type=int, default=4)
    ap.add_argument("--inflight", type=int, default=32)
    ap.add_argument("--requests", type=int, default=2048)
    ap.add_argument("--n", type=int, default=10**8)
    ap.add_argument("--procs", type=int, default=1, help="service processes sharing one broker")
    ap.add_argument("--broker-socket", default=None)
    args = ap.parse_args()
    from bee_code_interpreter_b200.executor import LocalGpuExecutor

    if args.procs > 1:
        import subprocess

        from bee_code_interpreter_b200.broker import BrokerProcess

        broker = BrokerProcess(device=0)
        cmd = [sys.executable, __file__, "--zygotes", str(args.zygotes), "--inflight", str(args.inflight),
               "--requests", str(args.requests), "--n", str(args.n), "--broker-socket", broker.socket_path]
        t0 = time.time()
        ps = [subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True) for _ in range(args.procs)]
        outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in ps]
        broker.close()
        for o in outs:
            print(json.dumps(o))
        print(json.dumps({"procs": args.procs, "sum_req_per_s": sum(o["req_per_s"] for o in outs)}))
        return
    if args.broker_socket:
        import types

        shared = types.SimpleNamespace(socket_path=args.broker_socket, close=lambda: None)
        ex = LocalGpuExecutor(broker=shared, max_inflight=args.inflight, timeout=300, zygotes=args.zygotes)
    else:
        ex = LocalGpuExecutor(max_inflight=args.inflight, timeout=300, zygotes=args.zygotes)
    src = SRC.format(n=args.n)
    rows = []

    async def one():
        ts = time.time()
        r = await ex.execute(src)
        te = time.time()
        f = r.stdout.split()
        rows.append((ts, float(f[1]), float(f[2]), float(f[3]), te))

    async def batch(k):
        await asyncio.gather(*(one() for _ in range(k)))

    asyncio.run(batch(4 * args.inflight))
    rows.clear()
    t0 = time.time()
    asyncio.run(batch(args.requests))
    dt = time.time() - t0
    ex.close()
    # requests queue on the executor's semaphore first: measure from the moment a slot was free,
    # i.e. use only the per-phase differences that do not include that wait
    med = lambda v: 1e3 * statistics.median(v)  # noqa: E731
    out = {
        "zygotes": args.zygotes, "inflight": args.inflight, "req_per_s": args.requests / dt,
        "script_rand_ms_p50": med([r[2] - r[1] for r in rows]),
        "script_sum_ms_p50": med([r[3] - r[2] for r in rows]),
        "script_end_to_result_ms_p50": med([r[4] - r[3] for r in rows]),
        "little_latency_ms": 1e3 * args.inflight / (args.requests / dt),
    }
    out["fork_and_setup_ms_est"] = out["little_latency_ms"] - out["script_rand_ms_p50"] - out["script_sum_ms_p50"] - out["script_end_to_result_ms_p50"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
