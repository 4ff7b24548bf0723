#!/usr/bin/env python
"""Host-side scaling of the executor's process model: W service processes, each with its own
zygotes, running a trivial script (`pass`) or the metric script.  Shows where the box's
fork/exec/tmpfs capacity sits relative to what the GPUs can absorb.
    python tools/spawn_scale.py [--script pass|metric] [--workers 1,2,4,8,16]
"""
from __future__ import annotations

import argparse
import asyncio
import json
import subprocess
import sys
import time
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def worker(args):
    from bee_code_interpreter_b200.executor import LocalGpuExecutor
    from bee_code_interpreter_b200 import workloads as W

    src = "pass" if args.script == "pass" else W.benchmark_numpy_source(10**8)
    shared = types.SimpleNamespace(socket_path=args.broker_socket, close=lambda: None)
    ex = LocalGpuExecutor(broker=shared, max_inflight=args.inflight, timeout=300, zygotes=args.zygotes)

    async def batch(k):
        return await asyncio.gather(*(ex.execute(src) for _ in range(k)))

    asyncio.run(batch(64))
    print("ready", flush=True)
    sys.stdin.readline()
    t0 = time.time()
    res = asyncio.run(batch(args.requests))
    t1 = time.time()
    ok = sum(r.exit_code == 0 for r in res)
    ex.close()
    print(json.dumps({"t0": t0, "t1": t1, "ok": ok}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--script", default="pass")
    ap.add_argument("--workers", default="1,2,4,8,16")
    ap.add_argument("--requests", type=int, default=1024)
    ap.add_argument("--inflight", type=int, default=32)
    ap.add_argument("--zygotes", type=int, default=4)
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--broker-socket", default=None)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    if args.worker:
        return worker(args)
    from bee_code_interpreter_b200.broker import BrokerProcess

    broker = BrokerProcess(device=args.device)
    out = []
    try:
        for w in [int(x) for x in args.workers.split(",")]:
            procs = [subprocess.Popen([sys.executable, __file__, "--worker", "--broker-socket", broker.socket_path,
                                       "--script", args.script, "--requests", str(args.requests), "--inflight", str(args.inflight),
                                       "--zygotes", str(args.zygotes)], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
                     for _ in range(w)]
            for p in procs:
                assert p.stdout.readline().strip() == "ready"
            for p in procs:
                p.stdin.write("go\n")
                p.stdin.flush()
            rs = [json.loads(p.stdout.readline()) for p in procs]
            for p in procs:
                p.wait(timeout=30)
            dt = max(r["t1"] for r in rs) - min(r["t0"] for r in rs)
            out.append({"workers": w, "req_per_s": w * args.requests / dt, "ok": sum(r["ok"] for r in rs)})
            print(out[-1], flush=True)
    finally:
        broker.close()
    print(json.dumps({"script": args.script, "rows": out}))


if __name__ == "__main__":
    main()
