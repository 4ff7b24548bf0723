#!/usr/bin/env python
"""Broker + GPU only, clients in several PROCESSES (no forks per request, no client-side GIL
contention): P processes x T threads, each thread loops rand(n) -> sum(square(x)) -> free.
Separates what the broker sustains at a given number of requests in flight from what the
per-request process costs.   python tools/broker_probe.py [P,T ...]"""
import json
import multiprocessing as mp
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


RECONNECT = "--reconnect" in sys.argv


def client(sock, threads, per_thread, n, start, out):
    import numpy as np
    from bee_code_interpreter_b200 import numpy_proxy as P
    from bee_code_interpreter_b200.broker_client import RemoteBackend

    lat = []

    def work():
        be = RemoteBackend(sock)
        for i in range(per_thread):
            t = time.perf_counter()
            if RECONNECT:            # like a script process: one connection per request
                be.close()
                be = RemoteBackend(sock)
            x = P.device_rand(be, (n,), seed=i)
            float(np.sum(np.square(x)))
            del x
            lat.append(time.perf_counter() - t)
        be.close()

    ts = [threading.Thread(target=work) for _ in range(threads)]
    start.wait()
    t0 = time.time()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    out.put((t0, time.time(), lat))


def main():
    from bee_code_interpreter_b200.broker import BrokerProcess

    configs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:] if not a.startswith("--")] or [(4, 4), (4, 8), (4, 16), (4, 32), (8, 16), (16, 8)]
    n = 10**8
    broker = BrokerProcess(device=0)
    ctx = mp.get_context("spawn")
    rows = []
    try:
        for P_, T in [(2, 4)] + configs:   # first = warm-up
            start, out = ctx.Event(), ctx.Queue()
            per = max(8, 2048 // (P_ * T))
            ps = [ctx.Process(target=client, args=(broker.socket_path, T, per, n, start, out)) for _ in range(P_)]
            for p in ps:
                p.start()
            time.sleep(2.0)
            start.set()
            res = [out.get() for _ in ps]
            for p in ps:
                p.join()
            dt = max(r[1] for r in res) - min(r[0] for r in res)
            lat = sorted(x for r in res for x in r[2])
            rows.append({"procs": P_, "threads": T, "in_flight": P_ * T, "req_per_s": P_ * T * per / dt, "p50_ms": 1e3 * lat[len(lat) // 2]})
            print(rows[-1], flush=True)
    finally:
        broker.close()
    print(json.dumps(rows[1:]))


if __name__ == "__main__":
    main()
