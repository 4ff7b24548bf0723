#!/usr/bin/env python
"""How many fork()+exit()+wait() cycles per second the box sustains with W forking processes,
for a bare interpreter and for one shaped like the zygote (numpy + proxy layer imported).
Separates the kernel's process-creation scalability from anything in the executor.
    python tools/fork_scale.py [--iters 300] [--workers 1,2,4,8,16,32,64]"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def worker(kind, iters, start, out):
    if kind == "zygote":
        import numpy  # noqa: F401
        from bee_code_interpreter_b200 import script_runner  # noqa: F401
    start.wait()
    t0 = time.perf_counter()
    fork_s = 0.0
    for _ in range(iters):
        a = time.perf_counter()
        pid = os.fork()
        if pid == 0:
            os._exit(0)
        fork_s += time.perf_counter() - a
        os.waitpid(pid, 0)
    out.put((time.perf_counter() - t0, fork_s))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--workers", default="1,2,4,8,16,32,64")
    args = ap.parse_args()
    ctx = mp.get_context("spawn")
    rows = []
    for kind in ("bare", "zygote"):
        for w in [int(x) for x in args.workers.split(",")]:
            start, out = ctx.Event(), ctx.Queue()
            ps = [ctx.Process(target=worker, args=(kind, args.iters, start, out)) for _ in range(w)]
            for p in ps:
                p.start()
            time.sleep(1.5 if kind == "zygote" else 0.5)
            start.set()
            res = [out.get() for _ in ps]
            for p in ps:
                p.join()
            wall = max(r[0] for r in res)
            row = {"kind": kind, "workers": w, "forks_per_s": w * args.iters / wall,
                   "fork_call_ms": 1e3 * sum(r[1] for r in res) / (w * args.iters)}
            rows.append(row)
            print(row, flush=True)
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
