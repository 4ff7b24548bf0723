#!/usr/bin/env python
"""Sweep of the e2e arm's host-side knobs on one GPU: service processes x admitted requests x warm
interpreters x pre-forked children (BCI_ZYGOTE_POOL).  Prints one JSON line per setting.
    python tools/e2e_sweep.py [--requests 3072] [--grid "4:16:4:4,8:16:2:4,..."]   (workers:inflight:zygotes:pool)"""
import argparse
import json
import os
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=3072)
    ap.add_argument("--n", type=int, default=10**8)
    ap.add_argument("--workload", default="c1")
    ap.add_argument("--grid", default="4:16:4:0,4:16:4:4,8:16:2:4,8:32:2:4,8:32:2:8,12:24:1:6,16:32:1:4,8:64:2:8")
    a = ap.parse_args()
    for cell in a.grid.split(","):
        w, inflight, zyg, pool = (int(x) for x in cell.split(":"))
        os.environ["BCI_ZYGOTE_POOL"] = str(pool)
        args = types.SimpleNamespace(e2e_workers=w, e2e_zygotes=zyg)
        h = bench.E2EHarness(args, 1, 0)
        try:
            r = h.run(a.workload, a.requests, None, 0, a.n, inflight)
        finally:
            h.close()
        print(json.dumps({"workers": w, "inflight_per_gpu": inflight, "zygotes_per_worker": zyg, "prefork_pool": pool,
                          "req_per_s": round(r["value"], 1), "p50_ms": round(1e3 * r["p50_latency_s"], 2), "parity": r["parity"],
                          "on_gpu": r["all_requests_on_gpu"]}), flush=True)


if __name__ == "__main__":
    main()
