#!/usr/bin/env python
"""What one request costs the host.  One service process (LocalGpuExecutor) runs K requests at a given
concurrency; every script reports its own rusage (user / system CPU, minor faults) and timestamps.
Three scripts: `pass` (host pipeline only), the metric script (one broker round trip), and a metric script that
also reports phase times.  Prints one JSON line per (script, inflight).
    python tools/e2e_cost.py [--requests 512] [--inflight 1,4,16,64] [--zygotes 4]"""
import argparse
import asyncio
import json
import os
import resource
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

TAIL = """
import resource as _r, time as _t
_u = _r.getrusage(_r.RUSAGE_SELF)
print("RU", repr(_t0), repr(_t.time()), _u.ru_utime, _u.ru_stime, _u.ru_minflt, _u.ru_nvcsw, _u.ru_nivcsw)
"""
SCRIPTS = {
    "pass": "import time as _tt\n_t0 = _tt.time()\n" + TAIL,
    "metric": "import time as _tt\n_t0 = _tt.time()\nimport numpy\nx = numpy.random.rand({n})\nr = numpy.sum(numpy.square(x))\nprint('Result:', r)\n" + TAIL,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=512)
    ap.add_argument("--inflight", default="1,4,16,64")
    ap.add_argument("--zygotes", type=int, default=4)
    ap.add_argument("--n", type=int, default=10**8)
    ap.add_argument("--procs", type=int, default=1, help="service processes sharing one broker (each runs --requests at each --inflight)")
    ap.add_argument("--broker-socket", default=None)
    ap.add_argument("--scripts", default="pass,metric")
    args = ap.parse_args()
    from bee_code_interpreter_b200.executor import LocalGpuExecutor

    if args.procs > 1:
        import subprocess

        from bee_code_interpreter_b200.broker import BrokerProcess

        broker = BrokerProcess(device=0)
        try:
            for name in args.scripts.split(","):
                for inflight in args.inflight.split(","):
                    cmd = [sys.executable, __file__, "--requests", str(args.requests), "--inflight", inflight, "--zygotes", str(args.zygotes),
                           "--n", str(args.n), "--broker-socket", broker.socket_path, "--scripts", name]
                    t0 = time.time()
                    ps = [subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True, env=dict(os.environ, BCI_TRACE="1")) for _ in range(args.procs)]
                    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in ps]
                    print(json.dumps({"procs": args.procs, "script": name, "inflight_per_proc": int(inflight),
                                      "sum_req_per_s": round(sum(o["req_per_s"] for o in outs), 1),
                                      "latency_ms_p50": [o["latency_ms"]["p50"] for o in outs][:4], "latency_ms_proc0": outs[0]["latency_ms"],
                                      "wall_s": round(time.time() - t0, 2), "total_requests": args.requests * args.procs,
                                      "start_body_end_ms": [(o["submit_to_script_start_ms_p50"], o["script_body_ms_p50"], o["script_end_to_result_ms_p50"]) for o in outs][:2],
                                      "service_cpu_ms": outs[0]["service_cpu_ms_per_request"], "child_cpu_ms": outs[0]["child_cpu_ms"],
                                      "zygote_phase_ms_mean": outs[0].get("zygote_phase_ms_mean")}), flush=True)
        finally:
            broker.close()
        return
    if args.broker_socket:
        import gc
        import types

        if not os.environ.get("BCI_NO_FREEZE"):
            gc.collect()
            gc.freeze()

        ex = LocalGpuExecutor(broker=types.SimpleNamespace(socket_path=args.broker_socket, close=lambda: None),
                              max_inflight=64, timeout=300, zygotes=args.zygotes)
    else:
        ex = LocalGpuExecutor(max_inflight=64, timeout=300, zygotes=args.zygotes)
    try:
        for name in args.scripts.split(","):
            src = SCRIPTS[name].format(n=args.n)
            for inflight in [int(x) for x in args.inflight.split(",")]:
                rows = []

                async def batch(k):
                    sem = asyncio.Semaphore(inflight)

                    async def one():
                        async with sem:
                            ts = time.time()
                            r = await ex.execute(src)
                            te = time.time()
                            ru = [ln for ln in r.stdout.splitlines() if ln.startswith("RU ")][0].split()
                            rows.append((ts, te, float(ru[1]), float(ru[2]), float(ru[3]), float(ru[4]), int(ru[5]), int(ru[6]), int(ru[7])))
                    await asyncio.gather(*(one() for _ in range(k)))

                asyncio.run(batch(max(16, 2 * inflight)))
                rows.clear()
                s0 = resource.getrusage(resource.RUSAGE_SELF)
                t0 = time.time()
                asyncio.run(batch(args.requests))
                dt = time.time() - t0
                s1 = resource.getrusage(resource.RUSAGE_SELF)
                lat = sorted(r[1] - r[0] for r in rows)
                ms = lambda v: round(1e3 * v, 3)  # noqa: E731
                print(json.dumps({
                    "script": name, "inflight": inflight, "req_per_s": round(args.requests / dt, 1),
                    "latency_ms": {"mean": ms(statistics.mean(lat)), "p50": ms(lat[len(lat) // 2]), "p90": ms(lat[int(len(lat) * 0.9)]),
                                   "p99": ms(lat[int(len(lat) * 0.99)]), "max": ms(lat[-1])},
                    "submit_to_script_start_ms_p50": ms(statistics.median(r[2] - r[0] for r in rows)),
                    "script_body_ms_p50": ms(statistics.median(r[3] - r[2] for r in rows)),
                    "script_end_to_result_ms_p50": ms(statistics.median(r[1] - r[3] for r in rows)),
                    "child_cpu_ms": {"user": ms(statistics.mean(r[4] for r in rows)), "sys": ms(statistics.mean(r[5] for r in rows))},
                    "child_minor_faults": round(statistics.mean(r[6] for r in rows)),
                    "child_ctx_switches": {"voluntary": round(statistics.mean(r[7] for r in rows), 1), "involuntary": round(statistics.mean(r[8] for r in rows), 1)},
                    "service_cpu_ms_per_request": {"user": ms((s1.ru_utime - s0.ru_utime) / args.requests), "sys": ms((s1.ru_stime - s0.ru_stime) / args.requests)},
                    "zygote_phase_ms_mean": (lambda tr: {"submit_to_child_pid": ms(statistics.mean(t[0] for t in tr)), "child_pid_to_exit": ms(statistics.mean(t[1] for t in tr)),
                                                         "p99_submit_to_child_pid": ms(sorted(t[0] for t in tr)[int(len(tr) * 0.99)]),
                                                         # of the SLOW requests (submit -> pid over 15 ms): where did the time go
                                                         "slow": (lambda sl: {"count": len(sl), "submit_to_zygote_accept": ms(statistics.mean(t[3] - t[2] for t in sl)),
                                                                              "zygote_accept_to_handoff": ms(statistics.mean(t[4] - t[3] for t in sl)),
                                                                              "handoff_to_child_running": ms(statistics.mean(t[5] - t[4] for t in sl))} if sl else None)(
                                                             [t for t in tr if len(t) >= 6 and t[0] > 0.015]),
                                                         "fast": (lambda sl: {"count": len(sl), "submit_to_zygote_accept": ms(statistics.mean(t[3] - t[2] for t in sl)),
                                                                              "zygote_accept_to_handoff": ms(statistics.mean(t[4] - t[3] for t in sl)),
                                                                              "handoff_to_child_running": ms(statistics.mean(t[5] - t[4] for t in sl))} if sl else None)(
                                                             [t for t in tr if len(t) >= 6 and t[0] <= 0.015])} if tr else None)(
                        [t for z in ex._zygotes if z.trace for t in z.trace[-args.requests:]]),
                }), flush=True)
    finally:
        ex.close()


if __name__ == "__main__":
    main()
