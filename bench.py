#!/usr/bin/env python
"""bench.py — requests/sec of `/v1/execute` on benchmark-numpy.py-shaped requests.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                      (the reference's CPU path)

A "step" = one batch of 1024 requests per GPU, each request the reference's metric script
(examples/benchmark-numpy.py:18-22): x = rand(1e8) float64; sum(square(x)).  Requests are
sharded round-robin (request i -> GPU i mod N), no data-path collective; per-GPU work is
fixed as N grows ("weak").  `value` times the batch-first C ABI with everything device
resident (waves of 64 requests, the next wave's launch queued before the host waits for the
current one); `e2e` times the same requests through the executor backend's public
`execute(source_code, files, env)` call (source text in, stdout text out) from several service
processes per GPU sharing that GPU's broker, with 16 (8 beyond 2 GPUs) requests admitted per GPU
at a time and up to 4 batches per GPU so that the window is seconds, not a burst.  Rank 0 prints
ONE JSON line on stdout (everything else goes to stderr).  See DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

PER_GPU_BATCH = 1024
DRY = bool(os.environ.get("BCI_BENCH_DRY"))
TDEV = "cpu" if DRY else "cuda"        # where the few scalars that cross ranks live (all_reduce of times)
METRIC = "requests/sec on benchmark-numpy.py at batch=1024 per GPU"


def env_int(name: str, default: int) -> int:
    v = os.environ.get(name)
    return int(v) if v not in (None, "") else default


def measured_bf16() -> float:
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["bf16_tflops"])
    except Exception:
        return 1590.0        # the profiling recipe's fallback (burst)


def measured_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ---- clocks -----------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu = gpu
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="bci-clocks-", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in Path(self.path).read_text().strip().splitlines() if r.strip()]
            sm = [float(r[0]) for r in rows]
            out["sm_mhz"] = statistics.median(sm) if sm else None
            out["sm_max_mhz"] = float(rows[0][1]) if rows else None
            out["power_w_max"] = max(float(r[2]) for r in rows) if rows else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for j, n in enumerate(names) if any(r[3 + j].strip().lower().startswith("active") for r in rows)]
            out["samples"] = len(rows)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


# ---- the GPU arm ---------------------------------------------------------------------------
def c1_step(dev, L, n: int, requests: int, wave: int, seed0: int, timing: dict | None):
    """One step on this rank: `requests` benchmark-numpy bodies, `wave` per launch.
    Per wave ONE kernel: rand(n) is generated in registers inside the fused sum(square(.)) reduction
    (bci_reduce_rand_f64) — no array is allocated, nothing but 8 B/request of results touches HBM."""
    results = []
    for w0 in range(0, requests, wave):
        b = min(wave, requests - w0)
        if timing is not None:
            dev.event_record(0)
        res = dev.reduce_rand_f64(L.RED_SUM, [n] * b, [seed0 + w0 + j for j in range(b)], pre=L.PRE_SQUARE)
        if timing is not None:
            dev.event_record(1)
            timing["fused_ms"].append(dev.event_elapsed(0, 1))
            timing["wave_b"].append(b)
        results.extend(res.tolist())
    return results


def c1_unfused_probe(dev, L, n: int, b: int, reps: int = 3) -> dict:
    """The two HBM-bound kernels the fused launch replaced, timed alone on a wave of `b` requests (outside
    the timed step): rand_f64_kernel writes 8n bytes per request, reduce_kernel<double,SUM,PRE_SQUARE> reads them."""
    import numpy as np

    ptrs = [dev.alloc(n * 8) for _ in range(b)]
    out = {}
    try:
        rand_ms, red_ms = [], []
        for r in range(reps + 1):
            dev.event_record(4)
            dev.rand_f64(ptrs, [n] * b, [1234 + r * b + j for j in range(b)])
            dev.event_record(5)
            dev.reduce(L.RED_SUM, np.float64, ptrs, [n] * b, pre=L.PRE_SQUARE)
            dev.event_record(6)
            if r:   # first round warms up
                rand_ms.append(dev.event_elapsed(4, 5))
                red_ms.append(dev.event_elapsed(5, 6))
        by = 8.0 * n * b
        out = {"rand_f64_kernel": {"avg_launch_ms": sum(rand_ms) / len(rand_ms), "achieved": by / (sum(rand_ms) / len(rand_ms) / 1e3) / 1e9},
               "reduce_kernel<double,SUM,PRE_SQUARE>": {"avg_launch_ms": sum(red_ms) / len(red_ms),
                                                        "achieved": by / (sum(red_ms) / len(red_ms) / 1e3) / 1e9}}
    finally:
        for p in ptrs:
            dev.free(p)
    return out


def run_ours(args) -> dict | None:
    import numpy as np
    from bee_code_interpreter_b200 import _lib as L
    from bee_code_interpreter_b200 import workloads as W
    if DRY:      # control-flow check on the CPU test double (tests/test_bench_contract.py): numbers are meaningless
        sys.path.insert(0, str(ROOT / "tests"))
        import fake_device as dev
    else:
        from bee_code_interpreter_b200 import device as dev

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        if DRY:      # control flow of the multi-rank arm on CPU (tests/test_bench_contract.py): gloo, no CUDA
            dist_mod.init_process_group("gloo")
        else:
            torch.cuda.set_device(local)
            dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    dev.init(local)
    n = args.n
    requests = args.requests
    wave = max(1, min(args.wave, requests))

    def barrier():
        dev.sync()
        if dist is not None:
            import torch

            if not DRY:
                torch.cuda.synchronize()
            dist.barrier()

    seed_base = 0x5EED0000 + rank * 10_000_019
    for s in range(args.warmup):
        c1_step(dev, L, n, requests, wave, seed_base + s * requests, None)
    timing = {"fused_ms": [], "wave_b": []}
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = dev.launch_count()
    dev.timer_begin()
    t0 = time.perf_counter()
    last = None
    for s in range(args.steps):
        last = c1_step(dev, L, n, requests, wave, seed_base + (args.warmup + s) * requests, timing)
    dev_ms = dev.timer_end()
    dev.sync()
    wall_s = time.perf_counter() - t0
    launches = dev.launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else {}

    # statistical parity of the last step (unseeded rand in the reference => distribution only)
    arr = np.asarray(last) / n
    sigma = np.sqrt(4.0 / 45.0 / n)
    parity_ok = bool(np.all(np.abs(arr - 1.0 / 3.0) < 6 * sigma))

    step_ms = dev_ms / args.steps
    if dist is not None:
        import torch

        t = torch.tensor([step_ms], device=TDEV, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
    value = requests * world / (step_ms / 1e3)

    unfused, unfused_b = None, 0
    if rank == 0:
        try:
            free_b, _t = dev.mem_info()
            unfused_b = max(1, min(64, int(free_b * 0.5) // (n * 8)))
            unfused = c1_unfused_probe(dev, L, n, unfused_b)
        except Exception as e:  # a probe must never cost the bench line
            print("bench.py: unfused probe failed: %r" % (e,), file=sys.stderr)
    # strong scaling datum (north star: "batch 1024 on 8 x B200"): the SAME global batch of 1024 split over the ranks
    strong = None
    if world > 1:
        try:
            share = max(1, PER_GPU_BATCH // world)
            c1_step(dev, L, n, share, min(wave, share), seed_base + 7_000_000, None)
            barrier()
            dev.timer_begin()
            for s_ in range(args.steps):
                c1_step(dev, L, n, share, min(wave, share), seed_base + 8_000_000 + s_ * share, None)
            s_ms = dev.timer_end() / args.steps
            t = torch.tensor([s_ms], device=TDEV, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            strong = {"global_batch": share * world, "per_gpu_batch": share, "ms_per_step": float(t.item()),
                      "value": share * world / (float(t.item()) / 1e3), "unit": "requests/s", "scaling": "strong"}
        except Exception as e:      # an extra datum must never cost the bench line
            strong = {"error": repr(e)}

    # hardware evidence of the split-operand collectives: only when >= 2 GPUs are here
    split_parity = None
    if world > 1 and not args.no_split_check and not DRY:
        try:
            from bee_code_interpreter_b200 import split_check

            split_parity = split_check.run(dist, dev, rank, world)
            if split_parity is not None:
                split_parity["ok"] = all(v for k, v in split_parity.items() if k.endswith("_ok"))
        except Exception as e:
            split_parity = {"ok": False, "error": repr(e)}

    # the other BASELINE configs, device resident (bench_configs.py)
    import bench_configs as BC

    which = [c for c in args.configs.split(",") if c in BC.DEVICE_ARMS]
    peaks_all = {"hbm_gbs": measured_peaks()[0], "bf16_tflops": measured_bf16()}
    cfg = BC.run_device_arms(dev, L, rank, world, dist, peaks_all, which, lambda m: print(m, file=sys.stderr)) if which else {}

    dev.shutdown()   # hand the GPU memory back before the service processes start
    e2e = {"value": None, "unit": "requests/s", "error": "skipped (--e2e-requests 0)"}
    if args.e2e_requests > 0:
        harness = None
        try:
            # admission per GPU: 16 is the optimum while the host CPUs keep up; with more GPUs than that the
            # job's CPU quota is shared — one admitted request per available CPU, at least 8 (N = 8 on a 96-CPU quota: 12)
            inflight = args.e2e_inflight if args.e2e_inflight > 0 else (16 if world <= 2 else max(8, min(16, effective_cores()[0] // world)))
            harness = E2EHarness(args, world, local, inflight)
            e2e = harness.run("c1", args.e2e_requests, dist, rank, n, inflight)
            e2e["parity_6sigma"] = e2e["parity"]
            for name in which:
                try:
                    sample = BC.E2E_SAMPLE[name] if not name.startswith("c4") else max(16, BC.E2E_SAMPLE[name] // world)
                    got = harness.run(name, sample, dist, rank, n, BC_INFLIGHT[name] if world <= 2 else max(inflight, BC_INFLIGHT[name] // 2))
                except Exception as e:
                    got = {"value": None, "error": repr(e)}
                if isinstance(cfg.get(name), dict):
                    cfg[name]["e2e"] = got
                if name == "c4" and isinstance(cfg.get(name), dict):
                    # the same config with the 64 MiB output born on the device: generated, hashed in HBM and written
                    # out through the pinned path (BRK_EXPORT) — the object store never reads the file back
                    try:
                        cfg[name]["e2e_device_born_output"] = harness.run("c4dev", max(16, BC.E2E_SAMPLE["c4dev"] // world), dist, rank, n,
                                                                          BC_INFLIGHT["c4dev"] if world <= 2 else BC_INFLIGHT["c4dev"] // 2)
                    except Exception as e:
                        cfg[name]["e2e_device_born_output"] = {"value": None, "error": repr(e)}
        except Exception as e:  # executor layer not usable: no end-to-end number
            e2e = {"value": None, "unit": "requests/s", "error": repr(e)}
        finally:
            if harness is not None:
                harness.close()

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None

    peak, peak_kind = measured_peaks()
    tot_b = sum(timing["wave_b"])
    fused_s = sum(timing["fused_ms"]) / 1e3
    sm_mhz = (clocks.get("sm_mhz") or 1965.0)
    # instruction roofline of the fused kernel.  ncu (profiles/r2_ncu_new_kernels.csv, `--set full` of a 16-request
    # launch): 0.986 warp-instructions = 31.5 thread-instructions per float64 (Philox4x32-10: 10 IMAD.WIDE + 10 LOP3,
    # numpy's 53-bit double: 2 SHF + 2 I2F + DMUL + DFMA, square + add); an SM issues at most 4 warp-instructions per
    # clock, so the ceiling is 4 x 32 / 31.5 = 4.06 float64 per clock per SM.
    instr_per_elem = 31.5
    elem_rate = n * tot_b / fused_s if fused_s else None
    elem_peak = 148 * 128.0 / instr_per_elem * sm_mhz * 1e6
    algo = float(W.C1_ALGO_BYTES) * n / W.C1_N      # numpy's unfused traffic per request (32 n bytes)
    k_fused = {"kernel": "reduce_rand_kernel<SUM,PRE_SQUARE>", "requests_per_launch": wave,
               "algorithmic_bytes_per_launch": algo * wave, "moved_bytes_per_launch": 8 * wave,
               "avg_launch_ms": 1e3 * fused_s / max(1, len(timing["fused_ms"])),
               "achieved": algo * tot_b / fused_s / 1e9 if fused_s else None}
    roofline = {"bound": "hbm", "kernel": k_fused["kernel"], "achieved": k_fused["achieved"], "peak": peak,
                "unit": "GB/s", "frac": (k_fused["achieved"] / peak) if k_fused["achieved"] else None,
                # dram__bytes_read + write of the ncu capture: 29.7 KB for a 16-request launch (descriptors, partials)
                "traffic": 29696.0 / 16 * wave,
                "traffic_note": "the kernel reads no array and writes 8 B per request + one partial per CTA: rand(n), its square "
                                "and the sum live in registers.  `achieved` counts numpy's UNFUSED algorithmic bytes (32 n per "
                                "request, SURVEY.md a9), which is why frac > 1; bytes actually moved are ~2 KB per request",
                "bound_note": "not HBM: instruction issue.  ncu: issue slots 59 % active with 6.5 eligible warps per scheduler per "
                              "cycle and 99 % occupancy — IMAD.WIDE and FP64 instructions hold the dispatch port for two cycles; fma "
                              "pipe 38 %, ALU 45 %, XU (I2F) 30 %, FP64 16 % busy — see `issue`",
                "issue": {"unit": "float64 elements/s", "achieved": elem_rate, "peak": elem_peak,
                          "frac": (elem_rate / elem_peak) if elem_rate else None, "thread_instructions_per_element": instr_per_elem,
                          "peak_how": "148 SMs x 4 schedulers x 32 lanes per clock / 31.5 instructions per element x median SM clock under load"},
                "peak_kind": peak_kind, "kernels": [k_fused],
                "share_of_step": {"fused": fused_s / (dev_ms / 1e3)}}
    if unfused:
        for name, k in unfused.items():
            k.update({"kernel": name, "frac": k["achieved"] / peak, "algorithmic_bytes_per_launch": 8 * n * unfused_b,
                      "note": "timed alone outside the step (the pair the fused launch replaced)"})
            roofline["kernels"].append(k)
    out = {
        "metric": METRIC, "value": value, "unit": "requests/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "benchmark-numpy.py x %d requests/GPU (rand(%d) f64 -> square -> sum)" % (requests, n),
                   "global_batch": requests * world, "per_gpu_batch": requests, "array_elems": n,
                   "wave": wave, "parallelism": "requests round-robin, dp%d, no collective" % world,
                   "cache": "no HBM inputs: the operand is generated in registers (compute-bound step); "
                            "the HBM-bound kernels it replaced are timed on 51 GB waves (>> L2) in roofline.kernels",
                   "fusion": "rand + square + sum in ONE kernel per wave (bci_reduce_rand_f64)"},
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
        "wall_ms_per_step": 1e3 * wall_s / args.steps, "parity": {"statistical_6sigma": parity_ok},
    }
    if strong is not None:
        out["strong_scaling"] = strong
    if split_parity is not None:
        out["split_operand_parity"] = split_parity
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(n, budget_s=args.cpu_budget)
    for name in which:
        c = cfg.get(name)
        if not isinstance(c, dict):
            continue
        c["workload"] = BC.DESCRIPTIONS[name]
        c["n_gpus"] = world
        if world == 1 and not args.no_cpu_baseline and "error" not in c:
            try:
                c["cpu_baseline"] = BC.cpu_baseline(name, budget_s=args.cfg_cpu_budget)
            except Exception as e:
                c["cpu_baseline"] = {"error": repr(e)}
    if cfg:
        out["configs"] = cfg
    return out


BC_INFLIGHT = {"c2": 32, "c3": 16, "c4": 48, "c4dev": 48, "c5": 32}      # admitted requests per GPU in the e2e arm of each config
for _kv in filter(None, os.environ.get("BCI_BENCH_CFG_INFLIGHT", "").split(",")):      # experiments: "c5=32,c2=24"
    BC_INFLIGHT[_kv.split("=")[0]] = int(_kv.split("=")[1])
MIN_JOBS = {"c1": 1, "c2": 6, "c3": 5, "c4": 1, "c4dev": 2, "c5": 5}     # broker jobs one request of each workload must cause


class E2EHarness:
    """The public-API arm: `--e2e-workers` service processes per GPU (the way several uvicorn workers would front
    one GPU), each with its own LocalGpuExecutor (zygotes, object store), all sharing this rank's broker process.
    Workers are started once and then run one workload after the other on command."""

    def __init__(self, args, world: int, local: int, inflight_per_gpu: int = 16):
        from bee_code_interpreter_b200.broker import BrokerProcess

        self.args, self.world = args, world
        # the host's fork/exit capacity (~4k processes/s on this 128-core box, tools/spawn_scale.py) is
        # shared by all ranks: fewer service processes per GPU when several GPUs are driven at once
        # (measured: 4 service processes in total give ~3.0-3.2 k req/s whether they feed 1 GPU or 2,
        # 8 give ~9 k, 16 give ~8.3 k — so 8 in total up to 4 GPUs, 2 per GPU beyond)
        self.workers = max(1, args.e2e_workers if args.e2e_workers > 0 else (4 if world <= 2 else 2))
        # forking processes box-wide: the kernel's fork path peaks at 16-32 of them and degrades beyond
        # (profiles/r1_fork_scale.md), so the warm interpreters per service process shrink with world size
        # one warm interpreter per admitted request of a worker (each is pinned to its own CPU, executor.py): a request
        # then never waits behind another one's fork, and children start on a CPU nobody else is starting on
        zygotes = args.e2e_zygotes if args.e2e_zygotes > 0 else max(1, min(8, inflight_per_gpu // self.workers))
        self.broker = (BrokerProcess(fake_device="fake_device", extra_pythonpath=str(ROOT / "tests")) if DRY
                       else BrokerProcess(device=local))
        self.procs = []
        try:
            for _ in range(self.workers):
                cmd = [sys.executable, str(ROOT / "bench.py"), "--e2e-worker", "--broker-socket", self.broker.socket_path,
                       "--e2e-zygotes", str(zygotes)]
                self.procs.append(subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True))
            for p in self.procs:
                line = p.stdout.readline()
                if line.strip() != "ready":
                    raise RuntimeError("e2e worker failed to start: %r" % line)
        except Exception:
            self.close()
            raise

    def _stats(self) -> dict:
        from bee_code_interpreter_b200.broker_client import RemoteBackend

        be = RemoteBackend(self.broker.socket_path)
        try:
            return be.ping()["stats"]
        finally:
            be.close()

    def _command(self, cmds):
        for p, c in zip(self.procs, cmds):
            p.stdin.write(json.dumps(c) + "\n")
            p.stdin.flush()
        outs = []
        for p in self.procs:
            line = p.stdout.readline()
            if not line:
                raise RuntimeError("e2e worker died")
            outs.append(json.loads(line))
        return outs

    def run(self, workload: str, requests_per_gpu: int, dist, rank: int, n: int, inflight_per_gpu: int, check: int = 2) -> dict:
        """`requests_per_gpu` requests of `workload` through execute() on this rank's GPU; request ids follow the
        round-robin placement (id = rank + world * k).  Timed from the first worker's start to the last one's end,
        max over ranks."""
        w, world = self.workers, self.world
        counts = [requests_per_gpu // w + (1 if i < requests_per_gpu % w else 0) for i in range(w)]
        inflight = max(1, inflight_per_gpu // w)
        base = {"workload": workload, "n": n, "inflight": inflight, "stride": world * w}
        warm_n = 4 if workload.startswith("c4") else max(8, 2 * inflight)      # C4 requests are 64 MiB each: a short warm-up
        warm = [dict(base, count=min(c, warm_n), first=rank + world * i + world * w * 100_000, check=0) for i, c in enumerate(counts)]
        self._command(warm)                                       # untimed: page cache, pool growth, zygote warm-up
        before = self._stats()
        if dist is not None:
            dist.barrier()
        outs = self._command([dict(base, count=c, first=rank + world * i, check=check) for i, c in enumerate(counts)])
        after = self._stats()
        dt = max(o["t_end"] for o in outs) - min(o["t_start"] for o in outs)
        if dist is not None:
            import torch

            t = torch.tensor([dt], device=TDEV, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        lat = sorted(x for o in outs for x in o["latencies"])
        jobs = after["jobs"] - before["jobs"]
        done = sum(o["requests"] for o in outs)
        return {"value": requests_per_gpu * world / dt, "unit": "requests/s", "requests_per_gpu": requests_per_gpu,
                "service_workers_per_gpu": w, "inflight_per_worker": inflight,
                "h2d_bytes_per_step": sum(o["h2d_bytes"] for o in outs), "d2h_bytes_per_step": sum(o["d2h_bytes"] for o in outs),
                "p50_latency_s": lat[len(lat) // 2] if lat else None,
                "parity": all(o["parity"] for o in outs), "parity_checked": sum(o["checked"] for o in outs),
                # every request must have reached the GPU: a script that silently fell back to host numpy
                # would make this number meaningless
                "broker_jobs": jobs, "broker_launch_groups": after["batches"] - before["batches"],
                "all_requests_on_gpu": jobs >= MIN_JOBS[workload] * done,
                "host_cores": effective_cores()[0], "host_cores_note": effective_cores()[1],
                "api": "LocalGpuExecutor.execute(source_code, files, env)"}

    def close(self):
        for p in self.procs:
            try:
                p.stdin.write(json.dumps({"op": "quit"}) + "\n")
                p.stdin.flush()
            except Exception:
                pass
        for p in self.procs:
            try:
                p.wait(timeout=10)
            except Exception:
                p.kill()
        self.broker.close()


def e2e_worker(args):
    """One service process: its own executor (zygotes, object store) on the shared broker; runs workloads on command."""
    import types

    import numpy as np
    from bee_code_interpreter_b200.executor import LocalGpuExecutor
    from bee_code_interpreter_b200 import workloads as W
    import bench_configs as BC

    import gc

    shared = types.SimpleNamespace(socket_path=args.broker_socket, close=lambda: None)
    gc.collect()
    gc.freeze()      # a service worker's module objects are permanent too: full collections cost tens of ms on the event loop
    # throughput arm: children are forked on arrival (zygote_pool=0).  The pre-forked pool (the default, 4) halves the
    # time to script start but costs ~8 % of the requests/s when the host CPU quota is the limit, which it is here
    # (profiles/r2_e2e_host_cost.md)
    execu = LocalGpuExecutor(broker=shared, max_inflight=64, timeout=600, zygotes=args.e2e_zygotes or None,
                             min_rand_elems=1000 if DRY else None, zygote_pool=0)
    print("ready", flush=True)
    while True:
        line = sys.stdin.readline()
        if not line:
            break
        cmd = json.loads(line)
        if cmd.get("op") == "quit":
            break
        name, n, count = cmd["workload"], cmd["n"], cmd["count"]
        ids = [cmd["first"] + j * cmd["stride"] for j in range(count)]
        srcs = [W.benchmark_numpy_source(n)] * count if name == "c1" else [BC.source_for(name, i) for i in ids]
        lat: list[float] = []

        async def batch():
            sem = asyncio.Semaphore(cmd["inflight"])

            async def one(src):
                async with sem:
                    t = time.perf_counter()
                    r = await execu.execute(src)
                    lat.append(time.perf_counter() - t)
                    return r
            return await asyncio.gather(*(one(s) for s in srcs))

        t0 = time.time()
        res = asyncio.run(batch())
        t1 = time.time()
        ok = all(r.exit_code == 0 for r in res)
        checked = 0
        if name == "c1":
            good = ok and all(r.stdout.startswith("Result:") for r in res)
            vals = np.array([float(r.stdout.split()[1]) for r in res]) / n if good else np.array([])
            parity = bool(good and np.all(np.abs(vals - 1.0 / 3.0) < 6 * np.sqrt(4.0 / 45.0 / n)))
            checked = len(res)
        else:
            parity = ok
            for i, r in list(zip(ids, res))[: cmd.get("check", 0)]:
                parity = parity and BC.check_result(name, i, r)
                checked += 1
        h2d = sum(len(s.encode()) for s in srcs) + (BC.C4_BYTES * count if name == "c4" else 0)       # c4: the file goes to the GPU to be hashed
        d2h = (sum(len(r.stdout.encode()) for r in res) if name != "c1" else 8 * count) + (BC.C4_BYTES * count if name == "c4dev" else 0)
        if name.startswith("c4"):      # 64 MiB objects: do not let the worker's temporary object store pile up on disk
            try:
                store = str(execu.file_storage.storage_path)
                for entry in os.scandir(store):
                    os.unlink(entry.path)
            except OSError:
                pass
        print(json.dumps({"t_start": t0, "t_end": t1, "requests": count, "parity": parity, "checked": checked, "latencies": lat,
                          "h2d_bytes": h2d, "d2h_bytes": d2h,
                          "first_error": next((r.stderr[-300:] for r in res if r.exit_code != 0), None)}), flush=True)
    execu.close()


# ---- the reference's CPU path ----------------------------------------------------------------
def cpu_round(n: int, procs: int) -> tuple[float, list[float]]:
    """`procs` concurrent requests on the reference's code path (oracle LocalCpuExecutor:
    python subprocess per request, stock numpy, single-threaded BLAS)."""
    from oracle.cpu_executor import LocalCpuExecutor, RefStorage
    from bee_code_interpreter_b200 import workloads as W

    src = W.benchmark_numpy_source(n)
    env = {"OMP_NUM_THREADS": "1", "OPENBLAS_NUM_THREADS": "1", "MKL_NUM_THREADS": "1"}
    with tempfile.TemporaryDirectory() as td:
        ex = LocalCpuExecutor(RefStorage(td), timeout=600)
        lat: list[float] = []

        async def one():
            t = time.perf_counter()
            r = await ex.execute(src, env=env)
            lat.append(time.perf_counter() - t)
            assert r.exit_code == 0, r.stderr
            return r

        async def batch():
            return await asyncio.gather(*(one() for _ in range(procs)))

        t0 = time.perf_counter()
        asyncio.run(batch())
        return time.perf_counter() - t0, lat


def effective_cores() -> tuple[int, str]:
    """Host cores this process tree may actually use: the cgroup CPU quota when there is one (the boxes of this pool
    show 128 logical CPUs but run the job under `cpu.max = 1600000 100000`, i.e. 16 CPUs — profiles/r2i_cgroup.log),
    else the affinity mask."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    note = "%d logical CPUs visible" % n
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            q = max(1, int(round(int(quota) / int(period))))
            if q < n:
                note += ", cgroup cpu.max quota = %d CPUs" % q
                n = q
    except (OSError, ValueError):
        pass
    return n, note


def cpu_procs(n: int) -> int:
    cores = effective_cores()[0]
    try:
        import psutil

        avail = psutil.virtual_memory().available
        cores = max(1, min(cores, int(avail * 0.6) // (n * 8 * 2 + (200 << 20))))
    except Exception:
        cores = min(cores, 16)
    return cores


def cpu_baseline(n: int, budget_s: float = 25.0) -> dict:
    procs = cpu_procs(n)
    t0 = time.perf_counter()
    done, lats, rounds = 0, [], 0
    wall = 0.0
    while True:
        w, lat = cpu_round(n, procs)
        wall += w
        done += procs
        lats += lat
        rounds += 1
        if time.perf_counter() - t0 + w > budget_s:
            break
    return {"value": done / wall, "unit": "requests/s", "cores": procs, "cores_note": effective_cores()[1], "kind": "port",
            "p50_latency_s": statistics.median(lats),
            "sample": "%d rounds x %d concurrent literal benchmark-numpy.py requests (n=%d), python subprocess + stock numpy per request" % (rounds, procs, n)}


def run_reference(args) -> dict | None:
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return None
    n = args.n
    procs = cpu_procs(n)
    for _ in range(min(args.warmup, 1)):
        cpu_round(n, procs)
    walls, lats = [], []
    budget_t0 = time.perf_counter()
    steps_done = 0
    for _ in range(args.steps):
        w, lat = cpu_round(n, procs)
        walls.append(w)
        lats += lat
        steps_done += 1
        if time.perf_counter() - budget_t0 > args.ref_budget:
            break
    value = procs * steps_done / sum(walls)
    sample = "each step = %d concurrent literal benchmark-numpy.py requests (n=%d) on host cores" % (procs, n)
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "requests/s", "n_gpus": world,
        "steps": steps_done, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(walls) / steps_done,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "benchmark-numpy.py (rand(%d) f64 -> square -> sum), reference CPU executor path" % n,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "requests/s", "cores": procs, "cores_note": effective_cores()[1], "kind": "port", "sample": sample,
                         "p50_latency_s": statistics.median(lats)},
        "e2e": {"value": value, "unit": "requests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--array-elems", dest="n", type=int, default=10**8,
                    help="array elements per request (benchmark-numpy.py: 1e8); use the long form under torchrun, "
                         "whose own parser treats --n as an abbreviation")
    ap.add_argument("--requests", type=int, default=PER_GPU_BATCH, help="requests per GPU per step")
    ap.add_argument("--wave", type=int, default=256, help="requests per launch")
    ap.add_argument("--e2e-requests", type=int, default=-1,
                    help="requests per GPU in the e2e arm (-1 = min(steps, 4) batches of 1024: a single batch lasts ~0.3 s, "
                         "which is mostly ramp-up and tail)")
    ap.add_argument("--e2e-inflight", type=int, default=0,
                    help="admitted requests per service process (0 = auto: 16 in flight per GPU)")
    ap.add_argument("--e2e-workers", type=int, default=0, help="service processes per GPU in the e2e arm (0 = auto)")
    ap.add_argument("--e2e-zygotes", type=int, default=0, help="warm interpreters per service process (0 = auto)")
    ap.add_argument("--e2e-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--broker-socket", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="c2,c3,c4,c5", help="other BASELINE configs to measure beside the headline ('' = none)")
    ap.add_argument("--cfg-cpu-budget", type=float, default=10.0, help="seconds of CPU baseline per config")
    ap.add_argument("--no-split-check", action="store_true", help="skip the split-operand NCCL self-check at N >= 2")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--ref-budget", type=float, default=150.0)
    args = ap.parse_args()
    world = env_int("WORLD_SIZE", 1)
    if args.e2e_requests < 0:
        args.e2e_requests = PER_GPU_BATCH * max(1, min(args.steps, 4))
    if args.gpus != world and world == 1 and args.gpus > 1:
        print("bench.py: --gpus %d needs torchrun (one process per GPU); running 1 GPU" % args.gpus, file=sys.stderr)
    if args.e2e_worker:
        e2e_worker(args)
        return
    # stdout carries ONE line, the JSON: whatever libraries print meanwhile (NCCL's version banner,
    # torchrun notices) is sent to stderr at the descriptor level
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        out = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
