#!/usr/bin/env python
"""Per-kernel roofline table on one B200: every op class of SURVEY.md §8a at the sizes of the
BASELINE configs, timed with CUDA events on the library's stream (warm-ups first; inputs
larger than L2 or L2 flushed between iterations).  All ctypes argument arrays are built
before the timed region, so the interval holds the launch(es) only.
    python tools/kernel_bench.py [--quick] [--md]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os

os.environ.setdefault("BCI_TUNING", "1")      # the kernels' A/B switches are honoured only in tuning mode
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from bee_code_interpreter_b200 import _lib as L  # noqa: E402
from bee_code_interpreter_b200 import device as dev  # noqa: E402

lib = None


def peak_gbs() -> float:
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def timed(fn, iters=5, warm=3, flush=True):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(iters):
        if flush:
            dev.flush_l2()
        dev.event_record(10)
        fn()
        dev.event_record(11)
        ms.append(dev.event_elapsed(10, 11))
    return float(np.median(ms))


# ---- closures over prebuilt argument arrays --------------------------------------------------
def k_ewise(op, dtype, in0, in1, scalars, out, ns):
    code, B = dev.dtype_code(dtype), len(in0)
    a0, a1, o, n = L.ptr_array(in0), (L.ptr_array(in1) if in1 is not None else None), L.ptr_array(out), L.size_array(ns)
    sc = dev._scalars(scalars, code, B)
    return lambda: L.check(lib.bci_ewise(op, code, B, a0, a1, sc, o, n))


def k_reduce(op, dtype, in0, ns, pre=L.PRE_NONE):
    code, B = dev.dtype_code(dtype), len(in0)
    a0, n = L.ptr_array(in0), L.size_array(ns)
    raw = np.zeros(B, dtype=np.uint64)
    return lambda: L.check(lib.bci_reduce(op, pre, code, B, a0, n, C.c_void_p(raw.ctypes.data)))


def k_reduce_fused(op, ew, dtype, in0, in1, ns, pre=L.PRE_NONE, addends=None):
    code, B = dev.dtype_code(dtype), len(in0)
    a0, a1, n = L.ptr_array(in0), L.ptr_array(in1), L.size_array(ns)
    ad = dev._scalars(addends, code, B)
    raw = np.zeros(B, dtype=np.uint64)
    return lambda: L.check(lib.bci_reduce_fused(op, pre, ew, code, B, a0, a1, None, ad, n, C.c_void_p(raw.ctypes.data)))


def k_argext(op, dtype, in0, ns):
    code, B = dev.dtype_code(dtype), len(in0)
    a0, n = L.ptr_array(in0), L.size_array(ns)
    out = np.zeros(B, dtype=np.int64)
    return lambda: L.check(lib.bci_argext(op, code, B, a0, n, out.ctypes.data_as(C.POINTER(C.c_int64))))


def k_inout(fn_name, dtype, in0, out, ns):
    code, B = dev.dtype_code(dtype), len(in0)
    a0, o, n = L.ptr_array(in0), L.ptr_array(out), L.size_array(ns)
    fn = getattr(lib, fn_name)
    return lambda: L.check(fn(code, B, a0, o, n))


def k_rand(out, ns):
    B = len(out)
    o, n = L.ptr_array(out), L.size_array(ns)
    seeds = (C.c_uint64 * B)(*range(B))
    return lambda: L.check(lib.bci_rand_f64(B, o, n, seeds))


def k_sha(ptrs, lens):
    B = len(ptrs)
    p, n = L.ptr_array(ptrs), L.size_array(lens)
    out = np.zeros(B * 32, dtype=np.uint8)
    return lambda: L.check(lib.bci_sha256(B, p, n, 1, C.c_void_p(out.ctypes.data)))


def k_matmul(a, b, c, m):
    B = len(a)
    pa, pb, pc = L.ptr_array(a), L.ptr_array(b), L.ptr_array(c)
    return lambda: L.check(lib.bci_matmul_f32(B, pa, pb, pc, m, m, m))


def main():
    global lib
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--md", action="store_true")
    args = ap.parse_args()
    dev.init(0)
    lib = L.load()
    peak = peak_gbs()
    rows = []

    def add(name, cfg, algo_bytes, ms, moved_bytes=None, extra=None):
        ach = algo_bytes / (ms / 1e3) / 1e9
        row = {"kernel": name, "config": cfg, "ms": ms, "algorithmic_GB": algo_bytes / 1e9, "achieved": ach,
               "unit": "GB/s", "frac_of_measured_hbm": ach / peak}
        if moved_bytes is not None:
            row["moved_GBps"] = moved_bytes / (ms / 1e3) / 1e9
            row["moved_frac"] = row["moved_GBps"] / peak
        if extra:
            row.update(extra)
        rows.append(row)

    # ---- C1: 1e8 f64, B requests per launch (12.8 GB per operand: no L2 flush needed)
    B1 = 4 if args.quick else 16
    n1 = 10**8
    bufs = [dev.alloc(n1 * 8) for _ in range(B1)]
    outs = [dev.alloc(n1 * 8) for _ in range(B1)]
    ns = [n1] * B1
    add("rand_f64 (K8)", f"C1 n=1e8 f64 x{B1}", 8 * n1 * B1, timed(k_rand(bufs, ns), flush=False))
    # seeded legacy stream: one CTA per request, so B >= SM count shows the device rate and B = 1 the per-stream rate
    rs0 = np.random.RandomState(1).get_state()
    st0 = np.concatenate([rs0[1], [rs0[2]]]).astype(np.uint32)
    for Bm, nm in ((1, 10**7), (len(bufs), 10**7), (296, 2 * 10**6)):
        tgt = [bufs[i % len(bufs)] + (i // len(bufs)) * nm * 8 for i in range(Bm)]
        sts = np.tile(st0, (Bm, 1))
        o_, n_ = L.ptr_array(tgt), L.size_array([nm] * Bm)
        stp = sts.ctypes.data_as(C.POINTER(C.c_uint32))
        ms = timed(lambda: L.check(lib.bci_rand_mt19937_f64(Bm, o_, n_, stp)), iters=3, warm=1, flush=False)
        add("rand_mt19937 seeded (K8b)", f"n={nm:.0e} x{Bm}", 8 * nm * Bm, ms,
            extra={"bound": "issue rate of one SM per stream (sequential generator)", "GBps_per_stream": 8 * nm / (ms / 1e3) / 1e9})
    add("ewise square f64 (K1)", f"C1 n=1e8 x{B1}", 16 * n1 * B1, timed(k_ewise(L.EW_SQUARE, np.float64, bufs, None, None, outs, ns), flush=False))
    add("reduce sum f64 (K2)", f"C1 n=1e8 x{B1}", 8 * n1 * B1, timed(k_reduce(L.RED_SUM, np.float64, bufs, ns), flush=False))
    add("reduce sum(square) f64 fused (K1+K2)", f"C1 n=1e8 x{B1}", 24 * n1 * B1,
        timed(k_reduce(L.RED_SUM, np.float64, bufs, ns, pre=L.PRE_SQUARE), flush=False), moved_bytes=8 * n1 * B1)
    add("ewise add f64 (K1)", f"n=1e8 x{B1}", 24 * n1 * B1, timed(k_ewise(L.EW_ADD, np.float64, bufs, outs, None, outs, ns), flush=False))
    # ---- round 2: generator fused into the reduction, PCG64 Generator streams, creation, casts
    seeds = (C.c_uint64 * B1)(*range(B1))
    raw = np.zeros(B1, dtype=np.uint64)
    nn = L.size_array(ns)
    ms = timed(lambda: L.check(lib.bci_reduce_rand_f64(L.RED_SUM, L.PRE_SQUARE, B1, nn, seeds, C.c_void_p(raw.ctypes.data))), flush=False)
    add("reduce sum(square(rand)) f64 fused (K1+K2+K8)", f"C1 n=1e8 x{B1}", 32 * n1 * B1, ms, moved_bytes=0,
        extra={"bound": "fma pipe (Philox4x32-10: 10 IMAD.WIDE per float64), no HBM traffic", "requests_per_s": B1 / (ms / 1e3),
               "imad_wide_lane_ops_per_s": 10.0 * n1 * B1 / (ms / 1e3)})
    st = np.random.PCG64(1).state["state"]
    gens = (L.Pcg64Gen * B1)()
    m64 = (1 << 64) - 1
    for i in range(B1):
        gens[i].state_lo, gens[i].state_hi = st["state"] & m64, st["state"] >> 64
        gens[i].inc_lo, gens[i].inc_hi = st["inc"] & m64, st["inc"] >> 64
        gens[i].low, gens[i].range = 0, 2**31
    o_ = L.ptr_array(bufs)
    rej = (C.c_uint32 * B1)()
    n32 = L.size_array([2 * n1] * B1)
    add("pcg64 random f64 (K8d)", f"n=1e8 x{B1}", 8 * n1 * B1, timed(lambda: L.check(lib.bci_rand_pcg64(L.PCG_F64, B1, o_, nn, gens, None)), flush=False))
    add("pcg64 random f32 (K8d)", f"n=2e8 x{B1}", 8 * n1 * B1, timed(lambda: L.check(lib.bci_rand_pcg64(L.PCG_F32, B1, o_, n32, gens, None)), flush=False))
    add("pcg64 integers(0, 2^31) int64 (K8d)", f"n=1e8 x{B1}", 8 * n1 * B1,
        timed(lambda: L.check(lib.bci_rand_pcg64(L.PCG_BOUNDED32_I64, B1, o_, nn, gens, rej)), flush=False))
    sa = dev._scalars([0.5] * B1, L.F64, B1)
    sb = dev._scalars([0.25] * B1, L.F64, B1)
    add("fill arange f64", f"n=1e8 x{B1}", 8 * n1 * B1, timed(lambda: L.check(lib.bci_fill(L.FILL_ARANGE, L.F64, B1, o_, nn, sa, sb, None)), flush=False))
    add("fill const f64", f"n=1e8 x{B1}", 8 * n1 * B1, timed(lambda: L.check(lib.bci_fill(L.FILL_CONST, L.F64, B1, o_, nn, sa, None, None)), flush=False))
    oo = L.ptr_array(outs)
    add("cast f64 -> f32", f"n=1e8 x{B1}", 12 * n1 * B1, timed(lambda: L.check(lib.bci_cast(L.F64, L.F32, B1, o_, oo, nn)), flush=False))
    add("cast i64 -> f64", f"n=1e8 x{B1}", 16 * n1 * B1, timed(lambda: L.check(lib.bci_cast(L.I64, L.F64, B1, o_, oo, nn)), flush=False))
    for p in bufs + outs:
        dev.free(p)

    # ---- C3: 2^20 f32, 1024 requests per launch (4.3 GB per operand)
    B3 = 256 if args.quick else 1024
    n3 = 2**20
    x = [dev.alloc(n3 * 4) for _ in range(B3)]
    y = [dev.alloc(n3 * 4) for _ in range(B3)]
    o = [dev.alloc(n3 * 4) for _ in range(B3)]
    ns = [n3] * B3
    host = np.random.default_rng(2000).random(n3, dtype=np.float32)
    for p in x + y:
        dev.h2d(p, host)
    dev.sync()
    u = n3 * 4 * B3
    add("ewise mul f32 (K1)", f"C3 n=2^20 f32 x{B3}", 3 * u, timed(k_ewise(L.EW_MUL, np.float32, x, y, None, o, ns), flush=False))
    add("ewise add-scalar f32 (K1)", f"C3 x{B3}", 2 * u, timed(k_ewise(L.EW_ADD, np.float32, x, None, [1.0] * B3, o, ns), flush=False))
    add("ewise abs f32 (K1)", f"C3 x{B3}", 2 * u, timed(k_ewise(L.EW_ABS, np.float32, x, None, None, o, ns), flush=False))
    add("reduce sum f32 (K2)", f"C3 x{B3}", u, timed(k_reduce(L.RED_SUM, np.float32, x, ns), flush=False))
    add("reduce max f32 (K2)", f"C3 x{B3}", u, timed(k_reduce(L.RED_MAX, np.float32, x, ns), flush=False))
    add("reduce sum(square) f32 fused", f"C3 x{B3}", 3 * u, timed(k_reduce(L.RED_SUM, np.float32, x, ns, pre=L.PRE_SQUARE), flush=False), moved_bytes=u)
    add("reduce mean(x*y+1) f32 fused", f"C3 x{B3}", 6 * u,
        timed(k_reduce_fused(L.RED_MEAN, L.EW_MUL, np.float32, x, y, ns, addends=[1.0] * B3), flush=False), moved_bytes=2 * u)
    add("reduce max(abs(x-y)) f32 fused", f"C3 x{B3}", 6 * u,
        timed(k_reduce_fused(L.RED_MAX, L.EW_SUB, np.float32, x, y, ns, pre=L.PRE_ABS), flush=False), moved_bytes=2 * u)
    for p in x + y + o:
        dev.free(p)

    # ---- C5: 1e7 rows, 32 requests per launch
    B5 = 8 if args.quick else 32
    n5 = 10**7
    v = [dev.alloc(n5 * 4) for _ in range(B5)]
    vo = [dev.alloc(n5 * 4) for _ in range(B5)]
    k = [dev.alloc(n5 * 8) for _ in range(B5)]
    ko = [dev.alloc(n5 * 8) for _ in range(B5)]
    idx = [dev.alloc(n5 * 8) for _ in range(B5)]
    ns = [n5] * B5
    rng = np.random.default_rng(4000)
    hv, hk = rng.random(n5, dtype=np.float32), rng.integers(0, 2**31, n5, dtype=np.int64)
    for i in range(B5):
        dev.h2d(v[i], np.roll(hv, i * 1001))
        dev.h2d(k[i], np.roll(hk, i * 1001))
    dev.sync()
    add("argmax f32 (K6)", f"C5 n=1e7 x{B5}", 4 * n5 * B5, timed(k_argext(L.ARGMAX, np.float32, v, ns)))
    add("cumsum i64 (K7)", f"C5 n=1e7 x{B5}", 16 * n5 * B5, timed(k_inout("bci_cumsum", np.int64, k, ko, ns)), moved_bytes=16 * n5 * B5)   # default = single-pass kernel: moves what it must
    for mode in ("legacy", "lookback", "pipe"):
        os.environ["BCI_SCAN"] = mode
        add(f"cumsum i64 [{mode}] (K7)", f"C5 n=1e7 x{B5}", 16 * n5 * B5, timed(k_inout("bci_cumsum", np.int64, k, ko, ns)))
    os.environ.pop("BCI_SCAN")
    add("sort f32 (K5)", f"C5 n=1e7 x{B5}", 36 * n5 * B5, timed(k_inout("bci_sort", np.float32, v, vo, ns), iters=3), moved_bytes=48 * n5 * B5)
    # C5's int64 column holds values below 2^31: the three constant middle sweeps are left out (5 of 8 run)
    add("sort i64 in [0, 2^31) (K5)", f"n=1e7 x{B5}", 17 * 8 * n5 * B5, timed(k_inout("bci_sort", np.int64, k, ko, ns), iters=3), moved_bytes=15 * 8 * n5 * B5)
    hfull = rng.integers(-2**63, 2**63 - 1, n5, dtype=np.int64)
    for i in range(B5):
        dev.h2d(ko[i], np.roll(hfull, i * 1001))
    add("sort i64 full range (K5)", f"n=1e7 x{B5}", 17 * 8 * n5 * B5, timed(k_inout("bci_sort", np.int64, ko, idx, ns), iters=3), moved_bytes=24 * 8 * n5 * B5)
    add("argsort f32 stable (K5)", f"n=1e7 x{B5}", (36 + 8) * n5 * B5, timed(k_inout("bci_argsort", np.float32, v, idx, ns), iters=3))
    for p in v + vo + k + ko + idx:
        dev.free(p)

    # ---- C4: SHA-256 on device-resident files
    B4 = 64 if args.quick else 512
    sz = 64 * 2**20
    slab = dev.alloc(sz)           # one 64 MiB payload hashed by B4 warps (content is irrelevant to timing)
    dev.rand_f64([slab], [sz // 8], [7])
    ms = timed(k_sha([slab] * B4, [sz] * B4), iters=2, warm=1, flush=False)
    add("sha256 warp-per-message (K4)", f"C4 {B4} x 64 MiB", sz * B4, ms,
        extra={"bound": "INT ALU / dependency chain, not HBM", "MBps_per_message": sz / (ms / 1e3) / 1e6})
    small, Bs = 4096, 65536
    ms = timed(k_sha([slab + (i % 8192) * small for i in range(Bs)], [small] * Bs), iters=3, warm=1, flush=False)
    add("sha256 thread-per-message (K4)", f"{Bs} x 4 KiB", small * Bs, ms, extra={"bound": "INT ALU"})
    dev.free(slab)

    # ---- C2: matmul 4096^3 f32
    Bm = 2 if args.quick else 8
    m = 4096
    a = [dev.alloc(m * m * 4) for _ in range(Bm)]
    b = [dev.alloc(m * m * 4) for _ in range(Bm)]
    c = [dev.alloc(m * m * 4) for _ in range(Bm)]
    hm = np.random.default_rng(1000).random((m, m), dtype=np.float32)
    for p in a + b:
        dev.h2d(p, hm)
    dev.sync()
    flops = 2.0 * m**3 * Bm
    for mode in ("tc", "simt"):
        os.environ["BCI_MATMUL"] = mode
        try:
            ms = timed(k_matmul(a, b, c, m), iters=3, warm=1)
            rows.append({"kernel": f"matmul f32 4096^3 [{mode}] (K3)", "config": f"C2 x{Bm}", "ms": ms,
                         "achieved": flops / (ms / 1e3) / 1e12, "unit": "TFLOP/s (algorithmic)",
                         "matmul_per_s": Bm / (ms / 1e3)})
        except Exception as e:  # noqa: BLE001
            rows.append({"kernel": f"matmul [{mode}]", "error": repr(e)})
    os.environ.pop("BCI_MATMUL", None)
    for p in a + b + c:
        dev.free(p)
    dev.sync()

    out = {"peak_hbm_GBps_measured": peak, "device": dev.device_info(), "rows": rows}
    if args.md:
        print("| kernel | config | ms | achieved (algorithmic) | frac of measured HBM | moved GB/s |")
        print("|---|---|---:|---:|---:|---:|")
        for r in rows:
            if "error" in r:
                print(f"| {r['kernel']} | error: {r['error']} | | | | |")
                continue
            frac = r.get("frac_of_measured_hbm")
            moved = r.get("moved_GBps")
            print(f"| {r['kernel']} | {r['config']} | {r['ms']:.3f} | {r['achieved']:.1f} {r['unit']} | "
                  f"{'' if frac is None else format(frac, '.2f')} | {'' if moved is None else format(moved, '.0f')} |")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
