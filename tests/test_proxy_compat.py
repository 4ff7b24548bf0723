"""Agent-code compatibility sweep: ~55 everyday numpy / pandas operations on device arrays (born from
`default_rng` through the hook) against the same operations on host twins — everything the kernels do not take
must come back exactly as numpy computes it (fallback on a host copy), with numpy's types and shapes."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_device  # noqa: E402

SCRIPT = r'''
import numpy as np
import pandas as pd
ok = []
def check(name, got, want):
    import numpy as _n
    good = _n.allclose(_n.asarray(got, dtype=float), _n.asarray(want, dtype=float), rtol=1e-12, atol=0, equal_nan=True)
    ok.append((name, bool(good)))
    if not good: print("MISMATCH", name, got, want)
import numpy as real   # same view; build host twins by seeding identically
rng = np.random.default_rng(1); x = rng.random(5000); y = rng.random(5000)
import importlib
rnp = importlib.import_module("numpy")          # library-style import: the real module
r2 = rnp.random.default_rng(1); xh = r2.random(5000); yh = r2.random(5000)
print(type(x).__name__, type(xh).__name__)
check("gt-mask-sum", (x > 0.5).sum(), (xh > 0.5).sum())
check("bool-index", x[x > 0.5].mean(), xh[xh > 0.5].mean())
check("where", np.where(x > y, x, y).sum(), rnp.where(xh > yh, xh, yh).sum())
check("concatenate", np.concatenate([x, y]).sum(), rnp.concatenate([xh, yh]).sum())
check("stack", np.stack([x, y]).shape, (2, 5000))
check("percentile", np.percentile(x, 90), rnp.percentile(xh, 90))
check("median", np.median(x), rnp.median(xh))
check("histogram", np.histogram(x, bins=10)[0], rnp.histogram(xh, bins=10)[0])
check("corrcoef", np.corrcoef(x, y)[0, 1], rnp.corrcoef(xh, yh)[0, 1])
check("linalg.norm", np.linalg.norm(x), rnp.linalg.norm(xh))
check("dot", x @ y, xh @ yh)
check("reshape-sum-axis", x.reshape(50, 100).sum(axis=0), xh.reshape(50, 100).sum(axis=0))
check("clip", np.clip(x, 0.2, 0.8).sum(), rnp.clip(xh, 0.2, 0.8).sum())
check("round", np.round(x, 2).sum(), rnp.round(xh, 2).sum())
check("power", (x ** 2).sum(), (xh ** 2).sum())
check("power3", (x ** 3).sum(), (xh ** 3).sum())
check("exp-log", np.log(np.exp(x)).sum(), rnp.log(rnp.exp(xh)).sum())
check("unique", len(np.unique(np.floor(x * 10))), len(rnp.unique(rnp.floor(xh * 10))))
check("cumprod", np.cumprod(x[:10])[-1], rnp.cumprod(xh[:10])[-1])
check("diff", np.diff(x).sum(), rnp.diff(xh).sum())
check("isnan", np.isnan(x).any(), False)
check("float()", float(x[3]), float(xh[3]))
check("len/iter", sum(1 for _ in x[:10]), 10)
check("f-string", float(f"{x.mean():.6f}"), float(f"{xh.mean():.6f}"))
check("tolist", x[:3].tolist(), xh[:3].tolist())
check("in-place", (lambda a: (a.__iadd__(1.0), a.sum())[1])(x * 1.0), (xh * 1.0 + 1.0).sum())
df = pd.DataFrame({"a": x, "b": y})
check("DataFrame", df["a"].mean() + df["b"].std(), pd.DataFrame({"a": xh, "b": yh})["a"].mean() + pd.Series(yh).std())
check("df.describe", df.describe().loc["mean", "a"], xh.mean())
check("df.corr", df.corr().iloc[0, 1], rnp.corrcoef(xh, yh)[0, 1])
check("argsort-take", x[np.argsort(x)[:5]], rnp.sort(xh)[:5])
check("sorted-search", np.searchsorted(np.sort(x), 0.5), rnp.searchsorted(rnp.sort(xh), 0.5))
check("astype-int", (x * 10).astype(int).sum(), (xh * 10).astype(int).sum())
check("max-axis", np.max(x.reshape(100, 50), axis=1).sum(), xh.reshape(100, 50).max(axis=1).sum())
check("std-ddof", np.std(x, ddof=1), rnp.std(xh, ddof=1))
check("mean-dtype", np.mean(x, dtype=np.float32), rnp.mean(xh, dtype=rnp.float32))
check("any/all", bool(np.all(x >= 0)), True)
check("zeros_like", np.zeros_like(x).sum(), 0.0)
check("ones*arr", (np.ones(5000) * x).sum(), xh.sum())
check("arange-idx", x[np.arange(10)].sum(), xh[:10].sum())
check("matmul2d", (x.reshape(50, 100) @ y.reshape(100, 50)).sum(), (xh.reshape(50, 100) @ yh.reshape(100, 50)).sum())
check("copy", x.copy().sum(), xh.sum())
check("T", x.reshape(50, 100).T.shape, (100, 50))
check("np.array(x)", np.array(x).sum(), xh.sum())
check("bytes", len(x.tobytes()), 40000)
import json as _j; check("json", _j.loads(_j.dumps(x[:3].tolist())), xh[:3].tolist())

import io as _io, pickle as _pk, copy as _cp
buf = _io.BytesIO(); np.save(buf, x); buf.seek(0); check("save-bytesio", rnp.load(buf).sum(), xh.sum())
check("pickle", _pk.loads(_pk.dumps(x)).sum(), xh.sum())
check("deepcopy", np.asarray(_cp.deepcopy(x)).sum(), xh.sum())
check("array_equal", np.array_equal(x, xh), True)
check("eq", (x == xh).all(), True)
check("repr", repr(x[:2]) == repr(xh[:2]), True)
check("min/max methods", x.min() + x.max(), xh.min() + xh.max())
check("sum of int draw", np.random.default_rng(3).integers(0, 2**31, 4000).sum(), rnp.random.default_rng(3).integers(0, 2**31, 4000).sum())
check("mixed host/device", (x + xh).sum(), 2 * xh.sum())
check("device/host matmul", x @ yh, xh @ yh)
check("sort method", (lambda a: (a.sort(), a[0], a[-1])[1:])(x.copy()), (xh.min(), xh.max()))
bad = [n for n, g in ok if not g]
print(len(ok), "checks", "FAILED: %s" % bad if bad else "all ok")
'''


def test_everyday_numpy_and_pandas_usage_on_device_arrays(built, monkeypatch):
    from bee_code_interpreter_b200 import numpy_hook as H, numpy_proxy as P
    from bee_code_interpreter_b200.backend import LocalBackend

    fake_device.reset()
    monkeypatch.setattr(P, "OFFLOAD_MIN_ELEMS", 1000)
    H._state["view"] = None
    H._state["backend"] = None
    H._state["lib_views"] = {}
    H.install(backend_factory=lambda: LocalBackend(fake_device), min_rand_elems=1000)
    glob = {"__name__": "__main__"}
    exec(compile(SCRIPT, "script.py", "exec"), glob)
    bad = [n for n, g in glob["ok"] if not g]
    assert len(glob["ok"]) >= 55 and not bad, bad
    kinds = {c[0] for c in fake_device.calls}
    assert {"rand_pcg", "reduce", "sort", "argsort"} <= kinds           # the device did take what it can
