"""Import hook for agent code (boundary ii): the reference wraps `builtins.__import__` in a
`sitecustomize` module and monkey-patches three libraries after import
(executor/sitecustomize.py:3-31).  This module keeps that contract — same wrapper signature,
same three post-import patches, module returned unchanged for everyone else — and adds the
numpy branch the north star asks for.

numpy branch.  When the importing code is the agent's script (its globals say
`__name__ == "__main__"`), `import numpy` / `import numpy as np` / `from numpy import sum` /
`import numpy.random` resolve to a thin *view* of the real numpy module in which a handful of
entry points are rebound:
    numpy.random.rand / random / random_sample / ranf / sample   -> Philox on the device
         (only while the legacy global generator is unseeded: after numpy.random.seed(...)
          the stream is a parity target — examples/using_imports.py:20-22 — and stays on numpy)
    numpy.sort / argsort / cumsum / matmul / dot                  -> upload large host arrays
Everything else on the view is the real numpy attribute; DeviceArray results flow through
numpy's own `__array_ufunc__` / `__array_function__` dispatch.  Libraries (pandas, scipy)
importing numpy get the untouched module, so their internals never see a proxy unless the
agent hands them one (then `__array__` materialises it).  The hook is idempotent: the view
is built once per process.
"""
from __future__ import annotations

import builtins
import os
import sys
import types

_real_import = builtins.__import__
_state = {"backend": None, "backend_factory": None, "view": None, "seeded": False, "installed": False,
          "min_rand": 1 << 16, "lib_views": {}}


# ---- the reference's three post-import patches (executor/sitecustomize.py:9-27) -------------
def _patch_pyplot(_g, _l, _f, _lvl):
    plt = sys.modules["matplotlib.pyplot"]
    plt.show = lambda: sys.modules["matplotlib.pyplot"].savefig("plot.png")


def _patch_moviepy(_g, _l, _f, _lvl):
    cls = sys.modules["moviepy.editor"].VideoClip
    original = cls.write_videofile
    cls.write_videofile = lambda self, *a, **kw: original(self, *a, verbose=False, logger=None, **kw)


def _patch_pil(g, l, f, lvl):
    _real_import("PIL.ImageShow", g, l, f, lvl)
    sys.modules["PIL.ImageShow"].show = lambda img, *a, **kw: img.save("image.png")


_POST_IMPORT = {"matplotlib.pyplot": _patch_pyplot, "moviepy.editor": _patch_moviepy, "PIL": _patch_pil}


# ---- numpy view ---------------------------------------------------------------------------------
class _ModuleView(types.ModuleType):
    """Attribute-transparent view of a module with a few names rebound."""

    def __init__(self, real, overrides):
        super().__init__(real.__name__, getattr(real, "__doc__", None))
        object.__setattr__(self, "_real", real)
        object.__setattr__(self, "_overrides", overrides)
        self.__dict__.update({"__file__": getattr(real, "__file__", None), "__path__": getattr(real, "__path__", None),
                              "__package__": getattr(real, "__package__", None), "__spec__": None,
                              "__version__": getattr(real, "__version__", None)})

    def __getattr__(self, name):
        ov = object.__getattribute__(self, "_overrides")
        if name in ov:
            return ov[name]
        return getattr(object.__getattribute__(self, "_real"), name)

    def __dir__(self):
        return dir(object.__getattribute__(self, "_real"))


def _backend():
    if _state["backend"] is None:
        factory = _state["backend_factory"]
        if factory is None:
            from .broker_client import RemoteBackend
            factory = RemoteBackend
        _state["backend"] = factory()
    return _state["backend"]


def _build_view():
    import numpy as real_np

    from . import _lib as L
    from . import numpy_proxy as P

    real_random = real_np.random

    def _size_of(shape):
        n = 1
        for s in shape:
            n *= int(s)
        return n

    def _device_uniform(shape, fallback):
        try:
            if _size_of(shape) < _state["min_rand"]:
                return fallback()
            if _state["seeded"]:   # deterministic stream: the device continues numpy's MT19937 sequence
                return P.device_rand_legacy(_backend(), tuple(int(s) for s in shape), real_random)
            seed = int.from_bytes(os.urandom(8), "little")
            return P.device_rand(_backend(), tuple(int(s) for s in shape), seed)
        except Exception as e:
            if os.environ.get("BCI_DEBUG_FALLBACK"):
                print(f"[bci] device rand fell back to numpy: {e!r}", file=sys.stderr)
            return fallback()

    def rand(*shape):
        if not shape or not all(isinstance(s, (int, real_np.integer)) for s in shape):
            return real_random.rand(*shape)
        return _device_uniform(shape, lambda: real_random.rand(*shape))

    def random_sample(size=None):
        if size is None:
            return real_random.random_sample()
        shape = (size,) if isinstance(size, (int, real_np.integer)) else tuple(size)
        return _device_uniform(shape, lambda: real_random.random_sample(size))

    def seed(*a, **kw):
        _state["seeded"] = True  # from here on the legacy stream is deterministic: numpy owns it
        return real_random.seed(*a, **kw)

    def set_state(*a, **kw):
        _state["seeded"] = True  # a restored state is as deterministic as a seed
        return real_random.set_state(*a, **kw)

    def _device_normal(shape, loc, scale, fallback):
        # unseeded only: the seeded legacy normal is MT19937 + polar rejection, numpy keeps that stream
        try:
            plain = all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in (loc, scale))
            if (_state["seeded"] or not plain or scale < 0 or _size_of(shape) < _state["min_rand"]
                    or not all(isinstance(s, (int, real_np.integer)) for s in shape)):
                return fallback()
            return P.device_normal(_backend(), tuple(int(s) for s in shape), int.from_bytes(os.urandom(8), "little"),
                                   loc, scale)
        except Exception as e:
            if os.environ.get("BCI_DEBUG_FALLBACK"):
                print(f"[bci] device normal fell back to numpy: {e!r}", file=sys.stderr)
            return fallback()

    def randn(*shape):
        if not shape:
            return real_random.randn()
        return _device_normal(shape, 0.0, 1.0, lambda: real_random.randn(*shape))

    def standard_normal(size=None):
        if size is None:
            return real_random.standard_normal()
        shape = (size,) if isinstance(size, (int, real_np.integer)) else tuple(size)
        return _device_normal(shape, 0.0, 1.0, lambda: real_random.standard_normal(size))

    def normal(loc=0.0, scale=1.0, size=None):
        if size is None:
            return real_random.normal(loc, scale)
        shape = (size,) if isinstance(size, (int, real_np.integer)) else tuple(size)
        return _device_normal(shape, loc, scale, lambda: real_random.normal(loc, scale, size))

    # ---- numpy.random.Generator over PCG64 (default_rng): the generator BASELINE configs C2/C3/C5 draw from
    f64, f32, i64, i32 = (real_np.dtype(t) for t in (real_np.float64, real_np.float32, real_np.int64, real_np.int32))

    def _shape_of(size):
        if size is None:
            return None
        if isinstance(size, (int, real_np.integer)):
            return (int(size),)
        try:
            shape = tuple(int(v) for v in size)
        except TypeError:
            return None
        return shape if all(v >= 0 for v in shape) else None

    def _plain_int(v):
        return isinstance(v, (int, real_np.integer)) and not isinstance(v, (bool, real_np.bool_))

    class DeviceGenerator(real_random.Generator):
        """numpy.random.Generator whose large draws are produced on the device, bit for bit numpy's
        stream (csrc/bci_rand_pcg.cu); everything else — and every case the kernel does not take — is
        the inherited numpy method on the same bit generator."""

        def _device(self, shape, kind, low=0, span=0):
            if shape is None or _size_of(shape) < _state["min_rand"]:
                return None
            try:
                return P.device_pcg(_backend(), self, shape, kind, low, span)
            except Exception as e:
                if os.environ.get("BCI_DEBUG_FALLBACK"):
                    print(f"[bci] device Generator draw fell back to numpy: {e!r}", file=sys.stderr)
                return None

        def random(self, size=None, dtype=real_np.float64, out=None):
            if out is None:
                try:
                    dt = real_np.dtype(dtype)
                except TypeError:
                    dt = None
                kind = L.PCG_F64 if dt == f64 else L.PCG_F32 if dt == f32 else None
                if kind is not None:
                    got = self._device(_shape_of(size), kind)
                    if got is not None:
                        return got
            return super().random(size, dtype, out)

        def integers(self, low, high=None, size=None, dtype=real_np.int64, endpoint=False):
            try:
                dt = real_np.dtype(dtype)
            except TypeError:
                dt = None
            if dt in (i64, i32) and _plain_int(low) and (high is None or _plain_int(high)):
                lo, hi = (0, int(low)) if high is None else (int(low), int(high))
                if endpoint:
                    hi += 1
                info = real_np.iinfo(dt)
                # numpy's 32-bit Lemire path: inclusive range 1 .. 2^32 - 2 (it validates the bounds itself otherwise)
                if info.min <= lo and hi - 1 <= info.max and 2 <= hi - lo <= 0xFFFFFFFF:
                    got = self._device(_shape_of(size), L.PCG_BOUNDED32_I64 if dt == i64 else L.PCG_BOUNDED32_I32, lo, hi - lo)
                    if got is not None:
                        return got
            return super().integers(low, high, size, dtype, endpoint)

        def uniform(self, low=0.0, high=1.0, size=None):
            plain = all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in (low, high))
            if plain and real_np.isfinite(float(high) - float(low)):
                got = self._device(_shape_of(size), L.PCG_F64)
                if got is not None:
                    # random_uniform(): off + rng * next_double, product and sum each rounded once
                    return float(low) + (float(high) - float(low)) * got
            return super().uniform(low, high, size)

    def default_rng(seed=None):
        if isinstance(seed, real_random.Generator) and not isinstance(seed, DeviceGenerator):
            return DeviceGenerator(seed.bit_generator)
        g = real_random.default_rng(seed)
        return g if isinstance(g, DeviceGenerator) else DeviceGenerator(g.bit_generator)

    random_view = _ModuleView(real_random, {"default_rng": default_rng, "Generator": DeviceGenerator, "rand": rand, "random": random_sample, "random_sample": random_sample,
                                            "ranf": random_sample, "sample": random_sample, "seed": seed, "set_state": set_state,
                                            "randn": randn, "standard_normal": standard_normal, "normal": normal})

    def _offloading(device_fn, real_fn, n_arrays):
        def wrapper(*args, **kwargs):
            try:
                if len(args) >= n_arrays and any(isinstance(a, (real_np.ndarray, P.DeviceArray)) for a in args[:n_arrays]):
                    be = _backend()
                    dev_args = [P.offload(be, a) for a in args[:n_arrays]]
                    return device_fn(*dev_args, *args[n_arrays:], **kwargs)
            except Exception:
                pass
            return real_fn(*P._to_host(args), **P._to_host(kwargs))
        wrapper.__name__ = real_fn.__name__
        wrapper.__doc__ = real_fn.__doc__
        return wrapper

    # ---- arrays born on the device: creation functions and numpy.load above the offload threshold
    def _creation(real_fn, device_fn):
        def wrapper(*args, **kwargs):
            try:
                got = device_fn(*args, **kwargs)
                if got is not None:
                    return got
            except Exception as e:
                if os.environ.get("BCI_DEBUG_FALLBACK"):
                    print(f"[bci] {real_fn.__name__} stayed on the host: {e!r}", file=sys.stderr)
            return real_fn(*args, **kwargs)
        wrapper.__name__ = real_fn.__name__
        wrapper.__doc__ = real_fn.__doc__
        return wrapper

    def _big_shape(shape):
        shp = _shape_of(shape)
        return shp if shp is not None and _size_of(shp) >= P.OFFLOAD_MIN_ELEMS else None

    def _full_like(value):
        def fn(shape, dtype=float, order="C", *, device=None, like=None):
            shp = _big_shape(shape)
            if shp is None or order not in ("C", "K", "A") or device is not None or like is not None:
                return None
            return P.device_full(_backend(), shp, real_np.dtype(dtype), value)
        return fn

    def _full(shape, fill_value, dtype=None, order="C", *, device=None, like=None):
        shp = _big_shape(shape)
        if shp is None or order not in ("C", "K", "A") or device is not None or like is not None or \
                not isinstance(fill_value, (int, float)) or isinstance(fill_value, bool):
            return None
        dt = real_np.dtype(dtype) if dtype is not None else real_np.asarray(fill_value).dtype
        return P.device_full(_backend(), shp, dt, fill_value)

    def _arange(*args, dtype=None, device=None, like=None):
        if device is not None or like is not None or not 1 <= len(args) <= 3 or \
                not all(isinstance(a, (int, float)) and not isinstance(a, bool) for a in args):
            return None
        start, stop, step = (0, args[0], 1) if len(args) == 1 else (args[0], args[1], 1) if len(args) == 2 else args
        if step == 0:
            return None
        dt = real_np.dtype(dtype) if dtype is not None else (
            real_np.dtype(real_np.int64) if all(isinstance(a, int) for a in (start, stop, step)) else f64)
        return P.device_arange(_backend(), start, stop, step, dt)

    def _linspace(start, stop, num=50, endpoint=True, retstep=False, dtype=None, axis=0, *, device=None):
        if retstep or dtype not in (None, float, real_np.float64) or axis != 0 or device is not None or \
                not all(isinstance(a, (int, float)) and not isinstance(a, bool) for a in (start, stop)) or \
                not isinstance(num, int) or num < 2:
            return None
        return P.device_linspace(_backend(), start, stop, num, bool(endpoint))

    def _load(file, mmap_mode=None, allow_pickle=False, fix_imports=True, encoding="ASCII", *, max_header_size=10000):
        if mmap_mode is not None or not (isinstance(file, str) or hasattr(file, "__fspath__")) or \
                not os.fspath(file).endswith(".npy"):
            return None
        return P.device_load(_backend(), os.fspath(file))

    overrides = {
        "random": random_view,
        "zeros": _creation(real_np.zeros, _full_like(0)),
        "ones": _creation(real_np.ones, _full_like(1)),
        "empty": _creation(real_np.empty, _full_like(0)),
        "full": _creation(real_np.full, _full),
        "arange": _creation(real_np.arange, _arange),
        "linspace": _creation(real_np.linspace, _linspace),
        "load": _creation(real_np.load, _load),
        "sort": _offloading(P.sort, real_np.sort, 1),
        "argsort": _offloading(P.argsort, real_np.argsort, 1),
        "cumsum": _offloading(P.cumsum, real_np.cumsum, 1),
        "matmul": _offloading(P.matmul, real_np.matmul, 2),
        "dot": _offloading(P._dot, real_np.dot, 2),
    }
    return _ModuleView(real_np, overrides)


def numpy_view():
    if _state["view"] is None:
        _state["view"] = _build_view()
    return _state["view"]


# ---- pandas / scipy entry points agent code reaches device arrays through -----------------
# examples/using_imports.py:25-36: pd.Series(x).mean() / .std(), scipy.stats.ttest_ind(a, b).  When `x` is a device
# array these run as the fused mean / variance reductions; anything else is pandas / scipy itself on a host copy.
def _build_pandas_view():
    import numpy as real_np
    import pandas as real_pd

    from . import _lib as L
    from . import numpy_proxy as P

    class DeviceSeries:
        """pd.Series over a 1-D float device array: the reductions of the reference's example run on the
        device (pandas' own formulas: nanops.nanmean = sum / count, nanvar = two-pass with ddof); every other
        attribute is the real pandas Series built from a host copy (once)."""

        def __init__(self, data, kwargs):
            self._dev, self._kwargs, self._real = data, kwargs, None

        def _series(self):
            if self._real is None:
                self._real = real_pd.Series(real_np.asarray(self._dev), **self._kwargs)
            return self._real

        def _plain(self, args, kw, allowed=()):
            return not args and all(k in allowed for k in kw)

        def _clean(self):
            """pandas skips NaN by default: only NaN-free data takes the device formulas (max propagates NaN)."""
            try:
                return not real_np.isnan(self._dev.max())
            except Exception:
                return False

        def mean(self, *a, **kw):
            if self._plain(a, kw) and self._clean():
                return float(self._dev.mean())
            return self._series().mean(*a, **kw)

        def sum(self, *a, **kw):
            if self._plain(a, kw) and self._clean():
                return float(self._dev.sum())
            return self._series().sum(*a, **kw)

        def var(self, *a, **kw):
            if self._plain(a, kw, ("ddof",)) and self._clean():
                return float(self._dev.var(ddof=kw.get("ddof", 1)))       # pandas' default is ddof=1
            return self._series().var(*a, **kw)

        def std(self, *a, **kw):
            if self._plain(a, kw, ("ddof",)) and self._clean():
                return float(real_np.sqrt(self._dev.var(ddof=kw.get("ddof", 1))))
            return self._series().std(*a, **kw)

        def max(self, *a, **kw):
            return float(self._dev.max()) if self._plain(a, kw) and self._clean() else self._series().max(*a, **kw)

        def min(self, *a, **kw):
            return float(self._dev.min()) if self._plain(a, kw) and self._clean() else self._series().min(*a, **kw)

        def median(self, *a, **kw):
            # nanops.nanmedian of NaN-free data is np.median of the values: one device sort, two elements back
            if self._plain(a, kw) and len(self._dev) >= 2 and self._clean():
                return real_np.median(self._dev)
            return self._series().median(*a, **kw)

        def quantile(self, q=0.5, *a, **kw):
            # pandas (core/array_algos/quantile.py): np.percentile(values, asarray(q) * 100.0, method=interpolation)
            if self._plain(a, kw) and isinstance(q, (int, float)) and not isinstance(q, bool) and 0 <= q <= 1 \
                    and len(self._dev) >= 2 and self._clean():
                return real_np.percentile(self._dev, real_np.asarray([q], dtype=real_np.float64) * 100.0)[0]
            return self._series().quantile(q, *a, **kw)

        def count(self):
            return int(self._dev.size) if self._clean() else self._series().count()

        def to_numpy(self, *a, **kw):
            return self._dev if not a and not kw else self._series().to_numpy(*a, **kw)

        @property
        def values(self):
            return self._dev

        def __len__(self):
            return len(self._dev)

        def __repr__(self):
            return repr(self._series())

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self._series(), name)

    class _SeriesMeta(type):
        def __instancecheck__(cls, obj):
            return isinstance(obj, (real_pd.Series, DeviceSeries))

        def __call__(cls, data=None, *args, **kwargs):
            if (isinstance(data, P.DeviceArray) and not args and data.ndim == 1 and data.dtype.kind == "f"
                    and all(k in ("name", "index") and (k == "name" or kwargs[k] is None) for k in kwargs)):
                return DeviceSeries(data, kwargs)
            return real_pd.Series(P._to_host(data), *P._to_host(args), **P._to_host(kwargs))

        def __getattr__(cls, name):
            return getattr(real_pd.Series, name)

    class Series(metaclass=_SeriesMeta):
        __doc__ = real_pd.Series.__doc__

    return _ModuleView(real_pd, {"Series": Series})


def _build_scipy_views():
    import numpy as real_np
    import scipy as real_scipy
    import scipy.stats as real_stats

    from . import numpy_proxy as P

    def ttest_ind(a, b, *args, **kwargs):
        """scipy.stats.ttest_ind for two 1-D float device arrays (default options, or equal_var / alternative):
        sizes, means and ddof=1 variances come from the fused reductions, the statistic and p-value from
        scipy's own ttest_ind_from_stats — the arrays themselves stay in HBM."""
        dev = all(isinstance(x, P.DeviceArray) and x.ndim == 1 and x.dtype.kind == "f" and x.size > 1 for x in (a, b))
        if dev and not args and all(k in ("equal_var", "alternative") for k in kwargs):
            try:
                m1, m2 = float(a.mean()), float(b.mean())
                v1, v2 = float(a.var(ddof=1)), float(b.var(ddof=1))
                if not (real_np.isnan(m1) or real_np.isnan(m2)):          # NaN handling (nan_policy) is scipy's business
                    return real_stats.ttest_ind_from_stats(m1, real_np.sqrt(v1), a.size, m2, real_np.sqrt(v2), b.size, **kwargs)
            except Exception:
                pass
        return real_stats.ttest_ind(P._to_host(a), P._to_host(b), *P._to_host(args), **P._to_host(kwargs))

    ttest_ind.__doc__ = real_stats.ttest_ind.__doc__
    stats_view = _ModuleView(real_stats, {"ttest_ind": ttest_ind})
    return _ModuleView(real_scipy, {"stats": stats_view}), stats_view


def _library_view(name: str, fromlist):
    """Views of pandas / scipy(.stats) for agent code; built on first use, None when the library is not involved."""
    if name == "pandas":
        if "pandas" not in _state["lib_views"]:
            _state["lib_views"]["pandas"] = _build_pandas_view()
        return _state["lib_views"]["pandas"]
    if name == "scipy" or name == "scipy.stats":
        if "scipy" not in _state["lib_views"]:
            _state["lib_views"]["scipy"], _state["lib_views"]["scipy.stats"] = _build_scipy_views()
        if name == "scipy.stats" and fromlist:
            return _state["lib_views"]["scipy.stats"]      # `from scipy.stats import ttest_ind`
        return _state["lib_views"]["scipy"]                # `import scipy.stats`, `from scipy import stats`
    return None


def _from_agent_code(globals_) -> bool:
    return isinstance(globals_, dict) and globals_.get("__name__") == "__main__"


def patched_import(name, globals=None, locals=None, fromlist=(), level=0):
    module = _real_import(name, globals, locals, fromlist, level)
    patch = _POST_IMPORT.get(name)
    if patch is not None:
        patch(globals, locals, fromlist, level)
    elif level == 0 and (name == "numpy" or name.startswith("numpy.")) and _from_agent_code(globals):
        try:
            view = numpy_view()
            if name == "numpy" or not fromlist:
                return view                      # `import numpy`, `import numpy.random` bind the top package
            if name == "numpy.random":
                return view.random               # `from numpy.random import rand`
        except Exception:
            return module                        # hook failure must never break the script
    elif level == 0 and name in ("pandas", "scipy", "scipy.stats") and _from_agent_code(globals):
        try:
            view = _library_view(name, fromlist)
            if view is not None:
                return view
        except Exception:
            return module
    return module


def install(backend_factory=None, min_rand_elems: int | None = None) -> None:
    """Installs the wrapper as builtins.__import__ (idempotent)."""
    if backend_factory is not None:
        _state["backend_factory"] = backend_factory
        _state["backend"] = None
    if min_rand_elems is not None:
        _state["min_rand"] = int(min_rand_elems)
    if not _state["installed"]:
        builtins.__import__ = patched_import
        _state["installed"] = True


def reset_request_state() -> None:
    """Per-request state of a forked script process (fresh connection, unseeded generator)."""
    _state["backend"] = None
    _state["seeded"] = False
