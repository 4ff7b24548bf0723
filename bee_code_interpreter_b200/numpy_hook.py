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
          "min_rand": 1 << 16}


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

    random_view = _ModuleView(real_random, {"rand": rand, "random": random_sample, "random_sample": random_sample,
                                            "ranf": random_sample, "sample": random_sample, "seed": seed,
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

    overrides = {
        "random": random_view,
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
