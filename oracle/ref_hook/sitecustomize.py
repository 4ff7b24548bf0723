"""oracle/ref_hook/sitecustomize.py — TEST INFRASTRUCTURE ONLY.

Restatement of the reference's post-import hook (executor/sitecustomize.py:3-31): the
builtin `__import__` is wrapped; after an import statement whose dotted name is exactly
one of three library names, that library is monkey-patched so scripts running headless
leave files behind instead of opening windows.  The wrapper returns the module untouched.
No numpy branch here: this is the CPU baseline's hook.
"""
import builtins
import sys

_real_import = builtins.__import__


def _after_pyplot(_g, _l, _f, _lvl):
    plt = sys.modules["matplotlib.pyplot"]
    plt.show = lambda: sys.modules["matplotlib.pyplot"].savefig("plot.png")  # sitecustomize.py:9-12


def _after_moviepy(_g, _l, _f, _lvl):
    clip_cls = sys.modules["moviepy.editor"].VideoClip
    write = clip_cls.write_videofile
    clip_cls.write_videofile = lambda self, *a, **kw: write(self, *a, verbose=False, logger=None, **kw)  # :13-22


def _after_pil(g, l, f, lvl):
    _real_import("PIL.ImageShow", g, l, f, lvl)
    sys.modules["PIL.ImageShow"].show = lambda img, *a, **kw: img.save("image.png")  # :23-27


_POST_IMPORT = {"matplotlib.pyplot": _after_pyplot, "moviepy.editor": _after_moviepy, "PIL": _after_pil}


def _hooked_import(name, globals=None, locals=None, fromlist=(), level=0):
    module = _real_import(name, globals, locals, fromlist, level)
    patch = _POST_IMPORT.get(name)
    if patch is not None:
        patch(globals, locals, fromlist, level)
    return module


builtins.__import__ = _hooked_import
