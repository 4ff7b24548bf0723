"""Cold-start installation of the import hook: put THIS directory on PYTHONPATH, where the
reference puts executor/sitecustomize.py (executor/Dockerfile:109-111).  CPython's `site`
imports it at interpreter start, which replaces builtins.__import__ with
numpy_hook.patched_import — the reference's three post-import patches plus the numpy view for
agent code.  The warm path (zygote.py) installs the same hook without an interpreter start;
bci-executor-server uses this one when no zygote socket is configured."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.append(_root)          # after the script's own directory and PYTHONPATH entries
try:
    from bee_code_interpreter_b200 import numpy_hook as _hook

    _hook.install()
except Exception as _e:             # a broken hook must never break the script
    if os.environ.get("BCI_DEBUG_FALLBACK"):
        print(f"[bci] import hook not installed: {_e!r}", file=sys.stderr)
