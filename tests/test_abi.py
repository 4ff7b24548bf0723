"""The C-ABI library loads and exports every symbol include/bci.h declares; entries fail
loudly (never silently fall back) when no GPU context exists.  No compute calls here."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (ROOT / "include" / "bci.h").read_text() + (ROOT / "include" / "bci_broker.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(bci_[a-z0-9_]+)\s*\(", text))


def test_binding_covers_exactly_the_header(built):
    from bee_code_interpreter_b200 import _lib

    assert set(_lib.SIGNATURES) == header_symbols()


def test_shared_library_exports_every_declared_symbol(built):
    from bee_code_interpreter_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\b(bci_[a-z0-9_]+)\b", out))
    assert header_symbols() <= exported
    lib = _lib.load()
    assert lib.bci_abi_version() == _lib.ABI_VERSION


def test_library_carries_sm_100a_code_only(built):
    from bee_code_interpreter_b200 import _lib

    out = subprocess.run(["cuobjdump", "--list-elf", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_entries_fail_loudly_without_a_device(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu tests")
    from bee_code_interpreter_b200 import _lib, device

    with pytest.raises(_lib.BciError):
        device.init(0)
    with pytest.raises(_lib.BciError) as e:
        device.alloc(1024)
    assert e.value.code == _lib.ERR_NOT_INIT
    with pytest.raises(_lib.BciError):
        device.sha256_host([b"abc"])
    st = _lib.Sha256State()
    assert _lib.load().bci_sha256_init(C.byref(st)) == 0  # pure host bookkeeping
    assert st.h[0] == 0x6A09E667 and st.total_len == 0
