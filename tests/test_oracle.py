"""The oracle checked against every pinned vector available for this path:
FIPS 180-4 / NIST known answers + hashlib for SHA-256, Random123 known answers for
Philox4x32-10, the numpy-generated numeric goldens, and the HTTP fixtures recorded through
the reference's own FastAPI app (tests/golden/make_golden.py)."""
import asyncio
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import numpy_oracle as O
from oracle.cpu_executor import LocalCpuExecutor, RefStorage

GOLDEN = Path(__file__).resolve().parent / "golden"

SHA_KATS = [
    (b"", "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"),
    (b"abc", "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"),
    (b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq",
     "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"),
    (b"abcdefghbcdefghicdefghijdefghijkefghijklfghijklmghijklmnhijklmnoijklmnopjklmnopqklmnopqrlmnopqrsmnopqrstnopqrstu",
     "cf5b16a778af8380036ce59e7b0492370b249b11e8f07a51afac45037afee9d1"),
    (b"Hello, World!", "dffd6021bb2bd5b0af676290809ec3a53191dd81c7f70a4b28688a362182986f"),  # test_http.py:48
    (b"a" * 1_000_000, "cdc76e5c9914fb9281a1c7e284d73e67f1809a48a497200e046d39ccc7112cd0"),
]


def test_sha256_c_known_answers(built):
    for msg, hexd in SHA_KATS:
        assert O.sha256_c(msg).hex() == hexd


def test_sha256_c_matches_hashlib_on_length_boundaries(built):
    rng = np.random.default_rng(1)
    for n in list(range(0, 260)) + [311, 447, 448, 511, 512, 513, 4095, 4096, 65537]:
        m = rng.bytes(n)
        assert O.sha256_c(m) == hashlib.sha256(m).digest()


PHILOX_KATS = [  # Random123 kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers(built):
    for ctr, key, want in PHILOX_KATS:
        assert tuple(O.philox4x32_10(ctr, key)) == want


def test_rand_f64_construction_and_distribution(built):
    x = O.rand_f64(200_001, 12345)
    r = O.philox4x32_10((0, 0, 0, 0), (12345, 0))
    assert x[0] == ((r[0] >> 5) * 67108864.0 + (r[1] >> 6)) / 9007199254740992.0  # numpy legacy random_double
    assert x[1] == ((r[2] >> 5) * 67108864.0 + (r[3] >> 6)) / 9007199254740992.0
    assert 0.0 <= x.min() and x.max() < 1.0
    n = x.size
    assert abs(x.mean() - 0.5) < 5 * np.sqrt(1 / 12 / n)
    assert abs(np.sum(np.square(x)) / n - 1 / 3) < 5 * np.sqrt(4 / 45 / n)  # the metric script's statistic
    assert not np.array_equal(x[:1000], O.rand_f64(1000, 12346))


def test_numeric_goldens_match_this_numpy(built):
    g = np.load(GOLDEN / "numeric.npz")
    for name in ("f32", "f64", "i32", "i64"):
        x, y = g[f"{name}_x"], g[f"{name}_y"]
        assert np.array_equal(O.ewise("add", x, y), g[f"{name}_add"])
        assert np.array_equal(O.ewise("multiply", x, y), g[f"{name}_mul"])
        assert np.array_equal(O.ewise("square", x), g[f"{name}_square"])
        assert np.array_equal(O.ewise("absolute", O.ewise("subtract", x, y)), g[f"{name}_abs_sub"])
        for red in ("sum", "mean", "max", "min"):
            want = g[f"{name}_{red}"]
            got = O.reduce(red, x)
            assert np.asarray(got).dtype == want.dtype
            assert np.isclose(got, want, rtol=1e-6 if name == "f32" else 1e-13)
        assert np.isclose(O.benchmark_numpy_body(x), g[f"{name}_sumsq"], rtol=1e-6 if name == "f32" else 1e-13)
        assert O.argext("argmax", x) == int(g[f"{name}_argmax"]) and O.argext("argmin", x) == int(g[f"{name}_argmin"])
        assert np.array_equal(O.sort(x), g[f"{name}_sort"])
        assert np.array_equal(O.argsort_stable(x), g[f"{name}_argsort"])
        if name != "f32":
            assert np.allclose(O.cumsum(x), g[f"{name}_cumsum"], rtol=1e-13, atol=0)
        else:  # documented deviation: float64 accumulation, numpy's float32 running sum drifts
            assert np.allclose(O.cumsum(x), g[f"{name}_cumsum"], rtol=1e-4, atol=1e-2)
    c = O.matmul_f64(g["mm_a"], g["mm_b"])
    assert np.abs(c - g["mm_c"]).max() <= 1e-5 * np.abs(c).max()


def _replay(executor, case):
    body = case["request"]
    return asyncio.run(executor.execute(body["source_code"], files=body.get("files", {}), env=body.get("env", {})))


def test_cpu_executor_reproduces_reference_app_fixtures(built, tmp_path):
    """Same requests that went through the reference's FastAPI app when the fixtures were
    recorded; the oracle executor alone must give the same Result fields."""
    cases = {c["name"]: c for c in json.loads((GOLDEN / "http_replay.json").read_text())["cases"]}
    ex = LocalCpuExecutor(RefStorage(tmp_path / "objects"), timeout=5)
    for name in ("hello", "env", "sys_exit_3", "stderr_and_unicode", "numpy_small"):
        r = _replay(ex, cases[name])
        want = cases[name]["response"]
        assert (r.stdout, r.stderr, r.exit_code, dict(r.files)) == (want["stdout"], want["stderr"], want["exit_code"], {})
    r = _replay(ex, cases["create_file"])
    assert r.exit_code == 0 and set(r.files) == {"/workspace/file.txt"}
    back = asyncio.run(ex.execute(cases["read_file_back"]["request"]["source_code"], files=dict(r.files)))
    assert back.stdout == cases["read_file_back"]["response"]["stdout"] and not back.files
    r = _replay(ex, cases["crash"])
    assert r.exit_code == cases["crash"]["response"]["exit_code"] == 1 and "ZeroDivisionError" in r.stderr
    r = _replay(ex, cases["subdir_not_listed"])
    assert set(r.files) == set(cases["subdir_not_listed"]["response"]["files"])
    r = _replay(ex, cases["timeout"])
    assert (r.stdout, r.stderr, r.exit_code) == ("", "Execution timed out", -1)
    with pytest.raises(FileNotFoundError):
        _replay(ex, cases["unknown_object"])


def test_pcg64_restatement_is_numpys_generator():
    """oracle.pcg64_* against numpy itself: raw stream, the three draw kinds, and the affine jump
    (P_k, G_k) against PCG64.advance — the arithmetic csrc/bci_rand_pcg.cu is built on."""
    for seed in (0, 7, 1000, 2**63 + 5):
        bg = np.random.PCG64(seed)
        st = bg.state["state"]
        state, inc = st["state"], st["inc"]
        assert O.pcg64_raw(state, inc, 64) == [int(v) for v in np.random.PCG64(seed).random_raw(64)]
        assert np.array_equal(O.pcg64_random(state, inc, 1001, "f64"), np.random.default_rng(seed).random(1001))
        assert np.array_equal(O.pcg64_random(state, inc, 1001, "f32"), np.random.default_rng(seed).random(1001, dtype=np.float32))
        assert np.array_equal(O.pcg64_random(state, inc, 999, "bounded", 0, 2**31),
                              np.random.default_rng(seed).integers(0, 2**31, 999, dtype=np.int64))
        assert O.pcg64_random(state, inc, 10, "u32").tobytes() == np.random.default_rng(seed).bytes(40)
        for k in (1, 8, 2040, 16384, 16384 * 256, 3 * 16384 * 65536 + 12345):
            P, G = O.pcg64_jump(k)
            ref = np.random.PCG64(seed)
            ref.advance(k)
            assert (P * state + inc * G) % 2**128 == ref.state["state"]["state"]
