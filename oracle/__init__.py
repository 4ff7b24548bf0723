"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's `/v1/execute` hot path, used as the checker by
`tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s CPU baseline / `--impl reference`
arm.  Nothing under `bee_code_interpreter_b200/` imports it; the product path has no CPU
fallback.

What is restated, and what it is pinned to:
  * executor semantics  — cpu_executor.py follows executor/server.rs:69-179 and
    services/kubernetes_code_executor.py:95-149; pinned by replaying the reference's own
    e2e test bodies (test/e2e/test_http.py:23-99) through the reference's own FastAPI app
    (services/http_server.py:77-106), recorded in tests/golden/http_*.json by
    tests/golden/make_golden.py (runs only where /root/reference exists).
  * numerics            — numpy_oracle.py calls stock numpy, which IS the reference's
    arithmetic (un-pinned Alpine py3-numpy, executor/Dockerfile:17,62; here numpy 2.3.x).
    The reference pins no numeric result (test_imports only checks a substring,
    test_http.py:31): numeric parity is "parity unpinned" in the reference and anchored on
    numpy itself + the committed .npz goldens.
  * SHA-256             — sha256.c (FIPS 180-4).  The reference computes no SHA-256
    (storage.py:52): "parity unpinned"; pinned to FIPS/NIST known answers and hashlib.
  * RNG                 — philox.c (Random123 Philox4x32-10 + numpy's 53-bit double);
    pinned to the Random123 known-answer vectors.
The Rust server cannot be compiled here (no rustc/cargo), so there is no oracle/_ref.
"""
