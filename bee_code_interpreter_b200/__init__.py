"""bee_code_interpreter_b200 — a B200 (sm_100a) executor backend for the `/v1/execute`
hot path of bee-code-interpreter: numpy array operations of agent-submitted code and
SHA-256 content hashing of workspace files run in hand-written CUDA kernels behind a
batch-first C ABI (include/bci.h).  See DESIGN.md."""

__version__ = "0.1.0"
