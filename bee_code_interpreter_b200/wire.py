"""Wire protocol of the per-GPU broker (include/bci_broker.h), shared by the script-side
client (broker_client.py), the native broker inside libbci.so and its Python stand-in used
with the test double (broker.py)."""
from __future__ import annotations

import struct

ALLOC, FREE, RAND, EWISE, REDUCE, REDUCE_FUSED, ARGEXT, CUMSUM, SORT, ARGSORT, MATMUL, COPY, UPLOAD, DOWNLOAD, \
    SHA256_DEV, SHA256_FILE, SHA256_BYTES, SHA256_UPDATE, SHA256_FINAL, PING, RAND_MT, RAND_NORMAL = range(1, 23)
MT_STATE_WORDS = 625   # MT19937 key[624] + pos, little-endian uint32

F_IN1, F_ADDEND, F_INPLACE, F_SHM, F_STATE = 1, 2, 4, 8, 16

# bci_wire_req: op, code, dtype, pre, aux, flags | p0, p1, n | i0, i1, i2 | s0, s1 | payload_len, reserved
REQ = struct.Struct("<6I3Q3q2QII")
REP = struct.Struct("<iIQ")
assert REQ.size == 96 and REP.size == 16

INLINE_MAX = 1 << 16   # array payloads up to 64 KiB travel inline; larger ones through /dev/shm files


def scalar_bits(value, is_float: bool) -> int:
    """bci_scalar bit pattern: double for float dtypes, int64 otherwise."""
    if value is None:
        return 0
    if is_float:
        return struct.unpack("<Q", struct.pack("<d", float(value)))[0]
    return int(value) & 0xFFFFFFFFFFFFFFFF


def scalar_from_bits(bits: int, is_float: bool):
    if is_float:
        return struct.unpack("<d", struct.pack("<Q", bits))[0]
    return struct.unpack("<q", struct.pack("<Q", bits))[0]


def pack_req(op, code=0, dtype=0, pre=0, aux=0, flags=0, p0=0, p1=0, n=0, i0=0, i1=0, i2=0, s0=0, s1=0, payload=b""):
    return REQ.pack(op, code, dtype, pre, aux, flags, p0 or 0, p1 or 0, n, i0, i1, i2, s0, s1, len(payload), 0) + payload


def recv_exact(sock, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise EOFError("connection closed")
        buf += chunk
    return bytes(buf)
