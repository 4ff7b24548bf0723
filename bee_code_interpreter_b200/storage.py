"""Content-addressed object store (boundary iii): the reference's `Storage` protocol
(services/storage.py:24-90) with the property its docstring promises and its code does not
deliver (storage.py:35-39 vs :52): an object's id IS the SHA-256 of its bytes.

Same surface: `writer()` async context manager yielding an object with `.write(data)` and
`.hash`; `write(data) -> id`; `reader(id)`; `read(id)`; `exists(id)`; unknown id ->
FileNotFoundError.  `.hash` is final when the `async with` block exits, which is when the
reference's only call site reads it (kubernetes_code_executor.py:127-134).

Hashing runs on the GPU: chunks from all concurrent writers are absorbed together, one
`bci_sha256_update_batch` launch per round (HashService).  There is no CPU hashing path: if
the device library is unavailable, writing fails loudly.
"""
from __future__ import annotations

import asyncio
import os
import queue
import re
import secrets
import threading
from contextlib import asynccontextmanager
from typing import AsyncIterator

from anyio import Path

HASH_RE = re.compile(r"^[0-9a-zA-Z_-]{1,255}$")  # utils/validation.py:19


class HashService:
    """Batches SHA-256 update/final calls from concurrent writers into single launches."""

    def __init__(self, device_module=None, coalesce_us: int = 200):
        if device_module is None:
            from . import device as device_module
            if not device_module.is_initialised():
                device_module.init()
        self.dev = device_module
        self.coalesce_s = coalesce_us / 1e6
        self._q: "queue.SimpleQueue" = queue.SimpleQueue()
        self._thread = threading.Thread(target=self._loop, name="bci-hash", daemon=True)
        self._running = True
        self.stats = {"launches": 0, "chunks": 0, "max_batch": 0}
        self._thread.start()

    def new_stream(self):
        return self.dev.Sha256Stream()

    def update(self, stream, data: bytes) -> None:
        self._submit(("update", stream, data))

    def final(self, stream) -> bytes:
        return self._submit(("final", stream, None))

    def _submit(self, item):
        done = threading.Event()
        box: list = [None, None]
        self._q.put((item, done, box))
        done.wait()
        if box[1] is not None:
            raise box[1]
        return box[0]

    def hash_fd(self, fd: int) -> bytes:
        """SHA-256 of an open regular file (the descriptor stays open and is read from offset 0)."""
        stream = self.new_stream()
        off = 0
        while True:
            chunk = os.pread(fd, 8 << 20, off)
            if not chunk:
                break
            off += len(chunk)
            self.update(stream, chunk)
        return self.final(stream)

    def close(self):
        self._running = False
        self._q.put(None)

    def _loop(self):
        import time

        while self._running:
            first = self._q.get()
            if first is None:
                break
            items = [first]
            if self.coalesce_s:
                time.sleep(self.coalesce_s)
            while True:
                try:
                    nxt = self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    self._running = False
                    break
                items.append(nxt)
            updates = [it for it in items if it[0][0] == "update"]
            finals = [it for it in items if it[0][0] == "final"]
            # a stream appears at most once per round (its writer awaits each call)
            for group, fn in ((updates, self._do_updates), (finals, self._do_finals)):
                if not group:
                    continue
                try:
                    fn(group)
                except Exception as e:
                    for _, _, box in group:
                        box[1] = e
                for _, done, _ in group:
                    done.set()

    def _do_updates(self, group):
        self.dev.sha256_update_many([it[0][1] for it in group], [it[0][2] for it in group])
        self.stats["launches"] += 1
        self.stats["chunks"] += len(group)
        self.stats["max_batch"] = max(self.stats["max_batch"], len(group))

    def _do_finals(self, group):
        digests = self.dev.sha256_final_many([it[0][1] for it in group])
        self.stats["launches"] += 1
        for (_, _, box), d in zip(group, digests):
            box[0] = d


class _RemoteStream:
    __slots__ = ("state",)

    def __init__(self):
        self.state = None  # serialized bci_sha256_state, travels with every chunk


class RemoteHashService:
    """Same interface as HashService, but the GPU belongs to the broker process: chunks are
    sent to it and batched there with the chunks of every other writer (broker.py).  One
    broker connection per thread (asyncio.to_thread workers)."""

    def __init__(self, socket_path: str):
        self.socket_path = socket_path
        self._local = threading.local()

    def _backend(self):
        be = getattr(self._local, "be", None)
        if be is None:
            from .broker_client import RemoteBackend
            be = self._local.be = RemoteBackend(self.socket_path)
        return be

    def new_stream(self):
        return _RemoteStream()

    def update(self, stream, data: bytes) -> None:
        stream.state = self._backend().sha256_update(stream.state, data)

    def final(self, stream) -> bytes:
        if stream.state is None:
            stream.state = self._backend().sha256_update(None, b"")
        return self._backend().sha256_final(stream.state)

    def hash_fd(self, fd: int) -> bytes:
        return self._backend().sha256_fd(fd)     # the descriptor crosses the socket (SCM_RIGHTS), never a path

    def close(self):
        pass


class ObjectWriter:
    """services/storage.py:28-31 — `.hash` + `async write(data)`."""

    FLUSH_BYTES = 4 << 20  # hash in >= 4 MiB slices so one launch carries real work

    def __init__(self, file, hasher: HashService):
        self._file = file
        self._hasher = hasher
        self._stream = hasher.new_stream()
        self._pending = bytearray()
        self.hash: str = ""

    async def write(self, data: bytes) -> None:
        await self._file.write(data)
        self._pending += data
        if len(self._pending) >= self.FLUSH_BYTES:
            await self._flush()

    async def _flush(self):
        if self._pending:
            chunk = bytes(self._pending)
            self._pending.clear()
            await asyncio.to_thread(self._hasher.update, self._stream, chunk)

    async def _finish(self) -> str:
        await self._flush()
        digest = await asyncio.to_thread(self._hasher.final, self._stream)
        self.hash = digest.hex()
        return self.hash


class ContentAddressedStorage:
    def __init__(self, storage_path: str, hasher: HashService | None = None):
        self.storage_path = Path(storage_path)
        self._hasher = hasher

    @property
    def hasher(self) -> HashService:
        if self._hasher is None:
            self._hasher = HashService()
        return self._hasher

    @asynccontextmanager
    async def writer(self) -> AsyncIterator[ObjectWriter]:
        await self.storage_path.mkdir(parents=True, exist_ok=True)
        tmp = self.storage_path / (".tmp-" + secrets.token_hex(16))
        try:
            async with await tmp.open("wb") as file:
                w = ObjectWriter(file, self.hasher)
                yield w
                await w._finish()
            target = self.storage_path / w.hash
            if await target.exists():
                await tmp.unlink()          # same content already stored: dedupe
            else:
                await tmp.rename(target)
        except BaseException:
            if await tmp.exists():
                await tmp.unlink()
            raise

    async def write(self, data: bytes) -> str:
        async with self.writer() as f:
            await f.write(data)
        return f.hash

    async def put_file(self, path: str) -> str:
        """Stores an existing file (an executor output) under its SHA-256: hashed where it
        lies, then moved into the store (copy across filesystems).  Not in the reference's
        Storage; the streaming `writer()` above is the protocol-compatible path.

        Only regular-file CONTENT gets in: the bytes that are hashed and the bytes that are stored are
        those of one open descriptor.  A symlink (the reference lists links to regular files as changed
        files, server.rs:103, and downloads what they point to) is read through, its bytes are COPIED,
        and the link itself never reaches the store, so it can neither dangle there nor redirect an id
        to a file that changes later.  Dedupe is a hard link with O_EXCL semantics (`os.link` fails if
        the id exists), so two writers of the same content cannot race."""
        import shutil
        import stat as stat_mod

        is_link = os.path.islink(str(path))
        fd = await asyncio.to_thread(os.open, str(path), os.O_RDONLY | os.O_CLOEXEC | (0 if is_link else os.O_NOFOLLOW))
        try:
            st = os.fstat(fd)
            if not stat_mod.S_ISREG(st.st_mode):
                raise ValueError(f"{path!r} is not a regular file")
            digest = (await asyncio.to_thread(self.hasher.hash_fd, fd)).hex()
            await self.storage_path.mkdir(parents=True, exist_ok=True)
            target = self.storage_path / digest
            if not await target.exists():
                tmp = str(self.storage_path / (".tmp-" + secrets.token_hex(16)))

                def place():
                    try:
                        if is_link:
                            raise OSError("links are copied, never moved")
                        os.rename(str(path), tmp)                       # same filesystem: the inode we hashed
                        if os.stat(tmp, follow_symlinks=False).st_ino != st.st_ino:
                            raise OSError("source changed under the store")
                    except OSError:
                        with open(tmp, "wb") as out, os.fdopen(os.dup(fd), "rb") as src:   # across filesystems: copy the hashed descriptor
                            shutil.copyfileobj(src, out, 8 << 20)
                    try:
                        os.link(tmp, str(target))                       # EEXIST = same content already stored
                    except FileExistsError:
                        pass
                    finally:
                        os.unlink(tmp)

                await asyncio.to_thread(place)
        finally:
            os.close(fd)
        return digest

    @asynccontextmanager
    async def reader(self, object_hash: str):
        _check_hash(object_hash)
        target = self.storage_path / object_hash
        if not object_hash or not await target.exists():
            raise FileNotFoundError(f"File not found: {object_hash}")
        async with await target.open("rb") as f:
            yield f

    async def read(self, object_hash: str) -> bytes:
        async with self.reader(object_hash) as f:
            return await f.read()

    async def exists(self, object_hash: str) -> bool:
        _check_hash(object_hash)
        return await (self.storage_path / object_hash).exists()


def _check_hash(object_hash: str) -> None:
    if not isinstance(object_hash, str) or not HASH_RE.match(object_hash):
        raise ValueError(f"invalid object id {object_hash!r}")  # pydantic ValidationError is a ValueError too
