// bci_runtime.cu — process-wide device context of libbci.so: one process = one GPU.
// Takes over what kubernetes_code_executor.py:151-264 (warm pod queue) provides in
// the reference: a ready execution context that requests are dispatched into.
#include "bci_common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <condition_variable>
#include <cerrno>
#include <unistd.h>

namespace bci {

static thread_local std::string g_err;

Ctx& ctx() {
  static Ctx c;
  return c;
}

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

static std::atomic<int> g_poisoned{0};

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
  switch (e) {   // errors that leave the context unusable: every later call would fail the same way
    case cudaErrorIllegalAddress: case cudaErrorLaunchFailure: case cudaErrorHardwareStackError:
    case cudaErrorIllegalInstruction: case cudaErrorMisalignedAddress: case cudaErrorInvalidAddressSpace:
    case cudaErrorInvalidPc: case cudaErrorECCUncorrectable: case cudaErrorLaunchTimeout:
    case cudaErrorAssert: case cudaErrorCudartUnloading: case cudaErrorContextIsDestroyed:
      g_poisoned.store((int)e);
      break;
    default: break;
  }
  cudaGetLastError();  // clear the sticky-less error state
  return e == cudaErrorMemoryAllocation ? BCI_ERR_OOM : BCI_ERR_CUDA;
}

static constexpr size_t kStageBytes = 8u << 20;  // descriptor ring

const char* tuning_env(const char* name) {
  static const bool enabled = [] { const char* t = getenv("BCI_TUNING"); return t && t[0] == '1'; }();
  return enabled ? getenv(name) : nullptr;
}

int stage_upload(const void* host, size_t bytes, void** dptr) {
  Ctx& c = ctx();
  size_t need = (bytes + 255) & ~size_t(255);
  if (need > c.stage_bytes) {
    // oversize descriptor table: grow the ring (rare; B > ~100k)
    BCI_CUDA(cudaStreamSynchronize(c.stream));
    if (c.h_stage) cudaFreeHost(c.h_stage);
    if (c.d_stage) cudaFree(c.d_stage);
    c.h_stage = nullptr; c.d_stage = nullptr;
    size_t nb = need * 2;
    BCI_CUDA(cudaHostAlloc((void**)&c.h_stage, nb, cudaHostAllocDefault));
    BCI_CUDA(cudaMalloc((void**)&c.d_stage, nb));
    c.stage_bytes = nb;
    c.stage_off = 0;
  }
  if (c.stage_off + need > c.stage_bytes) {
    // wrap: every earlier copy out of the pinned mirror must have executed
    BCI_CUDA(cudaStreamSynchronize(c.stream));
    c.stage_off = 0;
  }
  memcpy(c.h_stage + c.stage_off, host, bytes);
  BCI_CUDA(cudaMemcpyAsync(c.d_stage + c.stage_off, c.h_stage + c.stage_off, bytes,
                           cudaMemcpyHostToDevice, c.stream));
  *dptr = c.d_stage + c.stage_off;
  c.stage_off += need;
  return BCI_OK;
}

int scratch_get(size_t bytes, void** dptr) {
  Ctx& c = ctx();
  if (bytes > c.scratch_bytes) {
    BCI_CUDA(cudaStreamSynchronize(c.stream));
    if (c.scratch) cudaFree(c.scratch);
    c.scratch = nullptr; c.scratch_bytes = 0;
    size_t nb = (bytes + (bytes >> 2) + 4095) & ~size_t(4095);
    BCI_CUDA(cudaMalloc((void**)&c.scratch, nb));
    c.scratch_bytes = nb;
  }
  *dptr = c.scratch;
  return BCI_OK;
}

int result_acquire(std::unique_lock<std::mutex>& lk, size_t bytes, int* slot, void** hptr, void** dptr) {
  Ctx& c = ctx();
  int s = -1;
  for (;;) {
    for (int i = 0; i < Ctx::kResultSlots; ++i)
      if (!c.rslot[i].busy) { s = i; break; }
    if (s >= 0) break;
    lk.unlock();               // every slot is in flight: let their owners finish
    std::this_thread::yield();
    lk.lock();
  }
  Ctx::ResultSlot& r = c.rslot[s];
  if (bytes > r.bytes) {
    // the slot is free, so nothing on the stream refers to its buffers any more
    if (r.h) cudaFreeHost(r.h);
    if (r.d) cudaFree(r.d);
    r.h = nullptr; r.d = nullptr; r.bytes = 0;
    size_t nb = (bytes * 2 + 4095) & ~size_t(4095);
    BCI_CUDA(cudaHostAlloc((void**)&r.h, nb, cudaHostAllocDefault));
    BCI_CUDA(cudaMalloc((void**)&r.d, nb));
    r.bytes = nb;
  }
  if (!r.done) BCI_CUDA(cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming));
  r.busy = true;
  *slot = s;
  *hptr = r.h;
  *dptr = r.d;
  return BCI_OK;
}

int result_enqueue_copy(int slot, size_t bytes) {
  Ctx& c = ctx();
  Ctx::ResultSlot& r = c.rslot[slot];
  BCI_CUDA(cudaMemcpyAsync(r.h, r.d, bytes, cudaMemcpyDeviceToHost, c.stream));
  BCI_CUDA(cudaEventRecord(r.done, c.stream));
  return BCI_OK;
}

int result_wait(int slot) {
  BCI_CUDA(cudaEventSynchronize(ctx().rslot[slot].done));
  return BCI_OK;
}

void result_release(int slot) {
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  c.rslot[slot].busy = false;
}

void result_release_locked(int slot) { ctx().rslot[slot].busy = false; }

int upload_batch(Req* host_reqs, int B, size_t tile_elems, Batch* out) {
  // one upload: [Req x B][tile_start x (B+1)]
  size_t req_bytes = sizeof(Req) * (size_t)B;
  size_t ts_bytes = sizeof(unsigned long long) * (size_t)(B + 1);
  std::vector<uint8_t> buf(req_bytes + ts_bytes);
  unsigned long long* ts = reinterpret_cast<unsigned long long*>(buf.data() + req_bytes);
  unsigned long long t = 0;
  for (int i = 0; i < B; ++i) {
    host_reqs[i].tile_start = t;
    ts[i] = t;
    t += (host_reqs[i].n + tile_elems - 1) / tile_elems;
  }
  ts[B] = t;
  memcpy(buf.data(), host_reqs, req_bytes);
  void* d = nullptr;
  int rc = stage_upload(buf.data(), buf.size(), &d);
  if (rc != BCI_OK) return rc;
  out->d_reqs = reinterpret_cast<Req*>(d);
  out->d_tile_start = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(d) + req_bytes);
  out->total_tiles = t;
  return BCI_OK;
}

}  // namespace bci

using namespace bci;

extern "C" {

int bci_abi_version(void) { return BCI_ABI_VERSION; }

const char* bci_last_error(void) { return g_err.c_str(); }

int bci_healthy(void) { return g_poisoned.load() == 0 ? 1 : 0; }

int bci_device_count(int* count) {
  if (!count) { set_error("count is NULL"); return BCI_ERR_INVALID; }
  *count = 0;
  BCI_CUDA(cudaGetDeviceCount(count));
  return BCI_OK;
}

int bci_init(int device) {
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  if (c.ready) {
    if (c.device == device) return BCI_OK;
    set_error("bci_init: already bound to device %d (one process drives one GPU)", c.device);
    return BCI_ERR_INVALID;
  }
  int count = 0;
  BCI_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) {
    set_error("bci_init: device %d out of range (%d visible)", device, count);
    return BCI_ERR_INVALID;
  }
  BCI_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BCI_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("bci_init: device %d is sm_%d%d; libbci.so carries sm_100a code only", device,
              prop.major, prop.minor);
    return BCI_ERR_UNSUPPORTED;
  }
  c.device = device;
  c.sm_count = prop.multiProcessorCount;
  BCI_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  c.own_stream = true;
  BCI_CUDA(cudaDeviceGetDefaultMemPool(&c.pool, device));
  uint64_t thresh = UINT64_MAX;  // keep freed blocks: the next wave of requests reuses them
  BCI_CUDA(cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  BCI_CUDA(cudaEventCreate(&c.ev0));
  BCI_CUDA(cudaEventCreate(&c.ev1));
  BCI_CUDA(cudaHostAlloc((void**)&c.h_stage, kStageBytes, cudaHostAllocDefault));
  BCI_CUDA(cudaMalloc((void**)&c.d_stage, kStageBytes));
  c.stage_bytes = kStageBytes;
  c.stage_off = 0;
  c.launches = 0;
  c.ready = true;
  return BCI_OK;
}

void bci_shutdown(void) {
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  if (!c.ready) return;
  cudaStreamSynchronize(c.stream);
  if (c.pool) cudaMemPoolTrimTo(c.pool, 0);  // give cached blocks back to the driver
  if (c.h_stage) cudaFreeHost(c.h_stage);
  if (c.d_stage) cudaFree(c.d_stage);
  if (c.scratch) cudaFree(c.scratch);
  for (auto& r : c.rslot) {
    if (r.h) cudaFreeHost(r.h);
    if (r.d) cudaFree(r.d);
    if (r.done) cudaEventDestroy(r.done);
    r = Ctx::ResultSlot();
  }
  if (c.flush) cudaFree(c.flush);
  c.h_stage = c.d_stage = c.scratch = c.flush = nullptr;
  c.stage_bytes = c.scratch_bytes = c.flush_bytes = 0;
  cudaEventDestroy(c.ev0);
  cudaEventDestroy(c.ev1);
  if (c.own_stream) cudaStreamDestroy(c.stream);
  c.stream = nullptr;
  c.ready = false;
  c.device = -1;
}

int bci_device_info(bci_devinfo* info) {
  BCI_REQUIRE_INIT();
  if (!info) { set_error("info is NULL"); return BCI_ERR_INVALID; }
  Ctx& c = ctx();
  cudaDeviceProp prop;
  BCI_CUDA(cudaGetDeviceProperties(&prop, c.device));
  info->device = c.device;
  info->sm_count = prop.multiProcessorCount;
  info->cc_major = prop.major;
  info->cc_minor = prop.minor;
  info->total_mem = prop.totalGlobalMem;
  snprintf(info->name, sizeof info->name, "%s", prop.name);
  return BCI_OK;
}

int bci_mem_info(size_t* free_bytes, size_t* total_bytes) {
  BCI_REQUIRE_INIT();
  size_t f = 0, t = 0;
  BCI_CUDA(cudaMemGetInfo(&f, &t));
  // memory parked in the pool is reusable by bci_alloc: count it as free
  Ctx& c = ctx();
  uint64_t reserved = 0, used = 0;
  cudaMemPoolGetAttribute(c.pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
  cudaMemPoolGetAttribute(c.pool, cudaMemPoolAttrUsedMemCurrent, &used);
  if (free_bytes) *free_bytes = f + (size_t)(reserved - used);
  if (total_bytes) *total_bytes = t;
  return BCI_OK;
}

int bci_alloc(size_t bytes, void** dptr) {
  BCI_REQUIRE_INIT();
  if (!dptr) { set_error("dptr is NULL"); return BCI_ERR_INVALID; }
  *dptr = nullptr;
  if (bytes == 0) bytes = 256;  // numpy allows empty arrays; keep a unique non-NULL handle
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  cudaError_t e = cudaMallocFromPoolAsync(dptr, bytes, c.pool, c.stream);
  if (e != cudaSuccess) {
    *dptr = nullptr;
    return cuda_fail(e, "cudaMallocFromPoolAsync", __FILE__, __LINE__);
  }
  return BCI_OK;
}

int bci_free(void* dptr) {
  BCI_REQUIRE_INIT();
  if (!dptr) return BCI_OK;
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  BCI_CUDA(cudaFreeAsync(dptr, c.stream));
  return BCI_OK;
}

int bci_host_alloc(size_t bytes, void** hptr) {
  BCI_REQUIRE_INIT();
  if (!hptr) { set_error("hptr is NULL"); return BCI_ERR_INVALID; }
  BCI_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return BCI_OK;
}

int bci_host_free(void* hptr) {
  BCI_REQUIRE_INIT();
  if (!hptr) return BCI_OK;
  BCI_CUDA(cudaFreeHost(hptr));
  return BCI_OK;
}

int bci_h2d(void* dst_dev, const void* src_host, size_t bytes) {
  BCI_REQUIRE_INIT();
  if (bytes == 0) return BCI_OK;
  if (!dst_dev || !src_host) { set_error("bci_h2d: NULL pointer"); return BCI_ERR_INVALID; }
  if (bytes > (256u << 10)) {
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, src_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    // a large pageable source would be staged by the runtime WHILE the library lock is held: move it through a
    // transfer lane instead (returns when the copy is done); pinned sources keep the asynchronous path
    if (!pinned) return bci::h2d_pinned(dst_dev, src_host, bytes);
  }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  BCI_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, c.stream));
  return BCI_OK;
}

int bci_d2h(void* dst_host, const void* src_dev, size_t bytes) {
  BCI_REQUIRE_INIT();
  if (bytes == 0) return BCI_OK;
  if (!dst_host || !src_dev) { set_error("bci_d2h: NULL pointer"); return BCI_ERR_INVALID; }
  // Never wait for the GPU with the library lock held: a script printing `s[0], s[-1]` after a sort would otherwise
  // stall every other request's launches for as long as the op stream is busy.  Small reads (an element, a scalar
  // result) go through a pinned result slot and wait on its event; bulk reads through a transfer lane.
  if (bytes <= (64u << 10)) {
    Ctx& c = ctx();
    void *h = nullptr, *d = nullptr;
    ResultGuard guard;
    {
      std::unique_lock<std::mutex> lk(c.mu);
      int rc = result_acquire(lk, bytes, &guard.slot, &h, &d);
      if (rc != BCI_OK) return rc;
      Ctx::ResultSlot& r = c.rslot[guard.slot];
      BCI_CUDA(cudaMemcpyAsync(r.h, src_dev, bytes, cudaMemcpyDeviceToHost, c.stream));
      BCI_CUDA(cudaEventRecord(r.done, c.stream));
    }
    int rc = result_wait(guard.slot);
    if (rc != BCI_OK) return rc;
    memcpy(dst_host, h, bytes);
    return BCI_OK;
  }
  return bci::d2h_pinned(dst_host, src_dev, bytes);
}

int bci_d2d(void* dst_dev, const void* src_dev, size_t bytes) {
  BCI_REQUIRE_INIT();
  if (bytes == 0) return BCI_OK;
  if (!dst_dev || !src_dev) { set_error("bci_d2d: NULL pointer"); return BCI_ERR_INVALID; }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  BCI_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, c.stream));
  return BCI_OK;
}

int bci_memset(void* dst_dev, int byte, size_t bytes) {
  BCI_REQUIRE_INIT();
  if (bytes == 0) return BCI_OK;
  if (!dst_dev) { set_error("bci_memset: NULL pointer"); return BCI_ERR_INVALID; }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  BCI_CUDA(cudaMemsetAsync(dst_dev, byte, bytes, c.stream));
  return BCI_OK;
}

// ---- pinned file <-> device transfers ---------------------------------------------------------
namespace {
constexpr size_t kLaneChunk = 8u << 20;
constexpr int kLanes = 8;
struct Lane {
  cudaStream_t stream = nullptr;
  uint8_t* buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaEvent_t order = nullptr;
  bool busy = false;
};
Lane g_lanes[kLanes];
std::mutex g_lane_mu;
std::condition_variable g_lane_cv;

int lane_acquire(Lane** out) {
  std::unique_lock<std::mutex> lk(g_lane_mu);
  for (;;) {
    for (auto& l : g_lanes) {
      if (l.busy) continue;
      if (!l.stream) {
        BCI_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
          BCI_CUDA(cudaHostAlloc((void**)&l.buf[i], kLaneChunk, cudaHostAllocDefault));
          BCI_CUDA(cudaEventCreateWithFlags(&l.done[i], cudaEventDisableTiming));
        }
        BCI_CUDA(cudaEventCreateWithFlags(&l.order, cudaEventDisableTiming));
      }
      l.busy = true;
      *out = &l;
      return BCI_OK;
    }
    g_lane_cv.wait(lk);
  }
}
void lane_release(Lane* l) {
  { std::lock_guard<std::mutex> lk(g_lane_mu); l->busy = false; }
  g_lane_cv.notify_one();
}
struct LaneGuard { Lane* l = nullptr; ~LaneGuard() { if (l) { cudaStreamSynchronize(l->stream); lane_release(l); } } };

// the lane's stream starts after everything already enqueued on the op stream
int lane_order_after_ops(Lane* l) {
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  BCI_CUDA(cudaEventRecord(l->order, c.stream));
  BCI_CUDA(cudaStreamWaitEvent(l->stream, l->order, 0));
  return BCI_OK;
}
}  // namespace

}  // extern "C"

namespace bci {
// Host memory -> device through a transfer lane (pinned bounce buffers, copy stream): what cudaMemcpyAsync
// from pageable memory does internally, minus holding the library lock and the op stream while it happens.
// Several callers (threads) run in parallel on different lanes.
int h2d_pinned(void* dst_dev, const void* src_host, size_t bytes) {
  if (bytes == 0) return BCI_OK;
  LaneGuard g;
  int rc = lane_acquire(&g.l);
  if (rc != BCI_OK) return rc;
  if ((rc = lane_order_after_ops(g.l)) != BCI_OK) return rc;
  size_t done = 0;
  for (int k = 0; done < bytes; ++k) {
    const int b = k & 1;
    if (k >= 2) BCI_CUDA(cudaEventSynchronize(g.l->done[b]));
    const size_t want = bytes - done < kLaneChunk ? bytes - done : kLaneChunk;
    memcpy(g.l->buf[b], static_cast<const uint8_t*>(src_host) + done, want);
    BCI_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(dst_dev) + done, g.l->buf[b], want, cudaMemcpyHostToDevice, g.l->stream));
    BCI_CUDA(cudaEventRecord(g.l->done[b], g.l->stream));
    done += want;
  }
  BCI_CUDA(cudaStreamSynchronize(g.l->stream));
  return BCI_OK;
}

// Device -> host memory through a transfer lane (the mirror of h2d_pinned).
int d2h_pinned(void* dst_host, const void* src_dev, size_t bytes) {
  if (bytes == 0) return BCI_OK;
  LaneGuard g;
  int rc = lane_acquire(&g.l);
  if (rc != BCI_OK) return rc;
  if ((rc = lane_order_after_ops(g.l)) != BCI_OK) return rc;
  const size_t chunks = (bytes + kLaneChunk - 1) / kLaneChunk;
  auto issue = [&](size_t k) -> int {
    const size_t off = k * kLaneChunk, want = bytes - off < kLaneChunk ? bytes - off : kLaneChunk;
    BCI_CUDA(cudaMemcpyAsync(g.l->buf[k & 1], static_cast<const uint8_t*>(src_dev) + off, want, cudaMemcpyDeviceToHost, g.l->stream));
    BCI_CUDA(cudaEventRecord(g.l->done[k & 1], g.l->stream));
    return BCI_OK;
  };
  if ((rc = issue(0)) != BCI_OK) return rc;
  for (size_t k = 0; k < chunks; ++k) {
    if (k + 1 < chunks && (rc = issue(k + 1)) != BCI_OK) return rc;
    BCI_CUDA(cudaEventSynchronize(g.l->done[k & 1]));
    const size_t off = k * kLaneChunk, want = bytes - off < kLaneChunk ? bytes - off : kLaneChunk;
    memcpy(static_cast<uint8_t*>(dst_host) + off, g.l->buf[k & 1], want);
  }
  return BCI_OK;
}
}  // namespace bci

extern "C" {

int bci_h2d_fd(void* dst_dev, int fd, uint64_t file_offset, size_t bytes) {
  BCI_REQUIRE_INIT();
  if (bytes == 0) return BCI_OK;
  if (!dst_dev || fd < 0) { set_error("bci_h2d_fd: bad arguments"); return BCI_ERR_INVALID; }
  LaneGuard g;
  int rc = lane_acquire(&g.l);
  if (rc != BCI_OK) return rc;
  if ((rc = lane_order_after_ops(g.l)) != BCI_OK) return rc;
  size_t done = 0;
  for (int k = 0; done < bytes; ++k) {
    const int b = k & 1;
    if (k >= 2) BCI_CUDA(cudaEventSynchronize(g.l->done[b]));      // the DMA out of this bounce buffer has finished
    const size_t want = bytes - done < kLaneChunk ? bytes - done : kLaneChunk;
    size_t got = 0;
    while (got < want) {
      ssize_t r = ::pread(fd, g.l->buf[b] + got, want - got, (off_t)(file_offset + done + got));
      if (r < 0 && errno == EINTR) continue;
      if (r <= 0) { set_error("bci_h2d_fd: file ends %zu bytes short (%s)", bytes - done - got, r < 0 ? strerror(errno) : "EOF"); return BCI_ERR_INVALID; }
      got += (size_t)r;
    }
    BCI_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(dst_dev) + done, g.l->buf[b], want, cudaMemcpyHostToDevice, g.l->stream));
    BCI_CUDA(cudaEventRecord(g.l->done[b], g.l->stream));
    done += want;
  }
  BCI_CUDA(cudaStreamSynchronize(g.l->stream));
  return BCI_OK;
}

int bci_d2h_fd(int fd, uint64_t file_offset, const void* src_dev, size_t bytes) {
  BCI_REQUIRE_INIT();
  if (bytes == 0) return BCI_OK;
  if (!src_dev || fd < 0) { set_error("bci_d2h_fd: bad arguments"); return BCI_ERR_INVALID; }
  LaneGuard g;
  int rc = lane_acquire(&g.l);
  if (rc != BCI_OK) return rc;
  if ((rc = lane_order_after_ops(g.l)) != BCI_OK) return rc;
  const size_t chunks = (bytes + kLaneChunk - 1) / kLaneChunk;
  auto issue = [&](size_t k) -> int {
    const size_t off = k * kLaneChunk, want = bytes - off < kLaneChunk ? bytes - off : kLaneChunk;
    BCI_CUDA(cudaMemcpyAsync(g.l->buf[k & 1], static_cast<const uint8_t*>(src_dev) + off, want, cudaMemcpyDeviceToHost, g.l->stream));
    BCI_CUDA(cudaEventRecord(g.l->done[k & 1], g.l->stream));
    return BCI_OK;
  };
  if ((rc = issue(0)) != BCI_OK) return rc;
  for (size_t k = 0; k < chunks; ++k) {
    if (k + 1 < chunks && (rc = issue(k + 1)) != BCI_OK) return rc;   // DMA of the next chunk runs while this one is written
    BCI_CUDA(cudaEventSynchronize(g.l->done[k & 1]));
    const size_t off = k * kLaneChunk, want = bytes - off < kLaneChunk ? bytes - off : kLaneChunk;
    size_t put = 0;
    while (put < want) {
      ssize_t r = ::pwrite(fd, g.l->buf[k & 1] + put, want - put, (off_t)(file_offset + off + put));
      if (r < 0 && errno == EINTR) continue;
      if (r <= 0) { set_error("bci_d2h_fd: write failed: %s", strerror(errno)); return BCI_ERR_INVALID; }
      put += (size_t)r;
    }
  }
  return BCI_OK;
}

int bci_sync(void) {
  BCI_REQUIRE_INIT();
  Ctx& c = ctx();
  BCI_CUDA(cudaStreamSynchronize(c.stream));
  return BCI_OK;
}

int bci_timer_begin(void) {
  BCI_REQUIRE_INIT();
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  BCI_CUDA(cudaEventRecord(c.ev0, c.stream));
  return BCI_OK;
}

int bci_timer_end(float* elapsed_ms) {
  BCI_REQUIRE_INIT();
  if (!elapsed_ms) { set_error("elapsed_ms is NULL"); return BCI_ERR_INVALID; }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  BCI_CUDA(cudaEventRecord(c.ev1, c.stream));
  BCI_CUDA(cudaEventSynchronize(c.ev1));
  BCI_CUDA(cudaEventElapsedTime(elapsed_ms, c.ev0, c.ev1));
  return BCI_OK;
}

static cudaEvent_t g_slots[64];
static bool g_slot_made[64];

int bci_event_record(int slot) {
  BCI_REQUIRE_INIT();
  if (slot < 0 || slot >= 64) { set_error("bci_event_record: slot %d out of range", slot); return BCI_ERR_INVALID; }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  if (!g_slot_made[slot]) {
    BCI_CUDA(cudaEventCreate(&g_slots[slot]));
    g_slot_made[slot] = true;
  }
  BCI_CUDA(cudaEventRecord(g_slots[slot], c.stream));
  return BCI_OK;
}

int bci_event_elapsed(int slot_begin, int slot_end, float* elapsed_ms) {
  BCI_REQUIRE_INIT();
  if (slot_begin < 0 || slot_begin >= 64 || slot_end < 0 || slot_end >= 64 || !elapsed_ms ||
      !g_slot_made[slot_begin] || !g_slot_made[slot_end]) {
    set_error("bci_event_elapsed: bad slots");
    return BCI_ERR_INVALID;
  }
  BCI_CUDA(cudaEventSynchronize(g_slots[slot_end]));
  BCI_CUDA(cudaEventElapsedTime(elapsed_ms, g_slots[slot_begin], g_slots[slot_end]));
  return BCI_OK;
}

uint64_t bci_launch_count(void) { return ctx().launches.load(); }

int bci_flush_l2(void) {
  BCI_REQUIRE_INIT();
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  const size_t want = 256u << 20;  // 2x the 126 MB L2
  if (!c.flush) {
    BCI_CUDA(cudaMalloc((void**)&c.flush, want));
    c.flush_bytes = want;
  }
  BCI_CUDA(cudaMemsetAsync(c.flush, 0x5a, c.flush_bytes, c.stream));
  return BCI_OK;
}

}  // extern "C"
