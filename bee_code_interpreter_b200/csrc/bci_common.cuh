// bci_common.cuh — shared device/host plumbing of libbci.so (sm_100a only).
//
// Layout contract used by every batched kernel ("one launch per op class"):
//   * a launch covers B independent requests; request i has a descriptor `Req`
//     and owns the tile range [tile_start[i], tile_start[i+1]) of a flat 1-D grid;
//   * a CTA finds its request by binary search over tile_start (<= 11 steps for
//     B <= 2048, nothing for B == 1), so ragged batches cost no idle CTAs;
//   * all global traffic is 256-bit (LDG.E.256 / STG.E.256, sm_100a) on the aligned
//     path: 32 consecutive lanes touch 1 KiB contiguous per instruction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <mutex>
#include <thread>
#include <atomic>
#include "../../include/bci.h"

namespace bci {

// ---------------------------------------------------------------- host context
struct Ctx {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaMemPool_t pool = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // descriptor staging ring: pinned host mirror + device copy at the same offsets
  uint8_t* h_stage = nullptr;
  uint8_t* d_stage = nullptr;
  size_t stage_bytes = 0, stage_off = 0;
  // grow-only scratch (partials, histograms, tile flags)
  uint8_t* scratch = nullptr;
  size_t scratch_bytes = 0;
  // results coming back to the host: a few (pinned, device) slots, so that a call can drop the
  // context lock while it waits for ITS results and the next call can already launch
  struct ResultSlot {
    uint8_t* h = nullptr;
    uint8_t* d = nullptr;
    size_t bytes = 0;
    bool busy = false;
    cudaEvent_t done = nullptr;
  };
  static constexpr int kResultSlots = 32;
  ResultSlot rslot[kResultSlots];
  // L2 flush buffer
  uint8_t* flush = nullptr;
  size_t flush_bytes = 0;
  std::atomic<uint64_t> launches{0};
  std::mutex mu;
};
Ctx& ctx();
void set_error(const char* fmt, ...);
// A/B switches of the kernels (BCI_SCAN, BCI_SORT, BCI_MATMUL, BCI_GEMM_*, *_PREFETCH) are development tools:
// they are honoured only when the process was started with BCI_TUNING=1 (read once), so a stray variable in a
// production environment cannot select an untested path.  Returns nullptr otherwise.
const char* tuning_env(const char* name);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define BCI_CUDA(expr)                                                     \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) return ::bci::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define BCI_REQUIRE_INIT()                                                 \
  do {                                                                     \
    if (!::bci::ctx().ready) {                                             \
      ::bci::set_error("bci_init() has not been called (or failed)");      \
      return BCI_ERR_NOT_INIT;                                             \
    }                                                                      \
    /* the CUDA current device is per THREAD: broker/service threads other \
       than the one that called bci_init() start on device 0 */            \
    (void)cudaSetDevice(::bci::ctx().device);                              \
  } while (0)

#define BCI_CHECK_LAUNCH()                                                 \
  do {                                                                     \
    ::bci::ctx().launches.fetch_add(1, std::memory_order_relaxed);         \
    BCI_CUDA(cudaGetLastError());                                          \
  } while (0)

// Copies `bytes` of host data into the staging ring and schedules the H2D copy
// on the op stream; returns the DEVICE address (256-byte aligned).
int stage_upload(const void* host, size_t bytes, void** dptr);
// Host memory -> device through a pinned transfer lane; thread-safe, does not take the context lock for the copy.
int h2d_pinned(void* dst_dev, const void* src_host, size_t bytes);
int d2h_pinned(void* dst_host, const void* src_dev, size_t bytes);
// Grow-only device scratch; contents undefined. 256-byte aligned.
int scratch_get(size_t bytes, void** dptr);
// Result hand-back without holding the context lock while the GPU works:
//   { lock; result_acquire(); launch...; result_enqueue_copy(); }   // D2H + event on the op stream
//   result_wait();                                                     // no lock held
//   copy out of *hptr; result_release();
// result_acquire may drop and retake `lk` while it waits for a free slot.
int result_acquire(std::unique_lock<std::mutex>& lk, size_t bytes, int* slot, void** hptr, void** dptr);
int result_enqueue_copy(int slot, size_t bytes);
int result_wait(int slot);
void result_release(int slot);          // takes the context lock
void result_release_locked(int slot);   // caller holds it
struct ResultGuard {   // releases the slot on every exit path
  int slot = -1;
  bool lock_held = false;   // true when the guard dies inside the caller's locked region
  ~ResultGuard() {
    if (slot < 0) return;
    if (lock_held) result_release_locked(slot); else result_release(slot);
  }
};

// ---------------------------------------------------------------- descriptors
struct alignas(16) Req {
  const void* in0;
  const void* in1;  // nullptr -> use s0 as the second operand
  void* out;
  unsigned long long n;           // elements
  bci_scalar s0;                  // scalar operand
  bci_scalar s1;                  // addend of the fused forms / seed of rand
  unsigned long long tile_start;  // first tile of this request in the flat grid
  unsigned long long aux;         // op specific (partials offset, ...)
};
static_assert(sizeof(Req) == 64, "Req must stay 64 bytes");

static inline size_t elem_size(int dtype) {
  switch (dtype) {
    case BCI_F32: case BCI_I32: return 4;
    case BCI_F64: case BCI_I64: return 8;
    default: return 0;
  }
}

// ---------------------------------------------------------------- device side
#ifdef __CUDACC__

constexpr int kThreads = 256;

// 256-bit global access.  The pointer must be 32-byte aligned.
struct alignas(32) U256 { uint32_t r[8]; };

__device__ __forceinline__ U256 ldg256(const void* p) {
  U256 v;
  asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]),
                 "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7])
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256(void* p, const U256& v) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(v.r[0]), "r"(v.r[1]), "r"(v.r[2]), "r"(v.r[3]),
                  "r"(v.r[4]), "r"(v.r[5]), "r"(v.r[6]), "r"(v.r[7])
               : "memory");
}

template <typename T> struct VecOf {
  static constexpr int N = 32 / sizeof(T);
  union { U256 u; T v[N]; };
  __device__ __forceinline__ VecOf() {}
};

// Binary search: largest r with tile_start[r] <= tile.  tile_start has B+1 entries.
__device__ __forceinline__ int find_req(const unsigned long long* __restrict__ tile_start,
                                        int B, unsigned long long tile) {
  if (B == 1) return 0;
  int lo = 0, hi = B;  // invariant: tile_start[lo] <= tile < tile_start[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(tile_start + mid) <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

// Same result, found by the whole CTA in ONE memory round trip per 256 requests (each thread
// tests one boundary, the block counts the hits) instead of ~log2(B) dependent loads — for
// kernels whose tiles are short enough that the start-up chain shows (look-back kernels).
// Must be called by every thread of the CTA.
__device__ __forceinline__ int find_req_block(const unsigned long long* __restrict__ tile_start,
                                              int B, unsigned long long tile) {
  if (B == 1) return 0;
  int cnt = 0;  // boundaries tile_start[1..B-1] that are <= tile
  for (int i0 = 1; i0 < B; i0 += (int)blockDim.x) {
    const int i = i0 + (int)threadIdx.x;
    const bool le = i < B && __ldg(tile_start + i) <= tile;
    cnt += __syncthreads_count(le);
    if (cnt < i0 + (int)blockDim.x - 1) break;  // sorted boundaries: the rest are larger
  }
  return cnt;
}

__device__ __forceinline__ bool is_aligned32(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 31) == 0;
}

template <typename T> __device__ __forceinline__ T warp_shfl_down(T v, int d) {
  return __shfl_down_sync(0xffffffffu, v, d);
}
template <> __device__ __forceinline__ long long warp_shfl_down<long long>(long long v, int d) {
  return __shfl_down_sync(0xffffffffu, v, d);
}

#endif  // __CUDACC__

// Host helper: builds the Req table + tile_start table for B requests with a
// fixed tile size (in elements), uploads both, returns device pointers.
struct Batch {
  Req* d_reqs = nullptr;
  unsigned long long* d_tile_start = nullptr;
  unsigned long long total_tiles = 0;
};
int upload_batch(Req* host_reqs, int B, size_t tile_elems, Batch* out);

}  // namespace bci
