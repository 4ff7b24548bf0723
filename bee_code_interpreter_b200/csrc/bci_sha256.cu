// bci_sha256.cu — K4: SHA-256 (FIPS 180-4) of B messages per launch.
// Content addressing for services/storage.py:35-63 ("identified by their SHA-256 hash";
// the reference itself names objects with secrets.token_hex(32), storage.py:52).
//
// Not HBM-bound: one message is a strict chain of 64 rounds per 64-byte block, so the
// parallelism is the number of messages (SURVEY.md hard part 3).  Two kernels:
//  * warp-per-message (few, large files — BASELINE config C4): the 32 lanes take 32
//    consecutive blocks, each lane expands the message schedule W[0..63]+K of its own
//    block in parallel (phase 1, off the critical path, into shared memory), then the
//    warp runs the 64-round chain over the 32 blocks back to back (phase 2) reading
//    W+K as 128-bit broadcast LDS.  Phase 2 is ~14 ALU instructions per round and is
//    what bounds the kernel (INT-pipe issue + dependency latency);
//  * thread-per-message (many small files): 32 independent chains per warp, schedule
//    kept in a 16-word rolling register window.
// Padding (0x80, zeros, 64-bit big-endian bit length) is built in-kernel, so any byte
// length works; data is read with 128-bit loads when the pointer is 16-byte aligned.
#include "bci_common.cuh"
#include <thread>
#include <atomic>
#include <string>
#include <vector>
#include <condition_variable>
#include <cstring>

namespace bci {

__constant__ uint32_t kSha256K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static const uint32_t kSha256IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                      0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};

// One message (or one streaming chunk) of a launch.
struct alignas(16) ShaMsg {
  const uint8_t* data;        // device pointer
  unsigned long long len;     // bytes to absorb in this call
  unsigned long long prior;   // bytes absorbed before this call (multiple of 64)
  uint32_t h[8];              // chaining value going in
  uint32_t finalize;          // 1: append FIPS 180-4 padding for (prior + len) bytes
  uint32_t pad_;
};
static_assert(sizeof(ShaMsg) == 64, "ShaMsg must stay 64 bytes");

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ uint32_t sig0(uint32_t x) { return rotr(x, 7) ^ rotr(x, 18) ^ (x >> 3); }
__device__ __forceinline__ uint32_t sig1(uint32_t x) { return rotr(x, 17) ^ rotr(x, 19) ^ (x >> 10); }
__device__ __forceinline__ uint32_t Sig0(uint32_t x) { return rotr(x, 2) ^ rotr(x, 13) ^ rotr(x, 22); }
__device__ __forceinline__ uint32_t Sig1(uint32_t x) { return rotr(x, 6) ^ rotr(x, 11) ^ rotr(x, 25); }
__device__ __forceinline__ uint32_t Ch(uint32_t e, uint32_t f, uint32_t g) { return (e & f) ^ (~e & g); }
__device__ __forceinline__ uint32_t Maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) ^ (a & c) ^ (b & c); }

// number of 64-byte blocks this call compresses
__device__ __host__ __forceinline__ unsigned long long sha_blocks(unsigned long long len, bool finalize) {
  return finalize ? (len + 9 + 63) / 64 : len / 64;
}

// Loads block `b` of the call as 16 big-endian words, building the padding when the
// block reaches past the data.
__device__ __forceinline__ void load_block(const ShaMsg& m, unsigned long long b, bool aligned16,
                                           uint32_t w[16]) {
  const unsigned long long off = b * 64;
  if (off + 64 <= m.len && aligned16) {
    const uint4* p = reinterpret_cast<const uint4*>(m.data + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 v;
      asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + j));
      w[4 * j + 0] = bswap(v.x); w[4 * j + 1] = bswap(v.y);
      w[4 * j + 2] = bswap(v.z); w[4 * j + 3] = bswap(v.w);
    }
    return;
  }
  const unsigned long long bits = (m.prior + m.len) * 8ull;
  const unsigned long long nblk = sha_blocks(m.len, m.finalize != 0);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint32_t word = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned long long pos = off + 4 * i + k;
      uint32_t byte = 0;
      if (pos < m.len) byte = m.data[pos];
      else if (pos == m.len && m.finalize) byte = 0x80;
      word = (word << 8) | byte;
    }
    if (m.finalize && b == nblk - 1) {
      if (i == 14) word = (uint32_t)(bits >> 32);
      if (i == 15) word = (uint32_t)bits;
    }
    w[i] = word;
  }
}

#define SHA_ROUND(a, b, c, d, e, f, g, h, wk)                 \
  {                                                           \
    const uint32_t t1 = h + Sig1(e) + Ch(e, f, g) + (wk);     \
    const uint32_t t2 = Sig0(a) + Maj(a, b, c);               \
    d += t1;                                                  \
    h = t1 + t2;                                              \
  }

// Same round with every addition written as x * one + y, `one` being a kernel argument equal to
// 1 that ptxas cannot fold (it turns a literal `mad.lo x, 1, y` back into IADD3).  One warp per SM
// sub-partition makes the round chain ALU-pipe bound (6 SHF + 4 LOP3 + 2 IADD3, two issue cycles
// each, ncu: ALU pipe 79 %); IMAD issues on the FMA pipe, so the additions move there and 10 ALU
// instructions per round are left.
// The dependent chain per round is e -> Sig1(e) (SHF, LOP3) -> ONE three-input add -> new e:
// everything else that goes into the new e (d + h + W+K) and into the new a (Sig0 + Maj - d) is
// known a round earlier and is summed off the chain on the FMA pipe.
#define SHA_ROUND_FMA(a, b, c, d, e, f, g, h, wk, one)                        \
  {                                                                           \
    const uint32_t p = d * (one##_b) + (h * (one##_c) + (wk));               \
    const uint32_t r = d * (one##_m) + (Sig0(a) * (one) + Maj(a, b, c));     \
    const uint32_t en = Sig1(e) + Ch(e, f, g) + p;                            \
    h = en * (one) + r;                                                       \
    d = en;                                                                   \
  }

// ---- kernel A: warp per message ----------------------------------------------------
constexpr int kShaWarps = 4;          // warps (= messages) per CTA
constexpr int kShaRow = 68;           // words per block row in smem: 64 + 4 pad (keeps LDS.128 aligned)

__global__ void __launch_bounds__(kShaWarps * 32)
sha256_warp_kernel(const ShaMsg* __restrict__ msgs, int B, uint32_t* __restrict__ out_state, uint32_t one,
                   uint32_t one_b, uint32_t one_c, uint32_t one_m) {  // 1, 1, 1, -1: distinct arguments so the sums cannot be factored
  __shared__ __align__(16) uint32_t swk[kShaWarps][32 * kShaRow];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mi = blockIdx.x * kShaWarps + warp;
  if (mi >= B) return;
  const ShaMsg m = msgs[mi];
  const bool aligned16 = (reinterpret_cast<uintptr_t>(m.data) & 15) == 0;
  const unsigned long long nblk = sha_blocks(m.len, m.finalize != 0);
  uint32_t* wk = swk[warp];

  uint32_t s0 = m.h[0], s1 = m.h[1], s2 = m.h[2], s3 = m.h[3];
  uint32_t s4 = m.h[4], s5 = m.h[5], s6 = m.h[6], s7 = m.h[7];

  uint32_t w[16];
  bool have = false;
  if ((unsigned long long)lane < nblk) { load_block(m, lane, aligned16, w); have = true; }

#pragma unroll 1
  for (unsigned long long c0 = 0; c0 < nblk; c0 += 32) {
    // ---- phase 1: lane L expands the schedule of block c0+L into smem (W[t]+K[t])
    if (have) {
      uint32_t* row = wk + lane * kShaRow;
#pragma unroll
      for (int t = 0; t < 16; t += 4) {
        *reinterpret_cast<uint4*>(row + t) =
            make_uint4(w[t] + kSha256K[t], w[t + 1] + kSha256K[t + 1], w[t + 2] + kSha256K[t + 2],
                       w[t + 3] + kSha256K[t + 3]);
      }
#pragma unroll
      for (int t = 16; t < 64; t += 4) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = (t + k) & 15;
          const uint32_t v = w[i] + sig0(w[(i + 1) & 15]) + w[(i + 9) & 15] + sig1(w[(i + 14) & 15]);
          w[i] = v;
          o[k] = v + kSha256K[t + k];
        }
        *reinterpret_cast<uint4*>(row + t) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncwarp();
    // ---- prefetch this lane's block of the next chunk while the rounds run
    const unsigned long long nb = c0 + 32 + lane;
    have = nb < nblk;
    if (have) load_block(m, nb, aligned16, w);
    // ---- phase 2: the round chain over up to 32 blocks, all lanes in lock step
    const int cnt = (nblk - c0) < 32 ? (int)(nblk - c0) : 32;
#pragma unroll 1
    for (int bi = 0; bi < cnt; ++bi) {
      const uint4* row = reinterpret_cast<const uint4*>(wk + bi * kShaRow);
      uint32_t a = s0, b = s1, c = s2, d = s3, e = s4, f = s5, g = s6, h = s7;
#pragma unroll
      for (int t = 0; t < 64; t += 8) {
        const uint4 x = row[t / 4], y = row[t / 4 + 1];
        SHA_ROUND_FMA(a, b, c, d, e, f, g, h, x.x, one)
        SHA_ROUND_FMA(h, a, b, c, d, e, f, g, x.y, one)
        SHA_ROUND_FMA(g, h, a, b, c, d, e, f, x.z, one)
        SHA_ROUND_FMA(f, g, h, a, b, c, d, e, x.w, one)
        SHA_ROUND_FMA(e, f, g, h, a, b, c, d, y.x, one)
        SHA_ROUND_FMA(d, e, f, g, h, a, b, c, y.y, one)
        SHA_ROUND_FMA(c, d, e, f, g, h, a, b, y.z, one)
        SHA_ROUND_FMA(b, c, d, e, f, g, h, a, y.w, one)
      }
      s0 += a; s1 += b; s2 += c; s3 += d; s4 += e; s5 += f; s6 += g; s7 += h;
    }
    __syncwarp();
  }
  if (lane == 0) {
    uint32_t* o = out_state + (size_t)mi * 8;
    o[0] = s0; o[1] = s1; o[2] = s2; o[3] = s3; o[4] = s4; o[5] = s5; o[6] = s6; o[7] = s7;
  }
}

// ---- kernel B: thread per message ----------------------------------------------------
__global__ void __launch_bounds__(128)
sha256_thread_kernel(const ShaMsg* __restrict__ msgs, int B, uint32_t* __restrict__ out_state) {
  const int mi = blockIdx.x * blockDim.x + threadIdx.x;
  if (mi >= B) return;
  const ShaMsg m = msgs[mi];
  const bool aligned16 = (reinterpret_cast<uintptr_t>(m.data) & 15) == 0;
  const unsigned long long nblk = sha_blocks(m.len, m.finalize != 0);
  uint32_t s0 = m.h[0], s1 = m.h[1], s2 = m.h[2], s3 = m.h[3];
  uint32_t s4 = m.h[4], s5 = m.h[5], s6 = m.h[6], s7 = m.h[7];
#pragma unroll 1
  for (unsigned long long bi = 0; bi < nblk; ++bi) {
    uint32_t w[16];
    load_block(m, bi, aligned16, w);
    uint32_t a = s0, b = s1, c = s2, d = s3, e = s4, f = s5, g = s6, h = s7;
#pragma unroll
    for (int t = 0; t < 64; t += 8) {
      if (t >= 16) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int i = (t + k) & 15;
          w[i] = w[i] + sig0(w[(i + 1) & 15]) + w[(i + 9) & 15] + sig1(w[(i + 14) & 15]);
        }
      }
      SHA_ROUND(a, b, c, d, e, f, g, h, w[(t + 0) & 15] + kSha256K[t + 0])
      SHA_ROUND(h, a, b, c, d, e, f, g, w[(t + 1) & 15] + kSha256K[t + 1])
      SHA_ROUND(g, h, a, b, c, d, e, f, w[(t + 2) & 15] + kSha256K[t + 2])
      SHA_ROUND(f, g, h, a, b, c, d, e, w[(t + 3) & 15] + kSha256K[t + 3])
      SHA_ROUND(e, f, g, h, a, b, c, d, w[(t + 4) & 15] + kSha256K[t + 4])
      SHA_ROUND(d, e, f, g, h, a, b, c, w[(t + 5) & 15] + kSha256K[t + 5])
      SHA_ROUND(c, d, e, f, g, h, a, b, w[(t + 6) & 15] + kSha256K[t + 6])
      SHA_ROUND(b, c, d, e, f, g, h, a, w[(t + 7) & 15] + kSha256K[t + 7])
    }
    s0 += a; s1 += b; s2 += c; s3 += d; s4 += e; s5 += f; s6 += g; s7 += h;
  }
  uint32_t* o = out_state + (size_t)mi * 8;
  o[0] = s0; o[1] = s1; o[2] = s2; o[3] = s3; o[4] = s4; o[5] = s5; o[6] = s6; o[7] = s7;
}

// Side streams for hashing.  One SHA-256 is a serial chain: a launch over a few large messages runs for as long as
// its longest message (64 MiB: ~1 s) while using a handful of warps.  On the op stream it would stall every
// other request's launches for that long, so hashing launches go to one of a few side streams (ordered after
// the op stream by an event) and the library lock is released while the digests are awaited; launches of
// different callers overlap on the GPU.
namespace {
constexpr int kShaStreams = 64;
struct ShaLane {
  cudaStream_t stream = nullptr;
  cudaEvent_t order = nullptr, done = nullptr;
  uint32_t* h_out = nullptr;   // pinned
  uint32_t* d_out = nullptr;
  ShaMsg* d_msgs = nullptr;
  ShaMsg* h_msgs = nullptr;    // pinned
  size_t cap = 0;              // messages
  bool busy = false;
};
ShaLane g_sha_lanes[kShaStreams];
std::condition_variable g_sha_cv;

// called with the context lock held; may wait (lock released) for a free lane
int sha_lane_acquire(std::unique_lock<std::mutex>& lk, size_t B, ShaLane** out) {
  for (;;) {
    for (auto& l : g_sha_lanes) {
      if (l.busy) continue;
      if (!l.stream) {
        BCI_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        BCI_CUDA(cudaEventCreateWithFlags(&l.order, cudaEventDisableTiming));
        BCI_CUDA(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
      }
      if (B > l.cap) {
        if (l.h_out) cudaFreeHost(l.h_out);
        if (l.h_msgs) cudaFreeHost(l.h_msgs);
        if (l.d_out) cudaFree(l.d_out);
        if (l.d_msgs) cudaFree(l.d_msgs);
        l.h_out = nullptr; l.h_msgs = nullptr; l.d_out = nullptr; l.d_msgs = nullptr; l.cap = 0;
        const size_t cap = B * 2 < 64 ? 64 : B * 2;
        BCI_CUDA(cudaHostAlloc((void**)&l.h_out, cap * 32, cudaHostAllocDefault));
        BCI_CUDA(cudaHostAlloc((void**)&l.h_msgs, cap * sizeof(ShaMsg), cudaHostAllocDefault));
        BCI_CUDA(cudaMalloc((void**)&l.d_out, cap * 32));
        BCI_CUDA(cudaMalloc((void**)&l.d_msgs, cap * sizeof(ShaMsg)));
        l.cap = cap;
      }
      l.busy = true;
      *out = &l;
      return BCI_OK;
    }
    g_sha_cv.wait(lk);
  }
}
}  // namespace

// Runs one launch over `B` prepared messages; states come back to `h_states` (B x 8 words).
// Called with the context lock held (`lk`): the launch is enqueued under it, the wait for the digests is not.
static int sha_launch(std::unique_lock<std::mutex>& lk, std::vector<ShaMsg>& msgs, uint32_t* h_states) {
  Ctx& c = ctx();
  const int B = (int)msgs.size();
  ShaLane* lane = nullptr;
  int rc = sha_lane_acquire(lk, (size_t)B, &lane);
  if (rc != BCI_OK) return rc;
  struct Release {
    ShaLane* l; std::unique_lock<std::mutex>& lk;
    ~Release() { if (!lk.owns_lock()) lk.lock(); l->busy = false; g_sha_cv.notify_one(); }
  } release{lane, lk};
  memcpy(lane->h_msgs, msgs.data(), sizeof(ShaMsg) * (size_t)B);
  // everything the messages refer to (uploads, producers) is already enqueued on the op stream
  BCI_CUDA(cudaEventRecord(lane->order, c.stream));
  BCI_CUDA(cudaStreamWaitEvent(lane->stream, lane->order, 0));
  BCI_CUDA(cudaMemcpyAsync(lane->d_msgs, lane->h_msgs, sizeof(ShaMsg) * (size_t)B, cudaMemcpyHostToDevice, lane->stream));
  // many small messages: thread-per-message fills the lanes; few large: warp-per-message
  const int warp_slots = c.sm_count * 4 * 2;
  if (B > warp_slots) {
    const int threads = 128;
    sha256_thread_kernel<<<(B + threads - 1) / threads, threads, 0, lane->stream>>>(lane->d_msgs, B, lane->d_out);
  } else {
    sha256_warp_kernel<<<(B + kShaWarps - 1) / kShaWarps, kShaWarps * 32, 0, lane->stream>>>(
        lane->d_msgs, B, lane->d_out, 1u, 1u, 1u, 0xffffffffu);
  }
  BCI_CHECK_LAUNCH();
  BCI_CUDA(cudaMemcpyAsync(lane->h_out, lane->d_out, (size_t)B * 32, cudaMemcpyDeviceToHost, lane->stream));
  BCI_CUDA(cudaEventRecord(lane->done, lane->stream));
  lk.unlock();                                   // other requests launch while this hash runs
  cudaError_t e = cudaEventSynchronize(lane->done);
  lk.lock();
  if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize(sha)", __FILE__, __LINE__);
  memcpy(h_states, lane->h_out, (size_t)B * 32);
  return BCI_OK;
}

static void state_to_digest(const uint32_t* st, uint8_t* out) {
  for (int i = 0; i < 8; ++i) {
    out[4 * i + 0] = (uint8_t)(st[i] >> 24);
    out[4 * i + 1] = (uint8_t)(st[i] >> 16);
    out[4 * i + 2] = (uint8_t)(st[i] >> 8);
    out[4 * i + 3] = (uint8_t)(st[i]);
  }
}

}  // namespace bci

using namespace bci;

extern "C" int bci_sha256(int B, const void* const* data, const size_t* len, int on_device,
                          uint8_t* digests) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !data || !len || !digests) { set_error("bci_sha256: bad arguments"); return BCI_ERR_INVALID; }
  for (int i = 0; i < B; ++i)
    if (len[i] && !data[i]) { set_error("bci_sha256: NULL data in message %d", i); return BCI_ERR_INVALID; }
  Ctx& c = ctx();
  std::vector<ShaMsg> msgs((size_t)B);
  uint8_t* staging = nullptr;
  std::vector<size_t> offs((size_t)B, 0);
  if (!on_device) {
    // host messages: one device staging block, filled through the pinned transfer lanes by up to four copy
    // threads in parallel — the library lock is NOT held while the bytes cross PCIe
    size_t total = 0;
    for (int i = 0; i < B; ++i) { offs[i] = total; total += (len[i] + 255) & ~size_t(255); }
    if (total) {
      {
        std::lock_guard<std::mutex> lk(c.mu);
        cudaError_t e = cudaMallocFromPoolAsync((void**)&staging, total, c.pool, c.stream);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMallocFromPoolAsync(sha staging)", __FILE__, __LINE__);
      }
      const int workers = B < 4 ? B : 4;
      std::vector<int> rcs((size_t)workers, BCI_OK);
      std::vector<std::string> errs((size_t)workers);
      std::atomic<int> next{0};
      auto work = [&](int w) {
        (void)cudaSetDevice(c.device);
        for (int i = next.fetch_add(1); i < B; i = next.fetch_add(1)) {
          int rc = h2d_pinned(staging + offs[i], data[i], len[i]);
          if (rc != BCI_OK) { rcs[w] = rc; errs[w] = bci_last_error(); return; }
        }
      };
      std::vector<std::thread> th;
      for (int w = 1; w < workers; ++w) th.emplace_back(work, w);
      work(0);
      for (auto& t : th) t.join();
      for (int w = 0; w < workers; ++w)
        if (rcs[w] != BCI_OK) {
          std::lock_guard<std::mutex> lk(c.mu);
          cudaFreeAsync(staging, c.stream);
          set_error("%s", errs[w].c_str());
          return rcs[w];
        }
    }
  }
  for (int i = 0; i < B; ++i) {
    ShaMsg& m = msgs[i];
    memset(&m, 0, sizeof m);
    m.data = on_device ? reinterpret_cast<const uint8_t*>(data[i]) : staging + offs[i];
    m.len = len[i];
    m.prior = 0;
    memcpy(m.h, kSha256IV, sizeof m.h);
    m.finalize = 1;
  }
  std::vector<uint32_t> states((size_t)B * 8);
  std::unique_lock<std::mutex> lk(c.mu);
  int rc = sha_launch(lk, msgs, states.data());
  if (staging) cudaFreeAsync(staging, c.stream);
  if (rc != BCI_OK) return rc;
  for (int i = 0; i < B; ++i) state_to_digest(&states[(size_t)i * 8], digests + (size_t)i * 32);
  return BCI_OK;
}

extern "C" int bci_sha256_init(bci_sha256_state* st) {
  if (!st) { set_error("bci_sha256_init: NULL state"); return BCI_ERR_INVALID; }
  memcpy(st->h, kSha256IV, sizeof st->h);
  st->total_len = 0;
  st->buf_len = 0;
  memset(st->buf, 0, sizeof st->buf);
  return BCI_OK;
}

// Absorbs one host chunk per writer.  Whole 64-byte blocks go to the GPU in ONE launch
// over all B writers; the < 64-byte remainder waits in the state's buffer.
static int sha_absorb(int B, bci_sha256_state* const* st, const void* const* host_data,
                      const size_t* len, bool finalize, uint8_t* digests) {
  Ctx& c = ctx();
  std::unique_lock<std::mutex> lk(c.mu);
  // lay out each writer's bytes (buffered head + new data) contiguously in device staging
  std::vector<size_t> offs((size_t)B), take((size_t)B);
  size_t total = 0;
  for (int i = 0; i < B; ++i) {
    const size_t l = (len && host_data && host_data[i]) ? len[i] : 0;
    const size_t avail = st[i]->buf_len + l;
    take[i] = finalize ? avail : (avail / 64) * 64;
    offs[i] = total;
    total += (take[i] + 255) & ~size_t(255);
  }
  uint8_t* staging = nullptr;
  std::vector<uint8_t> heads;
  if (total) {
    cudaError_t e = cudaMallocFromPoolAsync((void**)&staging, total, c.pool, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocFromPoolAsync(sha staging)", __FILE__, __LINE__);
  }
  std::vector<ShaMsg> msgs;
  std::vector<int> who;
  for (int i = 0; i < B; ++i) {
    bci_sha256_state* s = st[i];
    const size_t l = (len && host_data && host_data[i]) ? len[i] : 0;
    const uint8_t* src = l ? reinterpret_cast<const uint8_t*>(host_data[i]) : nullptr;
    if (take[i] == 0 && !finalize) {
      // not even one block yet: keep buffering
      if (l) { memcpy(s->buf + s->buf_len, src, l); s->buf_len += (uint32_t)l; }
      continue;
    }
    const size_t from_buf = s->buf_len < take[i] ? s->buf_len : take[i];
    const size_t from_new = take[i] - from_buf;
    if (from_buf) {
      // buf lives in caller memory that may change after we return: stage via the ring
      void* dptr = nullptr;
      int rc = stage_upload(s->buf, from_buf, &dptr);
      if (rc != BCI_OK) { cudaFreeAsync(staging, c.stream); return rc; }
      BCI_CUDA(cudaMemcpyAsync(staging + offs[i], dptr, from_buf, cudaMemcpyDeviceToDevice, c.stream));
    }
    if (from_new) {
      cudaError_t e = cudaMemcpyAsync(staging + offs[i] + from_buf, src, from_new, cudaMemcpyHostToDevice, c.stream);
      if (e != cudaSuccess) { cudaFreeAsync(staging, c.stream); return cuda_fail(e, "cudaMemcpyAsync(sha h2d)", __FILE__, __LINE__); }
    }
    ShaMsg m;
    memset(&m, 0, sizeof m);
    m.data = staging + offs[i];
    m.len = take[i];
    m.prior = s->total_len;
    memcpy(m.h, s->h, sizeof m.h);
    m.finalize = finalize ? 1u : 0u;
    msgs.push_back(m);
    who.push_back(i);
    // remainder of the new data waits for the next call
    const size_t left = l - from_new;
    s->total_len += take[i];
    s->buf_len = 0;
    if (left) { memcpy(s->buf, src + from_new, left); s->buf_len = (uint32_t)left; }
  }
  int rc = BCI_OK;
  if (!msgs.empty()) {
    std::vector<uint32_t> states(msgs.size() * 8);
    rc = sha_launch(lk, msgs, states.data());
    if (rc == BCI_OK) {
      for (size_t k = 0; k < msgs.size(); ++k) memcpy(st[who[k]]->h, &states[k * 8], 32);
    }
  }
  if (staging) cudaFreeAsync(staging, c.stream);
  if (rc != BCI_OK) return rc;
  if (finalize && digests) {
    for (int i = 0; i < B; ++i) state_to_digest(st[i]->h, digests + (size_t)i * 32);
  }
  return BCI_OK;
}

extern "C" int bci_sha256_update_batch(int B, bci_sha256_state* const* st, const void* const* host_data,
                                       const size_t* len) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !st || !host_data || !len) { set_error("bci_sha256_update_batch: bad arguments"); return BCI_ERR_INVALID; }
  for (int i = 0; i < B; ++i) if (!st[i]) { set_error("bci_sha256_update_batch: NULL state %d", i); return BCI_ERR_INVALID; }
  return sha_absorb(B, st, host_data, len, false, nullptr);
}

extern "C" int bci_sha256_final_batch(int B, bci_sha256_state* const* st, uint8_t* digests) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !st || !digests) { set_error("bci_sha256_final_batch: bad arguments"); return BCI_ERR_INVALID; }
  for (int i = 0; i < B; ++i) if (!st[i]) { set_error("bci_sha256_final_batch: NULL state %d", i); return BCI_ERR_INVALID; }
  return sha_absorb(B, st, nullptr, nullptr, true, digests);
}
