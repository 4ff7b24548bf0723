// bci_rand_mt.cu — K8b: numpy's LEGACY generator on the device, bit-exact.
//
// `numpy.random.seed(s); numpy.random.rand(n)` (examples/using_imports.py:20-22 seeds the global
// generator) draws from MT19937: two tempered 32-bit outputs a, b per double,
// (a >> 5) * 2^26 + (b >> 6) scaled by 2^-53.  The stream is sequential by construction; what a
// GPU can do is run the 624-word state regeneration ("twist") of ONE stream wide, and B streams
// (requests) side by side, one CTA each.
//
// The textbook twist is three dependent sweeps (new[i] needs new[i-227]).  Substituting the
// recurrence into itself removes the dependency — every new word is an XOR of old words only:
//   i in [0,227):    new[i] = old[i+397] ^ f(old[i], old[i+1])
//   i in [227,454):  new[i] = old[i+170] ^ f(old[i-227], old[i-226]) ^ f(old[i], old[i+1])
//   i in [454,623):  new[i] = old[i-57]  ^ f(old[i-454], old[i-453]) ^ f(old[i-227], old[i-226]) ^ f(old[i], old[i+1])
//   new[623] = new[396] ^ f(old[623], new[0])   (both expanded the same way)
// so one regeneration is ONE barrier: 320 threads compute 624 words from buffer A into buffer B,
// then 312 of them temper a pair each and store one double (coalesced).  ~15 GB/s per stream —
// issue-bound on one SM, ~10x numpy on a host core — and B streams scale across the SMs.
//
// The caller passes the generator state (624 words + position, exactly numpy's get_state()) and
// gets the advanced state back, so host draws before and after stay on the same stream.
#include "bci_ops.cuh"
#include <vector>
#include <cstring>

namespace bci {

constexpr int kMtN = 624;
constexpr int kMtThreads = 640;   // one state word per thread: 20 warps hide the LDS -> XOR chain of a twist (320: 4.4 GB/s per stream)
constexpr int kMtStateWords = 625;  // key[624] + pos

struct MtReq {
  double* out;
  unsigned long long n;
};

__device__ __forceinline__ uint32_t mt_f(uint32_t a, uint32_t b) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return (y >> 1) ^ ((b & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
__device__ __forceinline__ uint32_t mt_twist_word(const uint32_t* __restrict__ o, int i) {
  if (i < 227) return o[i + 397] ^ mt_f(o[i], o[i + 1]);
  if (i < 454) return o[i + 170] ^ mt_f(o[i - 227], o[i - 226]) ^ mt_f(o[i], o[i + 1]);
  if (i < 623) return o[i - 57] ^ mt_f(o[i - 454], o[i - 453]) ^ mt_f(o[i - 227], o[i - 226]) ^ mt_f(o[i], o[i + 1]);
  const uint32_t n0 = o[397] ^ mt_f(o[0], o[1]);
  const uint32_t n396 = o[566] ^ mt_f(o[169], o[170]) ^ mt_f(o[396], o[397]);
  return n396 ^ mt_f(o[623], n0);
}

__global__ void __launch_bounds__(kMtThreads)
rand_mt19937_kernel(const MtReq* __restrict__ reqs, uint32_t* __restrict__ states) {
  __shared__ uint32_t buf[2][kMtN];
  const MtReq q = reqs[blockIdx.x];
  uint32_t* st = states + (size_t)blockIdx.x * kMtStateWords;
  const int t = threadIdx.x;
  for (int i = t; i < kMtN; i += kMtThreads) buf[0][i] = st[i];
  const long long pos = (long long)st[kMtN];  // 0..624: words of this block already consumed
  __syncthreads();
  const long long n = (long long)q.n;
  if (n == 0) return;
  const long long total_words = pos + 2 * n;          // flat index one past the last word used
  const long long b_last = (total_words - 1) / kMtN;  // block holding the last word used
  const int i2 = (pos & 1) ? 2 * t : 2 * t + 1;       // this thread's SECOND word of a pair
  int cur = 0;
  uint32_t carry = 0;  // word 623 of the previous block (thread 0, odd positions)
  for (long long b = 0;; ++b) {
    if (t < kMtN / 2) {
      const long long d = b * kMtN + i2 - pos - 1;  // 2 x output index when >= 0
      if (d >= 0 && (d >> 1) < n) {
        const uint32_t first = i2 > 0 ? buf[cur][i2 - 1] : carry;
        const uint32_t a = mt_temper(first) >> 5, c = mt_temper(buf[cur][i2]) >> 6;
        q.out[d >> 1] = ((double)a * 67108864.0 + (double)c) * (1.0 / 9007199254740992.0);
      }
    }
    if (b == b_last) break;
    if (t == 0) carry = buf[cur][kMtN - 1];
    for (int i = t; i < kMtN; i += kMtThreads) buf[cur ^ 1][i] = mt_twist_word(buf[cur], i);
    __syncthreads();
    cur ^= 1;
  }
  for (int i = t; i < kMtN; i += kMtThreads) st[i] = buf[cur][i];
  if (t == 0) st[kMtN] = (uint32_t)(total_words - b_last * kMtN);  // 1..624
}

}  // namespace bci

using namespace bci;

extern "C" int bci_rand_mt19937_f64(int B, void* const* out, const size_t* n, uint32_t* states) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !out || !n || !states) { set_error("bci_rand_mt19937_f64: bad arguments"); return BCI_ERR_INVALID; }
  std::vector<MtReq> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    if (n[i] && !out[i]) { set_error("bci_rand_mt19937_f64: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
    if (states[(size_t)i * kMtStateWords + kMtN] > (uint32_t)kMtN) {
      set_error("bci_rand_mt19937_f64: state position %u of request %d is out of range", states[(size_t)i * kMtStateWords + kMtN], i);
      return BCI_ERR_INVALID;
    }
    if (n[i] > (size_t)1 << 60) { set_error("bci_rand_mt19937_f64: request %d too large", i); return BCI_ERR_INVALID; }
    reqs[i].out = reinterpret_cast<double*>(out[i]);
    reqs[i].n = n[i];
  }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  void* d_reqs = nullptr;
  int rc = stage_upload(reqs.data(), sizeof(MtReq) * (size_t)B, &d_reqs);
  if (rc != BCI_OK) return rc;
  const size_t st_bytes = (size_t)B * kMtStateWords * sizeof(uint32_t);
  uint32_t* d_states = nullptr;
  cudaError_t e = cudaMallocFromPoolAsync((void**)&d_states, st_bytes, c.pool, c.stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocFromPoolAsync(mt19937 states)", __FILE__, __LINE__);
  e = cudaMemcpyAsync(d_states, states, st_bytes, cudaMemcpyHostToDevice, c.stream);
  if (e == cudaSuccess) {
    rand_mt19937_kernel<<<(unsigned)B, kMtThreads, 0, c.stream>>>(reinterpret_cast<const MtReq*>(d_reqs), d_states);
    c.launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(states, d_states, st_bytes, cudaMemcpyDeviceToHost, c.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
  cudaFreeAsync(d_states, c.stream);
  if (e != cudaSuccess) return cuda_fail(e, "bci_rand_mt19937_f64", __FILE__, __LINE__);
  return BCI_OK;
}
