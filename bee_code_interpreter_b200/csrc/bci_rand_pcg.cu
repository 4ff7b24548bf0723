// bci_rand_pcg.cu — K8d: numpy.random.Generator(PCG64) streams on the device, bit-exact.
//
// BASELINE configs C2/C3/C5 create their inputs with `numpy.random.default_rng(seed)`
// (SURVEY.md §8d): `.random(n, dtype=float32|float64)`, `.integers(0, 2**31, n)`; C4 with
// `.bytes(n)`.  numpy's arithmetic is the behavioural spec (numpy/random/src/pcg64/pcg64.h,
// numpy/random/src/distributions/distributions.c):
//   state' = state * 0x2360ED051FC65DA44385DF649FCCF645 + inc   (mod 2^128)
//   next64 = rotr64(hi ^ lo, state' >> 122)                      (XSL-RR of the NEW state)
//   next32 = low half of a next64, then its high half (has_uint32 buffering)
//   random f64 = (next64 >> 11) * 2^-53;  random f32 = (next32 >> 8) * 2^-24
//   integers over a range < 2^32: Lemire's multiply-shift on next32, rejecting when the low
//   product word is below (2^32 - range) % range.
//
// An LCG is sequential, but its k-step jump is affine: s_{i+k} = P_k s_i + inc * G_k with
// P_k = A^k and G_k = 1 + A + ... + A^(k-1), both independent of the request.  A table of
// (P_k, G_k) built once on the host gives every thread its own starting state in <= 5 jumps:
//   4 table levels over the tile index (256-ary digits) + 1 entry for the thread's offset;
// after that a thread emits 8 consecutive draws (64 output bytes, two STG.256) and hops to its
// next group with ONE more multiply (P_stride, and inc*G_stride kept in registers).
// Bound: 128-bit multiplies on the fma pipe (about 14 IMAD per 8 output bytes), not HBM.
// Draws a rejection would have consumed cannot be generated in parallel: the kernel counts
// them per request and the caller falls back to numpy when the count is not zero (never for
// power-of-two ranges such as C5's 2^31).
#include "bci_common.cuh"
#include <vector>
#include <cstring>

namespace bci {

typedef unsigned __int128 u128;

struct alignas(16) PcgJump { u128 P, G; };

constexpr int kPcgDrawsPerGroup = 8;
constexpr int kPcgGroups = 8;
constexpr int kPcgTileDraws = kThreads * kPcgDrawsPerGroup * kPcgGroups;   // 16384 draws = 128 KiB
constexpr int kPcgGroupStride = kThreads * kPcgDrawsPerGroup;              // 2048 draws
constexpr int kPcgLevels = 4;                                              // 256^4 tiles
constexpr int kPcgTabOffset = 0;                                           // [0,256): thread offsets
constexpr int kPcgTabLevel = 256;                                          // [256, 256+4*256): tile digits
constexpr int kPcgTabStride = kPcgTabLevel + kPcgLevels * 256;             // group hop
constexpr int kPcgTabSize = kPcgTabStride + 1;

struct alignas(16) PcgReq {
  u128 state, inc;
  void* out;
  unsigned long long n;           // elements
  unsigned long long tile_start;
  long long off;                  // bounded kinds: lowest value
  uint32_t rng_excl;              // bounded kinds: range size (high - low), 0 = 2^32
  uint32_t threshold;             // Lemire rejection threshold
  uint32_t pad[2];
};
static_assert(sizeof(PcgReq) == 80, "PcgReq layout");

#define PCG_MULT ((((u128)2549297995355413924ULL) << 64) | (u128)4865540595714422341ULL)

__device__ __forceinline__ uint64_t pcg_output(u128 s) {
  const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
  const uint64_t x = hi ^ lo;
  const unsigned rot = (unsigned)(hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}

template <int KIND> struct PcgOut;
template <> struct PcgOut<BCI_PCG_F64> { using T = double; static constexpr int PER_DRAW = 1; };
template <> struct PcgOut<BCI_PCG_F32> { using T = float; static constexpr int PER_DRAW = 2; };
template <> struct PcgOut<BCI_PCG_U32> { using T = uint32_t; static constexpr int PER_DRAW = 2; };
template <> struct PcgOut<BCI_PCG_BOUNDED32_I64> { using T = long long; static constexpr int PER_DRAW = 2; };
template <> struct PcgOut<BCI_PCG_BOUNDED32_I32> { using T = int; static constexpr int PER_DRAW = 2; };

// one 32-bit word of the stream -> one output element (kinds with two elements per draw)
template <int KIND>
__device__ __forceinline__ typename PcgOut<KIND>::T pcg_elem32(uint32_t x, const PcgReq& q, bool live, unsigned& rejected) {
  if constexpr (KIND == BCI_PCG_F32) {
    return (float)(x >> 8) * 5.9604644775390625e-08f;          // (next32 >> 8) * 2^-24, exact
  } else if constexpr (KIND == BCI_PCG_U32) {
    return x;
  } else {
    // buffered_bounded_lemire_uint32: m = x * rng_excl; reject while (uint32)m < threshold
    const unsigned long long m = (unsigned long long)x * (unsigned long long)q.rng_excl;
    if (live && (uint32_t)m < q.threshold) ++rejected;   // words past n are draws numpy never made
    const long long v = q.off + (long long)(m >> 32);
    return (typename PcgOut<KIND>::T)v;
  }
}

__device__ __forceinline__ double pcg_f64(uint64_t w) {
  // (w >> 11) * 2^-53 without a 64-bit I2F: the 53-bit integer is hi21 * 2^32 + lo32
  const double hi = (double)(uint32_t)(w >> 43);
  const double lo = (double)(uint32_t)(w >> 11);
  return __fma_rn(hi, 4.76837158203125e-07 /* 2^-21 */, __dmul_rn(lo, 1.1102230246251565404e-16 /* 2^-53 */));
}

template <int KIND>
__global__ void __launch_bounds__(kThreads)
pcg64_kernel(const PcgReq* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
             const PcgJump* __restrict__ tab, unsigned int* __restrict__ rejected_out) {
  using T = typename PcgOut<KIND>::T;
  constexpr int EPD = PcgOut<KIND>::PER_DRAW;
  constexpr int EPG = kPcgDrawsPerGroup * EPD;            // elements per thread per group
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const PcgReq q = reqs[r];
  const unsigned long long t = tile - q.tile_start;

  // this thread's first state: s_{t * tile + tid * 8} from s_0
  u128 s = q.state;
  const u128 inc = q.inc;
#pragma unroll
  for (int lvl = 0; lvl < kPcgLevels; ++lvl) {
    const unsigned d = (unsigned)(t >> (8 * lvl)) & 255u;
    if (d) {                                              // CTA-uniform branch
      const PcgJump j = tab[kPcgTabLevel + lvl * 256 + d];
      s = j.P * s + inc * j.G;
    }
  }
  {
    const PcgJump j = tab[kPcgTabOffset + threadIdx.x];
    s = j.P * s + inc * j.G;
  }
  const PcgJump hop = tab[kPcgTabStride];
  const u128 hop_c = inc * hop.G;

  T* __restrict__ out = reinterpret_cast<T*>(q.out);
  const unsigned long long n = q.n;
  unsigned rejected = 0;
#pragma unroll 1
  for (int g = 0; g < kPcgGroups; ++g) {
    const unsigned long long e0 =
        (t * (unsigned long long)kPcgTileDraws + (unsigned long long)g * kPcgGroupStride + threadIdx.x * kPcgDrawsPerGroup) * EPD;
    if (e0 >= n) break;                                   // later groups lie further out still
    T v[EPG];
#pragma unroll
    for (int k = 0; k < kPcgDrawsPerGroup; ++k) {
      s = s * PCG_MULT + inc;
      const uint64_t w = pcg_output(s);
      if constexpr (KIND == BCI_PCG_F64) {
        v[k] = pcg_f64(w);
      } else {
        v[2 * k] = pcg_elem32<KIND>((uint32_t)w, q, e0 + 2 * k < n, rejected);
        v[2 * k + 1] = pcg_elem32<KIND>((uint32_t)(w >> 32), q, e0 + 2 * k + 1 < n, rejected);
      }
    }
    T* o = out + e0;
    if (e0 + EPG <= n && is_aligned32(o)) {
      constexpr int NV = (int)(EPG * sizeof(T) / 32);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        U256 u;
        memcpy(&u, reinterpret_cast<const char*>(v) + 32 * j, 32);
        stg256(reinterpret_cast<char*>(o) + 32 * j, u);
      }
    } else {
#pragma unroll
      for (int k = 0; k < EPG; ++k) {
        if (e0 + k < n) o[k] = v[k];
      }
    }
    s = hop.P * s + hop_c;
  }
  if constexpr (KIND == BCI_PCG_BOUNDED32_I64 || KIND == BCI_PCG_BOUNDED32_I32) {
    if (rejected) atomicAdd(&rejected_out[r], rejected);
  }
}

// ---- host: jump tables ------------------------------------------------------------------------
static PcgJump jump_of(unsigned long long k) {
  // (P, G) of k steps by binary composition: (Pa,Ga) then (Pb,Gb) = (Pb*Pa, Pb*Ga + Gb)
  u128 P = 1, G = 0;              // identity
  u128 bp = PCG_MULT, bg = 1;     // one step
  while (k) {
    if (k & 1) { G = bp * G + bg; P = bp * P; }
    bg = bp * bg + bg;            // two of the current: (bp*bp, bp*bg + bg)
    bp = bp * bp;
    k >>= 1;
  }
  return PcgJump{P, G};
}

static PcgJump* g_pcg_tab = nullptr;

static int pcg_tables(const PcgJump** out) {
  if (!g_pcg_tab) {
    std::vector<PcgJump> h((size_t)kPcgTabSize);
    for (int t = 0; t < 256; ++t) h[kPcgTabOffset + t] = jump_of((unsigned long long)t * kPcgDrawsPerGroup);
    for (int lvl = 0; lvl < kPcgLevels; ++lvl) {
      // d * tile * 256^lvl steps: repeated composition of the d = 1 entry
      const PcgJump one = jump_of((unsigned long long)kPcgTileDraws << (8 * lvl));
      PcgJump acc{1, 0};
      for (int d = 0; d < 256; ++d) {
        h[kPcgTabLevel + lvl * 256 + d] = acc;
        acc = PcgJump{one.P * acc.P, one.P * acc.G + one.G};
      }
    }
    h[kPcgTabStride] = jump_of((unsigned long long)(kPcgGroupStride - kPcgDrawsPerGroup));
    PcgJump* d = nullptr;
    BCI_CUDA(cudaMalloc((void**)&d, h.size() * sizeof(PcgJump)));
    BCI_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(PcgJump), cudaMemcpyHostToDevice));
    g_pcg_tab = d;
  }
  *out = g_pcg_tab;
  return BCI_OK;
}

template <int KIND>
static int launch_pcg(const PcgReq* d_reqs, const unsigned long long* d_ts, int B, unsigned tiles,
                      const PcgJump* tab, unsigned int* rejected) {
  pcg64_kernel<KIND><<<tiles, kThreads, 0, ctx().stream>>>(d_reqs, d_ts, B, tab, rejected);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

}  // namespace bci

using namespace bci;

extern "C" int bci_rand_pcg64(int kind, int B, void* const* out, const size_t* n, const bci_pcg64_gen* gen,
                              uint32_t* rejected) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !out || !n || !gen) { set_error("bci_rand_pcg64: bad arguments"); return BCI_ERR_INVALID; }
  const bool bounded = kind == BCI_PCG_BOUNDED32_I64 || kind == BCI_PCG_BOUNDED32_I32;
  if (kind < BCI_PCG_F64 || kind > BCI_PCG_BOUNDED32_I32) { set_error("bci_rand_pcg64: unknown kind %d", kind); return BCI_ERR_INVALID; }
  if (bounded && !rejected) { set_error("bci_rand_pcg64: bounded kinds need the `rejected` array"); return BCI_ERR_INVALID; }
  const size_t per_draw = kind == BCI_PCG_F64 ? 1 : 2;
  const size_t tile_elems = (size_t)kPcgTileDraws * per_draw;
  const size_t req_bytes = sizeof(PcgReq) * (size_t)B, ts_bytes = 8 * (size_t)(B + 1);
  std::vector<uint8_t> buf(req_bytes + ts_bytes);
  PcgReq* reqs = reinterpret_cast<PcgReq*>(buf.data());
  unsigned long long* ts = reinterpret_cast<unsigned long long*>(buf.data() + req_bytes);
  unsigned long long t = 0;
  for (int i = 0; i < B; ++i) {
    PcgReq& q = reqs[i];
    memset(&q, 0, sizeof q);
    if (n[i] && !out[i]) { set_error("bci_rand_pcg64: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
    q.state = ((u128)gen[i].state_hi << 64) | gen[i].state_lo;
    q.inc = ((u128)gen[i].inc_hi << 64) | gen[i].inc_lo;
    q.out = out[i];
    q.n = n[i];
    q.off = gen[i].low;
    if (bounded) {
      // range = high - low (exclusive); the 32-bit Lemire path takes 1 < range < 2^32
      const uint64_t range = gen[i].range;
      if (range < 2 || range > 0xFFFFFFFFull) {
        set_error("bci_rand_pcg64: request %d has range %llu (need 2 <= range < 2^32)", i, (unsigned long long)range);
        return BCI_ERR_UNSUPPORTED;
      }
      q.rng_excl = (uint32_t)range;
      q.threshold = (uint32_t)((0xFFFFFFFFull - (range - 1)) % range);
    }
    q.tile_start = t;
    ts[i] = t;
    t += (n[i] + tile_elems - 1) / tile_elems;
  }
  ts[B] = t;
  if (t > 0x7fffffffull) { set_error("bci_rand_pcg64: batch too large for one launch"); return BCI_ERR_INVALID; }
  Ctx& c = ctx();
  void *h = nullptr, *d = nullptr;
  ResultGuard guard;
  {
    std::unique_lock<std::mutex> lk(c.mu);
    const PcgJump* tab = nullptr;
    int rc = pcg_tables(&tab);
    if (rc != BCI_OK) return rc;
    if (bounded) {
      rc = result_acquire(lk, (size_t)B * 4, &guard.slot, &h, &d);
      if (rc != BCI_OK) return rc;
      BCI_CUDA(cudaMemsetAsync(d, 0, (size_t)B * 4, c.stream));
    }
    if (t) {
      void* dd = nullptr;
      rc = stage_upload(buf.data(), buf.size(), &dd);
      if (rc != BCI_OK) return rc;
      const PcgReq* d_reqs = reinterpret_cast<const PcgReq*>(dd);
      const unsigned long long* d_ts = reinterpret_cast<const unsigned long long*>(reinterpret_cast<uint8_t*>(dd) + req_bytes);
      unsigned int* rej = reinterpret_cast<unsigned int*>(d);
      switch (kind) {
        case BCI_PCG_F64: rc = launch_pcg<BCI_PCG_F64>(d_reqs, d_ts, B, (unsigned)t, tab, rej); break;
        case BCI_PCG_F32: rc = launch_pcg<BCI_PCG_F32>(d_reqs, d_ts, B, (unsigned)t, tab, rej); break;
        case BCI_PCG_U32: rc = launch_pcg<BCI_PCG_U32>(d_reqs, d_ts, B, (unsigned)t, tab, rej); break;
        case BCI_PCG_BOUNDED32_I64: rc = launch_pcg<BCI_PCG_BOUNDED32_I64>(d_reqs, d_ts, B, (unsigned)t, tab, rej); break;
        case BCI_PCG_BOUNDED32_I32: rc = launch_pcg<BCI_PCG_BOUNDED32_I32>(d_reqs, d_ts, B, (unsigned)t, tab, rej); break;
      }
      if (rc != BCI_OK) return rc;
    }
    if (bounded && (rc = result_enqueue_copy(guard.slot, (size_t)B * 4)) != BCI_OK) return rc;
  }
  if (bounded) {
    int rc = result_wait(guard.slot);
    if (rc != BCI_OK) return rc;
    memcpy(rejected, h, (size_t)B * 4);
  } else if (rejected) {
    memset(rejected, 0, (size_t)B * 4);
  }
  return BCI_OK;
}
