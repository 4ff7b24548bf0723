// bci_reduce.cu — K2 full reductions (sum/mean/max/min, optional fused producer)
// and K6 argmax/argmin.
// Work the reference does inside numpy's pairwise-sum / min-max reduce loops when
// agent code runs `numpy.sum(numpy.square(large_array))` (examples/benchmark-numpy.py:21).
//
// HBM-bound: algorithmic bytes per request = n*sizeof(T) per input read + 8 written.
// One CTA owns a 128 KiB tile (4 rounds of 4 x 256-bit loads per thread, VN private
// accumulators per thread), reduces it with warp shuffles + one smem hop, writes ONE
// partial; the last CTA of each request (device counter) folds that request's
// partials in tile order, so the result is deterministic and independent of how
// the batch was packed.  No float atomics.
#include "bci_ops.cuh"
#include "bci_philox.cuh"
#include <vector>
#include <cstring>
#include <cmath>
#include <limits>

namespace bci {

constexpr int kRedUnroll = 4;
constexpr int kRedRounds = 4;
template <typename T> constexpr size_t red_tile_elems() {
  return (size_t)kThreads * kRedUnroll * kRedRounds * (32 / sizeof(T));
}

// ---- accumulator policies -------------------------------------------------------
template <typename T> struct SumAcc { using type = T; };
template <> struct SumAcc<int> { using type = long long; };  // numpy.sum(int32) -> int64

template <typename T, int OP> struct Red;
template <typename T> struct Red<T, BCI_RED_SUM> {
  using Acc = typename SumAcc<T>::type;
  static __device__ __forceinline__ Acc init() { return Acc(0); }
  static __device__ __forceinline__ Acc lift(T x) { return static_cast<Acc>(x); }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return fadd_(a, b); }
  static __device__ __forceinline__ void accum(Acc& a, Acc e, bool&) { a = fadd_(a, e); }
};
template <typename T> struct Lowest;
template <> struct Lowest<float> { static __device__ __forceinline__ float lo() { return -INFINITY; } static __device__ __forceinline__ float hi() { return INFINITY; } };
template <> struct Lowest<double> { static __device__ __forceinline__ double lo() { return -INFINITY; } static __device__ __forceinline__ double hi() { return INFINITY; } };
template <> struct Lowest<int> { static __device__ __forceinline__ int lo() { return INT_MIN; } static __device__ __forceinline__ int hi() { return INT_MAX; } };
template <> struct Lowest<long long> { static __device__ __forceinline__ long long lo() { return LLONG_MIN; } static __device__ __forceinline__ long long hi() { return LLONG_MAX; } };
template <typename T> struct Red<T, BCI_RED_MAX> {
  using Acc = T;
  static __device__ __forceinline__ Acc init() { return Lowest<T>::lo(); }
  static __device__ __forceinline__ Acc lift(T x) { return x; }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return nan_max(a, b); }
  // hot loop: FMNMX (ignores NaN) + a NaN flag folded in once per thread — 2 instructions per
  // element where the NaN-propagating select chain needs 5
  static __device__ __forceinline__ void accum(Acc& a, Acc e, bool& bad) {
    if constexpr (is_float_t<T>::value) { bad |= (e != e); a = fmax(a, e); } else { a = a > e ? a : e; }
  }
};
template <typename T> struct Red<T, BCI_RED_MIN> {
  using Acc = T;
  static __device__ __forceinline__ Acc init() { return Lowest<T>::hi(); }
  static __device__ __forceinline__ Acc lift(T x) { return x; }
  static __device__ __forceinline__ Acc combine(Acc a, Acc b) { return nan_min(a, b); }
  static __device__ __forceinline__ void accum(Acc& a, Acc e, bool& bad) {
    if constexpr (is_float_t<T>::value) { bad |= (e != e); a = fmin(a, e); } else { a = a < e ? a : e; }
  }
};

template <typename P> __device__ __forceinline__ typename P::Acc block_reduce(typename P::Acc v) {
  using Acc = typename P::Acc;
  __shared__ Acc warp_part[kThreads / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = P::combine(v, __shfl_down_sync(0xffffffffu, v, d));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // protect warp_part reuse between consecutive calls
  if (lane == 0) warp_part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < kThreads / 32 ? warp_part[lane] : P::init();
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) v = P::combine(v, __shfl_down_sync(0xffffffffu, v, d));
  }
  return v;  // valid in thread 0
}

// flags in Req.aux
constexpr unsigned long long kAuxFused = 1;      // in0 (op) in1|s0 producer
constexpr unsigned long long kAuxAddend = 2;     // + s1 after the producer
constexpr unsigned long long kAuxMean = 4;       // divide by n at the end
constexpr unsigned long long kAuxEwShift = 8;    // ew op in bits 8..15

template <typename T, int OP, int PRE, bool FUSED>
__global__ void __launch_bounds__(kThreads)
reduce_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
              typename Red<T, OP>::Acc* __restrict__ partials, unsigned int* __restrict__ counters,
              unsigned long long* __restrict__ results) {
  using P = Red<T, OP>;
  using Acc = typename P::Acc;
  constexpr int VN = 32 / sizeof(T);
  constexpr size_t TILE = (size_t)kThreads * kRedUnroll * kRedRounds * VN;
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * TILE;
  const size_t rem = (size_t)q.n - base;
  const T* __restrict__ a = reinterpret_cast<const T*>(q.in0) + base;
  const T* __restrict__ b = (FUSED && q.in1) ? reinterpret_cast<const T*>(q.in1) + base : nullptr;
  const T sc = scalar_as<T>(q.s0);
  const T addend = scalar_as<T>(q.s1);
  const bool has_add = (q.aux & kAuxAddend) != 0;
  const int ew = (int)((q.aux >> kAuxEwShift) & 0xff);

  auto elem = [&](T x, T y) -> Acc {
    if constexpr (FUSED) {
      x = ew_apply_rt<T>(ew, x, y);
      if (has_add) x = fadd_(x, addend);
    }
    return P::lift(pre_apply<T, PRE>(x));
  };

  Acc acc[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) acc[j] = P::init();
  bool bad = false;  // a NaN was seen (max/min only)

  const bool fast = rem >= TILE && is_aligned32(a) && (b == nullptr || is_aligned32(b));
  if (fast && FUSED && b != nullptr) {
    // two operand streams: 2 + 2 vector loads in flight per thread (the same 128 B as the
    // one-stream loop) instead of 4 + 4 — 8 live vectors cost 104 registers and half the
    // resident warps (ncu r1: 24 % occupancy, 5.7 TB/s moved)
    constexpr int H = kRedUnroll / 2;
#pragma unroll 1
    for (int it = 0; it < kRedRounds * 2; ++it) {
      VecOf<T> va[H], vb[H];
#pragma unroll
      for (int u = 0; u < H; ++u) {
        va[u].u = ldg256(a + ((size_t)(it * H + u) * kThreads + threadIdx.x) * VN);
        vb[u].u = ldg256(b + ((size_t)(it * H + u) * kThreads + threadIdx.x) * VN);
      }
#pragma unroll
      for (int u = 0; u < H; ++u) {
#pragma unroll
        for (int j = 0; j < VN; ++j) P::accum(acc[j], elem(va[u].v[j], vb[u].v[j]), bad);
      }
    }
  } else if (fast) {
#pragma unroll 1
    for (int it = 0; it < kRedRounds; ++it) {
      VecOf<T> va[kRedUnroll];
#pragma unroll
      for (int u = 0; u < kRedUnroll; ++u)
        va[u].u = ldg256(a + ((size_t)(it * kRedUnroll + u) * kThreads + threadIdx.x) * VN);
      {
#pragma unroll
        for (int u = 0; u < kRedUnroll; ++u) {
#pragma unroll
          for (int j = 0; j < VN; ++j) P::accum(acc[j], elem(va[u].v[j], sc), bad);
        }
      }
    }
  } else {
    const size_t m = rem < TILE ? rem : TILE;
    for (size_t i = threadIdx.x; i < m; i += kThreads) {
      const T y = (FUSED && b != nullptr) ? b[i] : sc;
      P::accum(acc[0], elem(a[i], y), bad);
    }
  }
  Acc v = acc[0];
#pragma unroll
  for (int j = 1; j < VN; ++j) v = P::combine(v, acc[j]);
  if constexpr (is_float_t<T>::value && OP != BCI_RED_SUM) {
    if (bad) v = Acc(NAN);
  }
  v = block_reduce<P>(v);

  // ---- second stage: last CTA of this request folds the partials in tile order
  __shared__ bool is_last;
  const unsigned long long t0 = q.tile_start;
  const unsigned long long t1 = __ldg(tile_start + r + 1);
  if (threadIdx.x == 0) {
    partials[tile] = v;
    __threadfence();
    const unsigned int done = atomicAdd(&counters[r], 1u);
    is_last = (done == (unsigned int)(t1 - t0 - 1));
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  Acc f = P::init();
  for (unsigned long long t = t0 + threadIdx.x; t < t1; t += kThreads) {
    f = P::combine(f, *reinterpret_cast<volatile Acc*>(&partials[t]));
  }
  f = block_reduce<P>(f);
  if (threadIdx.x == 0) {
    counters[r] = 0;  // self-reset: the counter table stays zero between launches
    unsigned long long bits = 0;
    if (q.aux & kAuxMean) {
      if constexpr (sizeof(T) == 4 && is_float_t<T>::value) {
        float m = (float)f / (float)q.n;  // numpy.mean(float32): f32 sum / count in f32
        bits = __float_as_uint(m);
      } else {
        double m = (double)f / (double)q.n;
        bits = (unsigned long long)__double_as_longlong(m);
      }
    } else {
      if constexpr (sizeof(Acc) == 4) {
        unsigned int lo;
        memcpy(&lo, &f, 4);
        bits = lo;
      } else {
        memcpy(&bits, &f, 8);
      }
    }
    results[r] = bits;
  }
}

// ---- reductions over a generated operand (K1+K2+K8 fused) -------------------------------------------
// `numpy.sum(numpy.square(numpy.random.rand(n)))` (examples/benchmark-numpy.py:20-21) with the
// uniform stream of rand_f64_kernel produced in registers: tile, thread -> element mapping and
// accumulator layout are those of reduce_kernel<double, OP, PRE, false> reading the materialised
// array, element e of a request is word pair (e & 1) of Philox counter e >> 1, so the result is
// bit-identical to bci_rand_f64 + bci_reduce — and no byte of the array touches HBM.
// Bound: the 20 IMAD.WIDE of Philox4x32-10 per two doubles (fma pipe), not memory.
template <int OP, int PRE>
__global__ void __launch_bounds__(kThreads)
reduce_rand_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
                   double* __restrict__ partials, unsigned int* __restrict__ counters,
                   unsigned long long* __restrict__ results) {
  using P = Red<double, OP>;
  constexpr int VN = 4;
  constexpr size_t TILE = (size_t)kThreads * kRedUnroll * kRedRounds * VN;
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * TILE;
  const size_t rem = (size_t)q.n - base;
  const PhiloxKeys keys = philox_keys((uint32_t)((unsigned long long)q.s1.i), (uint32_t)((unsigned long long)q.s1.i >> 32));

  double acc[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) acc[j] = P::init();
  bool bad = false;
  if (rem >= TILE) {
    const uint32_t phi = (uint32_t)(base >> 33);
    const uint32_t pbase = (uint32_t)(base >> 1);      // TILE is a multiple of 2^14: the low 13 bits are free
#pragma unroll 1
    for (int it = 0; it < kRedRounds; ++it) {
#pragma unroll
      for (int u = 0; u < kRedUnroll; ++u) {
        const uint32_t vec = (uint32_t)((it * kRedUnroll + u) * kThreads + threadIdx.x);   // 4 doubles = 2 counters
        const uint32_t plo = pbase | (vec << 1);
        const Philox4 a = philox4x32_10(plo, phi, 0u, 0u, keys);
        const Philox4 b = philox4x32_10(plo | 1u, phi, 0u, 0u, keys);
        P::accum(acc[0], pre_apply<double, PRE>(u53(a.x, a.y)), bad);
        P::accum(acc[1], pre_apply<double, PRE>(u53(a.z, a.w)), bad);
        P::accum(acc[2], pre_apply<double, PRE>(u53(b.x, b.y)), bad);
        P::accum(acc[3], pre_apply<double, PRE>(u53(b.z, b.w)), bad);
      }
    }
  } else {
    for (size_t i = threadIdx.x; i < rem; i += kThreads) {
      const size_t e = base + i;
      const unsigned long long p = e >> 1;
      const Philox4 a = philox4x32_10((uint32_t)p, (uint32_t)(p >> 32), 0u, 0u, keys);
      P::accum(acc[0], pre_apply<double, PRE>((e & 1) ? u53(a.z, a.w) : u53(a.x, a.y)), bad);
    }
  }
  double v = acc[0];
#pragma unroll
  for (int j = 1; j < VN; ++j) v = P::combine(v, acc[j]);
  v = block_reduce<P>(v);

  __shared__ bool is_last;
  const unsigned long long t0 = q.tile_start;
  const unsigned long long t1 = __ldg(tile_start + r + 1);
  if (threadIdx.x == 0) {
    partials[tile] = v;
    __threadfence();
    const unsigned int done = atomicAdd(&counters[r], 1u);
    is_last = (done == (unsigned int)(t1 - t0 - 1));
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double f = P::init();
  for (unsigned long long t = t0 + threadIdx.x; t < t1; t += kThreads)
    f = P::combine(f, *reinterpret_cast<volatile double*>(&partials[t]));
  f = block_reduce<P>(f);
  if (threadIdx.x == 0) {
    counters[r] = 0;
    if (q.aux & kAuxMean) f = f / (double)q.n;
    results[r] = (unsigned long long)__double_as_longlong(f);
  }
}

// ---- argmax / argmin --------------------------------------------------------------
template <typename T> struct ArgPair { T v; long long i; };

template <typename T, int OP> struct ArgRed {
  // "a beats b": numpy.argmax returns the first NaN, else the first maximum
  static __device__ __forceinline__ bool better(const ArgPair<T>& a, const ArgPair<T>& b) {
    if (b.i < 0) return a.i >= 0;
    if (a.i < 0) return false;
    if constexpr (is_float_t<T>::value) {
      const bool an = a.v != a.v, bn = b.v != b.v;
      if (an || bn) {
        if (an && bn) return a.i < b.i;
        return an;
      }
    }
    if (a.v == b.v) return a.i < b.i;
    return OP == BCI_ARGMAX ? (a.v > b.v) : (a.v < b.v);
  }
  static __device__ __forceinline__ ArgPair<T> combine(const ArgPair<T>& a, const ArgPair<T>& b) {
    return better(b, a) ? b : a;
  }
};

template <typename T, int OP>
__device__ __forceinline__ ArgPair<T> arg_block_reduce(ArgPair<T> p) {
  __shared__ T wv[kThreads / 32];
  __shared__ long long wi[kThreads / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    ArgPair<T> o;
    o.v = __shfl_down_sync(0xffffffffu, p.v, d);
    o.i = __shfl_down_sync(0xffffffffu, p.i, d);
    p = ArgRed<T, OP>::combine(p, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { wv[warp] = p.v; wi[warp] = p.i; }
  __syncthreads();
  if (warp == 0) {
    if (lane < kThreads / 32) { p.v = wv[lane]; p.i = wi[lane]; } else { p.v = T(0); p.i = -1; }
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) {
      ArgPair<T> o;
      o.v = __shfl_down_sync(0xffffffffu, p.v, d);
      o.i = __shfl_down_sync(0xffffffffu, p.i, d);
      p = ArgRed<T, OP>::combine(p, o);
    }
  }
  return p;
}

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads)
argext_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
              T* __restrict__ pval, long long* __restrict__ pidx, unsigned int* __restrict__ counters,
              unsigned long long* __restrict__ results) {
  constexpr int VN = 32 / sizeof(T);
  constexpr size_t TILE = (size_t)kThreads * kRedUnroll * kRedRounds * VN;
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * TILE;
  const size_t rem = (size_t)q.n - base;
  const T* __restrict__ a = reinterpret_cast<const T*>(q.in0) + base;

  ArgPair<T> best; best.v = T(0); best.i = -1;
  const bool fast = rem >= TILE && is_aligned32(a);
  if (fast) {
    // A thread visits its elements in increasing index order, so a strict compare keeps the
    // first occurrence and the first NaN is tracked on the side: ~5 instructions per element
    // instead of the full (value, index, NaN) comparator (ncu r1: argmax was issue-bound at 0.55).
    constexpr uint32_t kNone = 0xffffffffu;
    T bv = a[(size_t)threadIdx.x * VN];
    uint32_t bi = threadIdx.x * VN, nan_i = kNone;
#pragma unroll 1
    for (int it = 0; it < kRedRounds; ++it) {
      VecOf<T> va[kRedUnroll];
#pragma unroll
      for (int u = 0; u < kRedUnroll; ++u)
        va[u].u = ldg256(a + ((size_t)(it * kRedUnroll + u) * kThreads + threadIdx.x) * VN);
#pragma unroll
      for (int u = 0; u < kRedUnroll; ++u) {
        const uint32_t l0 = (uint32_t)((it * kRedUnroll + u) * kThreads + threadIdx.x) * VN;
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          const T x = va[u].v[j];
          if constexpr (is_float_t<T>::value) {
            if (x != x) nan_i = nan_i < l0 + j ? nan_i : l0 + j;
          }
          const bool take = OP == BCI_ARGMAX ? (x > bv) : (x < bv);
          bv = take ? x : bv;
          bi = take ? l0 + j : bi;
        }
      }
    }
    if (nan_i != kNone) {
      if constexpr (is_float_t<T>::value) best.v = T(NAN);
      best.i = (long long)(base + nan_i);
    } else {
      best.v = bv;
      best.i = (long long)(base + bi);
    }
  } else {
    const size_t m = rem < TILE ? rem : TILE;
    for (size_t i = threadIdx.x; i < m; i += kThreads) {
      ArgPair<T> c; c.v = a[i]; c.i = (long long)(base + i);
      best = ArgRed<T, OP>::combine(best, c);
    }
  }
  best = arg_block_reduce<T, OP>(best);

  __shared__ bool is_last;
  const unsigned long long t0 = q.tile_start;
  const unsigned long long t1 = __ldg(tile_start + r + 1);
  if (threadIdx.x == 0) {
    pval[tile] = best.v;
    pidx[tile] = best.i;
    __threadfence();
    const unsigned int done = atomicAdd(&counters[r], 1u);
    is_last = (done == (unsigned int)(t1 - t0 - 1));
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  ArgPair<T> f; f.v = T(0); f.i = -1;
  for (unsigned long long t = t0 + threadIdx.x; t < t1; t += kThreads) {
    ArgPair<T> c;
    c.v = *reinterpret_cast<volatile T*>(&pval[t]);
    c.i = *reinterpret_cast<volatile long long*>(&pidx[t]);
    f = ArgRed<T, OP>::combine(f, c);
  }
  f = arg_block_reduce<T, OP>(f);
  if (threadIdx.x == 0) {
    counters[r] = 0;
    results[r] = (unsigned long long)f.i;
  }
}

// ---- host side ------------------------------------------------------------------
static unsigned int* g_counters = nullptr;
static size_t g_counters_n = 0;

static int counters_get(size_t n, unsigned int** out) {
  Ctx& c = ctx();
  if (n > g_counters_n) {
    BCI_CUDA(cudaStreamSynchronize(c.stream));
    if (g_counters) cudaFree(g_counters);
    g_counters = nullptr; g_counters_n = 0;
    size_t nn = n * 2 < 4096 ? 4096 : n * 2;
    BCI_CUDA(cudaMalloc((void**)&g_counters, nn * sizeof(unsigned int)));
    BCI_CUDA(cudaMemsetAsync(g_counters, 0, nn * sizeof(unsigned int), c.stream));
    g_counters_n = nn;
  }
  *out = g_counters;
  return BCI_OK;
}

template <typename T, int OP, int PRE, bool FUSED>
static int launch_reduce(Req* reqs, int B, unsigned long long* d_results) {
  using Acc = typename Red<T, OP>::Acc;
  Ctx& c = ctx();
  Batch bt;
  int rc = upload_batch(reqs, B, red_tile_elems<T>(), &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_reduce: batch too large for one launch"); return BCI_ERR_INVALID; }
  void* partials = nullptr;
  rc = scratch_get(bt.total_tiles * sizeof(Acc), &partials);
  if (rc != BCI_OK) return rc;
  unsigned int* counters = nullptr;
  rc = counters_get((size_t)B, &counters);
  if (rc != BCI_OK) return rc;
  reduce_kernel<T, OP, PRE, FUSED><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(
      bt.d_reqs, bt.d_tile_start, B, reinterpret_cast<Acc*>(partials), counters, d_results);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

template <typename T, int OP>
static int dispatch_pre(int pre, bool fused, Req* reqs, int B, unsigned long long* d_results) {
  if (fused) {
    switch (pre) {
      case BCI_PRE_NONE: return launch_reduce<T, OP, BCI_PRE_NONE, true>(reqs, B, d_results);
      case BCI_PRE_SQUARE: return launch_reduce<T, OP, BCI_PRE_SQUARE, true>(reqs, B, d_results);
      case BCI_PRE_ABS: return launch_reduce<T, OP, BCI_PRE_ABS, true>(reqs, B, d_results);
    }
  } else {
    switch (pre) {
      case BCI_PRE_NONE: return launch_reduce<T, OP, BCI_PRE_NONE, false>(reqs, B, d_results);
      case BCI_PRE_SQUARE: return launch_reduce<T, OP, BCI_PRE_SQUARE, false>(reqs, B, d_results);
      case BCI_PRE_ABS: return launch_reduce<T, OP, BCI_PRE_ABS, false>(reqs, B, d_results);
    }
  }
  set_error("bci_reduce: unknown pre-transform %d", pre);
  return BCI_ERR_INVALID;
}

template <typename T>
static int dispatch_red(int op, int pre, bool fused, Req* reqs, int B, unsigned long long* d_results) {
  switch (op) {
    case BCI_RED_SUM:
    case BCI_RED_MEAN: return dispatch_pre<T, BCI_RED_SUM>(pre, fused, reqs, B, d_results);
    case BCI_RED_MAX: return dispatch_pre<T, BCI_RED_MAX>(pre, fused, reqs, B, d_results);
    case BCI_RED_MIN: return dispatch_pre<T, BCI_RED_MIN>(pre, fused, reqs, B, d_results);
  }
  set_error("bci_reduce: unknown op %d", op);
  return BCI_ERR_INVALID;
}

static int reduce_common(int op, int pre, int ew_op, bool fused, int dtype, int B,
                         const void* const* in0, const void* const* in1, const bci_scalar* scalar,
                         const bci_scalar* addend, const size_t* n, void* host_out) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !in0 || !n || !host_out) { set_error("bci_reduce: bad arguments"); return BCI_ERR_INVALID; }
  if (elem_size(dtype) == 0) { set_error("bci_reduce: unknown dtype %d", dtype); return BCI_ERR_INVALID; }
  if (fused && !(ew_op == BCI_EW_ADD || ew_op == BCI_EW_SUB || ew_op == BCI_EW_MUL || ew_op == BCI_EW_MAXIMUM ||
                 ew_op == BCI_EW_MINIMUM || ew_op == BCI_EW_RSUB)) {
    set_error("bci_reduce_fused: producer op %d is not supported", ew_op);
    return BCI_ERR_UNSUPPORTED;
  }
  std::vector<Req> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    Req& q = reqs[i];
    q.in0 = in0[i];
    q.in1 = (fused && in1) ? in1[i] : nullptr;
    q.out = nullptr;
    q.n = n[i];
    q.s0.i = 0; q.s1.i = 0;
    if (fused && scalar) q.s0 = scalar[i];
    q.aux = 0;
    if (fused) q.aux |= kAuxFused | ((unsigned long long)ew_op << kAuxEwShift);
    if (fused && addend) { q.s1 = addend[i]; q.aux |= kAuxAddend; }
    if (op == BCI_RED_MEAN) q.aux |= kAuxMean;
    if (n[i] == 0 && op != BCI_RED_SUM) {
      set_error("bci_reduce: request %d is empty (numpy raises/returns nan here)", i);
      return BCI_ERR_UNSUPPORTED;
    }
    if (n[i] && !in0[i]) { set_error("bci_reduce: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
    if (fused && !q.in1 && !scalar) { set_error("bci_reduce_fused: request %d has neither in1 nor scalar", i); return BCI_ERR_INVALID; }
  }
  Ctx& c = ctx();
  void *h = nullptr, *d = nullptr;
  ResultGuard guard;
  {
    std::unique_lock<std::mutex> lk(c.mu);
    int rc = result_acquire(lk, (size_t)B * 8, &guard.slot, &h, &d);
    if (rc != BCI_OK) return rc;
    BCI_CUDA(cudaMemsetAsync(d, 0, (size_t)B * 8, c.stream));  // empty sums stay 0
    unsigned long long* dres = reinterpret_cast<unsigned long long*>(d);
    switch (dtype) {
      case BCI_F32: rc = dispatch_red<float>(op, pre, fused, reqs.data(), B, dres); break;
      case BCI_F64: rc = dispatch_red<double>(op, pre, fused, reqs.data(), B, dres); break;
      case BCI_I32: rc = dispatch_red<int>(op, pre, fused, reqs.data(), B, dres); break;
      case BCI_I64: rc = dispatch_red<long long>(op, pre, fused, reqs.data(), B, dres); break;
    }
    if (rc != BCI_OK) return rc;
    if ((rc = result_enqueue_copy(guard.slot, (size_t)B * 8)) != BCI_OK) return rc;
  }
  // the lock is free while the GPU works: the next call (another broker thread) launches behind us
  int rc = result_wait(guard.slot);
  if (rc != BCI_OK) return rc;
  memcpy(host_out, h, (size_t)B * 8);
  return BCI_OK;
}

template <typename T>
static int launch_argext(int op, Req* reqs, int B, unsigned long long* d_results) {
  Ctx& c = ctx();
  Batch bt;
  int rc = upload_batch(reqs, B, red_tile_elems<T>(), &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_argext: batch too large for one launch"); return BCI_ERR_INVALID; }
  void* sc = nullptr;
  rc = scratch_get(bt.total_tiles * 16, &sc);
  if (rc != BCI_OK) return rc;
  long long* pidx = reinterpret_cast<long long*>(sc);
  T* pval = reinterpret_cast<T*>(pidx + bt.total_tiles);
  unsigned int* counters = nullptr;
  rc = counters_get((size_t)B, &counters);
  if (rc != BCI_OK) return rc;
  if (op == BCI_ARGMAX)
    argext_kernel<T, BCI_ARGMAX><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(
        bt.d_reqs, bt.d_tile_start, B, pval, pidx, counters, d_results);
  else
    argext_kernel<T, BCI_ARGMIN><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(
        bt.d_reqs, bt.d_tile_start, B, pval, pidx, counters, d_results);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

template <int OP>
static int launch_reduce_rand(int pre, Req* reqs, int B, unsigned long long* d_results) {
  Ctx& c = ctx();
  Batch bt;
  int rc = upload_batch(reqs, B, red_tile_elems<double>(), &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_reduce_rand_f64: batch too large for one launch"); return BCI_ERR_INVALID; }
  void* partials = nullptr;
  rc = scratch_get(bt.total_tiles * sizeof(double), &partials);
  if (rc != BCI_OK) return rc;
  unsigned int* counters = nullptr;
  rc = counters_get((size_t)B, &counters);
  if (rc != BCI_OK) return rc;
  const unsigned grid = (unsigned)bt.total_tiles;
  double* part = reinterpret_cast<double*>(partials);
  switch (pre) {
    case BCI_PRE_NONE: reduce_rand_kernel<OP, BCI_PRE_NONE><<<grid, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, part, counters, d_results); break;
    case BCI_PRE_SQUARE: reduce_rand_kernel<OP, BCI_PRE_SQUARE><<<grid, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, part, counters, d_results); break;
    case BCI_PRE_ABS: reduce_rand_kernel<OP, BCI_PRE_ABS><<<grid, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, part, counters, d_results); break;
    default: set_error("bci_reduce_rand_f64: unknown pre-transform %d", pre); return BCI_ERR_INVALID;
  }
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

}  // namespace bci

using namespace bci;

extern "C" int bci_reduce_rand_f64(int op, int pre, int B, const size_t* n, const uint64_t* seed, void* host_out) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !n || !seed || !host_out) { set_error("bci_reduce_rand_f64: bad arguments"); return BCI_ERR_INVALID; }
  std::vector<Req> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    Req& q = reqs[i];
    memset(&q, 0, sizeof q);
    q.n = n[i];
    q.s1.i = (int64_t)seed[i];
    if (op == BCI_RED_MEAN) q.aux |= kAuxMean;
    if (n[i] == 0 && op != BCI_RED_SUM) {
      set_error("bci_reduce_rand_f64: request %d is empty (numpy raises/returns nan here)", i);
      return BCI_ERR_UNSUPPORTED;
    }
  }
  Ctx& c = ctx();
  void *h = nullptr, *d = nullptr;
  ResultGuard guard;
  {
    std::unique_lock<std::mutex> lk(c.mu);
    int rc = result_acquire(lk, (size_t)B * 8, &guard.slot, &h, &d);
    if (rc != BCI_OK) return rc;
    BCI_CUDA(cudaMemsetAsync(d, 0, (size_t)B * 8, c.stream));
    unsigned long long* dres = reinterpret_cast<unsigned long long*>(d);
    switch (op) {
      case BCI_RED_SUM: case BCI_RED_MEAN: rc = launch_reduce_rand<BCI_RED_SUM>(pre, reqs.data(), B, dres); break;
      case BCI_RED_MAX: rc = launch_reduce_rand<BCI_RED_MAX>(pre, reqs.data(), B, dres); break;
      case BCI_RED_MIN: rc = launch_reduce_rand<BCI_RED_MIN>(pre, reqs.data(), B, dres); break;
      default: set_error("bci_reduce_rand_f64: unknown op %d", op); return BCI_ERR_INVALID;
    }
    if (rc != BCI_OK) return rc;
    if ((rc = result_enqueue_copy(guard.slot, (size_t)B * 8)) != BCI_OK) return rc;
  }
  int rc = result_wait(guard.slot);
  if (rc != BCI_OK) return rc;
  memcpy(host_out, h, (size_t)B * 8);
  return BCI_OK;
}

extern "C" int bci_reduce(int op, int pre, int dtype, int B, const void* const* in, const size_t* n,
                          void* host_out) {
  return reduce_common(op, pre, 0, false, dtype, B, in, nullptr, nullptr, nullptr, n, host_out);
}

extern "C" int bci_reduce_fused(int op, int pre, int ew_op, int dtype, int B, const void* const* in0,
                                const void* const* in1, const bci_scalar* scalar,
                                const bci_scalar* addend, const size_t* n, void* host_out) {
  return reduce_common(op, pre, ew_op, true, dtype, B, in0, in1, scalar, addend, n, host_out);
}

extern "C" int bci_argext(int op, int dtype, int B, const void* const* in, const size_t* n,
                          int64_t* host_out) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !in || !n || !host_out) { set_error("bci_argext: bad arguments"); return BCI_ERR_INVALID; }
  if (elem_size(dtype) == 0) { set_error("bci_argext: unknown dtype %d", dtype); return BCI_ERR_INVALID; }
  if (op != BCI_ARGMAX && op != BCI_ARGMIN) { set_error("bci_argext: unknown op %d", op); return BCI_ERR_INVALID; }
  std::vector<Req> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    Req& q = reqs[i];
    memset(&q, 0, sizeof q);
    q.in0 = in[i];
    q.n = n[i];
    if (n[i] == 0) { set_error("bci_argext: request %d is empty (numpy raises here)", i); return BCI_ERR_UNSUPPORTED; }
    if (!in[i]) { set_error("bci_argext: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
  }
  Ctx& c = ctx();
  void *h = nullptr, *d = nullptr;
  ResultGuard guard;
  {
    std::unique_lock<std::mutex> lk(c.mu);
    int rc = result_acquire(lk, (size_t)B * 8, &guard.slot, &h, &d);
    if (rc != BCI_OK) return rc;
    unsigned long long* dres = reinterpret_cast<unsigned long long*>(d);
    switch (dtype) {
      case BCI_F32: rc = launch_argext<float>(op, reqs.data(), B, dres); break;
      case BCI_F64: rc = launch_argext<double>(op, reqs.data(), B, dres); break;
      case BCI_I32: rc = launch_argext<int>(op, reqs.data(), B, dres); break;
      case BCI_I64: rc = launch_argext<long long>(op, reqs.data(), B, dres); break;
    }
    if (rc != BCI_OK) return rc;
    if ((rc = result_enqueue_copy(guard.slot, (size_t)B * 8)) != BCI_OK) return rc;
  }
  int rc = result_wait(guard.slot);
  if (rc != BCI_OK) return rc;
  memcpy(host_out, h, (size_t)B * 8);
  return BCI_OK;
}
