// bci_scan.cu — K7: numpy.cumsum over 1-D arrays, B requests per launch.
//
// HBM-bound: algorithmic bytes = n*(sizeof(in)+sizeof(out)).  Reduce-then-scan:
//   (1) tile_sums : one CTA sums one 128 KiB tile            (reads n)
//   (2) tile_scan : one CTA per request scans its tile sums  (tiny)
//   (3) tile_apply: one CTA re-reads its tile, scans it with warp shuffles and adds the
//                   tile prefix                               (reads n, writes n)
// => 3 passes of traffic for 2 algorithmic; no inter-CTA waiting, deterministic.  A single-pass
// chained scan with decoupled look-back is implemented below as well (BCI_SCAN=lookback): it
// moves 2 passes but, one tile per CTA, is latency-bound at 3.9 TB/s on B200 where this form
// streams at 6.6 TB/s — so this one is the default until the pipelined version lands.
// Accumulation type: int32/int64 -> int64 (bit-exact with numpy, two's complement wrap),
// float64 -> float64, float32 -> float64 rounded once on store (numpy accumulates float32
// sequentially in float32; see SURVEY.md hard part 2 — ours is the more accurate result).
#include "bci_ops.cuh"
#include <vector>
#include <cstring>
#include <cstdlib>

namespace bci {

template <typename T> struct ScanTypes;
template <> struct ScanTypes<float> { using Acc = double; using Out = float; };
template <> struct ScanTypes<double> { using Acc = double; using Out = double; };
template <> struct ScanTypes<int> { using Acc = long long; using Out = long long; };
template <> struct ScanTypes<long long> { using Acc = long long; using Out = long long; };

constexpr int kScanSub = 16;  // sub-tiles (of kThreads 256-bit vectors) per tile
template <typename T> constexpr size_t scan_tile_elems() { return (size_t)kThreads * kScanSub * (32 / sizeof(T)); }

template <typename A> __device__ __forceinline__ A block_sum(A v) {
  __shared__ A part[kThreads / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fadd_(v, __shfl_down_sync(0xffffffffu, v, d));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < kThreads / 32 ? part[lane] : A(0);
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) v = fadd_(v, __shfl_down_sync(0xffffffffu, v, d));
  }
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
scan_tile_sums(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
               typename ScanTypes<T>::Acc* __restrict__ tile_sum) {
  using A = typename ScanTypes<T>::Acc;
  constexpr int VN = 32 / sizeof(T);
  constexpr size_t TILE = (size_t)kThreads * kScanSub * VN;
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * TILE;
  const size_t rem = (size_t)q.n - base;
  const T* __restrict__ a = reinterpret_cast<const T*>(q.in0) + base;
  A acc[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) acc[j] = A(0);
  if (rem >= TILE && is_aligned32(a)) {
#pragma unroll 1
    for (int it = 0; it < kScanSub; it += 4) {
      VecOf<T> v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u].u = ldg256(a + ((size_t)(it + u) * kThreads + threadIdx.x) * VN);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < VN; ++j) acc[j] = fadd_(acc[j], (A)v[u].v[j]);
      }
    }
  } else {
    const size_t m = rem < TILE ? rem : TILE;
    for (size_t i = threadIdx.x; i < m; i += kThreads) acc[0] = fadd_(acc[0], (A)a[i]);
  }
  A s = acc[0];
#pragma unroll
  for (int j = 1; j < VN; ++j) s = fadd_(s, acc[j]);
  s = block_sum<A>(s);
  if (threadIdx.x == 0) tile_sum[tile] = s;
}

// one CTA per request: exclusive scan of that request's tile sums, in place
template <typename A>
__global__ void __launch_bounds__(kThreads)
scan_tile_prefix(const unsigned long long* __restrict__ tile_start, A* __restrict__ tile_sum) {
  __shared__ A warp_tot[kThreads / 32];
  __shared__ A carry_s;
  const int r = blockIdx.x;
  const unsigned long long t0 = tile_start[r], t1 = tile_start[r + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = A(0);
  __syncthreads();
  for (unsigned long long b = t0; b < t1; b += kThreads) {
    const unsigned long long t = b + threadIdx.x;
    const A x = t < t1 ? tile_sum[t] : A(0);
    A inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const A y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc = fadd_(inc, y);
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    A wprefix = A(0);
    for (int w = 0; w < warp; ++w) wprefix = fadd_(wprefix, warp_tot[w]);
    const A carry = carry_s;
    A excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = A(0);
    if (t < t1) tile_sum[t] = fadd_(fadd_(carry, wprefix), excl);
    __syncthreads();
    if (threadIdx.x == kThreads - 1) carry_s = fadd_(fadd_(carry, wprefix), inc);
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
scan_tile_apply(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
                const typename ScanTypes<T>::Acc* __restrict__ tile_prefix) {
  using A = typename ScanTypes<T>::Acc;
  using O = typename ScanTypes<T>::Out;
  constexpr int VN = 32 / sizeof(T);
  constexpr int VO = 32 / sizeof(O);       // outputs per 256-bit store
  constexpr int NST = VN / VO;             // stores per loaded vector (1 or 2)
  constexpr size_t TILE = (size_t)kThreads * kScanSub * VN;
  __shared__ A warp_tot[kThreads / 32];
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * TILE;
  const size_t rem = (size_t)q.n - base;
  const T* __restrict__ a = reinterpret_cast<const T*>(q.in0) + base;
  O* __restrict__ o = reinterpret_cast<O*>(q.out) + base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  A carry = tile_prefix[tile];
  const bool fast = rem >= TILE && is_aligned32(a) && is_aligned32(o);
  const int nsub = fast ? kScanSub : (int)(((rem < TILE ? rem : TILE) + (size_t)kThreads * VN - 1) / ((size_t)kThreads * VN));
#pragma unroll 1
  for (int s = 0; s < nsub; ++s) {
    const size_t e0 = ((size_t)s * kThreads + threadIdx.x) * VN;  // first element of this thread
    A x[VN];
    if (fast) {
      VecOf<T> v;
      v.u = ldg256(a + e0);
#pragma unroll
      for (int j = 0; j < VN; ++j) x[j] = (A)v.v[j];
    } else {
#pragma unroll
      for (int j = 0; j < VN; ++j) x[j] = (e0 + j < rem) ? (A)a[e0 + j] : A(0);
    }
    // thread-local inclusive scan
#pragma unroll
    for (int j = 1; j < VN; ++j) x[j] = fadd_(x[j - 1], x[j]);
    const A tot = x[VN - 1];
    A inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const A y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc = fadd_(inc, y);
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    A wprefix = A(0), all = A(0);
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      const A wt = warp_tot[w];
      if (w < warp) wprefix = fadd_(wprefix, wt);
      all = fadd_(all, wt);
    }
    A excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = A(0);
    const A off = fadd_(fadd_(carry, wprefix), excl);
    if (fast) {
#pragma unroll
      for (int k = 0; k < NST; ++k) {
        VecOf<O> vo;
#pragma unroll
        for (int j = 0; j < VO; ++j) vo.v[j] = (O)fadd_(off, x[k * VO + j]);
        stg256(o + e0 + (size_t)k * VO, vo.u);
      }
    } else {
#pragma unroll
      for (int j = 0; j < VN; ++j)
        if (e0 + j < rem) o[e0 + j] = (O)fadd_(off, x[j]);
    }
    carry = fadd_(carry, all);
    __syncthreads();
  }
}

// ---- single-pass form: chained scan with decoupled look-back (Merrill & Garland 2016) --------
// One read + one write of the data (the 2 algorithmic passes).  A tile (32 KiB of input held in
// registers) publishes its aggregate, warp 0 sums the aggregates of the preceding tiles of the
// SAME request 32 at a time until it meets one whose inclusive prefix is known, then the tile
// publishes its own inclusive prefix and writes its scanned values.  Flags and 64-bit values
// live in separate arrays (value first, release-store of the flag after).  Tile indices come
// from an atomic counter, so a tile only waits for tiles that already started.
// sub-tiles (256 threads x one 256-bit vector) per tile: every thread holds 32 elements, i.e.
// 32 KiB tiles for 4-byte inputs and 64 KiB tiles for 8-byte inputs
template <typename T> constexpr int lb_sub() { return 4; }
template <typename T> constexpr size_t lb_tile_elems() { return (size_t)kThreads * lb_sub<T>() * (32 / sizeof(T)); }

// One 16-byte status entry per tile = {flag (0 none, 1 aggregate, 2 inclusive prefix), value
// bits}, read and written as ONE 128-bit cache-global transaction, so flag and value can never
// be seen torn and no fence / L1 invalidation (ld.acquire = CCTL.IVALL + LD) is needed.
__device__ __forceinline__ ulonglong2 lb_load(const ulonglong2* p) {
  ulonglong2 v;
  asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void lb_store(ulonglong2* p, unsigned long long flag, unsigned long long bits) {
  asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(flag), "l"(bits) : "memory");
}
template <typename A> __device__ __forceinline__ unsigned long long lb_bits(A v) {
  unsigned long long b;
  memcpy(&b, &v, 8);
  return b;
}
template <typename A> __device__ __forceinline__ A lb_value(unsigned long long b) {
  A v;
  memcpy(&v, &b, 8);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
scan_lookback_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
                     ulonglong2* __restrict__ st, uint32_t* __restrict__ counter) {
  using A = typename ScanTypes<T>::Acc;
  using O = typename ScanTypes<T>::Out;
  constexpr int VN = 32 / sizeof(T);
  constexpr int VO = 32 / sizeof(O);
  constexpr int NST = VN / VO;
  constexpr int kLbSub = 4;
  constexpr size_t TILE = (size_t)kThreads * kLbSub * VN;
  __shared__ A warp_tot[kLbSub][kThreads / 32];
  __shared__ A tile_excl;
  __shared__ unsigned int tile_id;
  if (threadIdx.x == 0) tile_id = atomicAdd(counter, 1u);
  __syncthreads();
  const unsigned long long tile = tile_id;
  const int r = find_req_block(tile_start, B, tile);
  const Req q = reqs[r];
  const unsigned long long first_tile = q.tile_start;
  const size_t base = (size_t)(tile - first_tile) * TILE;
  const size_t rem = (size_t)q.n - base;
  const T* __restrict__ a = reinterpret_cast<const T*>(q.in0) + base;
  O* __restrict__ o = reinterpret_cast<O*>(q.out) + base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool fast = rem >= TILE && is_aligned32(a) && is_aligned32(o);

  // ---- load the tile, thread-local inclusive scans, warp scans of the thread totals
  A x[kLbSub][VN];
  A lane_excl[kLbSub];
  if (fast) {
    // all loads first, conversions after: a branch or a convert between the loads makes every
    // one of them expose its full latency (ncu r1d: 13 k long-scoreboard samples on this loop)
    VecOf<T> v[kLbSub];
#pragma unroll
    for (int s = 0; s < kLbSub; ++s) v[s].u = ldg256(a + ((size_t)s * kThreads + threadIdx.x) * VN);
#pragma unroll
    for (int s = 0; s < kLbSub; ++s) {
#pragma unroll
      for (int j = 0; j < VN; ++j) x[s][j] = (A)v[s].v[j];
    }
  } else {
#pragma unroll
    for (int s = 0; s < kLbSub; ++s) {
      const size_t e0 = ((size_t)s * kThreads + threadIdx.x) * VN;
#pragma unroll
      for (int j = 0; j < VN; ++j) x[s][j] = (e0 + j < rem) ? (A)a[e0 + j] : A(0);
    }
  }
#pragma unroll
  for (int s = 0; s < kLbSub; ++s) {
#pragma unroll
    for (int j = 1; j < VN; ++j) x[s][j] = fadd_(x[s][j - 1], x[s][j]);
    A inc = x[s][VN - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const A y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc = fadd_(inc, y);
    }
    if (lane == 31) warp_tot[s][warp] = inc;
    A ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = A(0);
    lane_excl[s] = ex;
  }
  __syncthreads();
  // ---- warp 0: tile aggregate, publish, look back
  if (warp == 0) {
    // the (sub-tile, warp) totals, one or two per lane
    A agg = A(0);
#pragma unroll
    for (int e = 0; e < kLbSub * (kThreads / 32); e += 32)
      agg = fadd_(agg, warp_tot[(e + lane) / (kThreads / 32)][(e + lane) % (kThreads / 32)]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) agg = fadd_(agg, __shfl_xor_sync(0xffffffffu, agg, d));
    const bool is_first = tile == first_tile;
    if (lane == 0) lb_store(st + tile, is_first ? 2ull : 1ull, lb_bits<A>(agg));
    A excl = A(0);
    if (!is_first) {
      long long p = (long long)tile - 1;
      for (;;) {
        const long long idx = p - lane;
        unsigned long long f = 2ull;
        A v = A(0);
        if (idx >= (long long)first_tile) {
          ulonglong2 e = lb_load(st + idx);
          for (unsigned spins = 0; e.x == 0ull; ++spins) {
            if (spins > (1u << 24)) __trap();  // a protocol bug must not hang the GPU
            e = lb_load(st + idx);
          }
          f = e.x;
          v = lb_value<A>(e.y);
        }
        const unsigned has_incl = __ballot_sync(0xffffffffu, f == 2ull);
        const int stop = has_incl ? (__ffs(has_incl) - 1) : 31;  // nearest tile whose prefix is complete
        A contrib = lane <= stop ? v : A(0);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) contrib = fadd_(contrib, __shfl_xor_sync(0xffffffffu, contrib, d));
        excl = fadd_(excl, contrib);
        if (has_incl) break;
        p -= 32;
      }
      if (lane == 0) lb_store(st + tile, 2ull, lb_bits<A>(fadd_(excl, agg)));
    }
    if (lane == 0) tile_excl = excl;
  }
  __syncthreads();
  // ---- write: tile prefix + earlier (sub-tile, warp) totals + lane prefix + local scan
  A carry = tile_excl;
#pragma unroll
  for (int s = 0; s < kLbSub; ++s) {
    A wprefix = A(0), all = A(0);
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      const A wt = warp_tot[s][w];
      if (w < warp) wprefix = fadd_(wprefix, wt);
      all = fadd_(all, wt);
    }
    const A off = fadd_(fadd_(carry, wprefix), lane_excl[s]);
    const size_t e0 = ((size_t)s * kThreads + threadIdx.x) * VN;
    if (fast) {
#pragma unroll
      for (int k = 0; k < NST; ++k) {
        VecOf<O> vo;
#pragma unroll
        for (int j = 0; j < VO; ++j) vo.v[j] = (O)fadd_(off, x[s][k * VO + j]);
        stg256(o + e0 + (size_t)k * VO, vo.u);
      }
    } else {
#pragma unroll
      for (int j = 0; j < VN; ++j)
        if (e0 + j < rem) o[e0 + j] = (O)fadd_(off, x[s][j]);
    }
    carry = fadd_(carry, all);
  }
}

template <typename T>
static int launch_cumsum_lookback(Req* reqs, int B) {
  Ctx& c = ctx();
  Batch bt;
  int rc = upload_batch(reqs, B, lb_tile_elems<T>(), &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_cumsum: batch too large for one launch"); return BCI_ERR_INVALID; }
  // scratch: [counter (256 B)][status T x 16 B]
  const size_t bytes = 256 + bt.total_tiles * sizeof(ulonglong2);
  void* sc = nullptr;
  rc = scratch_get(bytes, &sc);
  if (rc != BCI_OK) return rc;
  uint8_t* sb = reinterpret_cast<uint8_t*>(sc);
  uint32_t* counter = reinterpret_cast<uint32_t*>(sb);
  ulonglong2* st = reinterpret_cast<ulonglong2*>(sb + 256);
  BCI_CUDA(cudaMemsetAsync(sb, 0, bytes, c.stream));
  scan_lookback_kernel<T><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, st, counter);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

template <typename T>
static int launch_cumsum(Req* reqs, int B) {
  {
    // Measured on B200 (profiles/r1_kernel_table.md): the 3-pass form below runs at 6.6 TB/s of
    // traffic (0.67 of the algorithmic roofline), the single-pass look-back form at 3.9 TB/s
    // (0.59): one-tile-per-CTA look-back is latency-bound at 2-4 CTAs/SM.  The faster one is the
    // default; BCI_SCAN=lookback selects the single-pass kernel (kept for the pipelined rewrite).
    const char* force = getenv("BCI_SCAN");
    if (force && strcmp(force, "lookback") == 0) return launch_cumsum_lookback<T>(reqs, B);
  }
  using A = typename ScanTypes<T>::Acc;
  Ctx& c = ctx();
  Batch bt;
  int rc = upload_batch(reqs, B, scan_tile_elems<T>(), &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_cumsum: batch too large for one launch"); return BCI_ERR_INVALID; }
  void* sc = nullptr;
  rc = scratch_get(bt.total_tiles * sizeof(A), &sc);
  if (rc != BCI_OK) return rc;
  A* tile_sum = reinterpret_cast<A*>(sc);
  scan_tile_sums<T><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, tile_sum);
  BCI_CHECK_LAUNCH();
  scan_tile_prefix<A><<<(unsigned)B, kThreads, 0, c.stream>>>(bt.d_tile_start, tile_sum);
  BCI_CHECK_LAUNCH();
  scan_tile_apply<T><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, tile_sum);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

}  // namespace bci

using namespace bci;

extern "C" int bci_cumsum(int dtype, int B, const void* const* in, void* const* out, const size_t* n) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !in || !out || !n) { set_error("bci_cumsum: bad arguments"); return BCI_ERR_INVALID; }
  if (elem_size(dtype) == 0) { set_error("bci_cumsum: unknown dtype %d", dtype); return BCI_ERR_INVALID; }
  std::vector<Req> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    Req& q = reqs[i];
    memset(&q, 0, sizeof q);
    q.in0 = in[i];
    q.out = out[i];
    q.n = n[i];
    if (n[i] && (!in[i] || !out[i])) { set_error("bci_cumsum: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
  }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  switch (dtype) {
    case BCI_F32: return launch_cumsum<float>(reqs.data(), B);
    case BCI_F64: return launch_cumsum<double>(reqs.data(), B);
    case BCI_I32: return launch_cumsum<int>(reqs.data(), B);
    case BCI_I64: return launch_cumsum<long long>(reqs.data(), B);
  }
  return BCI_ERR_INVALID;
}
