// bci_scan.cu — K7: numpy.cumsum over 1-D arrays, B requests per launch.
//
// HBM-bound: algorithmic bytes = n*(sizeof(in)+sizeof(out)).  Reduce-then-scan:
//   (1) tile_sums : one CTA sums one 128 KiB tile            (reads n)
//   (2) tile_scan : one CTA per request scans its tile sums  (tiny)
//   (3) tile_apply: one CTA re-reads its tile, scans it with warp shuffles and adds the
//                   tile prefix                               (reads n, writes n)
// => 3 passes of traffic for 2 algorithmic; no inter-CTA waiting, deterministic: the form used
// for float inputs (0.67 of the algorithmic roofline, 6.6 TB/s of traffic).  Integer inputs take
// the warp-specialised single-pass kernel further down (`scan_pipe_kernel`, BCI_SCAN=pipe: 2
// passes, int64 0.88 / int32 0.86).  The plain one-tile-per-CTA chained scan (BCI_SCAN=lookback, 0.59) is kept as
// the reference point of that design.  BCI_SCAN=legacy forces the three-pass form.
// Accumulation type: int32/int64 -> int64 (bit-exact with numpy, two's complement wrap),
// float64 -> float64, float32 -> float64 rounded once on store (numpy accumulates float32
// sequentially in float32; see SURVEY.md hard part 2 — ours is the more accurate result).
#include "bci_ops.cuh"
#include <vector>
#include <cstring>
#include <cstdlib>
#include <type_traits>

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

// Warp-wide look-back for tile `tile` of a request starting at `first_tile` (tile > first_tile):
// returns the sum of everything before the tile.  Each round examines kLbWide*32 predecessors;
// all status loads of a round are issued before any is inspected (a round costs one L2 round
// trip whatever its width).  The persistent kernel has ~900 tiles in flight and uses 8 x 32; in
// the one-tile-per-CTA kernel the extra registers cost a resident CTA, so it stays at 1 x 32.
template <typename A, int kLbWide>
__device__ __forceinline__ A lb_exclusive(const ulonglong2* __restrict__ st, unsigned long long tile,
                                          unsigned long long first_tile) {
  const int lane = threadIdx.x & 31;
  A acc = A(0);  // per-lane partial; summed across the warp once at the end
  long long p = (long long)tile - 1;
  for (;;) {
    ulonglong2 e[kLbWide];
#pragma unroll
    for (int j = 0; j < kLbWide; ++j) {
      const long long idx = p - (j * 32 + lane);
      e[j] = idx >= (long long)first_tile ? lb_load(st + idx) : make_ulonglong2(2ull, 0ull);
    }
    bool done = false;
#pragma unroll
    for (int j = 0; j < kLbWide; ++j) {
      if (done) continue;
      const long long idx = p - (j * 32 + lane);
      for (unsigned spins = 0; e[j].x == 0ull; ++spins) {   // predecessor not published yet
        if (spins > (1u << 24)) __trap();                    // a protocol bug must not hang the GPU
        e[j] = lb_load(st + idx);
      }
      const unsigned has_incl = __ballot_sync(0xffffffffu, e[j].x == 2ull);
      const int stop = has_incl ? (__ffs(has_incl) - 1) : 31;  // nearest tile whose prefix is complete
      if (lane <= stop) acc = fadd_(acc, lb_value<A>(e[j].y));
      done = has_incl != 0u;
    }
    if (done) break;
    p -= 32 * kLbWide;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc = fadd_(acc, __shfl_xor_sync(0xffffffffu, acc, d));
  return acc;
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
      excl = lb_exclusive<A, 1>(st, tile, first_tile);  // registers are this form's limiter: narrow window
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

// ---------------------------------------------------------------------------------------------
// Warp-specialised single-pass scan: persistent CTAs (one per SM), a ring of input tiles kept in
// flight by bulk copies, and the chained-scan protocol run AHEAD of the data path.
//
//   warp 8  producer : next tile id from the atomic counter -> request lookup (32 boundaries per
//                      round trip, monotone) -> stage descriptor -> ONE cp.async.bulk global->shared
//                      of the whole tile, completing on full[s].  HBM reads never wait for a
//                      look-back.  (Ragged / unaligned tiles: the warp copies them itself, zero
//                      padded.)
//   warp 9  aggregate: as soon as full[s] fires, sums the tile out of shared memory and publishes
//                      the tile aggregate — i.e. as early as the data allows, never queued behind
//                      another tile's look-back (that queueing is what made a first pipelined
//                      version 3x slower than the plain one: 56 polls per tile, ncu r1e).
//   warps 10-13 look-back (round robin over tiles): exclusive prefix of the tile from its predecessors' status words
//                      (256 per round), publishes the inclusive prefix, hands `excl` to the scan
//                      warps through ready[s].
//   warps 0-7 scan   : tile from shared memory -> registers, release the stage (empty[s]), local +
//                      warp scans, add `excl`, STG.256.  They never see the look-back latency as
//                      long as the ring is deeper than it.
// Tile ids are handed out in increasing order and every role works through the ring in order, so
// the oldest unfinished tile never waits on a younger one: no deadlock, no co-residency assumption.
__device__ __forceinline__ uint32_t sc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sc_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = sc_smem_u32(bar);
  uint32_t done = 0;
  for (unsigned long long spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spins > (1ull << 26)) __trap();  // a protocol bug must not hang the GPU
  }
}
__device__ __forceinline__ void sc_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sc_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(sc_smem_u32(bar)) : "memory");
}

constexpr int kWsScanWarps = kThreads / 32;              // 8
constexpr int kWsLbWarps = 4;                            // look-backs of consecutive tiles overlap
constexpr int kWsThreads = kThreads + (2 + kWsLbWarps) * 32;  // + producer, aggregate, look-back warps
template <typename T> struct Pipe {
  static constexpr int SUB = 4;                                              // 32 KiB of input per tile (4096 or 8192 elements)
  static constexpr size_t TILE = (size_t)kThreads * SUB * (32 / sizeof(T)); // 4096 elements
  static constexpr size_t TILE_BYTES = TILE * sizeof(T);
  static constexpr int STAGES = (192 * 1024) / TILE_BYTES > 8 ? 8 : (int)((192 * 1024) / TILE_BYTES);
  static_assert(STAGES > kWsLbWarps, "the end markers must not wrap onto each other");
  static constexpr size_t SMEM = (size_t)STAGES * TILE_BYTES + 128;
};

struct alignas(16) PipeStage {
  const void* in;                  // first element of the tile
  void* out;
  unsigned long long rem;          // elements from `in` to the end of the request
  unsigned long long tile;
  unsigned long long first_tile;   // of the tile's request
  unsigned int valid, fast;
};

template <typename T>
__global__ void __launch_bounds__(kWsThreads, 1)
scan_pipe_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
                 unsigned long long total_tiles, ulonglong2* __restrict__ st, uint32_t* __restrict__ counter,
                 unsigned int prefetch_dist) {
  using A = typename ScanTypes<T>::Acc;
  using O = typename ScanTypes<T>::Out;
  constexpr int VN = 32 / sizeof(T);
  constexpr int VO = 32 / sizeof(O);
  constexpr int NST = VN / VO;
  constexpr int SUB = Pipe<T>::SUB;
  constexpr int STAGES = Pipe<T>::STAGES;
  constexpr size_t TILE = Pipe<T>::TILE;
  constexpr uint32_t TILE_BYTES = (uint32_t)Pipe<T>::TILE_BYTES;
  constexpr int NW = kWsScanWarps;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full[STAGES], aggd[STAGES], ready[STAGES], empty[STAGES];
  __shared__ PipeStage meta[STAGES];
  __shared__ A agg_s[STAGES], excl_s[STAGES];
  __shared__ A warp_tot2[2][SUB][NW];  // double-buffered by iteration parity: one named barrier per tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      sc_mbar_init(&full[s], 1);
      sc_mbar_init(&aggd[s], 1);
      sc_mbar_init(&ready[s], 1);
      sc_mbar_init(&empty[s], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NW) {
    // ------------------------------------------------------------------ producer
    int r = 0;   // request of the last tile fetched (tile ids only grow)
    int r2 = 0;  // request of the last tile prefetched
    int tail = 0;  // end markers written: one per look-back warp, so each of them meets one
    for (unsigned it = 0;; ++it) {
      const int s = (int)(it % STAGES);
      sc_mbar_wait(&empty[s], ((it / STAGES) & 1u) ^ 1u);  // first lap: passes immediately
      unsigned int id = 0;
      if (lane == 0) id = atomicAdd(counter, 1u);
      id = __shfl_sync(0xffffffffu, id, 0);
      const unsigned long long tile = id;
      if (tile >= total_tiles) {
        if (lane == 0) {
          meta[s].valid = 0u;
          sc_mbar_arrive(&full[s]);
        }
        if (++tail == kWsLbWarps) break;
        continue;
      }
      for (;;) {  // boundaries tile_start[r+1 ...] that are <= tile
        const int i = r + 1 + lane;
        const bool le = i < B && __ldg(tile_start + i) <= tile;
        const int cnt = __popc(__ballot_sync(0xffffffffu, le));
        r += cnt;
        if (cnt < 32) break;
      }
      const Req q = reqs[r];
      const size_t base = (size_t)(tile - q.tile_start) * TILE;
      const size_t rem = (size_t)q.n - base;
      const T* a = reinterpret_cast<const T*>(q.in0) + base;
      O* o = reinterpret_cast<O*>(q.out) + base;
      const bool fast = rem >= TILE && is_aligned32(a) && is_aligned32(o);
      if (lane == 0) {
        PipeStage m;
        m.in = a; m.out = o; m.rem = rem; m.tile = tile; m.first_tile = q.tile_start; m.valid = 1u; m.fast = fast ? 1u : 0u;
        meta[s] = m;
      }
      if (fast) {
        if (lane == 0) {
          sc_mbar_expect_tx(&full[s], TILE_BYTES);
          sc_bulk_load(ring + (size_t)s * TILE_BYTES, a, TILE_BYTES, &full[s]);
        }
      } else {
        T* dst = reinterpret_cast<T*>(ring + (size_t)s * TILE_BYTES);
#pragma unroll 8
        for (unsigned e = lane; e < (unsigned)TILE; e += 32) dst[e] = e < rem ? a[e] : T(0);
        __syncwarp();
        if (lane == 0) sc_mbar_arrive(&full[s]);
      }
      // L2 prefetch of the tile `prefetch_dist` ids ahead (whoever takes it): with the ring full
      // all the time a bulk copy waits ~9 us in the HBM queue (bytes in flight / bandwidth), which
      // is most of a stage's residence; a copy that hits L2 frees the stage sooner, and the bytes
      // in flight towards HBM are no longer bounded by shared memory.
      const unsigned long long ahead = tile + prefetch_dist;
      if (prefetch_dist != 0 && ahead < total_tiles) {
        for (;;) {
          const int i = r2 + 1 + lane;
          const bool le = i < B && __ldg(tile_start + i) <= ahead;
          const int cnt = __popc(__ballot_sync(0xffffffffu, le));
          r2 += cnt;
          if (cnt < 32) break;
        }
        if (lane == 0) {
          const Req q2 = reqs[r2];
          const size_t base2 = (size_t)(ahead - q2.tile_start) * TILE;
          const size_t rem2 = (size_t)q2.n - base2;
          const T* a2 = reinterpret_cast<const T*>(q2.in0) + base2;
          const uint32_t bytes2 = (uint32_t)((rem2 < TILE ? rem2 : TILE) * sizeof(T)) & ~15u;
          if (bytes2 != 0 && (reinterpret_cast<uintptr_t>(a2) & 15) == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a2), "r"(bytes2) : "memory");
        }
      }
    }
  } else if (warp == NW + 1) {
    // ------------------------------------------------------------------ aggregate
    int tail = 0;
    for (unsigned it = 0;; ++it) {
      const int s = (int)(it % STAGES);
      sc_mbar_wait(&full[s], (it / STAGES) & 1u);
      const PipeStage m = meta[s];
      if (!m.valid) {
        if (lane == 0) sc_mbar_arrive(&aggd[s]);
        if (++tail == kWsLbWarps) break;
        continue;
      }
      const uint4* src = reinterpret_cast<const uint4*>(ring + (size_t)s * TILE_BYTES);
      A part[4] = {A(0), A(0), A(0), A(0)};
#pragma unroll 4
      for (unsigned i = 0; i < TILE_BYTES / 16 / 32; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          union { uint4 q; T v[16 / sizeof(T)]; } w;
          w.q = src[(size_t)(i + u) * 32 + lane];
#pragma unroll
          for (int j = 0; j < (int)(16 / sizeof(T)); ++j) part[u] = fadd_(part[u], (A)w.v[j]);
        }
      }
      A agg = fadd_(fadd_(part[0], part[1]), fadd_(part[2], part[3]));
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) agg = fadd_(agg, __shfl_xor_sync(0xffffffffu, agg, d));
      if (lane == 0) {
        lb_store(st + m.tile, m.tile == m.first_tile ? 2ull : 1ull, lb_bits<A>(agg));
        agg_s[s] = agg;
        sc_mbar_arrive(&aggd[s]);
      }
    }
  } else if (warp >= NW + 2) {
    // ------------------------------------------------------------------ look-back (tile it -> warp it % kWsLbWarps)
    for (unsigned it = (unsigned)(warp - (NW + 2));; it += kWsLbWarps) {
      const int s = (int)(it % STAGES);
      sc_mbar_wait(&aggd[s], (it / STAGES) & 1u);
      const PipeStage m = meta[s];
      if (!m.valid) {
        if (lane == 0) sc_mbar_arrive(&ready[s]);
        break;
      }
      A excl = A(0);
      if (m.tile != m.first_tile) {
        excl = lb_exclusive<A, 8>(st, m.tile, m.first_tile);
        if (lane == 0) lb_store(st + m.tile, 2ull, lb_bits<A>(fadd_(excl, agg_s[s])));
      }
      if (lane == 0) {
        excl_s[s] = excl;
        sc_mbar_arrive(&ready[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ scan warps
    for (unsigned it = 0;; ++it) {
      const int s = (int)(it % STAGES);
      const uint32_t parity = (it / STAGES) & 1u;
      sc_mbar_wait(&full[s], parity);
      const PipeStage m = meta[s];
      if (!m.valid) break;
      O* __restrict__ o = reinterpret_cast<O*>(m.out);
      const size_t rem = (size_t)m.rem;
      const bool fast = m.fast != 0u;
      A (*warp_tot)[NW] = warp_tot2[it & 1u];
      A x[SUB][VN];
      A lane_excl[SUB];
      {
        const uint8_t* src = ring + (size_t)s * TILE_BYTES;
        VecOf<T> v[SUB];
#pragma unroll
        for (int k = 0; k < SUB; ++k) {
          const uint4* p = reinterpret_cast<const uint4*>(src + ((size_t)k * kThreads + threadIdx.x) * 32);
          const uint4 lo = p[0], hi = p[1];
          v[k].u.r[0] = lo.x; v[k].u.r[1] = lo.y; v[k].u.r[2] = lo.z; v[k].u.r[3] = lo.w;
          v[k].u.r[4] = hi.x; v[k].u.r[5] = hi.y; v[k].u.r[6] = hi.z; v[k].u.r[7] = hi.w;
        }
#pragma unroll
        for (int k = 0; k < SUB; ++k) {
#pragma unroll
          for (int j = 0; j < VN; ++j) x[k][j] = (A)v[k].v[j];
        }
      }
#pragma unroll
      for (int k = 0; k < SUB; ++k) {
#pragma unroll
        for (int j = 1; j < VN; ++j) x[k][j] = fadd_(x[k][j - 1], x[k][j]);
        A inc = x[k][VN - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const A y = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc = fadd_(inc, y);
        }
        if (lane == 31) warp_tot[k][warp] = inc;
        A ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = A(0);
        lane_excl[k] = ex;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");  // the 8 scan warps only
      sc_mbar_wait(&ready[s], parity);  // the look-back ran ahead (or finishes while we scanned)
      A carry = excl_s[s];
      __syncwarp();
      // release the stage only now: ready[s] / excl_s[s] must not be re-armed for a later tile
      // before every scan warp has consumed this one's
      if (lane == 0) sc_mbar_arrive(&empty[s]);
#pragma unroll
      for (int k = 0; k < SUB; ++k) {
        A wprefix = A(0), all = A(0);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const A wt = warp_tot[k][w];
          if (w < warp) wprefix = fadd_(wprefix, wt);
          all = fadd_(all, wt);
        }
        const A off = fadd_(fadd_(carry, wprefix), lane_excl[k]);
        const size_t e0 = ((size_t)k * kThreads + threadIdx.x) * VN;
        if (fast) {
#pragma unroll
          for (int kk = 0; kk < NST; ++kk) {
            VecOf<O> vo;
#pragma unroll
            for (int j = 0; j < VO; ++j) vo.v[j] = (O)fadd_(off, x[k][kk * VO + j]);
            stg256(o + e0 + (size_t)kk * VO, vo.u);
          }
        } else {
#pragma unroll
          for (int j = 0; j < VN; ++j)
            if (e0 + j < rem) o[e0 + j] = (O)fadd_(off, x[k][j]);
        }
        carry = fadd_(carry, all);
      }
    }
  }
}

template <typename T>
static int launch_cumsum_pipe(Req* reqs, int B) {
  Ctx& c = ctx();
  Batch bt;
  int rc = upload_batch(reqs, B, Pipe<T>::TILE, &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_cumsum: batch too large for one launch"); return BCI_ERR_INVALID; }
  const size_t bytes = 256 + bt.total_tiles * sizeof(ulonglong2);
  void* sc = nullptr;
  rc = scratch_get(bytes, &sc);
  if (rc != BCI_OK) return rc;
  uint8_t* sb = reinterpret_cast<uint8_t*>(sc);
  uint32_t* counter = reinterpret_cast<uint32_t*>(sb);
  ulonglong2* st = reinterpret_cast<ulonglong2*>(sb + 256);
  BCI_CUDA(cudaMemsetAsync(sb, 0, bytes, c.stream));
  static bool configured = false;
  if (!configured) {
    BCI_CUDA(cudaFuncSetAttribute(scan_pipe_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Pipe<T>::SMEM));
    configured = true;
  }
  unsigned long long grid = (unsigned long long)c.sm_count;
  if (grid > bt.total_tiles) grid = bt.total_tiles;
  // prefetch distance in tiles (BCI_SCAN_PREFETCH, 0 = off).  Measured, int64 1e7 x 32: off 0.79,
  // 74 tiles 0.86, 148-592 0.88 (flat), 888 0.81, 1776 0.67 (evicted before use) -> 2 x SMs
  unsigned int prefetch_dist = (unsigned int)(c.sm_count * 2);
  if (const char* pd = tuning_env("BCI_SCAN_PREFETCH")) prefetch_dist = (unsigned int)atoi(pd);
  scan_pipe_kernel<T><<<(unsigned)grid, kWsThreads, Pipe<T>::SMEM, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, bt.total_tiles, st, counter,
                                                                              prefetch_dist);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

template <typename T>
static int launch_cumsum(Req* reqs, int B) {
  {
    // Measured on B200, int64 1e7 x 32 (tools/prof_scan.py): three-pass 4.42 TB/s algorithmic
    // (0.67), one-tile-per-CTA look-back 3.88 (0.59), warp-specialised pipeline 5.15 (0.78), 5.75 (0.88) with the L2 prefetch.
    const char* force = tuning_env("BCI_SCAN");
    if (force && strcmp(force, "lookback") == 0) return launch_cumsum_lookback<T>(reqs, B);
    if (force && strcmp(force, "pipe") == 0) return launch_cumsum_pipe<T>(reqs, B);
    // Integers: the warp-specialised single-pass kernel (two's-complement sums are associative, so
    // the timing-dependent look-back depth cannot change a bit).  Floats stay on the three-pass
    // form: its summation order is fixed, a chained scan's is not.
    if (!(force && strcmp(force, "legacy") == 0) && std::is_integral<T>::value) return launch_cumsum_pipe<T>(reqs, B);
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
