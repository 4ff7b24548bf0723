// bci_sort.cu — K5: numpy.sort (values) and numpy.argsort(kind="stable") over 1-D arrays,
// B requests per launch.
//
// LSD radix sort, 8-bit digits, segmented by request.  Keys are mapped to order-preserving
// unsigned integers on the first pass and mapped back on the last one:
//   float : flip all bits of negatives, flip the sign bit of non-negatives; every NaN
//           becomes the all-ones key, so NaNs sort last like numpy (their payload is
//           canonicalised to +qNaN; -0.0 sorts before +0.0, value-equal to numpy);
//   int   : flip the sign bit.
// This file: the dispatcher and the multi-pass fallback (BCI_SORT=legacy, or n >= 2^30); the
// default is the one-sweep form in bci_sort_onesweep.cu.
// Per pass (reduce-then-scan, no inter-CTA waiting):
//   hist    : one CTA per 4096-key tile counts its 256 digits (uniform rows — the exponent
//             byte of uniform floats — collapse to one shared atomic) -> hist[req][digit][tile]
//   scan    : one CTA per (request, digit): exclusive scan over that digit's tiles, based
//             at the sum of the lower digits' totals -> global offsets, in place
//   scatter : ranks keys stably inside the tile (8 ballots per 32-key row, per-warp digit
//             counters), stages the tile in digit order in shared memory and writes each
//             digit run to its global offset with coalesced stores.
// HBM traffic: 3 key passes per digit pass (hist read + scatter read/write) where the
// one-sweep/look-back form needs 2: algorithmic bytes are 9*n*4 for
// float32, this implementation moves 12*n*4.  Bit-exact with numpy for all dtypes.
#include "bci_sort_common.cuh"
#include <vector>
#include <cstring>
#include <cstdlib>

namespace bci {

// Req fields for the sort kernels:
//   in0 = source keys, in1 = source indices (u32; nullptr on the first pass: generated),
//   out = destination keys (nullptr when only indices leave the last pass),
//   s0.i = destination indices (u32*, or int64* on the last pass), s1.i = offset (in u32
//   words) of this request's hist[256][T] block, aux = T (tiles of this request).

template <typename K>
__global__ void __launch_bounds__(kThreads)
sort_hist_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
                 int shift, int dtype, int first, uint32_t* __restrict__ hist, uint32_t* __restrict__ totals) {
  __shared__ uint32_t sh[256];
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const unsigned long long tl = tile - q.tile_start;
  const size_t base = (size_t)tl * kSortTile;
  const size_t rem = (size_t)q.n - base;
  const int m = rem < (size_t)kSortTile ? (int)rem : kSortTile;
  const K* __restrict__ src = reinterpret_cast<const K*>(q.in0) + base;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (m == kSortTile) {
    // all 16 loads are issued before the first shared atomic (ncu r1b: this kernel sat on
    // long-scoreboard stalls at 22 % of HBM with the loads interleaved with the atomics)
    K keys[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) keys[k] = src[k * kThreads + threadIdx.x];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
      K key = keys[k];
      if (first) key = to_key<K>(key, dtype);
      const uint32_t d = (uint32_t)(key >> shift) & 255u;
      // a constant digit (exponent byte of uniform floats) would serialise 32 same-address
      // shared atomics: collapse the row to one add (MATCH.ANY costs ~400 cycles, ncu r1)
      if (__all_sync(0xffffffffu, d == __shfl_sync(0xffffffffu, d, 0))) {
        if (lane == 0) atomicAdd(&sh[d], 32u);
      } else {
        atomicAdd(&sh[d], 1u);
      }
    }
  } else {
    for (int i = threadIdx.x; i < m; i += kThreads) {
      K key = src[i];
      if (first) key = to_key<K>(key, dtype);
      atomicAdd(&sh[(uint32_t)(key >> shift) & 255u], 1u);
    }
  }
  __syncthreads();
  const uint32_t c = sh[threadIdx.x];
  const unsigned long long T = q.aux;
  hist[(size_t)q.s1.i + (size_t)threadIdx.x * T + tl] = c;
  if (c) atomicAdd(&totals[(size_t)r * 256 + threadIdx.x], c);
}

// grid = (256 digits, B requests)
__global__ void __launch_bounds__(kThreads)
sort_scan_kernel(const Req* __restrict__ reqs, uint32_t* __restrict__ hist, const uint32_t* __restrict__ totals) {
  __shared__ uint32_t warp_tot[kSortWarps];
  __shared__ uint32_t carry_s;
  const int d = blockIdx.x, r = blockIdx.y;
  const Req q = reqs[r];
  const unsigned long long T = q.aux;
  if (T == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // base = number of keys of this request with a smaller digit
  uint32_t v = (int)threadIdx.x < d ? totals[(size_t)r * 256 + threadIdx.x] : 0u;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
  if (lane == 0) warp_tot[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t b = 0;
    for (int w = 0; w < kSortWarps; ++w) b += warp_tot[w];
    carry_s = b;
  }
  __syncthreads();
  uint32_t* h = hist + (size_t)q.s1.i + (size_t)d * T;
  for (unsigned long long b0 = 0; b0 < T; b0 += kThreads) {
    const unsigned long long t = b0 + threadIdx.x;
    const uint32_t x = t < T ? h[t] : 0u;
    uint32_t inc = x;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, inc, s);
      if (lane >= s) inc += y;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t wprefix = 0;
    for (int w = 0; w < warp; ++w) wprefix += warp_tot[w];
    const uint32_t carry = carry_s;
    if (t < T) h[t] = carry + wprefix + inc - x;
    __syncthreads();
    if (threadIdx.x == kThreads - 1) carry_s = carry + wprefix + inc;
    __syncthreads();
  }
}

template <typename K, bool PAIRS>
__global__ void __launch_bounds__(kThreads, (sizeof(K) == 4 && !PAIRS) ? 4 : 2)
sort_scatter_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
                    int shift, int dtype, int first, int last, const uint32_t* __restrict__ hist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K* skeys = reinterpret_cast<K*>(smem_raw);
  uint32_t* sidx = reinterpret_cast<uint32_t*>(smem_raw + sizeof(K) * kSortTile);
  uint32_t* warp_hist = sidx + (PAIRS ? kSortTile : 0);  // [kSortWarps][256]
  uint32_t* digit_base = warp_hist + kSortWarps * 256;    // [256] tile-local exclusive digit offsets
  uint32_t* goff = digit_base + 256;                      // [256] global offset - digit_base
  uint32_t* wsum = goff + 256;                            // [kSortWarps]

  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const unsigned long long tl = tile - q.tile_start;
  const size_t base = (size_t)tl * kSortTile;
  const size_t rem = (size_t)q.n - base;
  const int m = rem < (size_t)kSortTile ? (int)rem : kSortTile;
  const K* __restrict__ src = reinterpret_cast<const K*>(q.in0) + base;
  const uint32_t* __restrict__ srci = q.in1 ? reinterpret_cast<const uint32_t*>(q.in1) + base : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int i = threadIdx.x; i < kSortWarps * 256; i += kThreads) warp_hist[i] = 0;
  __syncthreads();

  // striped arrangement: row k of warp w covers tile positions w*512 + k*32 + [0,32)
  K key[kSortItems];
  uint32_t idx[kSortItems];
  uint32_t rank[kSortItems];
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int p = warp * (32 * kSortItems) + k * 32 + lane;
    K x = ~K(0);
    if (p < m) {
      x = src[p];
      if (first) x = to_key<K>(x, dtype);
    }
    key[k] = x;
    if (PAIRS) idx[k] = (p < m) ? (srci ? srci[p] : (uint32_t)(base + p)) : 0u;
  }
  uint32_t* wh = warp_hist + warp * 256;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int p = warp * (32 * kSortItems) + k * 32 + lane;
    const bool valid = p < m;
    const uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
    const unsigned peers = match_digit(d, valid, m != kSortTile);
    const int leader = __ffs(peers) - 1;
    uint32_t cnt = 0;
    if (lane == leader && valid) {
      cnt = wh[d];
      wh[d] = cnt + (uint32_t)__popc(peers);
    }
    cnt = __shfl_sync(0xffffffffu, cnt, leader);
    rank[k] = cnt + (uint32_t)__popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();
  // thread d: exclusive scan of digit d over the warps; then over the digits
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = warp_hist[w * 256 + d];
      warp_hist[w * 256 + d] = run;
      run += c;
    }
    uint32_t inc = run;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, inc, s);
      if (lane >= s) inc += y;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t wprefix = 0;
    for (int w = 0; w < warp; ++w) wprefix += wsum[w];
    const uint32_t excl = wprefix + inc - run;
    digit_base[d] = excl;
    goff[d] = hist[(size_t)q.s1.i + (size_t)d * q.aux + tl] - excl;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int p = warp * (32 * kSortItems) + k * 32 + lane;
    if (p < m) {
      const uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
      const uint32_t lp = digit_base[d] + wh[d] + rank[k];
      skeys[lp] = key[k];
      if (PAIRS) sidx[lp] = idx[k];
    }
  }
  __syncthreads();
  K* __restrict__ dst = reinterpret_cast<K*>(q.out);
  for (int i = threadIdx.x; i < m; i += kThreads) {
    const K x = skeys[i];
    const uint32_t d = (uint32_t)(x >> shift) & 255u;
    const size_t gp = (size_t)(uint32_t)(goff[d] + (uint32_t)i);
    if (dst) dst[gp] = last ? from_key<K>(x, dtype) : x;
    if (PAIRS) {
      if (last) reinterpret_cast<long long*>(q.s0.i)[gp] = (long long)sidx[i];
      else reinterpret_cast<uint32_t*>(q.s0.i)[gp] = sidx[i];
    }
  }
}

template <typename K, bool PAIRS>
static size_t scatter_smem() {
  return sizeof(K) * kSortTile + (PAIRS ? sizeof(uint32_t) * kSortTile : 0) +
         sizeof(uint32_t) * (kSortWarps * 256 + 256 + 256 + kSortWarps);
}

template <typename K, bool PAIRS>
static int sort_run(int dtype, int B, const void* const* in, void* const* out, const size_t* n) {
  Ctx& c = ctx();
  const int passes = (int)sizeof(K);
  static bool attr_set = false;
  if (!attr_set) {
    BCI_CUDA(cudaFuncSetAttribute(sort_scatter_kernel<K, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)scatter_smem<K, PAIRS>()));
    attr_set = true;
  }
  // scratch layout: [totals B*256][hist sum(256*T_r)] [per request: keysA (+keysB, idxA, idxB)]
  std::vector<size_t> T((size_t)B), hoff((size_t)B), koff((size_t)B);
  size_t hist_words = 0;
  for (int i = 0; i < B; ++i) {
    if (n[i] >= (1ull << 31)) { set_error("bci_sort: request %d has >= 2^31 elements", i); return BCI_ERR_UNSUPPORTED; }
    T[i] = (n[i] + kSortTile - 1) / kSortTile;
    hoff[i] = (size_t)B * 256 + hist_words;
    hist_words += 256 * T[i];
  }
  size_t bytes = ((size_t)B * 256 + hist_words) * sizeof(uint32_t);
  bytes = (bytes + 255) & ~size_t(255);
  const size_t hdr_bytes = bytes;
  for (int i = 0; i < B; ++i) {
    koff[i] = bytes;
    const size_t kb = (n[i] * sizeof(K) + 255) & ~size_t(255);
    const size_t ib = (n[i] * sizeof(uint32_t) + 255) & ~size_t(255);
    bytes += PAIRS ? 2 * kb + 2 * ib : kb;
  }
  void* sc = nullptr;
  int rc = scratch_get(bytes, &sc);
  if (rc != BCI_OK) return rc;
  uint8_t* sbase = reinterpret_cast<uint8_t*>(sc);
  uint32_t* words = reinterpret_cast<uint32_t*>(sbase);
  uint32_t* totals = words;
  (void)hdr_bytes;

  std::vector<Req> reqs((size_t)B);
  for (int p = 0; p < passes; ++p) {
    const bool first = p == 0, last = p == passes - 1;
    for (int i = 0; i < B; ++i) {
      Req& q = reqs[i];
      memset(&q, 0, sizeof q);
      const size_t kb = (n[i] * sizeof(K) + 255) & ~size_t(255);
      const size_t ib = (n[i] * sizeof(uint32_t) + 255) & ~size_t(255);
      uint8_t* keysA = sbase + koff[i];
      q.n = n[i];
      q.s1.i = (int64_t)hoff[i];
      q.aux = T[i];
      if (!PAIRS) {
        // in -> A -> out -> A -> out ... (even number of passes ends in out)
        const void* src = first ? in[i] : ((p & 1) ? (const void*)keysA : (const void*)out[i]);
        void* dst = (p & 1) ? out[i] : (void*)keysA;
        q.in0 = src;
        q.out = dst;
      } else {
        uint8_t* keysB = keysA + kb;
        uint8_t* idxA = keysB + kb;
        uint8_t* idxB = idxA + ib;
        q.in0 = first ? in[i] : ((p & 1) ? (const void*)keysA : (const void*)keysB);
        q.in1 = first ? nullptr : ((p & 1) ? (const void*)idxA : (const void*)idxB);
        q.out = last ? nullptr : ((p & 1) ? (void*)keysB : (void*)keysA);
        q.s0.i = last ? (int64_t)(intptr_t)out[i] : (int64_t)(intptr_t)((p & 1) ? idxB : idxA);
      }
    }
    Batch bt;
    rc = upload_batch(reqs.data(), B, kSortTile, &bt);
    if (rc != BCI_OK) return rc;
    if (bt.total_tiles == 0) return BCI_OK;
    if (bt.total_tiles > 0x7fffffffull) { set_error("bci_sort: batch too large for one launch"); return BCI_ERR_INVALID; }
    BCI_CUDA(cudaMemsetAsync(totals, 0, (size_t)B * 256 * sizeof(uint32_t), c.stream));
    sort_hist_kernel<K><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(
        bt.d_reqs, bt.d_tile_start, B, p * 8, dtype, first ? 1 : 0, words, totals);
    BCI_CHECK_LAUNCH();
    sort_scan_kernel<<<dim3(256, (unsigned)B), kThreads, 0, c.stream>>>(bt.d_reqs, words, totals);
    BCI_CHECK_LAUNCH();
    sort_scatter_kernel<K, PAIRS><<<(unsigned)bt.total_tiles, kThreads, scatter_smem<K, PAIRS>(), c.stream>>>(
        bt.d_reqs, bt.d_tile_start, B, p * 8, dtype, first ? 1 : 0, last ? 1 : 0, words);
    BCI_CHECK_LAUNCH();
  }
  return BCI_OK;
}

static int sort_common(bool pairs, int dtype, int B, const void* const* in, void* const* out, const size_t* n) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !in || !out || !n) { set_error("bci_sort: bad arguments"); return BCI_ERR_INVALID; }
  if (elem_size(dtype) == 0) { set_error("bci_sort: unknown dtype %d", dtype); return BCI_ERR_INVALID; }
  if (B > 65535) { set_error("bci_sort: at most 65535 requests per launch"); return BCI_ERR_INVALID; }
  for (int i = 0; i < B; ++i)
    if (n[i] && (!in[i] || !out[i])) { set_error("bci_sort: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  // default: one-sweep passes with decoupled look-back (bci_sort_onesweep.cu); the
  // reduce-then-scan passes below stay as the fallback for n >= 2^30 and for A/B runs
  const char* force = tuning_env("BCI_SORT");
  bool legacy = force && strcmp(force, "legacy") == 0;
  for (int i = 0; i < B && !legacy; ++i) legacy = n[i] >= (1ull << 30);
  if (!legacy) return sort_onesweep(pairs, dtype, B, in, out, n);
  const bool wide = elem_size(dtype) == 8;
  if (!pairs) return wide ? sort_run<unsigned long long, false>(dtype, B, in, out, n)
                          : sort_run<uint32_t, false>(dtype, B, in, out, n);
  return wide ? sort_run<unsigned long long, true>(dtype, B, in, out, n)
              : sort_run<uint32_t, true>(dtype, B, in, out, n);
}

}  // namespace bci

using namespace bci;

extern "C" int bci_sort(int dtype, int B, const void* const* in, void* const* out, const size_t* n) {
  return sort_common(false, dtype, B, in, out, n);
}

extern "C" int bci_argsort(int dtype, int B, const void* const* in, int64_t* const* out_idx, const size_t* n) {
  return sort_common(true, dtype, B, in, reinterpret_cast<void* const*>(out_idx), n);
}
