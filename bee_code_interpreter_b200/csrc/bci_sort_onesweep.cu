// bci_sort_onesweep.cu — K5 radix sort, one-sweep form (Adinets & Merrill, "Onesweep: A Faster
// Least Significant Digit Radix Sort for GPUs", 2022), segmented by request.
//
//   digit_hist_kernel : ONE read of the keys counts every 8-bit digit of every pass
//                       (4 histograms for 32-bit keys, 8 for 64-bit) per request;
//   digit_scan_kernel : exclusive scan over the 256 digits -> where each digit's run starts; also
//                       classifies each (request, pass): constant digit / skewed / ordinary;
//   onesweep_kernel   : per pass, a tile (512 or 256 threads x 20 keys) ranks its keys stably in
//                       half-warp units (one shared-memory word per unit and digit holds the row's
//                       peer mask and the running count; skewed passes use ballots over the digit
//                       bits that differ), publishes its 256 digit counts, places the keys in digit
//                       order in shared memory, obtains the counts of all earlier tiles of the SAME
//                       request by a CTA-wide decoupled look-back (status word = 2 flag bits + 30-bit
//                       count; warp w fetches the status row of the tile w+1 back), and writes every
//                       digit run with coalesced stores.  Passes whose digit is constant for every
//                       request of the batch are not launched at all.
// HBM traffic = 1 + 2 per pass key sweeps: 9*n*4 bytes for float32 = the algorithmic figure of
// SURVEY.md §8a K5 (the reduce-then-scan form in bci_sort.cu moves 12).  Tiles take their
// index from an atomic counter, so a tile only ever waits for tiles that already started:
// no deadlock; spins are bounded and trap instead of hanging the GPU.  Bit-exact with numpy.
#include "bci_sort_common.cuh"
#include <algorithm>
#include <vector>
#include <cstring>
#include <cstdlib>
#include <type_traits>

namespace bci {
namespace {

constexpr uint32_t kPartial = 1u << 30, kInclusive = 2u << 30, kFlagMask = 3u << 30, kValMask = (1u << 30) - 1u;

// A status word carries its own payload (flag bits + count), so no ordering against other
// memory is needed: cache-global relaxed accesses.  (ld.acquire.gpu compiles to CCTL.IVALL + LD
// and cost 20 % of the kernel's stall samples in the first version, ncu r1c.)
__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint4 ld_status4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// Tile shape of a sweep.  32-bit keys alone: 512 threads x 20 keys = 10240 keys per CTA, 2 CTAs per SM — the same 32
// resident warps as 4 CTAs of 256, but half the tiles: half the status words to publish and read back, half the
// look-back depth (the depth is set by how many tiles finish ranking per L2 round trip), half the per-tile prologues
// (ncu r2r: look-back was 28 % of the instructions and 25 % of the stall samples of the 4096-key form).  Measured,
// 32 x 1e7 float32: 4096-key tiles 8.28 ms, 8192 keys 7.67 ms, 16384 keys (1024 threads, 1 CTA per SM) 8.59 ms.
template <typename K, bool PAIRS> struct Sweep {
  static constexpr int T = (sizeof(K) == 4 && !PAIRS) ? 512 : kThreads;
  static constexpr int W = T / 32;
  // keys per thread: 20 (16 for the 64-bit pairs, whose staging buffers would otherwise cost the third CTA per SM).
  // Measured on 32 x 1e7 float32: 16 keys 6.84 ms, 20 keys 6.35 ms, 24 keys 6.28 ms but 7 % slower on 256 x 1e5 (the
  // longer ragged tail per request); stable argsort float32 8.98 -> 8.56 ms, int64 8.54 -> 8.39 ms with 20.
  static constexpr int ITEMS = (sizeof(K) == 8 && PAIRS) ? 16 : 20;
  static constexpr int TILE = T * ITEMS;
  static constexpr int CTAS = T == 512 ? 2 : 3;
};

// ---- all digit histograms in one read ---------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(kThreads)
digit_hist_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
                  unsigned long long total_tiles, int dtype, uint32_t* __restrict__ ghist) {
  constexpr int P = (int)sizeof(K);
  __shared__ uint32_t sh[P][256];
  const unsigned long long per = (total_tiles + gridDim.x - 1) / gridDim.x;
  const unsigned long long t0 = (unsigned long long)blockIdx.x * per;
  const unsigned long long t1 = t0 + per < total_tiles ? t0 + per : total_tiles;
  if (t0 >= t1) return;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int p = 0; p < P; ++p) sh[p][threadIdx.x] = 0;
  __syncthreads();
  int r = find_req(tile_start, B, t0);
  auto flush = [&](int req) {
    __syncthreads();
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const uint32_t v = sh[p][threadIdx.x];
      if (v) atomicAdd(&ghist[((size_t)req * P + p) * 256 + threadIdx.x], v);
      sh[p][threadIdx.x] = 0;
    }
    __syncthreads();
  };
  for (unsigned long long t = t0; t < t1; ++t) {
    if (t >= __ldg(tile_start + r + 1)) {
      flush(r);
      while (t >= __ldg(tile_start + r + 1)) ++r;
    }
    const Req q = reqs[r];
    const size_t base = (size_t)(t - q.tile_start) * kSortTile;
    const size_t rem = (size_t)q.n - base;
    const int m = rem < (size_t)kSortTile ? (int)rem : kSortTile;
    const K* __restrict__ src = reinterpret_cast<const K*>(q.in0) + base;
    if (m == kSortTile) {
      K keys[kSortItems];
#pragma unroll
      for (int k = 0; k < kSortItems; ++k) keys[k] = src[k * kThreads + threadIdx.x];
#pragma unroll
      for (int k = 0; k < kSortItems; ++k) {
        const K key = to_key<K>(keys[k], dtype);
        // Constant digits (exponent byte, high bytes of small integers) would serialise 32 same-address shared
        // atomics: one add per uniform row instead.  Which bytes are uniform across the row comes from ONE warp
        // reduction per 32 key bits (OR of key ^ lane 0's key) instead of a shuffle + vote per digit.
        const K x = key ^ __shfl_sync(0xffffffffu, key, 0);
        uint32_t diff_lo = __reduce_or_sync(0xffffffffu, (uint32_t)x), diff_hi = 0;
        if constexpr (sizeof(K) == 8) diff_hi = __reduce_or_sync(0xffffffffu, (uint32_t)((unsigned long long)x >> 32));
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const uint32_t d = (uint32_t)(key >> (8 * p)) & 255u;
          const uint32_t diff = ((p < 4 ? diff_lo : diff_hi) >> (8 * (p & 3))) & 255u;
          if (diff == 0) {  // warp-uniform
            if (lane == 0) atomicAdd(&sh[p][d], 32u);
          } else {
            atomicAdd(&sh[p][d], 1u);
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < m; i += kThreads) {
        const K key = to_key<K>(src[i], dtype);
#pragma unroll
        for (int p = 0; p < P; ++p) atomicAdd(&sh[p][(uint32_t)(key >> (8 * p)) & 255u], 1u);
      }
    }
  }
  flush(r);
}

// one CTA per (request, pass): ghist[..][256] counts -> exclusive offsets, in place
// Also: trivial[request * P + pass] = 1 when ONE digit holds every key of the request in that pass (the high bytes of
// small integers, the sign/exponent byte of same-magnitude floats): such a pass is the identity permutation and the
// sweep kernel copies its tiles instead of ranking them.
__global__ void __launch_bounds__(kThreads)
digit_scan_kernel(uint32_t* __restrict__ ghist, uint32_t* __restrict__ trivial) {
  __shared__ uint32_t wsum[kSortWarps];
  uint32_t* h = ghist + (size_t)blockIdx.x * 256;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t c = h[threadIdx.x];
  uint32_t inc = c;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, s);
    if (lane >= s) inc += y;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t wprefix = 0, total = 0;
  for (int w = 0; w < kSortWarps; ++w) { if (w < warp) wprefix += wsum[w]; total += wsum[w]; }
  h[threadIdx.x] = wprefix + inc - c;
  const int one_digit = __syncthreads_or(total != 0 && c == total);
  // 2 = skewed: one digit holds more than a quarter of the keys (sign/exponent byte of floats, a byte with a handful of
  // values in small integers): the sweep then matches peers with ballots over the few digit bits that differ in a row
  // instead of shared-memory atomics that would pile up on one or two words
  const int skewed = __syncthreads_or(total != 0 && c > total / 4);
  if (threadIdx.x == 0) trivial[blockIdx.x] = one_digit ? 1u : (skewed ? 2u : 0u);
}

// ---- one pass -----------------------------------------------------------------------------------
// Req fields: in0 = source keys, in1 = source indices (u32; nullptr on the first pass: generated),
// out = destination keys (nullptr when only indices leave the last pass), s0.i = destination
// indices (u32*, or int64* on the last pass).
// FIRST / LAST pass and full / ragged tile are compile-time: the first version carried them as
// runtime flags and spent 40 of its 145 instructions per key on predicates and branches (ncu r1c).
template <typename K, bool PAIRS, bool FIRST, bool LAST>
__global__ void __launch_bounds__((Sweep<K, PAIRS>::T), (Sweep<K, PAIRS>::CTAS))
onesweep_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B,
                int pass, int dtype, const uint32_t* __restrict__ gbase, const uint32_t* __restrict__ trivial,
                uint32_t* __restrict__ status, uint32_t* __restrict__ counter, unsigned int prefetch_tiles) {
  constexpr int P = (int)sizeof(K);
  constexpr int T = Sweep<K, PAIRS>::T, W = Sweep<K, PAIRS>::W, TILE = Sweep<K, PAIRS>::TILE, ITEMS = Sweep<K, PAIRS>::ITEMS;
  constexpr bool WRITE_KEYS = !(PAIRS && LAST);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K* skeys = reinterpret_cast<K*>(smem_raw);
  uint32_t* sidx = reinterpret_cast<uint32_t*>(smem_raw + sizeof(K) * TILE);
  uint32_t* tab = sidx + (PAIRS ? TILE : 0);  // [2W][256] one table per half-warp: peer mask (low 16 bits) | count << 16,
                                              // later the tile-local offset of the (half-warp, digit) run
  uint32_t* goff = tab + 2 * W * 256;         // [256] global offset - tile-local offset
  uint32_t* lb = goff + 256;                  // [W][256] status words of W earlier tiles (look-back)
  uint32_t* wsum = lb + W * 256;              // [W] (+1: dynamic tile id)

  // tiles are taken in the order CTAs start: whoever a tile waits for is already running
  if (threadIdx.x == 0) wsum[W] = atomicAdd(counter, 1u);
  {
    // 4 x STS.128 per thread
    static_assert((2 * W * 256) % (4 * T) == 0, "vector clear");
    uint4* t4 = reinterpret_cast<uint4*>(tab);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < 2 * W * 256 / (4 * T); ++i) t4[i * T + threadIdx.x] = z;
  }
  __syncthreads();
  const unsigned long long tile = wsum[W];
  const int r = find_req_block(tile_start, B, tile);
  const Req q = reqs[r];
  const unsigned long long tl = tile - q.tile_start;
  const size_t base = (size_t)tl * TILE;
  const size_t rem = (size_t)q.n - base;
  const int m = rem < (size_t)TILE ? (int)rem : TILE;
  const int shift = pass * 8;
  const K* __restrict__ src = reinterpret_cast<const K*>(q.in0) + base;
  const uint32_t* __restrict__ srci = reinterpret_cast<const uint32_t*>(q.in1) + base;  // !FIRST only
  const uint32_t pass_kind = __ldg(trivial + (size_t)r * P + pass);
  if (pass_kind == 1u) {
    // every key of this request carries the same digit in this pass: nothing moves.  Copy the tile (with the key
    // transform of a first / last pass) — ~10 instructions per key instead of a ranking pass; no status is published,
    // the other tiles of this request take the same branch.
    K* __restrict__ dstk = reinterpret_cast<K*>(q.out) + base;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int i = k * T + (int)threadIdx.x;
      if (i < m) {
        K x = src[i];
        if constexpr (FIRST) x = to_key<K>(x, dtype);
        if constexpr (WRITE_KEYS) dstk[i] = LAST ? from_key<K>(x, dtype) : x;
        if constexpr (PAIRS) {
          uint32_t ix;
          if constexpr (FIRST) ix = (uint32_t)(base + i); else ix = srci[i];
          if constexpr (LAST) reinterpret_cast<long long*>(q.s0.i)[base + i] = (long long)ix;
          else reinterpret_cast<uint32_t*>(q.s0.i)[base + i] = ix;
        }
      }
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // this request's run start for digit d of this pass (independent load, issued early)
  const uint32_t run_start = __ldg(gbase + ((size_t)r * P + pass) * 256 + (threadIdx.x & 255));
  // L2 prefetch of the tile `prefetch_tiles` further on in the same request (whoever takes it):
  // the key loads of a tile are then L2 hits instead of ~1 us HBM round trips under load
  if (prefetch_tiles != 0 && threadIdx.x == 32) {
    const size_t b2 = base + (size_t)prefetch_tiles * TILE;
    if (b2 < (size_t)q.n) {
      const size_t cnt = (size_t)q.n - b2 < (size_t)TILE ? (size_t)q.n - b2 : (size_t)TILE;
      const K* a2 = reinterpret_cast<const K*>(q.in0) + b2;
      const uint32_t bytes = (uint32_t)(cnt * sizeof(K)) & ~15u;
      if (bytes != 0 && (reinterpret_cast<uintptr_t>(a2) & 15) == 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a2), "r"(bytes) : "memory");
      if constexpr (PAIRS && !FIRST) {
        const uint32_t* i2 = reinterpret_cast<const uint32_t*>(q.in1) + b2;
        const uint32_t ibytes = (uint32_t)(cnt * 4) & ~15u;
        if (ibytes != 0 && (reinterpret_cast<uintptr_t>(i2) & 15) == 0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(i2), "r"(ibytes) : "memory");
      }
    }
  }
  // A ranking unit is a HALF-warp: its 16 lanes hold 16 rows of 16 consecutive keys (256 keys in tile order), and its
  // table word for a digit carries the row's peer mask in the low half and the running count in the high half, so a
  // row costs three shared-memory operations on ONE word (OR, read, leader's write of count+popc with the mask
  // cleared) where the warp-wide form with separate mask and counter arrays took five — and the digit-indexed
  // accesses are what saturates the shared-memory pipe (3.3-way bank conflicts each, ncu r2t).
  const int l16 = lane & 15;
  uint32_t* wh = tab + (threadIdx.x >> 4) * 256;
  const int p0 = (threadIdx.x >> 4) * (16 * ITEMS) + l16;  // row k of unit u = positions u*256 + k*16 + [0,16)

  auto body = [&](auto full_c) {
    constexpr bool FULL = decltype(full_c)::value;
    K key[ITEMS];
    uint32_t idx[ITEMS];
    uint32_t rank[ITEMS];
    // every load of the tile is issued before anything consumes one
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) key[k] = (FULL || p0 + k * 16 < m) ? src[p0 + k * 16] : K(0);
    if constexpr (PAIRS) {
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        if constexpr (FIRST) idx[k] = (uint32_t)(base + p0 + k * 16);
        else idx[k] = (FULL || p0 + k * 16 < m) ? srci[p0 + k * 16] : 0u;
      }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      if constexpr (FIRST) key[k] = to_key<K>(key[k], dtype);
      if constexpr (!FULL) key[k] = (p0 + k * 16 < m) ? key[k] : ~K(0);
    }
    {
      // Peers of a row = lanes of the half-warp with the same digit.  Each lane ORs its bit into the unit's word for
      // its digit and reads the word back; the leader (lowest peer lane) writes the new count with the mask cleared,
      // so the table is counts-only again when the tile is done.  Lanes past the end of a ragged tile sit out.
      const unsigned my_bit = 1u << l16;
      const unsigned lt16 = my_bit - 1u;
      if (FULL && pass_kind == 2u) {
        // skewed pass: in most rows only two or three digit BITS differ across the warp (one warp OR-reduction tells
        // which); peers then come from one ballot per differing bit and nobody touches the table with an atomic
        const unsigned half_mask = 0xffffu << (lane & 16);
        const unsigned below = (1u << lane) - 1u;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
          const uint32_t dg = (uint32_t)(key[k] >> shift) & 255u;
          uint32_t diff = __reduce_or_sync(0xffffffffu, dg ^ __shfl_sync(0xffffffffu, dg, 0));
          if (__popc(diff) <= 4) {
            unsigned peers = half_mask;
            while (diff) {
              const uint32_t bit = diff & (0u - diff);
              diff ^= bit;
              const unsigned mb = __ballot_sync(0xffffffffu, (dg & bit) != 0);
              peers &= (dg & bit) ? mb : ~mb;
            }
            const uint32_t cnt = wh[dg] >> 16;
            __syncwarp();  // every peer has read the count before the leader rewrites it
            if (lane == __ffs(peers) - 1) wh[dg] = (cnt + (uint32_t)__popc(peers)) << 16;
            rank[k] = cnt + (uint32_t)__popc(peers & below);
            __syncwarp();
          } else {
            atomicOr(&wh[dg], my_bit);
            __syncwarp();
            const uint32_t w = wh[dg];
            __syncwarp();
            const uint32_t peers = w & 0xffffu, cnt = w >> 16;
            if (my_bit == (peers & (0u - peers))) wh[dg] = (cnt + (uint32_t)__popc(peers)) << 16;
            rank[k] = cnt + (uint32_t)__popc(peers & lt16);
            __syncwarp();
          }
        }
      } else
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const uint32_t dg = (uint32_t)(key[k] >> shift) & 255u;
        const bool valid = FULL || p0 + k * 16 < m;
        if (valid) atomicOr(&wh[dg], my_bit);
        __syncwarp();
        const uint32_t w = wh[dg];
        __syncwarp();  // every peer has read the word before the leader rewrites it
        const uint32_t peers = w & 0xffffu, cnt = w >> 16;
        if (valid && my_bit == (peers & (0u - peers))) wh[dg] = (cnt + (uint32_t)__popc(peers)) << 16;
        rank[k] = cnt + (uint32_t)__popc(peers & lt16);
        __syncwarp();  // row k's count is visible to row k+1
      }
    }
    __syncthreads();
    // Thread d < 256 owns digit d.  Order of the rest of the tile: (1) counts over warps, publish them, tile-local
    // digit offsets; (2) every thread places its keys in digit order in shared memory — after which the 32 key and
    // rank registers are free; (3) look back for the counts of the earlier tiles; (4) write the digit runs out.
    const int d = threadIdx.x;
    const bool owner = T == 256 || d < 256;
    uint32_t* my_status = status + (size_t)tile * 256 + d;
    uint32_t run = 0, local = 0;
    {
      if (owner) {
#pragma unroll
        for (int u = 0; u < 2 * W; ++u) {
          const uint32_t c = tab[u * 256 + d] >> 16;
          tab[u * 256 + d] = run;
          run += c;
        }
        st_status(my_status, (tl == 0 ? kInclusive : kPartial) | run);
      }
      uint32_t inc = run;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += y;
      }
      if (lane == 31 && owner) wsum[warp] = inc;
      __syncthreads();
      if (owner) {
        uint32_t wprefix = 0;
        for (int w = 0; w < warp; ++w) wprefix += wsum[w];
        local = wprefix + inc - run;  // tile-local start of digit d
        // fold it into the per-warp prefixes: one shared load per key in the placement below
#pragma unroll
        for (int u = 0; u < 2 * W; ++u) tab[u * 256 + d] += local;
      }
    }
    __syncthreads();
    // Look-back by the whole CTA: warp w fetches all 256 status words of the tile (round * W + w + 1) back with two
    // 16-byte loads per lane and parks them in shared memory; digit d's owner then walks the W words of its digit in
    // order.  One L2 round trip covers W earlier tiles (the per-thread form took a round trip per 4 tiles and 85
    // instructions per round, ncu r2r).  Words beyond the first INCLUSIVE one are ignored; a word that is not
    // published yet is polled by its owner.  The first round's loads are in flight during the placement.
    const uint32_t tl32 = (uint32_t)tl;
    auto fetch = [&](uint32_t back0, uint4& v0, uint4& v1) {
      const uint32_t b = back0 + (uint32_t)warp;
      v0 = make_uint4(kInclusive, kInclusive, kInclusive, kInclusive);
      v1 = v0;
      if (b <= tl32) {
        const uint32_t* ps = status + ((size_t)tile - b) * 256 + 4 * lane;
        v0 = ld_status4(ps);
        v1 = ld_status4(ps + 128);
      }
    };
    uint4 v0, v1;
    if (tl32 > 0) fetch(1u, v0, v1);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      if (FULL || p0 + k * 16 < m) {
        const uint32_t dg = (uint32_t)(key[k] >> shift) & 255u;
        const uint32_t lp = wh[dg] + rank[k];
        skeys[lp] = key[k];
        if constexpr (PAIRS) sidx[lp] = idx[k];
      }
    }
    if (tl32 > 0) {
      uint32_t excl = 0;
      bool done = !owner;
      for (uint32_t back0 = 1;;) {
        *reinterpret_cast<uint4*>(lb + warp * 256 + 4 * lane) = v0;
        *reinterpret_cast<uint4*>(lb + warp * 256 + 128 + 4 * lane) = v1;
        __syncthreads();
        if (!done) {
#pragma unroll 4
          for (int j = 0; j < W; ++j) {
            uint32_t x = lb[j * 256 + d];
            if ((x & kFlagMask) == 0) {
              const uint32_t* ps = status + ((size_t)tile - (back0 + (uint32_t)j)) * 256 + d;
              for (unsigned spins = 0; (x & kFlagMask) == 0; ++spins) {
                if (spins > (1u << 24)) __trap();  // a protocol bug must not hang the GPU
                x = ld_status(ps);
              }
            }
            excl += x & kValMask;
            if (x & kInclusive) { done = true; break; }
          }
          if (done) {
            st_status(my_status, kInclusive | (excl + run));
            goff[d] = run_start + excl - local;
          }
        }
        if (__syncthreads_and(done)) break;  // also: goff and the staged keys are visible
        back0 += W;
        fetch(back0, v0, v1);
      }
    } else {
      if (owner) goff[d] = run_start - local;
      __syncthreads();
    }
    K* __restrict__ dst = reinterpret_cast<K*>(q.out);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int i = k * T + (int)threadIdx.x;
      if (FULL || i < m) {
        const K x = skeys[i];
        const uint32_t d = (uint32_t)(x >> shift) & 255u;
        const size_t gp = (size_t)(uint32_t)(goff[d] + (uint32_t)i);
        if constexpr (WRITE_KEYS) dst[gp] = LAST ? from_key<K>(x, dtype) : x;
        if constexpr (PAIRS) {
          if constexpr (LAST) reinterpret_cast<long long*>(q.s0.i)[gp] = (long long)sidx[i];
          else reinterpret_cast<uint32_t*>(q.s0.i)[gp] = sidx[i];
        }
      }
    }
  };
  if (m == TILE) body(std::true_type{});
  else body(std::false_type{});
}

template <typename K, bool PAIRS>
size_t onesweep_smem() {
  constexpr int W = Sweep<K, PAIRS>::W, TILE = Sweep<K, PAIRS>::TILE;
  return sizeof(K) * TILE + (PAIRS ? sizeof(uint32_t) * TILE : 0) + sizeof(uint32_t) * (W * 256 + 256 + W + 1 + W * 512);
}

template <typename K, bool PAIRS>
int onesweep_run(int dtype, int B, const void* const* in, void* const* out, const size_t* n) {
  Ctx& c = ctx();
  constexpr int P = (int)sizeof(K);
  static bool attr_set = false;
  if (!attr_set) {
    const int smem = (int)onesweep_smem<K, PAIRS>();
    BCI_CUDA(cudaFuncSetAttribute(onesweep_kernel<K, PAIRS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    BCI_CUDA(cudaFuncSetAttribute(onesweep_kernel<K, PAIRS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    BCI_CUDA(cudaFuncSetAttribute(onesweep_kernel<K, PAIRS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    BCI_CUDA(cudaFuncSetAttribute(onesweep_kernel<K, PAIRS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  constexpr int T = Sweep<K, PAIRS>::T, TILE = Sweep<K, PAIRS>::TILE;
  size_t total_tiles = 0, hist_tiles = 0;  // sweep tiles (TILE keys) and histogram tiles (kSortTile keys)
  for (int i = 0; i < B; ++i) {
    total_tiles += (n[i] + TILE - 1) / TILE;
    hist_tiles += (n[i] + kSortTile - 1) / kSortTile;
  }
  if (total_tiles == 0) return BCI_OK;
  if (total_tiles > 0x7fffffffull) { set_error("bci_sort: batch too large for one launch"); return BCI_ERR_INVALID; }
  // scratch: [counters P][ghist B*P*256][status total_tiles*256] [per request: keysA (+keysB, idxA, idxB)]
  const size_t flag_words = ((size_t)B * P + 63) & ~size_t(63);
  const size_t hdr_words = 64 + (size_t)B * P * 256 + flag_words;
  size_t bytes = (hdr_words + total_tiles * 256) * sizeof(uint32_t);
  bytes = (bytes + 255) & ~size_t(255);
  std::vector<size_t> koff((size_t)B);
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
  uint32_t* counters = reinterpret_cast<uint32_t*>(sbase);
  uint32_t* ghist = counters + 64;
  uint32_t* trivial = ghist + (size_t)B * P * 256;
  uint32_t* status = reinterpret_cast<uint32_t*>(sbase) + hdr_words;
  BCI_CUDA(cudaMemsetAsync(sbase, 0, hdr_words * sizeof(uint32_t), c.stream));

  std::vector<Req> reqs((size_t)B);
  // upfront histogram over the ORIGINAL keys
  for (int i = 0; i < B; ++i) {
    memset(&reqs[i], 0, sizeof(Req));
    reqs[i].in0 = in[i];
    reqs[i].n = n[i];
  }
  Batch bt;
  rc = upload_batch(reqs.data(), B, kSortTile, &bt);
  if (rc != BCI_OK) return rc;
  {
    const unsigned grid = (unsigned)(hist_tiles < (size_t)c.sm_count * 8 ? hist_tiles : (size_t)c.sm_count * 8);
    digit_hist_kernel<K><<<grid, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, hist_tiles, dtype, ghist);
    BCI_CHECK_LAUNCH();
    digit_scan_kernel<<<(unsigned)(B * P), kThreads, 0, c.stream>>>(ghist, trivial);
    BCI_CHECK_LAUNCH();
  }
  // Passes in which EVERY request of the batch has one digit only (the high bytes of small integers, for instance) move
  // nothing: they are left out altogether — int64 keys below 2^31 take 4 sweeps instead of 8.  The flags come back
  // with one small synchronous copy, so this is only done for batches big enough to pay for it.  The key transforms
  // belong to the first and the last sweep that RUN, whichever digits they are; a batch whose passes are all constant
  // keeps one sweep (the copy).  Mixed batches keep the in-kernel copy path.
  std::vector<int> passes;
  {
    size_t total_keys = 0;
    for (int i = 0; i < B; ++i) total_keys += n[i];
    std::vector<uint32_t> hflags;
    if (total_keys >= (1u << 20)) {
      hflags.resize((size_t)B * P);
      BCI_CUDA(cudaMemcpyAsync(hflags.data(), trivial, hflags.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
      BCI_CUDA(cudaStreamSynchronize(c.stream));
    }
    for (int p = 0; p < P; ++p) {
      bool skip = !hflags.empty();
      for (int i = 0; i < B && skip; ++i) skip = n[i] == 0 || hflags[(size_t)i * P + p] == 1u;
      if (!skip) passes.push_back(p);
    }
    if (passes.empty()) passes.push_back(0);
  }
  if (!PAIRS && (passes.size() & 1)) {
    // an odd number of sweeps makes the first one write `out` directly, which an in-place sort (in == out) cannot take:
    // put one of the left-out passes back (the kernel copies it through)
    bool in_place = false;
    for (int i = 0; i < B; ++i) in_place = in_place || in[i] == out[i];
    if (in_place) {
      for (int p = 0; p < P; ++p)
        if (std::find(passes.begin(), passes.end(), p) == passes.end()) {
          passes.insert(std::lower_bound(passes.begin(), passes.end(), p), p);
          break;
        }
    }
  }
  const int S = (int)passes.size();
  for (int j = 0; j < S; ++j) {
    const int p = passes[j];
    const bool first = j == 0, last = j == S - 1;
    for (int i = 0; i < B; ++i) {
      Req& q = reqs[i];
      memset(&q, 0, sizeof q);
      const size_t kb = (n[i] * sizeof(K) + 255) & ~size_t(255);
      const size_t ib = (n[i] * sizeof(uint32_t) + 255) & ~size_t(255);
      uint8_t* keysA = sbase + koff[i];
      q.n = n[i];
      if (!PAIRS) {
        // in -> ... -> A -> out: the buffers alternate so that the LAST sweep writes `out`
        auto dst_of = [&](int jj) -> void* { return ((S - 1 - jj) & 1) ? (void*)keysA : out[i]; };
        q.in0 = first ? in[i] : (const void*)dst_of(j - 1);
        q.out = dst_of(j);
      } else {
        uint8_t* keysB = keysA + kb;
        uint8_t* idxA = keysB + kb;
        uint8_t* idxB = idxA + ib;
        q.in0 = first ? in[i] : ((j & 1) ? (const void*)keysA : (const void*)keysB);
        q.in1 = first ? nullptr : ((j & 1) ? (const void*)idxA : (const void*)idxB);
        q.out = last ? nullptr : ((j & 1) ? (void*)keysB : (void*)keysA);
        q.s0.i = last ? (int64_t)(intptr_t)out[i] : (int64_t)(intptr_t)((j & 1) ? idxB : idxA);
      }
    }
    rc = upload_batch(reqs.data(), B, TILE, &bt);
    if (rc != BCI_OK) return rc;
    BCI_CUDA(cudaMemsetAsync(status, 0, total_tiles * 256 * sizeof(uint32_t), c.stream));
    const unsigned grid = (unsigned)total_tiles;
    const size_t smem = onesweep_smem<K, PAIRS>();
    // L2 prefetch distance: about half a wave of resident CTAs ahead.  Measured with 20 keys per thread, f32 32 x 1e7 /
    // i64 16 x 1e7 / argsort f32: off 6.42 / 8.91 / 8.98 ms, 120 tiles 6.30 / 8.23 / 8.44, 240 6.35 / 8.30 / 8.49,
    // 480 6.38 / 8.41 / 8.57, 960 = off
    unsigned int prefetch_tiles = (unsigned)(c.sm_count * Sweep<K, PAIRS>::CTAS / 2);
    if (const char* pf = tuning_env("BCI_SORT_PREFETCH")) prefetch_tiles = (unsigned int)atoi(pf);
    if (first && last)
      onesweep_kernel<K, PAIRS, true, true><<<grid, T, smem, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, p, dtype, ghist, trivial, status, counters + p, prefetch_tiles);
    else if (first)
      onesweep_kernel<K, PAIRS, true, false><<<grid, T, smem, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, p, dtype, ghist, trivial, status, counters + p, prefetch_tiles);
    else if (last)
      onesweep_kernel<K, PAIRS, false, true><<<grid, T, smem, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, p, dtype, ghist, trivial, status, counters + p, prefetch_tiles);
    else
      onesweep_kernel<K, PAIRS, false, false><<<grid, T, smem, c.stream>>>(bt.d_reqs, bt.d_tile_start, B, p, dtype, ghist, trivial, status, counters + p, prefetch_tiles);
    BCI_CHECK_LAUNCH();
  }
  return BCI_OK;
}

}  // namespace

int sort_onesweep(bool pairs, int dtype, int B, const void* const* in, void* const* out, const size_t* n) {
  const bool wide = elem_size(dtype) == 8;
  if (!pairs) return wide ? onesweep_run<unsigned long long, false>(dtype, B, in, out, n)
                          : onesweep_run<uint32_t, false>(dtype, B, in, out, n);
  return wide ? onesweep_run<unsigned long long, true>(dtype, B, in, out, n)
              : onesweep_run<uint32_t, true>(dtype, B, in, out, n);
}

}  // namespace bci
