// bci_sort_common.cuh — key transforms and warp-level digit matching shared by the two radix
// sort implementations (bci_sort.cu: reduce-then-scan passes; bci_sort_onesweep.cu: one-sweep
// passes with decoupled look-back).
#pragma once
#include "bci_common.cuh"

namespace bci {

constexpr int kSortItems = 16;
constexpr int kSortTile = kThreads * kSortItems;  // 4096 keys per CTA tile
constexpr int kSortWarps = kThreads / 32;

// Order-preserving map to unsigned keys (and back):
//   float : flip all bits of negatives, set the sign bit of non-negatives; every NaN becomes the
//           all-ones key so NaNs sort last like numpy (payload canonicalised to +qNaN; -0.0 sorts
//           before +0.0, value-equal to numpy's result);
//   int   : flip the sign bit.
template <typename K> __device__ __forceinline__ K to_key(K bits, int dtype);
template <> __device__ __forceinline__ uint32_t to_key<uint32_t>(uint32_t b, int dtype) {
  if (dtype == BCI_F32) {
    if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;  // NaN last
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  }
  return b ^ 0x80000000u;
}
template <> __device__ __forceinline__ unsigned long long to_key<unsigned long long>(unsigned long long b, int dtype) {
  if (dtype == BCI_F64) {
    if ((b & 0x7fffffffffffffffull) > 0x7ff0000000000000ull) return ~0ull;
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
  }
  return b ^ 0x8000000000000000ull;
}
template <typename K> __device__ __forceinline__ K from_key(K u, int dtype);
template <> __device__ __forceinline__ uint32_t from_key<uint32_t>(uint32_t u, int dtype) {
  if (dtype == BCI_F32) return (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return u ^ 0x80000000u;
}
template <> __device__ __forceinline__ unsigned long long from_key<unsigned long long>(unsigned long long u, int dtype) {
  if (dtype == BCI_F64) return (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  return u ^ 0x8000000000000000ull;
}

// Lanes of the warp holding the same digit as this lane (and the same validity).  Eight
// ballots, one per digit bit, each folded in with a single LOP3 — the MATCH.ANY instruction
// does the same job but measured ~400 cycles per warp (ncu, profiles/r1b), several times this
// loop.  A row whose 32 digits are all equal (exponent byte of uniform floats) short-circuits.
__device__ __forceinline__ unsigned match_digit(uint32_t d, bool valid, bool ragged_tile) {
  unsigned peers = 0xffffffffu;
  if (!__all_sync(0xffffffffu, d == __shfl_sync(0xffffffffu, d, 0))) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const bool bit = (d >> b) & 1u;
      const unsigned m = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? m : ~m;
    }
  }
  if (ragged_tile) {  // only the last tile of a request has lanes past the end
    const unsigned v = __ballot_sync(0xffffffffu, valid);
    peers &= valid ? v : ~v;
  }
  return peers;
}

// bci_sort_onesweep.cu
int sort_onesweep(bool pairs, int dtype, int B, const void* const* in, void* const* out, const size_t* n);

}  // namespace bci
