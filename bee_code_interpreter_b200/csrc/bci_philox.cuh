// bci_philox.cuh — Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14) and numpy's 53-bit
// double construction, shared by the K8 generator kernels (bci_rand.cu) and by the reductions
// that generate their operand in registers (bci_reduce.cu, `sum(square(rand(n)))` of
// examples/benchmark-numpy.py:20-21 without the 0.8 GB array ever reaching HBM).
#pragma once
#include "bci_common.cuh"

namespace bci {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

struct Philox4 { uint32_t x, y, z, w; };

// The key schedule (k + r*W) is the same for every counter of a request: it is
// expanded once per thread and kept in registers, leaving 2 IMAD.WIDE + 2 LOP3 per round.
struct PhiloxKeys { uint32_t a[10], b[10]; };

__device__ __forceinline__ PhiloxKeys philox_keys(uint32_t k0, uint32_t k1) {
  PhiloxKeys k;
#pragma unroll
  for (int r = 0; r < 10; ++r) { k.a[r] = k0 + (uint32_t)r * kPhiloxW0; k.b[r] = k1 + (uint32_t)r * kPhiloxW1; }
  return k;
}

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  const PhiloxKeys& k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)kPhiloxM0 * c0;
    const unsigned long long p1 = (unsigned long long)kPhiloxM1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.a[r];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.b[r];
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  return Philox4{c0, c1, c2, c3};
}

// ((a >> 5) * 2^26 + (b >> 6)) / 2^53, exactly (numpy legacy random_double).  Two 32-bit
// I2F (XU pipe) + DMUL + DFMA (FP64 pipe) instead of 64-bit shifts/OR: ncu (profiles/r1a)
// shows this kernel bound by the fmaheavy pipe (IMAD.WIDE of the Philox rounds, 82 %) with the
// integer ALU second, so the conversion is kept off both.  Every step is exact: hi*2^-27 and
// lo*2^-53 are exact scalings and their sum is a 53-bit multiple of 2^-53 below 1.
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  const double hi = (double)(a >> 5);
  const double lo = (double)(b >> 6);
  return __fma_rn(hi, 7.450580596923828125e-09, __dmul_rn(lo, 1.1102230246251565404e-16));
}

}  // namespace bci
