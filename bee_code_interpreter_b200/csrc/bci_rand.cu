// bci_rand.cu — K8: uniform float64 in [0,1) for the *unseeded* numpy.random.rand of
// examples/benchmark-numpy.py:20.
//
// Generator: Philox4x32-10 (Salmon et al., "Parallel Random Numbers: As Easy as 1,2,3",
// SC'11; Random123 v1.14) — counter based, so element i needs no state:
//   counter = (lo32(i/2), hi32(i/2), 0, 0), key = (lo32(seed), hi32(seed));
//   the 4 output words give elements 2*(i/2) and 2*(i/2)+1 with numpy's own 53-bit
//   construction  ((a >> 5) * 2^26 + (b >> 6)) / 2^53   (numpy legacy random_double).
// Same distribution and resolution as numpy's rand, not the MT19937 stream (SURVEY.md
// hard part 5); seeded legacy calls are never routed here.
//
// HBM-write-bound: algorithmic bytes = 8*n.  One CTA = one 32 KiB tile, 4 x STG.256
// per thread; the 10 round keys are warp-uniform so a round is 2 IMAD.WIDE + 2 LOP3.
#include "bci_philox.cuh"
#include <vector>
#include <cstring>

namespace bci {

// Standard normals by Box-Muller on the same counter layout (third counter word = 1, so the
// normal stream of a seed is not the uniform one): pair p -> u_a, u_b ->
//   r = sqrt(-2 ln(1 - u_a)),  z[2p] = r cos(2 pi u_b),  z[2p+1] = r sin(2 pi u_b);
// out = loc + scale * z with numpy's two roundings (multiply, then add).  numpy's legacy
// normal() is the polar method on MT19937 and unseeded, so only the distribution is a parity
// target; the stream itself is pinned against oracle/philox.c.
__device__ __forceinline__ void box_muller(const Philox4& p, double loc, double scale, double& z0, double& z1) {
  const double ua = u53(p.x, p.y), ub = u53(p.z, p.w);
  const double r = sqrt(-2.0 * log(1.0 - ua));
  double s, c;
  sincospi(2.0 * ub, &s, &c);
  z0 = __dadd_rn(loc, __dmul_rn(scale, r * c));
  z1 = __dadd_rn(loc, __dmul_rn(scale, r * s));
}

constexpr int kRandUnroll = 4;
constexpr size_t kRandTile = (size_t)kThreads * kRandUnroll * 4;  // doubles

__global__ void __launch_bounds__(kThreads)
rand_f64_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B) {
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * kRandTile;
  const size_t rem = (size_t)q.n - base;
  double* __restrict__ o = reinterpret_cast<double*>(q.out) + base;
  const PhiloxKeys keys = philox_keys((uint32_t)((unsigned long long)q.s1.i), (uint32_t)((unsigned long long)q.s1.i >> 32));

  if (rem >= kRandTile && is_aligned32(o)) {
#pragma unroll
    for (int u = 0; u < kRandUnroll; ++u) {
      // base is a multiple of the 4096-element tile and the in-tile offset is < 4096, so the
      // pair counter is (base >> 1) | (offset >> 1): no 64-bit carry chains in the hot loop
      const uint32_t off = (uint32_t)(u * kThreads + threadIdx.x) * 4u;
      const uint32_t plo = (uint32_t)(base >> 1) | (off >> 1), phi = (uint32_t)(base >> 33);
      const Philox4 a = philox4x32_10(plo, phi, 0u, 0u, keys);
      const Philox4 b = philox4x32_10(plo | 1u, phi, 0u, 0u, keys);
      VecOf<double> v;
      v.v[0] = u53(a.x, a.y);
      v.v[1] = u53(a.z, a.w);
      v.v[2] = u53(b.x, b.y);
      v.v[3] = u53(b.z, b.w);
      stg256(o + ((size_t)u * kThreads + threadIdx.x) * 4, v.u);
    }
  } else {
    const size_t m = rem < kRandTile ? rem : kRandTile;
    for (size_t i = threadIdx.x; i < m; i += kThreads) {
      const size_t e = base + i;
      const unsigned long long p = e >> 1;
      const Philox4 a = philox4x32_10((uint32_t)p, (uint32_t)(p >> 32), 0u, 0u, keys);
      o[i] = (e & 1) ? u53(a.z, a.w) : u53(a.x, a.y);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
rand_normal_f64_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B) {
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * kRandTile;
  const size_t rem = (size_t)q.n - base;
  double* __restrict__ o = reinterpret_cast<double*>(q.out) + base;
  const PhiloxKeys keys = philox_keys((uint32_t)((unsigned long long)q.s1.i), (uint32_t)((unsigned long long)q.s1.i >> 32));
  const double loc = q.s0.f;
  double scale;
  memcpy(&scale, &q.aux, 8);
  if (rem >= kRandTile && is_aligned32(o)) {
#pragma unroll
    for (int u = 0; u < kRandUnroll; ++u) {
      const uint32_t off = (uint32_t)(u * kThreads + threadIdx.x) * 4u;
      const uint32_t plo = (uint32_t)(base >> 1) | (off >> 1), phi = (uint32_t)(base >> 33);
      const Philox4 a = philox4x32_10(plo, phi, 1u, 0u, keys);
      const Philox4 b = philox4x32_10(plo | 1u, phi, 1u, 0u, keys);
      VecOf<double> v;
      box_muller(a, loc, scale, v.v[0], v.v[1]);
      box_muller(b, loc, scale, v.v[2], v.v[3]);
      stg256(o + ((size_t)u * kThreads + threadIdx.x) * 4, v.u);
    }
  } else {
    const size_t m = rem < kRandTile ? rem : kRandTile;
    for (size_t i = threadIdx.x; i < m; i += kThreads) {
      const size_t e = base + i;
      const unsigned long long p = e >> 1;
      const Philox4 a = philox4x32_10((uint32_t)p, (uint32_t)(p >> 32), 1u, 0u, keys);
      double z0, z1;
      box_muller(a, loc, scale, z0, z1);
      o[i] = (e & 1) ? z1 : z0;
    }
  }
}

}  // namespace bci

using namespace bci;

extern "C" int bci_rand_normal_f64(int B, void* const* out, const size_t* n, const uint64_t* seed,
                                   const double* loc, const double* scale) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !out || !n || !seed || !loc || !scale) { set_error("bci_rand_normal_f64: bad arguments"); return BCI_ERR_INVALID; }
  std::vector<Req> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    Req& q = reqs[i];
    memset(&q, 0, sizeof q);
    q.out = out[i];
    q.n = n[i];
    q.s1.i = (int64_t)seed[i];
    q.s0.f = loc[i];
    memcpy(&q.aux, &scale[i], 8);
    if (n[i] && !out[i]) { set_error("bci_rand_normal_f64: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
  }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  Batch bt;
  int rc = upload_batch(reqs.data(), B, kRandTile, &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_rand_normal_f64: batch too large for one launch"); return BCI_ERR_INVALID; }
  rand_normal_f64_kernel<<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}


extern "C" int bci_rand_f64(int B, void* const* out, const size_t* n, const uint64_t* seed) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !out || !n || !seed) { set_error("bci_rand_f64: bad arguments"); return BCI_ERR_INVALID; }
  std::vector<Req> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    Req& q = reqs[i];
    memset(&q, 0, sizeof q);
    q.out = out[i];
    q.n = n[i];
    q.s1.i = (int64_t)seed[i];
    if (n[i] && !out[i]) { set_error("bci_rand_f64: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
  }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  Batch bt;
  int rc = upload_batch(reqs.data(), B, kRandTile, &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_rand_f64: batch too large for one launch"); return BCI_ERR_INVALID; }
  rand_f64_kernel<<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}
