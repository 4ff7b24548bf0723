/*
 * oracle/philox.c — TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * CPU restatement of Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel Random Numbers:
 * As Easy as 1, 2, 3", SC'11; Random123 v1.14 philox.h) and of the uniform-double
 * construction numpy's legacy generator uses for numpy.random.rand
 * (random_double: (a >> 5, b >> 6) -> (a*2^26 + b) / 2^53), which is what the reference's
 * metric script calls (examples/benchmark-numpy.py:20).
 * Parity status: the reference call is UNSEEDED MT19937, so only the distribution is a
 * parity target (SURVEY.md hard part 5); this file pins the device generator bit-for-bit
 * and is itself pinned to the Random123 known-answer vectors in tests/test_oracle.py.
 */
#include <stddef.h>
#include <stdint.h>

#define M0 0xD2511F53u
#define M1 0xCD9E8D57u
#define W0 0x9E3779B9u
#define W1 0xBB67AE85u

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u53(uint32_t a, uint32_t b) {
  uint64_t v = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
  return (double)v / 9007199254740992.0;
}

/* out[i] for i in [0,n): element i comes from counter (i/2, 0, 0, 0), words (0,1) or (2,3). */
void oracle_rand_f64(double* out, size_t n, uint64_t seed) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (size_t p = 0; 2 * p < n; ++p) {
    uint32_t ctr[4] = {(uint32_t)p, (uint32_t)((uint64_t)p >> 32), 0, 0}, r[4];
    oracle_philox4x32_10(ctr, key, r);
    out[2 * p] = u53(r[0], r[1]);
    if (2 * p + 1 < n) out[2 * p + 1] = u53(r[2], r[3]);
  }
}

/* Box-Muller normals on counter (p, 0, 1, 0): z[2p] = r cos(2 pi u_b), z[2p+1] = r sin(2 pi u_b),
 * r = sqrt(-2 ln(1 - u_a)); out = loc + scale * z.  Restates bci_rand.cu:box_muller with libm. */
#include <math.h>
void oracle_rand_normal_f64(double* out, size_t n, uint64_t seed, double loc, double scale) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (size_t p = 0; 2 * p < n; ++p) {
    uint32_t ctr[4] = {(uint32_t)p, (uint32_t)((uint64_t)p >> 32), 1, 0}, r[4];
    oracle_philox4x32_10(ctr, key, r);
    double ua = u53(r[0], r[1]), ub = u53(r[2], r[3]);
    double rad = sqrt(-2.0 * log(1.0 - ua));
    double ang = 2.0 * 3.14159265358979323846 * ub;
    out[2 * p] = loc + scale * (rad * cos(ang));
    if (2 * p + 1 < n) out[2 * p + 1] = loc + scale * (rad * sin(ang));
  }
}
