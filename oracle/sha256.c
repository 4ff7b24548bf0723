/*
 * oracle/sha256.c — TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * CPU restatement of SHA-256 exactly as FIPS 180-4 §5.1.1, §5.3.3, §6.2 specifies it.
 * Parity status: the reference computes NO SHA-256 (services/storage.py:52 names objects
 * with secrets.token_hex(32); the `sha2 = "0.10"` crate of executor/Cargo.toml:20, locked
 * 0.10.8 at executor/Cargo.lock:1158-1159, has no call site in executor/server.rs:15-27),
 * so hash VALUES are "parity unpinned" in the reference.  This oracle is pinned instead to
 * the FIPS 180-4 / NIST CAVP known-answer vectors and to hashlib.sha256 (OpenSSL) in
 * tests/test_oracle.py, including the reference's own test payload "Hello, World!"
 * (test/e2e/test_http.py:48).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static uint32_t rotr32(uint32_t x, unsigned n) { return (x >> n) | (x << (32u - n)); }

/* FIPS 180-4 §6.2.2: one 512-bit block into the chaining value H. */
static void compress(uint32_t H[8], const uint8_t blk[64]) {
  uint32_t W[64];
  for (int t = 0; t < 16; ++t)
    W[t] = ((uint32_t)blk[4 * t] << 24) | ((uint32_t)blk[4 * t + 1] << 16) | ((uint32_t)blk[4 * t + 2] << 8) |
           (uint32_t)blk[4 * t + 3];
  for (int t = 16; t < 64; ++t) {
    uint32_t s0 = rotr32(W[t - 15], 7) ^ rotr32(W[t - 15], 18) ^ (W[t - 15] >> 3);
    uint32_t s1 = rotr32(W[t - 2], 17) ^ rotr32(W[t - 2], 19) ^ (W[t - 2] >> 10);
    W[t] = s1 + W[t - 7] + s0 + W[t - 16];
  }
  uint32_t a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5], g = H[6], h = H[7];
  for (int t = 0; t < 64; ++t) {
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t T1 = h + S1 + ch + K[t] + W[t];
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t T2 = S0 + mj;
    h = g; g = f; f = e; e = d + T1; d = c; c = b; b = a; a = T1 + T2;
  }
  H[0] += a; H[1] += b; H[2] += c; H[3] += d; H[4] += e; H[5] += f; H[6] += g; H[7] += h;
}

/* digest[32] = SHA-256(data[0..len)) */
void oracle_sha256(const uint8_t* data, size_t len, uint8_t digest[32]) {
  uint32_t H[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  size_t full = len / 64;
  for (size_t i = 0; i < full; ++i) compress(H, data + 64 * i);
  /* §5.1.1 padding: 0x80, zeros to 56 mod 64, 64-bit big-endian bit length */
  uint8_t tail[128];
  size_t r = len - 64 * full;
  memset(tail, 0, sizeof tail);
  if (r) memcpy(tail, data + 64 * full, r);
  tail[r] = 0x80;
  size_t tl = (r + 9 <= 64) ? 64 : 128;
  uint64_t bits = (uint64_t)len * 8u;
  for (int i = 0; i < 8; ++i) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
  compress(H, tail);
  if (tl == 128) compress(H, tail + 64);
  for (int i = 0; i < 8; ++i) {
    digest[4 * i] = (uint8_t)(H[i] >> 24);
    digest[4 * i + 1] = (uint8_t)(H[i] >> 16);
    digest[4 * i + 2] = (uint8_t)(H[i] >> 8);
    digest[4 * i + 3] = (uint8_t)H[i];
  }
}

/* B messages, one after the other (the CPU baseline loop of bench.py --workload sha256). */
void oracle_sha256_batch(int B, const uint8_t* const* data, const size_t* len, uint8_t* digests) {
  for (int i = 0; i < B; ++i) oracle_sha256(data[i], len[i], digests + 32 * (size_t)i);
}
