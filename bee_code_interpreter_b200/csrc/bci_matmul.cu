// bci_matmul.cu — K3: numpy.matmul / `@` for 2-D float32, B requests per launch
// (the reference reaches OpenBLAS sgemm through numpy here).
//
// Two device paths, both CUDA, selected per call:
//   * tcgen05 path (bci_matmul_tc.cu): TMA-staged shared-memory ring, tcgen05.mma
//     kind::tf32 with the fp32 accumulator in TMEM, 3xTF32 operand split
//     (hi*hi + hi*lo + lo*hi) so the result stays within ~1e-7 normwise of fp32;
//     used when M,N,K are multiples of its tile (128/128/32) and pointers are 16 B aligned;
//   * SIMT fp32 path (this file): 128x128x16 register-tiled FFMA kernel, any shape.
// Environment BCI_MATMUL=simt forces the SIMT path (used by the parity tests to check
// the two paths against each other).
#include "bci_common.cuh"
#include <vector>
#include <cstring>
#include <cstdlib>

namespace bci {

int matmul_tc_supported(int M, int N, int K, const void* const* A, const void* const* Bm, void* const* C, int B);
int matmul_tc_launch(int B, const void* const* A, const void* const* Bm, void* const* C, int M, int N, int K);

constexpr int BM = 128, BN = 128, BK = 16;

struct MatReq { const float* A; const float* Bm; float* C; };

__global__ void __launch_bounds__(256)
sgemm_simt_kernel(const MatReq* __restrict__ reqs, int M, int N, int K) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const MatReq q = reqs[blockIdx.z];
  const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8x8 outputs as 2x2 blocks of 4x4
  // global->smem mapping: A tile 128x16 = 512 float4 (2 per thread); B tile 16x128 = 512 float4
  const int a_row = tid >> 2, a_k4 = (tid & 3) * 4;          // rows 0..63 (+64), k offset 0,4,8,12
  const int b_k = tid >> 5, b_n4 = (tid & 31) * 4;           // k rows 0..7 (+8), n offset 0..124
  const bool a_vec = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(q.A) & 15) == 0);
  const bool b_vec = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(q.Bm) & 15) == 0);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto load_global = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = bm + a_row + h * 64, kk = k0 + a_k4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < M) {
        const float* p = q.A + (size_t)row * K + kk;
        if (a_vec && kk + 3 < K) v = *reinterpret_cast<const float4*>(p);
        else {
          if (kk + 0 < K) v.x = p[0];
          if (kk + 1 < K) v.y = p[1];
          if (kk + 2 < K) v.z = p[2];
          if (kk + 3 < K) v.w = p[3];
        }
      }
      ra[h] = v;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = k0 + b_k + h * 8, col = bn + b_n4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < K) {
        const float* p = q.Bm + (size_t)kk * N + col;
        if (b_vec && col + 3 < N) v = *reinterpret_cast<const float4*>(p);
        else {
          if (col + 0 < N) v.x = p[0];
          if (col + 1 < N) v.y = p[1];
          if (col + 2 < N) v.z = p[2];
          if (col + 3 < N) v.w = p[3];
        }
      }
      rb[h] = v;
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = a_row + h * 64;
      As[buf][a_k4 + 0][row] = ra[h].x;
      As[buf][a_k4 + 1][row] = ra[h].y;
      As[buf][a_k4 + 2][row] = ra[h].z;
      As[buf][a_k4 + 3][row] = ra[h].w;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
      *reinterpret_cast<float4*>(&Bs[buf][b_k + h * 8][b_n4]) = rb[h];
  };

  const int nk = (K + BK - 1) / BK;
  load_global(0);
  store_smem(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_global((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + 64]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + 64]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_smem(buf ^ 1);
      __syncthreads();
    }
  }
  const bool c_vec = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(q.C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = bm + ty * 4 + (i & 3) + (i >> 2) * 64;
    if (row >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = bn + tx * 4 + jh * 64;
      float* p = q.C + (size_t)row * N + col;
      if (c_vec && col + 3 < N) {
        *reinterpret_cast<float4*>(p) = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < N) p[j] = acc[i][jh * 4 + j];
      }
    }
  }
}

static int matmul_simt(int B, const void* const* A, const void* const* Bm, void* const* C, int M, int N, int K) {
  Ctx& c = ctx();
  std::vector<MatReq> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    reqs[i].A = reinterpret_cast<const float*>(A[i]);
    reqs[i].Bm = reinterpret_cast<const float*>(Bm[i]);
    reqs[i].C = reinterpret_cast<float*>(C[i]);
  }
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    void* d = nullptr;
    int rc = stage_upload(reqs.data() + b0, sizeof(MatReq) * (size_t)nb, &d);
    if (rc != BCI_OK) return rc;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, nb);
    sgemm_simt_kernel<<<grid, 256, 0, c.stream>>>(reinterpret_cast<const MatReq*>(d), M, N, K);
    BCI_CHECK_LAUNCH();
  }
  return BCI_OK;
}

}  // namespace bci

using namespace bci;

extern "C" int bci_matmul_f32(int B, const void* const* A, const void* const* Bm, void* const* C,
                              int M, int N, int K) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !A || !Bm || !C || M < 0 || N < 0 || K < 0) { set_error("bci_matmul_f32: bad arguments"); return BCI_ERR_INVALID; }
  for (int i = 0; i < B; ++i)
    if (!A[i] || !Bm[i] || !C[i]) { set_error("bci_matmul_f32: NULL matrix in request %d", i); return BCI_ERR_INVALID; }
  if (M == 0 || N == 0) return BCI_OK;
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  if (K == 0) {
    for (int i = 0; i < B; ++i) BCI_CUDA(cudaMemsetAsync(C[i], 0, (size_t)M * N * sizeof(float), c.stream));
    return BCI_OK;
  }
  const char* force = tuning_env("BCI_MATMUL");
  const bool want_simt = force && strcmp(force, "simt") == 0;
  if (!want_simt && matmul_tc_supported(M, N, K, A, Bm, C, B)) return matmul_tc_launch(B, A, Bm, C, M, N, K);
  if (force && strcmp(force, "tc") == 0) {
    set_error("bci_matmul_f32: BCI_MATMUL=tc but shape %dx%dx%d is not supported by the tcgen05 path", M, N, K);
    return BCI_ERR_UNSUPPORTED;
  }
  return matmul_simt(B, A, Bm, C, M, N, K);
}
