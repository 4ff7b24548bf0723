// bci_matmul_tc.cu — tcgen05 / TMA path of K3 (numpy.matmul, 2-D float32).
//
// fp32 accuracy on the tf32 tensor pipe: every operand is split x = hi + lo with
// hi = rna_tf32(x), lo = rna_tf32(x - hi) and the product is accumulated in fp32 (TMEM) as
//     A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi            (3xTF32; the dropped lo*lo term is
// 2^-22 relative), which keeps max|dC|/max|C| around 1e-7 where a single tf32 pass gives 3e-4
// (SURVEY.md hard part 1).
//
// Two kernels per batch slice:
//  1. split_kernel   : A -> A_hi, A_lo (row-major MxK, K contiguous); B -> Bt_hi, Bt_lo
//                      TRANSPOSED to NxK so that both MMA operands are K-major.  HBM-bound,
//                      (M*K + K*N) * 12 bytes per request.
//  2. gemm3x_kernel  : one CTA per 128x256 (or 128x128) output tile, K walked in 32-element
//                      (128-byte) blocks through a 2- (or 3-) stage shared-memory ring:
//        warp 0  (1 thread) TMA producer: 4 x cp.async.bulk.tensor (A_hi, A_lo, Bt_hi, Bt_lo
//                 tiles of 128 rows x 128 B, SWIZZLE_128B) per stage, mbarrier complete_tx;
//        warp 1  (1 thread) MMA issuer: 12 x tcgen05.mma.cta_group::1.kind::tf32 per stage
//                 (3 products x 4 K-steps of 8) into one of TWO 128-column TMEM accumulators;
//                 tcgen05.commit releases the stage and, every 4 k-blocks, hands the
//                 accumulator to the epilogue;
//        warps 2-5 epilogue: per chunk tcgen05.ld 32x32b.x32 -> fp32 REGISTER accumulators
//                 (round-to-nearest adds: the tensor core truncates when it accumulates, and
//                 1536 dependent adds at K=4096 cost 3e-5 of max|C|), 128-bit stores at the end.
//      Tensor maps are 3-D (k, row, request) over the split planes, which this file lays
//      out itself, so ONE launch covers every request of the slice (grid.z = request).
// Tensor-pipe bound in principle (137.4 GFLOP algorithmic per 4096^3 request, x3 issued);
// the 128x128 tile was limited by operand delivery (64 KiB per 128x128x32 step, 3 stages = 1.6 us
// of latency cover: tensor pipe 65 % busy); the 128x256 tile (GemmCfg<256>) moves 25 % fewer
// operand bytes per flop and runs the GEMM launches at 0.95 of the 3xTF32 ceiling.
#include "bci_common.cuh"
#include <cuda.h>
#include <vector>
#include <cstring>
#include <cstdlib>

namespace bci {

constexpr int TBM = 128, TBN = 128, TBK = 32;          // CTA tile; TBK * 4 B = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kTileBytes = TBM * TBK * 4;              // 16 KiB per operand plane per stage
constexpr int kStageBytes = 4 * kTileBytes;            // A_hi, A_lo, Bt_hi, Bt_lo
constexpr int kTmemCols = 256;                         // two 128-column accumulators (ping-pong)
constexpr int kChunkKB = 4;                            // k-blocks (K=128, 48 MMAs) accumulated inside TMEM before a flush
constexpr int kGemmThreads = 192;
constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

// The GEMM kernel exists for two CTA tiles.  128x128: 3 stages of 64 KiB.  128x256: one A tile
// feeds twice the MMA work, i.e. 96 KiB of operands per 128x256x32 step instead of 128 KiB for the
// same flops (-25 % L2->SM traffic), at the price of a 2-stage ring; 8 epilogue warps (the TMEM
// lane quadrant is warp % 4, the two warps of a quadrant take 128 columns each).
template <int BN> struct GemmCfg {
  static constexpr int kStg = BN == 128 ? 3 : 2;
  static constexpr int kATile = TBM * TBK * 4;
  static constexpr int kBTile = BN * TBK * 4;
  static constexpr int kStage = 2 * kATile + 2 * kBTile;
  static constexpr int kTmem = 2 * BN;
  static constexpr int kEpiWarps = BN / 32;
  static constexpr int kThr = 64 + 32 * kEpiWarps;
  static constexpr size_t kSmem = (size_t)kStg * kStage + 1024 + 256;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  // bounded spin: a protocol bug must end in a trap, not in a hung GPU
  for (unsigned long long spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spins > (1ull << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory operand descriptor (sm_100 UMMA format):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 = 1024 B
//   (8 rows x 128 B per swizzle atom) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32 instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- kernel 1: split (+ transpose B)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  const float r = __fsub_rn(x, hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
  lo = __uint_as_float(l);
}

struct SplitReq { const float* A; const float* Bm; };

// grid: (tiles, 1, nb).  tiles [0, ta) split A linearly; tiles [ta, ta+tb) transpose 32x32 blocks of B.
__global__ void __launch_bounds__(256)
split_kernel(const SplitReq* __restrict__ reqs, float* __restrict__ a_hi, float* __restrict__ a_lo,
             float* __restrict__ bt_hi, float* __restrict__ bt_lo, int M, int N, int K, unsigned ta) {
  __shared__ float th[32][33], tl[32][33];
  const SplitReq q = reqs[blockIdx.z];
  const size_t abase = (size_t)blockIdx.z * M * K, bbase = (size_t)blockIdx.z * N * K;
  if (blockIdx.x < ta) {
    // 256 threads x 4 float4 = 4096 elements per CTA
    const size_t e0 = (size_t)blockIdx.x * 4096;
    const size_t total = (size_t)M * K;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t e = e0 + ((size_t)u * 256 + threadIdx.x) * 4;
      if (e + 3 < total) {
        const float4 v = *reinterpret_cast<const float4*>(q.A + e);
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y);
        split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        *reinterpret_cast<float4*>(a_hi + abase + e) = h;
        *reinterpret_cast<float4*>(a_lo + abase + e) = l;
      }
    }
  } else {
    // B is K x N row-major; write Bt (N x K): 32x32 tile through padded smem
    const unsigned t = blockIdx.x - ta;
    const int tiles_n = N / 32;
    const int k0 = (t / tiles_n) * 32, n0 = (t % tiles_n) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = ty + r * 8;
      float h, l;
      split_tf32(q.Bm[(size_t)(k0 + kk) * N + n0 + tx], h, l);
      th[kk][tx] = h;
      tl[kk][tx] = l;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int nn = ty + r * 8;
      bt_hi[bbase + (size_t)(n0 + nn) * K + k0 + tx] = th[tx][nn];
      bt_lo[bbase + (size_t)(n0 + nn) * K + k0 + tx] = tl[tx][nn];
    }
  }
}

// ---------------------------------------------------------------- kernel 2: 3xTF32 GEMM
struct GemmOut { float* C; };

template <int BN>
__global__ void __launch_bounds__(GemmCfg<BN>::kThr, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
              const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
              const GemmOut* __restrict__ outs, int N, int K, int raster_group, int prefetch_kb) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStg;
  constexpr int kStageBytes = Cfg::kStage;
  constexpr int kOffAlo = Cfg::kATile, kOffBhi = 2 * Cfg::kATile, kOffBlo = 2 * Cfg::kATile + Cfg::kBTile;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* tmem_full_bar = bars + 2 * kStages;   // [2]
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Grouped rasterisation: CTAs are issued x-fastest, so co-resident CTAs would share ONE row
  // panel of A and sweep every column panel of B (B re-read from HBM once per ~4.6 tile rows).
  // Walking `raster_group` tile rows per column instead keeps a (group x ~18) block of tiles
  // resident: each B panel is reused by `raster_group` tile rows out of L2.  Measured at 4096^3:
  // 201.7 (off) -> 206.9 TFLOP/s (8); 16 and 32 are worse (195).
  int tile_n = blockIdx.x, tile_m = blockIdx.y;
  if (raster_group > 1) {
    const int tiles_n = gridDim.x, tiles_m = gridDim.y;
    const int lin = blockIdx.x + tiles_n * blockIdx.y;
    const int per_group = raster_group * tiles_n;
    const int first_m = (lin / per_group) * raster_group;
    const int rows = min(tiles_m - first_m, raster_group);
    tile_m = first_m + (lin % per_group) % rows;
    tile_n = (lin % per_group) / rows;
  }
  const int n0 = tile_n * BN, m0 = tile_m * TBM, req = blockIdx.z;
  const int num_kb = K / TBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_lo)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], Cfg::kEpiWarps); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::kTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* st = smem + (size_t)s * kStageBytes;
        mbar_expect_tx(&full_bar[s], kStageBytes);
        tma_load_3d(st, &map_a_hi, &full_bar[s], kb * TBK, m0, req);
        tma_load_3d(st + kOffAlo, &map_a_lo, &full_bar[s], kb * TBK, m0, req);
        tma_load_3d(st + kOffBhi, &map_b_hi, &full_bar[s], kb * TBK, n0, req);
        tma_load_3d(st + kOffBlo, &map_b_lo, &full_bar[s], kb * TBK, n0, req);
        // the 3-stage ring covers ~1.6 us of operand latency; pull the k-block `prefetch_kb` ahead
        // into L2 now so that its own load is an L2 hit by the time a stage frees up
        if (prefetch_kb > 0 && kb + prefetch_kb < num_kb) {
          const int kp = (kb + prefetch_kb) * TBK;
          tma_prefetch_3d(&map_a_hi, kp, m0, req);
          tma_prefetch_3d(&map_a_lo, kp, m0, req);
          tma_prefetch_3d(&map_b_hi, kp, n0, req);
          tma_prefetch_3d(&map_b_lo, kp, n0, req);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TBM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        // The tensor core adds into its fp32 accumulator with truncation; over K=4096 (1536
        // dependent adds) that bias reaches 3e-5 of max|C| (measured).  So only kChunkKB k-blocks
        // are accumulated inside TMEM; the epilogue warps then add the chunk into fp32 registers
        // with round-to-nearest while the MMAs continue into the other TMEM accumulator.
        const int chunk = kb / kChunkKB, buf = chunk & 1;
        const bool first_in_chunk = (kb % kChunkKB) == 0;
        if (first_in_chunk) {
          mbar_wait(&tmem_empty_bar[buf], (((uint32_t)chunk >> 1) & 1u) ^ 1u);
          tc_fence_after();
        }
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)s * kStageBytes);
        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kOffAlo);
        const uint64_t b_hi = umma_desc_sw128(st + kOffBhi), b_lo = umma_desc_sw128(st + kOffBlo);
        const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * BN);
#pragma unroll
        for (int k = 0; k < TBK / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 32 bytes per K-step of 8, in 16-byte units
          umma_tf32(tmem_acc, a_lo + adv, b_hi + adv, idesc, (first_in_chunk && k == 0) ? 0u : 1u);
          umma_tf32(tmem_acc, a_hi + adv, b_lo + adv, idesc, 1u);
          umma_tf32(tmem_acc, a_hi + adv, b_hi + adv, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);  // the stage is free once these MMAs have read it
        if ((kb % kChunkKB) == kChunkKB - 1 || kb == num_kb - 1) umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // epilogue warps 2..: TMEM lane quadrant = warp % 4; with 8 of them the two warps of a
    // quadrant split the accumulator's columns in halves of 128
    const int quad = warp & 3;
    const int col0 = ((warp - 2) >> 2) * 128;
    constexpr int EN = 128;  // columns per epilogue thread
    const int num_chunks = (num_kb + kChunkKB - 1) / kChunkKB;
    float acc[EN];
#pragma unroll
    for (int j = 0; j < EN; ++j) acc[j] = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < num_chunks; ++ch) {
      const int buf = ch & 1;
      mbar_wait(&tmem_full_bar[buf], ((uint32_t)ch >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < EN / 32; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + col0 + c * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c * 32 + j] = __fadd_rn(acc[c * 32 + j], __uint_as_float(v[j]));
      }
      // this accumulator may be overwritten by the MMAs of chunk ch+2
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[buf])) : "memory");
    }
    float* __restrict__ C = outs[req].C;
    const int row = m0 + quad * 32 + lane;
    float* crow = C + (size_t)row * N + n0 + col0;
#pragma unroll
    for (int j = 0; j < EN; j += 4)
      *reinterpret_cast<float4*>(crow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmem) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D map over a [nb][rows][K] float plane: box = 32 (k) x 128 (rows) x 1, 128-byte swizzle
static int make_map(CUtensorMap* map, float* base, int rows, int K, int nb, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable in this driver"); return BCI_ERR_CUDA; }
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)nb};
  const cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)rows * K * 4};
  const cuuint32_t box[3] = {(cuuint32_t)TBK, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return BCI_ERR_CUDA; }
  return BCI_OK;
}

int matmul_tc_supported(int M, int N, int K, const void* const* A, const void* const* Bm, void* const* C, int B) {
  if (M % TBM || N % TBN || K % TBK || M <= 0 || N <= 0 || K <= 0) return 0;
  for (int i = 0; i < B; ++i) {
    if ((reinterpret_cast<uintptr_t>(A[i]) | reinterpret_cast<uintptr_t>(Bm[i]) | reinterpret_cast<uintptr_t>(C[i])) & 15) return 0;
  }
  return encode_fn() != nullptr;
}

int matmul_tc_launch(int B, const void* const* A, const void* const* Bm, void* const* C, int M, int N, int K) {
  Ctx& c = ctx();
  static bool attr_set = false;
  if (!attr_set) {
    BCI_CUDA(cudaFuncSetAttribute(gemm3x_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<128>::kSmem));
    BCI_CUDA(cudaFuncSetAttribute(gemm3x_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<256>::kSmem));
    attr_set = true;
  }
  int raster_group = 8;
  if (const char* g = tuning_env("BCI_GEMM_GROUP")) raster_group = atoi(g);
  // L2 prefetch of the k-block `prefetch_kb` ahead (cp.async.bulk.prefetch.tensor): measured
  // 4096^3 x 8: off 204 TFLOP/s, 4/8/16/32 ahead 196 — the extra TMA traffic costs more than the
  // shorter load latency returns, so it stays off (kept as an experiment switch)
  int prefetch_kb = 0;
  if (const char* g = tuning_env("BCI_GEMM_PREFETCH")) prefetch_kb = atoi(g);
  // CTA tile width: 256 whenever N allows (measured 4096^3 x 8: 230.5 vs 205.9 TFLOP/s algorithmic
  // for the 128-wide tile; the GEMM launches alone run at 267 = 0.95 of the 3xTF32 ceiling, the
  // rest is the split pass); BCI_GEMM_TILE_N=128 forces the narrow tile
  int tile_n = N % 256 == 0 ? 256 : 128;
  if (const char* g = tuning_env("BCI_GEMM_TILE_N")) tile_n = (atoi(g) == 256 && N % 256 == 0) ? 256 : 128;
  // Optional two-deep pipeline over slices of the batch (BCI_GEMM_PIPE=1): split(i+1) on a second
  // stream under gemm(i).  Measured slower both times it was tried (8 x 4096^3: 5.35 vs 5.31 ms
  // with the 128-wide tile, 5.35 vs 4.75 ms with the 256-wide one): one-matrix slices leave a
  // 3.46-wave tail per GEMM launch and the HBM-saturating split delays the operand loads.  Off.
  bool pipe = false;
  if (const char* g = tuning_env("BCI_GEMM_PIPE")) pipe = atoi(g) != 0;

  // split planes: hi + lo of A and of B^T, 8 bytes per input element; two buffers when pipelined,
  // at most ~8 GiB of scratch in total
  const size_t per_req = ((size_t)M * K + (size_t)N * K) * 8;
  int cap = (int)((4ull << 30) / (per_req ? per_req : 1));
  if (cap < 1) cap = 1;
  int slice = pipe ? (B + 7) / 8 : B;
  if (slice > cap) slice = cap;
  if (slice > 65535) slice = 65535;
  const int n_slices = (B + slice - 1) / slice;
  const int n_bufs = (pipe && n_slices > 1) ? 2 : 1;
  void* sc = nullptr;
  int rc = scratch_get(per_req * (size_t)slice * n_bufs + 4096, &sc);
  if (rc != BCI_OK) return rc;
  float* plane0 = reinterpret_cast<float*>(sc);
  const size_t buf_floats = per_req * (size_t)slice / 4;

  static cudaStream_t split_stream = nullptr;
  static cudaEvent_t ev_in = nullptr, ev_split[2] = {nullptr, nullptr}, ev_gemm[2] = {nullptr, nullptr};
  static int pipe_device = -1;
  if (n_bufs == 2 && pipe_device != c.device) {
    int lo = 0, hi = 0;
    BCI_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    BCI_CUDA(cudaStreamCreateWithPriority(&split_stream, cudaStreamNonBlocking, hi));
    BCI_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) {
      BCI_CUDA(cudaEventCreateWithFlags(&ev_split[b], cudaEventDisableTiming));
      BCI_CUDA(cudaEventCreateWithFlags(&ev_gemm[b], cudaEventDisableTiming));
    }
    pipe_device = c.device;
  }

  // every descriptor goes up front: an upload queued behind gemm(i) would serialise split(i+1)
  std::vector<SplitReq> sreq((size_t)B);
  std::vector<GemmOut> outs((size_t)B);
  for (int i = 0; i < B; ++i) {
    sreq[i].A = reinterpret_cast<const float*>(A[i]);
    sreq[i].Bm = reinterpret_cast<const float*>(Bm[i]);
    outs[i].C = reinterpret_cast<float*>(C[i]);
  }
  void *d_sreq = nullptr, *d_outs = nullptr;
  rc = stage_upload(sreq.data(), sizeof(SplitReq) * (size_t)B, &d_sreq);
  if (rc != BCI_OK) return rc;
  rc = stage_upload(outs.data(), sizeof(GemmOut) * (size_t)B, &d_outs);
  if (rc != BCI_OK) return rc;
  cudaStream_t sstream = c.stream;
  if (n_bufs == 2) {
    sstream = split_stream;
    BCI_CUDA(cudaEventRecord(ev_in, c.stream));   // inputs, descriptors, earlier users of the scratch
    BCI_CUDA(cudaStreamWaitEvent(split_stream, ev_in, 0));
  }
  const unsigned ta = (unsigned)(((size_t)M * K + 4095) / 4096);
  const unsigned tb = (unsigned)((size_t)(K / 32) * (N / 32));
  for (int idx = 0; idx < n_slices; ++idx) {
    const int b0 = idx * slice;
    const int nb = B - b0 < slice ? B - b0 : slice;
    const int buf = idx % n_bufs;
    float* a_hi = plane0 + (size_t)buf * buf_floats;
    float* a_lo = a_hi + (size_t)slice * M * K;
    float* bt_hi = a_lo + (size_t)slice * M * K;
    float* bt_lo = bt_hi + (size_t)slice * N * K;
    if (n_bufs == 2 && idx >= 2) BCI_CUDA(cudaStreamWaitEvent(split_stream, ev_gemm[buf], 0));  // planes free again
    split_kernel<<<dim3(ta + tb, 1, (unsigned)nb), 256, 0, sstream>>>(
        reinterpret_cast<const SplitReq*>(d_sreq) + b0, a_hi, a_lo, bt_hi, bt_lo, M, N, K, ta);
    BCI_CHECK_LAUNCH();
    if (n_bufs == 2) {
      BCI_CUDA(cudaEventRecord(ev_split[buf], split_stream));
      BCI_CUDA(cudaStreamWaitEvent(c.stream, ev_split[buf], 0));
    }
    CUtensorMap m_ah, m_al, m_bh, m_bl;
    if ((rc = make_map(&m_ah, a_hi, M, K, nb, TBM)) != BCI_OK) return rc;
    if ((rc = make_map(&m_al, a_lo, M, K, nb, TBM)) != BCI_OK) return rc;
    if ((rc = make_map(&m_bh, bt_hi, N, K, nb, tile_n)) != BCI_OK) return rc;
    if ((rc = make_map(&m_bl, bt_lo, N, K, nb, tile_n)) != BCI_OK) return rc;
    const dim3 grid((unsigned)(N / tile_n), (unsigned)(M / TBM), (unsigned)nb);
    if (tile_n == 256)
      gemm3x_kernel<256><<<grid, GemmCfg<256>::kThr, GemmCfg<256>::kSmem, c.stream>>>(
          m_ah, m_al, m_bh, m_bl, reinterpret_cast<const GemmOut*>(d_outs) + b0, N, K, raster_group, prefetch_kb);
    else
      gemm3x_kernel<128><<<grid, GemmCfg<128>::kThr, GemmCfg<128>::kSmem, c.stream>>>(
          m_ah, m_al, m_bh, m_bl, reinterpret_cast<const GemmOut*>(d_outs) + b0, N, K, raster_group, prefetch_kb);
    BCI_CHECK_LAUNCH();
    if (n_bufs == 2) BCI_CUDA(cudaEventRecord(ev_gemm[buf], c.stream));
  }
  return BCI_OK;
}

}  // namespace bci
