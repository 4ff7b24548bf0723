// bci_ewise.cu — K1: batched elementwise ufuncs.
// Work the reference does inside numpy's ufunc loops when agent code runs
// `numpy.square(large_array)` (examples/benchmark-numpy.py:21) or `x*y+1.0`.
//
// HBM-bound: algorithmic bytes per request = (inputs + 1) * n * sizeof(T).
// One CTA = one 32 KiB tile per operand: 256 threads x 4 x 256-bit accesses, all
// loads issued before the first use (4..8 LDG.256 in flight per thread), then the
// arithmetic, then 4 STG.256.  Tails and unaligned views take a scalar path.
#include "bci_ops.cuh"
#include <vector>
#include <cstring>

namespace bci {

constexpr int kEwUnroll = 4;
template <typename T> constexpr size_t ew_tile_elems() { return (size_t)kThreads * kEwUnroll * (32 / sizeof(T)); }

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads)
ewise_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B) {
  constexpr int VN = 32 / sizeof(T);
  constexpr size_t TILE = (size_t)kThreads * kEwUnroll * VN;
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * TILE;
  const size_t rem = (size_t)q.n - base;
  const T* __restrict__ a = reinterpret_cast<const T*>(q.in0) + base;
  const T* __restrict__ b = q.in1 ? reinterpret_cast<const T*>(q.in1) + base : nullptr;
  T* __restrict__ o = reinterpret_cast<T*>(q.out) + base;
  const T sc = scalar_as<T>(q.s0);
  constexpr bool unary = ew_is_unary<OP>::value;

  const bool fast = rem >= TILE && is_aligned32(a) && is_aligned32(o) && (b == nullptr || is_aligned32(b));
  if (fast) {
    VecOf<T> va[kEwUnroll], vb[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u)
      va[u].u = ldg256(a + ((size_t)u * kThreads + threadIdx.x) * VN);
    if (!unary && b != nullptr) {
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u)
        vb[u].u = ldg256(b + ((size_t)u * kThreads + threadIdx.x) * VN);
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
#pragma unroll
        for (int j = 0; j < VN; ++j) va[u].v[j] = ew_apply<T, OP>(va[u].v[j], vb[u].v[j]);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
#pragma unroll
        for (int j = 0; j < VN; ++j) va[u].v[j] = ew_apply<T, OP>(va[u].v[j], sc);
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u)
      stg256(o + ((size_t)u * kThreads + threadIdx.x) * VN, va[u].u);
  } else {
    const size_t m = rem < TILE ? rem : TILE;
    for (size_t i = threadIdx.x; i < m; i += kThreads) {
      const T x = a[i];
      const T y = (!unary && b != nullptr) ? b[i] : sc;
      o[i] = ew_apply<T, OP>(x, y);
    }
  }
}


// ---- array creation: numpy.zeros/ones/full (CONST), numpy.arange (ARANGE), numpy.linspace (LINSPACE)
// Write-only, HBM-bound: n*sizeof(T) bytes.  ARANGE follows numpy's @TYPE@_fill
// (numpy/_core/src/multiarray/arraytypes.c.src): buffer[i] = start + i*delta with the index
// converted to the array type, product rounded, sum rounded.  LINSPACE follows
// numpy/_core/function_base.py: y = arange(num) * step + start, last element = stop.
template <typename T, int KIND>
__global__ void __launch_bounds__(kThreads)
fill_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B) {
  constexpr int VN = 32 / sizeof(T);
  constexpr size_t TILE = (size_t)kThreads * kEwUnroll * VN;
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * TILE;
  const size_t rem = (size_t)q.n - base;
  T* __restrict__ o = reinterpret_cast<T*>(q.out) + base;
  const T a = scalar_as<T>(q.s0), b = scalar_as<T>(q.s1);
  const bool has_last = (q.aux & 1) != 0;
  T last;
  {
    bci_scalar c;
    memcpy(&c, &q.in0, 8);                       // third scalar travels in the unused in0 slot
    last = scalar_as<T>(c);
  }
  auto value = [&](size_t idx) -> T {
    if constexpr (KIND == BCI_FILL_CONST) return a;
    else if constexpr (KIND == BCI_FILL_ARANGE) {
      if (has_last && idx == 1) return last;     // numpy stores start+step itself, the fill loop starts at 2
      if constexpr (is_float_t<T>::value) return fadd_(a, fmul_((T)(long long)idx, b));
      else return fadd_(a, fmul_((T)idx, b));
    } else {
      if (has_last && idx + 1 == (size_t)q.n) return last;
      if constexpr (is_float_t<T>::value) return fadd_(fmul_((T)(long long)idx, b), a);
      else return a;
    }
  };
  if (rem >= TILE && is_aligned32(o)) {
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      VecOf<T> v;
      const size_t i0 = ((size_t)u * kThreads + threadIdx.x) * VN;
#pragma unroll
      for (int j = 0; j < VN; ++j) v.v[j] = value(base + i0 + j);
      stg256(o + i0, v.u);
    }
  } else {
    const size_t m = rem < TILE ? rem : TILE;
    for (size_t i = threadIdx.x; i < m; i += kThreads) o[i] = value(base + i);
  }
}

// ---- numpy.ndarray.astype between device dtypes (value-preserving or round-to-nearest) -----------
template <typename S, typename D>
__global__ void __launch_bounds__(kThreads)
cast_kernel(const Req* __restrict__ reqs, const unsigned long long* __restrict__ tile_start, int B) {
  constexpr size_t TILE = (size_t)kThreads * 16;
  const unsigned long long tile = blockIdx.x;
  const int r = find_req(tile_start, B, tile);
  const Req q = reqs[r];
  const size_t base = (size_t)(tile - q.tile_start) * TILE;
  const size_t rem = (size_t)q.n - base;
  const S* __restrict__ a = reinterpret_cast<const S*>(q.in0) + base;
  D* __restrict__ o = reinterpret_cast<D*>(q.out) + base;
  constexpr int SV = 32 / sizeof(S), DV = 32 / sizeof(D);
  if (rem >= TILE && is_aligned32(a) && is_aligned32(o)) {
    // each thread converts 16 consecutive elements: whole 256-bit vectors on both sides
    const size_t e0 = (size_t)threadIdx.x * 16;
    S in[16];
    D out[16];
#pragma unroll
    for (int k = 0; k < 16 / SV; ++k) {
      const U256 u = ldg256(a + e0 + k * SV);
      memcpy(&in[k * SV], &u, 32);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) out[k] = static_cast<D>(in[k]);
#pragma unroll
    for (int k = 0; k < 16 / DV; ++k) {
      U256 u;
      memcpy(&u, &out[k * DV], 32);
      stg256(o + e0 + k * DV, u);
    }
  } else {
    const size_t m = rem < TILE ? rem : TILE;
    for (size_t i = threadIdx.x; i < m; i += kThreads) o[i] = static_cast<D>(a[i]);
  }
}

template <typename T, int KIND>
static int launch_fill(Req* reqs, int B) {
  Ctx& c = ctx();
  Batch bt;
  int rc = upload_batch(reqs, B, ew_tile_elems<T>(), &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_fill: batch too large for one launch"); return BCI_ERR_INVALID; }
  fill_kernel<T, KIND><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

template <typename T>
static int dispatch_fill(int kind, Req* reqs, int B) {
  switch (kind) {
    case BCI_FILL_CONST: return launch_fill<T, BCI_FILL_CONST>(reqs, B);
    case BCI_FILL_ARANGE: return launch_fill<T, BCI_FILL_ARANGE>(reqs, B);
    case BCI_FILL_LINSPACE:
      if constexpr (sizeof(T) == 8 && is_float_t<T>::value) return launch_fill<T, BCI_FILL_LINSPACE>(reqs, B);
      break;
  }
  set_error("bci_fill: kind %d is not supported for this dtype", kind);
  return BCI_ERR_UNSUPPORTED;
}

template <typename S, typename D>
static int launch_cast(Req* reqs, int B) {
  Ctx& c = ctx();
  Batch bt;
  int rc = upload_batch(reqs, B, (size_t)kThreads * 16, &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_cast: batch too large for one launch"); return BCI_ERR_INVALID; }
  cast_kernel<S, D><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

template <typename T, int OP>
static int launch_ewise(Req* reqs, int B) {
  Ctx& c = ctx();
  Batch bt;
  int rc = upload_batch(reqs, B, ew_tile_elems<T>(), &bt);
  if (rc != BCI_OK) return rc;
  if (bt.total_tiles == 0) return BCI_OK;
  if (bt.total_tiles > 0x7fffffffull) { set_error("bci_ewise: batch too large for one launch"); return BCI_ERR_INVALID; }
  ewise_kernel<T, OP><<<(unsigned)bt.total_tiles, kThreads, 0, c.stream>>>(bt.d_reqs, bt.d_tile_start, B);
  BCI_CHECK_LAUNCH();
  return BCI_OK;
}

template <typename T>
static int dispatch_op(int op, Req* reqs, int B) {
  switch (op) {
    case BCI_EW_ADD: return launch_ewise<T, BCI_EW_ADD>(reqs, B);
    case BCI_EW_SUB: return launch_ewise<T, BCI_EW_SUB>(reqs, B);
    case BCI_EW_MUL: return launch_ewise<T, BCI_EW_MUL>(reqs, B);
    case BCI_EW_MAXIMUM: return launch_ewise<T, BCI_EW_MAXIMUM>(reqs, B);
    case BCI_EW_MINIMUM: return launch_ewise<T, BCI_EW_MINIMUM>(reqs, B);
    case BCI_EW_RSUB: return launch_ewise<T, BCI_EW_RSUB>(reqs, B);
    case BCI_EW_SQUARE: return launch_ewise<T, BCI_EW_SQUARE>(reqs, B);
    case BCI_EW_ABS: return launch_ewise<T, BCI_EW_ABS>(reqs, B);
    case BCI_EW_NEG: return launch_ewise<T, BCI_EW_NEG>(reqs, B);
    default: break;
  }
  if constexpr (is_float_t<T>::value) {
    switch (op) {
      case BCI_EW_DIV: return launch_ewise<T, BCI_EW_DIV>(reqs, B);
      case BCI_EW_RDIV: return launch_ewise<T, BCI_EW_RDIV>(reqs, B);
      case BCI_EW_SQRT: return launch_ewise<T, BCI_EW_SQRT>(reqs, B);
      case BCI_EW_EXP: return launch_ewise<T, BCI_EW_EXP>(reqs, B);
      case BCI_EW_LOG: return launch_ewise<T, BCI_EW_LOG>(reqs, B);
      case BCI_EW_SIN: return launch_ewise<T, BCI_EW_SIN>(reqs, B);
      case BCI_EW_COS: return launch_ewise<T, BCI_EW_COS>(reqs, B);
      case BCI_EW_TANH: return launch_ewise<T, BCI_EW_TANH>(reqs, B);
      case BCI_EW_FLOOR: return launch_ewise<T, BCI_EW_FLOOR>(reqs, B);
      case BCI_EW_CEIL: return launch_ewise<T, BCI_EW_CEIL>(reqs, B);
      case BCI_EW_LOG1P: return launch_ewise<T, BCI_EW_LOG1P>(reqs, B);
      case BCI_EW_EXPM1: return launch_ewise<T, BCI_EW_EXPM1>(reqs, B);
      default: break;
    }
  }
  set_error("bci_ewise: op %d is not supported for this dtype", op);
  return BCI_ERR_UNSUPPORTED;
}

}  // namespace bci

using namespace bci;

extern "C" int bci_ewise(int op, int dtype, int B, const void* const* in0, const void* const* in1,
                         const bci_scalar* scalar, void* const* out, const size_t* n) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !in0 || !out || !n) { set_error("bci_ewise: bad arguments"); return BCI_ERR_INVALID; }
  if (elem_size(dtype) == 0) { set_error("bci_ewise: unknown dtype %d", dtype); return BCI_ERR_INVALID; }
  const bool unary = op >= 32;
  std::vector<Req> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    Req& q = reqs[i];
    q.in0 = in0[i];
    q.in1 = (in1 && !unary) ? in1[i] : nullptr;
    q.out = out[i];
    q.n = n[i];
    q.s0.i = 0;
    if (scalar) q.s0 = scalar[i];
    q.s1.i = 0;
    q.aux = 0;
    if (n[i] && (!in0[i] || !out[i])) { set_error("bci_ewise: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
    if (!unary && q.in1 == nullptr && !scalar) {
      set_error("bci_ewise: request %d has neither in1 nor a scalar", i);
      return BCI_ERR_INVALID;
    }
  }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  switch (dtype) {
    case BCI_F32: return dispatch_op<float>(op, reqs.data(), B);
    case BCI_F64: return dispatch_op<double>(op, reqs.data(), B);
    case BCI_I32: return dispatch_op<int>(op, reqs.data(), B);
    case BCI_I64: return dispatch_op<long long>(op, reqs.data(), B);
  }
  return BCI_ERR_INVALID;
}

extern "C" int bci_fill(int kind, int dtype, int B, void* const* out, const size_t* n, const bci_scalar* a,
                        const bci_scalar* b, const bci_scalar* cc) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !out || !n || !a) { set_error("bci_fill: bad arguments"); return BCI_ERR_INVALID; }
  if (elem_size(dtype) == 0) { set_error("bci_fill: unknown dtype %d", dtype); return BCI_ERR_INVALID; }
  if (kind != BCI_FILL_CONST && !b) { set_error("bci_fill: kind %d needs the step scalar", kind); return BCI_ERR_INVALID; }
  std::vector<Req> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    Req& q = reqs[i];
    memset(&q, 0, sizeof q);
    q.out = out[i];
    q.n = n[i];
    q.s0 = a[i];
    if (b) q.s1 = b[i];
    if (cc && kind != BCI_FILL_CONST) { memcpy(&q.in0, &cc[i], 8); q.aux = 1; }
    if (n[i] && !out[i]) { set_error("bci_fill: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
  }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  switch (dtype) {
    case BCI_F32: return dispatch_fill<float>(kind, reqs.data(), B);
    case BCI_F64: return dispatch_fill<double>(kind, reqs.data(), B);
    case BCI_I32: return dispatch_fill<int>(kind, reqs.data(), B);
    case BCI_I64: return dispatch_fill<long long>(kind, reqs.data(), B);
  }
  return BCI_ERR_INVALID;
}

extern "C" int bci_cast(int src_dtype, int dst_dtype, int B, const void* const* in, void* const* out, const size_t* n) {
  BCI_REQUIRE_INIT();
  if (B <= 0 || !in || !out || !n) { set_error("bci_cast: bad arguments"); return BCI_ERR_INVALID; }
  if (elem_size(src_dtype) == 0 || elem_size(dst_dtype) == 0) { set_error("bci_cast: unknown dtype"); return BCI_ERR_INVALID; }
  std::vector<Req> reqs((size_t)B);
  for (int i = 0; i < B; ++i) {
    Req& q = reqs[i];
    memset(&q, 0, sizeof q);
    q.in0 = in[i];
    q.out = out[i];
    q.n = n[i];
    if (n[i] && (!in[i] || !out[i])) { set_error("bci_cast: NULL buffer in request %d", i); return BCI_ERR_INVALID; }
  }
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  const int key = src_dtype * 4 + dst_dtype;
  switch (key) {
    case BCI_F32 * 4 + BCI_F64: return launch_cast<float, double>(reqs.data(), B);
    case BCI_F64 * 4 + BCI_F32: return launch_cast<double, float>(reqs.data(), B);
    case BCI_I32 * 4 + BCI_I64: return launch_cast<int, long long>(reqs.data(), B);
    case BCI_I64 * 4 + BCI_I32: return launch_cast<long long, int>(reqs.data(), B);
    case BCI_I32 * 4 + BCI_F32: return launch_cast<int, float>(reqs.data(), B);
    case BCI_I32 * 4 + BCI_F64: return launch_cast<int, double>(reqs.data(), B);
    case BCI_I64 * 4 + BCI_F32: return launch_cast<long long, float>(reqs.data(), B);
    case BCI_I64 * 4 + BCI_F64: return launch_cast<long long, double>(reqs.data(), B);
  }
  set_error("bci_cast: dtype %d -> %d is left to numpy", src_dtype, dst_dtype);
  return BCI_ERR_UNSUPPORTED;
}
