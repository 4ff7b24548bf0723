// bci_ops.cuh — per-element semantics shared by the elementwise and fused-reduction
// kernels.  Each op rounds exactly once, like the numpy ufunc loop it mirrors
// (numpy evaluates `np.sum(np.square(x))`, examples/benchmark-numpy.py:21, as separate
// correctly-rounded ufunc passes), so float results are bit-identical to numpy for
// add/sub/mul/div/sqrt/abs/neg/square/maximum/minimum; exp/log are within 2 ulp.
#pragma once
#include "bci_common.cuh"

namespace bci {

template <typename T> struct is_float_t { static constexpr bool value = false; };
template <> struct is_float_t<float> { static constexpr bool value = true; };
template <> struct is_float_t<double> { static constexpr bool value = true; };

template <typename T> __device__ __forceinline__ T scalar_as(const bci_scalar& s) {
  if constexpr (is_float_t<T>::value) return static_cast<T>(s.f);
  else return static_cast<T>(s.i);
}

// explicit single-rounding arithmetic: never contracted into FMA by nvcc
__device__ __forceinline__ float fadd_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double fadd_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fsub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fmul_(double a, double b) { return __dmul_rn(a, b); }
// integers wrap like numpy (two's complement); do it in unsigned to stay defined
__device__ __forceinline__ int fadd_(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int fsub_(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int fmul_(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__device__ __forceinline__ long long fadd_(long long a, long long b) {
  return (long long)((unsigned long long)a + (unsigned long long)b);
}
__device__ __forceinline__ long long fsub_(long long a, long long b) {
  return (long long)((unsigned long long)a - (unsigned long long)b);
}
__device__ __forceinline__ long long fmul_(long long a, long long b) {
  return (long long)((unsigned long long)a * (unsigned long long)b);
}

template <typename T> __device__ __forceinline__ T nan_max(T a, T b) {
  if constexpr (is_float_t<T>::value) {
    if (a != a) return a;
    if (b != b) return b;
  }
  return a > b ? a : b;
}
template <typename T> __device__ __forceinline__ T nan_min(T a, T b) {
  if constexpr (is_float_t<T>::value) {
    if (a != a) return a;
    if (b != b) return b;
  }
  return a < b ? a : b;
}

// integers: abs(INT_MIN) wraps to INT_MIN like numpy
template <typename T> __device__ __forceinline__ T abs_(T a) {
  return a < T(0) ? fsub_(T(0), a) : a;
}
template <> __device__ __forceinline__ float abs_<float>(float a) { return fabsf(a); }
template <> __device__ __forceinline__ double abs_<double>(double a) { return fabs(a); }

template <int OP> struct ew_is_unary { static constexpr bool value = OP >= 32; };

// binary: a OP b;  unary: OP a (b ignored)
template <typename T, int OP> __device__ __forceinline__ T ew_apply(T a, T b) {
  if constexpr (OP == BCI_EW_ADD) return fadd_(a, b);
  else if constexpr (OP == BCI_EW_SUB) return fsub_(a, b);
  else if constexpr (OP == BCI_EW_MUL) return fmul_(a, b);
  else if constexpr (OP == BCI_EW_MAXIMUM) return nan_max(a, b);
  else if constexpr (OP == BCI_EW_MINIMUM) return nan_min(a, b);
  else if constexpr (OP == BCI_EW_RSUB) return fsub_(b, a);
  else if constexpr (OP == BCI_EW_SQUARE) return fmul_(a, a);
  else if constexpr (OP == BCI_EW_ABS) return abs_(a);
  else if constexpr (OP == BCI_EW_NEG) {
    if constexpr (is_float_t<T>::value) return -a;
    else return fsub_(T(0), a);
  } else if constexpr (is_float_t<T>::value) {
    if constexpr (OP == BCI_EW_DIV) return a / b;     // IEEE (-prec-div=true default)
    else if constexpr (OP == BCI_EW_RDIV) return b / a;
    else if constexpr (OP == BCI_EW_SQRT) return sqrt(a);  // IEEE (-prec-sqrt=true default)
    else if constexpr (OP == BCI_EW_EXP) return exp(a);
    else if constexpr (OP == BCI_EW_LOG) return log(a);
    else if constexpr (OP == BCI_EW_SIN) return sin(a);     // full-range libdevice paths (no -use_fast_math)
    else if constexpr (OP == BCI_EW_COS) return cos(a);
    else if constexpr (OP == BCI_EW_TANH) return tanh(a);
    else if constexpr (OP == BCI_EW_FLOOR) return floor(a);
    else if constexpr (OP == BCI_EW_CEIL) return ceil(a);
    else if constexpr (OP == BCI_EW_LOG1P) return log1p(a);
    else if constexpr (OP == BCI_EW_EXPM1) return expm1(a);
    else return a;
  } else {
    return a;
  }
}

// runtime-dispatched variants for the fused reduction producers (uniform branch)
template <typename T> __device__ __forceinline__ T ew_apply_rt(int op, T a, T b) {
  switch (op) {
    case BCI_EW_ADD: return fadd_(a, b);
    case BCI_EW_SUB: return fsub_(a, b);
    case BCI_EW_MUL: return fmul_(a, b);
    case BCI_EW_MAXIMUM: return nan_max(a, b);
    case BCI_EW_MINIMUM: return nan_min(a, b);
    case BCI_EW_RSUB: return fsub_(b, a);
    default: return a;
  }
}
template <typename T, int PRE> __device__ __forceinline__ T pre_apply(T a) {
  if constexpr (PRE == BCI_PRE_SQUARE) return fmul_(a, a);
  else if constexpr (PRE == BCI_PRE_ABS) return abs_(a);
  else return a;
}

}  // namespace bci
