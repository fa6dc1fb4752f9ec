// Shared device helpers: per-precision math, limits shared with include/pyves_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pyves_b200.h"

#define PV_DEV __device__ __forceinline__

namespace pv {

constexpr int kMaxK1 = PYVES_MAX_DEGREE + 1;   // basis functions per axis
constexpr int kMaxPath = PYVES_MAX_PATH;

template <typename T> struct Math;

template <> struct Math<float> {
  static PV_DEV float fma(float a, float b, float c) { return fmaf(a, b, c); }
  static PV_DEV float exp(float x) { return __expf(x); }
  static PV_DEV float log(float x) { return __logf(x); }
  static PV_DEV float tanh(float x) { return tanhf(x); }
  static PV_DEV float min(float a, float b) { return fminf(a, b); }
  static PV_DEV float max(float a, float b) { return fmaxf(a, b); }
};

template <> struct Math<double> {
  static PV_DEV double fma(double a, double b, double c) { return ::fma(a, b, c); }
  static PV_DEV double exp(double x) { return ::exp(x); }
  static PV_DEV double log(double x) { return ::log(x); }
  static PV_DEV double tanh(double x) { return ::tanh(x); }
  static PV_DEV double min(double a, double b) { return ::fmin(a, b); }
  static PV_DEV double max(double a, double b) { return ::fmax(a, b); }
};

// Legendre recursion coefficients: P_{n+1} = A_n s P_n - B_n P_{n-1}, A_n = (2n+1)/(n+1),
// B_n = n/(n+1); derivative P'_{n+1} = P'_{n-1} + C_n P_n, C_n = 2n+1  (ves/basis.py:111-112).
template <typename T> PV_DEV constexpr T leg_a(int n) { return T(2 * n + 1) / T(n + 1); }
template <typename T> PV_DEV constexpr T leg_b(int n) { return T(n) / T(n + 1); }
template <typename T> PV_DEV constexpr T leg_c(int n) { return T(2 * n + 1); }

}  // namespace pv
