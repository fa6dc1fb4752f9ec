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

// Two values updated by one instruction.  For float this is Blackwell's packed FP32x2 datapath
// (PTX fma.rn.f32x2 -> SASS FFMA2, new on sm_100): one issue slot performs two FMAs, and a scalar
// operand is broadcast to both halves, so (A, B) += w * (P, P') costs a single instruction.
// For double it is simply two DFMAs.
template <typename T> struct Pair;

template <> struct Pair<float> {
  unsigned long long v;
  static PV_DEV Pair make(float lo, float hi) {
    Pair p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p.v) : "f"(lo), "f"(hi));
    return p;
  }
  PV_DEV float lo() const { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
  PV_DEV float hi() const { return __uint_as_float((unsigned)(v >> 32)); }
  // a * w + c with the scalar w broadcast to both halves
  static PV_DEV Pair fma_bcast(Pair a, float w, Pair c) {
    Pair d, ww = make(w, w);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(ww.v), "l"(c.v));
    return d;
  }
  static PV_DEV Pair fma(Pair a, Pair b, Pair c) {
    Pair d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return d;
  }
  // acc += a * w in place (the accumulator is tied to one register pair)
  static PV_DEV void acc_bcast(Pair& acc, Pair a, float w) {
    const Pair ww = make(w, w);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.v) : "l"(a.v), "l"(ww.v));
  }
  static PV_DEV Pair mul(Pair a, Pair b) {
    Pair d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
  }
  static PV_DEV Pair mul_bcast(Pair a, float w) { return mul(a, make(w, w)); }
  static PV_DEV Pair add(Pair a, Pair b) {
    Pair d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
  }
  static PV_DEV Pair add_bcast(Pair a, float w) { return add(a, make(w, w)); }
};

template <> struct Pair<double> {
  double x, y;
  static PV_DEV Pair make(double lo, double hi) { return Pair{lo, hi}; }
  PV_DEV double lo() const { return x; }
  PV_DEV double hi() const { return y; }
  static PV_DEV Pair fma_bcast(Pair a, double w, Pair c) { return Pair{::fma(a.x, w, c.x), ::fma(a.y, w, c.y)}; }
  static PV_DEV Pair fma(Pair a, Pair b, Pair c) { return Pair{::fma(a.x, b.x, c.x), ::fma(a.y, b.y, c.y)}; }
  static PV_DEV void acc_bcast(Pair& acc, Pair a, double w) { acc.x = ::fma(a.x, w, acc.x); acc.y = ::fma(a.y, w, acc.y); }
  static PV_DEV Pair mul(Pair a, Pair b) { return Pair{a.x * b.x, a.y * b.y}; }
  static PV_DEV Pair mul_bcast(Pair a, double w) { return Pair{a.x * w, a.y * w}; }
  static PV_DEV Pair add(Pair a, Pair b) { return Pair{a.x + b.x, a.y + b.y}; }
  static PV_DEV Pair add_bcast(Pair a, double w) { return Pair{a.x + w, a.y + w}; }
};

// 128-bit broadcast load of loop-invariant coefficients from shared memory.  It is `asm volatile` on
// purpose: the coefficients do not change during a launch, so the compiler would otherwise hoist all
// of them out of the time loop and keep (then spill) hundreds of registers.
PV_DEV void lds_vec(const float* p, float (&v)[4]) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
}
PV_DEV void lds_vec(const double* p, double (&v)[2]) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a));
}

// Legendre recursion coefficients: P_{n+1} = A_n s P_n - B_n P_{n-1}, A_n = (2n+1)/(n+1),
// B_n = n/(n+1); derivative P'_{n+1} = P'_{n-1} + C_n P_n, C_n = 2n+1  (ves/basis.py:111-112).
template <typename T> PV_DEV constexpr T leg_a(int n) { return T(2 * n + 1) / T(n + 1); }
template <typename T> PV_DEV constexpr T leg_b(int n) { return T(n) / T(n + 1); }
template <typename T> PV_DEV constexpr T leg_c(int n) { return T(2 * n + 1); }

}  // namespace pv
