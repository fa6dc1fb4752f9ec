// Analytic potential energy surfaces: gradient (and energy) per walker.
// Replaces the Lepton-compiled openmm.CustomExternalForce of ves/potentials.py
// (expressions at :59,72,135,151,183,229,253-254,283-286).  CPU twin: oracle/potentials.py.
#pragma once
#include "common.cuh"

namespace pv {

template <typename T> struct PotParams {
  int kind;
  T p[11];
};

// Each functor: grad<WANT_E>(P, x, y, e, gx, gy) ACCUMULATES dU/dx, dU/dy (and U when WANT_E).

// Two walkers per thread in packed form (Pair<T>: lo = walker a, hi = walker b; FP32x2 instructions for float):
// grad2 returns the gradient pairs.  Functors without a packed implementation are evaluated half by half.
template <class Pot, typename T, typename = void> struct PotPair {
  static PV_DEV void grad2(const PotParams<T>& P, Pair<T> x, Pair<T> y, Pair<T>& gx, Pair<T>& gy) {
    T e = T(0), ax = T(0), ay = T(0), bx = T(0), by = T(0);
    Pot::template grad<false>(P, x.lo(), y.lo(), e, ax, ay);
    Pot::template grad<false>(P, x.hi(), y.hi(), e, bx, by);
    gx = Pair<T>::make(ax, bx);
    gy = Pair<T>::make(ay, by);
  }
};

template <typename T> struct PotDoubleWell2D {   // 10 (x^4 - 4 x^2 + 0.7 x) + 5 y^2
  static PV_DEV void grad2(const PotParams<T>&, Pair<T> x, Pair<T> y, Pair<T>& gx, Pair<T>& gy) {
    const Pair<T> x2 = Pair<T>::mul(x, x);
    const Pair<T> t = Pair<T>::fma_bcast(x2, T(40), Pair<T>::make(T(-80), T(-80)));
    gx = Pair<T>::fma(x, t, Pair<T>::make(T(7), T(7)));
    gy = Pair<T>::mul_bcast(y, T(10));
  }
  template <bool WANT_E> static PV_DEV void grad(const PotParams<T>&, T x, T y, T& e, T& gx, T& gy) {
    const T x2 = x * x;
    gx += x * Math<T>::fma(T(40), x2, T(-80)) + T(7);
    gy = Math<T>::fma(T(10), y, gy);
    if (WANT_E) e += T(10) * (x2 * (x2 - T(4)) + T(0.7) * x) + T(5) * y * y;
  }
};

template <typename T> struct PotSlipBond2D {     // ((y-y0)^2/ys - ysh)^2 + (x-y-xy0)^2/xys - fx x - fy y
  static PV_DEV void grad2(const PotParams<T>& P, Pair<T> x, Pair<T> y, Pair<T>& gx, Pair<T>& gy) {
    using Q = Pair<T>;
    const Q dy = Q::add_bcast(y, -P.p[2]);
    const Q t = Q::fma_bcast(Q::mul(dy, dy), P.p[3], Q::make(-P.p[4], -P.p[4]));
    const Q u = Q::add_bcast(Q::add(x, Q::mul_bcast(y, T(-1))), -P.p[5]);
    const Q du = Q::mul_bcast(u, T(2) * P.p[6]);
    gx = Q::add_bcast(du, -P.p[0]);
    gy = Q::add_bcast(Q::add(Q::mul(Q::mul_bcast(t, T(4) * P.p[3]), dy), Q::mul_bcast(du, T(-1))), -P.p[1]);
  }
  template <bool WANT_E> static PV_DEV void grad(const PotParams<T>& P, T x, T y, T& e, T& gx, T& gy) {
    // p: fx fy y0 1/ys ysh xy0 1/xys
    const T dy = y - P.p[2];
    const T t = Math<T>::fma(dy * dy, P.p[3], -P.p[4]);
    const T u = x - y - P.p[5];
    const T du = T(2) * u * P.p[6];
    gx += du - P.p[0];
    gy += T(4) * t * dy * P.p[3] - du - P.p[1];
    if (WANT_E) e += t * t + u * u * P.p[6] - P.p[0] * x - P.p[1] * y;
  }
};

template <typename T> struct PotSzaboBerezhkovskii {   // Omega2/2 (x-y)^2 + piecewise(x)
  template <bool WANT_E> static PV_DEV void grad(const PotParams<T>& P, T x, T y, T& e, T& gx, T& gy) {
    // p: x0 omega2 Omega2 Delta
    const T x0 = P.p[0], w2 = P.p[1], W2 = P.p[2];
    const T half = T(0.5) * x0;
    // Lepton select(step(x + x0/2), select(step(x - x0/2), right, middle), left)
    const T shift = (x < -half) ? x0 : ((x >= half) ? -x0 : T(0));
    const bool mid = (shift == T(0));
    const T xs = x + shift;
    const T coupling = W2 * (x - y);
    gx += (mid ? -w2 * x : w2 * xs) + coupling;
    gy -= coupling;
    if (WANT_E) e += (mid ? T(-0.5) * w2 * x * x : T(0.5) * w2 * xs * xs - P.p[3]) + T(0.5) * W2 * (x - y) * (x - y);
  }
};

// functors with a packed implementation
template <typename T> struct PotPair<PotDoubleWell2D<T>, T, void> {
  static PV_DEV void grad2(const PotParams<T>& P, Pair<T> x, Pair<T> y, Pair<T>& gx, Pair<T>& gy) { PotDoubleWell2D<T>::grad2(P, x, y, gx, gy); }
};
template <typename T> struct PotPair<PotSlipBond2D<T>, T, void> {
  static PV_DEV void grad2(const PotParams<T>& P, Pair<T> x, Pair<T> y, Pair<T>& gx, Pair<T>& gy) { PotSlipBond2D<T>::grad2(P, x, y, gx, gy); }
};

// Runtime-dispatched potential for every kind (warp-uniform switch).
template <typename T> struct PotGeneric {
  template <bool WANT_E> static PV_DEV void grad(const PotParams<T>& P, T x, T y, T& e, T& gx, T& gy) {
    switch (P.kind) {
      case PYVES_POT_SINGLE_WELL_1D:
        gx += T(2) * x;
        gy += T(2000) * y;
        if (WANT_E) e += x * x + T(1000) * y * y;
        break;
      case PYVES_POT_DOUBLE_WELL_1D: {
        const T x2 = x * x;
        gx += x * (T(4) * x2 - T(8)) + T(0.7);
        gy += T(2000) * y;
        if (WANT_E) e += x2 * (x2 - T(4)) + T(0.7) * x + T(1000) * y * y;
      } break;
      case PYVES_POT_SINGLE_WELL_2D:
        gx += T(2) * x;
        gy += T(2) * y;
        if (WANT_E) e += x * x + y * y;
        break;
      case PYVES_POT_DOUBLE_WELL_2D: PotDoubleWell2D<T>::template grad<WANT_E>(P, x, y, e, gx, gy); break;
      case PYVES_POT_SLIP_BOND_2D: PotSlipBond2D<T>::template grad<WANT_E>(P, x, y, e, gx, gy); break;
      case PYVES_POT_SZABO_BEREZHKOVSKII: PotSzaboBerezhkovskii<T>::template grad<WANT_E>(P, x, y, e, gx, gy); break;
      case PYVES_POT_CATCH_BOND_2D: {
        // -log(exp(-A) + exp(-B)) - fx x - fy y;  p: fx fy y0 1/ys ysh xy0 1/xys gx0 1/gxs gy0 1/gys
        const T dy = y - P.p[2];
        const T t = Math<T>::fma(dy * dy, P.p[3], -P.p[4]);
        const T u = x - y - P.p[5];
        const T A = t * t + u * u * P.p[6];
        const T dAx = T(2) * u * P.p[6];
        const T dAy = T(4) * t * dy * P.p[3] - dAx;
        const T bx = x - P.p[7], by = y - P.p[9];
        const T B = bx * bx * P.p[8] + by * by * P.p[10];
        const T dBx = T(2) * bx * P.p[8], dBy = T(2) * by * P.p[10];
        const T m = Math<T>::min(A, B);
        const T ea = Math<T>::exp(m - A), eb = Math<T>::exp(m - B);
        const T inv = T(1) / (ea + eb);
        gx += (ea * dAx + eb * dBx) * inv - P.p[0];
        gy += (ea * dAy + eb * dBy) * inv - P.p[1];
        if (WANT_E) e += m - Math<T>::log(ea + eb) - P.p[0] * x - P.p[1] * y;
      } break;
      case PYVES_POT_MULLER_BROWN: {
        const T A[4] = {T(-200), T(-100), T(-170), T(15)};
        const T a[4] = {T(-1), T(-1), T(-6.5), T(0.7)};
        const T b[4] = {T(0), T(0), T(11), T(0.6)};
        const T c[4] = {T(-10), T(-10), T(-6.5), T(0.7)};
        const T X[4] = {T(1), T(0), T(-0.5), T(-1)};
        const T Y[4] = {T(0), T(0.5), T(1.5), T(1)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const T dx = x - X[i], dy = y - Y[i];
          const T ee = A[i] * Math<T>::exp(a[i] * dx * dx + b[i] * dx * dy + c[i] * dy * dy);
          gx += ee * (T(2) * a[i] * dx + b[i] * dy);
          gy += ee * (b[i] * dx + T(2) * c[i] * dy);
          if (WANT_E) e += ee;
        }
      } break;
      default: break;
    }
  }
};

}  // namespace pv
