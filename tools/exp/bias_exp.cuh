// Experimental 16 x 16 functors for the step-kernel harness (tools/exp/step16.cu).
#pragma once
#include "../../pyves_b200/csrc/bias.cuh"

namespace pv {

// j-outer ("column sweep") evaluation: the y recursion produces Q_j = (P_j(y'), P'_j(y')) one at
// a time and every Q_j immediately updates all KX row accumulators (A_i, B_i) += w_ij Q_j, so
//  * there are KX independent FFMA2 chains (no dependent-issue waits),
//  * the Q_j never have to be live together (KX pairs of accumulators instead of KY pairs of Q),
//  * the weights of one column are contiguous (wt[j][i]): four 128-bit uniform loads per column.
// Row i = 0 (P_0(x) = 1, P'_0(x) = 0) only needs B_0 and column j = 0 (Q_0 = (1, 0)) only feeds
// the A_i: both are done with scalar FMAs.
template <typename T, int KX, int KY> struct Legendre2DStaticTParams {
  T cx, ihx, cy, ihy;
  T wt[KY * KX];   // wt[j * KX + i] = w_ij
};

template <typename T, int KX, int KY, int GCH = 2> struct BiasLegendre2DStaticT {
  using Params = Legendre2DStaticTParams<T, KX, KY>;
  static size_t smem_bytes(const Params&) { return 0; }
  static PV_DEV void stage(const Params&, void*) {}
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void*, T x, T y, T& e, T& gx, T& gy) {
    bool inx, iny;
    const T sx = scale_clamp(x, P.cx, P.ihx, inx);
    const T sy = scale_clamp(y, P.cy, P.ihy, iny);
    Pair<T> AB[KX];
    // column 1: Q_1 = (sy, 1);  column 0 contributes (w_i0, 0)
    T pm = T(1), p = sy, dpm = T(0), dp = T(1);
    {
      const Pair<T> Q1 = Pair<T>::make(sy, T(1));
#pragma unroll
      for (int i = 0; i < KX; ++i) AB[i] = Pair<T>::fma_bcast(Q1, P.wt[1 * KX + i], Pair<T>::make(P.wt[i], T(0)));
    }
#pragma unroll
    for (int j = 2; j < KY; ++j) {
      const T pn = leg_a<T>(j - 1) * sy * p - leg_b<T>(j - 1) * pm;
      const T dpn = Math<T>::fma(leg_c<T>(j - 1), p, dpm);
      pm = p; p = pn; dpm = dp; dp = dpn;
      const Pair<T> Q = Pair<T>::make(p, dp);
#pragma unroll
      for (int i = 0; i < KX; ++i) AB[i] = Pair<T>::fma_bcast(Q, P.wt[j * KX + i], AB[i]);
    }
    // x sweep: G = sum_i (P'_i(x), P_i(x)) * (A_i, B_i)
    Pair<T> G[GCH];
    G[0] = Pair<T>::make(T(0), AB[0].hi());          // i = 0: (0, 1) * (A_0, B_0)
#pragma unroll
    for (int c = 1; c < GCH; ++c) G[c] = Pair<T>::make(T(0), T(0));
    T E = WANT_E ? AB[0].lo() : T(0);
    T pxm = T(1), px = sx, dpxm = T(0), dpx = T(1);
#pragma unroll
    for (int i = 1; i < KX; ++i) {
      if (i >= 2) {
        const T pn = leg_a<T>(i - 1) * sx * px - leg_b<T>(i - 1) * pxm;
        const T dpn = Math<T>::fma(leg_c<T>(i - 1), px, dpxm);
        pxm = px; px = pn; dpxm = dpx; dpx = dpn;
      }
      G[i % GCH] = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G[i % GCH]);
      if (WANT_E) E = Math<T>::fma(px, AB[i].lo(), E);
    }
    T glo = G[0].lo(), ghi = G[0].hi();
#pragma unroll
    for (int c = 1; c < GCH; ++c) { glo += G[c].lo(); ghi += G[c].hi(); }
    gx += inx ? glo * P.ihx : T(0);
    gy += iny ? ghi * P.ihy : T(0);
    if (WANT_E) e += E;
  }
};

// Rolled column sweep: a runtime loop over the columns j (UNR columns per trip) keeps the register
// footprint small (KX accumulator pairs + one column of weights), so 3-4 blocks of 128 threads fit
// per SM.  SRC = 0: weights read from the constant bank with a register index (LDC.128);
// SRC = 1: weights staged in shared memory (LDS.128 broadcast).
template <typename T, int KX, int KY> struct Legendre2DRolledTParams {
  T cx, ihx, cy, ihy;
  T ca[KY], cb[KY], cc[KY];   // recursion coefficients a_n, b_n, c_n
  alignas(16) T wt[KY * KX];  // wt[j * KX + i] = w_ij
};

template <typename T, int KX, int KY, int UNR, int SRC> struct BiasLegendre2DRolledT {
  using Params = Legendre2DRolledTParams<T, KX, KY>;
  static constexpr int V = 16 / sizeof(T);
  struct alignas(16) Vec { T v[V]; };
  static size_t smem_bytes(const Params&) { return SRC == 1 ? sizeof(T) * KX * KY : 0; }
  static PV_DEV void stage(const Params& P, void* smem) {
    if (SRC == 1) {
      T* d = reinterpret_cast<T*>(smem);
      for (int i = threadIdx.x; i < KX * KY; i += blockDim.x) d[i] = P.wt[i];
    }
  }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    bool inx, iny;
    const T sx = scale_clamp(x, P.cx, P.ihx, inx);
    const T sy = scale_clamp(y, P.cy, P.ihy, iny);
    Pair<T> AB[KX];
    T pm = T(1), p = sy, dpm = T(0), dp = T(1);
    {
      const Pair<T> Q1 = Pair<T>::make(sy, T(1));
#pragma unroll
      for (int i = 0; i < KX; ++i) AB[i] = Pair<T>::fma_bcast(Q1, P.wt[1 * KX + i], Pair<T>::make(P.wt[i], T(0)));
    }
    static_assert((KY - 2) % UNR == 0, "UNR must divide KY - 2");
#pragma unroll 1
    for (int j0 = 2; j0 < KY; j0 += UNR) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int j = j0 + u;
        const T pn = P.ca[j - 1] * sy * p - P.cb[j - 1] * pm;
        const T dpn = Math<T>::fma(P.cc[j - 1], p, dpm);
        pm = p; p = pn; dpm = dp; dp = dpn;
        const Pair<T> Q = Pair<T>::make(p, dp);
#pragma unroll
        for (int iv = 0; iv < KX / V; ++iv) {
          Vec w;
          if (SRC == 1) {
            lds_vec(reinterpret_cast<const T*>(smem) + j * KX + iv * V, w.v);
          } else {
            w = *reinterpret_cast<const Vec*>(&P.wt[j * KX + iv * V]);
          }
#pragma unroll
          for (int q = 0; q < V; ++q) Pair<T>::acc_bcast(AB[iv * V + q], Q, w.v[q]);
        }
      }
    }
    Pair<T> G[2];
    G[0] = Pair<T>::make(T(0), AB[0].hi());
    G[1] = Pair<T>::make(T(0), T(0));
    T E = WANT_E ? AB[0].lo() : T(0);
    T pxm = T(1), px = sx, dpxm = T(0), dpx = T(1);
#pragma unroll
    for (int i = 1; i < KX; ++i) {
      if (i >= 2) {
        const T pn = leg_a<T>(i - 1) * sx * px - leg_b<T>(i - 1) * pxm;
        const T dpn = Math<T>::fma(leg_c<T>(i - 1), px, dpxm);
        pxm = px; px = pn; dpxm = dpx; dpx = dpn;
      }
      G[i % 2] = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G[i % 2]);
      if (WANT_E) E = Math<T>::fma(px, AB[i].lo(), E);
    }
    gx += inx ? (G[0].lo() + G[1].lo()) * P.ihx : T(0);
    gy += iny ? (G[0].hi() + G[1].hi()) * P.ihy : T(0);
    if (WANT_E) e += E;
  }
};

}  // namespace pv
