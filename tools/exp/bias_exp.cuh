// Experimental 16 x 16 functors for the step-kernel harness (tools/exp/step16.cu).
#pragma once
#include "../../pyves_b200/csrc/bias.cuh"

namespace pv {

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

// Column sweep for runtime sizes: weights transposed and zero-padded to wt[kyp][KXP] in shared memory
// (kyp = 2 + roundup(ky - 2, UNR)), recursion coefficients in shared memory as well.  A rolled loop
// over the columns (UNR per trip) keeps KXP accumulator pairs + UNR * 4 weights live.
template <typename T> struct Legendre2DSmemTParams {
  T cx, ihx, cy, ihy;
  int kx, ky;
  const T* w;   // device, row-major [kx][ky]
};

template <typename T, int KXP, int UNR> struct BiasLegendre2DSmemT {
  using Params = Legendre2DSmemTParams<T>;
  static constexpr int V = 16 / sizeof(T);
  static __host__ __device__ int kyp(int ky) { return 2 + (ky > 2 ? (ky - 2 + UNR - 1) / UNR * UNR : 0); }
  static size_t smem_bytes(const Params& P) { return (size_t)kyp(P.ky) * (KXP + 4) * sizeof(T); }
  static PV_DEV void stage(const Params& P, void* smem) {
    T* wt = reinterpret_cast<T*>(smem);
    const int n = kyp(P.ky);
    for (int idx = threadIdx.x; idx < n * KXP; idx += blockDim.x) {
      const int j = idx / KXP, i = idx % KXP;
      wt[idx] = (i < P.kx && j < P.ky) ? P.w[i * P.ky + j] : T(0);
    }
    T* co = wt + n * KXP;   // co[4 j + {0,1,2}] = a_{j-1}, b_{j-1}, c_{j-1}: P_j from P_{j-1}, P_{j-2}
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const T fn = T(j > 0 ? j - 1 : 0);
      co[4 * j + 0] = (T(2) * fn + T(1)) / (fn + T(1));
      co[4 * j + 1] = fn / (fn + T(1));
      co[4 * j + 2] = T(2) * fn + T(1);
      co[4 * j + 3] = T(0);
    }
  }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    bool inx, iny;
    const T sx = scale_clamp(x, P.cx, P.ihx, inx);
    const T sy = scale_clamp(y, P.cy, P.ihy, iny);
    const T* wt = reinterpret_cast<const T*>(smem);
    const int n = kyp(P.ky);
    const T* co = wt + n * KXP;
    Pair<T> AB[KXP];
    T pm = T(1), p = sy, dpm = T(0), dp = T(1);
    {
      const Pair<T> Q1 = Pair<T>::make(sy, T(1));
#pragma unroll
      for (int iv = 0; iv < KXP / V; ++iv) {
        T w0[V], w1[V];
        lds_vec(wt + iv * V, w0);
        lds_vec(wt + KXP + iv * V, w1);
#pragma unroll
        for (int q = 0; q < V; ++q) AB[iv * V + q] = Pair<T>::fma_bcast(Q1, w1[q], Pair<T>::make(w0[q], T(0)));
      }
    }
#pragma unroll 1
    for (int j0 = 2; j0 < n; j0 += UNR) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int j = j0 + u;
        T c[V];
        lds_vec(co + 4 * j, c);
        const T pn = c[0] * sy * p - c[1] * pm;
        const T dpn = Math<T>::fma(sizeof(T) == 4 ? c[2] : T(2 * j - 1), p, dpm);
        pm = p; p = pn; dpm = dp; dp = dpn;
        const Pair<T> Q = Pair<T>::make(p, dp);
#pragma unroll
        for (int iv = 0; iv < KXP / V; ++iv) {
          T w[V];
          lds_vec(wt + j * KXP + iv * V, w);
#pragma unroll
          for (int q = 0; q < V; ++q) Pair<T>::acc_bcast(AB[iv * V + q], Q, w[q]);
        }
      }
    }
    Pair<T> G0 = Pair<T>::make(T(0), AB[0].hi()), G1 = Pair<T>::make(T(0), T(0));
    T E = WANT_E ? AB[0].lo() : T(0);
    T pxm = T(1), px = sx, dpxm = T(0), dpx = T(1);
#pragma unroll
    for (int i = 1; i < KXP; ++i) {
      if (i >= 2) {
        const T pn = leg_a<T>(i - 1) * sx * px - leg_b<T>(i - 1) * pxm;
        const T dpn = Math<T>::fma(leg_c<T>(i - 1), px, dpxm);
        pxm = px; px = pn; dpxm = dpx; dpx = dpn;
      }
      if (i & 1) G1 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G1);
      else G0 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G0);
      if (WANT_E) E = Math<T>::fma(px, AB[i].lo(), E);
    }
    gx += inx ? (G0.lo() + G1.lo()) * P.ihx : T(0);
    gy += iny ? (G0.hi() + G1.hi()) * P.ihy : T(0);
    if (WANT_E) e += E;
  }
};

// Experiment: hybrid column sweep -- the first J0 columns take their weights from the constant bank (static),
// the others from shared memory (volatile LDS.128 right before use), to trade LDC / ADU traffic for LSU traffic.
template <int KX, int KY> struct Legendre2DStaticTHParams {
  float cx, ihx, cy, ihy;
  float wt[KY * KX];        // constant-bank copy
  const float* wdev;        // device copy of the same transposed weights (staged into shared memory)
};
template <int KX, int KY, int J0, bool TAIL = false, bool NV = false> struct BiasLegendre2DStaticTH {
  using T = float;
  using Params = Legendre2DStaticTHParams<KX, KY>;
  static size_t smem_bytes(const Params&) { return sizeof(T) * KX * KY; }
  static PV_DEV void stage(const Params& P, void* smem) {
    T* d = reinterpret_cast<T*>(smem);
    for (int i = threadIdx.x; i < KX * KY; i += blockDim.x) d[i] = P.wdev[i];
  }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    bool inx, iny;
    const T sx = scale_clamp(x, P.cx, P.ihx, inx);
    const T sy = scale_clamp(y, P.cy, P.ihy, iny);
    const T* ws = reinterpret_cast<const T*>(smem);
    Pair<T> AB[KX];
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
      if (TAIL ? (j >= KY - J0 + 2) : (j < J0)) {
#pragma unroll
        for (int i = 0; i < KX; ++i) AB[i] = Pair<T>::fma_bcast(Q, P.wt[j * KX + i], AB[i]);
      } else {
#pragma unroll
        for (int iv = 0; iv < KX / 4; ++iv) {
          T w[4];
          if (NV) {   // plain loads: the compiler may keep the weights in registers across the time loop
            const float4 v = *reinterpret_cast<const float4*>(ws + j * KX + iv * 4);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
          } else {
            lds_vec(ws + j * KX + iv * 4, w);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) Pair<T>::acc_bcast(AB[iv * 4 + q], Q, w[q]);
        }
      }
    }
    Pair<T> G0 = Pair<T>::make(T(0), AB[0].hi()), G1 = Pair<T>::make(T(0), T(0));
    T pxm = T(1), px = sx, dpxm = T(0), dpx = T(1);
#pragma unroll
    for (int i = 1; i < KX; ++i) {
      if (i >= 2) {
        const T pn = leg_a<T>(i - 1) * sx * px - leg_b<T>(i - 1) * pxm;
        const T dpn = Math<T>::fma(leg_c<T>(i - 1), px, dpxm);
        pxm = px; px = pn; dpxm = dpx; dpx = dpn;
      }
      if (i & 1) G1 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G1);
      else G0 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G0);
    }
    gx += inx ? (G0.lo() + G1.lo()) * P.ihx : T(0);
    gy += iny ? (G0.hi() + G1.hi()) * P.ihy : T(0);
  }
};

// Experiment: like StaticTH, but the shared-memory copy holds every weight TWICE (w, w), so the FFMA2s of the
// shared-memory columns take a plain 64-bit register operand instead of a broadcast scalar (2 KB, two weights
// per LDS.128).
template <int KX, int KY, int J0> struct BiasLegendre2DStaticTHD {
  using T = float;
  using Params = Legendre2DStaticTHParams<KX, KY>;
  static size_t smem_bytes(const Params&) { return sizeof(T) * KX * KY * 2; }
  static PV_DEV void stage(const Params& P, void* smem) {
    T* d = reinterpret_cast<T*>(smem);
    for (int i = threadIdx.x; i < KX * KY; i += blockDim.x) {
      const T v = P.wdev[i];
      d[2 * i] = v;
      d[2 * i + 1] = v;
    }
  }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    bool inx, iny;
    const T sx = scale_clamp(x, P.cx, P.ihx, inx);
    const T sy = scale_clamp(y, P.cy, P.ihy, iny);
    const T* ws = reinterpret_cast<const T*>(smem);
    Pair<T> AB[KX];
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
      if (j < J0) {
#pragma unroll
        for (int i = 0; i < KX; ++i) AB[i] = Pair<T>::fma_bcast(Q, P.wt[j * KX + i], AB[i]);
      } else {
#pragma unroll
        for (int iv = 0; iv < KX / 2; ++iv) {
          unsigned long long w0, w1;
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"((unsigned)__cvta_generic_to_shared(ws + 2 * (j * KX + iv * 2))));
          asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AB[iv * 2].v) : "l"(Q.v), "l"(w0));
          asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AB[iv * 2 + 1].v) : "l"(Q.v), "l"(w1));
        }
      }
    }
    Pair<T> G0 = Pair<T>::make(T(0), AB[0].hi()), G1 = Pair<T>::make(T(0), T(0));
    T pxm = T(1), px = sx, dpxm = T(0), dpx = T(1);
#pragma unroll
    for (int i = 1; i < KX; ++i) {
      if (i >= 2) {
        const T pn = leg_a<T>(i - 1) * sx * px - leg_b<T>(i - 1) * pxm;
        const T dpn = Math<T>::fma(leg_c<T>(i - 1), px, dpxm);
        pxm = px; px = pn; dpxm = dpx; dpx = dpn;
      }
      if (i & 1) G1 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G1);
      else G0 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G0);
    }
    gx += inx ? (G0.lo() + G1.lo()) * P.ihx : T(0);
    gy += iny ? (G0.hi() + G1.hi()) * P.ihy : T(0);
  }
};

// THD with an explicit software pipeline: loads and FFMA2s of the shared-memory columns are all `asm volatile`
// (program order kept), the loads run DEPTH slots ahead of their use.
template <int KX, int KY, int J0, int DEPTH> struct BiasLegendre2DStaticTHP {
  using T = float;
  using Params = Legendre2DStaticTHParams<KX, KY>;
  static size_t smem_bytes(const Params&) { return sizeof(T) * KX * KY * 2; }
  static PV_DEV void stage(const Params& P, void* smem) { BiasLegendre2DStaticTHD<KX, KY, J0>::stage(P, smem); }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    bool inx, iny;
    const T sx = scale_clamp(x, P.cx, P.ihx, inx);
    const T sy = scale_clamp(y, P.cy, P.ihy, iny);
    const unsigned ws = (unsigned)__cvta_generic_to_shared(smem);
    Pair<T> AB[KX];
    T pm = T(1), p = sy, dpm = T(0), dp = T(1);
    constexpr int HX = KX / 2, NL = (KY - J0) * HX;
    unsigned long long buf[DEPTH][2];
#pragma unroll
    for (int l = 0; l < DEPTH; ++l)
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(buf[l][0]), "=l"(buf[l][1]) : "r"(ws + 8u * (unsigned)((J0 + l / HX) * KX + (l % HX) * 2)));
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
      if (j < J0) {
#pragma unroll
        for (int i = 0; i < KX; ++i) AB[i] = Pair<T>::fma_bcast(Q, P.wt[j * KX + i], AB[i]);
      } else {
#pragma unroll
        for (int iv = 0; iv < HX; ++iv) {
          const int l = (j - J0) * HX + iv;
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AB[iv * 2].v) : "l"(Q.v), "l"(buf[l % DEPTH][0]));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(AB[iv * 2 + 1].v) : "l"(Q.v), "l"(buf[l % DEPTH][1]));
          if (l + DEPTH < NL) {
            const int ln = l + DEPTH;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(buf[l % DEPTH][0]), "=l"(buf[l % DEPTH][1]) : "r"(ws + 8u * (unsigned)((J0 + ln / HX) * KX + (ln % HX) * 2)));
          }
        }
      }
    }
    Pair<T> G0 = Pair<T>::make(T(0), AB[0].hi()), G1 = Pair<T>::make(T(0), T(0));
    T pxm = T(1), px = sx, dpxm = T(0), dpx = T(1);
#pragma unroll
    for (int i = 1; i < KX; ++i) {
      if (i >= 2) {
        const T pn = leg_a<T>(i - 1) * sx * px - leg_b<T>(i - 1) * pxm;
        const T dpn = Math<T>::fma(leg_c<T>(i - 1), px, dpxm);
        pxm = px; px = pn; dpxm = dpx; dpx = dpn;
      }
      if (i & 1) G1 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G1);
      else G0 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[i], G0);
    }
    gx += inx ? (G0.lo() + G1.lo()) * P.ihx : T(0);
    gy += iny ? (G0.hi() + G1.hi()) * P.ihy : T(0);
  }
};

}  // namespace pv
