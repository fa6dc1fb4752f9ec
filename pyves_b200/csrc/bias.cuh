// Bias energy functions evaluated per walker: energy and gradient w.r.t. (x, y).
// Replaces the TorchScript module + autograd that openmmtorch.TorchForce runs every step
// (ves/bias.py:52-58) for the modules of ves/basis.py.  CPU twins: oracle/legendre.py,
// oracle/mlp.py, oracle/pathcv.py.
//
// Functor interface:
//   Params                       POD passed BY VALUE as a kernel parameter (constant bank)
//   smem_bytes(P)                dynamic shared memory the functor needs
//   stage(P, smem)               cooperative fill of shared memory (all threads; caller syncs)
//   grad<WANT_E>(P, smem, x, y, e, gx, gy)   ACCUMULATES dV/dx, dV/dy (and V when WANT_E)
#pragma once
#include <type_traits>

#include "common.cuh"

namespace pv {

// s = (q - centre) * inv_half clamped to [-1, 1]; `inside` is where torch.clamp passes the
// gradient (inclusive at both ends)  -- ves/basis.py:143,147.
template <typename T> PV_DEV T scale_clamp(T q, T centre, T inv_half, bool& inside) {
  const T s = (q - centre) * inv_half;
  inside = (s >= T(-1)) && (s <= T(1));
  return Math<T>::min(Math<T>::max(s, T(-1)), T(1));
}

// ------------------------------------------------------------------------------------------------
template <typename T> struct BiasNone {
  struct Params {};
  static size_t smem_bytes(const Params&) { return 0; }
  static PV_DEV void stage(const Params&, void*) {}
  template <bool WANT_E> static PV_DEV void grad(const Params&, const void*, T, T, T&, T&, T&) {}
};

// ------------------------------------------------------------------------------------------------
// Device-resident coefficient updates (pyves_set_bias_legendre{1d,2d}_device): kernel parameters come from the
// host, so the static functors can instead read their weights from a __constant__ array that is refreshed with a
// stream-ordered device-to-device cudaMemcpyToSymbolAsync -- same constant-bank operands (bank 3 instead of
// the parameter bank 0), no host round trip between two launches.  One copy per translation unit; the kernels
// and launch_upload_static_weights of langevin_inst.cu (group 0) share theirs.
constexpr int kConstWeights = 256;
static __constant__ float c_wt_f32[kConstWeights];
static __constant__ double c_wt_f64[kConstWeights];
template <typename T> PV_DEV const T* const_weights();
template <> PV_DEV const float* const_weights<float>() { return c_wt_f32; }
template <> PV_DEV const double* const_weights<double>() { return c_wt_f64; }

// ------------------------------------------------------------------------------------------------
// LegendreBasis1D (ves/basis.py:86-156).  DEG > 0: compile-time degree, recursion fully
// unrolled, weights are constant-bank operands.  DEG == 0: runtime degree.
template <typename T> struct Legendre1DParams {
  int axis;
  int degree;
  T centre;
  T inv_half;
  T w[kMaxK1];
};

template <typename T, int DEG, bool CONSTSYM = false> struct BiasLegendre1D {
  using Params = Legendre1DParams<T>;
  static PV_DEV T wgt(const Params& P, int n) { return CONSTSYM ? const_weights<T>()[n] : P.w[n]; }
  static size_t smem_bytes(const Params&) { return 0; }
  static PV_DEV void stage(const Params&, void*) {}
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void*, T x, T y, T& e, T& gx, T& gy) {
    bool inside;
    const T s = scale_clamp(P.axis ? y : x, P.centre, P.inv_half, inside);
    const int deg = DEG > 0 ? DEG : P.degree;
    T pm = T(1), p = s, dpm = T(0), dp = T(1);
    T E = wgt(P, 0), g = T(0);
    if (deg >= 1) {
      E = Math<T>::fma(wgt(P, 1), s, E);
      g = wgt(P, 1);
    }
    if (DEG > 0) {
#pragma unroll
      for (int n = 1; n < DEG; ++n) {
        const T pn = leg_a<T>(n) * s * p - leg_b<T>(n) * pm;
        const T dpn = Math<T>::fma(leg_c<T>(n), p, dpm);
        g = Math<T>::fma(wgt(P, n + 1), dpn, g);
        if (WANT_E) E = Math<T>::fma(wgt(P, n + 1), pn, E);
        pm = p; p = pn; dpm = dp; dp = dpn;
      }
    } else {
      for (int n = 1; n < deg; ++n) {
        const T fn = T(n), inv = T(1) / (fn + T(1)), c = T(2) * fn + T(1);
        const T pn = (c * s * p - fn * pm) * inv;
        const T dpn = Math<T>::fma(c, p, dpm);
        g = Math<T>::fma(wgt(P, n + 1), dpn, g);
        if (WANT_E) E = Math<T>::fma(wgt(P, n + 1), pn, E);
        pm = p; p = pn; dpm = dp; dp = dpn;
      }
    }
    g = inside ? g * P.inv_half : T(0);
    gx += P.axis ? T(0) : g;
    gy += P.axis ? g : T(0);
    if (WANT_E) e += E;
  }
};

// ------------------------------------------------------------------------------------------------
// Two walkers per thread in packed form (lo = walker a, hi = walker b), for the light biases whose step is
// issue bound: every FP32 operation of the pair is ONE packed FP32x2 instruction.  The default evaluates the
// scalar functor half by half.
template <class Bias, typename T> struct BiasPair {
  static PV_DEV void grad2(const typename Bias::Params& P, const void* smem, Pair<T> x, Pair<T> y, Pair<T>& gx, Pair<T>& gy) {
    T e = T(0), ax = gx.lo(), ay = gy.lo(), bx = gx.hi(), by = gy.hi();
    Bias::template grad<false>(P, smem, x.lo(), y.lo(), e, ax, ay);
    Bias::template grad<false>(P, smem, x.hi(), y.hi(), e, bx, by);
    gx = Pair<T>::make(ax, bx);
    gy = Pair<T>::make(ay, by);
  }
};

template <typename T> struct BiasPair<BiasNone<T>, T> {
  static PV_DEV void grad2(const typename BiasNone<T>::Params&, const void*, Pair<T>, Pair<T>, Pair<T>&, Pair<T>&) {}
};

template <typename T, int DEG, bool CONSTSYM> struct BiasPair<BiasLegendre1D<T, DEG, CONSTSYM>, T> {
  using Q = Pair<T>;
  using B1 = BiasLegendre1D<T, DEG, CONSTSYM>;
  static PV_DEV void grad2(const Legendre1DParams<T>& P, const void*, Q x, Q y, Q& gx, Q& gy) {
    const Q q = P.axis ? y : x;
    const Q sraw = Q::mul_bcast(Q::add_bcast(q, -P.centre), P.inv_half);
    const T s0 = sraw.lo(), s1 = sraw.hi();
    const bool in0 = (s0 >= T(-1)) && (s0 <= T(1)), in1 = (s1 >= T(-1)) && (s1 <= T(1));
    const Q s = Q::make(Math<T>::min(Math<T>::max(s0, T(-1)), T(1)), Math<T>::min(Math<T>::max(s1, T(-1)), T(1)));
    const int deg = DEG > 0 ? DEG : P.degree;
    Q pm = Q::make(T(1), T(1)), p = s, dpm = Q::make(T(0), T(0)), dp = Q::make(T(1), T(1));
    Q g = deg >= 1 ? Q::make(B1::wgt(P, 1), B1::wgt(P, 1)) : Q::make(T(0), T(0));
    if (DEG > 0) {
#pragma unroll
      for (int n = 1; n < DEG; ++n) {
        const Q pn = Q::fma(Q::mul_bcast(s, leg_a<T>(n)), p, Q::mul_bcast(pm, -leg_b<T>(n)));
        const Q dpn = Q::fma_bcast(p, leg_c<T>(n), dpm);
        g = Q::fma_bcast(dpn, B1::wgt(P, n + 1), g);
        pm = p; p = pn; dpm = dp; dp = dpn;
      }
    } else {
      for (int n = 1; n < deg; ++n) {
        const T fn = T(n), inv = T(1) / (fn + T(1)), c = T(2) * fn + T(1);
        const Q pn = Q::fma(Q::mul_bcast(s, c * inv), p, Q::mul_bcast(pm, -fn * inv));
        const Q dpn = Q::fma_bcast(p, c, dpm);
        g = Q::fma_bcast(dpn, B1::wgt(P, n + 1), g);
        pm = p; p = pn; dpm = dp; dp = dpn;
      }
    }
    g = Q::mul_bcast(g, P.inv_half);
    const Q gm = Q::make(in0 ? g.lo() : T(0), in1 ? g.hi() : T(0));
    if (P.axis) gy = Q::add(gy, gm);
    else gx = Q::add(gx, gm);
  }
};

// ------------------------------------------------------------------------------------------------
// 2D tensor-product Legendre basis, V = sum_ij w_ij P_i(x') P_j(y')  (per-axis
// conventions of ves/basis.py:143-147).  Fills Py[0..KY), dPy[0..KY) by recursion.
template <typename T, int KY> PV_DEV void legendre_fill(T s, T (&P)[KY], T (&dP)[KY]) {
  P[0] = T(1);
  dP[0] = T(0);
  if (KY > 1) {
    P[1] = s;
    dP[1] = T(1);
  }
#pragma unroll
  for (int n = 1; n + 1 < KY; ++n) {
    P[n + 1] = leg_a<T>(n) * s * P[n] - leg_b<T>(n) * P[n - 1];
    dP[n + 1] = Math<T>::fma(leg_c<T>(n), P[n], dP[n - 1]);
  }
}

// Static sizes, weights by value in the constant bank, everything unrolled: the 2 KX KY FMAs
// of A_i = sum_j w_ij P_j(y'), B_i = sum_j w_ij P'_j(y') take their weight as a c[bank][imm]
// operand, so no load instruction is issued for them.
template <typename T, int KX, int KY> struct Legendre2DStaticParams {
  T cx, ihx, cy, ihy;
  T w[KX * KY];
};

// Packed evaluation: Q[j] = (P_j(y'), P'_j(y')) pairs, so that one packed FP32x2 FMA (FFMA2, with
// the weight broadcast to both halves) updates (A_i, B_i) together.  ROWS rows are accumulated at a
// time (2 * ROWS independent chains sharing the Q[j] operand).  Measured on B200 (profiles/): the
// 16 x 16 kernel issues 533 instructions per walker-step instead of 748 with scalar FFMAs.
template <typename T, int KX, int KY, int ROWS = 1, bool SPLIT = true> struct BiasLegendre2DStatic {
  using Params = Legendre2DStaticParams<T, KX, KY>;
  static_assert(KX % ROWS == 0, "ROWS must divide KX");
  static size_t smem_bytes(const Params&) { return 0; }
  static PV_DEV void stage(const Params&, void*) {}
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void*, T x, T y, T& e, T& gx, T& gy) {
    bool inx, iny;
    const T sx = scale_clamp(x, P.cx, P.ihx, inx);
    const T sy = scale_clamp(y, P.cy, P.ihy, iny);
    T Py[KY], dPy[KY];
    legendre_fill<T, KY>(sy, Py, dPy);
    Pair<T> Q[KY];
#pragma unroll
    for (int j = 0; j < KY; ++j) Q[j] = Pair<T>::make(Py[j], dPy[j]);
    const Pair<T> zero = Pair<T>::make(T(0), T(0)), one = Pair<T>::make(T(1), T(1));
    Pair<T> G = zero;   // (dV/dsx, dV/dsy)
    T E = T(0);
    T pxm = T(1), px = T(1), dpxm = T(0), dpx = T(0);
#pragma unroll
    for (int i0 = 0; i0 < KX; i0 += ROWS) {
      Pair<T> AB0[ROWS], AB1[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) AB0[r] = AB1[r] = zero;
#pragma unroll
      for (int j = 0; j < KY; ++j)
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          if (SPLIT && (j & 1)) AB1[r] = Pair<T>::fma_bcast(Q[j], P.w[(i0 + r) * KY + j], AB1[r]);
          else AB0[r] = Pair<T>::fma_bcast(Q[j], P.w[(i0 + r) * KY + j], AB0[r]);
        }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const int i = i0 + r;
        if (i == 1) {
          px = sx;
          dpx = T(1);
        } else if (i >= 2) {
          const T pn = leg_a<T>(i - 1) * sx * px - leg_b<T>(i - 1) * pxm;
          const T dpn = Math<T>::fma(leg_c<T>(i - 1), px, dpxm);
          pxm = px; px = pn; dpxm = dpx; dpx = dpn;
        }
        const Pair<T> AB = SPLIT ? Pair<T>::fma(AB0[r], one, AB1[r]) : AB0[r];
        G = Pair<T>::fma(Pair<T>::make(dpx, px), AB, G);
        if (WANT_E) E = Math<T>::fma(px, AB.lo(), E);
      }
    }
    gx += inx ? G.lo() * P.ihx : T(0);
    gy += iny ? G.hi() * P.ihy : T(0);
    if (WANT_E) e += E;
  }
};

// Column sweep ("j-outer") evaluation of the same sum, the default for the static sizes: the y
// recursion yields Q_j = (P_j(y'), P'_j(y')) and every Q_j updates all KX row accumulators
// (A_i, B_i) += w_ij Q_j with one packed FP32x2 FMA each (FFMA2, weight broadcast to both halves).
//  * KX independent accumulation chains: no dependent-issue waits inside the contraction;
//  * the weights of one column are contiguous (wt[j][i]) in the constant bank;
//  * no FADD2 to merge split chains, no spills at 255 registers.
// Measured on B200, 16 x 16, 10^6 walkers x 500 steps: 11.77 ms against 12.6-13.0 ms for the
// row-major functor above (profiles/r01_variants.md).
template <typename T, int KX, int KY> struct Legendre2DStaticTParams {
  T cx, ihx, cy, ihy;
  T wt[KY * KX];   // wt[j * KX + i] = w_ij
};

template <typename T, int KX, int KY> struct Legendre2DStaticTCParams {
  T cx, ihx, cy, ihy;
};

// J0 < KY makes the sweep HYBRID: columns j < J0 take their weights from the constant bank, columns j >= J0 from a
// shared-memory copy read with volatile LDS.128 broadcasts right before use.  ptxas feeds only ~64 of the 256
// weights of the 16 x 16 kernel through cheap uniform loads and fetches the rest with LDC.64 into vector registers
// (ADU path, one per ~8.7 cycles per scheduler); moving those columns to the LSU path is worth 5 %:
// 16 x 16, 10^6 walkers x 500 steps: 11.83 ms all constant bank, 11.21 ms with J0 = 4 (profiles/r01_variants.md).
template <typename T, int KX, int KY, bool CONSTSYM = false, int J0 = KY> struct BiasLegendre2DStaticT {
  using Params = typename std::conditional<CONSTSYM, Legendre2DStaticTCParams<T, KX, KY>, Legendre2DStaticTParams<T, KX, KY>>::type;
  static_assert(!CONSTSYM || KX * KY <= kConstWeights, "constant weight array too small");
  static_assert(J0 >= 2 && J0 <= KY, "columns 0 and 1 always come from the constant bank");
  static constexpr int V = 16 / sizeof(T);
  template <bool C = CONSTSYM> static PV_DEV typename std::enable_if<C, T>::type weight(const Params&, int idx) { return const_weights<T>()[idx]; }
  template <bool C = CONSTSYM> static PV_DEV typename std::enable_if<!C, T>::type weight(const Params& P, int idx) { return P.wt[idx]; }
  static_assert(KX >= 2 && KY >= 2, "static sizes start at 2 x 2");
  static size_t smem_bytes(const Params&) { return J0 < KY ? sizeof(T) * KX * KY : 0; }
  static PV_DEV void stage(const Params& P, void* smem) {
    if (J0 < KY) {
      T* d = reinterpret_cast<T*>(smem);
      for (int i = threadIdx.x; i < KX * KY; i += blockDim.x) d[i] = weight(P, i);
    }
  }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    const T* ws = reinterpret_cast<const T*>(smem);
    bool inx, iny;
    const T sx = scale_clamp(x, P.cx, P.ihx, inx);
    const T sy = scale_clamp(y, P.cy, P.ihy, iny);
    Pair<T> AB[KX];
    // columns 0 and 1: Q_0 = (1, 0), Q_1 = (sy, 1)
    T pm = T(1), p = sy, dpm = T(0), dp = T(1);
    {
      const Pair<T> Q1 = Pair<T>::make(sy, T(1));
#pragma unroll
      for (int i = 0; i < KX; ++i) AB[i] = Pair<T>::fma_bcast(Q1, weight(P, 1 * KX + i), Pair<T>::make(weight(P, i), T(0)));
    }
#pragma unroll
    for (int j = 2; j < KY; ++j) {
      const T pn = leg_a<T>(j - 1) * sy * p - leg_b<T>(j - 1) * pm;
      const T dpn = Math<T>::fma(leg_c<T>(j - 1), p, dpm);
      pm = p; p = pn; dpm = dp; dp = dpn;
      const Pair<T> Q = Pair<T>::make(p, dp);
      if (j < J0) {
#pragma unroll
        for (int i = 0; i < KX; ++i) AB[i] = Pair<T>::fma_bcast(Q, weight(P, j * KX + i), AB[i]);
      } else {
#pragma unroll
        for (int iv = 0; iv < KX / V; ++iv) {
          T w[V];
          lds_vec(ws + j * KX + iv * V, w);
#pragma unroll
          for (int q = 0; q < V; ++q) Pair<T>::acc_bcast(AB[iv * V + q], Q, w[q]);
        }
      }
    }
    // row sweep: (dV/dsx, dV/dsy) = sum_i (P'_i(x'), P_i(x')) * (A_i, B_i), two chains
    Pair<T> G0 = Pair<T>::make(T(0), AB[0].hi()), G1 = Pair<T>::make(T(0), T(0));
    T E = WANT_E ? AB[0].lo() : T(0);
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
      if (WANT_E) E = Math<T>::fma(px, AB[i].lo(), E);
    }
    gx += inx ? (G0.lo() + G1.lo()) * P.ihx : T(0);
    gy += iny ? (G0.hi() + G1.hi()) * P.ihy : T(0);
    if (WANT_E) e += E;
  }
};

// Runtime sizes: column sweep with the weights in shared memory.  Rows are handled in NRB blocks of RB
// (kx <= RB * NRB); per row block the weights are transposed and zero-padded to wt[rb][kyp][RB]
// (kyp = 2 + roundup(ky - 2, UNR)) and read with 128-bit broadcast loads.  A rolled loop over the
// columns, UNR per trip, keeps RB accumulator pairs and one column of weights live (no spills, two to
// three blocks per SM); each row block repeats the cheap y recursion.  Measured on B200 (FP32, 10^6
// walkers): 32 x 32 1.25e10 walker-steps/s (0.80 of the FMA peak; the row-major version reached
// 0.53), 64 x 64 0.75 (0.57), runtime 16 x 16 0.73 (0.48).
template <typename T> struct Legendre2DSmemParams {
  T cx, ihx, cy, ihy;
  int kx, ky;
  const T* w;   // device, row-major [kx][ky]
};

template <typename T, int RB, int NRB, int UNR> struct BiasLegendre2DSmem {
  using Params = Legendre2DSmemParams<T>;
  static constexpr int V = 16 / sizeof(T);   // elements per 128-bit load
  static __host__ __device__ int kyp(int ky) { return 2 + (ky > 2 ? (ky - 2 + UNR - 1) / UNR * UNR : 0); }
  static __host__ __device__ int nrb(int kx) { return (kx + RB - 1) / RB; }
  static size_t smem_bytes(const Params& P) { return (size_t)kyp(P.ky) * (nrb(P.kx) * RB + 4) * sizeof(T); }
  static PV_DEV void stage(const Params& P, void* smem) {
    T* wt = reinterpret_cast<T*>(smem);
    const int n = kyp(P.ky), nb = nrb(P.kx);
    for (int idx = threadIdx.x; idx < nb * n * RB; idx += blockDim.x) {
      const int r = idx % RB, j = (idx / RB) % n, i = (idx / (RB * n)) * RB + r;
      wt[idx] = (i < P.kx && j < P.ky) ? P.w[i * P.ky + j] : T(0);
    }
    T* co = wt + nb * n * RB;   // co[4 j + {0, 1, 2}] = A_{j-1}, B_{j-1}, C_{j-1}: P_j from P_{j-1}, P_{j-2}
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
    const int n = kyp(P.ky), nb = nrb(P.kx);
    const T* co = reinterpret_cast<const T*>(smem) + nb * n * RB;
    Pair<T> G0 = Pair<T>::make(T(0), T(0)), G1 = G0;   // (dV/dsx, dV/dsy), two chains
    T E = T(0);
    T pxm = T(1), px = T(1), dpxm = T(0), dpx = T(0);
#pragma unroll
    for (int rb = 0; rb < NRB; ++rb) {
      if (rb < nb) {
        const T* wt = reinterpret_cast<const T*>(smem) + rb * n * RB;
        Pair<T> AB[RB];
        // columns 0 and 1: Q_0 = (1, 0), Q_1 = (sy, 1)
        T pm = T(1), p = sy, dpm = T(0), dp = T(1);
        {
          const Pair<T> Q1 = Pair<T>::make(sy, T(1));
#pragma unroll
          for (int iv = 0; iv < RB / V; ++iv) {
            T w0[V], w1[V];
            lds_vec(wt + iv * V, w0);
            lds_vec(wt + RB + iv * V, w1);
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
            const T cc = sizeof(T) == 4 ? c[2] : T(2 * j - 1);   // FP64: the 128-bit load holds A, B only
            const T pn = c[0] * sy * p - c[1] * pm;
            const T dpn = Math<T>::fma(cc, p, dpm);
            pm = p; p = pn; dpm = dp; dp = dpn;
            const Pair<T> Q = Pair<T>::make(p, dp);
#pragma unroll
            for (int iv = 0; iv < RB / V; ++iv) {
              T w[V];
              lds_vec(wt + j * RB + iv * V, w);
#pragma unroll
              for (int q = 0; q < V; ++q) Pair<T>::acc_bcast(AB[iv * V + q], Q, w[q]);
            }
          }
        }
        // row sweep over this block, continuing the x recursion
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const int i = rb * RB + r;
          if (i == 1) {
            px = sx;
            dpx = T(1);
          } else if (i >= 2) {
            const T pn = leg_a<T>(i - 1) * sx * px - leg_b<T>(i - 1) * pxm;
            const T dpn = Math<T>::fma(leg_c<T>(i - 1), px, dpxm);
            pxm = px; px = pn; dpxm = dpx; dpx = dpn;
          }
          if (i & 1) G1 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[r], G1);
          else G0 = Pair<T>::fma(Pair<T>::make(dpx, px), AB[r], G0);
          if (WANT_E) E = Math<T>::fma(px, AB[r].lo(), E);
        }
      }
    }
    gx += inx ? (G0.lo() + G1.lo()) * P.ihx : T(0);
    gy += iny ? (G0.hi() + G1.hi()) * P.ihy : T(0);
    if (WANT_E) e += E;
  }
};

// ------------------------------------------------------------------------------------------------
// NNBasis1D (ves/basis.py:184-232) in folded form.  The first two Linear layers have no
// activation between them (:197-202), so the host folds them: z1 = u s + c.  The derivative
// dV/ds is carried in forward mode next to the value, so each weight read from shared memory
// feeds two FMAs.
//   H == 0 : two hidden layers (no matmul left after folding), any width (padded to 4):
//            streamed, no per-thread arrays.       shared (T): u[W] c[W] wout[W] bout pad[3]
//   H  > 0 : three hidden layers = one H x H matmul, widths padded to H.
//                                                  shared (T): u[H] c[H] wout[H] b2[H] W2T[H][H] bout pad[3] (W2T[i][j] = W2[j][i])
// A single hidden layer is the reference's degenerate linear case V = u0 s + bout (scalars by value).
// The output bias travels at the end of the packed buffer, not by value, so that a device-resident optimiser
// (pyves_set_bias_mlp_device) can refresh every parameter without a host copy.
template <typename T> struct MlpParams {
  int axis;
  int linear;      // 1: single hidden layer
  int width;       // padded width of the streamed (H == 0) variant
  T lo, inv_range;
  T bout, u0;       // single hidden layer only
  const T* packed;  // device, layout above
};

template <typename T, int ACT> PV_DEV T act_fn(T z, T& d) {
  if (ACT == PYVES_ACT_RELU) {
    d = z > T(0) ? T(1) : T(0);
    return Math<T>::max(z, T(0));
  } else {
    const T t = Math<T>::tanh(z);
    d = T(1) - t * t;
    return t;
  }
}

template <typename T, int H, int ACT> struct BiasMlp {
  using Params = MlpParams<T>;
  static constexpr int V = 16 / sizeof(T);
  struct alignas(16) Vec { T v[V]; };
  static size_t smem_bytes(const Params& P) {
    if (P.linear) return 0;
    return (size_t)packed_count(P) * sizeof(T);
  }
  static __host__ __device__ int packed_count(const Params& P) { return (H > 0 ? 4 * H + H * H : 3 * P.width) + 4; }
  static PV_DEV void stage(const Params& P, void* smem) {
    if (P.linear) return;
    T* d = reinterpret_cast<T*>(smem);
    const int n = packed_count(P);
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = P.packed[i];
  }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    const T s = ((P.axis ? y : x) - P.lo) * P.inv_range;
    const T* sm = reinterpret_cast<const T*>(smem);
    T val = P.linear ? P.bout : sm[packed_count(P) - 4], dval = T(0);
    if (P.linear) {
      val = Math<T>::fma(P.u0, s, P.bout);
      dval = P.u0;
    } else if (H == 0) {
      const int nv = P.width / V;
      const Vec* uv = reinterpret_cast<const Vec*>(sm);
      const Vec* cv = uv + nv;
      const Vec* wv = cv + nv;
      for (int iv = 0; iv < nv; ++iv) {
        const Vec u = uv[iv], c = cv[iv], w = wv[iv];
#pragma unroll
        for (int q = 0; q < V; ++q) {
          T d;
          const T a = act_fn<T, ACT>(Math<T>::fma(u.v[q], s, c.v[q]), d);
          val = Math<T>::fma(w.v[q], a, val);
          dval = Math<T>::fma(w.v[q] * d, u.v[q], dval);
        }
      }
    } else {
      // input sweep: every hidden unit i of the folded first layer is evaluated once, (h_i, dh_i/ds),
      // and immediately feeds all HH outputs (z_j, dz_j/ds) += W2[j][i] * (h_i, dh_i/ds): HH independent
      // FFMA2 chains, the transposed weights W2T[i][.] come as 128-bit shared-memory broadcasts.
      constexpr int HH = H > 0 ? H : 4;
      const T* b2 = sm + 3 * HH;
      const T* WT = sm + 4 * HH;
      Pair<T> Z[HH];
#pragma unroll
      for (int jv = 0; jv < HH / V; ++jv) {
        T b[V];
        lds_vec(b2 + jv * V, b);
#pragma unroll
        for (int q = 0; q < V; ++q) Z[jv * V + q] = Pair<T>::make(b[q], T(0));
      }
#pragma unroll 1
      for (int i0 = 0; i0 < HH; i0 += V) {
        T u[V], c[V];
        lds_vec(sm + i0, u);
        lds_vec(sm + HH + i0, c);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          T d;
          const T a = act_fn<T, ACT>(Math<T>::fma(u[k], s, c[k]), d);
          const Pair<T> HD = Pair<T>::make(a, d * u[k]);
#pragma unroll
          for (int jv = 0; jv < HH / V; ++jv) {
            T w[V];
            lds_vec(WT + (i0 + k) * HH + jv * V, w);
#pragma unroll
            for (int q = 0; q < V; ++q) Pair<T>::acc_bcast(Z[jv * V + q], HD, w[q]);
          }
        }
      }
#pragma unroll
      for (int jv = 0; jv < HH / V; ++jv) {
        T wo[V];
        lds_vec(sm + 2 * HH + jv * V, wo);
#pragma unroll
        for (int q = 0; q < V; ++q) {
          T d;
          const T a = act_fn<T, ACT>(Z[jv * V + q].lo(), d);
          val = Math<T>::fma(wo[q], a, val);
          dval = Math<T>::fma(wo[q] * d, Z[jv * V + q].hi(), dval);
        }
      }
    }
    const T g = dval * P.inv_range;
    gx += P.axis ? T(0) : g;
    gy += P.axis ? g : T(0);
    if (WANT_E) e += val;
  }
};

// ------------------------------------------------------------------------------------------------
// ReLU NNBasis1D as what it is: a piecewise-LINEAR function of its one input.  Every unit of the folded first layer switches at
// s = -c_i / u_i; between two such points the second layer's pre-activations are linear in s and switch at most once each; so
// dV/ds is piecewise constant on at most (H + 1)^2 pieces.  mlp_pwl_build_kernel (misc_kernels.cu) compiles the parameters into
// the sorted breakpoints and the slope of every piece (FP64 sums, exact up to the rounding of the breakpoints) whenever they
// change; the step kernel's bias force is then ONE binary search in shared memory -- no matmul, no activations, no tensor core.
// Step kernels only (forces); energies and parameter gradients go through the network itself (BiasMlp, moments_mlp2_kernel).
constexpr int kPwlBuckets = 2048;   // uniform grid over [first, last breakpoint]: where the binary search starts

template <typename T> struct PwlParams {
  int axis;
  T lo, inv_range;
  int cap;             // capacity of brk (slope has cap + 1 entries)
  const T* brk;        // device: breakpoints in the scaled coordinate, ascending
  const T* slope;      // device: dV/ds on the pieces: slope[k] for brk[k-1] < s <= brk[k]
  const int* count;    // device: number of breakpoints
  const int* grid;     // device: grid[b] = number of breakpoints below the lower edge of bucket b (kPwlBuckets + 1 entries)
  const T* grid_map;   // device: {origin, buckets per unit of s}
};

template <typename T> struct BiasPwl1D {
  using Params = PwlParams<T>;
  static size_t smem_bytes(const Params& P) { return (size_t)(2 * P.cap + 1) * sizeof(T) + (size_t)(kPwlBuckets + 1) * sizeof(int) + 32; }
  static PV_DEV void stage(const Params& P, void* smem) {
    T* d = reinterpret_cast<T*>(smem);
    int* g = reinterpret_cast<int*>(d + 2 * P.cap + 1);
    const int n = *P.count;
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = P.brk[i];
    for (int i = threadIdx.x; i <= n; i += blockDim.x) d[P.cap + i] = P.slope[i];
    for (int i = threadIdx.x; i <= kPwlBuckets; i += blockDim.x) g[i] = P.grid[i];
  }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    const T s = ((P.axis ? y : x) - P.lo) * P.inv_range;
    const T* brk = reinterpret_cast<const T*>(smem);
    const int* g = reinterpret_cast<const int*>(brk + 2 * P.cap + 1);
    // number of breakpoints below s: the bucket of s (one bucket of margin on either side against rounding at the edges)
    // bounds it from both sides, a branch-free halving finishes inside -- usually zero to two probes instead of twelve
    const T t = (s - P.grid_map[0]) * P.grid_map[1];
    int b = (int)Math<T>::min(Math<T>::max(t, T(0)), T(kPwlBuckets - 1));
    int lo = g[b > 0 ? b - 1 : 0];
    int len = g[b + 2 <= kPwlBuckets ? b + 2 : kPwlBuckets] - lo;
    while (len > 0) {
      const int half = len >> 1;
      const bool right = brk[lo + half] < s;
      lo = right ? lo + half + 1 : lo;
      len = right ? len - half - 1 : half;
    }
    const T gr = brk[P.cap + lo] * P.inv_range;
    gx += P.axis ? T(0) : gr;
    gy += P.axis ? gr : T(0);
  }
};

// ------------------------------------------------------------------------------------------------
// Everything else behind one runtime switch: LegendreBasis2DPathCV (ves/basis.py:509-548),
// HarmonicBias (ves/bias.py:133), and MLPs of any depth/width up to kGenH (unfolded, layer by
// layer, activations in local memory).
constexpr int kGenH = 128;

template <typename T> struct OtherParams {
  int kind;                  // PYVES_BIAS_PATHCV | PYVES_BIAS_HARMONIC | PYVES_BIAS_MLP_1D
  // path CV
  int degree, npath;
  T lam, phi;
  T w[kMaxK1];
  const T* path;             // device: x_i[npath], y_i[npath]
  // harmonic / mlp
  int axis, act;
  T k, s0;                   // harmonic; mlp: lo = s0, inv_range = k
  int n_layers;              // mlp: number of Linear layers
  int sizes[10];             // mlp: widths incl. input 1 and output 1
  const T* theta;            // mlp: device, torch parameter order
};

template <typename T> struct BiasOther {
  using Params = OtherParams<T>;
  static size_t smem_bytes(const Params& P) { return P.kind == PYVES_BIAS_PATHCV ? 2 * (size_t)P.npath * sizeof(T) : 0; }
  static PV_DEV void stage(const Params& P, void* smem) {
    if (P.kind == PYVES_BIAS_PATHCV) {
      T* d = reinterpret_cast<T*>(smem);
      for (int i = threadIdx.x; i < 2 * P.npath; i += blockDim.x) d[i] = P.path[i];
    }
  }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    if (P.kind == PYVES_BIAS_HARMONIC) {
      const T d = (P.axis ? y : x) - P.s0;
      gx += P.axis ? T(0) : P.k * d;
      gy += P.axis ? P.k * d : T(0);
      if (WANT_E) e += T(0.5) * P.k * d * d;
    } else if (P.kind == PYVES_BIAS_PATHCV) {
      const T* xi = reinterpret_cast<const T*>(smem);
      const T* yi = xi + P.npath;
      // nearest image m first; the exponents are then -lam (d_i^2 - d_m^2) with the difference of the squared distances
      // taken as (x_m - x_i)(2x - x_i - x_m) + (y_m - y_i)(2y - y_i - y_m): no cancellation between two rounded squares
      // (lam = 100 turned the FP32 rounding of d^2 into 1e-3 relative errors of the weights)
      T dmin = T(1e30), xm = xi[0], ym = yi[0];
      for (int i = 0; i < P.npath; ++i) {
        const T dx = x - xi[i], dy = y - yi[i];
        const T d2 = Math<T>::fma(dx, dx, dy * dy);
        if (d2 < dmin) { dmin = d2; xm = xi[i]; ym = yi[i]; }
      }
      const T amax = -P.lam * dmin;
      const T tx = (x - xm) + x, ty = (y - ym) + y;       // 2x - x_m, 2y - y_m
      T S0 = 0, S1 = 0, G0x = 0, G0y = 0, G1x = 0, G1y = 0;
      for (int i = 0; i < P.npath; ++i) {
        const T dx = x - xi[i], dy = y - yi[i];
        const T dd = Math<T>::fma(xm - xi[i], tx - xi[i], (ym - yi[i]) * (ty - yi[i]));
        const T ee = Math<T>::exp(-P.lam * dd);
        const T ei = ee * T(i + 1);
        S0 += ee; S1 += ei;
        G0x += ee * dx; G0y += ee * dy;
        G1x += ei * dx; G1y += ei * dy;
      }
      const T m2l = T(-2) * P.lam;
      const T dL0x = m2l * G0x / S0, dL0y = m2l * G0y / S0;
      const T dL1x = m2l * G1x / S1, dL1y = m2l * G1y / S1;
      const T s = S1 / (S0 * T(P.npath));
      bool inside;
      const T sp = scale_clamp(s, T(0.5), T(2), inside);
      T pm = T(1), p = sp, dpm = T(0), dp = T(1);
      T E = P.w[0], g = T(0);
      if (P.degree >= 1) { E += P.w[1] * sp; g = P.w[1]; }
      for (int n = 1; n < P.degree; ++n) {
        const T fn = T(n), inv = T(1) / (fn + T(1)), c = T(2) * fn + T(1);
        const T pn = (c * sp * p - fn * pm) * inv;
        const T dpn = Math<T>::fma(c, p, dpm);
        g = Math<T>::fma(P.w[n + 1], dpn, g);
        E = Math<T>::fma(P.w[n + 1], pn, E);
        pm = p; p = pn; dpm = dp; dp = dpn;
      }
      g = inside ? T(2) * g : T(0);
      const T il = T(1) / P.lam;
      gx += g * s * (dL1x - dL0x) - P.phi * dL0x * il;
      gy += g * s * (dL1y - dL0y) - P.phi * dL0y * il;
      if (WANT_E) e += E - P.phi * (amax + Math<T>::log(S0)) * il;
    } else if (P.kind == PYVES_BIAS_MLP_1D) {
      const T s = ((P.axis ? y : x) - P.s0) * P.k;
      T h[kGenH], dh[kGenH], hn[kGenH], dhn[kGenH];
      h[0] = s;
      dh[0] = T(1);
      const T* th = P.theta;
      for (int l = 0; l < P.n_layers; ++l) {
        const int nin = P.sizes[l], nout = P.sizes[l + 1];
        const T* W = th;
        const T* b = th + nin * nout;
        th = b + nout;
        const bool activate = (l >= 1) && (l < P.n_layers - 1);   // ves/basis.py:197-205
        for (int j = 0; j < nout; ++j) {
          T z = b[j], dz = T(0);
          for (int i = 0; i < nin; ++i) {
            z = Math<T>::fma(W[j * nin + i], h[i], z);
            dz = Math<T>::fma(W[j * nin + i], dh[i], dz);
          }
          if (activate) {
            T d;
            z = P.act == PYVES_ACT_RELU ? act_fn<T, PYVES_ACT_RELU>(z, d) : act_fn<T, PYVES_ACT_TANH>(z, d);
            dz *= d;
          }
          hn[j] = z;
          dhn[j] = dz;
        }
        for (int j = 0; j < nout; ++j) { h[j] = hn[j]; dh[j] = dhn[j]; }
      }
      const T g = dh[0] * P.k;
      gx += P.axis ? T(0) : g;
      gy += P.axis ? g : T(0);
      if (WANT_E) e += h[0];
    }
  }
};

}  // namespace pv
