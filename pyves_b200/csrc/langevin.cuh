// Fused Langevin step kernels: potential force + bias force + Philox noise + integrator update
// for `nsteps` consecutive steps per launch with the walker state held in registers.
// Replaces the per-step round trip integrator.step(1) -> CustomExternalForce + TorchForce ->
// LangevinIntegrator kernels of ves/langevin_dynamics.py:188 (OpenMM update restated in
// SURVEY Appendix A.1; CPU twin oracle/langevin.py).
#pragma once
#include "bias.cuh"
#include "philox.cuh"
#include "potentials.cuh"

namespace pv {

constexpr int kStepThreads = 128;
// leading columns of the static 16 x 16 sweep that keep their weights in the constant bank (bias.cuh, hybrid sweep)
constexpr int kHybrid16 = 4;

template <typename T> struct IntegratorParams {
  T a;              // exp(-gamma dt)
  T b_over_m;       // (1 - a) / gamma / m
  T c_over_sqrtm;   // sqrt(kT (1 - a^2) / m)
  T dt;
  T dt_over_m;      // middle scheme
  int scheme;
};

template <typename T> struct StepArgs {
  T* pos;           // SoA [dims][n]
  T* vel;
  int64_t n;        // walkers on this GPU
  uint64_t gid0;    // global id of local walker 0
  int64_t step0;    // global index of the first step of this launch
  int64_t nsteps;
  const T* noise;   // NULL (Philox) or injected N(0,1) draws [nsteps][dims][n]
  // general kernel only
  T* traj;          // NULL or [n_frames][dims][n_rec]
  int64_t every, n_rec;
};

// ------------------------------------------------------------------------------------------------
// Fast path: 2 simulated dimensions, OpenMM Langevin scheme, WPT walkers per thread, at least
// MINB resident blocks per SM (register cap).  One
// Philox4x32-10 call yields two Box-Muller pairs = the noise of two consecutive steps.
// INJECT selects the noise source at compile time (injected draws for the parity tests, Philox
// otherwise): with no branch inside the time loop the whole two-step body is one basic block, so
// ptxas interleaves the integer Philox rounds and the MUFU chain of Box-Muller with the FMA stream
// of the bias instead of running them back to back.
template <typename T, class Pot, class Bias, int WPT, int MINB, bool INJECT>
__global__ void __launch_bounds__(kStepThreads, MINB)
langevin2d_kernel(const StepArgs<T> A, const IntegratorParams<T> I, const PotParams<T> PP,
                  const typename Bias::Params BP, const RngKeys K) {
  extern __shared__ __align__(16) unsigned char smem[];
  Bias::stage(BP, smem);
  __syncthreads();

  const int64_t base = (int64_t)blockIdx.x * (kStepThreads * WPT) + threadIdx.x;
  int64_t w[WPT];
  bool live[WPT];
  T x[WPT], y[WPT], vx[WPT], vy[WPT];
#pragma unroll
  for (int k = 0; k < WPT; ++k) {
    const int64_t idx = base + (int64_t)k * kStepThreads;
    live[k] = idx < A.n;
    w[k] = live[k] ? idx : A.n - 1;
    x[k] = A.pos[w[k]];
    y[k] = A.pos[A.n + w[k]];
    vx[k] = A.vel[w[k]];
    vy[k] = A.vel[A.n + w[k]];
  }

  auto advance = [&](int k, T zx, T zy) {
    T gx = T(0), gy = T(0), e = T(0);
    Pot::template grad<false>(PP, x[k], y[k], e, gx, gy);
    Bias::template grad<false>(BP, smem, x[k], y[k], e, gx, gy);
    vx[k] = Math<T>::fma(I.c_over_sqrtm, zx, Math<T>::fma(-I.b_over_m, gx, I.a * vx[k]));
    vy[k] = Math<T>::fma(I.c_over_sqrtm, zy, Math<T>::fma(-I.b_over_m, gy, I.a * vy[k]));
    x[k] = Math<T>::fma(I.dt, vx[k], x[k]);
    y[k] = Math<T>::fma(I.dt, vy[k], y[k]);
  };
  // the heavy one-walker kernels sit at the register limit: they keep the plain call (Philox is < 5 % there)
  constexpr bool kSeq = WPT >= 2;
  PhiloxSeq SQ[WPT];
  if (!INJECT && kSeq) {
#pragma unroll
    for (int k = 0; k < WPT; ++k) SQ[k].init(A.gid0 + (uint64_t)w[k], (uint64_t)(A.step0 >> 1), kStreamDyn2, K);
  }
  // noise of steps 2c and 2c+1 for walker k
  auto draw = [&](int k, int64_t t_even, T (&z)[4]) {
    if (INJECT) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t ts = t_even + (q >> 1) - A.step0;
        z[q] = (ts >= 0 && ts < A.nsteps) ? A.noise[(ts * 2 + (q & 1)) * A.n + w[k]] : T(0);
      }
    } else {
      // call t_even / 2; the calls of a launch are consecutive, which the light kernels exploit (PhiloxSeq)
      const uint4 r = kSeq ? SQ[k].next(K) : philox_call(A.gid0 + (uint64_t)w[k], (uint64_t)(t_even >> 1), kStreamDyn2, K);
      box_muller(r.x, r.y, z[0], z[1]);
      box_muller(r.z, r.w, z[2], z[3]);
    }
  };

  int64_t t = A.step0;
  const int64_t end = A.step0 + A.nsteps;
  if ((t & 1) && t < end) {   // odd head: second half of its Philox call
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
      T z[4];
      draw(k, t - 1, z);
      advance(k, z[2], z[3]);
    }
    ++t;
  }
  for (; t + 1 < end; t += 2) {
    T z[WPT][4];
#pragma unroll
    for (int k = 0; k < WPT; ++k) draw(k, t, z[k]);
#pragma unroll
    for (int k = 0; k < WPT; ++k) advance(k, z[k][0], z[k][1]);
#pragma unroll
    for (int k = 0; k < WPT; ++k) advance(k, z[k][2], z[k][3]);
  }
  if (t < end) {              // even tail: first half of its Philox call
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
      T z[4];
      draw(k, t, z);
      advance(k, z[0], z[1]);
    }
  }

#pragma unroll
  for (int k = 0; k < WPT; ++k) {
    if (live[k]) {
      A.pos[w[k]] = x[k];
      A.pos[A.n + w[k]] = y[k];
      A.vel[w[k]] = vx[k];
      A.vel[A.n + w[k]] = vy[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Packed fast path for the light biases (none, Legendre 1D): two walkers per thread held as Pair<T> (lo, hi),
// so the potential, the bias recursion, the Box-Muller scaling and the integrator are packed FP32x2
// instructions -- one issue slot per two walkers.  These steps are issue bound (77 instructions per walker-step
// for 49 FMA-pipe cycles in the scalar kernel); Philox and the MUFU calls stay scalar.  Same arithmetic per
// walker as langevin2d_kernel (each half of an FFMA2 is an IEEE FMA).
template <typename T> PV_DEV void box_muller_pair(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, Pair<T>& z0, Pair<T>& z1);

template <> PV_DEV void box_muller_pair<float>(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, Pair<float>& z0, Pair<float>& z1) {
  using Q = Pair<float>;
  const Q u1 = Q::fma_bcast(Q::make(__uint2float_rn(a0), __uint2float_rn(b0)), 2.3283064365386963e-10f,
                            Q::make(1.1641532182693481e-10f, 1.1641532182693481e-10f));
  float la, lb;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la) : "f"(u1.lo()));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lb) : "f"(u1.hi()));
  const Q m2ln = Q::mul_bcast(Q::make(la, lb), -1.3862943611198906f);
  float ra, rb;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(m2ln.lo()));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(m2ln.hi()));
  const Q th = Q::fma_bcast(Q::make(__uint2float_rn(a1), __uint2float_rn(b1)), 1.4629180792671596e-09f,
                            Q::make(-3.1415926528583340f, -3.1415926528583340f));
  float sa, ca, sb, cb;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sa) : "f"(th.lo()));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(ca) : "f"(th.lo()));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sb) : "f"(th.hi()));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cb) : "f"(th.hi()));
  const Q rad = Q::make(ra, rb);
  z0 = Q::mul(rad, Q::make(ca, cb));
  z1 = Q::mul(rad, Q::make(sa, sb));
}

template <> PV_DEV void box_muller_pair<double>(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, Pair<double>& z0, Pair<double>& z1) {
  double za0, za1, zb0, zb1;
  box_muller(a0, a1, za0, za1);
  box_muller(b0, b1, zb0, zb1);
  z0 = Pair<double>::make(za0, zb0);
  z1 = Pair<double>::make(za1, zb1);
}

template <typename T, class Pot, class Bias, int MINB, bool INJECT>
__global__ void __launch_bounds__(kStepThreads, MINB)
langevin2d_pair_kernel(const StepArgs<T> A, const IntegratorParams<T> I, const PotParams<T> PP,
                       const typename Bias::Params BP, const RngKeys K) {
  using Q = Pair<T>;
  extern __shared__ __align__(16) unsigned char smem[];
  Bias::stage(BP, smem);
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * (kStepThreads * 2) + threadIdx.x;
  int64_t w[2];
  bool live[2];
  T x0[2], y0[2], vx0[2], vy0[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int64_t idx = base + (int64_t)k * kStepThreads;
    live[k] = idx < A.n;
    w[k] = live[k] ? idx : A.n - 1;
    x0[k] = A.pos[w[k]];
    y0[k] = A.pos[A.n + w[k]];
    vx0[k] = A.vel[w[k]];
    vy0[k] = A.vel[A.n + w[k]];
  }
  Q X = Q::make(x0[0], x0[1]), Y = Q::make(y0[0], y0[1]), VX = Q::make(vx0[0], vx0[1]), VY = Q::make(vy0[0], vy0[1]);

  auto advance = [&](Q ZX, Q ZY) {
    Q GX = Q::make(T(0), T(0)), GY = GX;
    PotPair<Pot, T>::grad2(PP, X, Y, GX, GY);
    BiasPair<Bias, T>::grad2(BP, smem, X, Y, GX, GY);
    VX = Q::fma_bcast(ZX, I.c_over_sqrtm, Q::fma_bcast(GX, -I.b_over_m, Q::mul_bcast(VX, I.a)));
    VY = Q::fma_bcast(ZY, I.c_over_sqrtm, Q::fma_bcast(GY, -I.b_over_m, Q::mul_bcast(VY, I.a)));
    X = Q::fma_bcast(VX, I.dt, X);
    Y = Q::fma_bcast(VY, I.dt, Y);
  };
  PhiloxSeq SA, SB;
  if (!INJECT) {
    SA.init(A.gid0 + (uint64_t)w[0], (uint64_t)(A.step0 >> 1), kStreamDyn2, K);
    SB.init(A.gid0 + (uint64_t)w[1], (uint64_t)(A.step0 >> 1), kStreamDyn2, K);
  }
  // noise of steps 2c and 2c+1 for both walkers: Z[0], Z[1] = (zx, zy) of the even step, Z[2], Z[3] of the odd one
  auto draw = [&](int64_t t_even, Q (&Z)[4]) {
    if (INJECT) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t ts = t_even + (q >> 1) - A.step0;
        const bool ok = ts >= 0 && ts < A.nsteps;
        Z[q] = Q::make(ok ? A.noise[(ts * 2 + (q & 1)) * A.n + w[0]] : T(0), ok ? A.noise[(ts * 2 + (q & 1)) * A.n + w[1]] : T(0));
      }
    } else {
      const uint4 ra = SA.next(K);     // call t_even / 2: the calls of a launch are consecutive
      const uint4 rb = SB.next(K);
      box_muller_pair<T>(ra.x, ra.y, rb.x, rb.y, Z[0], Z[1]);
      box_muller_pair<T>(ra.z, ra.w, rb.z, rb.w, Z[2], Z[3]);
    }
  };

  int64_t t = A.step0;
  const int64_t end = A.step0 + A.nsteps;
  if ((t & 1) && t < end) {   // odd head: second half of its Philox call
    Q Z[4];
    draw(t - 1, Z);
    advance(Z[2], Z[3]);
    ++t;
  }
  for (; t + 1 < end; t += 2) {
    Q Z[4];
    draw(t, Z);
    advance(Z[0], Z[1]);
    advance(Z[2], Z[3]);
  }
  if (t < end) {              // even tail: first half of its Philox call
    Q Z[4];
    draw(t, Z);
    advance(Z[0], Z[1]);
  }
  const T xo[2] = {X.lo(), X.hi()}, yo[2] = {Y.lo(), Y.hi()}, vxo[2] = {VX.lo(), VX.hi()}, vyo[2] = {VY.lo(), VY.hi()};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (live[k]) {
      A.pos[w[k]] = xo[k];
      A.pos[A.n + w[k]] = yo[k];
      A.vel[w[k]] = vxo[k];
      A.vel[A.n + w[k]] = vyo[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// General path: 2 or 3 dimensions, both schemes, optional trajectory recording, one walker per
// thread, any potential (PotGeneric) and any bias (BiasAny).  Used for reference-format runs
// (n_walkers = 1, z simulated, traj.dat) and for the cases the fast table does not cover.
template <typename T> struct BiasAnyParams {
  int kind;
  Legendre1DParams<T> l1;
  Legendre2DSmemParams<T> l2;
  OtherParams<T> other;
};

template <typename T> struct BiasAny {
  using Params = BiasAnyParams<T>;
  using L1 = BiasLegendre1D<T, 0>;
  using L2 = BiasLegendre2DSmem<T, 32, 2, 5>;
  using Ot = BiasOther<T>;
  static size_t smem_bytes(const Params& P) {
    if (P.kind == PYVES_BIAS_LEGENDRE_2D) return L2::smem_bytes(P.l2);
    if (P.kind == PYVES_BIAS_NONE || P.kind == PYVES_BIAS_LEGENDRE_1D) return 0;
    return Ot::smem_bytes(P.other);
  }
  static PV_DEV void stage(const Params& P, void* smem) {
    if (P.kind == PYVES_BIAS_LEGENDRE_2D) L2::stage(P.l2, smem);
    else if (P.kind != PYVES_BIAS_NONE && P.kind != PYVES_BIAS_LEGENDRE_1D) Ot::stage(P.other, smem);
  }
  template <bool WANT_E> static PV_DEV void grad(const Params& P, const void* smem, T x, T y, T& e, T& gx, T& gy) {
    switch (P.kind) {
      case PYVES_BIAS_NONE: break;
      case PYVES_BIAS_LEGENDRE_1D: L1::template grad<WANT_E>(P.l1, smem, x, y, e, gx, gy); break;
      case PYVES_BIAS_LEGENDRE_2D: L2::template grad<WANT_E>(P.l2, smem, x, y, e, gx, gy); break;
      default: Ot::template grad<WANT_E>(P.other, smem, x, y, e, gx, gy); break;
    }
  }
};

template <typename T, int D>
__global__ void __launch_bounds__(kStepThreads)
langevin_general_kernel(const StepArgs<T> A, const IntegratorParams<T> I, const PotParams<T> PP,
                        const BiasAnyParams<T> BP, const RngKeys K) {
  extern __shared__ __align__(16) unsigned char smem[];
  BiasAny<T>::stage(BP, smem);
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
  const bool live = idx < A.n;
  const int64_t w = live ? idx : A.n - 1;
  T q[3] = {T(0), T(0), T(0)}, v[3] = {T(0), T(0), T(0)};
#pragma unroll
  for (int d = 0; d < D; ++d) {
    q[d] = A.pos[d * A.n + w];
    v[d] = A.vel[d * A.n + w];
  }
  for (int64_t ts = 0; ts < A.nsteps; ++ts) {
    const int64_t t = A.step0 + ts;
    if (A.traj != nullptr && live && w < A.n_rec && ts % A.every == 0) {
      const int64_t f = ts / A.every;
#pragma unroll
      for (int d = 0; d < D; ++d) A.traj[(f * D + d) * A.n_rec + w] = q[d];
    }
    T z[4] = {T(0), T(0), T(0), T(0)};
    if (A.noise != nullptr) {
#pragma unroll
      for (int d = 0; d < D; ++d) z[d] = A.noise[(ts * D + d) * A.n + w];
    } else if (D == 2) {
      const uint4 r = philox_call(A.gid0 + (uint64_t)w, (uint64_t)(t >> 1), kStreamDyn2, K);
      if (t & 1) box_muller(r.z, r.w, z[0], z[1]);
      else box_muller(r.x, r.y, z[0], z[1]);
    } else {
      const uint4 r = philox_call(A.gid0 + (uint64_t)w, (uint64_t)t, kStreamDyn3, K);
      box_muller(r.x, r.y, z[0], z[1]);
      box_muller(r.z, r.w, z[2], z[3]);
    }
    T g[3] = {T(0), T(0), T(0)}, e = T(0);
    PotGeneric<T>::template grad<false>(PP, q[0], q[1], e, g[0], g[1]);
    BiasAny<T>::template grad<false>(BP, smem, q[0], q[1], e, g[0], g[1]);
    if (D == 3) g[2] = T(2000) * q[2];   // 1000 z^2, ves/potentials.py:105
    if (I.scheme == PYVES_SCHEME_OPENMM_LANGEVIN) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        v[d] = Math<T>::fma(I.c_over_sqrtm, z[d], Math<T>::fma(-I.b_over_m, g[d], I.a * v[d]));
        q[d] = Math<T>::fma(I.dt, v[d], q[d]);
      }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        v[d] = Math<T>::fma(-I.dt_over_m, g[d], v[d]);
        q[d] = Math<T>::fma(T(0.5) * I.dt, v[d], q[d]);
        v[d] = Math<T>::fma(I.c_over_sqrtm, z[d], I.a * v[d]);
        q[d] = Math<T>::fma(T(0.5) * I.dt, v[d], q[d]);
      }
    }
  }
  if (live) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      A.pos[d * A.n + w] = q[d];
      A.vel[d * A.n + w] = v[d];
    }
  }
}

// copies n converted weights (device, element type T, transposed wt[j][i]) into the __constant__ array the
// BiasLegendre2DStaticT<.., true> kernels read (same translation unit as those kernels); stream ordered
template <typename T> cudaError_t upload_static_weights(const T* dev_src, int n, cudaStream_t st);

// experimental tensor-core step kernel for the 2D Legendre bias (langevin_tc.cu); cudaErrorNotSupported when it does not apply
// `scratch`: device buffer of langevin2d_tc_scratch_bytes() bytes owned by the caller: the operand image of the folded coefficient
// matrices, built by langevin2d_tc_prepare (a one-block kernel on `st`) whenever the coefficients change and read by every launch
size_t langevin2d_tc_scratch_bytes();
cudaError_t langevin2d_tc_prepare(const Legendre2DSmemParams<float>& BP, void* scratch, cudaStream_t st);
int langevin2d_tc_walkers_per_cta(int kx, int ky);     // of the shape the launcher picks (one CTA per SM)
int langevin2d_tc_mlp_walkers_per_cta(int H, int act);
cudaError_t launch_langevin2d_tc(const StepArgs<float>& A, const IntegratorParams<float>& I, const PotParams<float>& PP,
                                 const Legendre2DSmemParams<float>& BP, const RngKeys& K, void* scratch, cudaStream_t st);

cudaError_t launch_langevin2d_tc_mlp(const StepArgs<float>& A, const IntegratorParams<float>& I, const PotParams<float>& PP,
                                     const MlpParams<float>& BP, int H, int act, const RngKeys& K, cudaStream_t st);

// launchers implemented in the langevin_*.cu translation units
template <typename T, class Pot, class Bias, int WPT, int MINB>
cudaError_t launch_langevin2d(const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,
                              const typename Bias::Params& BP, const RngKeys& K, cudaStream_t st);
template <typename T, int D>
cudaError_t launch_langevin_general(const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,
                                    const BiasAnyParams<T>& BP, const RngKeys& K, cudaStream_t st);

template <typename T, class Pot, class Bias, int MINB>
cudaError_t launch_langevin2d_pair(const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,
                                   const typename Bias::Params& BP, const RngKeys& K, cudaStream_t st);

template <typename T, class Pot, class Bias, int MINB>
cudaError_t launch_langevin2d_pair_impl(const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,
                                        const typename Bias::Params& BP, const RngKeys& K, cudaStream_t st) {
  auto kern = A.noise != nullptr ? langevin2d_pair_kernel<T, Pot, Bias, MINB, true> : langevin2d_pair_kernel<T, Pot, Bias, MINB, false>;
  const size_t sm = Bias::smem_bytes(BP);
  const int64_t per_block = (int64_t)kStepThreads * 2;
  const unsigned grid = (unsigned)((A.n + per_block - 1) / per_block);
  kern<<<grid, kStepThreads, sm, st>>>(A, I, PP, BP, K);
  return cudaGetLastError();
}

template <typename T, class Pot, class Bias, int WPT, int MINB>
cudaError_t launch_langevin2d_impl(const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,
                                   const typename Bias::Params& BP, const RngKeys& K, cudaStream_t st) {
  auto kern = A.noise != nullptr ? langevin2d_kernel<T, Pot, Bias, WPT, MINB, true>
                                 : langevin2d_kernel<T, Pot, Bias, WPT, MINB, false>;
  const size_t sm = Bias::smem_bytes(BP);
  if (sm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  const int64_t per_block = (int64_t)kStepThreads * WPT;
  const unsigned grid = (unsigned)((A.n + per_block - 1) / per_block);
  kern<<<grid, kStepThreads, sm, st>>>(A, I, PP, BP, K);
  return cudaGetLastError();
}

template <typename T, int D>
cudaError_t launch_langevin_general_impl(const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,
                                         const BiasAnyParams<T>& BP, const RngKeys& K, cudaStream_t st) {
  auto kern = langevin_general_kernel<T, D>;
  const size_t sm = BiasAny<T>::smem_bytes(BP);
  if (sm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  const unsigned grid = (unsigned)((A.n + kStepThreads - 1) / kStepThreads);
  kern<<<grid, kStepThreads, sm, st>>>(A, I, PP, BP, K);
  return cudaGetLastError();
}

}  // namespace pv
