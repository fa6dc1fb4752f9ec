#include "misc_kernels.cuh"

namespace pv {

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) eval_bias_kernel(const T* __restrict__ pos, int64_t n, int ncols,
                                                        T* __restrict__ E, T* __restrict__ F,
                                                        const BiasAnyParams<T> BP) {
  extern __shared__ __align__(16) unsigned char smem[];
  BiasAny<T>::stage(BP, smem);
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const T x = pos[r * ncols];
  const T y = ncols > 1 ? pos[r * ncols + 1] : T(0);
  T e = T(0), gx = T(0), gy = T(0);
  BiasAny<T>::template grad<true>(BP, smem, x, y, e, gx, gy);
  if (E != nullptr) E[r] = e;
  if (F != nullptr) {
    F[r * ncols] = -gx;
    if (ncols > 1) F[r * ncols + 1] = -gy;
    if (ncols > 2) F[r * ncols + 2] = T(0);
  }
}

template <typename T>
cudaError_t launch_eval_bias(const T* pos, int64_t n, int ncols, T* E, T* F, const BiasAnyParams<T>& BP,
                             cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  auto kern = eval_bias_kernel<T>;
  const size_t sm = BiasAny<T>::smem_bytes(BP);
  if (sm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)((n + 127) / 128), 128, sm, st>>>(pos, n, ncols, E, F, BP);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) energies_kernel(const T* __restrict__ pos, const T* __restrict__ vel, int64_t n,
                                                       int dims, T* __restrict__ PE, T* __restrict__ KE, T mass, T dt,
                                                       const PotParams<T> PP, const BiasAnyParams<T> BP) {
  extern __shared__ __align__(16) unsigned char smem[];
  BiasAny<T>::stage(BP, smem);
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const T x = pos[r], y = pos[n + r];
  const T z = dims > 2 ? pos[2 * n + r] : T(0);
  T e = T(0), g[3] = {T(0), T(0), T(0)};
  PotGeneric<T>::template grad<true>(PP, x, y, e, g[0], g[1]);
  BiasAny<T>::template grad<true>(BP, smem, x, y, e, g[0], g[1]);
  if (dims > 2) {
    e += T(1000) * z * z;
    g[2] = T(2000) * z;
  }
  T ke = T(0);
  for (int d = 0; d < dims; ++d) {
    const T vs = vel[d * n + r] - T(0.5) * dt * g[d] / mass;   // leap-frog half-step shift [UPSTREAM OpenMM]
    ke += vs * vs;
  }
  if (PE != nullptr) PE[r] = e;
  if (KE != nullptr) KE[r] = T(0.5) * mass * ke;
}

template <typename T>
cudaError_t launch_energies(const T* pos, const T* vel, int64_t n, int dims, T* PE, T* KE, T mass, T dt,
                            const PotParams<T>& PP, const BiasAnyParams<T>& BP, cudaStream_t st) {
  auto kern = energies_kernel<T>;
  const size_t sm = BiasAny<T>::smem_bytes(BP);
  if (sm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)((n + 127) / 128), 128, sm, st>>>(pos, vel, n, dims, PE, KE, mass, dt, PP, BP);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) init_positions_kernel(T* __restrict__ pos, int64_t n, int dims, uint64_t gid0,
                                                             double xmin, double xmax, double ymin, double ymax,
                                                             double beta, int ntrials, const PotParams<double> PP,
                                                             const RngKeys K, unsigned long long* fail_count) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  // Acceptance is decided in double so that every precision places walkers identically.
  const double thresh = 0.6931471805599453 / beta;   // exp(-beta U) >= 0.5  <=>  U <= ln 2 / beta
  double px = 0.0, py = 0.0;
  bool ok = false;
  for (int trial = 0; trial < ntrials && !ok; ++trial) {
    const uint4 o = philox_call(gid0 + (uint64_t)r, (uint64_t)trial, kStreamPos, K);
    const double tx = xmin + (xmax - xmin) * uniform01(o.x, 0.0);
    const double ty = ymin + (ymax - ymin) * uniform01(o.y, 0.0);
    double e = 0.0, gx = 0.0, gy = 0.0;
    PotGeneric<double>::template grad<true>(PP, tx, ty, e, gx, gy);
    if (e <= thresh) {
      px = tx;
      py = ty;
      ok = true;
    }
  }
  if (!ok) atomicAdd(fail_count, 1ull);
  pos[r] = (T)px;
  pos[n + r] = (T)py;
  if (dims > 2) pos[2 * n + r] = T(0);
}

template <typename T>
cudaError_t launch_init_positions(T* pos, int64_t n, int dims, uint64_t gid0, const double* box, double beta,
                                  int ntrials, const PotParams<double>& PD, const RngKeys& K,
                                  unsigned long long* fail_count, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(fail_count, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  init_positions_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pos, n, dims, gid0, box[0], box[1], box[2],
                                                                        box[3], beta, ntrials, PD, K, fail_count);
  return cudaGetLastError();
}

template <typename T>
__global__ void __launch_bounds__(128) init_velocities_kernel(T* __restrict__ vel, int64_t n, int dims, uint64_t gid0,
                                                              T sigma_v, const RngKeys K) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint4 o = philox_call(gid0 + (uint64_t)r, 0ull, kStreamVel, K);
  T z[4];
  box_muller(o.x, o.y, z[0], z[1]);
  box_muller(o.z, o.w, z[2], z[3]);
  for (int d = 0; d < dims; ++d) vel[d * n + r] = sigma_v * z[d];
}

template <typename T>
cudaError_t launch_init_velocities(T* vel, int64_t n, int dims, uint64_t gid0, T sigma_v, const RngKeys& K,
                                   cudaStream_t st) {
  init_velocities_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(vel, n, dims, gid0, sigma_v, K);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Histogram: per-block shared-memory bins (float counts) when they fit, flushed with one double
// atomic per non-empty bin per block; global double atomics otherwise.
constexpr int kHistSmemBins = 11 * 1024;

__device__ __forceinline__ int hist_bin(double v, double lo, double hi, int nb) {
  if (!(v >= lo && v <= hi)) return -1;
  int b = (int)((v - lo) / (hi - lo) * nb);
  return b >= nb ? nb - 1 : b;   // right edge belongs to the last bin (numpy)
}

template <typename T>
__global__ void __launch_bounds__(256) histogram_kernel(const T* __restrict__ pos, int64_t n, int64_t stride_row,
                                                        int64_t stride_col, int ndim, int axis, double lo0, double hi0,
                                                        double lo1, double hi1, int nb0, int nb1,
                                                        const T* __restrict__ w, double* __restrict__ counts,
                                                        int use_smem) {
  extern __shared__ float bins[];
  const int nb = nb0 * nb1;
  if (use_smem) {
    for (int i = threadIdx.x; i < nb; i += blockDim.x) bins[i] = 0.f;
    __syncthreads();
  }
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    int b;
    if (ndim == 1) {
      b = hist_bin((double)pos[r * stride_row + axis * stride_col], lo0, hi0, nb0);
    } else {
      const int bx = hist_bin((double)pos[r * stride_row], lo0, hi0, nb0);
      const int by = hist_bin((double)pos[r * stride_row + stride_col], lo1, hi1, nb1);
      b = (bx < 0 || by < 0) ? -1 : bx * nb1 + by;
    }
    if (b < 0) continue;
    const float wt = w != nullptr ? (float)w[r] : 1.f;
    if (use_smem) atomicAdd(&bins[b], wt);
    else atomicAdd(&counts[b], (double)wt);
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
      if (bins[i] != 0.f) atomicAdd(&counts[i], (double)bins[i]);
  }
}

template <typename T>
cudaError_t launch_histogram(const T* pos, int64_t n, int64_t stride_row, int64_t stride_col, int ndim, int axis,
                             const double* range, const int* nbins, const T* row_weights, double* counts,
                             cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const int nb0 = nbins[0], nb1 = ndim == 2 ? nbins[1] : 1;
  const int nb = nb0 * nb1;
  const int use_smem = nb <= kHistSmemBins;
  int64_t blocks = (n + 255) / 256;
  // few blocks per SM so that each flush amortises over many rows; float bins stay exact (< 2^24 per block)
  if (blocks > 148 * 4) blocks = 148 * 4;
  histogram_kernel<T><<<(unsigned)blocks, 256, use_smem ? nb * sizeof(float) : 0, st>>>(
      pos, n, stride_row, stride_col, ndim, axis, range[0], range[1], ndim == 2 ? range[2] : 0.0,
      ndim == 2 ? range[3] : 1.0, nb0, nb1, row_weights, counts, use_smem);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// FP32 FMA peak: 16 independent FFMA chains per thread, register operands only.
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, float b, float c, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;   // never true; keeps the chains alive
}

// Same probe with the packed FP32x2 instruction (FFMA2): 8 chains of pairs per thread.
__global__ void __launch_bounds__(256) fma2_peak_kernel(float* out, float b, float c, int iters) {
  Pair<float> a[8];
  const Pair<float> cc = Pair<float>::make(c, c);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = Pair<float>::make((float)(threadIdx.x + i) * 1e-3f, (float)i * 2e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = Pair<float>::fma_bcast(a[i], b, cc);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].lo() + a[i].hi();
  if (s == 123.456f) out[0] = s;
}

cudaError_t fp32_fma_peak(double* tflops, int packed) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* out = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(float));
  if (e != cudaSuccess) return e;
  const int iters = 4096, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    if (packed) fma2_peak_kernel<<<blocks, threads>>>(out, 0.999f, 1e-3f, iters);
    else fma_peak_kernel<<<blocks, threads>>>(out, 0.999f, 1e-3f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * 8 * (double)iters * blocks * threads;
    const double tf = flops / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return cudaGetLastError();
}

template <typename T>
__global__ void convert_weights2d_kernel(const double* __restrict__ w64, int kx, int ky, T* __restrict__ w_rm, T* __restrict__ w_t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kx * ky) return;
  const int i = idx / ky, j = idx - i * ky;
  const T v = (T)w64[idx];
  w_rm[idx] = v;
  w_t[j * kx + i] = v;
}

template <typename T>
cudaError_t launch_convert_weights2d(const double* w64, int kx, int ky, T* w_rm, T* w_t, cudaStream_t st) {
  convert_weights2d_kernel<T><<<(kx * ky + 255) / 256, 256, 0, st>>>(w64, kx, ky, w_rm, w_t);
  return cudaGetLastError();
}

// one warp per coefficient k: the lanes split the Hessian row (fixed order, butterfly), lane 0 applies the update
__global__ void __launch_bounds__(256) averaged_sgd_kernel(int K, const double* __restrict__ sum_w, const double* __restrict__ sum_wf,
                                                           const double* __restrict__ sum_wff, const double* __restrict__ f_target,
                                                           const double* __restrict__ alpha, const double* __restrict__ abar,
                                                           const double* __restrict__ scale, double mu, double beta, int averaging, double decay,
                                                           long long n_after, double* __restrict__ alpha_out, double* __restrict__ abar_out) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  if (!(sum_w[0] > 0.0)) {   // no sample in the averaging domain (every walker outside the basis box): keep the state
    if (lane == 0) {
      alpha_out[k] = alpha[k];
      abar_out[k] = abar[k];
    }
    return;
  }
  const double inv_w = 1.0 / sum_w[0];
  const double mean_k = sum_wf[k] * inv_w;
  double h = 0.0;
  if (sum_wff != nullptr) {
    for (int l = lane; l < K; l += 32) h += (sum_wff[(size_t)k * K + l] * inv_w - mean_k * (sum_wf[l] * inv_w)) * (alpha[l] - abar[l]);
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  }
  if (lane != 0) return;
  double d = f_target[k] - mean_k + beta * h;
  if (scale != nullptr) d *= scale[k];
  const double a = alpha[k] - mu * d;
  alpha_out[k] = a;
  abar_out[k] = averaging == 0 ? decay * abar[k] + (1.0 - decay) * a : abar[k] + (a - abar[k]) / (double)(n_after + 1);
}

cudaError_t launch_averaged_sgd_step(int K, const double* sum_w, const double* sum_wf, const double* sum_wff, const double* f_target,
                                     const double* alpha, const double* abar, const double* scale, double mu, double beta, int averaging,
                                     double decay, long long n_after, double* alpha_out, double* abar_out, cudaStream_t st) {
  averaged_sgd_kernel<<<(K + 7) / 8, 256, 0, st>>>(K, sum_w, sum_wf, sum_wff, f_target, alpha, abar, scale, mu, beta, averaging, decay, n_after,
                                                   alpha_out, abar_out);
  return cudaGetLastError();
}

template <typename T>
__global__ void convert_weights1d_kernel(const double* __restrict__ w64, int n, T* __restrict__ w_t, double* __restrict__ keep64) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = w64[i];
  w_t[i] = (T)v;
  keep64[i] = v;
}

template <typename T>
cudaError_t launch_convert_weights1d(const double* w64, int n, T* w_t, double* keep64, cudaStream_t st) {
  convert_weights1d_kernel<T><<<1, 64, 0, st>>>(w64, n, w_t, keep64);
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// One optimiser step of the NN-bias parameters on the device (FP64, elementwise, in place): the gradient of the
// VES functional from the ensemble sums, its optional running / exponentially weighted average over updates, and
// the torch.optim rule the reference delegates to (ves/bias.py:193-200 picks SGD / RMSprop / Adam, :236-245 steps).
// Same arithmetic, operation by operation, as torch's single-tensor CPU implementations (lerp as an FMA, addcmul,
// addcdiv), so the device loop tracks the host loop to rounding.
//   g = f_target - sum_wf / sum_w
//   gavg 1: G = n == 1 ? g : G + (g - G) / n          gavg 2: G = n == 1 ? g : decay G + (1 - decay) g      g <- G
//   SGD     : buf = t == 1 ? g : momentum buf + g (momentum != 0);  theta -= lr (buf | g)
//   RMSprop : sq = alpha sq + (1 - alpha) g g;  theta -= lr g / (sqrt(sq) + eps)
//   Adam    : m = m + (1 - b1)(g - m);  v = b2 v + (1 - b2) g g;  theta -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) optimizer_step_kernel(int kind, long long P, const double* __restrict__ sum_w,
                                                             const double* __restrict__ sum_wf, const double* __restrict__ f_target,
                                                             int gavg, double gdecay, long long n_avg, double* __restrict__ gstate,
                                                             double* __restrict__ theta, double* __restrict__ s1, double* __restrict__ s2,
                                                             OptHyper H, long long t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double sw = sum_w[0];
  if (!(sw > 0.0)) return;   // no sample in the averaging domain: parameters and optimiser state stay as they are
  double g = f_target[i] - sum_wf[i] / sw;
  if (gavg != 0) {
    double G = gstate[i];
    if (n_avg == 1) G = g;
    else if (gavg == 1) G = G + (g - G) / (double)n_avg;
    else G = gdecay * G + (1.0 - gdecay) * g;
    gstate[i] = G;
    g = G;
  }
  double th = theta[i];
  if (kind == PYVES_OPT_SGD) {
    if (H.a != 0.0) {                       // momentum (dampening 0, no Nesterov)
      const double buf = t == 1 ? g : H.a * s1[i] + g;
      s1[i] = buf;
      g = buf;
    }
    th = th + (-H.lr) * g;
  } else if (kind == PYVES_OPT_RMSPROP) {
    const double sq = s1[i] * H.a + (1.0 - H.a) * g * g;
    s1[i] = sq;
    th = th + (-H.lr) * g / (sqrt(sq) + H.eps);
  } else {
    const double m = fma(1.0 - H.a, g - s1[i], s1[i]);
    const double v = s2[i] * H.b + (1.0 - H.b) * g * g;
    s1[i] = m;
    s2[i] = v;
    const double denom = sqrt(v) / H.bc2_sqrt + H.eps;
    th = th + (-(H.lr / H.bc1)) * m / denom;
  }
  theta[i] = th;
}

cudaError_t launch_optimizer_step(int kind, long long P, const double* sum_w, const double* sum_wf, const double* f_target, int gavg,
                                  double gdecay, long long n_avg, double* gstate, double* theta, double* s1, double* s2,
                                  const OptHyper& H, long long t, cudaStream_t st) {
  optimizer_step_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(kind, P, sum_w, sum_wf, f_target, gavg, gdecay, n_avg, gstate, theta,
                                                                     s1, s2, H, t);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// NN parameters from a device-resident optimiser: theta64 (FP64, torch parameters() order) -> element-type copy
// theta_t (generic functor, gradient-moment kernels) and the packed layout of BiasMlp (bias.cuh): the activation-free
// pair Linear(1, h0) -> Linear(h0, h1) folded to z1 = u s + c in FP64 (same summation order as the host fold in
// pyves_set_bias_mlp), the third layer transposed, zero padding, the output bias at the end.  One thread per output.
template <typename T>
__global__ void __launch_bounds__(256) mlp_pack_kernel(const double* __restrict__ th, MlpPackPlan Q, T* __restrict__ theta_t,
                                                       T* __restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < Q.np) theta_t[idx] = (T)th[idx];
  const int p = idx - Q.np;
  if (p < 0 || p >= Q.n_packed) return;
  const int h0 = Q.h0, h1 = Q.h1, h2 = Q.h2;
  const double* W0 = th + Q.w0;
  const double* b0 = th + Q.b0;
  const double* W1 = th + Q.w1;
  const double* b1 = th + Q.b1;
  auto fold_u = [&](int j) {
    double su = 0.0;
    for (int i = 0; i < h0; ++i) su = __dadd_rn(su, __dmul_rn(W1[j * h0 + i], W0[i]));   // unfused, as the host compiler rounds it
    return su;
  };
  auto fold_c = [&](int j) {
    double sc = b1[j];
    for (int i = 0; i < h0; ++i) sc = __dadd_rn(sc, __dmul_rn(W1[j * h0 + i], b0[i]));
    return sc;
  };
  double v = 0.0;
  if (Q.H == 0) {                         // two hidden layers: u[Wp] c[Wp] wout[Wp] bout
    const int Wp = Q.width, r = p / Wp, j = p - r * Wp;
    if (p == 3 * Wp) v = th[Q.bo];
    else if (p < 3 * Wp && j < h1) v = r == 0 ? fold_u(j) : (r == 1 ? fold_c(j) : th[Q.wo + j]);
  } else {                                // three hidden layers: u[H] c[H] wout[H] b2[H] W2T[H][H] bout
    const int H = Q.H;
    if (p < 4 * H) {
      const int r = p / H, j = p - r * H;
      if (r == 0 && j < h1) v = fold_u(j);
      else if (r == 1 && j < h1) v = fold_c(j);
      else if (r == 2 && j < h2) v = th[Q.wo + j];
      else if (r == 3 && j < h2) v = th[Q.b2 + j];
    } else if (p < 4 * H + H * H) {
      const int q = p - 4 * H, i = q / H, j = q - i * H;
      if (i < h1 && j < h2) v = th[Q.w2 + j * h1 + i];      // transposed: W2T[i][j]
    } else if (p == 4 * H + H * H) {
      v = th[Q.bo];
    }
  }
  packed[p] = (T)v;
}

template <typename T>
cudaError_t launch_mlp_pack(const double* theta64, const MlpPackPlan& Q, T* theta_t, T* packed, cudaStream_t st) {
  const int total = Q.np + Q.n_packed;
  mlp_pack_kernel<T><<<(total + 255) / 256, 256, 0, st>>>(theta64, Q, theta_t, packed);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
#define PV_INST(T)                                                                                                 \
  template cudaError_t launch_eval_bias<T>(const T*, int64_t, int, T*, T*, const BiasAnyParams<T>&, cudaStream_t); \
  template cudaError_t launch_convert_weights2d<T>(const double*, int, int, T*, T*, cudaStream_t);                 \
  template cudaError_t launch_convert_weights1d<T>(const double*, int, T*, double*, cudaStream_t);                 \
  template cudaError_t launch_mlp_pack<T>(const double*, const MlpPackPlan&, T*, T*, cudaStream_t);                \
  template cudaError_t launch_energies<T>(const T*, const T*, int64_t, int, T*, T*, T, T, const PotParams<T>&,     \
                                          const BiasAnyParams<T>&, cudaStream_t);                                  \
  template cudaError_t launch_init_positions<T>(T*, int64_t, int, uint64_t, const double*, double, int,            \
                                                const PotParams<double>&, const RngKeys&, unsigned long long*,     \
                                                cudaStream_t);                                                     \
  template cudaError_t launch_init_velocities<T>(T*, int64_t, int, uint64_t, T, const RngKeys&, cudaStream_t);     \
  template cudaError_t launch_histogram<T>(const T*, int64_t, int64_t, int64_t, int, int, const double*,           \
                                           const int*, const T*, double*, cudaStream_t);
PV_INST(float)
PV_INST(double)

// ------------------------------------------------------------------------------------------------
// ReLU NNBasis1D -> piecewise-constant dV/ds (see BiasPwl1D in bias.cuh).  One block; all sums in FP64.
//   first layer (folded): unit i is active where u_i s + c_i > 0, i.e. on one side of tau_i = -c_i / u_i; the sorted tau cut the
//   line into intervals r = 0 .. M on which the active set is constant;
//   H == 0 (two hidden layers): V = bout + sum_i wout_i relu(u_i s + c_i), slope_r = sum over the active i of wout_i u_i;
//   H  > 0 (three hidden layers): on interval r, z2_j = alpha_rj s + beta_rj with alpha_rj = sum_active W2[j][i] u_i,
//   beta_rj = b2_j + sum_active W2[j][i] c_i; unit j of the second layer switches at kappa_rj = -beta_rj / alpha_rj if that lies
//   inside the interval; between two consecutive switch points slope = sum over the active j of wout_j alpha_rj.
// Active sets are evaluated at the midpoint of every piece (robust against ties), never inferred from the order of the switches.
constexpr int kPwlMaxUnits = 128;     // first-layer units (H <= 64, streamed width <= 128)
constexpr int kPwlThreads = 1024;

int mlp_pwl_capacity(int H, int width) {
  if (H > 0) return (H + 1) * (H + 1);
  return width + 1;
}

// grid[b] = number of breakpoints below the lower edge of bucket b of a uniform grid over [brk[0], brk[n - 1]] (block-wide; brk was
// written by this block and a barrier separates the writes from these reads); grid[kPwlBuckets] = n
template <typename T> PV_DEV void pwl_grid(const T* brk, int n, int* grid, T* grid_map) {
  const T g0 = n > 0 ? brk[0] : T(0);
  const T width = n > 1 ? brk[n - 1] - brk[0] : T(0);
  const T per = width > T(0) ? T(kPwlBuckets) / width : T(0);
  if (threadIdx.x == 0) {
    grid_map[0] = g0;
    grid_map[1] = per;
  }
  for (int b = threadIdx.x; b <= kPwlBuckets; b += blockDim.x) {
    int lo = 0, len = n;
    if (b == kPwlBuckets || per == T(0)) {
      lo = b == 0 ? 0 : n;
    } else {
      const T edge = g0 + T(b) / per;
      while (len > 0) {
        const int half = len >> 1;
        const bool right = brk[lo + half] < edge;
        lo = right ? lo + half + 1 : lo;
        len = right ? len - half - 1 : half;
      }
    }
    grid[b] = lo;
  }
}

template <typename T>
__global__ void __launch_bounds__(kPwlThreads) mlp_pwl_build_kernel(const T* __restrict__ packed, const int H, const int width,
                                                                   T* __restrict__ brk, T* __restrict__ slope, int* __restrict__ count,
                                                                   int* __restrict__ grid, T* __restrict__ grid_map) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n1 = H > 0 ? H : width;                   // first-layer units
  const int n2 = H > 0 ? H : 0;                       // second-layer units
  double* tau = reinterpret_cast<double*>(smem_raw);  // [n1] sorted thresholds
  double* alpha = tau + kPwlMaxUnits;                 // [(n1 + 1) * n2]
  double* beta = alpha + (size_t)(n1 + 1) * n2;       // [(n1 + 1) * n2]
  double* kink = beta + (size_t)(n1 + 1) * n2;        // [(n1 + 1) * n2] sorted switch points of the second layer inside interval r
  int* kcount = reinterpret_cast<int*>(kink + (size_t)(n1 + 1) * n2);   // [n1 + 1]
  int* base = kcount + (n1 + 2);                      // [n1 + 2] first piece of interval r
  __shared__ int M_sh;
  const T* u = packed;
  const T* c = packed + n1;
  const T* wout = packed + 2 * n1;
  const T* b2 = packed + 3 * n1;                      // H > 0 only
  const T* W2T = packed + 4 * n1;                     // H > 0 only: W2T[i][j] = W2[j][i]
  const int tid = threadIdx.x;
  if (tid == 0) M_sh = 0;
  __syncthreads();
  // ---- thresholds of the first layer, rank sort
  if (tid < n1) {
    const double ui = (double)u[tid], ci = (double)c[tid];
    if (ui != 0.0) {
      const double t = -ci / ui;
      int rank = 0;
      for (int k = 0; k < n1; ++k) {
        const double uk = (double)u[k];
        if (uk == 0.0) continue;
        const double tk = -(double)c[k] / uk;
        rank += (tk < t || (tk == t && k < tid)) ? 1 : 0;
      }
      tau[rank] = t;
      atomicAdd(&M_sh, 1);
    }
  }
  __syncthreads();
  const int M = M_sh;
  auto mid = [&](double lo, bool has_lo, double hi, bool has_hi) {   // a point strictly inside (lo, hi) where possible
    if (has_lo && has_hi) return 0.5 * (lo + hi);
    if (has_lo) return lo + 1.0;
    if (has_hi) return hi - 1.0;
    return 0.0;
  };
  if (n2 == 0) {
    for (int r = tid; r <= M; r += blockDim.x) {
      const double sp = mid(r > 0 ? tau[r - 1] : 0.0, r > 0, r < M ? tau[r] : 0.0, r < M);
      double g = 0.0;
      for (int i = 0; i < n1; ++i) {
        const double ui = (double)u[i];
        if (ui * sp + (double)c[i] > 0.0) g += (double)wout[i] * ui;
      }
      slope[r] = (T)g;
      if (r < M) brk[r] = (T)tau[r];
    }
    if (tid == 0) *count = M;
    __syncthreads();
    pwl_grid(brk, M, grid, grid_map);
    return;
  }
  // ---- second layer on every interval of the first
  for (int item = tid; item < (M + 1) * n2; item += blockDim.x) {
    const int r = item / n2, j = item - r * n2;
    const double sp = mid(r > 0 ? tau[r - 1] : 0.0, r > 0, r < M ? tau[r] : 0.0, r < M);
    double a = 0.0, b = (double)b2[j];
    for (int i = 0; i < n1; ++i) {
      const double ui = (double)u[i], ci = (double)c[i];
      if (ui * sp + ci > 0.0) {
        const double w = (double)W2T[i * n2 + j];
        a += w * ui;
        b += w * ci;
      }
    }
    alpha[item] = a;
    beta[item] = b;
  }
  __syncthreads();
  // ---- switch points of the second layer inside each interval, rank sort per interval
  for (int item = tid; item < (M + 1) * n2; item += blockDim.x) {
    const int r = item / n2, j = item - r * n2;
    const bool has_lo = r > 0, has_hi = r < M;
    const double lo = has_lo ? tau[r - 1] : 0.0, hi = has_hi ? tau[r] : 0.0;
    auto inside = [&](int jj, double& kap) {
      const double a = alpha[r * n2 + jj];
      if (a == 0.0) return false;
      kap = -beta[r * n2 + jj] / a;
      return (!has_lo || kap > lo) && (!has_hi || kap < hi);
    };
    double kap;
    if (inside(j, kap)) {
      int rank = 0;
      for (int jj = 0; jj < n2; ++jj) {
        double kk;
        if (jj != j && inside(jj, kk)) rank += (kk < kap || (kk == kap && jj < j)) ? 1 : 0;
      }
      kink[r * n2 + rank] = kap;
    }
  }
  for (int r = tid; r <= M; r += blockDim.x) {
    const bool has_lo = r > 0, has_hi = r < M;
    const double lo = has_lo ? tau[r - 1] : 0.0, hi = has_hi ? tau[r] : 0.0;
    int cnt = 0;
    for (int jj = 0; jj < n2; ++jj) {
      const double a = alpha[r * n2 + jj];
      if (a == 0.0) continue;
      const double kap = -beta[r * n2 + jj] / a;
      cnt += ((!has_lo || kap > lo) && (!has_hi || kap < hi)) ? 1 : 0;
    }
    kcount[r] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int r = 0; r <= M; ++r) {
      base[r] = acc;
      acc += kcount[r] + 1;
    }
    base[M + 1] = acc;                                // pieces in total
    *count = acc - 1;                                 // breakpoints
  }
  __syncthreads();
  // ---- pieces: (r, q), q = 0 .. kcount[r]; its upper end is breakpoint base[r] + q (the last piece has none)
  const int pieces = base[M + 1];
  for (int pidx = tid; pidx < pieces; pidx += blockDim.x) {
    int r = 0;
    while (base[r + 1] <= pidx) ++r;                  // <= 129 intervals
    const int q = pidx - base[r], kc = kcount[r];
    const bool has_lo = q > 0 || r > 0, has_hi = q < kc || r < M;
    const double lo = q > 0 ? kink[r * n2 + q - 1] : (r > 0 ? tau[r - 1] : 0.0);
    const double hi = q < kc ? kink[r * n2 + q] : (r < M ? tau[r] : 0.0);
    const double sp = mid(lo, has_lo, hi, has_hi);
    double g = 0.0;
    for (int j = 0; j < n2; ++j) {
      const double a = alpha[r * n2 + j];
      if (a * sp + beta[r * n2 + j] > 0.0) g += (double)wout[j] * a;
    }
    slope[pidx] = (T)g;
    if (has_hi) brk[pidx] = (T)hi;
  }
  __syncthreads();
  pwl_grid(brk, pieces - 1, grid, grid_map);
}

template <typename T>
cudaError_t launch_mlp_pwl_build(const T* packed, int H, int width, T* brk, T* slope, int* count, int* grid, T* grid_map, cudaStream_t st) {
  const int n1 = H > 0 ? H : width, n2 = H > 0 ? H : 0;
  if (n1 < 1 || n1 > kPwlMaxUnits) return cudaErrorInvalidValue;
  const size_t sm = sizeof(double) * (kPwlMaxUnits + 3 * (size_t)(n1 + 1) * n2) + sizeof(int) * (2 * (size_t)n1 + 8);
  auto kern = mlp_pwl_build_kernel<T>;
  if (sm > 48 * 1024 - 64) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  kern<<<1, kPwlThreads, sm, st>>>(packed, H, width, brk, slope, count, grid, grid_map);
  return cudaGetLastError();
}
template cudaError_t launch_mlp_pwl_build<float>(const float*, int, int, float*, float*, int*, int*, float*, cudaStream_t);
template cudaError_t launch_mlp_pwl_build<double>(const double*, int, int, double*, double*, int*, int*, double*, cudaStream_t);

}  // namespace pv
