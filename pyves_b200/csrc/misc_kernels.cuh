// Batched bias evaluation, per-walker energies, ensemble initialisation, histograms and the
// FP32-FMA peak probe.  Launchers are implemented in misc_kernels.cu for float and double.
#pragma once
#include "langevin.cuh"

namespace pv {

// TorchForce evaluation at caller positions pos[n][ncols] (ves/bias.py:58, ves/basis.py:123-156)
template <typename T>
cudaError_t launch_eval_bias(const T* pos, int64_t n, int ncols, T* E, T* F, const BiasAnyParams<T>& BP, cudaStream_t st);

// device-resident coefficient update: w64[kx][ky] (FP64, device) -> row-major T copy w_rm[kx][ky] (runtime
// functors, eval) and transposed T copy w_t[ky][kx] (static column-sweep functors)
template <typename T>
cudaError_t launch_convert_weights2d(const double* w64, int kx, int ky, T* w_rm, T* w_t, cudaStream_t st);

// w64[n] (FP64, device) -> element-type copy (source of the __constant__ upload of the 1D functors) + FP64 keep
template <typename T>
cudaError_t launch_convert_weights1d(const double* w64, int n, T* w_t, double* keep64, cudaStream_t st);

// One averaged-SGD update of the VES coefficients on the device (FP64), the host twin is pyves_b200/optim.py:
//   mean = sum_wf / sum_w;  d = f_target - mean  (+ beta (sum_wff / sum_w - mean mean^T) (alpha - abar) when sum_wff)
//   d *= scale (optional);  alpha' = alpha - mu d;  abar' = decay abar + (1 - decay) alpha'   (averaging 0, EWMA)
//                                                   abar' = abar + (alpha' - abar) / (n + 1)  (averaging 1, running)
// alpha_out / abar_out must not alias the inputs (the Hessian term reads every old entry).  sum_w <= 0 (no sample
// in the averaging domain): alpha and abar are copied through unchanged.
cudaError_t launch_averaged_sgd_step(int K, const double* sum_w, const double* sum_wf, const double* sum_wff, const double* f_target,
                                     const double* alpha, const double* abar, const double* scale, double mu, double beta, int averaging,
                                     double decay, long long n_after, double* alpha_out, double* abar_out, cudaStream_t st);

// torch.optim step on device-resident NN parameters (see optimizer_step_kernel).  a / b: SGD momentum | RMSprop alpha |
// Adam beta1, beta2;  bc1 = 1 - beta1^t, bc2_sqrt = sqrt(1 - beta2^t) (Adam bias corrections, computed by the caller).
struct OptHyper {
  double lr, a, b, eps, bc1, bc2_sqrt;
};
cudaError_t launch_optimizer_step(int kind, long long P, const double* sum_w, const double* sum_wf, const double* f_target, int gavg,
                                  double gdecay, long long n_avg, double* gstate, double* theta, double* s1, double* s2,
                                  const OptHyper& H, long long t, cudaStream_t st);

// layout of an NNBasis1D parameter vector (offsets into theta) and of the packed BiasMlp buffer it folds into
struct MlpPackPlan {
  int np;                 // parameters
  int n_packed;           // elements of the packed buffer (0: no fast functor, only the converted copy is written)
  int H, width;           // BiasMlp variant: H > 0 one H x H matmul; H == 0 streamed with padded width
  int h0, h1, h2;
  int w0, b0, w1, b1, w2, b2, wo, bo;
};
template <typename T>
cudaError_t launch_mlp_pack(const double* theta64, const MlpPackPlan& Q, T* theta_t, T* packed, cudaStream_t st);

// Compiles the packed ReLU NNBasis1D parameters (BiasMlp layout: H > 0 one H x H layer, H == 0 streamed with `width` units) into
// the piecewise-constant derivative dV/ds that BiasPwl1D evaluates: brk[count] ascending, slope[count + 1], count <= mlp_pwl_capacity.
int mlp_pwl_capacity(int H, int width);
template <typename T>
cudaError_t launch_mlp_pwl_build(const T* packed, int H, int width, T* brk, T* slope, int* count, int* grid, T* grid_map, cudaStream_t st);

// PE = U + V_bias (+1000 z^2), KE = m/2 |v + dt/2 F/m|^2 (ves/langevin_dynamics.py:227-228)
template <typename T>
cudaError_t launch_energies(const T* pos, const T* vel, int64_t n, int dims, T* PE, T* KE, T mass, T dt,
                            const PotParams<T>& PP, const BiasAnyParams<T>& BP, cudaStream_t st);

// ves/config_creation.py:56-68 per walker; fail_count (device int64, zeroed by the launcher) counts
// walkers that were never accepted.
template <typename T>
cudaError_t launch_init_positions(T* pos, int64_t n, int dims, uint64_t gid0, const double* box, double beta,
                                  int ntrials, const PotParams<double>& PD, const RngKeys& K,
                                  unsigned long long* fail_count, cudaStream_t st);

// ves/langevin_dynamics.py:70-74 setVelocitiesToTemperature
template <typename T>
cudaError_t launch_init_velocities(T* vel, int64_t n, int dims, uint64_t gid0, T sigma_v, const RngKeys& K,
                                   cudaStream_t st);

// ves/visualization.py:475 (np.histogram2d semantics: last bin closed on the right)
template <typename T>
cudaError_t launch_histogram(const T* pos, int64_t n, int64_t stride_row, int64_t stride_col, int ndim, int axis,
                             const double* range, const int* nbins, const T* row_weights, double* counts,
                             cudaStream_t st);

cudaError_t fp32_fma_peak(double* tflops, int packed);

}  // namespace pv
