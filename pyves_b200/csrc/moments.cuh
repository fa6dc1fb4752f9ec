// Ensemble moments of the basis functions: sum_w, sum_w f_k, sum_w f_k f_l.
// Replaces the two Python loops of VESBias_SingleParticle_1D.update (ves/bias.py:219-233:
// one module call per trajectory frame and per target node, then autograd) and provides the
// ensemble averages <f_k>_V, <f_k f_l>_V of textbook VES.
#pragma once
#include "bias.cuh"

namespace pv {

// rows: element (r, c) at pos[r * stride_row + c * stride_col]
template <typename T> struct RowSource {
  const T* pos;
  int64_t n;
  int64_t stride_row, stride_col;
  int ncols;
  const T* w;   // NULL: all ones
};

// Product bases f_k = FA_a(sa) FB_b(sb), k = a * kb + b.  1D kinds have kb = 1.
template <typename T> struct BasisParams {
  int kind;            // PYVES_BIAS_LEGENDRE_1D | _2D | _PATHCV
  int ka, kb;
  int axis;            // 1D: which column feeds factor A
  T ca, iha, cb, ihb;  // centre / inverse half width per factor
  int npath;           // path CV
  T lam;
  const T* path;       // device x_i[npath], y_i[npath]
  int inside_only;     // 1: rows outside the basis interval (where the scaled coordinate clamps) get weight 0
};

template <typename T> struct MlpGradParams {
  int axis, act, n_layers;
  int sizes[10];       // widths incl. input 1 and output 1
  T lo, inv_range;
  const T* theta;      // device, torch parameter order
  int n_params;
  int inside_only;     // 1: rows whose scaled coordinate leaves [0, 1] (the support of the target) get weight 0
};

// scratch must hold moments_scratch_doubles(K, cov, n, weighted) doubles (n rows, with / without row weights)
size_t moments_scratch_doubles(int64_t K, bool cov, int64_t n, bool weighted);

template <typename T>
cudaError_t launch_moments_basis(const RowSource<T>& S, const BasisParams<T>& B, double* sum_w, double* sum_wf,
                                 double* sum_wff, double* scratch, cudaStream_t st, int* launches);

// tcgen05 covariance (cov_tc.cu, FP16 operand tiles = the 10-bit mantissa of TF32 on |f| <= 1, FP32 accumulate): used
// for FP32 handles and 2D Legendre bases with 64 < K <= 4096, with or without row weights; everything else runs
// moments_small_kernel / moments2_kernel on the CUDA cores.
bool cov_tc_supported(const RowSource<float>& S, const BasisParams<float>& B);
size_t cov_tc_scratch_floats(int K, int64_t n, int n_sm, bool weighted);
cudaError_t launch_cov_tc(const RowSource<float>& S, const BasisParams<float>& B, double* sum_wff, float* scratch, int n_sm,
                          cudaStream_t st, int* launches);
void cov_tc_set_max_rows(int64_t rows);   // rows one CTA accumulates in FP32 before its partial goes to the FP64 sum (tuning)
constexpr int kMaxSmForScratch = 160;
bool no_tensor_cores();
// ReLU NNBasis1D: every parameter gradient dV/dtheta is AFFINE in the input on each piece of the piecewise-linear V (bias.cuh,
// BiasPwl1D), so sum_r w_r dV/dtheta(s_r) = sum over the pieces p of n_p dV/dtheta(mean_p): the rows are binned into the pieces
// (n_p = sum of weights, S_p = sum of weights times s), and the parameter-gradient kernel runs on cap + 1 pseudo rows instead
// of n.  hist: 2 (cap + 1) doubles; rows: 2 (cap + 1) T (two columns, both the coordinate); row_w: cap + 1 T.  Stream ordered.
template <typename T>
cudaError_t launch_pwl_compress(const RowSource<T>& S, int axis, T lo, T inv_range, int inside_only, const T* brk, const int* count, int cap,
                                double* hist, T* rows, T* row_w, cudaStream_t st, int* launches);
void set_tiled_first_moments(bool v);   // tuning / test switch: register-tiled first moments of large 2D bases (default on)
void set_no_tensor_cores(bool v);   // tuning / test switch: force the CUDA-core covariance

template <typename T>
cudaError_t launch_moments_mlp(const RowSource<T>& S, const MlpGradParams<T>& M, double* sum_w, double* sum_wf,
                               double* scratch, cudaStream_t st, int* launches);

}  // namespace pv
