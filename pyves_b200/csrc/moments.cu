#include "moments.cuh"

namespace pv {

constexpr int kMomThreads = 256;
constexpr int kMaxBlocks = 148 * 4;
constexpr int kPairTile = 64;       // covariance: pairs handled per block = 64 x 64
constexpr int kKpt = 16;            // first moments: basis functions per thread (K <= 4096)

template <typename T> struct Tile {
  static constexpr int kRows1 = 1024 / sizeof(T);   // rows per tile, first moments (256 / 128)
  static constexpr int kRows2 = 256 / sizeof(T);    // rows per tile, covariance (64 / 32)
};

// ------------------------------------------------------------------------------------------------
// Per-row factor tables: FA_a into fa[a] (a < ka), FB_b into fb[b] (b < kb), row weight into wt.
template <typename T>
PV_DEV void row_factors(const RowSource<T>& S, const BasisParams<T>& B, int64_t r, T* fa, T* fb, T& wt) {
  const T x = S.pos[r * S.stride_row];
  const T y = S.ncols > 1 ? S.pos[r * S.stride_row + S.stride_col] : T(0);
  wt = S.w != nullptr ? S.w[r] : T(1);
  T sa, sb = T(0);
  bool inside, inside_b = true;
  if (B.kind == PYVES_BIAS_LEGENDRE_2D) {
    sa = scale_clamp(x, B.ca, B.iha, inside);
    sb = scale_clamp(y, B.cb, B.ihb, inside_b);
  } else if (B.kind == PYVES_BIAS_LEGENDRE_1D) {
    sa = scale_clamp(B.axis ? y : x, B.ca, B.iha, inside);
  } else {   // path CV, ves/basis.py:524-535
    T amax = T(-1e30);
    for (int i = 0; i < B.npath; ++i) {
      const T dx = x - B.path[i], dy = y - B.path[B.npath + i];
      amax = Math<T>::max(amax, -B.lam * (dx * dx + dy * dy));
    }
    T S0 = 0, S1 = 0;
    for (int i = 0; i < B.npath; ++i) {
      const T dx = x - B.path[i], dy = y - B.path[B.npath + i];
      const T ee = Math<T>::exp(-B.lam * (dx * dx + dy * dy) - amax);
      S0 += ee;
      S1 += ee * T(i + 1);
    }
    sa = scale_clamp(S1 / (S0 * T(B.npath)), T(0.5), T(2), inside);
  }
  // conditional ensemble averages <f>_{V | s in the box}: the estimator of the VES functional restricted
  // to the support of the target (walkers that left the interval would otherwise be lumped onto its edge)
  if (B.inside_only && !(inside && inside_b)) wt = T(0);
  T pm = T(1), p = sa;
  fa[0] = T(1);
  if (B.ka > 1) fa[1] = sa;
  for (int n = 1; n + 1 < B.ka; ++n) {
    const T fn = T(n);
    const T pn = ((T(2) * fn + T(1)) * sa * p - fn * pm) / (fn + T(1));
    fa[n + 1] = pn;
    pm = p;
    p = pn;
  }
  if (B.kb > 1) {
    pm = T(1);
    p = sb;
    fb[0] = T(1);
    fb[1] = sb;
    for (int n = 1; n + 1 < B.kb; ++n) {
      const T fn = T(n);
      const T pn = ((T(2) * fn + T(1)) * sb * p - fn * pm) / (fn + T(1));
      fb[n + 1] = pn;
      pm = p;
      p = pn;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// First moments.  part_f[block][K], part_w[block].
template <typename T>
__global__ void __launch_bounds__(kMomThreads) moments1_kernel(const RowSource<T> S, const BasisParams<T> B,
                                                               double* __restrict__ part_f,
                                                               double* __restrict__ part_w) {
  constexpr int R = Tile<T>::kRows1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* fa = reinterpret_cast<T*>(smem_raw);        // [R][kap]
  const int kap = B.ka | 1, kbp = B.kb | 1;
  T* fb = fa + R * kap;                          // [R][kbp]
  __shared__ double wred[kMomThreads / 32];
  const int K = B.ka * B.kb;
  double acc[kKpt];
#pragma unroll
  for (int m = 0; m < kKpt; ++m) acc[m] = 0.0;
  double wsum = 0.0;
  const int64_t ntiles = (S.n + R - 1) / R;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * R;
    const int rows = (int)((S.n - r0) < R ? (S.n - r0) : R);
    if ((int)threadIdx.x < R) {
      T* ra = fa + threadIdx.x * kap;
      T* rb = fb + threadIdx.x * kbp;
      if ((int)threadIdx.x < rows) {
        T wt;
        row_factors(S, B, r0 + threadIdx.x, ra, rb, wt);
        for (int a = 0; a < B.ka; ++a) ra[a] *= wt;   // fold the row weight into factor A
        wsum += (double)wt;
      } else {
        for (int a = 0; a < B.ka; ++a) ra[a] = T(0);
        for (int b = 0; b < B.kb; ++b) rb[b] = T(0);
      }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < kKpt; ++m) {
      const int k = threadIdx.x + m * kMomThreads;
      if (k < K) {
        const int a = k / B.kb, b = k - a * B.kb;
        T s0 = T(0), s1 = T(0);
        if (B.kb > 1) {
          for (int r = 0; r < R; r += 2) {
            s0 = Math<T>::fma(fa[r * kap + a], fb[r * kbp + b], s0);
            s1 = Math<T>::fma(fa[(r + 1) * kap + a], fb[(r + 1) * kbp + b], s1);
          }
        } else {
          for (int r = 0; r < R; r += 2) {
            s0 += fa[r * kap + a];
            s1 += fa[(r + 1) * kap + a];
          }
        }
        acc[m] += (double)(s0 + s1);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < kKpt; ++m) {
    const int k = threadIdx.x + m * kMomThreads;
    if (k < K) part_f[(int64_t)blockIdx.x * K + k] = acc[m];
  }
  for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < kMomThreads / 32; ++i) s += wred[i];
    part_w[blockIdx.x] = s;
  }
}

// out[m] = sum_g part[g][m] (m < M) and out_w[0] = sum_g part_w[g], in a fixed order (deterministic):
// one warp per column, lane l adds the partials g = l, l + 32, ... (independent loads, all in flight
// together), then a butterfly over the lanes.  A single thread walking the G = 592 partials of a column
// was a chain of 592 dependent L2 round trips (~100 us per call, as long as the moment kernel itself).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ part, int64_t G, int64_t M,
                                                              double* __restrict__ out, const double* __restrict__ part_w,
                                                              double* __restrict__ out_w) {
  const int64_t col = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (col > M || (col == M && part_w == nullptr)) return;
  const double* src = col < M ? part + col : part_w;
  const int64_t stride = col < M ? M : 1;
  double s = 0.0;
  for (int64_t g = lane; g < G; g += 32) s += src[g * stride];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (col < M) out[col] = s;
    else out_w[0] = s;
  }
}

static inline void launch_reduce_partials(const double* part, int64_t G, int64_t M, double* out, const double* part_w,
                                          double* out_w, cudaStream_t st) {
  const int64_t cols = M + (part_w != nullptr ? 1 : 0);
  reduce_partials_kernel<<<(unsigned)((cols + 7) / 8), 256, 0, st>>>(part, G, M, out, part_w, out_w);
}

// ------------------------------------------------------------------------------------------------
// Second moments sum_w f_k f_l.  grid = (row chunks, upper-triangular 64x64 pair tiles).
// Each thread owns a 4 x 4 register tile; the f values of 64 rows (32 in FP64) are staged in
// shared memory as fA[r][64] (weight folded in) and fB[r][64], read with 128-bit loads.
template <typename T>
__global__ void __launch_bounds__(kMomThreads) moments2_kernel(const RowSource<T> S, const BasisParams<T> B,
                                                               int K64, double* __restrict__ part) {
  constexpr int R = Tile<T>::kRows2;
  constexpr int V = 16 / sizeof(T);
  struct alignas(16) Vec { T v[V]; };
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* fA = reinterpret_cast<T*>(smem_raw);   // [R][64]
  T* fB = fA + R * kPairTile;               // [R][64]
  const int kap = B.ka | 1, kbp = B.kb | 1;
  T* ta = fB + R * kPairTile;               // [R][kap]  FA
  T* tb = ta + R * kap;                     // [R][kbp]  FB
  T* wrow = tb + R * kbp;                   // [R]       row weight (0 for dead rows)
  // pair tile (ti <= tj) from the linear index
  int ti = 0, rem = blockIdx.y;
  while (rem >= K64 - ti) {
    rem -= K64 - ti;
    ++ti;
  }
  const int tj = ti + rem;
  const int kA0 = ti * kPairTile, kB0 = tj * kPairTile;
  const int K = B.ka * B.kb;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  T acc[4][4];
  double acc64[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j] = T(0);
      acc64[i][j] = 0.0;
    }
  const int64_t ntiles = (S.n + R - 1) / R;
  int since_flush = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * R;
    const int rows = (int)((S.n - r0) < R ? (S.n - r0) : R);
    if ((int)threadIdx.x < R) {
      T* ra = ta + threadIdx.x * kap;
      T* rb = tb + threadIdx.x * kbp;
      if ((int)threadIdx.x < rows) {
        T wt;
        row_factors(S, B, r0 + threadIdx.x, ra, rb, wt);
        wrow[threadIdx.x] = wt;
      } else {
        for (int a = 0; a < B.ka; ++a) ra[a] = T(0);
        for (int b = 0; b < B.kb; ++b) rb[b] = T(0);
        wrow[threadIdx.x] = T(0);
      }
    }
    __syncthreads();
    // fA[r][kk] = w_r f_{kA0+kk}(r);  fB[r][ll] = f_{kB0+ll}(r);  f_k = FA_a FB_b
    for (int e = threadIdx.x; e < R * kPairTile; e += kMomThreads) {
      const int r = e / kPairTile, kk = e - r * kPairTile;
      const int ka_ = kA0 + kk, kb_ = kB0 + kk;
      T va = T(0), vb = T(0);
      if (ka_ < K) {
        const int a = ka_ / B.kb, b = ka_ - a * B.kb;
        va = wrow[r] * ta[r * kap + a] * (B.kb > 1 ? tb[r * kbp + b] : T(1));
      }
      if (kb_ < K) {
        const int a = kb_ / B.kb, b = kb_ - a * B.kb;
        vb = ta[r * kap + a] * (B.kb > 1 ? tb[r * kbp + b] : T(1));
      }
      fA[e] = va;
      fB[e] = vb;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      T a4[4], b4[4];
      if (V == 4) {
        const Vec av = *reinterpret_cast<const Vec*>(fA + r * kPairTile + ty * 4);
        const Vec bv = *reinterpret_cast<const Vec*>(fB + r * kPairTile + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a4[i] = av.v[i % V];
          b4[i] = bv.v[i % V];
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const Vec av = *reinterpret_cast<const Vec*>(fA + r * kPairTile + ty * 4 + h * 2);
          const Vec bv = *reinterpret_cast<const Vec*>(fB + r * kPairTile + tx * 4 + h * 2);
          a4[h * 2] = av.v[0];
          a4[h * 2 + 1] = av.v[1 % V];
          b4[h * 2] = bv.v[0];
          b4[h * 2 + 1] = bv.v[1 % V];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = Math<T>::fma(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
    if (++since_flush == 8) {
      since_flush = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc64[i][j] += (double)acc[i][j];
          acc[i][j] = T(0);
        }
    }
  }
  double* out = part + ((int64_t)blockIdx.x * gridDim.y + blockIdx.y) * (kPairTile * kPairTile);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      out[(ty * 4 + i) * kPairTile + tx * 4 + j] = acc64[i][j] + (double)acc[i][j];
}

// sum_wff[k][l] (and mirror) = sum over row chunks of the pair-tile partials
__global__ void reduce_cov_kernel(const double* __restrict__ part, int G, int npt, int K64, int K,
                                  double* __restrict__ out) {
  const int pt = blockIdx.y;
  int ti = 0, rem = pt;
  while (rem >= K64 - ti) {
    rem -= K64 - ti;
    ++ti;
  }
  const int tj = ti + rem;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kPairTile * kPairTile) return;
  const int k = ti * kPairTile + e / kPairTile, l = tj * kPairTile + e % kPairTile;
  if (k >= K || l >= K) return;
  double s = 0.0;
  for (int g = 0; g < G; ++g) s += part[((int64_t)g * npt + pt) * (kPairTile * kPairTile) + e];
  if (ti == tj) {
    if (l >= k) {   // use the upper triangle of a diagonal tile for both halves: exact symmetry
      out[(int64_t)k * K + l] = s;
      out[(int64_t)l * K + k] = s;
    }
  } else {
    out[(int64_t)k * K + l] = s;
    out[(int64_t)l * K + k] = s;
  }
}

static int cov_chunks(int64_t n, int rows_per_tile, int npt) {
  int64_t tiles = (n + rows_per_tile - 1) / rows_per_tile;
  int64_t g = kMaxBlocks / npt;
  if (g < 1) g = 1;
  if (g > tiles) g = tiles;
  if (g < 1) g = 1;
  return (int)g;
}

size_t moments_scratch_doubles(int64_t K, bool cov) {
  size_t first = (size_t)kMaxBlocks * (size_t)(K + 1);
  if (!cov) return first;
  const int K64 = (int)((K + kPairTile - 1) / kPairTile);
  const int npt = K64 * (K64 + 1) / 2;
  int g = kMaxBlocks / npt;
  if (g < 1) g = 1;
  size_t total = first + (size_t)g * npt * kPairTile * kPairTile;
  if (K > 128 && K <= 256) {   // room for the tensor-core path's per-CTA FP32 partials
    const size_t tc = first + (cov_tc_scratch_floats(kMaxSmForScratch) + 1) / 2;
    if (tc > total) total = tc;
  }
  return total;
}

template <typename T> struct TcDispatch {
  static bool use(const RowSource<T>&, const BasisParams<T>&, int) { return false; }
  static cudaError_t run(const RowSource<T>&, const BasisParams<T>&, double*, double*, int, cudaStream_t, int*) { return cudaErrorNotSupported; }
};
template <> struct TcDispatch<float> {
  static bool use(const RowSource<float>& S, const BasisParams<float>& B, int n_sm) {
    return n_sm <= kMaxSmForScratch && cov_tc_supported(S, B);
  }
  static cudaError_t run(const RowSource<float>& S, const BasisParams<float>& B, double* sum_wff, double* scratch, int n_sm,
                         cudaStream_t st, int* launches) {
    return launch_cov_tc(S, B, sum_wff, reinterpret_cast<float*>(scratch), n_sm, st, launches);
  }
};

static bool g_no_tc = false;
bool no_tensor_cores() { return g_no_tc; }
void set_no_tensor_cores(bool v) { g_no_tc = v; }

template <typename T>
cudaError_t launch_moments_basis(const RowSource<T>& S, const BasisParams<T>& B, double* sum_w, double* sum_wf,
                                 double* sum_wff, double* scratch, cudaStream_t st, int* launches) {
  const int K = B.ka * B.kb;
  cudaError_t e;
  if (S.n == 0) {
    cudaMemsetAsync(sum_w, 0, sizeof(double), st);
    cudaMemsetAsync(sum_wf, 0, sizeof(double) * K, st);
    if (sum_wff) cudaMemsetAsync(sum_wff, 0, sizeof(double) * K * K, st);
    return cudaGetLastError();
  }
  {
    constexpr int R = Tile<T>::kRows1;
    const int64_t tiles = (S.n + R - 1) / R;
    const int G = (int)(tiles < kMaxBlocks ? tiles : kMaxBlocks);
    double* part_f = scratch;
    double* part_w = scratch + (size_t)kMaxBlocks * K;
    const size_t sm = (size_t)R * ((B.ka | 1) + (B.kb | 1)) * sizeof(T);
    auto kern = moments1_kernel<T>;
    if (sm > 48 * 1024) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e != cudaSuccess) return e;
    }
    kern<<<G, kMomThreads, sm, st>>>(S, B, part_f, part_w);
    launch_reduce_partials(part_f, G, K, sum_wf, part_w, sum_w, st);
    *launches += 2;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (sum_wff != nullptr && !no_tensor_cores() && TcDispatch<T>::use(S, B, n_sm)) {
    e = TcDispatch<T>::run(S, B, sum_wff, scratch + (size_t)kMaxBlocks * (K + 1), n_sm, st, launches);
    if (e != cudaSuccess) return e;
  } else if (sum_wff != nullptr) {
    constexpr int R = Tile<T>::kRows2;
    const int K64 = (K + kPairTile - 1) / kPairTile;
    const int npt = K64 * (K64 + 1) / 2;
    const int G = cov_chunks(S.n, R, npt);
    double* part = scratch + (size_t)kMaxBlocks * (K + 1);
    const size_t sm = (size_t)R * (2 * kPairTile + (B.ka | 1) + (B.kb | 1) + 1) * sizeof(T);
    auto kern = moments2_kernel<T>;
    if (sm > 48 * 1024) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e != cudaSuccess) return e;
    }
    kern<<<dim3(G, npt), kMomThreads, sm, st>>>(S, B, K64, part);
    reduce_cov_kernel<<<dim3(kPairTile * kPairTile / 256, npt), 256, 0, st>>>(part, G, npt, K64, K, sum_wff);
    *launches += 2;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// MLP parameter-gradient moments: sum_r w_r dV(s_r)/dtheta for NNBasis1D (ves/basis.py:184-232),
// i.e. what autograd accumulates over the loops of ves/bias.py:219-233.
// Phase 1: one thread per row runs forward + backward with activations and deltas in shared
// memory ([unit][row] so lanes hit consecutive banks).  Phase 2: all threads accumulate the outer
// products delta_l[j] h_{l-1}[i] over the tile's rows into the block's private partial vector.
template <typename T> struct MlpTile { static constexpr int kRows = 128 / sizeof(T); };   // 32 / 16

template <typename T>
__global__ void __launch_bounds__(kMomThreads) moments_mlp_kernel(const RowSource<T> S, const MlpGradParams<T> M,
                                                                  int total_units, double* __restrict__ part,
                                                                  double* __restrict__ part_w) {
  constexpr int R = MlpTile<T>::kRows;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* h = reinterpret_cast<T*>(smem_raw);   // inputs of each Linear: [unit offset][R]
  T* d = h + (size_t)total_units * R;      // deltas at the outputs of each Linear: [unit offset][R]
  __shared__ double wred;
  if (threadIdx.x == 0) wred = 0.0;
  double* mypart = part + (int64_t)blockIdx.x * M.n_params;
  const int64_t ntiles = (S.n + R - 1) / R;
  const int L = M.n_layers;
  double wsum = 0.0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * R;
    const int rows = (int)((S.n - r0) < R ? (S.n - r0) : R);
    if ((int)threadIdx.x < R) {
      const int r = threadIdx.x;
      if (r < rows) {
        const int64_t gr = r0 + r;
        const T q = (M.axis && S.ncols > 1) ? S.pos[gr * S.stride_row + S.stride_col] : S.pos[gr * S.stride_row];
        const T wt = S.w != nullptr ? S.w[gr] : T(1);
        wsum += (double)wt;
        // forward.  h block l (offset hoff) = inputs of Linear l; d block l (offset doff) = its
        // outputs: first the activation gates are stashed there, then the deltas overwrite them.
        h[r] = (q - M.lo) * M.inv_range;
        const T* th = M.theta;
        int hoff = 0, doff = 0;
        for (int l = 0; l < L; ++l) {
          const int nin = M.sizes[l], nout = M.sizes[l + 1];
          const T* W = th;
          const T* b = th + nin * nout;
          th = b + nout;
          const bool activate = (l >= 1) && (l < L - 1);   // ves/basis.py:197-205
          for (int j = 0; j < nout; ++j) {
            T z = b[j];
            for (int i = 0; i < nin; ++i) z = Math<T>::fma(W[j * nin + i], h[(hoff + i) * R + r], z);
            T gate = T(1);
            if (activate) z = M.act == PYVES_ACT_RELU ? act_fn<T, PYVES_ACT_RELU>(z, gate) : act_fn<T, PYVES_ACT_TANH>(z, gate);
            if (l < L - 1) h[(hoff + nin + j) * R + r] = z;
            d[(doff + j) * R + r] = gate;
          }
          hoff += nin;
          doff += nout;
        }
        // backward: delta of the output unit is the row weight; delta_{l-1} = (W_l^T delta_l) * gate_{l-1}
        doff -= 1;                       // offset of the last Linear's (single) output
        d[doff * R + r] = wt;
        for (int l = L - 1; l >= 1; --l) {
          const int nin = M.sizes[l], nout = M.sizes[l + 1];
          th -= nin * nout + nout;
          const T* W = th;
          const int dprev = doff - nin;
          for (int i = 0; i < nin; ++i) {
            T sacc = T(0);
            for (int j = 0; j < nout; ++j) sacc = Math<T>::fma(W[j * nin + i], d[(doff + j) * R + r], sacc);
            d[(dprev + i) * R + r] *= sacc;
          }
          doff = dprev;
        }
      } else {
        // dead rows of the last tile: zero deltas so phase 2 adds nothing
        int nd = 0;
        for (int l = 0; l < L; ++l) nd += M.sizes[l + 1];
        for (int u = 0; u < nd; ++u) d[u * R + r] = T(0);
        for (int u = 0; u < total_units; ++u) h[u * R + r] = T(0);
      }
    }
    __syncthreads();
    // phase 2
    {
      int poff = 0, hoff = 0, doff = 0;
      for (int l = 0; l < L; ++l) {
        const int nin = M.sizes[l], nout = M.sizes[l + 1];
        for (int p = threadIdx.x; p < nin * nout + nout; p += kMomThreads) {
          T s = T(0);
          if (p < nin * nout) {
            const int j = p / nin, i = p - j * nin;
            for (int r = 0; r < R; ++r) s = Math<T>::fma(d[(doff + j) * R + r], h[(hoff + i) * R + r], s);
          } else {
            const int j = p - nin * nout;
            for (int r = 0; r < R; ++r) s += d[(doff + j) * R + r];
          }
          mypart[poff + p] += (double)s;
        }
        poff += nin * nout + nout;
        hoff += nin;
        doff += nout;
      }
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  if ((threadIdx.x & 31) == 0 && wsum != 0.0) atomicAdd(&wred, wsum);
  __syncthreads();
  if (threadIdx.x == 0) part_w[blockIdx.x] = wred;
}

template <typename T>
cudaError_t launch_moments_mlp(const RowSource<T>& S, const MlpGradParams<T>& M, double* sum_w, double* sum_wf,
                               double* scratch, cudaStream_t st, int* launches) {
  constexpr int R = MlpTile<T>::kRows;
  const int P = M.n_params;
  if (S.n == 0) {
    cudaMemsetAsync(sum_w, 0, sizeof(double), st);
    cudaMemsetAsync(sum_wf, 0, sizeof(double) * P, st);
    return cudaGetLastError();
  }
  int units_in = 0, units_out = 0;
  for (int l = 0; l < M.n_layers; ++l) {
    units_in += M.sizes[l];
    units_out += M.sizes[l + 1];
  }
  const int64_t tiles = (S.n + R - 1) / R;
  const int G = (int)(tiles < kMaxBlocks ? tiles : kMaxBlocks);
  double* part = scratch;
  double* part_w = scratch + (size_t)kMaxBlocks * P;
  cudaError_t e = cudaMemsetAsync(part, 0, sizeof(double) * (size_t)G * P, st);
  if (e != cudaSuccess) return e;
  const size_t sm = (size_t)R * (units_in + units_out) * sizeof(T);
  auto kern = moments_mlp_kernel<T>;
  if (sm > 48 * 1024) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  kern<<<G, kMomThreads, sm, st>>>(S, M, units_in, part, part_w);
  launch_reduce_partials(part, G, P, sum_wf, part_w, sum_w, st);
  *launches += 2;
  return cudaGetLastError();
}

#define PV_INST(T)                                                                                               \
  template cudaError_t launch_moments_basis<T>(const RowSource<T>&, const BasisParams<T>&, double*, double*,     \
                                               double*, double*, cudaStream_t, int*);                            \
  template cudaError_t launch_moments_mlp<T>(const RowSource<T>&, const MlpGradParams<T>&, double*, double*,     \
                                             double*, cudaStream_t, int*);
PV_INST(float)
PV_INST(double)

}  // namespace pv
