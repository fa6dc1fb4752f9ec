#include <cstring>

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
    // nearest image first, then exponents from the difference of squares (see BiasOther, bias.cuh)
    T dmin = T(1e30), xm = B.path[0], ym = B.path[B.npath];
    for (int i = 0; i < B.npath; ++i) {
      const T dx = x - B.path[i], dy = y - B.path[B.npath + i];
      const T d2 = Math<T>::fma(dx, dx, dy * dy);
      if (d2 < dmin) { dmin = d2; xm = B.path[i]; ym = B.path[B.npath + i]; }
    }
    const T tx = (x - xm) + x, ty = (y - ym) + y;
    T S0 = 0, S1 = 0;
    for (int i = 0; i < B.npath; ++i) {
      const T xi = B.path[i], yi = B.path[B.npath + i];
      const T dd = Math<T>::fma(xm - xi, tx - xi, (ym - yi) * (ty - yi));
      const T ee = Math<T>::exp(-B.lam * dd);
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

// First moments of a LARGE 2D product basis (K > 256), register tiled.  moments1_kernel gives every thread one basis function
// per pass and pays two shared-memory loads per FMA (K = 4096: 1.25 ms per 10^6 rows, LSU bound).  Here
//   phase 1: ALL threads build the factor tables of a 128-row tile -- thread r < 128 the A (x) factors of row r, row weight
//            folded in, thread 128 + r the B (y) factors -- four values per 128-bit store (row stride = k + 4 words: conflict free);
//   phase 2: a thread owns a 4 x 4 tile of functions (a in 4 ta .., b in 4 tb ..) and, when the basis has fewer than 256 such
//            tiles, one of RG row groups: per row two 128-bit loads feed 16 FMAs;
// partial sums leave the tile loop in FP64 (one add per accumulator and tile), the row groups are summed through shared memory
// at the end.  part_f[block][K], part_w[block] as for moments1_kernel.
constexpr int kTileRows = 128;
template <typename T>
__global__ void __launch_bounds__(kMomThreads) moments1_tiled_kernel(const RowSource<T> S, const BasisParams<T> B,
                                                                     double* __restrict__ part_f, double* __restrict__ part_w) {
  constexpr int R = kTileRows;
  constexpr int V = 16 / sizeof(T);
  struct alignas(16) Vec { T v[V]; };
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int kap = ((B.ka + 3) & ~3) + 4, kbp = ((B.kb + 3) & ~3) + 4;
  T* fa = reinterpret_cast<T*>(smem_raw);        // [R][kap]
  T* fb = fa + R * kap;                          // [R][kbp]
  __shared__ double wred[kMomThreads / 32];
  __shared__ T rec_a[64], rec_b[64];                 // P_{n+1} = a_n s P_n - b_n P_{n-1}
  if (threadIdx.x < 64) {
    rec_a[threadIdx.x] = T(2 * threadIdx.x + 1) / T(threadIdx.x + 1);
    rec_b[threadIdx.x] = T(threadIdx.x) / T(threadIdx.x + 1);
  }
  __syncthreads();
  const int ka4 = (B.ka + 3) >> 2, kb4 = (B.kb + 3) >> 2;
  const int nt = ka4 * kb4, RG = kMomThreads / nt;
  const int tile_id = threadIdx.x % nt, rg = threadIdx.x / nt;
  const bool active = rg < RG;
  const int ta = tile_id / kb4, tb = tile_id - ta * kb4;
  double acc64[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc64[i][j] = 0.0;
  double wsum = 0.0;
  const int64_t ntiles = (S.n + R - 1) / R;
  const int row = threadIdx.x & (R - 1), side = threadIdx.x >> 7;          // phase 1 role (R = 128, 256 threads)
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r = tile * R + row;
    {
      const int kf = side ? B.kb : B.ka, kfp = side ? kbp : kap;
      T* dst = (side ? fb : fa) + row * kfp;
      T wt = T(0), s = T(0);
      if (r < S.n) {
        const T x = S.pos[r * S.stride_row], y = S.pos[r * S.stride_row + S.stride_col];
        bool inx, iny;
        const T sx = scale_clamp(x, B.ca, B.iha, inx), sy = scale_clamp(y, B.cb, B.ihb, iny);
        wt = S.w != nullptr ? S.w[r] : T(1);
        if (B.inside_only && !(inx && iny)) wt = T(0);
        s = side ? sy : sx;
        if (side == 0) wsum += (double)wt;
      }
      const T scale = side ? (r < S.n ? T(1) : T(0)) : wt;                 // the row weight rides on factor A; rows past the end are zero
      T pm = T(1), p = s;
      for (int n0 = 0; n0 < kf; n0 += 4) {
        Vec out[4 / V];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = n0 + q;
          T val;
          if (n == 0) {
            val = T(1);
          } else if (n == 1) {
            val = s;
          } else {
            T pn;
            if (sizeof(T) == 8) {
              const T fn = T(n - 1);
              pn = ((T(2) * fn + T(1)) * s * p - fn * pm) / (fn + T(1));         // the arithmetic of row_factors (FP64 parity)
            } else {
              pn = rec_a[n - 1] * s * p - rec_b[n - 1] * pm;                     // FP32: tabulated coefficients, no division
            }
            pm = p;
            p = pn;
            val = pn;
          }
          out[q / V].v[q % V] = n < kf ? val * scale : T(0);
        }
#pragma unroll
        for (int h = 0; h < 4 / V; ++h) *reinterpret_cast<Vec*>(dst + n0 + h * V) = out[h];
      }
    }
    __syncthreads();
    if (active) {
      T acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
#pragma unroll 4
      for (int rr = rg; rr < R; rr += RG) {
        T a4[4], b4[4];
#pragma unroll
        for (int h = 0; h < 4 / V; ++h) {
          const Vec av = *reinterpret_cast<const Vec*>(fa + rr * kap + ta * 4 + h * V);
          const Vec bv = *reinterpret_cast<const Vec*>(fb + rr * kbp + tb * 4 + h * V);
#pragma unroll
          for (int q = 0; q < V; ++q) {
            a4[h * V + q] = av.v[q];
            b4[h * V + q] = bv.v[q];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = Math<T>::fma(a4[i], b4[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc64[i][j] += (double)acc[i][j];
    }
    __syncthreads();
  }
  // row groups -> one vector per block, through shared memory: red[rg][tile][16]
  double* red = reinterpret_cast<double*>(smem_raw);
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[((size_t)rg * nt + tile_id) * 16 + i * 4 + j] = acc64[i][j];
  }
  __syncthreads();
  const int K = B.ka * B.kb;
  for (int e = threadIdx.x; e < nt * 16; e += kMomThreads) {
    const int tl = e >> 4, ij = e & 15;
    const int a = (tl / kb4) * 4 + (ij >> 2), b = (tl % kb4) * 4 + (ij & 3);
    if (a < B.ka && b < B.kb) {
      double sum = 0.0;
      for (int g = 0; g < RG; ++g) sum += red[((size_t)g * nt + tl) * 16 + ij];
      part_f[(int64_t)blockIdx.x * K + a * B.kb + b] = sum;
    }
  }
  for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sw = 0.0;
    for (int i = 0; i < kMomThreads / 32; ++i) sw += wred[i];
    part_w[blockIdx.x] = sw;
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
  // factor indices (a, b) of the 64 + 64 basis functions of this pair tile: no divisions in the row loop
  __shared__ short idx_a[2][kPairTile], idx_b[2][kPairTile];
  if (threadIdx.x < 2 * kPairTile) {
    const int side = threadIdx.x / kPairTile, kk = threadIdx.x % kPairTile;
    const int kf = (side ? kB0 : kA0) + kk;
    idx_a[side][kk] = (short)(kf < K ? kf / B.kb : 0);
    idx_b[side][kk] = (short)(kf < K ? kf % B.kb : 0);
  }
  __syncthreads();
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
      const int r = e / kPairTile, kk = e % kPairTile;
      T va = T(0), vb = T(0);
      if (kA0 + kk < K) va = wrow[r] * ta[r * kap + idx_a[0][kk]] * (B.kb > 1 ? tb[r * kbp + idx_b[0][kk]] : T(1));
      if (kB0 + kk < K) vb = ta[r * kap + idx_a[1][kk]] * (B.kb > 1 ? tb[r * kbp + idx_b[1][kk]] : T(1));
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

// sum_wff[k][l] (and mirror) = sum over row chunks of the pair-tile partials, fixed order: LPO lanes per
// output element split the G partials (independent loads in flight), then a butterfly over those lanes.
template <int LPO>
__global__ void __launch_bounds__(256) reduce_cov_kernel(const double* __restrict__ part, int G, int npt, int K64, int K,
                                                          double* __restrict__ out) {
  const int pt = blockIdx.y;
  int ti = 0, rem = pt;
  while (rem >= K64 - ti) {
    rem -= K64 - ti;
    ++ti;
  }
  const int tj = ti + rem;
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) / LPO;   // grid.x covers kPairTile^2 * LPO threads exactly
  const int sub = threadIdx.x % LPO;
  const int k = ti * kPairTile + e / kPairTile, l = tj * kPairTile + e % kPairTile;
  if (k >= K || l >= K) return;                                  // uniform over the LPO lanes of one element
  double s = 0.0;
  for (int g = sub; g < G; g += LPO) s += part[((int64_t)g * npt + pt) * (kPairTile * kPairTile) + e];
  const unsigned lane = threadIdx.x & 31;
  const unsigned mask = LPO == 32 ? 0xffffffffu : (((1u << LPO) - 1u) << (lane & ~(unsigned)(LPO - 1)));
#pragma unroll
  for (int o = LPO / 2; o > 0; o >>= 1) s += __shfl_xor_sync(mask, s, o);
  if (sub != 0) return;
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

static void launch_reduce_cov(const double* part, int G, int npt, int K64, int K, double* out, cudaStream_t st) {
  const int E = kPairTile * kPairTile;
  if (G >= 64) reduce_cov_kernel<32><<<dim3(E * 32 / 256, npt), 256, 0, st>>>(part, G, npt, K64, K, out);
  else if (G >= 8) reduce_cov_kernel<8><<<dim3(E * 8 / 256, npt), 256, 0, st>>>(part, G, npt, K64, K, out);
  else reduce_cov_kernel<1><<<dim3(E / 256, npt), 256, 0, st>>>(part, G, npt, K64, K, out);
}

// ------------------------------------------------------------------------------------------------
// Small bases (K <= 32: the 1D Legendre biases of configs[0] / configs[2], K = 6 / 13): first and second
// moments in one pass.  The 64 x 64 pair-tile kernel above pads such a basis 25-fold (0.6 ms per 10^6 rows,
// a quarter of the 500 MD steps of configs[2]).  Here every thread stages the K basis values of one row
// (f_0 = P_0 = 1), then thread t owns the pair p = t % NPP (k <= l) on the row slice t / NPP of the tile;
// per-tile FP32 sums are promoted to FP64, the slices are merged in shared memory in a fixed order.
constexpr int kSmallK = 32;
constexpr int kSmallRows = 256;

template <typename T>
__global__ void __launch_bounds__(kMomThreads) moments_small_kernel(const RowSource<T> S, const BasisParams<T> B, int NPP,
                                                                    double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = B.ka * B.kb, Kp = K | 1, NP = K * (K + 1) / 2;
  T* f = reinterpret_cast<T*>(smem_raw);            // [kSmallRows][Kp] basis values
  T* fw = f + kSmallRows * Kp;                      // [kSmallRows][Kp] weight * basis values
  const int tid = threadIdx.x;
  const int p = tid % NPP, slice = tid / NPP, nslices = kMomThreads / NPP;
  int pk = 0, pl = 0;
  {   // pair index -> (k, l), k <= l, row-major over the upper triangle
    int rem = p;
    while (pk < K && rem >= K - pk) {
      rem -= K - pk;
      ++pk;
    }
    pl = pk + rem;
  }
  const bool owner = p < NP && slice < nslices;
  double acc = 0.0;
  const int64_t ntiles = (S.n + kSmallRows - 1) / kSmallRows;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r = tile * kSmallRows + tid;
    {
      T fa[kSmallK], fb[kSmallK], wt = T(0);
      if (r < S.n) {
        row_factors(S, B, r, fa, fb, wt);
      } else {
        for (int a = 0; a < B.ka; ++a) fa[a] = T(0);
        for (int b = 0; b < B.kb; ++b) fb[b] = T(0);
      }
      for (int a = 0; a < B.ka; ++a)
        for (int b = 0; b < B.kb; ++b) {
          const T v = B.kb > 1 ? fa[a] * fb[b] : fa[a];
          f[tid * Kp + a * B.kb + b] = v;
          fw[tid * Kp + a * B.kb + b] = wt * v;
        }
    }
    __syncthreads();
    if (owner) {
      T s0 = T(0), s1 = T(0);
      const int rows_per = kSmallRows / nslices;
      const T* fk = fw + (slice * rows_per) * Kp + pk;
      const T* fl = f + (slice * rows_per) * Kp + pl;
      for (int q = 0; q < rows_per; q += 2) {
        s0 = Math<T>::fma(fk[q * Kp], fl[q * Kp], s0);
        s1 = Math<T>::fma(fk[(q + 1) * Kp], fl[(q + 1) * Kp], s1);
      }
      acc += (double)(s0 + s1);
    }
    __syncthreads();
  }
  double* red = reinterpret_cast<double*>(smem_raw);   // [nslices][NPP] (the row tables are free now)
  if (slice < nslices) red[slice * NPP + p] = acc;
  __syncthreads();
  if (tid < NP) {
    double t = 0.0;
    for (int q = 0; q < nslices; ++q) t += red[q * NPP + tid];
    part[(int64_t)blockIdx.x * NP + tid] = t;
  }
}

// sums the block partials (one warp per pair) and scatters: sum_wff (both triangles), sum_wf = row 0, sum_w = [0][0]
__global__ void __launch_bounds__(256) reduce_small_kernel(const double* __restrict__ part, int G, int K, double* __restrict__ sum_w,
                                                           double* __restrict__ sum_wf, double* __restrict__ sum_wff) {
  const int NP = K * (K + 1) / 2;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= NP) return;
  double s = 0.0;
  for (int g = lane; g < G; g += 32) s += part[(int64_t)g * NP + p];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane != 0) return;
  int k = 0, rem = p;
  while (rem >= K - k) {
    rem -= K - k;
    ++k;
  }
  const int l = k + rem;
  sum_wff[(int64_t)k * K + l] = s;
  sum_wff[(int64_t)l * K + k] = s;
  if (k == 0) {
    sum_wf[l] = s;
    if (l == 0) sum_w[0] = s;
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

size_t moments_scratch_doubles(int64_t K, bool cov, int64_t n, bool weighted) {
  size_t first = (size_t)kMaxBlocks * (size_t)(K + 1);
  if (!cov) return first;
  const int K64 = (int)((K + kPairTile - 1) / kPairTile);
  const int npt = K64 * (K64 + 1) / 2;
  int g = kMaxBlocks / npt;
  if (g < 1) g = 1;
  size_t total = first + (size_t)g * npt * kPairTile * kPairTile;
  if (K > 64 && K <= 4096) {   // room for the tensor-core path's FP32 partial tiles
    const size_t tc = first + (cov_tc_scratch_floats((int)K, n, kMaxSmForScratch, weighted) + 1) / 2;
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
static int g_tiled_min_k = 64;       // below: moments1_kernel (few 4 x 4 tiles per block)
static bool g_tiled_first = true;    // experiment knob (pyves_set_tuning 96 / 95): register-tiled first moments of large 2D bases
void set_tiled_first_moments(bool v) { g_tiled_first = v; }
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
  if (sum_wff != nullptr && K <= kSmallK) {   // one pass for first and second moments
    const int NP = K * (K + 1) / 2;
    int NPP = 32;
    while (NPP < NP && NPP < kMomThreads) NPP <<= 1;          // pairs padded to a power of two <= 256
    if (NP > kMomThreads) NPP = 0;
    if (NPP) {
      const int64_t tiles = (S.n + kSmallRows - 1) / kSmallRows;
      const int G = (int)(tiles < kMaxBlocks ? tiles : kMaxBlocks);
      const size_t sm = (size_t)2 * kSmallRows * (K | 1) * sizeof(T);
      auto kern = moments_small_kernel<T>;
      if (sm > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
      }
      kern<<<G, kMomThreads, sm, st>>>(S, B, NPP, scratch);
      reduce_small_kernel<<<(NP + 7) / 8, 256, 0, st>>>(scratch, G, K, sum_w, sum_wf, sum_wff);
      *launches += 2;
      return cudaGetLastError();
    }
  }
  if (B.kind == PYVES_BIAS_LEGENDRE_2D && B.kb > 1 && B.ka <= 64 && B.kb <= 64 && K >= g_tiled_min_k && K <= 4096 && g_tiled_first) {   // product bases: register-tiled
    constexpr int R = kTileRows;
    const int64_t tiles = (S.n + R - 1) / R;
    const int G = (int)(tiles < kMaxBlocks ? tiles : kMaxBlocks);
    double* part_f = scratch;
    double* part_w = scratch + (size_t)kMaxBlocks * K;
    const int kap = ((B.ka + 3) & ~3) + 4, kbp = ((B.kb + 3) & ~3) + 4;
    size_t sm = (size_t)R * (kap + kbp) * sizeof(T);
    const size_t red = (size_t)kMomThreads * 16 * sizeof(double);
    if (sm < red) sm = red;
    auto kern = moments1_tiled_kernel<T>;
    if (sm + 1024 > 48 * 1024) {   // static shared memory counts against the default limit too
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e != cudaSuccess) return e;
    }
    kern<<<G, kMomThreads, sm, st>>>(S, B, part_f, part_w);
    launch_reduce_partials(part_f, G, K, sum_wf, part_w, sum_w, st);
    *launches += 2;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  } else {
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
    launch_reduce_cov(part, G, npt, K64, K, sum_wff, st);
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
        T wt = S.w != nullptr ? S.w[gr] : T(1);
        const T sq = (q - M.lo) * M.inv_range;
        if (M.inside_only && !(sq >= T(0) && sq <= T(1))) wt = T(0);
        wsum += (double)wt;
        // forward.  h block l (offset hoff) = inputs of Linear l; d block l (offset doff) = its
        // outputs: first the activation gates are stashed there, then the deltas overwrite them.
        h[r] = sq;
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

// ------------------------------------------------------------------------------------------------
// Fast path of the MLP parameter-gradient moments (widths multiples of 4, <= 128).  The version above
// runs forward + backward on one thread per row (1 warp of 8 busy) and read-modify-writes every
// parameter partial in global memory once per 32 rows; measured on B200 for 10^6 rows: 17 ms for the
// 3 x 48 net and 107 ms for the reference example's [128, 128] (8 x the 500 MD steps between updates).
// Here a block keeps a tile of R rows in shared memory ([unit][R + 4]) and
//   phase 1: all 256 threads run each Linear as a small GEMM (4 units x 8 rows per thread, weights as
//            128-bit broadcast loads through L1, activations as 128-bit shared loads), forward and backward;
//   phase 2: the outer products delta_l h_{l-1}^T of ONE "big" layer (nin, nout > 1; blockIdx.y selects it)
//            are accumulated in a TS x TS register tile per thread (strided ownership: conflict-free
//            shared loads) over ALL tiles of the block and written once; biases and the 1 -> h / h -> 1
//            weights ("small" parameters) are handled by the blockIdx.y == 0 blocks.
struct MlpPlan {
  int L;                       // number of Linear layers
  int nin[10], nout[10];
  int hoff[10], doff[10];      // unit offsets of the inputs / outputs of Linear l
  int poff[10];                // parameter offset of W_l (bias follows at poff + nin * nout)
  int nbig, big[10];           // layers with nin > 1 and nout > 1
  int nsmall;                  // biases + weights of the layers with nin == 1 or nout == 1
  int units_in, units_out;
};

template <typename T> struct Mlp2 {
  static constexpr int kRows = 256 / sizeof(T);      // 64 / 32
  static constexpr int kStride = kRows + 4;          // row stride in shared memory (bank spread)
  static constexpr int V = 16 / sizeof(T);
  static constexpr int kRG = kRows / 8;              // row groups of 8
};
constexpr int kMlp2Blocks = 148 * 2;
constexpr int kMlp2Small = 4;                        // small parameters per thread (<= 1024 in total)

template <typename T> PV_DEV void ldg_vec(const T* p, T (&v)[16 / sizeof(T)]) {
  if (sizeof(T) == 4) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = (T)q.x; v[1] = (T)q.y; v[2] = (T)q.z; v[3] = (T)q.w;
  } else {
    const double2 q = __ldg(reinterpret_cast<const double2*>(p));
    v[0] = (T)q.x; v[1] = (T)q.y;
  }
}
template <typename T> PV_DEV void lds8(const T* p, T (&v)[8]) {
  constexpr int V = 16 / sizeof(T);
#pragma unroll
  for (int c = 0; c < 8 / V; ++c) {
    T t[V];
    lds_vec(p + c * V, t);
#pragma unroll
    for (int q = 0; q < V; ++q) v[c * V + q] = t[q];
  }
}

// NB big layers per block (blockIdx.y picks the group): with NB = 2 the 3 x 48 net of configs[3] runs its forward / backward
// pass ONCE per tile instead of once per big layer (2.7 -> 1.7 ms per 10^6 rows).
template <typename T, int TS, int NB>
__global__ void __launch_bounds__(kMomThreads) moments_mlp2_kernel(const RowSource<T> S, const MlpGradParams<T> M, const MlpPlan P,
                                                                   double* __restrict__ part, double* __restrict__ part_w) {
  using X = Mlp2<T>;
  constexpr int R = X::kRows, RS = X::kStride, RG = X::kRG;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* h = reinterpret_cast<T*>(smem_raw);        // inputs of each Linear: [unit][RS]
  T* d = h + (size_t)P.units_in * RS;           // gates, then deltas, at the outputs of each Linear: [unit][RS]
  T* wrow = d + (size_t)P.units_out * RS;       // [R] row weights
  const int tid = threadIdx.x;
  const int y = blockIdx.y;
  int bls[NB];
#pragma unroll
  for (int g = 0; g < NB; ++g) bls[g] = (y * NB + g) < P.nbig ? P.big[y * NB + g] : -1;
  const int tj = tid >> 4, ti = tid & 15;

  T acc[NB][TS][TS];
#pragma unroll
  for (int g = 0; g < NB; ++g)
#pragma unroll
    for (int a = 0; a < TS; ++a)
#pragma unroll
      for (int b = 0; b < TS; ++b) acc[g][a][b] = T(0);
  // small parameters of this thread (y == 0 only): gradient = sum_r d[srow_d][r] * (srow_h >= 0 ? h[srow_h][r] : 1)
  int srow_d[kMlp2Small], srow_h[kMlp2Small], sidx[kMlp2Small];
  T sacc[kMlp2Small];
#pragma unroll
  for (int q = 0; q < kMlp2Small; ++q) {
    sacc[q] = T(0);
    sidx[q] = -1;
    srow_d[q] = srow_h[q] = 0;
    int sp = tid + q * kMomThreads;
    if (y == 0 && sp < P.nsmall) {
      for (int l = 0; l < P.L; ++l) {
        const int nw = P.nin[l] * P.nout[l];
        const bool small_w = P.nin[l] == 1 || P.nout[l] == 1;
        if (small_w) {
          if (sp < nw) {
            const int j = sp / P.nin[l], i = sp - j * P.nin[l];
            srow_d[q] = P.doff[l] + j; srow_h[q] = P.hoff[l] + i; sidx[q] = P.poff[l] + sp;
            break;
          }
          sp -= nw;
        }
        if (sp < P.nout[l]) {
          srow_d[q] = P.doff[l] + sp; srow_h[q] = -1; sidx[q] = P.poff[l] + nw + sp;
          break;
        }
        sp -= P.nout[l];
      }
    }
  }
  double wsum = 0.0;

  const int64_t ntiles = (S.n + R - 1) / R;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0g = tile * R;
    // inputs: scaled coordinate and row weight (0 for the dead rows of the last tile)
    if (tid < R) {
      const int64_t gr = r0g + tid;
      T sv = T(0), wt = T(0);
      if (gr < S.n) {
        const T q = (M.axis && S.ncols > 1) ? S.pos[gr * S.stride_row + S.stride_col] : S.pos[gr * S.stride_row];
        sv = (q - M.lo) * M.inv_range;
        wt = S.w != nullptr ? S.w[gr] : T(1);
        if (M.inside_only && !(sv >= T(0) && sv <= T(1))) wt = T(0);   // conditional average on the support of the target
      }
      h[tid] = sv;
      wrow[tid] = wt;
      d[(P.doff[P.L - 1]) * RS + tid] = wt;     // delta of the single output unit
      wsum += (double)wt;
    }
    __syncthreads();
    // ---- phase 1, forward (the last Linear's output is not needed)
    for (int l = 0; l + 1 < P.L; ++l) {
      const int nin = P.nin[l], nout = P.nout[l];
      const T* W = M.theta + P.poff[l];
      const T* bias = W + nin * nout;
      const T* hin = h + (size_t)P.hoff[l] * RS;
      T* hout = h + (size_t)P.hoff[l + 1] * RS;
      T* gate = d + (size_t)P.doff[l] * RS;
      const bool activate = l >= 1;               // ves/basis.py:197-205 (no activation after the first Linear)
      const int items = (nout / 4) * RG;
      for (int item = tid; item < items; item += kMomThreads) {
        const int jg = item / RG, rg = item - jg * RG;
        const int j0 = jg * 4, r0 = rg * 8;
        T z[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const T bj = __ldg(bias + j0 + a);
#pragma unroll
          for (int c = 0; c < 8; ++c) z[a][c] = bj;
        }
        if (nin == 1) {
          T hv[8];
          lds8(hin + r0, hv);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const T w = __ldg(W + j0 + a);
#pragma unroll
            for (int c = 0; c < 8; ++c) z[a][c] = Math<T>::fma(w, hv[c], z[a][c]);
          }
        } else {
          for (int i = 0; i < nin; i += X::V) {
            T w[4][X::V];
#pragma unroll
            for (int a = 0; a < 4; ++a) ldg_vec(W + (size_t)(j0 + a) * nin + i, w[a]);
#pragma unroll
            for (int bb = 0; bb < X::V; ++bb) {
              T hv[8];
              lds8(hin + (size_t)(i + bb) * RS + r0, hv);
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 8; ++c) z[a][c] = Math<T>::fma(w[a][bb], hv[c], z[a][c]);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            T g = T(1), v = z[a][c];
            if (activate) v = M.act == PYVES_ACT_RELU ? act_fn<T, PYVES_ACT_RELU>(v, g) : act_fn<T, PYVES_ACT_TANH>(v, g);
            hout[(size_t)(j0 + a) * RS + r0 + c] = v;
            gate[(size_t)(j0 + a) * RS + r0 + c] = g;
          }
      }
      __syncthreads();
    }
    // ---- phase 1, backward: delta_{l-1} = gate_{l-1} * (W_l^T delta_l)
    for (int l = P.L - 1; l >= 1; --l) {
      const int nin = P.nin[l], nout = P.nout[l];
      const T* W = M.theta + P.poff[l];
      const T* dl = d + (size_t)P.doff[l] * RS;
      T* dprev = d + (size_t)P.doff[l - 1] * RS;
      const int items = (nin / 4) * RG;
      for (int item = tid; item < items; item += kMomThreads) {
        const int ig = item / RG, rg = item - ig * RG;
        const int i0 = ig * 4, r0 = rg * 8;
        T sacc4[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 8; ++c) sacc4[a][c] = T(0);
        for (int j = 0; j < nout; ++j) {
          T w[4];
          if (sizeof(T) == 4) {
            T t[X::V];
            ldg_vec(W + (size_t)j * nin + i0, t);
#pragma unroll
            for (int a = 0; a < 4; ++a) w[a] = t[a % X::V];
          } else {
            T t0[X::V], t1[X::V];
            ldg_vec(W + (size_t)j * nin + i0, t0);
            ldg_vec(W + (size_t)j * nin + i0 + 2, t1);
            w[0] = t0[0]; w[1] = t0[1 % X::V]; w[2] = t1[0]; w[3] = t1[1 % X::V];
          }
          T dv[8];
          lds8(dl + (size_t)j * RS + r0, dv);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 8; ++c) sacc4[a][c] = Math<T>::fma(w[a], dv[c], sacc4[a][c]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 8; ++c) dprev[(size_t)(i0 + a) * RS + r0 + c] *= sacc4[a][c];
      }
      __syncthreads();
    }
    // ---- phase 2: outer products of the big layers of this block, register tiles, strided ownership
#pragma unroll
    for (int g = 0; g < NB; ++g) {
      const int bl = bls[g];
      if (bl >= 0) {
        const int nin = P.nin[bl], nout = P.nout[bl];
        const T* dl = d + (size_t)P.doff[bl] * RS;
        const T* hl = h + (size_t)P.hoff[bl] * RS;
        for (int r = 0; r < R; r += X::V) {
          T dv[TS][X::V], hv[TS][X::V];
#pragma unroll
          for (int a = 0; a < TS; ++a) {
            const int j = tj + 16 * a;
            if (j < nout) lds_vec(dl + (size_t)j * RS + r, dv[a]);
            else
#pragma unroll
              for (int q = 0; q < X::V; ++q) dv[a][q] = T(0);
          }
#pragma unroll
          for (int b = 0; b < TS; ++b) {
            const int i = ti + 16 * b;
            if (i < nin) lds_vec(hl + (size_t)i * RS + r, hv[b]);
            else
#pragma unroll
              for (int q = 0; q < X::V; ++q) hv[b][q] = T(0);
          }
#pragma unroll
          for (int a = 0; a < TS; ++a)
#pragma unroll
            for (int b = 0; b < TS; ++b)
#pragma unroll
              for (int q = 0; q < X::V; ++q) acc[g][a][b] = Math<T>::fma(dv[a][q], hv[b][q], acc[g][a][b]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kMlp2Small; ++q) {
      if (sidx[q] >= 0) {
        const T* dr = d + (size_t)srow_d[q] * RS;
        T s0 = T(0);
        if (srow_h[q] >= 0) {
          const T* hr = h + (size_t)srow_h[q] * RS;
          for (int r = 0; r < R; ++r) s0 = Math<T>::fma(dr[r], hr[r], s0);
        } else {
          for (int r = 0; r < R; ++r) s0 += dr[r];
        }
        sacc[q] += s0;
      }
    }
    __syncthreads();
  }

  double* mypart = part + (int64_t)blockIdx.x * M.n_params;
#pragma unroll
  for (int g = 0; g < NB; ++g) {
    const int bl = bls[g];
    if (bl >= 0) {
      const int nin = P.nin[bl], nout = P.nout[bl];
#pragma unroll
      for (int a = 0; a < TS; ++a)
#pragma unroll
        for (int b = 0; b < TS; ++b) {
          const int j = tj + 16 * a, i = ti + 16 * b;
          if (j < nout && i < nin) mypart[P.poff[bl] + j * nin + i] = (double)acc[g][a][b];
        }
    }
  }
#pragma unroll
  for (int q = 0; q < kMlp2Small; ++q)
    if (sidx[q] >= 0) mypart[sidx[q]] = (double)sacc[q];
  if (y == 0) {
    __shared__ double wred[kMomThreads / 32];
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if ((tid & 31) == 0) wred[tid >> 5] = wsum;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int i = 0; i < kMomThreads / 32; ++i) t += wred[i];
      part_w[blockIdx.x] = t;
    }
  }
}

// true when the fast kernel applies; fills the plan
template <typename T> static bool make_mlp_plan(const MlpGradParams<T>& M, MlpPlan& P) {
  std::memset(&P, 0, sizeof(P));
  P.L = M.n_layers;
  if (P.L < 2 || P.L > 10 || M.sizes[0] != 1 || M.sizes[P.L] != 1) return false;
  int hoff = 0, doff = 0, poff = 0, wmax = 0;
  for (int l = 0; l < P.L; ++l) {
    const int nin = M.sizes[l], nout = M.sizes[l + 1];
    P.nin[l] = nin; P.nout[l] = nout; P.hoff[l] = hoff; P.doff[l] = doff; P.poff[l] = poff;
    if ((l == 0) != (nin == 1) || (l == P.L - 1) != (nout == 1)) return false;   // hidden widths >= 4
    if (nin > 1 && nin % 4 != 0) return false;
    if (nout > 1 && nout % 4 != 0) return false;
    if (nin > 128 || nout > 128) return false;
    if (poff % 4 != 0) return false;                      // 128-bit weight loads
    if (nin > 1 && nout > 1) P.big[P.nbig++] = l;
    else P.nsmall += nin * nout;
    P.nsmall += nout;
    if (nin > wmax) wmax = nin;
    if (nout > wmax) wmax = nout;
    hoff += nin; doff += nout; poff += nin * nout + nout;
  }
  P.units_in = hoff;
  P.units_out = doff;
  if (P.nsmall > kMlp2Small * kMomThreads) return false;
  const size_t sm = ((size_t)(P.units_in + P.units_out) * Mlp2<T>::kStride + Mlp2<T>::kRows) * sizeof(T);
  return sm <= 200 * 1024;
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
  MlpPlan plan;
  if (make_mlp_plan<T>(M, plan)) {
    using X = Mlp2<T>;
    const int64_t tiles2 = (S.n + X::kRows - 1) / X::kRows;
    const int G2 = (int)(tiles2 < kMlp2Blocks ? tiles2 : kMlp2Blocks);
    double* part2 = scratch;
    double* part_w2 = scratch + (size_t)kMaxBlocks * P;
    const size_t sm2 = ((size_t)(plan.units_in + plan.units_out) * X::kStride + X::kRows) * sizeof(T);
    int wmax = 0;
    for (int l = 0; l <= M.n_layers; ++l) wmax = M.sizes[l] > wmax ? M.sizes[l] : wmax;
    // narrow nets with several big layers: two of them per block (one forward / backward pass feeds both outer products)
    const bool pair = plan.nbig >= 2 && wmax <= 64;
    auto kern2 = pair ? (wmax <= 32 ? moments_mlp2_kernel<T, 2, 2> : moments_mlp2_kernel<T, 4, 2>)
                      : (wmax <= 32 ? moments_mlp2_kernel<T, 2, 1> : (wmax <= 64 ? moments_mlp2_kernel<T, 4, 1> : moments_mlp2_kernel<T, 8, 1>));
    const int ngroups = plan.nbig > 0 ? (pair ? (plan.nbig + 1) / 2 : plan.nbig) : 1;
    cudaError_t e2 = cudaSuccess;
    if (sm2 > 48 * 1024) e2 = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    if (e2 != cudaSuccess) return e2;
    kern2<<<dim3(G2, ngroups), kMomThreads, sm2, st>>>(S, M, plan, part2, part_w2);
    launch_reduce_partials(part2, G2, P, sum_wf, part_w2, sum_w, st);
    *launches += 2;
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

// ------------------------------------------------------------------------------------------------
// Rows -> pieces of the piecewise-linear ReLU network (see moments.cuh).  Block-private histograms in shared memory (FP64), flushed
// with one global atomic per non-empty bin.
template <typename T>
__global__ void __launch_bounds__(256) pwl_hist_kernel(const RowSource<T> S, const int axis, const T lo, const T inv_range, const int inside_only,
                                                       const T* __restrict__ brk, const int* __restrict__ count, const int cap,
                                                       double* __restrict__ hist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* hw = reinterpret_cast<double*>(smem_raw);          // [cap + 1] sum of weights
  double* hs = hw + (cap + 1);                               // [cap + 1] sum of weights times s
  T* b = reinterpret_cast<T*>(hs + (cap + 1));               // [cap] breakpoints
  const int nb = *count;
  for (int i = threadIdx.x; i <= cap; i += blockDim.x) hw[i] = hs[i] = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) b[i] = brk[i];
  __syncthreads();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < S.n; r += (int64_t)gridDim.x * blockDim.x) {
    const T q = S.pos[r * S.stride_row + (axis && S.ncols > 1 ? S.stride_col : 0)];
    const T s = (q - lo) * inv_range;
    T w = S.w != nullptr ? S.w[r] : T(1);
    if (inside_only && !(s >= T(0) && s <= T(1))) w = T(0);
    if (w == T(0)) continue;
    int p = 0, len = nb;
    while (len > 0) {
      const int half = len >> 1;
      const bool right = b[p + half] < s;
      p = right ? p + half + 1 : p;
      len = right ? len - half - 1 : half;
    }
    atomicAdd(&hw[p], (double)w);
    atomicAdd(&hs[p], (double)w * (double)s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= cap; i += blockDim.x) {
    if (hw[i] != 0.0) {
      atomicAdd(&hist[i], hw[i]);
      atomicAdd(&hist[cap + 1 + i], hs[i]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pwl_rows_kernel(const double* __restrict__ hist, const int cap, const T lo, const T inv_range,
                                                       T* __restrict__ rows, T* __restrict__ row_w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > cap) return;
  const double n = hist[i], sw = hist[cap + 1 + i];
  const double sbar = n != 0.0 ? sw / n : 0.0;
  const T x = (T)((double)lo + sbar / (double)inv_range);
  rows[2 * i] = x;
  rows[2 * i + 1] = x;
  row_w[i] = (T)n;
}

template <typename T>
cudaError_t launch_pwl_compress(const RowSource<T>& S, int axis, T lo, T inv_range, int inside_only, const T* brk, const int* count, int cap,
                                double* hist, T* rows, T* row_w, cudaStream_t st, int* launches) {
  cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(double) * 2 * (size_t)(cap + 1), st);
  if (e != cudaSuccess) return e;
  const size_t sm = sizeof(double) * 2 * (size_t)(cap + 1) + sizeof(T) * (size_t)cap + 16;
  auto kern = pwl_hist_kernel<T>;
  if (sm > 48 * 1024 - 64) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  int64_t blocks = (S.n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, 256, sm, st>>>(S, axis, lo, inv_range, inside_only, brk, count, cap, hist);
  pwl_rows_kernel<T><<<(cap + 1 + 255) / 256, 256, 0, st>>>(hist, cap, lo, inv_range, rows, row_w);
  *launches += 2;
  return cudaGetLastError();
}
template cudaError_t launch_pwl_compress<float>(const RowSource<float>&, int, float, float, int, const float*, const int*, int, double*, float*,
                                                float*, cudaStream_t, int*);
template cudaError_t launch_pwl_compress<double>(const RowSource<double>&, int, double, double, int, const double*, const int*, int, double*,
                                                 double*, double*, cudaStream_t, int*);

#define PV_INST(T)                                                                                               \
  template cudaError_t launch_moments_basis<T>(const RowSource<T>&, const BasisParams<T>&, double*, double*,     \
                                               double*, double*, cudaStream_t, int*);                            \
  template cudaError_t launch_moments_mlp<T>(const RowSource<T>&, const MlpGradParams<T>&, double*, double*,     \
                                             double*, cudaStream_t, int*);
PV_INST(float)
PV_INST(double)

}  // namespace pv
