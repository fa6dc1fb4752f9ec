// Basis covariance  C = sum_rows w f f^T  on the 5th-generation tensor cores (tcgen05, FP32 accumulate in
// TMEM) -- the one GEMM-shaped step of the path (textbook VES Hessian beta Cov_V(f), SURVEY Appendix
// A.4; the reference only names the option, README.md:24).
//
// Operand precision: the factors are products of two Legendre values, |P_a P_b| <= 1, and FP16 has the same
// 10-bit mantissa as TF32 on that range, so the tiles are FP16 (kind::f16): half the shared-memory bytes per
// element for the threads that write them AND for the tensor core that reads them, and twice the MMA rate of
// kind::tf32.  The kernel is bound by shared-memory bandwidth (tile writes + table reads + operand reads), not by
// the tensor pipe, so bytes per element are what counts (round 1, TF32 tiles: tensor pipe 40-50 % busy).
//
// One CTA per SM and per (row chunk, 256 x 256 output tile).  Per trip a CTA turns 64 walkers into a
// [256 basis functions] x [64 walkers] FP16 sub-tile written straight into shared memory in the canonical K-major
// SWIZZLE_128B UMMA layout (the basis matrix B[N, K] never exists in HBM): the per-walker Legendre tables are
// kept as FP16, every thread owns one basis function and builds 8 walkers per step from two 128-bit table loads,
// four packed HMUL2 and one 128-bit store.  One elected thread issues
//     D0[128 x 256] += F[0:128, :] F^T        D1[128 x 256] += F[128:256, :] F^T
// as 2 x 4 tcgen05.mma (M = 128, N = 256, K = 16); on a diagonal tile A and B are the SAME sub-tile and the
// lower-left 128 x 128 block of D1 is skipped (N = 128: symmetry).  The two accumulators fill the CTA's 512 TMEM
// columns.  Two shared-memory stages: the tensor core consumes stage s while all threads generate stage s^1;
// completion is tracked with tcgen05.commit -> mbarrier.  At the end the accumulators are read back with
// tcgen05.ld and written as per-CTA FP32 partials that a second kernel sums in fixed order into the FP64 result.
//
// Row weights w_r enter as sqrt(|w_r| / max|w|) folded into the x-table (both operands carry it, so the product
// carries w); rows of negative weight are accumulated by a second pass of the same kernel that runs only when a
// device-side scan found one, and the reduction subtracts it.
#include <cuda_fp16.h>

#include "moments.cuh"

namespace pv {

namespace {

constexpr int kTcThreads = 256;
constexpr int kTcRows = 64;                 // walkers per sub-tile = 128 bytes of FP16 = one swizzle atom
constexpr int kTcK = 256;                   // padded number of basis functions (UMMA N, 2 x UMMA M)

PV_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

PV_DEV uint64_t umma_desc_k_sw128(uint32_t saddr) {
  // start address [0,14) (>>4) | LBO [16,30) = 1 (unused for swizzled K-major) | SBO [32,46) = 1024 B >> 4
  // | version [46,48) = 1 (Blackwell) | layout type [61,64) = 2 (SWIZZLE_128B)
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// kind::f16 (A, B = FP16: format 0), FP32 accumulate, A and B K-major, M = 128, N = 256 | 128
constexpr uint32_t kIdesc256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((128u >> 4) << 24);

PV_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
PV_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
PV_DEV void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
PV_DEV bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
PV_DEV void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The shared-memory struct is reached through a pointer rounded up to 1024 bytes, so the compiler falls back to
// generic loads and stores; these keep the hot accesses in the shared window.
PV_DEV uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
PV_DEV void sts16(uint32_t addr, __half v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(v)) : "memory");
}
PV_DEV uint32_t hmul2(uint32_t a, uint32_t b) {     // packed FP16 product of two register pairs
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
PV_DEV void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Legendre recursion P_{n+1} = A_n s P_n - B_n P_{n-1}: A_n = (2n+1)/(n+1), B_n = n/(n+1)
__constant__ float c_leg_a[PYVES_MAX_DEGREE + 1] = {1.000000000e+00f, 1.500000000e+00f, 1.666666667e+00f, 1.750000000e+00f, 1.800000000e+00f, 1.833333333e+00f, 1.857142857e+00f, 1.875000000e+00f, 1.888888889e+00f, 1.900000000e+00f, 1.909090909e+00f, 1.916666667e+00f, 1.923076923e+00f, 1.928571429e+00f, 1.933333333e+00f, 1.937500000e+00f, 1.941176471e+00f, 1.944444444e+00f, 1.947368421e+00f, 1.950000000e+00f, 1.952380952e+00f, 1.954545455e+00f, 1.956521739e+00f, 1.958333333e+00f, 1.960000000e+00f, 1.961538462e+00f, 1.962962963e+00f, 1.964285714e+00f, 1.965517241e+00f, 1.966666667e+00f, 1.967741935e+00f, 1.968750000e+00f, 1.969696970e+00f, 1.970588235e+00f, 1.971428571e+00f, 1.972222222e+00f, 1.972972973e+00f, 1.973684211e+00f, 1.974358974e+00f, 1.975000000e+00f, 1.975609756e+00f, 1.976190476e+00f, 1.976744186e+00f, 1.977272727e+00f, 1.977777778e+00f, 1.978260870e+00f, 1.978723404e+00f, 1.979166667e+00f, 1.979591837e+00f, 1.980000000e+00f, 1.980392157e+00f, 1.980769231e+00f, 1.981132075e+00f, 1.981481481e+00f, 1.981818182e+00f, 1.982142857e+00f, 1.982456140e+00f, 1.982758621e+00f, 1.983050847e+00f, 1.983333333e+00f, 1.983606557e+00f, 1.983870968e+00f, 1.984126984e+00f, 1.984375000e+00f};
__constant__ float c_leg_b[PYVES_MAX_DEGREE + 1] = {0.000000000e+00f, 5.000000000e-01f, 6.666666667e-01f, 7.500000000e-01f, 8.000000000e-01f, 8.333333333e-01f, 8.571428571e-01f, 8.750000000e-01f, 8.888888889e-01f, 9.000000000e-01f, 9.090909091e-01f, 9.166666667e-01f, 9.230769231e-01f, 9.285714286e-01f, 9.333333333e-01f, 9.375000000e-01f, 9.411764706e-01f, 9.444444444e-01f, 9.473684211e-01f, 9.500000000e-01f, 9.523809524e-01f, 9.545454545e-01f, 9.565217391e-01f, 9.583333333e-01f, 9.600000000e-01f, 9.615384615e-01f, 9.629629630e-01f, 9.642857143e-01f, 9.655172414e-01f, 9.666666667e-01f, 9.677419355e-01f, 9.687500000e-01f, 9.696969697e-01f, 9.705882353e-01f, 9.714285714e-01f, 9.722222222e-01f, 9.729729730e-01f, 9.736842105e-01f, 9.743589744e-01f, 9.750000000e-01f, 9.756097561e-01f, 9.761904762e-01f, 9.767441860e-01f, 9.772727273e-01f, 9.777777778e-01f, 9.782608696e-01f, 9.787234043e-01f, 9.791666667e-01f, 9.795918367e-01f, 9.800000000e-01f, 9.803921569e-01f, 9.807692308e-01f, 9.811320755e-01f, 9.814814815e-01f, 9.818181818e-01f, 9.821428571e-01f, 9.824561404e-01f, 9.827586207e-01f, 9.830508475e-01f, 9.833333333e-01f, 9.836065574e-01f, 9.838709677e-01f, 9.841269841e-01f, 9.843750000e-01f};

constexpr int kSubBytes = kTcK * kTcRows * 2;    // one [256 basis functions] x [64 walkers] FP16 sub-tile, 32 KB
constexpr int kSuperRows = 128;                  // walkers per table build (thread = (walker, axis): all 256 threads)
constexpr int kTabRS = kSuperRows + 8;           // table row stride (halves): 272 B, conflict-free 128-bit reads
constexpr int kTabRows = 2 * (PYVES_MAX_DEGREE + 1);
int64_t g_max_rows_per_chunk = 1 << 14;          // rows accumulated in the FP32 TMEM accumulators between two drains (tunable: cov_tc_set_max_rows)

struct CovSmem {
  alignas(1024) unsigned char tile[2][2 * kSubBytes];   // 2 stages x 2 sub-tiles
  alignas(16) __half tab[2][kTabRows * kTabRS];         // tab[.][a][r] = sqrt(w_r) P_a(sx_r) (a < ka), tab[.][ka + b][r] = P_b(sy_r)
  alignas(8) uint64_t stage_bar[2];
  alignas(8) uint64_t done_bar;
  uint32_t tmem_base;
};

// wstat[0] = max |w|, wstat[1] = 1 when some weight is negative (float bits); one block, fixed order
__global__ void __launch_bounds__(1024) cov_weight_scan_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ wstat) {
  __shared__ float smax[32];
  __shared__ int sneg[32];
  float m = 0.f;
  int neg = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = w[i];
    m = fmaxf(m, fabsf(v));
    neg |= v < 0.f;
  }
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    neg |= __shfl_xor_sync(0xffffffffu, neg, o);
  }
  if ((threadIdx.x & 31) == 0) {
    smax[threadIdx.x >> 5] = m;
    sneg[threadIdx.x >> 5] = neg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
      m = fmaxf(m, smax[i]);
      neg |= sneg[i];
    }
    wstat[0] = m;
    wstat[1] = neg ? 1.f : 0.f;
  }
}

// Output tiling: C is cut into 256 x 256 tiles (tm <= tn); blockIdx.y picks the tile, blockIdx.x the chunk of
// rows.  A diagonal tile needs ONE operand tile F_m (A = B = F_m): a trip covers 128 walkers as two
// [256 x 64] sub-tiles.  An off-diagonal tile needs F_m and F_n: a trip covers 64 walkers, sub-tile 0 = F_m,
// sub-tile 1 = F_n.  Either way a stage is 64 KB and every thread writes 2 x 64 values per trip.
// pass 0 accumulates the rows of positive weight, pass 1 (launched only with weights) those of negative weight and
// returns at once when the scan found none.
__global__ void __launch_bounds__(kTcThreads, 1)
cov_f16_kernel(const RowSource<float> S, const BasisParams<float> B, int T256, int64_t rows_per_chunk, int64_t seg_rows,
               const float* __restrict__ wstat, int pass, float* __restrict__ part) {
  if (pass == 1 && wstat[1] == 0.f) return;      // uniform over the grid: no row of negative weight
  extern __shared__ unsigned char smem_raw[];
  CovSmem& sm = *reinterpret_cast<CovSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = B.ka * B.kb;
  int tm = 0, rem = blockIdx.y;
  while (rem >= T256 - tm) {
    rem -= T256 - tm;
    ++tm;
  }
  const int tn = tm + rem;
  const bool diag = tm == tn;
  float winv = 1.f, wsign = 1.f;
  if (S.w != nullptr) {
    const float wabs = wstat[0];
    winv = wabs > 0.f ? 1.f / wabs : 0.f;
    wsign = pass == 0 ? 1.f : -1.f;
  }

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    mbar_init(&sm.stage_bar[0], 1);
    mbar_init(&sm.stage_bar[1], 1);
    mbar_init(&sm.done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  // this thread owns row `tid` of both sub-tiles: basis function k = (fa, fb) factors, or none (zero row)
  int fa[2], fb[2];
  bool kvalid[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int k = ((q == 1 && !diag) ? tn : tm) * kTcK + tid;
    kvalid[q] = k < K;
    fa[q] = kvalid[q] ? k / B.kb : 0;
    fb[q] = kvalid[q] ? B.ka + (k - (k / B.kb) * B.kb) : 0;
  }
  const uint32_t row_off = (uint32_t)(tid >> 3) * 1024u + (uint32_t)(tid & 7) * 128u;

  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t row_end = (row_begin + rows_per_chunk) < S.n ? (row_begin + rows_per_chunk) : S.n;
  // Pipeline per 128 walkers: [table build, thread = (walker, axis)] -> sync -> [all threads fill a stage] -> sync ->
  // [thread kIssuer queues the MMAs].  The coordinates of the NEXT 128 walkers are loaded one iteration ahead and the
  // table is double buffered (no third barrier), so the tensor core works on stage s while the warps already build the
  // next tables and generate stage s ^ 1.
  constexpr int kIssuer = 192;
  const bool issuing_warp = __shfl_sync(0xffffffffu, tid >> 5, 0) == (kIssuer >> 5);
  const int rr = tid & (kSuperRows - 1), ax = tid >> 7;
  float px = 0.f, py = 0.f, pw = 1.f;
  bool plive = false;
  auto prefetch = [&](int64_t r0) {
    const int64_t r = r0 + rr;
    plive = r < row_end;
    if (plive) {
      px = S.pos[r * S.stride_row];
      py = S.pos[r * S.stride_row + S.stride_col];
      if (S.w != nullptr) pw = S.w[r];
    }
  };
  prefetch(row_begin);
  int trip = 0;   // stage uses so far
  int it = 0;
  // The rows of the chunk are cut into segments of seg_rows: the tensor core truncates when it aligns the addends of its FP32
  // accumulation (measured: the diagonal comes out low by 2^-27 per accumulated row, tools/diag/cov_accum_diag.py), so the
  // TMEM accumulators are drained into this CTA's partial tile (rounded FP32 adds, L2 resident) after every segment.
  int seg = 0;
  const int acc = warp >> 2, lq = warp & 3;
  float* const out = part + (((size_t)blockIdx.x * gridDim.y + blockIdx.y) * kTcK + (size_t)(acc * 128 + lq * 32 + lane)) * kTcK;
  for (int64_t s0 = row_begin; s0 < row_end || seg == 0; s0 += seg_rows, ++seg) {
  const int64_t seg_end = (s0 + seg_rows) < row_end ? (s0 + seg_rows) : row_end;
  int seg_trip = 0;
  for (int64_t r0 = s0; r0 < seg_end; r0 += kSuperRows, ++it) {
    const uint32_t tab = smem_u32(sm.tab[it & 1]);
    // ---- Legendre tables of 128 walkers (FP32 recursion, stored as FP16): no divisions
    {
      const int kn = ax ? B.kb : B.ka;
      const uint32_t t = tab + (uint32_t)(((ax ? B.ka : 0) * kTabRS + rr) * 2);
      bool live = plive;
      float sv = 0.f, amp = 1.f;
      if (live) {
        bool in_a, in_b;
        const float sx = scale_clamp(px, B.ca, B.iha, in_a), sy = scale_clamp(py, B.cb, B.ihb, in_b);
        if (B.inside_only && !(in_a && in_b)) live = false;   // weight 0: the row contributes nothing to F F^T
        sv = ax ? sy : sx;
        if (S.w != nullptr) {
          const float wv = wsign * pw * winv;                 // this pass's share of the weight, in [0, 1]
          if (!(wv > 0.f)) live = false;
          else if (!ax) amp = sqrtf(wv);                      // both operands carry sqrt(w): the product carries w
        }
      }
      prefetch(r0 + kSuperRows);                               // in flight during the recursion and the fill
      if (live) {
        float pm = 1.f, p = sv;
        sts16(t, __float2half_rn(amp));
        if (kn > 1) sts16(t + kTabRS * 2, __float2half_rn(amp * p));
        for (int n = 1; n + 1 < kn; ++n) {
          const float pn = fmaf(c_leg_a[n] * sv, p, -c_leg_b[n] * pm);   // uniform index: constant-bank operands
          sts16(t + (uint32_t)((n + 1) * kTabRS * 2), __float2half_rn(amp * pn));
          pm = p;
          p = pn;
        }
      } else {
        for (int n = 0; n < kn; ++n) sts16(t + (uint32_t)(n * kTabRS * 2), __ushort_as_half((unsigned short)0));
      }
    }
    __syncthreads();
    // ---- one trip (diagonal: 128 walkers) or two trips (off-diagonal: 64 walkers each)
    const int ntr = diag ? 1 : 2;
    for (int hh = 0; hh < ntr; ++hh, ++trip, ++seg_trip) {
      const int s = trip & 1, use = trip >> 1;
      if (use >= 1) mbar_wait(&sm.stage_bar[s], (uint32_t)((use - 1) & 1));   // the MMAs that read this stage are done
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int rbase = diag ? kTcRows * q : kTcRows * hh;
        const uint32_t ta = tab + (uint32_t)((fa[q] * kTabRS + rbase) * 2);
        const uint32_t tb = tab + (uint32_t)((fb[q] * kTabRS + rbase) * 2);
        const uint32_t base = smem_u32(sm.tile[s]) + (uint32_t)q * kSubBytes + row_off;
        if (kvalid[q]) {
          uint4 a4[8], b4[8];     // all 16 loads in flight before the first use (the asm loads are not reordered)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            a4[c] = lds128(ta + c * 16);
            b4[c] = lds128(tb + c * 16);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 v = make_uint4(hmul2(a4[c].x, b4[c].x), hmul2(a4[c].y, b4[c].y), hmul2(a4[c].z, b4[c].z), hmul2(a4[c].w, b4[c].w));
            sts128(base + (uint32_t)((c ^ (tid & 7)) << 4), v);   // Swizzle<3,4,3>: 16-byte chunk index ^ (row % 8)
          }
        } else {     // padding rows of the last tile (k >= K)
#pragma unroll
          for (int c = 0; c < 8; ++c) sts128(base + (uint32_t)(c << 4), make_uint4(0u, 0u, 0u, 0u));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      __syncthreads();
      if (issuing_warp && elect_one()) {                            // warp-uniform branch, one elected lane: descriptors stay in uniform registers
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(sm.tile[s]);
        const int nsub = diag ? 2 : 1;
        for (int sub = 0; sub < nsub; ++sub) {
          const uint32_t abase = a0 + (uint32_t)(diag ? sub : 0) * kSubBytes;
          const uint32_t bbase = a0 + (uint32_t)(diag ? sub : 1) * kSubBytes;
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            // rows 128 m .. 128 m + 127 of A.  Diagonal tile, m = 1: only the columns 128 .. 255 are needed (symmetry):
            // B starts at its row 128, N = 128, D at column 128 of the second accumulator
            const bool half_n = diag && m == 1;
            const uint64_t da = umma_desc_k_sw128(abase + (uint32_t)m * 128u * 128u);
            const uint64_t db = umma_desc_k_sw128(bbase + (half_n ? 128u * 128u : 0u));
            const uint32_t d = tmem + (uint32_t)m * 256u + (half_n ? 128u : 0u);
#pragma unroll
            for (int kk = 0; kk < kTcRows / 16; ++kk)       // K = 16 FP16 = 32 bytes per instruction: +2 in 16-byte units
              umma_f16(d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), half_n ? kIdesc128 : kIdesc256,
                       (seg_trip > 0 || sub > 0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(&sm.stage_bar[s]);
      }
    }
  }
  // ---- drain: all MMAs of the segment done?  (the issuer's commit covers every MMA it queued before)
  if (seg_trip > 0) {
    if (issuing_warp && elect_one()) umma_commit(&sm.done_bar);
    mbar_wait(&sm.done_bar, (uint32_t)(seg & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // TMEM -> registers -> partial [256][256] floats of this (chunk, tile).  Warp w reads lanes 32 (w % 4) .. +31
  // of accumulator w / 4; thread = one row of D.  (The skipped block of a diagonal tile holds stale TMEM: never read.)
#pragma unroll 1
  for (int c0 = 0; c0 < kTcK; c0 += 32) {
    uint32_t v[32];
    if (seg_trip > 0 && !(diag && acc == 1 && c0 < 128)) {
      const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * 256 + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = 0u;      // a chunk without rows contributes zeros
    }
#pragma unroll
    for (int q = 0; q < 32; q += 4) {
      float4 o = make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
      if (seg > 0) {
        const float4 p = *reinterpret_cast<const float4*>(out + c0 + q);
        o = make_float4(o.x + p.x, o.y + p.y, o.z + p.z, o.w + p.w);
      }
      *reinterpret_cast<float4*>(out + c0 + q) = o;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();                                   // every accumulator read is done before the next segment overwrites it
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }   // segments
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// out[k][l] (FP64, K x K) = wabs (sum over the row chunks of part[chunk][tile][i][j] - the same of the negative pass);
// only the upper triangle of a tile is read and mirrored (exact symmetry).  One thread per element of the upper
// triangle, four independent partial sums (fixed order).
__global__ void __launch_bounds__(256) cov_tc_reduce_kernel(const float* __restrict__ part, const float* __restrict__ part_neg,
                                                            const float* __restrict__ wstat, int G, int T256, int K, double* __restrict__ out) {
  const int tile = blockIdx.y;
  int tm = 0, rem = tile;
  while (rem >= T256 - tm) {
    rem -= T256 - tm;
    ++tm;
  }
  const int tn = tm + rem;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = e / kTcK, j = e % kTcK;
  const int k = tm * kTcK + i, l = tn * kTcK + j;
  if (k >= K || l >= K || l < k) return;
  const int ntile = gridDim.y;
  const size_t stride = (size_t)ntile * kTcK * kTcK;
  const float* p0 = part + ((size_t)tile * kTcK + i) * kTcK + j;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int g = 0; g < G; ++g) s[g & 3] += (double)p0[g * stride];
  double t = (s[0] + s[1]) + (s[2] + s[3]);
  if (wstat != nullptr) {
    if (wstat[1] != 0.f) {
      const float* p1 = part_neg + ((size_t)tile * kTcK + i) * kTcK + j;
      double u[4] = {0.0, 0.0, 0.0, 0.0};
      for (int g = 0; g < G; ++g) u[g & 3] += (double)p1[g * stride];
      t -= (u[0] + u[1]) + (u[2] + u[3]);
    }
    t *= (double)wstat[0];
  }
  out[(size_t)k * K + l] = t;
  out[(size_t)l * K + k] = t;
}

}  // namespace

// smallest basis that goes to the tensor cores (a tile is padded to 256 functions)
constexpr int kCovTcMinK = 64;   // measured: K = 100 0.39 ms (tensor) vs 1.2 ms (CUDA cores) per 10^6 walkers; K <= 64 equal

bool cov_tc_supported(const RowSource<float>& S, const BasisParams<float>& B) {
  const int K = B.ka * B.kb;
  return B.kind == PYVES_BIAS_LEGENDRE_2D && K > kCovTcMinK && K <= 4096 && S.ncols >= 2;
}

// grid: base = n_sm / ntile CTAs per tile (one CTA per SM), each with a contiguous chunk of the rows
static void cov_tc_grid(int K, int64_t n, int n_sm, int& T256, int& ntile, int& chunks, int64_t& rows_per_chunk) {
  T256 = (K + kTcK - 1) / kTcK;
  ntile = T256 * (T256 + 1) / 2;
  chunks = n_sm / ntile;
  if (chunks < 1) chunks = 1;
  const int64_t supers = (n + kSuperRows - 1) / kSuperRows;
  if (chunks > supers) chunks = (int)supers;
  if (chunks < 1) chunks = 1;
  rows_per_chunk = ((supers + chunks - 1) / chunks) * kSuperRows;
}

void cov_tc_set_max_rows(int64_t rows) { g_max_rows_per_chunk = rows < kSuperRows ? kSuperRows : rows; }

// floats of partial tiles the kernel writes (x 2 with row weights: the negative pass has its own) + the weight scan's result
size_t cov_tc_scratch_floats(int K, int64_t n, int n_sm, bool weighted) {
  int T256, ntile, chunks;
  int64_t rpc;
  cov_tc_grid(K, n, n_sm, T256, ntile, chunks, rpc);
  return (size_t)chunks * ntile * kTcK * kTcK * (weighted ? 2 : 1) + 4;
}

cudaError_t launch_cov_tc(const RowSource<float>& S, const BasisParams<float>& B, double* sum_wff, float* scratch, int n_sm,
                          cudaStream_t st, int* launches) {
  const int K = B.ka * B.kb;
  int T256, ntile, chunks;
  int64_t rows_per_chunk;
  cov_tc_grid(K, S.n, n_sm, T256, ntile, chunks, rows_per_chunk);
  const size_t smem = sizeof(CovSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(cov_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const size_t per_pass = (size_t)chunks * ntile * kTcK * kTcK;
  const int64_t seg_rows = (g_max_rows_per_chunk / kSuperRows) * kSuperRows;
  float* wstat = nullptr;
  float* part_neg = nullptr;
  if (S.w != nullptr) {
    wstat = scratch + 2 * per_pass;
    part_neg = scratch + per_pass;
    cov_weight_scan_kernel<<<1, 1024, 0, st>>>(S.w, S.n, wstat);
    *launches += 1;
  }
  cov_f16_kernel<<<dim3(chunks, ntile), kTcThreads, smem, st>>>(S, B, T256, rows_per_chunk, seg_rows, wstat, 0, scratch);
  *launches += 1;
  if (S.w != nullptr) {
    cov_f16_kernel<<<dim3(chunks, ntile), kTcThreads, smem, st>>>(S, B, T256, rows_per_chunk, seg_rows, wstat, 1, part_neg);
    *launches += 1;
  }
  cov_tc_reduce_kernel<<<dim3(kTcK * kTcK / 256, ntile), 256, 0, st>>>(scratch, part_neg, wstat, chunks, T256, K, sum_wff);
  *launches += 1;
  return cudaGetLastError();
}

}  // namespace pv
