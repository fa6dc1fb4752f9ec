// Experimental step kernel: the per-step contraction of the 2D tensor-product Legendre bias on the tensor cores.
//
//   A_i = sum_j w_ij P_j(y'),  B_i = sum_j w_ij P'_j(y')   for the 128 walkers of a warp group
//
// is a [128 walkers x KY] x [KY x KX] product per step and per group.  The CUDA-core kernels (bias.cuh) spend 4 KX KY flops per
// walker-step on it (64 x 64: 94 % of the step); here a group of 128 threads (thread = walker) writes its rows
// (P_j, P'_j), split into FP16 hi + lo parts (22 mantissa bits together: FP32-level accuracy), straight into shared memory
// in the canonical K-major SWIZZLE_128B UMMA layout, one thread issues
//     D_A = P_hi W_hi^T + P_hi W_lo^T + P_lo W_hi^T            D_B = the same with P'
// as tcgen05.mma kind::f16 (M = 128, N = KX, K = 16) into the group's TMEM columns, and after the commit every thread
// reads ITS walker's row (A_i, B_i) back with tcgen05.ld (TMEM lane = walker) and finishes the step (x recursion, row
// sweep, potential, noise, integrator) in registers.  A walker's steps are sequential, so the MMA round trip is latency
// for that group; G groups per CTA (each with its own tiles, TMEM columns and mbarrier) keep the SM busy meanwhile.
//
// Selected with pyves_set_tuning(h, 7) for FP32 handles, kx, ky <= 64 (profiles/r02_tc_step.md has the measurements).
#include <cuda_fp16.h>

#include <cstdlib>

#include "langevin.cuh"

namespace pv {

namespace {

constexpr int kGT = 128;                    // threads (= walkers) per group
constexpr int kTileBytes = 128 * 128;       // [128 walkers][64 FP16] operand tile, 16 KB
constexpr int kWBytes = 64 * 128;           // [64 i][64 j FP16] weight tile, 8 KB
constexpr float kScaleP = 1024.f;           // P_j is carried as 1024 P_j, P'_j as 16 P'_j (|P'_63| = 2016): the FP16 lo parts stay normal
constexpr float kScaleD = 16.f;

PV_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

PV_DEV uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
PV_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
PV_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TCS_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TCS_DONE;\n\t"
      "bra TCS_WAIT;\n\t"
      "TCS_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
PV_DEV bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
PV_DEV void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
PV_DEV void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
PV_DEV void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
PV_DEV void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

// 8 FP32 values -> 8 FP16 hi + 8 FP16 lo (x = hi + lo up to 2^-22 |x|), packed in memory order.  The remainder x - hi is ONE
// mixed-precision instruction per value (PTX fma.rn.f32.f16 -> SASS FHFMA: FP16 hi times -1 plus FP32 x, exact), not a
// conversion followed by a subtraction.
PV_DEV float sub_half(float x, uint32_t packed, int upper) {
  const unsigned short h = (unsigned short)(upper ? packed >> 16 : packed & 0xffffu), m1 = 0xBC00;   // -1.0 in FP16
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(h), "h"(m1), "f"(x));
  return d;
}
PV_DEV void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __half2 hh = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
    h[q] = *reinterpret_cast<const uint32_t*>(&hh);
    const __half2 ll = __floats2half2_rn(sub_half(v[2 * q], h[q], 0), sub_half(v[2 * q + 1], h[q], 1));
    l[q] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

PV_DEV void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]);
}

PV_DEV void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(r[q]);
}

// Scaled monic recursion.  With k_j the leading coefficient of P_j and e_j = round(log2 k_j), q_j = P_j 2^e_j / k_j obeys
//     q_{j+1} = 2^d_j s q_j - cc_j q_{j-1},   d_j = e_{j+1} - e_j in {0, 1},   cc_j = j^2 / (4 j^2 - 1) * 2^(e_{j+1} - e_{j-1})
// -- one FMUL and one FFMA per function (the textbook form needs a second FMUL for A_j s) with |q_j| <= 1.41, and
// P_j = kappa_j q_j, kappa_j = k_j / 2^e_j in [0.71, 1.41], is folded into the coefficient matrices.
struct MonicTab {
  float cc[64];
  int d[64];
  double kappa[64];
};
__host__ __device__ constexpr MonicTab make_monic_tab() {
  MonicTab t{};
  double kap = 1.0;
  int e_prev = 0, e_cur = 0;                          // e_{j-1}, e_j
  t.kappa[0] = 1.0;
  for (int j = 0; j < 63; ++j) {                      // step j -> j + 1
    double kn = kap * (double)(2 * j + 1) / (double)(j + 1);
    int dn = 0;
    if (kn >= 1.4142135623730951) {
      kn *= 0.5;
      dn = 1;
    }
    const int e_next = e_cur + dn;
    double c = j == 0 ? 0.0 : (double)j * j / (4.0 * j * j - 1.0);
    for (int q = 0; q < e_next - e_prev; ++q) c *= 2.0;
    t.d[j] = dn;
    t.cc[j] = (float)c;
    t.kappa[j + 1] = kn;
    kap = kn;
    e_prev = e_cur;
    e_cur = e_next;
  }
  return t;
}
PV_DEV double monic_kappa(int n) {                    // the same recurrence at run time (prologue only)
  double kap = 1.0;
  for (int j = 0; j < n; ++j) {
    kap = kap * (double)(2 * j + 1) / (double)(j + 1);
    if (kap >= 1.4142135623730951) kap *= 0.5;
  }
  return kap;
}

// Folded derivatives.  P'_n = sum over k = n-1, n-3, ... of (2k+1) P_k, so BOTH force components are contractions with the
// Legendre VALUES only, against two matrices derived from the coefficients once per launch:
//     dV/dsx = sum_k P_k(sx) A'_k,   A'_k = sum_j WA[k][j] P_j(sy),   WA[k][j] = (2k+1) sum_{i>k, i-k odd} w_ij
//     dV/dsy = sum_i P_i(sx) B_i,    B_i  = sum_k WB[i][k] P_k(sy),   WB[i][k] = (2k+1) sum_{j>k, j-k odd} w_ij
// A thread therefore writes ONE row per step (P_j(sy) as FP16 hi + lo; no derivative recursion, no second split), the
// tensor core multiplies it with the stacked [2 KX x KY] matrix (WA over WB) in one N = 2 KX instruction per term, and the
// x sweep needs P_i(sx) only.
//
// The two operand matrices of a group (P_hi, P_lo: [128 walkers] x [KYP functions] FP16 each) share 128-byte rows:
// operand s starts at byte 2 s KYP of the concatenated row -- one 16 KB tile per group for KYP <= 32, two for 64.  A UMMA
// descriptor may start anywhere inside the swizzle row (that is how the K slabs of one matrix are addressed as well).
template <int KYP> struct TcLayout {
  static constexpr int kTiles = (2 * KYP * 2 + 127) / 128;                           // 16 KB tiles per group
  static __host__ __device__ constexpr int tile(int s) { return (s * KYP * 2) / 128; }
  static __host__ __device__ constexpr int chunk0(int s) { return ((s * KYP * 2) % 128) / 16; }   // first 16-byte chunk of operand s
};

constexpr int kW2Bytes = 128 * 128;         // [WA rows 0..KXP) | WB rows KXP..2 KXP)][64 j FP16], 128-byte rows: 16 KB

template <int KYP, int G, int WPT> struct TcSmem {
  alignas(1024) unsigned char a[G][WPT][TcLayout<KYP>::kTiles][kTileBytes];
  alignas(1024) unsigned char w[2][kW2Bytes];        // stacked weights, hi and lo parts
  alignas(8) uint64_t bar[G][WPT];                   // contraction done (tcgen05.commit)
  alignas(8) uint64_t ready[G][WPT];                 // all 128 rows of the set written and last step's TMEM reads done (128 arrivals)
  uint32_t tmem_base;
};

__host__ __device__ constexpr uint32_t tmem_cols(int need) { return need <= 32 ? 32u : need <= 64 ? 64u : need <= 128 ? 128u : need <= 256 ? 256u : 512u; }

// How the sets of 128 walkers are dealt to the CTAs of a launch (see the kernel and launch_tc)
struct TcGrid {
  int full, tail;
};

// One CTA (or `resident` CTAs) per SM and every CTA of a round takes the same time, so a launch costs whole rounds: full CTAs
// (cap sets each) are dealt in whole rounds, the remainder goes into a last round of small CTAs spread evenly over the SMs.
// Returns the grid size.
inline unsigned tc_deal_sets(int64_t n, int cap, int wpt, int resident, TcGrid& GR) {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  if (resident < 1) resident = 1;
  const int64_t sets = (n + kGT - 1) / kGT, slots = (int64_t)n_sm * resident;
  const int64_t full = (sets / (cap * slots)) * slots, rem = sets - full * cap;
  int64_t tail_sets = 1, tail_ctas = 0;
  if (rem > 0) {
    tail_sets = (rem + slots - 1) / slots;
    if (wpt > 1) tail_sets = (tail_sets + wpt - 1) / wpt * wpt;     // whole groups
    if (tail_sets > cap) tail_sets = cap;
    tail_ctas = (rem + tail_sets - 1) / tail_sets;
  }
  GR.full = (int)full;
  GR.tail = (int)tail_sets;
  return (unsigned)(full + tail_ctas);
}

// Once per launch, one block: folded matrices WA[k][j], WB[i][k] (FP64 sums of the FP32 coefficients, times kappa_row
// kappa_column), their common power-of-two scale (max |W| into [2^12, 2^13): FP16 range; found on the device so that coefficients
// handed over on the device need no host copy), and the stacked B operand image -- row r < KXP = WA[r][.], row KXP + i = WB[i][.],
// 64 j (K index) per 128-byte row, hi and lo parts, K-major SWIZZLE_128B -- exactly as the step kernel's shared memory holds it.
// img: [hi 16 KB][lo 16 KB][wscale].
__global__ void __launch_bounds__(1024, 1) tc_weight_prep_kernel(const Legendre2DSmemParams<float> BP, const int KXP, unsigned char* __restrict__ img) {
  __shared__ float wab[2 * 64 * 64];
  __shared__ double kappa[64];
  __shared__ float red[32];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 64) kappa[tid] = monic_kappa(tid);
  __syncthreads();
  float wmax = 0.f;
  for (int item = tid; item < 2 * 64 * 64; item += blockDim.x) {
    const int which = item >> 12, r = (item >> 6) & 63, c = item & 63;
    double acc = 0.0;
    if (which == 0) {
      // WA[k = r][j = c]: (2k+1) times the sum over i = k + 1, k + 3, ... of w_ij
      if (r < BP.kx && c < BP.ky)
        for (int i = r + 1; i < BP.kx; i += 2) acc += (double)BP.w[i * BP.ky + c];
      acc *= (double)(2 * r + 1);
    } else {
      // WB[i = r][k = c]: (2k+1) times the sum over j = k + 1, k + 3, ... of w_ij
      if (r < BP.kx && c < BP.ky)
        for (int j = c + 1; j < BP.ky; j += 2) acc += (double)BP.w[r * BP.ky + j];
      acc *= (double)(2 * c + 1);
    }
    const float v = (float)(acc * kappa[r] * kappa[c]);
    wab[item] = v;
    wmax = fmaxf(wmax, fabsf(v));
  }
  for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if ((tid & 31) == 0) red[warp] = wmax;
  __syncthreads();
  float m = 0.f;
  for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, red[q]);
  int ex = 0;
  if (m > 0.f && m < 3e38f) frexpf(m, &ex);
  const float wscale = ldexpf(1.f, 13 - ex);
  if (tid == 0) *reinterpret_cast<float*>(img + 2 * kW2Bytes) = wscale;
  for (int item = tid; item < 2 * KXP * 8; item += blockDim.x) {
    const int r = item >> 3, c = item & 7;
    const float* src = (r < KXP ? wab + r * 64 : wab + 64 * 64 + (r - KXP) * 64) + 8 * c;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = src[q] * wscale;
    uint4 hi, lo;
    split8(v, hi, lo);
    const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(img + off) = hi;
    *reinterpret_cast<uint4*>(img + kW2Bytes + off) = lo;
  }
}

// KXP / KYP = number of x / y functions padded to 16, 32 or 64 (UMMA N = 2 KXP, K extent KYP; weights beyond kx, ky are zero).
// WPT walker sets per group (thread = one walker of each set; with WPT = 2 the two recursions of a thread run as one packed
// FP32x2 chain).  The MMA round trip (store -> proxy fence -> barrier -> MMA -> commit -> TMEM load, ~1.3 us) is latency
// that only independent groups can hide.
template <int KXP, int KYP, int G, int WPT, bool INJECT, class POT>
__global__ void __launch_bounds__(kGT* G, (G <= 2 ? 3 : G == 3 ? 2 : 1))
langevin2d_tc_kernel(const StepArgs<float> A, const IntegratorParams<float> I, const PotParams<float> PP,
                     const Legendre2DSmemParams<float> BP, const RngKeys K, const unsigned char* __restrict__ wimg,
                     const TcGrid GR) {
  using LY = TcLayout<KYP>;
  extern __shared__ unsigned char smem_raw[];
  TcSmem<KYP, G, WPT>& sm = *reinterpret_cast<TcSmem<KYP, G, WPT>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = __shfl_sync(0xffffffffu, tid / kGT, 0), tl = tid % kGT;   // g through a shuffle: the compiler keeps it (and the descriptors built from it) in uniform registers
  const bool issuing_warp = __shfl_sync(0xffffffffu, (tid % kGT) >> 5, 0) == 0;
  constexpr uint32_t kCols = tmem_cols(G * WPT * 2 * KXP);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int q = 0; q < G * WPT; ++q) {
      mbar_init(&sm.bar[0][0] + q, 1);
      mbar_init(&sm.ready[0][0] + q, kGT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr MonicTab tab = make_monic_tab();
  // stacked B operand (WA rows over WB rows, hi and lo parts), prepared once per launch by tc_weight_prep_kernel in the final
  // swizzled layout: a straight copy of 2 KXP rows of 128 bytes each
  for (int item = tid; item < 2 * (2 * KXP * 8); item += blockDim.x) {
    const int part = item / (2 * KXP * 8), off = (item % (2 * KXP * 8)) * 16;
    const uint4 v = *reinterpret_cast<const uint4*>(wimg + part * kW2Bytes + off);
    sts128(smem_u32(sm.w[part]) + (uint32_t)off, v);
  }
  const float wscale = *reinterpret_cast<const float*>(wimg + 2 * kW2Bytes);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;           // this warp's TMEM lanes
  const uint32_t row_off = (uint32_t)(tl >> 3) * 1024u + (uint32_t)(tl & 7) * 128u;
  const uint64_t dWhi = umma_desc_k_sw128(smem_u32(sm.w[0])), dWlo = umma_desc_k_sw128(smem_u32(sm.w[1]));
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)((2 * KXP) >> 3) << 17) | ((128u >> 4) << 24);
  const float inv_w = 1.f / (kScaleP * wscale);                       // a power of two: exact

  // operand s of set k: tile LY::tile(s), starting LY::chunk0(s) 16-byte chunks into the 128-byte row
  uint64_t dsc[WPT][2];
#pragma unroll
  for (int k = 0; k < WPT; ++k)
#pragma unroll
    for (int q = 0; q < 2; ++q) dsc[k][q] = umma_desc_k_sw128(smem_u32(sm.a[g][k][0]) + LY::tile(q) * kTileBytes) + (uint64_t)LY::chunk0(q);

  // sets of 128 walkers per CTA (TcGrid, filled by launch_tc): CTAs [0, full) advance G * WPT sets each -- whole rounds of the
  // grid over the SMs --, CTAs [full, grid) `tail` sets: what is left over is spread evenly over the SMs, and the groups of such
  // a CTA that have no set skip the time loop (a round of one-group CTAs is latency bound and costs about half a full round,
  // instead of a whole round for a fraction of the SMs)
  const int b = (int)blockIdx.x;
  const int nsets = b < GR.full ? G * WPT : GR.tail;
  const int64_t set0 = b < GR.full ? (int64_t)b * (G * WPT) : (int64_t)GR.full * (G * WPT) + (int64_t)(b - GR.full) * GR.tail;
  int64_t wk[WPT];
  bool live[WPT];
  float x[WPT], y[WPT], vx[WPT], vy[WPT], zodd_x[WPT], zodd_y[WPT];
#pragma unroll
  for (int k = 0; k < WPT; ++k) {
    const int64_t idx = (set0 + g * WPT + k) * kGT + tl;
    live[k] = g * WPT + k < nsets && idx < A.n;
    wk[k] = live[k] ? idx : A.n - 1;
    x[k] = A.pos[wk[k]];
    y[k] = A.pos[A.n + wk[k]];
    vx[k] = A.vel[wk[k]];
    vy[k] = A.vel[A.n + wk[k]];
    zodd_x[k] = zodd_y[k] = 0.f;
  }
  auto row_addr = [&](uint32_t tiles, int sidx, int c) {   // 16-byte chunk c of operand sidx in this walker's row (Swizzle<3,4,3>: chunk ^ row % 8)
    return tiles + (uint32_t)(LY::tile(sidx) * kTileBytes) + row_off + (uint32_t)(((LY::chunk0(sidx) + c) ^ (tl & 7)) << 4);
  };

  uint32_t parity = 0;
  const uint32_t tmem_base = sm.tmem_base;                         // once: inside the loop it would be a generic load per step
  const int64_t nsteps = g * WPT < nsets ? A.nsteps : 0;           // a group without a set idles until the end of the kernel
  for (int64_t ts = 0; ts < nsteps; ++ts) {
    const int64_t t = A.step0 + ts;
    float sx[WPT], gx[WPT], gy[WPT], zx[WPT], zy[WPT];
    bool inx[WPT], iny[WPT];
    // ---- rows 1024 q_j(sy), j < KYP, as FP16 hi / lo, 8 functions per 128-bit store
    if constexpr (WPT == 2) {
      using Q = Pair<float>;
      sx[0] = scale_clamp(x[0], BP.cx, BP.ihx, inx[0]);
      sx[1] = scale_clamp(x[1], BP.cx, BP.ihx, inx[1]);
      const float sy0 = scale_clamp(y[0], BP.cy, BP.ihy, iny[0]), sy1 = scale_clamp(y[1], BP.cy, BP.ihy, iny[1]);
      const Q yy = Q::make(sy0, sy1), yy2 = Q::add(yy, yy);
      Q pm = Q::make(kScaleP, kScaleP), p = Q::mul_bcast(yy, kScaleP);
      const uint32_t tiles0 = smem_u32(sm.a[g][0][0]), tiles1 = smem_u32(sm.a[g][1][0]);
#pragma unroll
      for (int c = 0; c < KYP / 8; ++c) {
        float pv0[8], pv1[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = 8 * c + q;
          Q P;
          if (j == 0) {
            P = pm;
          } else if (j == 1) {
            P = p;
          } else {
            const Q pn = Q::fma(tab.d[j - 1] ? yy2 : yy, p, Q::mul_bcast(pm, -tab.cc[j - 1]));
            pm = p; p = pn;
            P = pn;
          }
          pv0[q] = P.lo();
          pv1[q] = P.hi();
        }
        uint4 hi, lo;
        split8(pv0, hi, lo);
        sts128(row_addr(tiles0, 0, c), hi);
        sts128(row_addr(tiles0, 1, c), lo);
        split8(pv1, hi, lo);
        sts128(row_addr(tiles1, 0, c), hi);
        sts128(row_addr(tiles1, 1, c), lo);
      }
    } else {
#pragma unroll
      for (int k = 0; k < WPT; ++k) {
        const uint32_t tiles = smem_u32(sm.a[g][k][0]);
        sx[k] = scale_clamp(x[k], BP.cx, BP.ihx, inx[k]);
        const float sy = scale_clamp(y[k], BP.cy, BP.ihy, iny[k]);
        const float sy2 = sy + sy;
        float pm = kScaleP, p = kScaleP * sy;              // 1024 q_0, 1024 q_1
#pragma unroll
        for (int c = 0; c < KYP / 8; ++c) {
          float pv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int j = 8 * c + q;
            if (j == 0) {
              pv[q] = pm;
            } else if (j == 1) {
              pv[q] = p;
            } else {
              const float pn = fmaf(tab.d[j - 1] ? sy2 : sy, p, -(tab.cc[j - 1] * pm));
              pm = p; p = pn;
              pv[q] = pn;
            }
          }
          uint4 hi, lo;
          split8(pv, hi, lo);
          sts128(row_addr(tiles, 0, c), hi);
          sts128(row_addr(tiles, 1, c), lo);
        }
      }
    }
    // rows of every walker set are written back to back, then ONE proxy fence; no block-wide barrier: every thread ARRIVES
    // (its rows are written, its TMEM reads of the last step are done) and goes on with the work that does not need the
    // contraction; only the issuing thread waits for the 128 arrivals
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#pragma unroll
    for (int k = 0; k < WPT; ++k) mbar_arrive(&sm.ready[g][k]);
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
      const uint32_t tmem = tmem_base + (uint32_t)((g * WPT + k) * 2 * KXP);   // A'_k in [0, KXP), B_i in [KXP, 2 KXP)
      if (issuing_warp) {                                            // warp-uniform branch; one elected lane issues
        mbar_wait(&sm.ready[g][k], parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t dPhi = dsc[k][0], dPlo = dsc[k][1];
#pragma unroll
          for (int kk = 0; kk < KYP / 16; ++kk) {                    // K = 16 per instruction: 32 bytes = +2 in 16-byte units
            const uint64_t o = (uint64_t)(kk * 2);
            umma_f16(tmem, dPhi + o, dWhi + o, kIdesc, kk > 0 ? 1u : 0u);
            umma_f16(tmem, dPhi + o, dWlo + o, kIdesc, 1u);
            umma_f16(tmem, dPlo + o, dWhi + o, kIdesc, 1u);
          }
          umma_commit(&sm.bar[g][k]);
        }
        __syncwarp();
      }
      // ---- independent of the contraction: potential gradient and noise (overlaps the MMA round trip)
      float e = 0.f;
      gx[k] = gy[k] = 0.f;
      POT::template grad<false>(PP, x[k], y[k], e, gx[k], gy[k]);
      if (INJECT) {
        zx[k] = A.noise[(ts * 2) * A.n + wk[k]];
        zy[k] = A.noise[(ts * 2 + 1) * A.n + wk[k]];
      } else if (!(t & 1) || ts == 0) {
        const uint4 r = philox_call(A.gid0 + (uint64_t)wk[k], (uint64_t)(t >> 1), kStreamDyn2, K);
        float z0, z1, z2, z3;
        box_muller(r.x, r.y, z0, z1);
        box_muller(r.z, r.w, z2, z3);
        if (t & 1) {
          zx[k] = z2;
          zy[k] = z3;
        } else {
          zx[k] = z0;
          zy[k] = z1;
          zodd_x[k] = z2;
          zodd_y[k] = z3;
        }
      } else {
        zx[k] = zodd_x[k];
        zy[k] = zodd_y[k];
      }
    }
    // ---- this walker's (A'_i, B_i) from TMEM, x recursion and the two row sums
    float Gx[WPT], Gy[WPT];
    if constexpr (WPT == 2) {
      using Q = Pair<float>;
      const uint32_t tmem0 = tmem_base + (uint32_t)((g * 2) * 2 * KXP) + tlane, tmem1 = tmem0 + (uint32_t)(2 * KXP);
      mbar_wait(&sm.bar[g][0], parity);
      mbar_wait(&sm.bar[g][1], parity);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float s0a = 0.f, s0b = 0.f, s1a = 0.f, s1b = 0.f;
      const Q xx = Q::make(sx[0], sx[1]), xx2 = Q::add(xx, xx);
      Q pxm = Q::make(1.f, 1.f), px = xx;
#pragma unroll
      for (int i0 = 0; i0 < KXP; i0 += 8) {
        float a0[8], b0[8], a1[8], b1[8];
        tmem_ld8(tmem0 + (uint32_t)i0, a0);
        tmem_ld8(tmem0 + (uint32_t)(KXP + i0), b0);
        tmem_ld8(tmem1 + (uint32_t)i0, a1);
        tmem_ld8(tmem1 + (uint32_t)(KXP + i0), b1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = i0 + q;
          Q P;
          if (i == 0) {
            P = pxm;
          } else if (i == 1) {
            P = px;
          } else {
            const Q pn = Q::fma(tab.d[i - 1] ? xx2 : xx, px, Q::mul_bcast(pxm, -tab.cc[i - 1]));
            pxm = px; px = pn;
            P = pn;
          }
          s0a = fmaf(P.lo(), a0[q], s0a);
          s0b = fmaf(P.lo(), b0[q], s0b);
          s1a = fmaf(P.hi(), a1[q], s1a);
          s1b = fmaf(P.hi(), b1[q], s1b);
        }
      }
      Gx[0] = s0a; Gy[0] = s0b; Gx[1] = s1a; Gy[1] = s1b;
    } else {
#pragma unroll
      for (int k = 0; k < WPT; ++k) {
        const uint32_t tmem = tmem_base + (uint32_t)((g * WPT + k) * 2 * KXP);
        float Gx0 = 0.f, Gx1 = 0.f, Gy0 = 0.f, Gy1 = 0.f;
        const float sx2 = sx[k] + sx[k];
        float pxm = 1.f, px = sx[k];
        if constexpr (KXP <= 16) {
          // the x recursion does not need the contraction: it runs in the shadow of the MMA round trip (16 registers), so that
          // only the loads, the two row sums and the integrator are left on the group's critical path after the commit
          float qx[KXP];
#pragma unroll
          for (int i = 0; i < KXP; ++i) {
            if (i == 0) {
              qx[i] = 1.f;
            } else if (i == 1) {
              qx[i] = px;
            } else {
              const float pn = fmaf(tab.d[i - 1] ? sx2 : sx[k], px, -(tab.cc[i - 1] * pxm));
              pxm = px; px = pn;
              qx[i] = pn;
            }
          }
          mbar_wait(&sm.bar[g][k], parity);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          float av[16], bv[16];
          tmem_ld16(tmem + tlane, av);
          tmem_ld16(tmem + tlane + (uint32_t)KXP, bv);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          Gx0 = av[0];                                               // q_0 = 1
          Gy0 = bv[0];
#pragma unroll
          for (int q = 1; q < 16; ++q) {
            if (q & 1) {
              Gx1 = fmaf(qx[q], av[q], Gx1);
              Gy1 = fmaf(qx[q], bv[q], Gy1);
            } else {
              Gx0 = fmaf(qx[q], av[q], Gx0);
              Gy0 = fmaf(qx[q], bv[q], Gy0);
            }
          }
        } else {
        mbar_wait(&sm.bar[g][k], parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int i0 = 0; i0 < KXP; i0 += 16) {
          float av[16], bv[16];
          tmem_ld16(tmem + tlane + (uint32_t)i0, av);
          tmem_ld16(tmem + tlane + (uint32_t)(KXP + i0), bv);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int i = i0 + q;
            float pi;
            if (i == 0) {
              pi = 1.f;
            } else if (i == 1) {
              pi = px;
            } else {
              const float pn = fmaf(tab.d[i - 1] ? sx2 : sx[k], px, -(tab.cc[i - 1] * pxm));
              pxm = px; px = pn;
              pi = pn;
            }
            if (q & 1) {
              Gx1 = fmaf(pi, av[q], Gx1);
              Gy1 = fmaf(pi, bv[q], Gy1);
            } else {
              Gx0 = fmaf(pi, av[q], Gx0);
              Gy0 = fmaf(pi, bv[q], Gy0);
            }
          }
        }
        }
        Gx[k] = Gx0 + Gx1;
        Gy[k] = Gy0 + Gy1;
      }
    }
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
      const float fx = gx[k] + (inx[k] ? Gx[k] * inv_w * BP.ihx : 0.f);
      const float fy = gy[k] + (iny[k] ? Gy[k] * inv_w * BP.ihy : 0.f);
      vx[k] = fmaf(I.c_over_sqrtm, zx[k], fmaf(-I.b_over_m, fx, I.a * vx[k]));
      vy[k] = fmaf(I.c_over_sqrtm, zy[k], fmaf(-I.b_over_m, fy, I.a * vy[k]));
      x[k] = fmaf(I.dt, vx[k], x[k]);
      y[k] = fmaf(I.dt, vy[k], y[k]);
    }
    parity ^= 1u;
  }
#pragma unroll
  for (int k = 0; k < WPT; ++k) {
    if (live[k]) {
      A.pos[wk[k]] = x[k];
      A.pos[A.n + wk[k]] = y[k];
      A.vel[wk[k]] = vx[k];
      A.vel[A.n + wk[k]] = vy[k];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem_base), "r"(kCols));
}

template <int KXP, int KYP, int G, int WPT, class POT = PotGeneric<float>>
cudaError_t launch_tc(const StepArgs<float>& A, const IntegratorParams<float>& I, const PotParams<float>& PP,
                      const Legendre2DSmemParams<float>& BP, const RngKeys& K, unsigned char* img, cudaStream_t st) {
  if (img == nullptr) return cudaErrorInvalidValue;
  auto kern = A.noise != nullptr ? langevin2d_tc_kernel<KXP, KYP, G, WPT, true, POT> : langevin2d_tc_kernel<KXP, KYP, G, WPT, false, POT>;
  const size_t smem = sizeof(TcSmem<KYP, G, WPT>) + 1024;
  static_assert(sizeof(TcSmem<KYP, G, WPT>) + 1024 <= 227 * 1024, "operand tiles do not fit the shared memory of an SM");
  static_assert(G * WPT * 2 * KXP <= 512, "accumulators do not fit the TMEM columns of an SM");
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // (letting the remainder ride as a fifth group on four-group CTAs was measured too: 5.27 against 5.09 ms, the larger block
  // lowers the register cap)
  int resident = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kGT * G, smem);
  TcGrid GR;
  const unsigned grid = tc_deal_sets(A.n, G * WPT, WPT, resident, GR);
  kern<<<grid, kGT * G, smem, st>>>(A, I, PP, BP, K, img, GR);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// NN bias (NNBasis1D with three hidden layers, folded as in bias.cuh: z1 = u s + c): the one H x H layer
//     z2_j = b2_j + sum_i W2[j][i] h_i,   dz2_j/ds = sum_i W2[j][i] dh_i/ds          (h_i = act(z1_i), dh_i = act'(z1_i) u_i)
// is the same [128 walkers x H] x [H x H] product per step: rows (h_i | dh_i) as FP16 hi + lo, B = W2 (row j = output unit),
// epilogue per thread: activation of z2_j, output layer, chain rule.  The vectors u, c, wout, b2 are read from shared memory
// as 128-bit broadcasts.  H = 32 / 48 / 64 (the padded width the host packs, pyves_set_bias_mlp).
template <int H, int G> struct TcMlpSmem {
  alignas(1024) unsigned char a[G][4][kTileBytes];   // per group: h_hi, h_lo, dh_hi, dh_lo
  alignas(1024) unsigned char w[2][kWBytes];         // W2_hi, W2_lo: [H j][H i] FP16, 128-byte rows
  alignas(16) float vec[4 * H + 4];                  // u, c, wout, b2, bout
  alignas(8) uint64_t bar[G];
  alignas(8) uint64_t ready[G];
  float red[32];
  uint32_t tmem_base;
};

template <int H, int ACT, int G, bool INJECT>
__global__ void __launch_bounds__(kGT* G, 1)
langevin2d_tc_mlp_kernel(const StepArgs<float> A, const IntegratorParams<float> I, const PotParams<float> PP,
                         const MlpParams<float> BP, const RngKeys K) {
  extern __shared__ unsigned char smem_raw[];
  TcMlpSmem<H, G>& sm = *reinterpret_cast<TcMlpSmem<H, G>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = __shfl_sync(0xffffffffu, tid / kGT, 0), tl = tid % kGT;   // g through a shuffle: the descriptors built from it stay in uniform registers
  const bool issuing_warp = __shfl_sync(0xffffffffu, (tid % kGT) >> 5, 0) == 0;
  constexpr uint32_t kCols = tmem_cols(G * 2 * H);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int q = 0; q < G; ++q) {
      mbar_init(&sm.bar[q], 1);
      mbar_init(&sm.ready[q], kGT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const float* W2T = BP.packed + 4 * H;              // W2T[i][j] = W2[j][i]
  for (int q = tid; q < 4 * H + 4; q += blockDim.x) sm.vec[q] = q < 4 * H ? BP.packed[q] : BP.packed[4 * H + H * H + (q - 4 * H)];
  float wscale;
  {
    float m = 0.f;
    for (int q = tid; q < H * H; q += blockDim.x) m = fmaxf(m, fabsf(W2T[q]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) sm.red[warp] = m;
    __syncthreads();
    m = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, sm.red[q]);
    int ex = 0;
    if (m > 0.f && m < 3e38f) frexpf(m, &ex);
    wscale = ldexpf(1.f, 13 - ex);
  }
  for (int item = tid; item < H * (H / 8); item += blockDim.x) {     // row j (N index), chunk c of 8 inputs i (K index)
    const int j = item / (H / 8), c = item % (H / 8);
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = W2T[(8 * c + q) * H + j] * wscale;
    uint4 hi, lo;
    split8(v, hi, lo);
    const uint32_t off = (uint32_t)(j >> 3) * 1024u + (uint32_t)(j & 7) * 128u + (uint32_t)((c ^ (j & 7)) << 4);
    sts128(smem_u32(sm.w[0]) + off, hi);
    sts128(smem_u32(sm.w[1]) + off, lo);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t row_off = (uint32_t)(tl >> 3) * 1024u + (uint32_t)(tl & 7) * 128u;
  const uint64_t dWhi = umma_desc_k_sw128(smem_u32(sm.w[0])), dWlo = umma_desc_k_sw128(smem_u32(sm.w[1]));
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(H >> 3) << 17) | ((128u >> 4) << 24);
  const float inv_w = 1.f / wscale;
  const uint32_t tiles = smem_u32(sm.a[g][0]);
  uint64_t dsc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) dsc[q] = umma_desc_k_sw128(tiles + q * kTileBytes);
  const uint32_t tmem = sm.tmem_base + (uint32_t)(g * 2 * H);       // D_z [0, H), D_dz [H, 2 H)
  const float* uvec = sm.vec;
  const float* cvec = sm.vec + H;
  const float* wout = sm.vec + 2 * H;
  const float* b2 = sm.vec + 3 * H;
  const float bout = sm.vec[4 * H];

  const int64_t idx = ((int64_t)blockIdx.x * G + g) * kGT + tl;
  const bool live = idx < A.n;
  const int64_t w = live ? idx : A.n - 1;
  float x = A.pos[w], y = A.pos[A.n + w], vx = A.vel[w], vy = A.vel[A.n + w];
  float zodd_x = 0.f, zodd_y = 0.f;
  uint32_t parity = 0;
  for (int64_t ts = 0; ts < A.nsteps; ++ts) {
    const int64_t t = A.step0 + ts;
    const float sc = ((BP.axis ? y : x) - BP.lo) * BP.inv_range;
    // ---- rows (h_i, dh_i/ds), i < H, as FP16 hi / lo
#pragma unroll
    for (int c = 0; c < H / 8; ++c) {
      float hv[8], dv[8];
#pragma unroll
      for (int q4 = 0; q4 < 2; ++q4) {
        float u4[4], c4[4];
        lds_vec(uvec + 8 * c + 4 * q4, u4);
        lds_vec(cvec + 8 * c + 4 * q4, c4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float d;
          hv[4 * q4 + q] = act_fn<float, ACT>(fmaf(u4[q], sc, c4[q]), d);
          dv[4 * q4 + q] = d * u4[q];
        }
      }
      uint4 hi, lo;
      const uint32_t off = row_off + (uint32_t)((c ^ (tl & 7)) << 4);
      split8(hv, hi, lo);
      sts128(tiles + off, hi);
      sts128(tiles + kTileBytes + off, lo);
      split8(dv, hi, lo);
      sts128(tiles + 2 * kTileBytes + off, hi);
      sts128(tiles + 3 * kTileBytes + off, lo);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(&sm.ready[g]);
    if (issuing_warp) {                                              // warp-uniform branch; one elected lane issues
      mbar_wait(&sm.ready[g], parity);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < H / 16; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);
          umma_f16(tmem, dsc[0] + o, dWhi + o, kIdesc, kk > 0 ? 1u : 0u);
          umma_f16(tmem, dsc[0] + o, dWlo + o, kIdesc, 1u);
          umma_f16(tmem, dsc[1] + o, dWhi + o, kIdesc, 1u);
          umma_f16(tmem + (uint32_t)H, dsc[2] + o, dWhi + o, kIdesc, kk > 0 ? 1u : 0u);
          umma_f16(tmem + (uint32_t)H, dsc[2] + o, dWlo + o, kIdesc, 1u);
          umma_f16(tmem + (uint32_t)H, dsc[3] + o, dWhi + o, kIdesc, 1u);
        }
        umma_commit(&sm.bar[g]);
      }
      __syncwarp();
    }
    float gx = 0.f, gy = 0.f, e = 0.f;
    PotGeneric<float>::template grad<false>(PP, x, y, e, gx, gy);
    float zx, zy;
    if (INJECT) {
      zx = A.noise[(ts * 2) * A.n + w];
      zy = A.noise[(ts * 2 + 1) * A.n + w];
    } else if (!(t & 1) || ts == 0) {
      const uint4 r = philox_call(A.gid0 + (uint64_t)w, (uint64_t)(t >> 1), kStreamDyn2, K);
      float z0, z1, z2, z3;
      box_muller(r.x, r.y, z0, z1);
      box_muller(r.z, r.w, z2, z3);
      if (t & 1) {
        zx = z2;
        zy = z3;
      } else {
        zx = z0;
        zy = z1;
        zodd_x = z2;
        zodd_y = z3;
      }
    } else {
      zx = zodd_x;
      zy = zodd_y;
    }
    // ---- epilogue: this walker's (z2_j, dz2_j/ds), activation, output layer
    mbar_wait(&sm.bar[g], parity);
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float dval0 = 0.f, dval1 = 0.f;
#pragma unroll
    for (int j0 = 0; j0 < H; j0 += 16) {
      float zv[16], dz[16];
      tmem_ld16(tmem + tlane + (uint32_t)j0, zv);
      tmem_ld16(tmem + tlane + (uint32_t)(H + j0), dz);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float b4[4], w4[4];
        lds_vec(b2 + j0 + 4 * q4, b4);
        lds_vec(wout + j0 + 4 * q4, w4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float d;
          (void)act_fn<float, ACT>(fmaf(zv[4 * q4 + q], inv_w, b4[q]), d);
          if (q & 1) dval1 = fmaf(w4[q] * d, dz[4 * q4 + q], dval1);
          else dval0 = fmaf(w4[q] * d, dz[4 * q4 + q], dval0);
        }
      }
    }
    (void)bout;
    const float gb = (dval0 + dval1) * inv_w * BP.inv_range;
    gx += BP.axis ? 0.f : gb;
    gy += BP.axis ? gb : 0.f;
    vx = fmaf(I.c_over_sqrtm, zx, fmaf(-I.b_over_m, gx, I.a * vx));
    vy = fmaf(I.c_over_sqrtm, zy, fmaf(-I.b_over_m, gy, I.a * vy));
    x = fmaf(I.dt, vx, x);
    y = fmaf(I.dt, vy, y);
  }
  if (live) {
    A.pos[w] = x;
    A.pos[A.n + w] = y;
    A.vel[w] = vx;
    A.vel[A.n + w] = vy;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem_base), "r"(kCols));
}

// ReLU specialisation.  With m_i = [z1_i > 0] (z1_i = u_i s + c_i) the hidden layer is h_i = m_i z1_i, dh_i/ds = m_i u_i, so
//     dz2_j/ds = sum_i (W2[j][i] u_i) m_i,     z2_j = b2_j + s dz2_j/ds + sum_i (W2[j][i] c_i) m_i :
// ONE operand row per walker, the 0 / 1 mask -- exact in FP16, no lo part, no split --, against the stacked matrices
// WU = W2 diag(u) over WC = W2 diag(c) (N = 2 H, hi + lo): two MMAs per K slab instead of six, a 16 KB tile per group
// instead of four (five groups per SM; the 512 TMEM columns are the limit), 2.5 instructions per hidden unit instead of ~9.
template <int H, int G> struct TcReluSmem {
  alignas(1024) unsigned char a[G][kTileBytes];      // per group: mask rows [128 walkers][H] FP16
  alignas(1024) unsigned char w[2][kW2Bytes];        // (WU over WC) hi, lo: [2 H rows][H i] FP16, 128-byte rows
  alignas(16) float vec[4 * H + 4];                  // u, c, wout, b2, bout
  alignas(8) uint64_t bar[G];
  alignas(8) uint64_t ready[G];
  float red[32];
  uint32_t tmem_base;
};

template <int H, int G, bool INJECT, class POT>
__global__ void __launch_bounds__(kGT* G, 1)
langevin2d_tc_relu_kernel(const StepArgs<float> A, const IntegratorParams<float> I, const PotParams<float> PP,
                          const MlpParams<float> BP, const RngKeys K, const TcGrid GR) {
  extern __shared__ unsigned char smem_raw[];
  TcReluSmem<H, G>& sm = *reinterpret_cast<TcReluSmem<H, G>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = __shfl_sync(0xffffffffu, tid / kGT, 0), tl = tid % kGT;   // g through a shuffle: the descriptors built from it stay in uniform registers
  const bool issuing_warp = __shfl_sync(0xffffffffu, (tid % kGT) >> 5, 0) == 0;
  constexpr uint32_t kCols = tmem_cols(G * 2 * H);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int q = 0; q < G; ++q) {
      mbar_init(&sm.bar[q], 1);
      mbar_init(&sm.ready[q], kGT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const float* W2T = BP.packed + 4 * H;              // W2T[i][j] = W2[j][i]
  for (int q = tid; q < 4 * H + 4; q += blockDim.x) sm.vec[q] = q < 4 * H ? BP.packed[q] : BP.packed[4 * H + H * H + (q - 4 * H)];
  float wscale;
  {
    float m = 0.f;
    for (int q = tid; q < H * H; q += blockDim.x) {
      const int i = q / H;
      const float w2 = W2T[q];
      m = fmaxf(m, fmaxf(fabsf(w2 * BP.packed[i]), fabsf(w2 * BP.packed[H + i])));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) sm.red[warp] = m;
    __syncthreads();                                 // also: sm.vec is complete
    m = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, sm.red[q]);
    int ex = 0;
    if (m > 0.f && m < 3e38f) frexpf(m, &ex);
    wscale = ldexpf(1.f, 13 - ex);
  }
  for (int item = tid; item < 2 * H * (H / 8); item += blockDim.x) {   // row r (N index), chunk c of 8 inputs i (K index)
    const int r = item / (H / 8), c = item % (H / 8);
    const int j = r < H ? r : r - H;
    const float* fac = sm.vec + (r < H ? 0 : H);     // u for the WU rows, c for the WC rows
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = (float)((double)W2T[(8 * c + q) * H + j] * (double)fac[8 * c + q]) * wscale;
    uint4 hi, lo;
    split8(v, hi, lo);
    const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    sts128(smem_u32(sm.w[0]) + off, hi);
    sts128(smem_u32(sm.w[1]) + off, lo);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t row_off = (uint32_t)(tl >> 3) * 1024u + (uint32_t)(tl & 7) * 128u;
  const uint64_t dWhi = umma_desc_k_sw128(smem_u32(sm.w[0])), dWlo = umma_desc_k_sw128(smem_u32(sm.w[1]));
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)((2 * H) >> 3) << 17) | ((128u >> 4) << 24);
  const float inv_w = 1.f / wscale;
  const uint32_t tile = smem_u32(sm.a[g]);
  const uint64_t dM = umma_desc_k_sw128(tile);
  const uint32_t tmem = sm.tmem_base + (uint32_t)(g * 2 * H);       // wscale dz2_j/ds in [0, H), wscale sum_i WC m_i in [H, 2 H)
  const float* uvec = sm.vec;
  const float* cvec = sm.vec + H;
  const float* wout = sm.vec + 2 * H;
  const float* b2 = sm.vec + 3 * H;

  // sets of 128 walkers per CTA as in langevin2d_tc_kernel: full CTAs in whole rounds, the remainder in small CTAs
  const int bI = (int)blockIdx.x;
  const int nsets = bI < GR.full ? G : GR.tail;
  const int64_t set0 = bI < GR.full ? (int64_t)bI * G : (int64_t)GR.full * G + (int64_t)(bI - GR.full) * GR.tail;
  const int64_t idx = (set0 + g) * kGT + tl;
  const bool live = g < nsets && idx < A.n;
  const int64_t w = live ? idx : A.n - 1;
  float x = A.pos[w], y = A.pos[A.n + w], vx = A.vel[w], vy = A.vel[A.n + w];
  float zodd_x = 0.f, zodd_y = 0.f;
  uint32_t parity = 0;
  const int64_t nsteps = g < nsets ? A.nsteps : 0;             // a group without a set idles until the end of the kernel
  for (int64_t ts = 0; ts < nsteps; ++ts) {
    const int64_t t = A.step0 + ts;
    const float sc = ((BP.axis ? y : x) - BP.lo) * BP.inv_range;
    // ---- mask row m_i = [u_i s + c_i > 0], i < H, as FP16 0 / 1
#pragma unroll
    for (int c = 0; c < H / 8; ++c) {
      uint32_t pk[4];
#pragma unroll
      for (int q4 = 0; q4 < 2; ++q4) {
        float u4[4], c4[4], m4[4];
        lds_vec(uvec + 8 * c + 4 * q4, u4);
        lds_vec(cvec + 8 * c + 4 * q4, c4);
#pragma unroll
        for (int q = 0; q < 4; ++q) m4[q] = fmaf(u4[q], sc, c4[q]) > 0.f ? 1.f : 0.f;
        const __half2 h0 = __floats2half2_rn(m4[0], m4[1]), h1 = __floats2half2_rn(m4[2], m4[3]);
        pk[2 * q4] = *reinterpret_cast<const uint32_t*>(&h0);
        pk[2 * q4 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
      }
      sts128(tile + row_off + (uint32_t)((c ^ (tl & 7)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(&sm.ready[g]);
    if (issuing_warp) {                                              // warp-uniform branch; one elected lane issues
      mbar_wait(&sm.ready[g], parity);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < H / 16; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);
          umma_f16(tmem, dM + o, dWhi + o, kIdesc, kk > 0 ? 1u : 0u);
          umma_f16(tmem, dM + o, dWlo + o, kIdesc, 1u);
        }
        umma_commit(&sm.bar[g]);
      }
      __syncwarp();
    }
    float gx = 0.f, gy = 0.f, e = 0.f;
    POT::template grad<false>(PP, x, y, e, gx, gy);
    float zx, zy;
    if (INJECT) {
      zx = A.noise[(ts * 2) * A.n + w];
      zy = A.noise[(ts * 2 + 1) * A.n + w];
    } else if (!(t & 1) || ts == 0) {
      const uint4 r = philox_call(A.gid0 + (uint64_t)w, (uint64_t)(t >> 1), kStreamDyn2, K);
      float z0, z1, z2, z3;
      box_muller(r.x, r.y, z0, z1);
      box_muller(r.z, r.w, z2, z3);
      if (t & 1) {
        zx = z2;
        zy = z3;
      } else {
        zx = z0;
        zy = z1;
        zodd_x = z2;
        zodd_y = z3;
      }
    } else {
      zx = zodd_x;
      zy = zodd_y;
    }
    // ---- epilogue: z2_j = b2_j + (s D_j + D_{H+j}) / wscale, second ReLU, output layer, chain rule
    mbar_wait(&sm.bar[g], parity);
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float dval0 = 0.f, dval1 = 0.f;
#pragma unroll
    for (int j0 = 0; j0 < H; j0 += 16) {
      float dz[16], zc[16];
      tmem_ld16(tmem + tlane + (uint32_t)j0, dz);
      tmem_ld16(tmem + tlane + (uint32_t)(H + j0), zc);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float b4[4], w4[4];
        lds_vec(b2 + j0 + 4 * q4, b4);
        lds_vec(wout + j0 + 4 * q4, w4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float z2 = fmaf(fmaf(sc, dz[4 * q4 + q], zc[4 * q4 + q]), inv_w, b4[q]);
          const float wd = z2 > 0.f ? w4[q] : 0.f;
          if (q & 1) dval1 = fmaf(wd, dz[4 * q4 + q], dval1);
          else dval0 = fmaf(wd, dz[4 * q4 + q], dval0);
        }
      }
    }
    const float gb = (dval0 + dval1) * inv_w * BP.inv_range;
    gx += BP.axis ? 0.f : gb;
    gy += BP.axis ? gb : 0.f;
    vx = fmaf(I.c_over_sqrtm, zx, fmaf(-I.b_over_m, gx, I.a * vx));
    vy = fmaf(I.c_over_sqrtm, zy, fmaf(-I.b_over_m, gy, I.a * vy));
    x = fmaf(I.dt, vx, x);
    y = fmaf(I.dt, vy, y);
  }
  if (live) {
    A.pos[w] = x;
    A.pos[A.n + w] = y;
    A.vel[w] = vx;
    A.vel[A.n + w] = vy;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem_base), "r"(kCols));
}

template <int H, int G, class POT>
cudaError_t launch_tc_relu_pot(const StepArgs<float>& A, const IntegratorParams<float>& I, const PotParams<float>& PP, const MlpParams<float>& BP,
                               const RngKeys& K, cudaStream_t st) {
  auto kern = A.noise != nullptr ? langevin2d_tc_relu_kernel<H, G, true, POT> : langevin2d_tc_relu_kernel<H, G, false, POT>;
  const size_t smem = sizeof(TcReluSmem<H, G>) + 1024;
  static_assert(sizeof(TcReluSmem<H, G>) + 1024 <= 227 * 1024, "operand tiles do not fit the shared memory of an SM");
  static_assert(G * 2 * H <= 512, "accumulators do not fit the TMEM columns of an SM");
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int resident = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kGT * G, smem);
  TcGrid GR;
  const unsigned grid = tc_deal_sets(A.n, G, 1, resident, GR);
  kern<<<grid, kGT * G, smem, st>>>(A, I, PP, BP, K, GR);
  return cudaGetLastError();
}
template <int H, int G>
cudaError_t launch_tc_relu(const StepArgs<float>& A, const IntegratorParams<float>& I, const PotParams<float>& PP, const MlpParams<float>& BP,
                           const RngKeys& K, cudaStream_t st) {
  // the reference's NN example (examples/szabo_berezhkovskii) without the runtime switch over the potentials
  if (PP.kind == PYVES_POT_SZABO_BEREZHKOVSKII) return launch_tc_relu_pot<H, G, PotSzaboBerezhkovskii<float>>(A, I, PP, BP, K, st);
  return launch_tc_relu_pot<H, G, PotGeneric<float>>(A, I, PP, BP, K, st);
}

template <int H, int ACT>
cudaError_t launch_tc_mlp(const StepArgs<float>& A, const IntegratorParams<float>& I, const PotParams<float>& PP, const MlpParams<float>& BP,
                          const RngKeys& K, cudaStream_t st) {
  constexpr int G = 3;
  auto kern = A.noise != nullptr ? langevin2d_tc_mlp_kernel<H, ACT, G, true> : langevin2d_tc_mlp_kernel<H, ACT, G, false>;
  const size_t smem = sizeof(TcMlpSmem<H, G>) + 1024;
  static_assert(sizeof(TcMlpSmem<H, G>) + 1024 <= 227 * 1024, "operand tiles do not fit the shared memory of an SM");
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t per_block = (int64_t)kGT * G;
  kern<<<(unsigned)((A.n + per_block - 1) / per_block), kGT * G, smem, st>>>(A, I, PP, BP, K);
  return cudaGetLastError();
}

template <int KXP>
cudaError_t launch_tc_ky(const StepArgs<float>& A, const IntegratorParams<float>& I, const PotParams<float>& PP,
                         const Legendre2DSmemParams<float>& BP, const RngKeys& K, unsigned char* img, cudaStream_t st) {
  // groups per CTA x walker sets per group, measured on B200 (10^6 walkers, profiles/r02_tc_step_shapes.md).  Since the MMA issue
  // path became short (descriptors in uniform registers, one elected lane) four groups at <= 128 registers step fastest per walker
  // at 16 x 16 (5.09 ms per 10^6 x 500 against 5.19 / 5.42 ms for five / six groups), three CTAs of two groups at 32 x 32; one
  // walker set per thread beats two (packed FP32x2 recursions do not pay: FFMA2 holds the FMA pipe two cycles)
  static const int shape = getenv("PYVES_TC_SHAPE") ? atoi(getenv("PYVES_TC_SHAPE")) : 0;     // experiment knob: G * 10 + WPT
  if (BP.ky <= 16) {
    if (KXP == 16) {
      switch (shape) {
        case 21: return launch_tc<16, 16, 2, 1>(A, I, PP, BP, K, img, st);
        case 31: return launch_tc<16, 16, 3, 1>(A, I, PP, BP, K, img, st);
        case 41: return launch_tc<16, 16, 4, 1>(A, I, PP, BP, K, img, st);
        case 42: return launch_tc<16, 16, 4, 2>(A, I, PP, BP, K, img, st);
        case 51: return launch_tc<16, 16, 5, 1>(A, I, PP, BP, K, img, st);
        case 71: return launch_tc<16, 16, 7, 1>(A, I, PP, BP, K, img, st);
        case 81: return launch_tc<16, 16, 8, 1>(A, I, PP, BP, K, img, st);
        case 62: return launch_tc<16, 16, 6, 2>(A, I, PP, BP, K, img, st);
        case 61: return launch_tc<16, 16, 6, 1>(A, I, PP, BP, K, img, st);
        default:
          // the double well (every BASELINE 2D configuration) without the runtime switch over the potentials
          if (PP.kind == PYVES_POT_DOUBLE_WELL_2D) return launch_tc<16, 16, 4, 1, PotDoubleWell2D<float>>(A, I, PP, BP, K, img, st);
          return launch_tc<16, 16, 4, 1>(A, I, PP, BP, K, img, st);
      }
    }
    return launch_tc<KXP, 16, 4, (KXP <= 32 ? 2 : 1)>(A, I, PP, BP, K, img, st);
  }
  if (BP.ky <= 32) {
    if (KXP == 32) {
      switch (shape) {
        case 21: return launch_tc<32, 32, 2, 1>(A, I, PP, BP, K, img, st);
        case 31: return launch_tc<32, 32, 3, 1>(A, I, PP, BP, K, img, st);
        case 41: return launch_tc<32, 32, 4, 1>(A, I, PP, BP, K, img, st);
        case 51: return launch_tc<32, 32, 5, 1>(A, I, PP, BP, K, img, st);
        case 71: return launch_tc<32, 32, 7, 1>(A, I, PP, BP, K, img, st);
        case 81: return launch_tc<32, 32, 8, 1>(A, I, PP, BP, K, img, st);
        case 42: return launch_tc<32, 32, 4, 2>(A, I, PP, BP, K, img, st);
        case 61: return launch_tc<32, 32, 6, 1>(A, I, PP, BP, K, img, st);
        default:
          if (PP.kind == PYVES_POT_DOUBLE_WELL_2D) return launch_tc<32, 32, 2, 1, PotDoubleWell2D<float>>(A, I, PP, BP, K, img, st);
          return launch_tc<32, 32, 2, 1>(A, I, PP, BP, K, img, st);
      }
    }
    return launch_tc<KXP, 32, 4, (KXP <= 32 ? 2 : 1)>(A, I, PP, BP, K, img, st);
  }
  if (KXP == 64 && shape == 31) return launch_tc<KXP, 64, 3, 1>(A, I, PP, BP, K, img, st);
  if (KXP == 64 && shape == 21) return launch_tc<KXP, 64, 2, 1>(A, I, PP, BP, K, img, st);
  if (KXP == 64 && PP.kind == PYVES_POT_DOUBLE_WELL_2D) return launch_tc<KXP, 64, 4, 1, PotDoubleWell2D<float>>(A, I, PP, BP, K, img, st);
  return launch_tc<KXP, 64, 4, 1>(A, I, PP, BP, K, img, st);
}

}  // namespace

size_t langevin2d_tc_scratch_bytes() { return 2 * kW2Bytes + 256; }

// mirrors launch_tc_ky / launch_langevin2d_tc_mlp (default shapes)
int langevin2d_tc_walkers_per_cta(int kx, int ky) {       // per SM: walkers per CTA x resident CTAs of the default shapes
  if (kx > 64 || ky > 64 || kx < 1 || ky < 1) return 0;
  if (ky > 32) return kGT * 4;
  if (kx <= 16 && ky <= 16) return kGT * 4;               // <16,16,4,1>, one CTA per SM
  if (kx > 16 && kx <= 32 && ky > 16) return kGT * 2 * 3; // <32,32,2,1>, three CTAs per SM
  return kGT * 4 * (kx <= 32 ? 2 : 1);
}
int langevin2d_tc_mlp_walkers_per_cta(int H, int act) {
  const bool relu = act == PYVES_ACT_RELU;
  switch (H) {
    case 32: return kGT * (relu ? 8 : 3);
    case 48: return kGT * (relu ? 5 : 3);
    case 64: return kGT * (relu ? 4 : 3);
    default: return 0;
  }
}

// the operand image of the folded coefficient matrices (see tc_weight_prep_kernel), rebuilt whenever the coefficients change
cudaError_t langevin2d_tc_prepare(const Legendre2DSmemParams<float>& BP, void* scratch, cudaStream_t st) {
  if (BP.kx > 64 || BP.ky > 64 || BP.kx < 1 || BP.ky < 1 || scratch == nullptr) return cudaErrorNotSupported;
  const int KXP = BP.kx <= 16 ? 16 : BP.kx <= 32 ? 32 : 64;
  tc_weight_prep_kernel<<<1, 1024, 0, st>>>(BP, KXP, static_cast<unsigned char*>(scratch));
  return cudaGetLastError();
}

cudaError_t launch_langevin2d_tc(const StepArgs<float>& A, const IntegratorParams<float>& I, const PotParams<float>& PP,
                                 const Legendre2DSmemParams<float>& BP, const RngKeys& K, void* scratch, cudaStream_t st) {
  if (BP.kx > 64 || BP.ky > 64 || BP.kx < 1 || BP.ky < 1 || A.traj != nullptr || scratch == nullptr) return cudaErrorNotSupported;
  unsigned char* img = static_cast<unsigned char*>(scratch);
  if (BP.kx <= 16) return launch_tc_ky<16>(A, I, PP, BP, K, img, st);
  if (BP.kx <= 32) return launch_tc_ky<32>(A, I, PP, BP, K, img, st);
  return launch_tc_ky<64>(A, I, PP, BP, K, img, st);
}

// NN bias with three hidden layers packed at width H = 32 / 48 / 64 (MlpParams::packed of pyves_set_bias_mlp[_device])
cudaError_t launch_langevin2d_tc_mlp(const StepArgs<float>& A, const IntegratorParams<float>& I, const PotParams<float>& PP,
                                     const MlpParams<float>& BP, int H, int act, const RngKeys& K, cudaStream_t st) {
  if (A.traj != nullptr || BP.linear) return cudaErrorNotSupported;
  const bool relu = act == PYVES_ACT_RELU;
  static const int general = getenv("PYVES_TC_MLP_GENERAL") ? atoi(getenv("PYVES_TC_MLP_GENERAL")) : 0;   // experiment knob: 1 = no ReLU specialisation
  static const int rg = getenv("PYVES_TC_RELU_GROUPS") ? atoi(getenv("PYVES_TC_RELU_GROUPS")) : 0;        // experiment knob
  if (relu && !general) {
    switch (H) {
      case 32: return rg == 4 ? launch_tc_relu<32, 4>(A, I, PP, BP, K, st) : rg == 6 ? launch_tc_relu<32, 6>(A, I, PP, BP, K, st) : launch_tc_relu<32, 8>(A, I, PP, BP, K, st);
      case 48: return rg == 3 ? launch_tc_relu<48, 3>(A, I, PP, BP, K, st) : rg == 4 ? launch_tc_relu<48, 4>(A, I, PP, BP, K, st) : launch_tc_relu<48, 5>(A, I, PP, BP, K, st);
      case 64: return rg == 3 ? launch_tc_relu<64, 3>(A, I, PP, BP, K, st) : launch_tc_relu<64, 4>(A, I, PP, BP, K, st);
      default: return cudaErrorNotSupported;
    }
  }
  switch (H) {
    case 32: return relu ? launch_tc_mlp<32, PYVES_ACT_RELU>(A, I, PP, BP, K, st) : launch_tc_mlp<32, PYVES_ACT_TANH>(A, I, PP, BP, K, st);
    case 48: return relu ? launch_tc_mlp<48, PYVES_ACT_RELU>(A, I, PP, BP, K, st) : launch_tc_mlp<48, PYVES_ACT_TANH>(A, I, PP, BP, K, st);
    case 64: return relu ? launch_tc_mlp<64, PYVES_ACT_RELU>(A, I, PP, BP, K, st) : launch_tc_mlp<64, PYVES_ACT_TANH>(A, I, PP, BP, K, st);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace pv
