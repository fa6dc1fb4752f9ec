// Basis covariance  C = sum_rows f f^T  on the 5th-generation tensor cores (tcgen05, TF32 in,
// FP32 accumulate in TMEM) -- the one GEMM-shaped step of the path (textbook VES Hessian
// beta Cov_V(f); the reference only names the option, README.md:24).
//
// One CTA per SM.  Per trip a CTA turns 32 walkers into a [256 basis functions] x [32 walkers] TF32
// tile written straight into shared memory in the canonical K-major SWIZZLE_128B UMMA layout (the
// basis matrix B[N, K] never exists in HBM), then one elected thread issues
//     D0[128 x 256] += F[0:128, :] F^T        D1[128 x 256] += F[128:256, :] F^T
// as 2 x 4 tcgen05.mma (M = 128, N = 256, K = 8) reading the SAME tile as A and B operand.  The two
// accumulators fill the CTA's 512 TMEM columns.  Two shared-memory stages: the tensor core consumes
// stage s while all threads generate stage s^1; completion is tracked with tcgen05.commit ->
// mbarrier.  At the end the accumulators are read back with tcgen05.ld and written as per-CTA FP32
// partials that a second kernel sums in fixed order into the FP64 result.
#include "moments.cuh"

namespace pv {

namespace {

constexpr int kTcThreads = 256;
constexpr int kTcRows = 32;                 // walkers per tile = 128 bytes of TF32 = one swizzle atom
constexpr int kTcK = 256;                   // padded number of basis functions (UMMA N, 2 x UMMA M)
constexpr int kTileBytes = kTcK * kTcRows * 4;   // 32 KB
constexpr int kTabStride = 33;

PV_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

PV_DEV uint64_t umma_desc_k_sw128(uint32_t saddr) {
  // start address [0,14) (>>4) | LBO [16,30) = 1 (unused for swizzled K-major) | SBO [32,46) = 1024 B >> 4
  // | version [46,48) = 1 (Blackwell) | layout type [61,64) = 2 (SWIZZLE_128B)
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcK >> 3) << 17) | ((128u >> 4) << 24);

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
PV_DEV void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
PV_DEV void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
PV_DEV uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

struct CovSmem {
  alignas(1024) unsigned char tile[2][kTileBytes];
  float tab[kTcRows * kTabStride];
  alignas(8) uint64_t stage_bar[2];
  alignas(8) uint64_t done_bar;
  uint32_t tmem_base;
};

// Legendre tables of one row: tab[r][a] = P_a(sx) (a < ka), tab[r][ka + b] = P_b(sy) (b < kb)
PV_DEV void row_table(const RowSource<float>& S, const BasisParams<float>& B, int64_t r, bool live, float* t) {
  if (!live) {
    for (int a = 0; a < B.ka + B.kb; ++a) t[a] = 0.f;
    return;
  }
  const float x = S.pos[r * S.stride_row];
  const float y = S.pos[r * S.stride_row + S.stride_col];
  bool in_a, in_b;
  const float s[2] = {scale_clamp(x, B.ca, B.iha, in_a), scale_clamp(y, B.cb, B.ihb, in_b)};
  if (B.inside_only && !(in_a && in_b)) {   // weight 0: the row contributes nothing to B^T B
    for (int a = 0; a < B.ka + B.kb; ++a) t[a] = 0.f;
    return;
  }
  const int k[2] = {B.ka, B.kb};
  for (int ax = 0; ax < 2; ++ax) {
    float pm = 1.f, p = s[ax];
    t[0] = 1.f;
    if (k[ax] > 1) t[1] = p;
    for (int n = 1; n + 1 < k[ax]; ++n) {
      const float fn = (float)n;
      const float pn = ((2.f * fn + 1.f) * s[ax] * p - fn * pm) / (fn + 1.f);
      t[n + 1] = pn;
      pm = p;
      p = pn;
    }
    t += k[ax];
  }
}

__global__ void __launch_bounds__(kTcThreads, 1)
cov_tf32_kernel(const RowSource<float> S, const BasisParams<float> B, float* __restrict__ part) {
  extern __shared__ unsigned char smem_raw[];
  CovSmem& sm = *reinterpret_cast<CovSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = B.ka * B.kb;

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

  // this thread owns basis function k = tid: row tid of the [256][32] tile
  const int k = tid;
  const bool kvalid = k < K;
  const int fa = kvalid ? k / B.kb : 0, fb = kvalid ? B.ka + (k - (k / B.kb) * B.kb) : 0;
  const uint32_t row_off = (uint32_t)(k >> 3) * 1024u + (uint32_t)(k & 7) * 128u;

  const int64_t ntiles = (S.n + kTcRows - 1) / kTcRows;
  int trip = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++trip) {
    const int s = trip & 1;
    const int use = trip >> 1;
    if (use >= 1) mbar_wait(&sm.stage_bar[s], (uint32_t)((use - 1) & 1));   // the MMAs that read this stage are done
    // per-row Legendre tables (one warp), then every thread fills its own swizzled 128-byte row
    if (warp == 1) {
      const int64_t r = tile * kTcRows + lane;
      row_table(S, B, r, r < S.n, sm.tab + lane * kTabStride);
    }
    __syncthreads();
    {
      unsigned char* base = sm.tile[s] + row_off;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 v;
        uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = c * 4 + q;
          const float f = kvalid ? sm.tab[r * kTabStride + fa] * sm.tab[r * kTabStride + fb] : 0.f;
          pv[q] = to_tf32(f);
        }
        *reinterpret_cast<uint4*>(base + ((c ^ (k & 7)) << 4)) = v;   // Swizzle<3,4,3>: 16-byte chunk index ^ (row % 8)
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(sm.tile[s]);
      const uint64_t db = umma_desc_k_sw128(a0);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const uint64_t da = umma_desc_k_sw128(a0 + (uint32_t)m * 128u * 128u);   // rows 128 m .. 128 m + 127
#pragma unroll
        for (int kk = 0; kk < kTcRows / 8; ++kk)       // K = 8 TF32 = 32 bytes per instruction: +2 in 16-byte units
          umma_tf32(tmem + (uint32_t)m * 256u, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), (trip > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(&sm.stage_bar[s]);
    }
  }
  // all MMAs done?
  if (tid == 0) umma_commit(&sm.done_bar);
  mbar_wait(&sm.done_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // TMEM -> registers -> per-CTA partial [256][256] floats.  Warp w reads lanes 32 (w % 4) .. +31 of
  // accumulator w / 4; thread = one row of D.
  if (trip > 0) {
    const int acc = warp >> 2, lq = warp & 3;
    float* out = part + ((size_t)blockIdx.x * kTcK + (size_t)(acc * 128 + lq * 32 + lane)) * kTcK;
#pragma unroll 1
    for (int c0 = 0; c0 < kTcK; c0 += 32) {
      uint32_t v[32];
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
#pragma unroll
      for (int q = 0; q < 32; q += 4)
        *reinterpret_cast<uint4*>(out + c0 + q) = make_uint4(v[q], v[q + 1], v[q + 2], v[q + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// out[k][l] (FP64, K x K) = sum over CTAs of part[cta][k][l]; symmetrised exactly
__global__ void cov_tc_reduce_kernel(const float* __restrict__ part, int G, int K, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * K) return;
  const int k = idx / K, l = idx - k * K;
  if (l < k) return;
  double s = 0.0;
  for (int g = 0; g < G; ++g)
    s += 0.5 * ((double)part[((size_t)g * kTcK + k) * kTcK + l] + (double)part[((size_t)g * kTcK + l) * kTcK + k]);
  out[(size_t)k * K + l] = s;
  out[(size_t)l * K + k] = s;
}

}  // namespace

bool cov_tc_supported(const RowSource<float>& S, const BasisParams<float>& B) {
  const int K = B.ka * B.kb;
  return S.w == nullptr && B.kind == PYVES_BIAS_LEGENDRE_2D && K > 128 && K <= kTcK && S.ncols >= 2;
}

size_t cov_tc_scratch_floats(int n_sm) { return (size_t)n_sm * kTcK * kTcK; }

cudaError_t launch_cov_tc(const RowSource<float>& S, const BasisParams<float>& B, double* sum_wff, float* scratch, int n_sm,
                          cudaStream_t st, int* launches) {
  const int K = B.ka * B.kb;
  const int64_t ntiles = (S.n + kTcRows - 1) / kTcRows;
  const int G = (int)(ntiles < n_sm ? ntiles : n_sm);
  const size_t smem = sizeof(CovSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(cov_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cov_tf32_kernel<<<G, kTcThreads, smem, st>>>(S, B, scratch);
  cov_tc_reduce_kernel<<<(K * K + 255) / 256, 256, 0, st>>>(scratch, G, K, sum_wff);
  *launches += 2;
  return cudaGetLastError();
}

}  // namespace pv
