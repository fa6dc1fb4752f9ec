// Micro-benchmark: TMEM read throughput of tcgen05.ld.32x32b.x16 with W warps per SM (all four lane quadrants used).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tmem_bw tools/exp/tmem_bw.cu ; run: gpurun_out/tmem_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) k(int iters, int unroll_cols, float* out, long long* cycles) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 16 % 512);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr + (uint32_t)((it & 7) * 16 % unroll_cols)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(r[0] & 0x3f800000u) + __uint_as_float(r[15] & 0x3f800000u);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 20000;
  for (int warps : {1, 4, 8, 16, 24, 32}) {
    k<<<148, warps * 32>>>(iters, 128, out, cyc);
    cudaDeviceSynchronize();
    k<<<148, warps * 32>>>(iters, 128, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    const double bytes = (double)warps * iters * 32 * 16 * 4;
    printf("warps %2d: %.0f cycles, %.1f B/cycle/SM, %.1f cycles per LDTM.x16 per warp (%s)\n", warps, c, bytes / c, c / iters, cudaGetErrorString(e));
  }
  return 0;
}
