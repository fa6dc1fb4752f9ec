// Micro-benchmark 2: the 16-accumulator column sweep  AB[i] += w[j][i] * Q[j]  with the weights coming
// from (0) kernel params = constant bank (compiler picks UR / LDC), (1) shared memory LDS.128.
#include <cstdio>
#include <cuda_runtime.h>
struct W { float w[256]; };
struct W2 { unsigned long long w[256]; };
__device__ __forceinline__ void f2p(unsigned long long& acc, unsigned long long q, unsigned long long ww) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(q), "l"(ww));
}
template <int NQ> __global__ void __launch_bounds__(128) kp(float* out, const W2 P, float a, int iters) {
  unsigned long long AB[16], Q[NQ];
  for (int i = 0; i < 16; ++i) AB[i] = i;
  for (int j = 0; j < NQ; ++j) asm("mov.b64 %0, {%1, %2};" : "=l"(Q[j]) : "f"(a + j), "f"(a * j));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int i = 0; i < 16; ++i) f2p(AB[i], Q[j % NQ], P.w[j * 16 + i]);
  }
  unsigned long long s = 0;
  for (int i = 0; i < 16; ++i) s += AB[i];
  if (s == 12345) out[0] = 1.f;
}
__device__ __forceinline__ void f2b(unsigned long long& acc, unsigned long long q, float w) {
  unsigned long long ww;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(q), "l"(ww));
}
template <int SRC, int NQ> __global__ void __launch_bounds__(128) k(float* out, const W P, float a, int iters) {
  __shared__ __align__(16) float sw[SRC == 2 ? 512 : 256];
  if (SRC == 2) { for (int i = threadIdx.x; i < 256; i += 128) { sw[2 * i] = P.w[i]; sw[2 * i + 1] = P.w[i]; } }
  else { for (int i = threadIdx.x; i < 256; i += 128) sw[i] = P.w[i]; }
  __syncthreads();
  // SRC 3 / 4: loop-invariant weights in registers, as pairs (w, w) / as scalars broadcast by the FFMA2
  unsigned long long WP[16];
  float WS[16];
  for (int i = 0; i < 16; ++i) { WS[i] = P.w[i] * a; asm("mov.b64 %0, {%1, %2};" : "=l"(WP[i]) : "f"(WS[i]), "f"(WS[i] + a)); }
  unsigned long long AB[16], Q[NQ];
  for (int i = 0; i < 16; ++i) AB[i] = i;
  for (int j = 0; j < NQ; ++j) asm("mov.b64 %0, {%1, %2};" : "=l"(Q[j]) : "f"(a + j), "f"(a * j));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (SRC == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f2b(AB[i], Q[j % NQ], P.w[j * 16 + i]);
      } else if (SRC == 3) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f2p(AB[i], Q[j % NQ], WP[i]);
      } else if (SRC == 4) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f2b(AB[i], Q[j % NQ], WS[i]);
      } else if (SRC == 2) {
#pragma unroll
        for (int iv = 0; iv < 8; ++iv) {
          unsigned long long w0, w1;
          const unsigned ad = (unsigned)__cvta_generic_to_shared(&sw[2 * (j * 16 + iv * 2)]);
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(ad));
          f2p(AB[iv * 2 + 0], Q[j % NQ], w0); f2p(AB[iv * 2 + 1], Q[j % NQ], w1);
        }
      } else {
#pragma unroll
        for (int iv = 0; iv < 4; ++iv) {
          float4 w;
          const unsigned ad = (unsigned)__cvta_generic_to_shared(&sw[j * 16 + iv * 4]);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "r"(ad));
          f2b(AB[iv * 4 + 0], Q[j % NQ], w.x); f2b(AB[iv * 4 + 1], Q[j % NQ], w.y);
          f2b(AB[iv * 4 + 2], Q[j % NQ], w.z); f2b(AB[iv * 4 + 3], Q[j % NQ], w.w);
        }
      }
    }
  }
  unsigned long long s = 0;
  for (int i = 0; i < 16; ++i) s += AB[i];
  if (s == 12345) out[0] = 1.f;
}
template <int SRC, int NQ> void run(const char* name, int bps) {
  float* out; cudaMalloc(&out, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  W P; for (int i = 0; i < 256; ++i) P.w[i] = 1e-3f * i;
  const int iters = 2000, blocks = sms * bps, threads = 128;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k<SRC, NQ><<<blocks, threads>>>(out, P, 0.5f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double warps_per_sched = (double)bps;   // 128 threads = 4 warps = 1 per scheduler per block
  const double cyc = best * 1e-3 * 1.965e9 / (iters * warps_per_sched);
  printf("%-24s blocks/SM %d: %7.3f ms  cycles per 256-FFMA2 sweep per warp %.1f (ideal 512) -> %.3f\n", name, bps, best, cyc, 512 / cyc);
  cudaFree(out);
}
template <int NQ> void runp(const char* name, int bps) {
  float* out; cudaMalloc(&out, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  W2 P; for (int i = 0; i < 256; ++i) P.w[i] = 0x3a83126f3a83126full + i;
  const int iters = 2000, blocks = sms * bps, threads = 128;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    kp<NQ><<<blocks, threads>>>(out, P, 0.5f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double cyc = best * 1e-3 * 1.965e9 / (iters * (double)bps);
  printf("%-24s blocks/SM %d: %7.3f ms  cycles per 256-FFMA2 sweep per warp %.1f (ideal 512) -> %.3f\n", name, bps, best, cyc, 512 / cyc);
  cudaFree(out);
}
int main() {
  run<3, 16>("regs pair, 16 Q", 2);
  run<4, 16>("regs bcast, 16 Q", 2);
  run<3, 1>("regs pair, 1 Q", 2);
  run<4, 1>("regs bcast, 1 Q", 2);
  run<2, 16>("LDS.128 dup, 16 Q", 2);
  run<2, 16>("LDS.128 dup, 16 Q", 4);
  run<2, 1>("LDS.128 dup, 1 Q", 2);
  runp<16>("pair weights const, 16 Q", 2);
  runp<16>("pair weights const, 16 Q", 4);
  run<0, 16>("const-bank, 16 Q", 2);
  run<0, 16>("const-bank, 16 Q", 4);
  run<0, 1>("const-bank, 1 Q", 2);
  run<0, 1>("const-bank, 1 Q", 4);
  run<1, 16>("LDS.128, 16 Q", 2);
  run<1, 16>("LDS.128, 16 Q", 4);
  run<1, 1>("LDS.128, 1 Q", 4);
  return 0;
}
