// Micro-benchmark: throughput of FFMA2 mixed with scalar FFMA in different patterns (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#define F2(acc, a, b) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b))
#define F1(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))
#define M1(acc, a) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(acc) : "f"(a))
#define I1(acc, a) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(acc) : "r"(a))

template <int PAT> __global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  unsigned long long p[8], A, B;
  float s[8];
  unsigned u[4] = {1, 2, 3, 4};
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b), "f"(b));
  for (int i = 0; i < 8; ++i) { p[i] = A + i; s[i] = a + i; }
  unsigned ua = __float_as_uint(a);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (PAT == 0) { F2(p[r], A, B); F2(p[(r + 4) & 7], A, B); }                       // 2 packed
      if (PAT == 1) { F1(s[r], a, b); F1(s[(r + 4) & 7], a, b); F1(s[(r + 2) & 7], a, b); F1(s[(r + 6) & 7], a, b); }  // 4 scalar
      if (PAT == 2) { F2(p[r], A, B); F1(s[r], a, b); }                                  // P S
      if (PAT == 3) { F2(p[r], A, B); F1(s[r], a, b); F1(s[(r + 4) & 7], a, b); }         // P S S
      if (PAT == 4) { F2(p[r], A, B); F2(p[(r + 4) & 7], A, B); F1(s[r], a, b); F1(s[(r + 4) & 7], a, b); }   // P P S S
      if (PAT == 5) { F2(p[r], A, B); F2(p[(r + 4) & 7], A, B); F1(s[r], a, b); }         // P P S
      if (PAT == 6) { F2(p[r], A, B); F2(p[(r + 4) & 7], A, B); F2(p[(r + 2) & 7], A, B); F1(s[r], a, b); }   // P P P S
      if (PAT == 7) { F2(p[r], A, B); I1(u[r & 3], ua); }                                 // P + ALU op
      if (PAT == 8) { F2(p[r], A, B); I1(u[r & 3], ua); I1(u[(r + 2) & 3], ua); }          // P + 2 ALU ops
      if (PAT == 9) { F1(s[r], a, b); I1(u[r & 3], ua); }                                 // S + ALU
    }
  }
  float acc = 0;
  for (int i = 0; i < 8; ++i) { acc += s[i] + (float)(p[i] & 0xffff); }
  acc += u[0] + u[1] + u[2] + u[3];
  if (acc == 123.456f) out[0] = acc;
}

template <int PAT> void run(const char* name, double fma_lanes_per_iter, int ninstr) {
  float* out; cudaMalloc(&out, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2048, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k<PAT><<<blocks, threads>>>(out, 0.999f, 1e-3f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  // warp-instructions per scheduler: blocks*threads/32 warps / (sms*4) schedulers
  const double warps_per_sched = (double)blocks * threads / 32 / (sms * 4.0);
  const double groups = (double)iters * 8 * warps_per_sched;
  const double cycles = best * 1e-3 * 1.965e9;
  printf("%-10s %7.3f ms  cycles per group %.2f  (instr per group %d, full-width FMA cycles if ideal %.1f)\n", name, best, cycles / groups, ninstr,
         fma_lanes_per_iter);
  cudaFree(out);
}

int main() {
  run<0>("PP", 4, 2);
  run<1>("SSSS", 4, 4);
  run<2>("PS", 3, 2);
  run<3>("PSS", 4, 3);
  run<4>("PPSS", 6, 4);
  run<5>("PPS", 5, 3);
  run<6>("PPPS", 7, 4);
  run<7>("P+alu", 2, 2);
  run<8>("P+2alu", 2, 3);
  run<9>("S+alu", 1, 2);
  return 0;
}
