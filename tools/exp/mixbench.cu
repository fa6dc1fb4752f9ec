// Micro-benchmark: throughput of FFMA2 mixed with scalar FFMA in different patterns (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#define F2(acc, a, b) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b))
#define F1(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))
#define M1(acc, a) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(acc) : "f"(a))
#define I1(acc, a) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(acc) : "r"(a))
#define W1(acc, a) asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %0; xor.b32 lo, lo, hi; mul.wide.u32 %0, lo, 0xD2511F53; }" : "+l"(acc))

template <int PAT> __global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  unsigned long long p[8], A, B;
  float s[8];
  unsigned u[4] = {threadIdx.x + 1, threadIdx.x * 3 + 2, threadIdx.x ^ 0x55, threadIdx.x + 77};
  unsigned long long wq[8];
  for (int i = 0; i < 8; ++i) wq[i] = (unsigned long long)(threadIdx.x + i) * 0x9E3779B97F4A7C15ull + __float_as_uint(a);
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
      if (PAT == 10) { W1(wq[r], 0); }                                          // IMAD.WIDE alone
      if (PAT == 11) { F2(p[r], A, B); W1(wq[r], 0); }                          // P + W
      if (PAT == 12) { F1(s[r], a, b); W1(wq[r], 0); }                          // S + W
      if (PAT == 13) { F1(s[r], a, b); F1(s[(r + 4) & 7], a, b); W1(wq[r], 0); }   // S S + W
      if (PAT == 14) { F1(s[r], a, b); F1(s[(r + 4) & 7], a, b); F1(s[(r + 2) & 7], a, b); F1(s[(r + 6) & 7], a, b); W1(wq[r], 0); }   // 4 S + W
      if (PAT == 16) { asm volatile("{ .reg .b32 t; mul.lo.u32 t, %0, 0xD2511F53; xor.b32 %0, %0, t; }" : "+r"(u[r & 3])); }     // IMAD lo
      if (PAT == 17) { asm volatile("{ .reg .b32 t; mul.hi.u32 t, %0, 0xD2511F53; xor.b32 %0, %0, t; }" : "+r"(u[r & 3])); }     // IMAD.HI
      if (PAT == 18) { asm volatile("{ .reg .b32 t, v; mul.hi.u32 t, %0, 0xD2511F53; mul.lo.u32 v, %0, 0xD2511F53; xor.b32 %0, v, t; }" : "+r"(u[r & 3])); }   // hi + lo
      if (PAT == 15) { F2(p[r], A, B); F2(p[(r + 4) & 7], A, B); W1(wq[r], 0); }   // P P + W
    }
  }
  float acc = 0;
  for (int i = 0; i < 8; ++i) { acc += s[i] + (float)(p[i] & 0xffff); }
  acc += u[0] + u[1] + u[2] + u[3] + (float)((wq[0] ^ wq[1] ^ wq[2] ^ wq[3] ^ wq[4] ^ wq[5] ^ wq[6] ^ wq[7]) >> 40);
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
  run<10>("W", 0, 1);
  run<11>("P+W", 2, 2);
  run<12>("S+W", 1, 2);
  run<13>("SS+W", 2, 3);
  run<14>("SSSS+W", 4, 5);
  run<15>("PP+W", 4, 3);
  run<16>("mul.lo+xor", 0, 2);
  run<17>("mul.hi+xor", 0, 2);
  run<18>("hi+lo+xor", 0, 3);
  return 0;
}
