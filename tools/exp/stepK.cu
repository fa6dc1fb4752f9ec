// Stand-alone harness for the runtime-size (shared-memory) 2D Legendre functors.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DKP=32 -DUNR_T=2 -DMINB_T=1 -o stepK stepK.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include "../../pyves_b200/csrc/langevin.cuh"
#include "bias_exp.cuh"
using namespace pv;
#ifndef KP
#define KP 32
#endif
#ifndef UNR_T
#define UNR_T 2
#endif
#ifndef MINB_T
#define MINB_T 1
#endif
#ifndef TY
#define TY float
#endif
typedef TY R;

template <class Bias, int WPT, int MINB>
double run(const char* tag, const typename Bias::Params& BP, int64_t n, int64_t steps, const std::vector<R>& p0,
           const std::vector<R>& v0, std::vector<R>& out, int reps, double flops) {
  R *pos, *vel;
  cudaMalloc(&pos, 2 * n * sizeof(R));
  cudaMalloc(&vel, 2 * n * sizeof(R));
  StepArgs<R> A{};
  A.pos = pos; A.vel = vel; A.n = n; A.gid0 = 0; A.step0 = 0; A.nsteps = steps; A.noise = nullptr; A.traj = nullptr; A.every = 1; A.n_rec = 0;
  IntegratorParams<R> I{};
  const double kT = 8.31446261815324e-3 * 300, g = 20.0, dt = 0.01, a = std::exp(-g * dt);
  I.a = (R)a; I.b_over_m = (R)((1 - a) / g); I.c_over_sqrtm = (R)std::sqrt(kT * (1 - a * a)); I.dt = (R)dt; I.dt_over_m = (R)dt; I.scheme = 0;
  PotParams<R> PP{};
  RngKeys K = make_keys(1234567ull);
  auto kern = langevin2d_kernel<R, PotDoubleWell2D<R>, Bias, WPT, MINB, false>;
  const size_t sm = Bias::smem_bytes(BP);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  const int64_t per_block = (int64_t)kStepThreads * WPT;
  const unsigned grid = (unsigned)((n + per_block - 1) / per_block);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    cudaMemcpy(pos, p0.data(), 2 * n * sizeof(R), cudaMemcpyHostToDevice);
    cudaMemcpy(vel, v0.data(), 2 * n * sizeof(R), cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    kern<<<grid, kStepThreads, sm>>>(A, I, PP, BP, K);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  out.resize(2 * n);
  cudaMemcpy(out.data(), pos, 2 * n * sizeof(R), cudaMemcpyDeviceToHost);
  const double rate = (double)n * steps / (best * 1e-3);
  printf("%-22s regs %3d local %4zu smem %6zu  %9.3f ms  %.3e walker-steps/s  frac(72.5TF) %.3f  %s\n", tag, fa.numRegs, fa.localSizeBytes, sm, best, rate,
         rate * flops * 1e-12 / 72.5, err == cudaSuccess ? "" : cudaGetErrorString(err));
  cudaFree(pos); cudaFree(vel);
  return rate;
}

double maxdiff(const std::vector<R>& a, const std::vector<R>& b) {
  double m = 0;
  for (size_t i = 0; i < a.size(); ++i) m = fmax(m, fabs((double)a[i] - b[i]));
  return m;
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 1000000;
  const int64_t steps = argc > 2 ? atoll(argv[2]) : 100;
  const int kx = argc > 3 ? atoi(argv[3]) : KP, ky = argc > 4 ? atoi(argv[4]) : KP;
  const char* only = argc > 5 ? argv[5] : nullptr;
  std::vector<R> p0(2 * n), v0(2 * n), w(kx * ky);
  srand(1);
  auto u = []() { return rand() / (double)RAND_MAX; };
  for (int64_t i = 0; i < n; ++i) {
    p0[i] = (R)((u() < 0.5 ? -1.4 : 1.3) + 0.2 * (u() - 0.5));
    p0[n + i] = (R)(0.6 * (u() - 0.5));
    v0[i] = (R)(1.5 * (u() - 0.5));
    v0[n + i] = (R)(1.5 * (u() - 0.5));
  }
  for (int i = 0; i < kx * ky; ++i) w[i] = (R)(0.05 * (u() - 0.5));
  R* wd; cudaMalloc(&wd, kx * ky * sizeof(R));
  cudaMemcpy(wd, w.data(), kx * ky * sizeof(R), cudaMemcpyHostToDevice);
  Legendre2DSmemParams<R> S{}; S.cx = 0; S.ihx = 1 / 2.5f; S.cy = 0; S.ihy = 1 / 2.0f; S.kx = kx; S.ky = ky; S.w = wd;
  Legendre2DSmemTParams<R> Tt{}; Tt.cx = 0; Tt.ihx = 1 / 2.5f; Tt.cy = 0; Tt.ihy = 1 / 2.0f; Tt.kx = kx; Tt.ky = ky; Tt.w = wd;
  const double flops = 4.0 * kx * ky + 4 * kx + 6 * (kx + ky - 2) + 32;
  std::vector<R> ref, out;
  if (!only) {
    const int64_t nc = 4096;
    std::vector<R> pc(2 * nc), vc(2 * nc);
    for (int64_t i = 0; i < nc; ++i) { pc[i] = p0[i]; pc[nc + i] = p0[n + i]; vc[i] = v0[i]; vc[nc + i] = v0[n + i]; }
    run<BiasLegendre2DSmem<R, KP>, 1, 1>("check base", S, nc, 40, pc, vc, ref, 1, flops);
    run<BiasLegendre2DSmemT<R, KP, UNR_T>, 1, MINB_T>("check T", Tt, nc, 40, pc, vc, out, 1, flops);
    printf("max |dx| after 40 steps, T vs base: %.3e\n", maxdiff(ref, out));
  }
  if (!only || std::string(only) == "base") run<BiasLegendre2DSmem<R, KP>, 1, 1>("base", S, n, steps, p0, v0, ref, 2, flops);
  if (!only || std::string(only) == "T") run<BiasLegendre2DSmemT<R, KP, UNR_T>, 1, MINB_T>("T", Tt, n, steps, p0, v0, out, 2, flops);
  return 0;
}
