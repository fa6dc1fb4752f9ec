// Stand-alone harness: times the 16 x 16 step kernel variants without rebuilding the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o step16 step16.cu
//   ./step16 [n_walkers] [steps] 
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include "../../pyves_b200/csrc/langevin.cuh"
#include "bias_exp.cuh"

using namespace pv;

#ifndef WPT_T
#define WPT_T 1
#endif
#ifndef MINB_T
#define MINB_T 1
#endif

template <class Bias, int WPT, int MINB>
double run(const char* tag, const typename Bias::Params& BP, int64_t n, int64_t steps, const std::vector<float>& p0,
           const std::vector<float>& v0, std::vector<float>& out, int reps) {
  float *pos, *vel;
  cudaMalloc(&pos, 2 * n * sizeof(float));
  cudaMalloc(&vel, 2 * n * sizeof(float));
  StepArgs<float> A{};
  A.pos = pos; A.vel = vel; A.n = n; A.gid0 = 0; A.step0 = 0; A.nsteps = steps; A.noise = nullptr; A.traj = nullptr; A.every = 1; A.n_rec = 0;
  IntegratorParams<float> I{};
  const double kT = 8.31446261815324e-3 * 300, g = 20.0, dt = 0.01, a = std::exp(-g * dt);
  I.a = (float)a; I.b_over_m = (float)((1 - a) / g); I.c_over_sqrtm = (float)std::sqrt(kT * (1 - a * a)); I.dt = (float)dt; I.dt_over_m = (float)dt; I.scheme = 0;
  PotParams<float> PP{};
  RngKeys K = make_keys(1234567ull);
  auto kern = langevin2d_kernel<float, PotDoubleWell2D<float>, Bias, WPT, MINB, false>;
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  const int64_t per_block = (int64_t)kStepThreads * WPT;
  const unsigned grid = (unsigned)((n + per_block - 1) / per_block);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    cudaMemcpy(pos, p0.data(), 2 * n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(vel, v0.data(), 2 * n * sizeof(float), cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    kern<<<grid, kStepThreads, Bias::smem_bytes(BP)>>>(A, I, PP, BP, K);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  out.resize(2 * n);
  cudaMemcpy(out.data(), pos, 2 * n * sizeof(float), cudaMemcpyDeviceToHost);
  const double rate = (double)n * steps / (best * 1e-3);
  printf("%-28s regs %3d local %4zu  %8.3f ms  %.3e walker-steps/s  frac(72.5TF) %.3f  %s\n", tag, fa.numRegs, fa.localSizeBytes, best, rate,
         rate * 1300 * 1e-12 / 72.5, err == cudaSuccess ? "" : cudaGetErrorString(err));
  cudaFree(pos); cudaFree(vel);
  return rate;
}

double maxdiff(const std::vector<float>& a, const std::vector<float>& b) {
  double m = 0;
  for (size_t i = 0; i < a.size(); ++i) m = fmax(m, fabs((double)a[i] - b[i]));
  return m;
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 1000000;
  const int64_t steps = argc > 2 ? atoll(argv[2]) : 500;
  std::vector<float> p0(2 * n), v0(2 * n), w(256);
  srand(1);
  auto u = []() { return rand() / (double)RAND_MAX; };
  for (int64_t i = 0; i < n; ++i) {
    p0[i] = (float)((u() < 0.5 ? -1.4 : 1.3) + 0.2 * (u() - 0.5));
    p0[n + i] = (float)(0.6 * (u() - 0.5));
    v0[i] = (float)(1.5 * (u() - 0.5));
    v0[n + i] = (float)(1.5 * (u() - 0.5));
  }
  for (int i = 0; i < 256; ++i) w[i] = (float)(0.2 * (u() - 0.5));
  Legendre2DStaticParams<float, 16, 16> S{};
  S.cx = 0; S.ihx = 1 / 2.5f; S.cy = 0; S.ihy = 1 / 2.0f;
  Legendre2DStaticTParams<float, 16, 16> Tt{};
  Tt.cx = 0; Tt.ihx = 1 / 2.5f; Tt.cy = 0; Tt.ihy = 1 / 2.0f;
  for (int i = 0; i < 16; ++i)
    for (int j = 0; j < 16; ++j) { S.w[i * 16 + j] = w[i * 16 + j]; Tt.wt[j * 16 + i] = w[i * 16 + j]; }
  static Legendre2DRolledTParams<float, 16, 16> R{};
  R.cx = 0; R.ihx = 1 / 2.5f; R.cy = 0; R.ihy = 1 / 2.0f;
  for (int k = 0; k < 16; ++k) { R.ca[k] = (2.f * k + 1) / (k + 1); R.cb[k] = (float)k / (k + 1); R.cc[k] = 2.f * k + 1; }
  for (int k = 0; k < 256; ++k) R.wt[k] = Tt.wt[k];
  const char* only = argc > 3 ? argv[3] : nullptr;
  std::vector<float> ref, out;
  if (only && std::string(only) == "thd") {   // duplicated shared-memory weights vs broadcast scalars
    static Legendre2DStaticTHParams<16, 16> TH{};
    TH.cx = 0; TH.ihx = 1 / 2.5f; TH.cy = 0; TH.ihy = 1 / 2.0f;
    for (int k = 0; k < 256; ++k) TH.wt[k] = Tt.wt[k];
    { float* wd; cudaMalloc(&wd, 1024); cudaMemcpy(wd, Tt.wt, 1024, cudaMemcpyHostToDevice); TH.wdev = wd; }
    const int64_t nc = 4096;
    std::vector<float> pc(2 * nc), vc(2 * nc), o2;
    for (int64_t i = 0; i < nc; ++i) { pc[i] = p0[i]; pc[nc + i] = p0[n + i]; vc[i] = v0[i]; vc[nc + i] = v0[n + i]; }
    run<BiasLegendre2DStaticTH<16, 16, 4>, 1, 1>("check TH4", TH, nc, 40, pc, vc, ref, 1);
    run<BiasLegendre2DStaticTHD<16, 16, 4>, 1, 1>("check THD4", TH, nc, 40, pc, vc, o2, 1);
    printf("max |dx| after 40 steps, THD vs TH: %.3e\n", maxdiff(ref, o2));
    run<BiasLegendre2DStaticTH<16, 16, 4>, 1, 1>("TH J0=4", TH, n, steps, p0, v0, out, 3);
    run<BiasLegendre2DStaticTH<16, 16, 4, false, true>, 1, 1>("check TH4 nonvolatile", TH, nc, 40, pc, vc, o2, 1);
    printf("max |dx| after 40 steps, TH nonvolatile vs TH: %.3e\n", maxdiff(ref, o2));
    run<BiasLegendre2DStaticTH<16, 16, 4, false, true>, 1, 1>("TH J0=4 nonvolatile", TH, n, steps, p0, v0, out, 3);
    run<BiasLegendre2DStaticTH<16, 16, 5, false, true>, 1, 1>("TH J0=5 nonvolatile", TH, n, steps, p0, v0, out, 3);
    run<BiasLegendre2DStaticTH<16, 16, 6, false, true>, 1, 1>("TH J0=6 nonvolatile", TH, n, steps, p0, v0, out, 3);
    if (argc > 4) return 0;
    run<BiasLegendre2DStaticTHD<16, 16, 4>, 1, 1>("THD J0=4", TH, n, steps, p0, v0, out, 3);
    run<BiasLegendre2DStaticTHP<16, 16, 4, 4>, 1, 1>("check THP4", TH, nc, 40, pc, vc, o2, 1);
    printf("max |dx| after 40 steps, THP vs TH: %.3e\n", maxdiff(ref, o2));
    run<BiasLegendre2DStaticTHP<16, 16, 4, 4>, 1, 1>("THP J0=4 D4", TH, n, steps, p0, v0, out, 3);
    run<BiasLegendre2DStaticTHP<16, 16, 4, 8>, 1, 1>("THP J0=4 D8", TH, n, steps, p0, v0, out, 3);
    run<BiasLegendre2DStaticTHP<16, 16, 2, 6>, 1, 1>("THP J0=2 D6", TH, n, steps, p0, v0, out, 3);
    run<BiasLegendre2DStaticTHP<16, 16, 6, 6>, 1, 1>("THP J0=6 D6", TH, n, steps, p0, v0, out, 3);
    return 0;
  }
  if (only) {
    std::string o(only);
#ifdef ROLLED
    if (o == "rolled") run<BiasLegendre2DRolledT<float, 16, 16, UNR_T, SRC_T>, WPT_T, MINB_T>("rolled", R, n, steps, p0, v0, out, 2);
#endif
    if (o == "base") run<BiasLegendre2DStatic<float, 16, 16, 4, false>, 1, 1>("base rows4 nosplit", S, n, steps, p0, v0, ref, 2);
    if (o == "T2") run<BiasLegendre2DStaticT<float, 16, 16>, WPT_T, MINB_T>("T gch2", Tt, n, steps, p0, v0, out, 2);
    return 0;
  }
  static Legendre2DStaticTHParams<16, 16> TH{};
  TH.cx = 0; TH.ihx = 1 / 2.5f; TH.cy = 0; TH.ihy = 1 / 2.0f;
  for (int k = 0; k < 256; ++k) TH.wt[k] = Tt.wt[k];
  { float* wd; cudaMalloc(&wd, 1024); cudaMemcpy(wd, Tt.wt, 1024, cudaMemcpyHostToDevice); TH.wdev = wd; }
  // numerics check on a short run
  const int64_t nc = 4096;
  std::vector<float> pc(2 * nc), vc(2 * nc);
  for (int64_t i = 0; i < nc; ++i) { pc[i] = p0[i]; pc[nc + i] = p0[n + i]; vc[i] = v0[i]; vc[nc + i] = v0[n + i]; }
  run<BiasLegendre2DStatic<float, 16, 16, 4, false>, 1, 1>("check base", S, nc, 40, pc, vc, ref, 1);
  run<BiasLegendre2DStaticT<float, 16, 16>, WPT_T, MINB_T>("check T", Tt, nc, 40, pc, vc, out, 1);
  printf("max |dx| after 40 steps, T vs base: %.3e\n", maxdiff(ref, out));
#ifdef ROLLED
  run<BiasLegendre2DRolledT<float, 16, 16, UNR_T, SRC_T>, WPT_T, MINB_T>("check rolled", R, nc, 40, pc, vc, out, 1);
  printf("max |dx| after 40 steps, rolled vs base: %.3e\n", maxdiff(ref, out));
  run<BiasLegendre2DRolledT<float, 16, 16, UNR_T, SRC_T>, WPT_T, MINB_T>("rolled", R, n, steps, p0, v0, out, 3);
  return 0;
#endif
  { std::vector<float> tmp; run<BiasLegendre2DStatic<float, 16, 16, 4, false>, 1, 1>("base rows4 nosplit", S, n, steps, p0, v0, tmp, 3); }
  { std::vector<float> tmp; run<BiasLegendre2DStaticT<float, 16, 16>, WPT_T, MINB_T>("T gch2", Tt, n, steps, p0, v0, tmp, 3); }
  run<BiasLegendre2DStaticTH<16, 16, 6>, 1, 1>("check TH6", TH, nc, 40, pc, vc, out, 1);
  printf("max |dx| after 40 steps, TH vs base: %.3e\n", maxdiff(ref, out));
  run<BiasLegendre2DStaticTH<16, 16, 4>, 1, 1>("TH J0=4", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 6>, 1, 1>("TH J0=6", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 10>, 1, 1>("TH J0=10", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 4>, 2, 1>("TH J0=4 WPT2", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 8>, 2, 1>("TH J0=8 WPT2", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 16>, 2, 1>("TH J0=16 WPT2", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 5>, 1, 1>("TH J0=5", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 8>, 1, 1>("TH J0=8", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 4, true>, 1, 1>("TH tail 4 (2 cols)", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 6, true>, 1, 1>("TH tail 6 (4 cols)", TH, n, steps, p0, v0, out, 3);
  run<BiasLegendre2DStaticTH<16, 16, 8, true>, 1, 1>("TH tail 8 (6 cols)", TH, n, steps, p0, v0, out, 3);
  return 0;
}
