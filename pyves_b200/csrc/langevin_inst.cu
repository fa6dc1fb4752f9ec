// Explicit instantiations of the step kernels, split into groups so the translation units
// compile in parallel:  nvcc -DPV_T=float|double -DPV_GROUP=0..3 langevin_inst.cu
#include "langevin.cuh"

#ifndef PV_T
#define PV_T float
#endif
#ifndef PV_GROUP
#define PV_GROUP 0
#endif

namespace pv {
using T = PV_T;

#define PV_FAST(POT, BIAS, WPT, MINB)                                                                     \
  template <>                                                                                             \
  cudaError_t launch_langevin2d<T, POT, BIAS, WPT, MINB>(                                                 \
      const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,                         \
      const typename BIAS::Params& BP, const RngKeys& K, cudaStream_t st) {                               \
    return launch_langevin2d_impl<T, POT, BIAS, WPT, MINB>(A, I, PP, BP, K, st);                          \
  }
#define PV_PAIR(POT, BIAS, MINB)                                                                          \
  template <>                                                                                             \
  cudaError_t launch_langevin2d_pair<T, POT, BIAS, MINB>(                                                 \
      const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,                         \
      const typename BIAS::Params& BP, const RngKeys& K, cudaStream_t st) {                               \
    return launch_langevin2d_pair_impl<T, POT, BIAS, MINB>(A, I, PP, BP, K, st);                          \
  }
#define PV_COMMA ,

#if PV_GROUP == 0   // double well: the BASELINE.json headline family
PV_FAST(PotDoubleWell2D<T>, BiasNone<T>, 2, 1)
PV_FAST(PotDoubleWell2D<T>, BiasLegendre1D<T PV_COMMA 5>, 2, 1)
PV_PAIR(PotDoubleWell2D<T>, BiasNone<T>, 1)
PV_PAIR(PotDoubleWell2D<T>, BiasLegendre1D<T PV_COMMA 5>, 1)
PV_FAST(PotDoubleWell2D<T>, BiasLegendre2DStatic<T PV_COMMA 8 PV_COMMA 8>, 2, 1)
PV_FAST(PotDoubleWell2D<T>, BiasLegendre2DStatic<T PV_COMMA 16 PV_COMMA 16>, 1, 1)
PV_FAST(PotDoubleWell2D<T>, BiasLegendre2DStaticT<T PV_COMMA 8 PV_COMMA 8>, 2, 1)
PV_FAST(PotDoubleWell2D<T>, BiasLegendre2DStaticT<T PV_COMMA 16 PV_COMMA 16 PV_COMMA false PV_COMMA kHybrid16>, 1, 1)
PV_FAST(PotDoubleWell2D<T>, BiasLegendre2DStaticT<T PV_COMMA 8 PV_COMMA 8 PV_COMMA true>, 2, 1)
PV_FAST(PotDoubleWell2D<T>, BiasLegendre2DStaticT<T PV_COMMA 16 PV_COMMA 16 PV_COMMA true PV_COMMA kHybrid16>, 1, 1)
// every kernel that reads the __constant__ weights lives in THIS translation unit (the array is per unit)
PV_PAIR(PotDoubleWell2D<T>, BiasLegendre1D<T PV_COMMA 5 PV_COMMA true>, 1)
PV_PAIR(PotSlipBond2D<T>, BiasLegendre1D<T PV_COMMA 12 PV_COMMA true>, 1)
PV_PAIR(PotGeneric<T>, BiasLegendre1D<T PV_COMMA 0 PV_COMMA true>, 1)
template <> cudaError_t upload_static_weights<T>(const T* dev_src, int n, cudaStream_t st) {
  if (n > kConstWeights) return cudaErrorInvalidValue;
  if (sizeof(T) == 4) return cudaMemcpyToSymbolAsync(c_wt_f32, dev_src, (size_t)n * sizeof(T), 0, cudaMemcpyDeviceToDevice, st);
  return cudaMemcpyToSymbolAsync(c_wt_f64, dev_src, (size_t)n * sizeof(T), 0, cudaMemcpyDeviceToDevice, st);
}
#elif PV_GROUP == 1  // slip bond, Szabo-Berezhkovskii
PV_FAST(PotDoubleWell2D<T>, BiasLegendre2DSmem<T PV_COMMA 32 PV_COMMA 1 PV_COMMA 5>, 1, 2)
PV_FAST(PotDoubleWell2D<T>, BiasLegendre2DSmem<T PV_COMMA 32 PV_COMMA 2 PV_COMMA 5>, 1, 2)
PV_FAST(PotSlipBond2D<T>, BiasNone<T>, 2, 1)
PV_FAST(PotSlipBond2D<T>, BiasLegendre1D<T PV_COMMA 12>, 2, 1)
PV_PAIR(PotSlipBond2D<T>, BiasNone<T>, 1)
PV_PAIR(PotSlipBond2D<T>, BiasLegendre1D<T PV_COMMA 12>, 1)
PV_PAIR(PotSzaboBerezhkovskii<T>, BiasNone<T>, 1)
PV_FAST(PotSzaboBerezhkovskii<T>, BiasNone<T>, 2, 1)
PV_FAST(PotSzaboBerezhkovskii<T>, BiasMlp<T PV_COMMA 0 PV_COMMA PYVES_ACT_RELU>, 2, 1)
PV_FAST(PotSzaboBerezhkovskii<T>, BiasMlp<T PV_COMMA 48 PV_COMMA PYVES_ACT_RELU>, 1, 1)
PV_FAST(PotSzaboBerezhkovskii<T>, BiasPwl1D<T>, 2, 1)
#elif PV_GROUP == 2  // any potential x each bias family
PV_FAST(PotGeneric<T>, BiasNone<T>, 2, 1)
PV_FAST(PotGeneric<T>, BiasLegendre1D<T PV_COMMA 0>, 2, 1)
PV_PAIR(PotGeneric<T>, BiasNone<T>, 1)
PV_PAIR(PotGeneric<T>, BiasLegendre1D<T PV_COMMA 0>, 1)
PV_FAST(PotGeneric<T>, BiasLegendre2DSmem<T PV_COMMA 16 PV_COMMA 1 PV_COMMA 7>, 1, 3)
PV_FAST(PotGeneric<T>, BiasLegendre2DStaticT<T PV_COMMA 16 PV_COMMA 16 PV_COMMA false PV_COMMA kHybrid16>, 1, 1)
PV_FAST(PotGeneric<T>, BiasLegendre2DSmem<T PV_COMMA 32 PV_COMMA 1 PV_COMMA 5>, 1, 2)
PV_FAST(PotGeneric<T>, BiasLegendre2DSmem<T PV_COMMA 32 PV_COMMA 2 PV_COMMA 5>, 1, 2)
PV_FAST(PotGeneric<T>, BiasOther<T>, 1, 1)
PV_FAST(PotGeneric<T>, BiasPwl1D<T>, 2, 1)
#elif PV_GROUP == 3  // any potential x MLP widths; the general kernel
PV_FAST(PotGeneric<T>, BiasMlp<T PV_COMMA 0 PV_COMMA PYVES_ACT_RELU>, 2, 1)
PV_FAST(PotGeneric<T>, BiasMlp<T PV_COMMA 0 PV_COMMA PYVES_ACT_TANH>, 2, 1)
PV_FAST(PotGeneric<T>, BiasMlp<T PV_COMMA 32 PV_COMMA PYVES_ACT_RELU>, 1, 1)
PV_FAST(PotGeneric<T>, BiasMlp<T PV_COMMA 48 PV_COMMA PYVES_ACT_RELU>, 1, 1)
PV_FAST(PotGeneric<T>, BiasMlp<T PV_COMMA 64 PV_COMMA PYVES_ACT_RELU>, 1, 1)
PV_FAST(PotGeneric<T>, BiasMlp<T PV_COMMA 32 PV_COMMA PYVES_ACT_TANH>, 1, 1)
PV_FAST(PotGeneric<T>, BiasMlp<T PV_COMMA 48 PV_COMMA PYVES_ACT_TANH>, 1, 1)
PV_FAST(PotGeneric<T>, BiasMlp<T PV_COMMA 64 PV_COMMA PYVES_ACT_TANH>, 1, 1)
template <>
cudaError_t launch_langevin_general<T, 2>(const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,
                                          const BiasAnyParams<T>& BP, const RngKeys& K, cudaStream_t st) {
  return launch_langevin_general_impl<T, 2>(A, I, PP, BP, K, st);
}
template <>
cudaError_t launch_langevin_general<T, 3>(const StepArgs<T>& A, const IntegratorParams<T>& I, const PotParams<T>& PP,
                                          const BiasAnyParams<T>& BP, const RngKeys& K, cudaStream_t st) {
  return launch_langevin_general_impl<T, 3>(A, I, PP, BP, K, st);
}
#endif

}  // namespace pv
