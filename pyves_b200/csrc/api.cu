// C ABI of pyves_b200 (include/pyves_b200.h): handle, parameter upload, kernel dispatch.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "langevin.cuh"
#include "misc_kernels.cuh"
#include "moments.cuh"

using namespace pv;

static std::atomic<int64_t> g_launches{0};
// content tag of device-resident coefficients: a fresh value per pyves_set_bias_legendre{1d,2d}_device call,
// copied by pyves_set_bias_like (0 = none)
std::atomic<uint64_t> g_const_tag{1};
static thread_local std::string g_err;

struct pyves_ctx {
  int device = 0;
  int64_t n = 0;
  int dims = 2;
  uint64_t seed = 0;
  int prec = PYVES_F32;
  size_t esz = 4;
  void* pos = nullptr;
  void* vel = nullptr;
  bool have_int = false, have_pot = false;
  double mass = 1, kT = 0, friction = 0, dt = 0;
  int scheme = 0;
  PotParams<double> pot{};
  uint64_t gid0 = 0;
  int64_t step = 0;
  int bias_kind = PYVES_BIAS_NONE;
  // Legendre 1D
  int l1_axis = 0, l1_degree = 0;
  double l1_lo = -1, l1_hi = 1;
  double l1_w[kMaxK1] = {0};
  // Legendre 2D
  int l2_dx = 0, l2_dy = 0;
  double l2_mm[4] = {-1, 1, -1, 1};
  std::vector<double> l2_w;
  void* l2_w_dev = nullptr;
  // device-resident coefficients (pyves_set_bias_legendre2d_device)
  bool l2_dev_mode = false;      // weights live on the device only: the host copy l2_w is stale
  void* l2_wt_dev = nullptr;     // transposed element-type copy wt[ky][kx] (source of the __constant__ upload)
  size_t l2_dev_count = 0;       // elements allocated in l2_w_dev / l2_wt_dev
  bool l1_dev_mode = false;      // pyves_set_bias_legendre1d_device: l1_w (host) may be stale
  bool l1_host_stale = false;
  void* l1_wt_dev = nullptr;     // element-type copy, kMaxK1 entries
  double* l1_w64_dev = nullptr;  // FP64 copy (host refresh for the consumers that take weights by value)
  uint64_t const_tag = 0;        // identifies the content of l1_wt_dev / l2_wt_dev for the __constant__ array's owner check
  std::map<void**, size_t> alloc_count;   // element counts of the buffers managed by upload()
  // MLP
  int mlp_axis = 0, mlp_act = 0;
  double mlp_lo = 0, mlp_hi = 1;
  std::vector<int> mlp_sizes;       // hidden widths
  size_t mlp_np = 0;                // number of parameters (torch parameters() order)
  void* mlp_theta_dev = nullptr;
  void* mlp_packed_dev = nullptr;
  void* pwl_brk = nullptr;         // ReLU NNBasis1D compiled to a piecewise-constant dV/ds (BiasPwl1D): breakpoints, slopes, count
  void* pwl_slope = nullptr;
  int* pwl_count = nullptr;
  int* pwl_grid = nullptr;         // bucket grid of the breakpoints (kPwlBuckets + 1 ints) and its {origin, buckets per unit}
  void* pwl_grid_map = nullptr;
  int pwl_cap = 0;
  bool pwl_valid = false;          // the handle's bias is a ReLU net this applies to
  bool pwl_stale = true;           // the table does not reflect the current parameters yet (built on first use, ensure_pwl)
  double* pwl_hist = nullptr;      // moments through the pieces: histogram, pseudo rows, their weights
  void* pwl_rows = nullptr;
  void* pwl_row_w = nullptr;
  int pwl_hist_cap = 0;
  void* tc_img = nullptr;          // operand image of the tensor-core Legendre step kernel (langevin2d_tc_scratch_bytes())
  bool tc_img_valid = false;       // built from the current 2D coefficients (reset by every pyves_set_bias_legendre2d*)
  int mlp_H = -1;                   // -1: no fast functor; 0: streamed; 32/48/64: one matmul
  int mlp_width = 0, mlp_linear = 0;
  double mlp_bout = 0, mlp_u0 = 0;
  // path CV
  int pc_degree = 0, pc_npath = 0;
  double pc_lam = 0, pc_phi = 0;
  double pc_w[kMaxK1] = {0};
  void* pc_path_dev = nullptr;
  // harmonic
  int h_axis = 0;
  double h_k = 0, h_s0 = 0;
  // scratch
  double* scratch = nullptr;
  size_t scratch_doubles = 0;
  void* stage = nullptr;
  size_t stage_bytes = 0;
  unsigned long long* fail_dev = nullptr;
  int variant = -1;
  int mom_inside = 0;   // pyves_set_moment_domain
  std::string err;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int fail(pyves_ctx* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_err = msg;
  return code;
}

#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(h, PYVES_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                   \
  } while (0)

int ensure_scratch(pyves_ctx* h, size_t doubles) {
  if (doubles <= h->scratch_doubles) return PYVES_OK;
  if (h->scratch) cudaFree(h->scratch);
  h->scratch = nullptr;
  h->scratch_doubles = 0;
  CU(cudaMalloc(&h->scratch, doubles * sizeof(double)));
  h->scratch_doubles = doubles;
  return PYVES_OK;
}

int ensure_stage(pyves_ctx* h, size_t bytes) {
  if (bytes <= h->stage_bytes) return PYVES_OK;
  if (h->stage) cudaFree(h->stage);
  h->stage = nullptr;
  h->stage_bytes = 0;
  CU(cudaMalloc(&h->stage, bytes));
  h->stage_bytes = bytes;
  return PYVES_OK;
}

// upload a double host array as the handle's element type, ordered on the caller's stream `st`: the copy comes
// from pageable host memory, so cudaMemcpyAsync returns once the source has been staged (the temporary may die)
// and the transfer itself is ordered after the work already enqueued on `st` -- a step / eval / moments kernel
// still reading the previous parameters on that stream finishes first.  Launches on OTHER streams must be ordered
// by the caller (wait for `st`).
// device buffer of `count` elements of the handle's type behind *dev; reused when the size is unchanged
// (coefficient updates): cudaFree would synchronise the device
int ensure_buffer(pyves_ctx* h, void** dev, size_t count) {
  auto it = h->alloc_count.find(dev);
  const bool reuse = *dev != nullptr && it != h->alloc_count.end() && it->second == count && count > 0;
  if (reuse) return PYVES_OK;
  if (*dev) cudaFree(*dev);
  *dev = nullptr;
  h->alloc_count.erase(dev);
  if (count == 0) return PYVES_OK;
  CU(cudaMalloc(dev, count * h->esz));
  h->alloc_count[dev] = count;
  return PYVES_OK;
}

int upload(pyves_ctx* h, void** dev, const double* src, size_t count, cudaStream_t st) {
  { const int rc = ensure_buffer(h, dev, count); if (rc) return rc; }
  if (count == 0) return PYVES_OK;
  if (h->prec == PYVES_F64) {
    CU(cudaMemcpyAsync(*dev, src, count * sizeof(double), cudaMemcpyHostToDevice, st));
  } else {
    std::vector<float> tmp(count);
    for (size_t i = 0; i < count; ++i) tmp[i] = (float)src[i];
    CU(cudaMemcpyAsync(*dev, tmp.data(), count * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  return PYVES_OK;
}

template <typename T> PotParams<T> pot_params(const pyves_ctx* h) {
  PotParams<T> P;
  P.kind = h->pot.kind;
  for (int i = 0; i < 11; ++i) P.p[i] = (T)h->pot.p[i];
  return P;
}

template <typename T> IntegratorParams<T> integrator_params(const pyves_ctx* h) {
  const double a = std::exp(-h->friction * h->dt);
  const double b = h->friction != 0.0 ? (1.0 - a) / h->friction : h->dt;
  const double c = std::sqrt(h->kT * (1.0 - a * a));
  IntegratorParams<T> I;
  I.a = (T)a;
  I.b_over_m = (T)(b / h->mass);
  I.c_over_sqrtm = (T)(c / std::sqrt(h->mass));
  I.dt = (T)h->dt;
  I.dt_over_m = (T)(h->dt / h->mass);
  I.scheme = h->scheme;
  return I;
}

template <typename T> Legendre1DParams<T> l1_params(const pyves_ctx* h) {
  Legendre1DParams<T> P;
  P.axis = h->l1_axis;
  P.degree = h->l1_degree;
  P.centre = (T)((h->l1_lo + h->l1_hi) / 2);
  P.inv_half = (T)(1.0 / ((h->l1_hi - h->l1_lo) / 2));
  for (int i = 0; i < kMaxK1; ++i) P.w[i] = (T)h->l1_w[i];
  return P;
}

template <typename T> Legendre2DSmemParams<T> l2_params(const pyves_ctx* h) {
  Legendre2DSmemParams<T> P;
  P.cx = (T)((h->l2_mm[0] + h->l2_mm[1]) / 2);
  P.ihx = (T)(1.0 / ((h->l2_mm[1] - h->l2_mm[0]) / 2));
  P.cy = (T)((h->l2_mm[2] + h->l2_mm[3]) / 2);
  P.ihy = (T)(1.0 / ((h->l2_mm[3] - h->l2_mm[2]) / 2));
  P.kx = h->l2_dx + 1;
  P.ky = h->l2_dy + 1;
  P.w = static_cast<const T*>(h->l2_w_dev);
  return P;
}

template <typename T, int KX, int KY> Legendre2DStaticParams<T, KX, KY> l2_static_params(const pyves_ctx* h) {
  const Legendre2DSmemParams<T> S = l2_params<T>(h);
  Legendre2DStaticParams<T, KX, KY> P;
  P.cx = S.cx;
  P.ihx = S.ihx;
  P.cy = S.cy;
  P.ihy = S.ihy;
  for (int i = 0; i < KX * KY; ++i) P.w[i] = (T)h->l2_w[i];
  return P;
}

template <typename T, int KX, int KY> Legendre2DStaticTParams<T, KX, KY> l2_static_t_params(const pyves_ctx* h) {
  const Legendre2DSmemParams<T> S = l2_params<T>(h);
  Legendre2DStaticTParams<T, KX, KY> P;
  P.cx = S.cx;
  P.ihx = S.ihx;
  P.cy = S.cy;
  P.ihy = S.ihy;
  for (int i = 0; i < KX; ++i)
    for (int j = 0; j < KY; ++j) P.wt[j * KX + i] = (T)h->l2_w[i * KY + j];
  return P;
}

template <typename T, int KX, int KY> Legendre2DStaticTCParams<T, KX, KY> l2_static_tc_params(const pyves_ctx* h) {
  const Legendre2DSmemParams<T> S = l2_params<T>(h);
  Legendre2DStaticTCParams<T, KX, KY> P;
  P.cx = S.cx;
  P.ihx = S.ihx;
  P.cy = S.cy;
  P.ihy = S.ihy;
  return P;
}

// ReLU nets with two or three hidden layers can be compiled into the piecewise-constant dV/ds of BiasPwl1D.  refresh_pwl (called by
// the parameter setters) only decides whether that applies and marks the table stale; ensure_pwl builds it on the stream of the first
// step / moments call that wants it -- an FP64 evaluator handle that never asks never pays for it.
int refresh_pwl(pyves_ctx* h, cudaStream_t) {
  h->pwl_valid = false;
  h->pwl_stale = true;
  if (h->mlp_act != PYVES_ACT_RELU || h->mlp_linear || h->mlp_H < 0) return PYVES_OK;
  const int n1 = h->mlp_H > 0 ? h->mlp_H : h->mlp_width;
  if (n1 < 1 || n1 > 128) return PYVES_OK;
  h->pwl_valid = true;
  return PYVES_OK;
}

int ensure_pwl(pyves_ctx* h, cudaStream_t st) {
  if (!h->pwl_valid || !h->pwl_stale) return PYVES_OK;
  const int cap = mlp_pwl_capacity(h->mlp_H, h->mlp_width);
  if (cap > h->pwl_cap) {
    if (h->pwl_brk) cudaFree(h->pwl_brk);
    if (h->pwl_slope) cudaFree(h->pwl_slope);
    h->pwl_brk = h->pwl_slope = nullptr;
    CU(cudaMalloc(&h->pwl_brk, (size_t)cap * h->esz));
    CU(cudaMalloc(&h->pwl_slope, (size_t)(cap + 1) * h->esz));
    h->pwl_cap = cap;
  }
  if (!h->pwl_count) {
    CU(cudaMalloc(reinterpret_cast<void**>(&h->pwl_count), sizeof(int)));
    CU(cudaMalloc(reinterpret_cast<void**>(&h->pwl_grid), sizeof(int) * (kPwlBuckets + 1)));
    CU(cudaMalloc(&h->pwl_grid_map, 2 * sizeof(double)));
  }
  if (h->prec == PYVES_F64) {
    CU(launch_mlp_pwl_build<double>(static_cast<const double*>(h->mlp_packed_dev), h->mlp_H, h->mlp_width, static_cast<double*>(h->pwl_brk),
                                    static_cast<double*>(h->pwl_slope), h->pwl_count, h->pwl_grid, static_cast<double*>(h->pwl_grid_map), st));
  } else {
    CU(launch_mlp_pwl_build<float>(static_cast<const float*>(h->mlp_packed_dev), h->mlp_H, h->mlp_width, static_cast<float*>(h->pwl_brk),
                                   static_cast<float*>(h->pwl_slope), h->pwl_count, h->pwl_grid, static_cast<float*>(h->pwl_grid_map), st));
  }
  g_launches++;
  h->pwl_stale = false;
  return PYVES_OK;
}

template <typename T> PwlParams<T> pwl_params(const pyves_ctx* h) {
  PwlParams<T> P;
  P.axis = h->mlp_axis;
  P.lo = (T)h->mlp_lo;
  P.inv_range = (T)(1.0 / (h->mlp_hi - h->mlp_lo));
  P.cap = h->pwl_cap;
  P.brk = static_cast<const T*>(h->pwl_brk);
  P.slope = static_cast<const T*>(h->pwl_slope);
  P.count = h->pwl_count;
  P.grid = h->pwl_grid;
  P.grid_map = static_cast<const T*>(h->pwl_grid_map);
  return P;
}

template <typename T> MlpParams<T> mlp_params(const pyves_ctx* h) {
  MlpParams<T> P;
  P.axis = h->mlp_axis;
  P.linear = h->mlp_linear;
  P.width = h->mlp_width;
  P.lo = (T)h->mlp_lo;
  P.inv_range = (T)(1.0 / (h->mlp_hi - h->mlp_lo));
  P.bout = (T)h->mlp_bout;
  P.u0 = (T)h->mlp_u0;
  P.packed = static_cast<const T*>(h->mlp_packed_dev);
  return P;
}

template <typename T> OtherParams<T> other_params(const pyves_ctx* h) {
  OtherParams<T> P;
  std::memset(&P, 0, sizeof(P));
  P.kind = h->bias_kind;
  if (h->bias_kind == PYVES_BIAS_PATHCV) {
    P.degree = h->pc_degree;
    P.npath = h->pc_npath;
    P.lam = (T)h->pc_lam;
    P.phi = (T)h->pc_phi;
    for (int i = 0; i < kMaxK1; ++i) P.w[i] = (T)h->pc_w[i];
    P.path = static_cast<const T*>(h->pc_path_dev);
  } else if (h->bias_kind == PYVES_BIAS_HARMONIC) {
    P.axis = h->h_axis;
    P.k = (T)h->h_k;
    P.s0 = (T)h->h_s0;
  } else if (h->bias_kind == PYVES_BIAS_MLP_1D) {
    P.axis = h->mlp_axis;
    P.act = h->mlp_act;
    P.s0 = (T)h->mlp_lo;
    P.k = (T)(1.0 / (h->mlp_hi - h->mlp_lo));
    P.n_layers = (int)h->mlp_sizes.size() + 1;
    P.sizes[0] = 1;
    for (size_t i = 0; i < h->mlp_sizes.size(); ++i) P.sizes[i + 1] = h->mlp_sizes[i];
    P.sizes[h->mlp_sizes.size() + 1] = 1;
    P.theta = static_cast<const T*>(h->mlp_theta_dev);
  }
  return P;
}

template <typename T> BiasAnyParams<T> any_params(const pyves_ctx* h) {
  BiasAnyParams<T> P;
  std::memset(&P, 0, sizeof(P));
  P.kind = h->bias_kind;
  if (h->bias_kind == PYVES_BIAS_LEGENDRE_1D) P.l1 = l1_params<T>(h);
  else if (h->bias_kind == PYVES_BIAS_LEGENDRE_2D) P.l2 = l2_params<T>(h);
  else if (h->bias_kind != PYVES_BIAS_NONE) P.other = other_params<T>(h);
  return P;
}

int64_t basis_size(const pyves_ctx* h) {
  switch (h->bias_kind) {
    case PYVES_BIAS_LEGENDRE_1D: return h->l1_degree + 1;
    case PYVES_BIAS_LEGENDRE_2D: return (int64_t)(h->l2_dx + 1) * (h->l2_dy + 1);
    case PYVES_BIAS_PATHCV: return h->pc_degree + 1;
    case PYVES_BIAS_MLP_1D: return (int64_t)h->mlp_np;
    default: return 0;
  }
}

// ------------------------------------------------------------------- the __constant__ weight array
// One array per device and process, shared by every handle in device mode.  `owner` is the content tag of the
// coefficients it holds.  Launches and re-uploads may come from different streams, so the slot keeps
// an event per reading stream and one for the last upload: an upload waits for every kernel that still reads the
// old content (other streams only; the own stream is ordered anyway), a launch on another stream than the
// upload's waits for the upload.  The mutex covers bind + launch + record, so handles driven from different host
// threads cannot interleave between the upload and the launch that depends on it.
constexpr int kMaxConstDevices = 64;
struct ConstSlot {
  std::mutex mu;
  uint64_t owner = 0;
  cudaStream_t up_stream = nullptr;
  cudaEvent_t up_event = nullptr;
  struct Reader {
    cudaStream_t st;
    cudaEvent_t ev;
    bool pending;
  };
  std::vector<Reader> readers;
};
ConstSlot g_const[kMaxConstDevices];

template <typename T, class Launch>
int with_const_weights(pyves_ctx* h, const T* src, int count, cudaStream_t st, Launch&& launch) {
  ConstSlot& S = g_const[h->device];
  std::lock_guard<std::mutex> lock(S.mu);
  const uint64_t tag = h->const_tag;
  if (S.owner != tag) {
    for (auto& r : S.readers) {
      if (r.pending && r.st != st) CU(cudaStreamWaitEvent(st, r.ev, 0));
      r.pending = false;
    }
    if (S.readers.size() > 64) {               // streams come and go: forget the idle ones
      for (auto& r : S.readers) cudaEventDestroy(r.ev);
      S.readers.clear();
    }
    S.owner = 0;
    CU(upload_static_weights<T>(src, count, st));
    if (!S.up_event) CU(cudaEventCreateWithFlags(&S.up_event, cudaEventDisableTiming));
    CU(cudaEventRecord(S.up_event, st));
    S.up_stream = st;
    S.owner = tag;
  } else if (S.up_stream != st) {
    CU(cudaStreamWaitEvent(st, S.up_event, 0));
  }
  const int rc = launch();
  if (rc) return rc;
  ConstSlot::Reader* me = nullptr;
  for (auto& r : S.readers)
    if (r.st == st) me = &r;
  if (!me) {
    cudaEvent_t ev = nullptr;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    S.readers.push_back({st, ev, false});
    me = &S.readers.back();
  }
  CU(cudaEventRecord(me->ev, st));
  me->pending = true;
  return PYVES_OK;
}

// ---------------------------------------------------------------------------------------- stepping
template <typename T, class Pot, class Bias, int WPT, int MINB>
int run_fast(pyves_ctx* h, const StepArgs<T>& A, const typename Bias::Params& BP, cudaStream_t st) {
  CU((launch_langevin2d<T, Pot, Bias, WPT, MINB>(A, integrator_params<T>(h), pot_params<T>(h), BP, make_keys(h->seed), st)));
  g_launches++;
  return PYVES_OK;
}

template <typename T, class Pot, class Bias>
int run_pair(pyves_ctx* h, const StepArgs<T>& A, const typename Bias::Params& BP, cudaStream_t st) {
  if (h->variant == 1) return run_fast<T, Pot, Bias, 2, 1>(h, A, BP, st);   // scalar two-walker kernel (tuning runs)
  CU((launch_langevin2d_pair<T, Pot, Bias, 1>(A, integrator_params<T>(h), pot_params<T>(h), BP, make_keys(h->seed), st)));
  g_launches++;
  return PYVES_OK;
}

// Device-mode 1D weights reach the pair kernels through the __constant__ array; every other consumer takes the
// weights by value from the host copy, which is refreshed here (synchronous, rare: eval_bias / energies /
// general kernel on a handle whose coefficients were set from the device).
int refresh_l1_host(pyves_ctx* h, cudaStream_t st) {
  if (!h->l1_dev_mode || !h->l1_host_stale) return PYVES_OK;
  double tmp[kMaxK1];
  // on the caller's stream: the conversion that filled l1_w64_dev was enqueued there (or on a stream it waits for)
  CU(cudaMemcpyAsync(tmp, h->l1_w64_dev, (size_t)(h->l1_degree + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  std::memset(h->l1_w, 0, sizeof(h->l1_w));
  for (int i = 0; i <= h->l1_degree; ++i) h->l1_w[i] = tmp[i];
  h->l1_host_stale = false;
  return PYVES_OK;
}

template <typename T, class Pot, class Bias>
int run_pair_c(pyves_ctx* h, const StepArgs<T>& A, const typename Bias::Params& BP, cudaStream_t st) {
  CU((launch_langevin2d_pair<T, Pot, Bias, 1>(A, integrator_params<T>(h), pot_params<T>(h), BP, make_keys(h->seed), st)));
  g_launches++;
  return PYVES_OK;
}

template <typename T> int run_general(pyves_ctx* h, const StepArgs<T>& A, cudaStream_t st) {
  if (h->bias_kind == PYVES_BIAS_MLP_1D) {
    for (int w : h->mlp_sizes)
      if (w > kGenH) return fail(h, PYVES_EUNSUPPORTED, "MLP hidden width > 128 is not supported");
  }
  { const int rc = refresh_l1_host(h, st); if (rc) return rc; }
  if (h->dims == 2) {
    CU((launch_langevin_general<T, 2>(A, integrator_params<T>(h), pot_params<T>(h), any_params<T>(h), make_keys(h->seed), st)));
  } else {
    CU((launch_langevin_general<T, 3>(A, integrator_params<T>(h), pot_params<T>(h), any_params<T>(h), make_keys(h->seed), st)));
  }
  g_launches++;
  return PYVES_OK;
}

// tuning variant 7: the 2D Legendre contraction on the tensor cores (langevin_tc.cu); returns 1 when it did not apply
template <typename T> int run_tc(pyves_ctx*, const StepArgs<T>&, cudaStream_t) { return 1; }
// builds the operand image of the folded coefficient matrices from the handle's current 2D coefficients (FP32 handles, bases <= 64 x 64)
int refresh_tc_img(pyves_ctx* h, cudaStream_t st) {
  if (h->prec != PYVES_F32) return PYVES_OK;
  if (!h->tc_img) CU(cudaMalloc(&h->tc_img, langevin2d_tc_scratch_bytes()));
  const cudaError_t rc = langevin2d_tc_prepare(l2_params<float>(h), h->tc_img, st);
  if (rc == cudaErrorNotSupported) return PYVES_OK;
  if (rc != cudaSuccess) return fail(h, PYVES_ECUDA, std::string("langevin2d_tc_prepare: ") + cudaGetErrorString(rc));
  g_launches++;
  h->tc_img_valid = true;
  return PYVES_OK;
}

template <> int run_tc<float>(pyves_ctx* h, const StepArgs<float>& A, cudaStream_t st) {
  if (!h->tc_img_valid) {
    const int rc0 = refresh_tc_img(h, st);
    if (rc0) return rc0;
    if (!h->tc_img_valid) return 1;
  }
  const cudaError_t rc = launch_langevin2d_tc(A, integrator_params<float>(h), pot_params<float>(h), l2_params<float>(h), make_keys(h->seed), h->tc_img, st);
  if (rc == cudaErrorNotSupported) return 1;
  if (rc != cudaSuccess) return fail(h, PYVES_ECUDA, std::string("launch_langevin2d_tc: ") + cudaGetErrorString(rc));
  g_launches++;
  return PYVES_OK;
}

// the H x H layer of the three-hidden-layer NN bias on the tensor cores (langevin_tc.cu); same conventions as run_tc
template <typename T> int run_tc_mlp(pyves_ctx*, const StepArgs<T>&, cudaStream_t) { return 1; }
template <> int run_tc_mlp<float>(pyves_ctx* h, const StepArgs<float>& A, cudaStream_t st) {
  const cudaError_t rc = launch_langevin2d_tc_mlp(A, integrator_params<float>(h), pot_params<float>(h), mlp_params<float>(h), h->mlp_H, h->mlp_act,
                                                  make_keys(h->seed), st);
  if (rc == cudaErrorNotSupported) return 1;
  if (rc != cudaSuccess) return fail(h, PYVES_ECUDA, std::string("launch_langevin2d_tc_mlp: ") + cudaGetErrorString(rc));
  g_launches++;
  return PYVES_OK;
}

template <typename T>
int do_step(pyves_ctx* h, int64_t nsteps, const void* noise, void* traj, int64_t every, int64_t n_rec, cudaStream_t st) {
  StepArgs<T> A;
  A.pos = static_cast<T*>(h->pos);
  A.vel = static_cast<T*>(h->vel);
  A.n = h->n;
  A.gid0 = h->gid0;
  A.step0 = h->step;
  A.nsteps = nsteps;
  A.noise = static_cast<const T*>(noise);
  A.traj = static_cast<T*>(traj);
  A.every = every > 0 ? every : 1;
  A.n_rec = n_rec;
  const bool fast = h->dims == 2 && h->scheme == PYVES_SCHEME_OPENMM_LANGEVIN && traj == nullptr && h->variant != 99;
  if (!fast) return run_general<T>(h, A, st);
  const int pk = h->pot.kind;
  using DW = PotDoubleWell2D<T>;
  using SL = PotSlipBond2D<T>;
  using SB = PotSzaboBerezhkovskii<T>;
  using GP = PotGeneric<T>;
  switch (h->bias_kind) {
    case PYVES_BIAS_NONE: {
      typename BiasNone<T>::Params BP;
      if (pk == PYVES_POT_DOUBLE_WELL_2D) return run_pair<T, DW, BiasNone<T>>(h, A, BP, st);
      if (pk == PYVES_POT_SLIP_BOND_2D) return run_pair<T, SL, BiasNone<T>>(h, A, BP, st);
      if (pk == PYVES_POT_SZABO_BEREZHKOVSKII) return run_pair<T, SB, BiasNone<T>>(h, A, BP, st);
      return run_pair<T, GP, BiasNone<T>>(h, A, BP, st);
    }
    case PYVES_BIAS_LEGENDRE_1D: {
      if (h->l1_dev_mode && h->variant != 1 && h->device < kMaxConstDevices) {   // weights in the __constant__ array
        const Legendre1DParams<T> CP = l1_params<T>(h);            // centre / scale / axis / degree; weights unused
        return with_const_weights<T>(h, static_cast<const T*>(h->l1_wt_dev), h->l1_degree + 1, st, [&]() {
          if (pk == PYVES_POT_DOUBLE_WELL_2D && h->l1_degree == 5) return run_pair_c<T, DW, BiasLegendre1D<T, 5, true>>(h, A, CP, st);
          if (pk == PYVES_POT_SLIP_BOND_2D && h->l1_degree == 12) return run_pair_c<T, SL, BiasLegendre1D<T, 12, true>>(h, A, CP, st);
          return run_pair_c<T, GP, BiasLegendre1D<T, 0, true>>(h, A, CP, st);
        });
      }
      { const int rc = refresh_l1_host(h, st); if (rc) return rc; }
      const Legendre1DParams<T> BP = l1_params<T>(h);
      if (pk == PYVES_POT_DOUBLE_WELL_2D && h->l1_degree == 5) return run_pair<T, DW, BiasLegendre1D<T, 5>>(h, A, BP, st);
      if (pk == PYVES_POT_SLIP_BOND_2D && h->l1_degree == 12) return run_pair<T, SL, BiasLegendre1D<T, 12>>(h, A, BP, st);
      return run_pair<T, GP, BiasLegendre1D<T, 0>>(h, A, BP, st);
    }
    case PYVES_BIAS_LEGENDRE_2D: {
      const int kx = h->l2_dx + 1, ky = h->l2_dy + 1;
      // FP32 default: the tensor-core kernel (6.4 ms per 10^6 x 500 steps for any basis up to 16 x 16); the one CUDA-core kernel
      // that beats it is the static 8 x 8 column sweep on the double well (3.9 ms)
      if (h->variant == 7 || (h->variant == -1 && !(kx == 8 && ky == 8 && pk == PYVES_POT_DOUBLE_WELL_2D))) {
        const int rc = run_tc<T>(h, A, st);
        if (rc <= 0) return rc;
      }
      const Legendre2DSmemParams<T> BP = l2_params<T>(h);
      // FP64 16 x 16: the static functor feeds every DFMA with its own LDC.64 (196 per step, ADU bound); the
      // shared-memory column sweep is 19 % faster there (1.83e10 vs 1.54e10 walker-steps/s), so doubles skip it
      const bool static_ok = !(sizeof(T) == 8 && kx == 16 && ky == 16 && h->variant != 1);
      if (pk == PYVES_POT_DOUBLE_WELL_2D && h->variant != 2 && static_ok) {   // variant 2: runtime-size functor for tuning runs
        if (h->l2_dev_mode && ((kx == 8 && ky == 8) || (kx == 16 && ky == 16)) && h->device < kMaxConstDevices) {
          // coefficients were set from a device pointer: the kernels read them from the __constant__ array
          return with_const_weights<T>(h, static_cast<const T*>(h->l2_wt_dev), kx * ky, st, [&]() {
            if (kx == 8) return run_fast<T, DW, BiasLegendre2DStaticT<T, 8, 8, true>, 2, 1>(h, A, l2_static_tc_params<T, 8, 8>(h), st);
            return run_fast<T, DW, BiasLegendre2DStaticT<T, 16, 16, true, kHybrid16>, 1, 1>(h, A, l2_static_tc_params<T, 16, 16>(h), st);
          });
        }
        // variant 1 selects the row-major functor (kept for tuning runs); default = column sweep
        if (kx == 8 && ky == 8 && !h->l2_dev_mode) {
          if (h->variant == 1) return run_fast<T, DW, BiasLegendre2DStatic<T, 8, 8>, 2, 1>(h, A, l2_static_params<T, 8, 8>(h), st);
          // hybrid constant / shared weights (the 16 x 16 trick) do not pay at 8 x 8: 8.07-8.14 ms against 7.87 ms per 10^6 x 1000 steps
          return run_fast<T, DW, BiasLegendre2DStaticT<T, 8, 8>, 2, 1>(h, A, l2_static_t_params<T, 8, 8>(h), st);
        }
        if (kx == 16 && ky == 16 && !h->l2_dev_mode) {
          if (h->variant == 1) return run_fast<T, DW, BiasLegendre2DStatic<T, 16, 16>, 1, 1>(h, A, l2_static_params<T, 16, 16>(h), st);
          return run_fast<T, DW, BiasLegendre2DStaticT<T, 16, 16, false, kHybrid16>, 1, 1>(h, A, l2_static_t_params<T, 16, 16>(h), st);
        }
        if (kx > 16 && kx <= 32) return run_fast<T, DW, BiasLegendre2DSmem<T, 32, 1, 5>, 1, 2>(h, A, BP, st);
        if (kx > 32) return run_fast<T, DW, BiasLegendre2DSmem<T, 32, 2, 5>, 1, 2>(h, A, BP, st);
      }
      // any other potential with a 16 x 16 basis: the same static hybrid sweep behind the runtime potential switch
      if (kx == 16 && ky == 16 && !h->l2_dev_mode && static_ok && h->variant != 2 && pk != PYVES_POT_DOUBLE_WELL_2D)
        return run_fast<T, GP, BiasLegendre2DStaticT<T, 16, 16, false, kHybrid16>, 1, 1>(h, A, l2_static_t_params<T, 16, 16>(h), st);
      if (kx <= 16) return run_fast<T, GP, BiasLegendre2DSmem<T, 16, 1, 7>, 1, 3>(h, A, BP, st);
      if (kx <= 32) return run_fast<T, GP, BiasLegendre2DSmem<T, 32, 1, 5>, 1, 2>(h, A, BP, st);
      return run_fast<T, GP, BiasLegendre2DSmem<T, 32, 2, 5>, 1, 2>(h, A, BP, st);
    }
    case PYVES_BIAS_MLP_1D: {
      const MlpParams<T> BP = mlp_params<T>(h);
      const bool relu = h->mlp_act == PYVES_ACT_RELU;
      const bool sb = pk == PYVES_POT_SZABO_BEREZHKOVSKII && relu;
      // ReLU nets with two / three hidden layers, FP32 default: dV/ds is piecewise constant in the one input -- a binary search
      // (variant 8 forces it for FP64 handles too; 6 / 7 keep the CUDA-core / tensor-core evaluation of the network)
      if (h->pwl_valid && ((h->variant == -1 && sizeof(T) == 4) || h->variant == 8)) {
        { const int rc = ensure_pwl(h, st); if (rc) return rc; }
        if (pk == PYVES_POT_SZABO_BEREZHKOVSKII) return run_fast<T, SB, BiasPwl1D<T>, 2, 1>(h, A, pwl_params<T>(h), st);
        return run_fast<T, GP, BiasPwl1D<T>, 2, 1>(h, A, pwl_params<T>(h), st);
      }
      if (h->mlp_H > 0 && (h->variant == -1 || h->variant == 7)) {      // one H x H layer: on the tensor cores (variant 6: CUDA cores)
        const int rc = run_tc_mlp<T>(h, A, st);
        if (rc <= 0) return rc;
      }
      switch (h->mlp_H) {
        case 0:
          if (sb) return run_fast<T, SB, BiasMlp<T, 0, PYVES_ACT_RELU>, 2, 1>(h, A, BP, st);
          if (relu) return run_fast<T, GP, BiasMlp<T, 0, PYVES_ACT_RELU>, 2, 1>(h, A, BP, st);
          return run_fast<T, GP, BiasMlp<T, 0, PYVES_ACT_TANH>, 2, 1>(h, A, BP, st);
        case 32:
          if (relu) return run_fast<T, GP, BiasMlp<T, 32, PYVES_ACT_RELU>, 1, 1>(h, A, BP, st);
          return run_fast<T, GP, BiasMlp<T, 32, PYVES_ACT_TANH>, 1, 1>(h, A, BP, st);
        case 48:
          if (sb) return run_fast<T, SB, BiasMlp<T, 48, PYVES_ACT_RELU>, 1, 1>(h, A, BP, st);
          if (relu) return run_fast<T, GP, BiasMlp<T, 48, PYVES_ACT_RELU>, 1, 1>(h, A, BP, st);
          return run_fast<T, GP, BiasMlp<T, 48, PYVES_ACT_TANH>, 1, 1>(h, A, BP, st);
        case 64:
          if (relu) return run_fast<T, GP, BiasMlp<T, 64, PYVES_ACT_RELU>, 1, 1>(h, A, BP, st);
          return run_fast<T, GP, BiasMlp<T, 64, PYVES_ACT_TANH>, 1, 1>(h, A, BP, st);
        default: return run_general<T>(h, A, st);
      }
    }
    default: return run_fast<T, GP, BiasOther<T>, 1, 1>(h, A, other_params<T>(h), st);
  }
}

bool valid_axis(int axis) { return axis == 0 || axis == 1; }

}  // namespace

template <typename T>
static int moments_impl(pyves_ctx* h, const void* pos, int64_t n, int ncols, const void* w, double* sum_w, double* sum_wf, double* sum_wff, cudaStream_t st) {
  RowSource<T> S;
  if (pos) {
    S.pos = static_cast<const T*>(pos);
    S.n = n;
    S.stride_row = ncols;
    S.stride_col = 1;
    S.ncols = ncols;
  } else {
    S.pos = static_cast<const T*>(h->pos);
    S.n = h->n;
    S.stride_row = 1;
    S.stride_col = h->n;
    S.ncols = h->dims;
  }
  S.w = static_cast<const T*>(w);
  const int64_t K = basis_size(h);
  int rc = ensure_scratch(h, moments_scratch_doubles(K, sum_wff != nullptr, S.n, S.w != nullptr));
  if (rc) return rc;
  int launches = 0;
  if (h->bias_kind == PYVES_BIAS_MLP_1D) {
    if (sum_wff) return fail(h, PYVES_EUNSUPPORTED, "second moments are not defined for the MLP bias");
    MlpGradParams<T> M;
    M.axis = h->mlp_axis;
    M.act = h->mlp_act;
    M.n_layers = (int)h->mlp_sizes.size() + 1;
    M.sizes[0] = 1;
    for (size_t i = 0; i < h->mlp_sizes.size(); ++i) M.sizes[i + 1] = h->mlp_sizes[i];
    M.sizes[h->mlp_sizes.size() + 1] = 1;
    M.lo = (T)h->mlp_lo;
    M.inv_range = (T)(1.0 / (h->mlp_hi - h->mlp_lo));
    M.theta = static_cast<const T*>(h->mlp_theta_dev);
    M.n_params = (int)K;
    M.inside_only = h->mom_inside;
    // ReLU nets: rows binned into the pieces of the piecewise-linear network, gradients on one pseudo row per piece (moments.cuh);
    // FP32 default for ensembles much larger than the table, variant 8 always
    const int cap = h->pwl_valid ? mlp_pwl_capacity(h->mlp_H, h->mlp_width) : 0;
    if (h->pwl_valid && S.ncols <= 2 && (h->variant == 8 || (h->variant == -1 && sizeof(T) == 4 && S.n >= 8 * (int64_t)(cap + 1)))) {
      { const int rc = ensure_pwl(h, st); if (rc) return rc; }
      if (h->pwl_hist_cap < cap) {
        if (h->pwl_hist) cudaFree(h->pwl_hist);
        if (h->pwl_rows) cudaFree(h->pwl_rows);
        if (h->pwl_row_w) cudaFree(h->pwl_row_w);
        h->pwl_hist = nullptr;
        h->pwl_rows = h->pwl_row_w = nullptr;
        CU(cudaMalloc(reinterpret_cast<void**>(&h->pwl_hist), sizeof(double) * 2 * (size_t)(cap + 1)));
        CU(cudaMalloc(&h->pwl_rows, sizeof(T) * 2 * (size_t)(cap + 1)));
        CU(cudaMalloc(&h->pwl_row_w, sizeof(T) * (size_t)(cap + 1)));
        h->pwl_hist_cap = cap;
      }
      CU(launch_pwl_compress<T>(S, M.axis, M.lo, M.inv_range, M.inside_only, static_cast<const T*>(h->pwl_brk), h->pwl_count, cap, h->pwl_hist,
                                static_cast<T*>(h->pwl_rows), static_cast<T*>(h->pwl_row_w), st, &launches));
      RowSource<T> S2;
      S2.pos = static_cast<const T*>(h->pwl_rows);
      S2.n = cap + 1;
      S2.stride_row = 2;
      S2.stride_col = 1;
      S2.ncols = 2;
      S2.w = static_cast<const T*>(h->pwl_row_w);
      M.inside_only = 0;                           // already applied to the rows; the pseudo rows are means of admitted rows
      CU(launch_moments_mlp<T>(S2, M, sum_w, sum_wf, h->scratch, st, &launches));
    } else {
      CU(launch_moments_mlp<T>(S, M, sum_w, sum_wf, h->scratch, st, &launches));
    }
  } else {
    BasisParams<T> B;
    std::memset(&B, 0, sizeof(B));
    B.kind = h->bias_kind;
    B.kb = 1;
    if (h->bias_kind == PYVES_BIAS_LEGENDRE_1D) {
      const Legendre1DParams<T> P = l1_params<T>(h);
      B.ka = h->l1_degree + 1;
      B.axis = h->l1_axis;
      B.ca = P.centre;
      B.iha = P.inv_half;
    } else if (h->bias_kind == PYVES_BIAS_LEGENDRE_2D) {
      const Legendre2DSmemParams<T> P = l2_params<T>(h);
      B.ka = P.kx;
      B.kb = P.ky;
      B.ca = P.cx;
      B.iha = P.ihx;
      B.cb = P.cy;
      B.ihb = P.ihy;
    } else {
      B.ka = h->pc_degree + 1;
      B.npath = h->pc_npath;
      B.lam = (T)h->pc_lam;
      B.path = static_cast<const T*>(h->pc_path_dev);
    }
    B.inside_only = h->mom_inside;
    CU(launch_moments_basis<T>(S, B, sum_w, sum_wf, sum_wff, h->scratch, st, &launches));
  }
  g_launches += launches;
  return PYVES_OK;
}

template <typename T> static int set_l2_device(pyves_ctx* h, const double* w_dev, int kx, int ky, cudaStream_t st) {
  const size_t K = (size_t)kx * ky;
  if (h->l2_dev_count != K || !h->l2_w_dev || !h->l2_wt_dev) {
    if (h->l2_w_dev) cudaFree(h->l2_w_dev);
    if (h->l2_wt_dev) cudaFree(h->l2_wt_dev);
    h->l2_w_dev = h->l2_wt_dev = nullptr;
    h->alloc_count.erase(&h->l2_w_dev);
    CU(cudaMalloc(&h->l2_w_dev, K * sizeof(T)));
    CU(cudaMalloc(&h->l2_wt_dev, K * sizeof(T)));
    h->l2_dev_count = K;
  }
  CU(launch_convert_weights2d<T>(w_dev, kx, ky, static_cast<T*>(h->l2_w_dev), static_cast<T*>(h->l2_wt_dev), st));
  g_launches++;
  h->tc_img_valid = false;
  return PYVES_OK;
}

// =================================================================================================
extern "C" {

int pyves_abi_version(void) { return PYVES_ABI_VERSION; }
int64_t pyves_launch_count(void) { return g_launches.load(); }

const char* pyves_last_error(const pyves_t* h) { return h ? h->err.c_str() : g_err.c_str(); }

int pyves_create(pyves_t** out, int device, int64_t n_walkers, int dims, uint64_t seed, int precision) {
  pyves_ctx* h = nullptr;
  if (!out) return fail(h, PYVES_EINVAL, "out is NULL");
  *out = nullptr;
  if (n_walkers < 1) return fail(h, PYVES_EINVAL, "n_walkers must be >= 1");
  if (dims != 2 && dims != 3) return fail(h, PYVES_EINVAL, "dims must be 2 or 3");
  if (precision != PYVES_F32 && precision != PYVES_F64) return fail(h, PYVES_EINVAL, "precision must be PYVES_F32 or PYVES_F64");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(h, PYVES_ECUDA, "no CUDA device available");
  }
  if (device < 0 || device >= ndev) return fail(h, PYVES_EINVAL, "invalid device ordinal");
  h = new (std::nothrow) pyves_ctx();
  if (!h) return fail(nullptr, PYVES_ENOMEM, "out of host memory");
  h->device = device;
  h->n = n_walkers;
  h->dims = dims;
  h->seed = seed;
  h->prec = precision;
  h->esz = precision == PYVES_F64 ? 8 : 4;
  DeviceGuard g(device);
  const size_t bytes = (size_t)n_walkers * dims * h->esz;
  cudaError_t e = cudaMalloc(&h->pos, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->vel, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->fail_dev, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->pos, 0, bytes);
  if (e == cudaSuccess) e = cudaMemset(h->vel, 0, bytes);
  if (e != cudaSuccess) {
    const std::string msg = std::string("pyves_create: ") + cudaGetErrorString(e);
    pyves_destroy(h);
    return fail(nullptr, e == cudaErrorMemoryAllocation ? PYVES_ENOMEM : PYVES_ECUDA, msg);
  }
  *out = h;
  return PYVES_OK;
}

int pyves_destroy(pyves_t* h) {
  if (!h) return PYVES_OK;
  DeviceGuard g(h->device);
  void* ptrs[] = {h->pos, h->vel, h->l2_w_dev, h->l2_wt_dev, h->l1_wt_dev, h->l1_w64_dev, h->mlp_theta_dev, h->mlp_packed_dev, h->pc_path_dev,
                  h->scratch, h->stage, h->fail_dev, h->tc_img, h->pwl_brk, h->pwl_slope, h->pwl_count, h->pwl_hist, h->pwl_rows, h->pwl_row_w, h->pwl_grid, h->pwl_grid_map};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete h;
  return PYVES_OK;
}

int pyves_set_integrator(pyves_t* h, double mass, double kT, double friction, double dt, int scheme) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!(mass > 0) || !(kT >= 0) || !(friction >= 0) || !(dt > 0)) return fail(h, PYVES_EINVAL, "mass, dt must be > 0; kT, friction >= 0");
  if (scheme != PYVES_SCHEME_OPENMM_LANGEVIN && scheme != PYVES_SCHEME_MIDDLE) return fail(h, PYVES_EINVAL, "unknown integration scheme");
  h->mass = mass;
  h->kT = kT;
  h->friction = friction;
  h->dt = dt;
  h->scheme = scheme;
  h->have_int = true;
  return PYVES_OK;
}

int pyves_set_potential(pyves_t* h, int kind, const double* params, int nparams) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  PotParams<double> P;
  std::memset(&P, 0, sizeof(P));
  P.kind = kind;
  auto need = [&](int k) { return params != nullptr && nparams == k; };
  switch (kind) {
    case PYVES_POT_SINGLE_WELL_1D:
    case PYVES_POT_DOUBLE_WELL_1D:
    case PYVES_POT_SINGLE_WELL_2D:
    case PYVES_POT_DOUBLE_WELL_2D:
    case PYVES_POT_MULLER_BROWN:
      if (nparams != 0) return fail(h, PYVES_EINVAL, "this potential takes no parameters");
      break;
    case PYVES_POT_SLIP_BOND_2D:
    case PYVES_POT_CATCH_BOND_2D: {
      const int k = kind == PYVES_POT_SLIP_BOND_2D ? 7 : 11;
      if (!need(k)) return fail(h, PYVES_EINVAL, "slip/catch bond potential needs 7/11 parameters");
      for (int i = 0; i < k; ++i) P.p[i] = params[i];
      if (params[3] == 0 || params[6] == 0) return fail(h, PYVES_EINVAL, "y_scale and xy_scale must be non-zero");
      P.p[3] = 1.0 / params[3];   // kernels multiply by the reciprocal scales
      P.p[6] = 1.0 / params[6];
      if (k == 11) {
        if (params[8] == 0 || params[10] == 0) return fail(h, PYVES_EINVAL, "gx_scale and gy_scale must be non-zero");
        P.p[8] = 1.0 / params[8];
        P.p[10] = 1.0 / params[10];
      }
    } break;
    case PYVES_POT_SZABO_BEREZHKOVSKII:
      if (!need(4)) return fail(h, PYVES_EINVAL, "Szabo-Berezhkovskii potential needs 4 parameters");
      for (int i = 0; i < 4; ++i) P.p[i] = params[i];
      break;
    default: return fail(h, PYVES_EINVAL, "unknown potential kind");
  }
  h->pot = P;
  h->have_pot = true;
  return PYVES_OK;
}

int pyves_set_walker_offset(pyves_t* h, int64_t gid0) {
  if (!h || gid0 < 0) return fail(h, PYVES_EINVAL, "bad walker offset");
  h->gid0 = (uint64_t)gid0;
  return PYVES_OK;
}

int pyves_set_step_counter(pyves_t* h, int64_t step) {
  if (!h || step < 0) return fail(h, PYVES_EINVAL, "bad step counter");
  h->step = step;
  return PYVES_OK;
}

int pyves_get_step_counter(const pyves_t* h, int64_t* step) {
  if (!h || !step) return PYVES_EINVAL;
  *step = h->step;
  return PYVES_OK;
}

int pyves_set_moment_domain(pyves_t* h, int inside_only) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (inside_only != 0 && inside_only != 1) return fail(h, PYVES_EINVAL, "inside_only must be 0 or 1");
  h->mom_inside = inside_only;
  return PYVES_OK;
}

int pyves_averaged_sgd_step(int device, int K, const double* sum_w, const double* sum_wf, const double* sum_wff, const double* f_target,
                            const double* alpha, const double* abar, const double* scale, double mu, double beta, int averaging,
                            double decay, int64_t n_after, double* alpha_out, double* abar_out, void* stream) {
  pyves_ctx* h = nullptr;
  if (K < 1 || !sum_w || !sum_wf || !f_target || !alpha || !abar || !alpha_out || !abar_out) return fail(h, PYVES_EINVAL, "averaged_sgd_step: NULL argument or K < 1");
  if (alpha_out == alpha || abar_out == abar || alpha_out == abar || abar_out == alpha) return fail(h, PYVES_EINVAL, "averaged_sgd_step: outputs must not alias the inputs");
  if (averaging != 0 && averaging != 1) return fail(h, PYVES_EINVAL, "averaged_sgd_step: averaging must be 0 (ewma) or 1 (running)");
  DeviceGuard g(device);
  cudaError_t e = launch_averaged_sgd_step(K, sum_w, sum_wf, sum_wff, f_target, alpha, abar, scale, mu, beta, averaging, decay,
                                           (long long)n_after, alpha_out, abar_out, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(h, PYVES_ECUDA, cudaGetErrorString(e));
  g_launches++;
  return PYVES_OK;
}

int pyves_iterate_pinned(pyves_t* h, const void* pos_in, const void* vel_in, int64_t nsteps, double* sum_w, double* sum_wf, void* pos_out,
                         void* vel_out, void* stream) {
  if (!h) return PYVES_EINVAL;
  if (!pos_in || !vel_in || !sum_w || !sum_wf || !pos_out || !vel_out) return fail(h, PYVES_EINVAL, "iterate_pinned: NULL buffer");
  int rc = pyves_set_state(h, pos_in, vel_in, 2, stream);
  if (rc) return rc;
  rc = pyves_step(h, nsteps, nullptr, stream);
  if (rc) return rc;
  rc = pyves_moments(h, nullptr, 0, 0, nullptr, sum_w, sum_wf, nullptr, 0, stream);
  if (rc) return rc;
  return pyves_get_state(h, pos_out, vel_out, 2, stream);
}

int pyves_step_wave_walkers(pyves_t* h, int64_t* walkers) {
  if (!h || !walkers) return PYVES_EINVAL;
  *walkers = 0;
  if (h->prec != PYVES_F32 || h->dims != 2 || h->scheme != PYVES_SCHEME_OPENMM_LANGEVIN) return PYVES_OK;
  int per_cta = 0;
  if (h->bias_kind == PYVES_BIAS_LEGENDRE_2D) {
    const int kx = h->l2_dx + 1, ky = h->l2_dy + 1;
    const bool tc = h->variant == 7 || (h->variant == -1 && !(kx == 8 && ky == 8 && h->pot.kind == PYVES_POT_DOUBLE_WELL_2D));
    if (tc) per_cta = langevin2d_tc_walkers_per_cta(kx, ky);
  } else if (h->bias_kind == PYVES_BIAS_MLP_1D && h->mlp_H > 0 && (h->variant == 7 || (h->variant == -1 && !h->pwl_valid))) {
    per_cta = langevin2d_tc_mlp_walkers_per_cta(h->mlp_H, h->mlp_act);
  }
  if (per_cta > 0) {
    int n_sm = 0;
    DeviceGuard g(h->device);
    CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device));
    *walkers = (int64_t)n_sm * per_cta;
  }
  return PYVES_OK;
}

int pyves_set_tuning(pyves_t* h, int variant) {
  if (!h) return PYVES_EINVAL;
  if (variant == 98 || variant == 97) {   // 98: covariance on CUDA cores only; 97: tensor cores allowed again
    set_no_tensor_cores(variant == 98);
    return PYVES_OK;
  }
  if (variant == 96 || variant == 95) {   // 96: first moments of large 2D bases with the one-function-per-thread kernel; 95: register-tiled kernel again
    set_tiled_first_moments(variant == 95);
    return PYVES_OK;
  }
  if (variant >= 1000 && variant <= 1030) {   // rows per CTA of the tensor-core covariance: 2^(variant - 1000)
    cov_tc_set_max_rows((int64_t)1 << (variant - 1000));
    return PYVES_OK;
  }
  h->variant = variant;
  return PYVES_OK;
}

int pyves_set_bias_none(pyves_t* h) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  h->bias_kind = PYVES_BIAS_NONE;
  return PYVES_OK;
}

int pyves_set_bias_legendre1d(pyves_t* h, int axis, int degree, double lo, double hi, const double* w) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!valid_axis(axis)) return fail(h, PYVES_EINVAL, "Invalid axis");
  if (degree < 0 || degree > PYVES_MAX_DEGREE) return fail(h, PYVES_EINVAL, "degree must be in [0, 63]");
  if (!(hi > lo) || !w) return fail(h, PYVES_EINVAL, "need max > min and weights");
  h->l1_axis = axis;
  h->l1_degree = degree;
  h->l1_lo = lo;
  h->l1_hi = hi;
  std::memset(h->l1_w, 0, sizeof(h->l1_w));
  for (int i = 0; i <= degree; ++i) h->l1_w[i] = w[i];
  h->l1_dev_mode = false;
  h->l1_host_stale = false;
  h->bias_kind = PYVES_BIAS_LEGENDRE_1D;
  return PYVES_OK;
}

int pyves_set_bias_legendre1d_device(pyves_t* h, int axis, int degree, double lo, double hi, const double* w_dev, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!valid_axis(axis)) return fail(h, PYVES_EINVAL, "Invalid axis");
  if (degree < 0 || degree > PYVES_MAX_DEGREE) return fail(h, PYVES_EINVAL, "degree must be in [0, 63]");
  if (!(hi > lo) || !w_dev) return fail(h, PYVES_EINVAL, "need max > min and weights");
  DeviceGuard g(h->device);
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, w_dev) != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
    cudaGetLastError();
    return fail(h, PYVES_EINVAL, "w_dev must be a device pointer (FP64, degree + 1 entries)");
  }
  if (!h->l1_wt_dev) {
    CU(cudaMalloc(&h->l1_wt_dev, kMaxK1 * sizeof(double)));
    CU(cudaMalloc(reinterpret_cast<void**>(&h->l1_w64_dev), kMaxK1 * sizeof(double)));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->prec == PYVES_F64) {
    CU(launch_convert_weights1d<double>(w_dev, degree + 1, static_cast<double*>(h->l1_wt_dev), h->l1_w64_dev, st));
  } else {
    CU(launch_convert_weights1d<float>(w_dev, degree + 1, static_cast<float*>(h->l1_wt_dev), h->l1_w64_dev, st));
  }
  g_launches++;
  h->l1_axis = axis;
  h->l1_degree = degree;
  h->l1_lo = lo;
  h->l1_hi = hi;
  h->l1_dev_mode = true;
  h->l1_host_stale = true;
  h->const_tag = g_const_tag.fetch_add(1);
  h->bias_kind = PYVES_BIAS_LEGENDRE_1D;
  return PYVES_OK;
}

int pyves_set_bias_legendre2d(pyves_t* h, int deg_x, int deg_y, const double* mm, const double* w, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (deg_x < 0 || deg_x > PYVES_MAX_DEGREE || deg_y < 0 || deg_y > PYVES_MAX_DEGREE) return fail(h, PYVES_EINVAL, "degrees must be in [0, 63]");
  if (!mm || !w || !(mm[1] > mm[0]) || !(mm[3] > mm[2])) return fail(h, PYVES_EINVAL, "need max > min per axis and weights");
  DeviceGuard g(h->device);
  const size_t K = (size_t)(deg_x + 1) * (deg_y + 1);
  h->l2_dx = deg_x;
  h->l2_dy = deg_y;
  std::memcpy(h->l2_mm, mm, sizeof(h->l2_mm));
  h->l2_w.assign(w, w + K);
  const int rc = upload(h, &h->l2_w_dev, w, K, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  h->l2_dev_mode = false;
  h->l2_dev_count = 0;           // upload() reallocated l2_w_dev
  h->tc_img_valid = false;
  h->bias_kind = PYVES_BIAS_LEGENDRE_2D;
  return PYVES_OK;
}

int pyves_set_bias_legendre2d_device(pyves_t* h, int deg_x, int deg_y, const double* mm, const double* w_dev, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (deg_x < 0 || deg_x > PYVES_MAX_DEGREE || deg_y < 0 || deg_y > PYVES_MAX_DEGREE) return fail(h, PYVES_EINVAL, "degrees must be in [0, 63]");
  if (!mm || !w_dev || !(mm[1] > mm[0]) || !(mm[3] > mm[2])) return fail(h, PYVES_EINVAL, "need max > min per axis and weights");
  DeviceGuard g(h->device);
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, w_dev) != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
    cudaGetLastError();
    return fail(h, PYVES_EINVAL, "w_dev must be a device pointer (FP64, [deg_x + 1][deg_y + 1])");
  }
  h->l2_dx = deg_x;
  h->l2_dy = deg_y;
  std::memcpy(h->l2_mm, mm, sizeof(h->l2_mm));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rc = h->prec == PYVES_F64 ? set_l2_device<double>(h, w_dev, deg_x + 1, deg_y + 1, st)
                                      : set_l2_device<float>(h, w_dev, deg_x + 1, deg_y + 1, st);
  if (rc) return rc;
  h->l2_dev_mode = true;
  h->const_tag = g_const_tag.fetch_add(1);
  h->l2_w.clear();               // no host copy in this mode
  h->bias_kind = PYVES_BIAS_LEGENDRE_2D;
  return PYVES_OK;
}

int pyves_set_bias_like(pyves_t* h, const pyves_t* leader, void* stream) {
  if (!h || !leader) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (h == leader) return PYVES_OK;
  if (h->device != leader->device || h->prec != leader->prec) return fail(h, PYVES_EINVAL, "set_bias_like: handles must share device and precision");
  const bool l1 = leader->bias_kind == PYVES_BIAS_LEGENDRE_1D && leader->l1_dev_mode;
  const bool l2 = leader->bias_kind == PYVES_BIAS_LEGENDRE_2D && leader->l2_dev_mode;
  if (!l1 && !l2) return fail(h, PYVES_ESTATE, "set_bias_like: the leader's bias was not set by pyves_set_bias_legendre{1d,2d}_device");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (l1) {
    if (!h->l1_wt_dev) {
      CU(cudaMalloc(&h->l1_wt_dev, kMaxK1 * sizeof(double)));
      CU(cudaMalloc(reinterpret_cast<void**>(&h->l1_w64_dev), kMaxK1 * sizeof(double)));
    }
    CU(cudaMemcpyAsync(h->l1_wt_dev, leader->l1_wt_dev, kMaxK1 * h->esz, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(h->l1_w64_dev, leader->l1_w64_dev, kMaxK1 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    h->l1_axis = leader->l1_axis;
    h->l1_degree = leader->l1_degree;
    h->l1_lo = leader->l1_lo;
    h->l1_hi = leader->l1_hi;
    h->l1_dev_mode = true;
    h->l1_host_stale = true;
  } else {
    const size_t K = (size_t)(leader->l2_dx + 1) * (leader->l2_dy + 1);
    if (h->l2_dev_count != K || !h->l2_w_dev || !h->l2_wt_dev) {
      if (h->l2_w_dev) cudaFree(h->l2_w_dev);
      if (h->l2_wt_dev) cudaFree(h->l2_wt_dev);
      h->l2_w_dev = h->l2_wt_dev = nullptr;
      h->l2_dev_count = 0;
      h->alloc_count.erase(&h->l2_w_dev);
      CU(cudaMalloc(&h->l2_w_dev, K * h->esz));
      CU(cudaMalloc(&h->l2_wt_dev, K * h->esz));
      h->l2_dev_count = K;
    }
    CU(cudaMemcpyAsync(h->l2_w_dev, leader->l2_w_dev, K * h->esz, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(h->l2_wt_dev, leader->l2_wt_dev, K * h->esz, cudaMemcpyDeviceToDevice, st));
    h->l2_dx = leader->l2_dx;
    h->l2_dy = leader->l2_dy;
    std::memcpy(h->l2_mm, leader->l2_mm, sizeof(h->l2_mm));
    h->l2_dev_mode = true;
    h->l2_w.clear();
    // the operand image of the tensor-core step kernel: built ONCE by the leader, copied to every follower (a follower that built
    // its own in front of its step kernel would wait for an SM that the previous shard's step kernel still holds)
    h->tc_img_valid = false;
    if (h->prec == PYVES_F32) {
      pyves_ctx* L = const_cast<pyves_ctx*>(leader);
      if (!L->tc_img_valid) {
        const int rc = refresh_tc_img(L, st);
        if (rc) return rc;
      }
      if (L->tc_img_valid) {
        if (!h->tc_img) CU(cudaMalloc(&h->tc_img, langevin2d_tc_scratch_bytes()));
        CU(cudaMemcpyAsync(h->tc_img, L->tc_img, langevin2d_tc_scratch_bytes(), cudaMemcpyDeviceToDevice, st));
        h->tc_img_valid = true;
      }
    }
  }
  h->const_tag = leader->const_tag;   // same content: a launch of either handle finds the __constant__ array up to date
  h->bias_kind = leader->bias_kind;
  return PYVES_OK;
}

int pyves_set_bias_mlp(pyves_t* h, int axis, double lo, double hi, int n_hidden, const int* sizes, int act, const double* theta, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!valid_axis(axis)) return fail(h, PYVES_EINVAL, "Invalid axis");
  if (n_hidden < 1) return fail(h, PYVES_EINVAL, "NNBasis needs at least one hidden layer");
  if (n_hidden > 8) return fail(h, PYVES_EUNSUPPORTED, "at most 8 hidden layers are supported");
  if (act != PYVES_ACT_RELU && act != PYVES_ACT_TANH) return fail(h, PYVES_EUNSUPPORTED, "activation must be ReLU or Tanh");
  if (!sizes || !theta || hi == lo) return fail(h, PYVES_EINVAL, "need sizes, parameters and max != min");
  for (int i = 0; i < n_hidden; ++i)
    if (sizes[i] < 1) return fail(h, PYVES_EINVAL, "hidden layer sizes must be >= 1");
  DeviceGuard g(h->device);
  std::vector<int> dims(1, 1);
  dims.insert(dims.end(), sizes, sizes + n_hidden);
  dims.push_back(1);
  size_t np = 0;
  std::vector<size_t> woff, boff;
  for (size_t l = 0; l + 1 < dims.size(); ++l) {
    woff.push_back(np);
    np += (size_t)dims[l] * dims[l + 1];
    boff.push_back(np);
    np += dims[l + 1];
  }
  h->mlp_axis = axis;
  h->mlp_act = act;
  h->mlp_lo = lo;
  h->mlp_hi = hi;
  h->mlp_sizes.assign(sizes, sizes + n_hidden);
  h->mlp_np = np;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = upload(h, &h->mlp_theta_dev, theta, np, st);
  if (rc) return rc;
  // fold the activation-free pair Linear(1,h0) -> Linear(h0,h1)  (ves/basis.py:197-202)
  const int L = n_hidden;
  h->mlp_linear = 0;
  h->mlp_H = -1;
  h->mlp_width = 0;
  const double* W0 = theta + woff[0];
  const double* b0 = theta + boff[0];
  std::vector<double> packed;
  if (L == 1) {
    const double* Wo = theta + woff[1];
    double u0 = 0, bo = theta[boff[1]];
    for (int j = 0; j < dims[1]; ++j) {
      u0 += Wo[j] * W0[j];
      bo += Wo[j] * b0[j];
    }
    h->mlp_linear = 1;
    h->mlp_u0 = u0;
    h->mlp_bout = bo;
    h->mlp_H = 0;
  } else if (L == 2 || L == 3) {
    const int h0 = dims[1], h1 = dims[2];
    const double* W1 = theta + woff[1];
    const double* b1 = theta + boff[1];
    std::vector<double> u(h1), c(h1);
    for (int j = 0; j < h1; ++j) {
      double su = 0, sc = b1[j];
      for (int i = 0; i < h0; ++i) {
        su += W1[j * h0 + i] * W0[i];
        sc += W1[j * h0 + i] * b0[i];
      }
      u[j] = su;
      c[j] = sc;
    }
    if (L == 2) {
      const int Wp = (h1 + 3) / 4 * 4;
      packed.assign((size_t)3 * Wp + 4, 0.0);
      const double* Wo = theta + woff[2];
      for (int j = 0; j < h1; ++j) {
        packed[j] = u[j];
        packed[Wp + j] = c[j];
        packed[2 * Wp + j] = Wo[j];
      }
      packed[(size_t)3 * Wp] = theta[boff[2]];
      h->mlp_bout = theta[boff[2]];
      h->mlp_width = Wp;
      h->mlp_H = 0;
    } else {
      const int h2 = dims[3];
      const int m = h1 > h2 ? h1 : h2;
      const int H = m <= 32 ? 32 : (m <= 48 ? 48 : (m <= 64 ? 64 : -1));
      if (H > 0) {
        packed.assign((size_t)4 * H + (size_t)H * H + 4, 0.0);
        const double* W2 = theta + woff[2];
        const double* b2 = theta + boff[2];
        const double* Wo = theta + woff[3];
        for (int j = 0; j < h1; ++j) {
          packed[j] = u[j];
          packed[H + j] = c[j];
        }
        for (int j = 0; j < h2; ++j) {
          packed[2 * H + j] = Wo[j];
          packed[3 * H + j] = b2[j];
          for (int i = 0; i < h1; ++i) packed[4 * H + (size_t)i * H + j] = W2[j * h1 + i];   // transposed: W2T[i][j]
        }
        packed[(size_t)4 * H + (size_t)H * H] = theta[boff[3]];
        h->mlp_bout = theta[boff[3]];
        h->mlp_H = H;
      }
    }
  }
  rc = upload(h, &h->mlp_packed_dev, packed.data(), packed.size(), st);
  if (rc) return rc;
  h->bias_kind = PYVES_BIAS_MLP_1D;
  return refresh_pwl(h, st);
}

int pyves_set_bias_mlp_device(pyves_t* h, int axis, double lo, double hi, int n_hidden, const int* sizes, int act, const double* theta_dev,
                              void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!valid_axis(axis)) return fail(h, PYVES_EINVAL, "Invalid axis");
  if (n_hidden < 1) return fail(h, PYVES_EINVAL, "NNBasis needs at least one hidden layer");
  if (n_hidden < 2) return fail(h, PYVES_EUNSUPPORTED, "device-resident NN parameters need at least two hidden layers (a single one is the linear case: use pyves_set_bias_mlp)");
  if (n_hidden > 8) return fail(h, PYVES_EUNSUPPORTED, "at most 8 hidden layers are supported");
  if (act != PYVES_ACT_RELU && act != PYVES_ACT_TANH) return fail(h, PYVES_EUNSUPPORTED, "activation must be ReLU or Tanh");
  if (!sizes || !theta_dev || hi == lo) return fail(h, PYVES_EINVAL, "need sizes, parameters and max != min");
  for (int i = 0; i < n_hidden; ++i)
    if (sizes[i] < 1) return fail(h, PYVES_EINVAL, "hidden layer sizes must be >= 1");
  DeviceGuard g(h->device);
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, theta_dev) != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
    cudaGetLastError();
    return fail(h, PYVES_EINVAL, "theta_dev must be a device pointer (FP64, torch parameters() order)");
  }
  std::vector<int> dims(1, 1);
  dims.insert(dims.end(), sizes, sizes + n_hidden);
  dims.push_back(1);
  size_t np = 0;
  std::vector<size_t> woff, boff;
  for (size_t l = 0; l + 1 < dims.size(); ++l) {
    woff.push_back(np);
    np += (size_t)dims[l] * dims[l + 1];
    boff.push_back(np);
    np += dims[l + 1];
  }
  MlpPackPlan Q;
  std::memset(&Q, 0, sizeof(Q));
  Q.np = (int)np;
  Q.h0 = dims[1];
  Q.h1 = dims[2];
  Q.w0 = (int)woff[0]; Q.b0 = (int)boff[0]; Q.w1 = (int)woff[1]; Q.b1 = (int)boff[1];
  int H = -1, width = 0;
  if (n_hidden == 2) {
    width = (dims[2] + 3) / 4 * 4;
    H = 0;
    Q.wo = (int)woff[2]; Q.bo = (int)boff[2];
    Q.n_packed = 3 * width + 4;
  } else if (n_hidden == 3) {
    const int m = dims[2] > dims[3] ? dims[2] : dims[3];
    H = m <= 32 ? 32 : (m <= 48 ? 48 : (m <= 64 ? 64 : -1));
    if (H > 0) {
      Q.h2 = dims[3];
      Q.w2 = (int)woff[2]; Q.b2 = (int)boff[2]; Q.wo = (int)woff[3]; Q.bo = (int)boff[3];
      Q.n_packed = 4 * H + H * H + 4;
    }
  }
  Q.H = H > 0 ? H : 0;
  Q.width = width;
  int rc = ensure_buffer(h, &h->mlp_theta_dev, np);
  if (rc) return rc;
  rc = ensure_buffer(h, &h->mlp_packed_dev, (size_t)Q.n_packed);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->prec == PYVES_F64) {
    CU(launch_mlp_pack<double>(theta_dev, Q, static_cast<double*>(h->mlp_theta_dev), static_cast<double*>(h->mlp_packed_dev), st));
  } else {
    CU(launch_mlp_pack<float>(theta_dev, Q, static_cast<float*>(h->mlp_theta_dev), static_cast<float*>(h->mlp_packed_dev), st));
  }
  g_launches++;
  h->mlp_axis = axis;
  h->mlp_act = act;
  h->mlp_lo = lo;
  h->mlp_hi = hi;
  h->mlp_sizes.assign(sizes, sizes + n_hidden);
  h->mlp_np = np;
  h->mlp_linear = 0;
  h->mlp_H = H;
  h->mlp_width = width;
  h->mlp_bout = 0;
  h->mlp_u0 = 0;
  h->bias_kind = PYVES_BIAS_MLP_1D;
  return refresh_pwl(h, st);
}

int pyves_optimizer_step(int device, int kind, int64_t P, const double* sum_w, const double* sum_wf, const double* f_target, int grad_average,
                         double grad_decay, int64_t n_average, double* grad_state, double* theta, double* state1, double* state2,
                         const double* hyper, int64_t step, void* stream) {
  pyves_ctx* h = nullptr;
  if (P < 1 || !sum_w || !sum_wf || !f_target || !theta || !hyper) return fail(h, PYVES_EINVAL, "optimizer_step: NULL argument or P < 1");
  if (kind != PYVES_OPT_SGD && kind != PYVES_OPT_RMSPROP && kind != PYVES_OPT_ADAM) return fail(h, PYVES_EINVAL, "Requested optimizer not yet supported.");
  if (grad_average < 0 || grad_average > 2 || (grad_average != 0 && (!grad_state || n_average < 1))) return fail(h, PYVES_EINVAL, "optimizer_step: gradient averaging needs its state vector and n_average >= 1");
  if (step < 1) return fail(h, PYVES_EINVAL, "optimizer_step: step counts from 1");
  OptHyper H;
  H.lr = hyper[0];
  H.a = hyper[1];
  H.b = hyper[2];
  H.eps = hyper[3];
  H.bc1 = H.bc2_sqrt = 1.0;
  if (kind == PYVES_OPT_ADAM) {
    if (!state1 || !state2) return fail(h, PYVES_EINVAL, "optimizer_step: Adam needs exp_avg and exp_avg_sq");
    H.bc1 = 1.0 - std::pow(H.a, (double)step);
    H.bc2_sqrt = std::sqrt(1.0 - std::pow(H.b, (double)step));
  } else if (kind == PYVES_OPT_RMSPROP) {
    if (!state1) return fail(h, PYVES_EINVAL, "optimizer_step: RMSprop needs square_avg");
  } else if (H.a != 0.0 && !state1) {
    return fail(h, PYVES_EINVAL, "optimizer_step: SGD with momentum needs its buffer");
  }
  DeviceGuard g(device);
  cudaError_t e = launch_optimizer_step(kind, (long long)P, sum_w, sum_wf, f_target, grad_average, grad_decay, (long long)n_average, grad_state,
                                        theta, state1, state2, H, (long long)step, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(h, PYVES_ECUDA, cudaGetErrorString(e));
  g_launches++;
  return PYVES_OK;
}

int pyves_set_bias_pathcv(pyves_t* h, int degree, int npath, const double* x_i, const double* y_i, double lam, const double* w, double phi, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (degree < 0 || degree > PYVES_MAX_DEGREE) return fail(h, PYVES_EINVAL, "degree must be in [0, 63]");
  if (npath < 1 || npath > PYVES_MAX_PATH) return fail(h, PYVES_EINVAL, "path needs 1..256 images");
  if (!x_i || !y_i || !w || !(lam > 0)) return fail(h, PYVES_EINVAL, "need path images, weights and lam > 0");
  DeviceGuard g(h->device);
  std::vector<double> path(x_i, x_i + npath);
  path.insert(path.end(), y_i, y_i + npath);
  const int rc = upload(h, &h->pc_path_dev, path.data(), path.size(), static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  h->pc_degree = degree;
  h->pc_npath = npath;
  h->pc_lam = lam;
  h->pc_phi = phi;
  std::memset(h->pc_w, 0, sizeof(h->pc_w));
  for (int i = 0; i <= degree; ++i) h->pc_w[i] = w[i];
  h->bias_kind = PYVES_BIAS_PATHCV;
  return PYVES_OK;
}

int pyves_set_bias_harmonic(pyves_t* h, int axis, double k, double s0) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!valid_axis(axis)) return fail(h, PYVES_EINVAL, "Invalid axis");
  h->h_axis = axis;
  h->h_k = k;
  h->h_s0 = s0;
  h->bias_kind = PYVES_BIAS_HARMONIC;
  return PYVES_OK;
}

int pyves_basis_size(const pyves_t* h, int64_t* K) {
  if (!h || !K) return PYVES_EINVAL;
  *K = basis_size(h);
  return PYVES_OK;
}

// ------------------------------------------------------------------------------------------ state
int pyves_set_state(pyves_t* h, const void* pos, const void* vel, int on_host, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = (size_t)h->n * h->dims * h->esz;
  const cudaMemcpyKind kind = on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  if (pos) CU(cudaMemcpyAsync(h->pos, pos, bytes, kind, st));
  if (vel) CU(cudaMemcpyAsync(h->vel, vel, bytes, kind, st));
  if (on_host == 1) CU(cudaStreamSynchronize(st));   // on_host == 2: pinned host buffers, the caller synchronises
  return PYVES_OK;
}

int pyves_get_state(pyves_t* h, void* pos, void* vel, int on_host, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = (size_t)h->n * h->dims * h->esz;
  const cudaMemcpyKind kind = on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (pos) CU(cudaMemcpyAsync(pos, h->pos, bytes, kind, st));
  if (vel) CU(cudaMemcpyAsync(vel, h->vel, bytes, kind, st));
  if (on_host == 1) CU(cudaStreamSynchronize(st));   // on_host == 2: pinned host buffers, the caller synchronises
  return PYVES_OK;
}

int pyves_init_positions_mc(pyves_t* h, const double* box, double beta, int ntrials, int64_t* n_failed, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!h->have_pot) return fail(h, PYVES_ESTATE, "set_potential must be called first");
  if (!box || !(box[1] >= box[0]) || !(box[3] >= box[2]) || !(beta > 0) || ntrials < 1) return fail(h, PYVES_EINVAL, "bad placement box / beta / ntrials");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->prec == PYVES_F64) {
    CU(launch_init_positions<double>(static_cast<double*>(h->pos), h->n, h->dims, h->gid0, box, beta, ntrials, h->pot, make_keys(h->seed), h->fail_dev, st));
  } else {
    CU(launch_init_positions<float>(static_cast<float*>(h->pos), h->n, h->dims, h->gid0, box, beta, ntrials, h->pot, make_keys(h->seed), h->fail_dev, st));
  }
  g_launches++;
  unsigned long long nf = 0;
  CU(cudaMemcpyAsync(&nf, h->fail_dev, sizeof(nf), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (n_failed) *n_failed = (int64_t)nf;
  return PYVES_OK;
}

int pyves_init_velocities(pyves_t* h, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!h->have_int) return fail(h, PYVES_ESTATE, "set_integrator must be called first");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double sigma = std::sqrt(h->kT / h->mass);
  if (h->prec == PYVES_F64) {
    CU(launch_init_velocities<double>(static_cast<double*>(h->vel), h->n, h->dims, h->gid0, sigma, make_keys(h->seed), st));
  } else {
    CU(launch_init_velocities<float>(static_cast<float*>(h->vel), h->n, h->dims, h->gid0, (float)sigma, make_keys(h->seed), st));
  }
  g_launches++;
  return PYVES_OK;
}

// --------------------------------------------------------------------------------------- hot path
int pyves_step_record(pyves_t* h, int64_t nsteps, int64_t every, int64_t n_rec, void* traj_out, const void* injected_noise, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!h->have_int || !h->have_pot) return fail(h, PYVES_ESTATE, "set_integrator and set_potential must be called before stepping");
  if (nsteps < 0) return fail(h, PYVES_EINVAL, "nsteps must be >= 0");
  if (traj_out && (every < 1 || n_rec < 1 || n_rec > h->n)) return fail(h, PYVES_EINVAL, "recording needs every >= 1 and 1 <= n_rec <= n_walkers");
  if (nsteps == 0) return PYVES_OK;
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Philox launches are cut at multiples of 2^33 steps: PhiloxSeq (philox.cuh) keeps the upper word of the call
  // index fixed within a launch
  int64_t left = nsteps;
  while (left > 0) {
    int64_t chunk = left;
    if (!injected_noise && !traj_out) {
      const int64_t boundary = ((h->step >> 33) + 1) << 33;
      if (boundary - h->step < chunk) chunk = boundary - h->step;
    }
    const int rc = h->prec == PYVES_F64 ? do_step<double>(h, chunk, injected_noise, traj_out, every, n_rec, st)
                                        : do_step<float>(h, chunk, injected_noise, traj_out, every, n_rec, st);
    if (rc != PYVES_OK) return rc;
    h->step += chunk;
    left -= chunk;
  }
  return PYVES_OK;
}

int pyves_step(pyves_t* h, int64_t nsteps, const void* injected_noise, void* stream) {
  return pyves_step_record(h, nsteps, 1, 0, nullptr, injected_noise, stream);
}

int pyves_eval_bias(pyves_t* h, const void* pos, int64_t n, int ncols, void* E, void* F, int on_host, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (n < 0 || ncols < 1 || ncols > 3 || (!pos && n > 0)) return fail(h, PYVES_EINVAL, "need positions [n][ncols], ncols in 1..3");
  if (h->bias_kind == PYVES_BIAS_MLP_1D)
    for (int w : h->mlp_sizes)
      if (w > kGenH) return fail(h, PYVES_EUNSUPPORTED, "MLP hidden width > 128 is not supported");
  if (n == 0) return PYVES_OK;
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t pb = (size_t)n * ncols * h->esz, eb = (size_t)n * h->esz;
  const void* dpos = pos;
  void *dE = E, *dF = F;
  if (on_host) {
    const int rc = ensure_stage(h, 2 * pb + eb);
    if (rc) return rc;
    char* base = static_cast<char*>(h->stage);
    CU(cudaMemcpyAsync(base, pos, pb, cudaMemcpyHostToDevice, st));
    dpos = base;
    dF = F ? base + pb : nullptr;
    dE = E ? base + 2 * pb : nullptr;
  }
  { const int rc = refresh_l1_host(h, st); if (rc) return rc; }
  if (h->prec == PYVES_F64) {
    CU(launch_eval_bias<double>(static_cast<const double*>(dpos), n, ncols, static_cast<double*>(dE), static_cast<double*>(dF), any_params<double>(h), st));
  } else {
    CU(launch_eval_bias<float>(static_cast<const float*>(dpos), n, ncols, static_cast<float*>(dE), static_cast<float*>(dF), any_params<float>(h), st));
  }
  g_launches++;
  if (on_host) {
    if (E) CU(cudaMemcpyAsync(E, dE, eb, cudaMemcpyDeviceToHost, st));
    if (F) CU(cudaMemcpyAsync(F, dF, pb, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return PYVES_OK;
}

int pyves_energies(pyves_t* h, void* PE, void* KE, int on_host, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if (!h->have_int || !h->have_pot) return fail(h, PYVES_ESTATE, "set_integrator and set_potential must be called first");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t eb = (size_t)h->n * h->esz;
  void *dPE = PE, *dKE = KE;
  if (on_host) {
    const int rc = ensure_stage(h, 2 * eb);
    if (rc) return rc;
    dPE = PE ? h->stage : nullptr;
    dKE = KE ? static_cast<char*>(h->stage) + eb : nullptr;
  }
  { const int rc = refresh_l1_host(h, st); if (rc) return rc; }
  if (h->prec == PYVES_F64) {
    CU(launch_energies<double>(static_cast<const double*>(h->pos), static_cast<const double*>(h->vel), h->n, h->dims, static_cast<double*>(dPE), static_cast<double*>(dKE), h->mass, h->dt, pot_params<double>(h), any_params<double>(h), st));
  } else {
    CU(launch_energies<float>(static_cast<const float*>(h->pos), static_cast<const float*>(h->vel), h->n, h->dims, static_cast<float*>(dPE), static_cast<float*>(dKE), (float)h->mass, (float)h->dt, pot_params<float>(h), any_params<float>(h), st));
  }
  g_launches++;
  if (on_host) {
    if (PE) CU(cudaMemcpyAsync(PE, dPE, eb, cudaMemcpyDeviceToHost, st));
    if (KE) CU(cudaMemcpyAsync(KE, dKE, eb, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return PYVES_OK;
}

int pyves_moments(pyves_t* h, const void* pos, int64_t n, int ncols, const void* row_weights, double* sum_w, double* sum_wf, double* sum_wff, int on_host, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  const int64_t K = basis_size(h);
  if (K == 0) return fail(h, PYVES_ESTATE, "the current bias has no basis functions / parameters");
  if (!sum_w || !sum_wf) return fail(h, PYVES_EINVAL, "sum_w and sum_wf are required");
  if (pos && (n < 0 || ncols < 1 || ncols > 3)) return fail(h, PYVES_EINVAL, "need positions [n][ncols], ncols in 1..3");
  if (h->bias_kind == PYVES_BIAS_LEGENDRE_2D && pos && ncols < 2) return fail(h, PYVES_EINVAL, "the 2D basis needs at least 2 columns");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nrows = pos ? n : h->n;
  const void *dpos = pos, *dw = row_weights;
  double *dsw = sum_w, *dsf = sum_wf, *dsff = sum_wff;
  if (on_host) {
    const size_t pb = pos ? (size_t)n * ncols * h->esz : 0, wb = row_weights ? (size_t)nrows * h->esz : 0;
    const size_t pbp = (pb + 15) / 16 * 16, wbp = (wb + 15) / 16 * 16;
    const size_t ob = (size_t)(1 + K + (sum_wff ? K * K : 0)) * sizeof(double);
    const int rc = ensure_stage(h, pbp + wbp + ob);
    if (rc) return rc;
    char* base = static_cast<char*>(h->stage);
    if (pos) {
      CU(cudaMemcpyAsync(base, pos, pb, cudaMemcpyHostToDevice, st));
      dpos = base;
    }
    if (row_weights) {
      CU(cudaMemcpyAsync(base + pbp, row_weights, wb, cudaMemcpyHostToDevice, st));
      dw = base + pbp;
    }
    dsw = reinterpret_cast<double*>(base + pbp + wbp);
    dsf = dsw + 1;
    dsff = sum_wff ? dsf + K : nullptr;
  }
  const int rc = h->prec == PYVES_F64 ? moments_impl<double>(h, dpos, n, ncols, dw, dsw, dsf, dsff, st)
                                      : moments_impl<float>(h, dpos, n, ncols, dw, dsw, dsf, dsff, st);
  if (rc) return rc;
  if (on_host) {
    CU(cudaMemcpyAsync(sum_w, dsw, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(sum_wf, dsf, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
    if (sum_wff) CU(cudaMemcpyAsync(sum_wff, dsff, sizeof(double) * K * K, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return PYVES_OK;
}

int pyves_histogram(pyves_t* h, const void* pos, int64_t n, int ncols, int ndim, int axis, const double* range, const int* nbins, const void* row_weights, double* counts, int on_host, void* stream) {
  if (!h) return fail(h, PYVES_EINVAL, "handle is NULL");
  if ((ndim != 1 && ndim != 2) || !range || !nbins || !counts) return fail(h, PYVES_EINVAL, "need ndim 1|2, range, nbins, counts");
  if (ndim == 1 && !valid_axis(axis) && !(axis == 2)) return fail(h, PYVES_EINVAL, "Invalid axis");
  if (nbins[0] < 1 || (ndim == 2 && nbins[1] < 1) || !(range[1] > range[0]) || (ndim == 2 && !(range[3] > range[2]))) return fail(h, PYVES_EINVAL, "bad bins / range");
  if (pos && (n < 0 || ncols < 1 || ncols > 3 || (ndim == 2 && ncols < 2) || (ndim == 1 && axis >= ncols))) return fail(h, PYVES_EINVAL, "positions do not have the requested columns");
  if (!pos && ndim == 1 && axis >= h->dims) return fail(h, PYVES_EINVAL, "Invalid axis");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nrows = pos ? n : h->n;
  const size_t nb = (size_t)nbins[0] * (ndim == 2 ? nbins[1] : 1);
  const void *dpos = pos ? pos : h->pos, *dw = row_weights;
  double* dcounts = counts;
  if (on_host) {
    const size_t pb = pos ? (size_t)n * ncols * h->esz : 0, wb = row_weights ? (size_t)nrows * h->esz : 0;
    const size_t pbp = (pb + 15) / 16 * 16, wbp = (wb + 15) / 16 * 16;
    const int rc = ensure_stage(h, pbp + wbp + nb * sizeof(double));
    if (rc) return rc;
    char* base = static_cast<char*>(h->stage);
    if (pos) {
      CU(cudaMemcpyAsync(base, pos, pb, cudaMemcpyHostToDevice, st));
      dpos = base;
    }
    if (row_weights) {
      CU(cudaMemcpyAsync(base + pbp, row_weights, wb, cudaMemcpyHostToDevice, st));
      dw = base + pbp;
    }
    dcounts = reinterpret_cast<double*>(base + pbp + wbp);
    CU(cudaMemcpyAsync(dcounts, counts, nb * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  const int64_t srow = pos ? ncols : 1, scol = pos ? 1 : h->n;
  if (h->prec == PYVES_F64) {
    CU(launch_histogram<double>(static_cast<const double*>(dpos), nrows, srow, scol, ndim, axis, range, nbins, static_cast<const double*>(dw), dcounts, st));
  } else {
    CU(launch_histogram<float>(static_cast<const float*>(dpos), nrows, srow, scol, ndim, axis, range, nbins, static_cast<const float*>(dw), dcounts, st));
  }
  g_launches++;
  if (on_host) {
    CU(cudaMemcpyAsync(counts, dcounts, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return PYVES_OK;
}

int pyves_fp32_fma_peak(int device, double* tflops) {
  pyves_ctx* h = nullptr;
  if (!tflops) return fail(h, PYVES_EINVAL, "tflops is NULL");
  DeviceGuard g(device);
  CU(fp32_fma_peak(tflops, 0));
  g_launches += 5;
  return PYVES_OK;
}

int pyves_fp32_fma2_peak(int device, double* tflops) {
  pyves_ctx* h = nullptr;
  if (!tflops) return fail(h, PYVES_EINVAL, "tflops is NULL");
  DeviceGuard g(device);
  CU(fp32_fma_peak(tflops, 1));
  g_launches += 5;
  return PYVES_OK;
}

}  // extern "C"
