/* pyves_b200 -- C ABI of the B200-native biased-Langevin VES hot path.
 *
 * One opaque handle = one ensemble of independent walkers on one GPU.  Every entry point
 * returns 0 on success or a negative PYVES_E* code; nothing throws across the boundary;
 * pyves_last_error() returns the message of the last failure on that handle (or the last
 * handle-less failure when h == NULL).  All array arguments are CALLER-OWNED and are not
 * retained past the call, except the state that pyves_set_state() copies in and the bias
 * parameters that pyves_set_bias_*() copy to the device.  A handle is not thread-safe;
 * distinct handles are independent.  Work is enqueued on the cudaStream_t passed as
 * `stream` (NULL = legacy default stream).  "on_host" selects host pointers (the library
 * stages through device memory and synchronises the stream) or device pointers.
 *
 * Element type of `void*` state/energy/force arrays is the handle's precision
 * (float for PYVES_F32, double for PYVES_F64); moment / histogram outputs are always double.
 *
 * Each declaration cites the reference interface (apallath/pyves) it replaces.
 */
#ifndef PYVES_B200_H
#define PYVES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pyves_ctx pyves_t;

#define PYVES_ABI_VERSION 2

/* error codes */
#define PYVES_OK 0
#define PYVES_EINVAL (-1)   /* bad argument (Python shim raises ValueError) */
#define PYVES_ECUDA (-2)    /* CUDA runtime failure, sticky (RuntimeError) */
#define PYVES_ESTATE (-3)   /* call sequence error, e.g. stepping before set_integrator */
#define PYVES_ENOMEM (-4)
#define PYVES_EUNSUPPORTED (-5) /* NotImplementedError */

/* precision of state and force arithmetic */
#define PYVES_F32 0
#define PYVES_F64 1

/* integration scheme: ves/langevin_dynamics.py:49-51 builds openmm.LangevinIntegrator */
#define PYVES_SCHEME_OPENMM_LANGEVIN 0 /* v <- a v + b F/m + c xi/sqrt(m); x <- x + dt v */
#define PYVES_SCHEME_MIDDLE 1          /* openmm.LangevinMiddleIntegrator (BAOAB-like), optional */

/* potential kinds: ves/potentials.py.  params (doubles):
 *   SLIP_BOND_2D  : force_x force_y y_0 y_scale y_shift xy_0 xy_scale          (:166,183)
 *   CATCH_BOND_2D : the 7 above + gx_0 gx_scale gy_0 gy_scale                  (:199,229)
 *   SZABO_BEREZHKOVSKII : x0 omega2 Omega2 Delta                               (:241-254)
 *   others take none.  z always feels 1000 z^2 (:105); 1D kinds add 1000 y^2 (:32).     */
#define PYVES_POT_SINGLE_WELL_1D 0
#define PYVES_POT_DOUBLE_WELL_1D 1
#define PYVES_POT_SINGLE_WELL_2D 2
#define PYVES_POT_DOUBLE_WELL_2D 3
#define PYVES_POT_SLIP_BOND_2D 4
#define PYVES_POT_CATCH_BOND_2D 5
#define PYVES_POT_SZABO_BEREZHKOVSKII 6
#define PYVES_POT_MULLER_BROWN 7

/* bias kinds */
#define PYVES_BIAS_NONE 0
#define PYVES_BIAS_LEGENDRE_1D 1 /* ves/basis.py:16-156  LegendreBasis1D */
#define PYVES_BIAS_LEGENDRE_2D 2 /* tensor product, new (BASELINE.json configs[1]) */
#define PYVES_BIAS_MLP_1D 3      /* ves/basis.py:159-232 NNBasis1D */
#define PYVES_BIAS_PATHCV 4      /* ves/basis.py:386-548 LegendreBasis2DPathCV */
#define PYVES_BIAS_HARMONIC 5    /* ves/bias.py:97-133   HarmonicBias_SingleParticle_1D */

#define PYVES_ACT_RELU 0
#define PYVES_ACT_TANH 1

#define PYVES_MAX_DEGREE 63   /* Legendre degree per axis */
#define PYVES_MAX_PATH 256    /* images on a path CV */

/* ---- lifetime ------------------------------------------------------------------------ */
/* replaces SingleParticleSimulation.__init__ (ves/langevin_dynamics.py:21-83): System +
 * Context creation.  dims = 2 (z dropped; it is decoupled) or 3.  `seed` keys the Philox
 * stream (setRandomNumberSeed, :52-53). */
int pyves_create(pyves_t** out, int device, int64_t n_walkers, int dims, uint64_t seed, int precision);
int pyves_destroy(pyves_t* h);
const char* pyves_last_error(const pyves_t* h);
int pyves_abi_version(void);

/* ---- configuration ------------------------------------------------------------------- */
/* LangevinIntegrator(temp, friction, timestep) + addParticle(mass), :34-51.  kT in kJ/mol,
 * friction in 1/ps, dt in ps, mass in Da. */
int pyves_set_integrator(pyves_t* h, double mass, double kT, double friction, double dt, int scheme);
/* system.addForce(potential), :44-47 (CustomExternalForce from ves/potentials.py) */
int pyves_set_potential(pyves_t* h, int kind, const double* params, int nparams);
/* global id of local walker 0 (walker sharding across GPUs); Philox counters use global ids */
int pyves_set_walker_offset(pyves_t* h, int64_t gid0);
int pyves_set_step_counter(pyves_t* h, int64_t step);
int pyves_get_step_counter(const pyves_t* h, int64_t* step);

/* system.addForce(bias.force) + context.reinitialize, :143-145,155-159,167-169; the
 * TorchForce(model_loc) round trip of ves/bias.py:52-58 becomes a parameter upload. */
int pyves_set_bias_none(pyves_t* h);
/* axis 0 = x, 1 = y; w[degree+1] */
int pyves_set_bias_legendre1d(pyves_t* h, int axis, int degree, double lo, double hi, const double* w);
/* The three setters below keep their parameters in device buffers: the host arrays are copied with
 * cudaMemcpyAsync on `stream` (pageable source: staged before the call returns, so the caller may reuse it at
 * once), i.e. ordered after the launches already enqueued on that stream; launches of this handle on OTHER
 * streams must wait for `stream` themselves.
 * minmax = {xlo, xhi, ylo, yhi}; w row-major [deg_x+1][deg_y+1] */
int pyves_set_bias_legendre2d(pyves_t* h, int deg_x, int deg_y, const double* minmax, const double* w, void* stream);
/* theta in torch parameters() order: W0[h0,1] b0[h0] W1[h1,h0] b1[h1] ... Wout[1,hL] bout[1] */
int pyves_set_bias_mlp(pyves_t* h, int axis, double lo, double hi, int n_hidden, const int* sizes, int act,
                       const double* theta, void* stream);
int pyves_set_bias_pathcv(pyves_t* h, int degree, int npath, const double* x_i, const double* y_i, double lam,
                          const double* w, double phi, void* stream);
int pyves_set_bias_harmonic(pyves_t* h, int axis, double k, double s0);
/* number of basis functions (Legendre) or parameters (MLP) of the current bias */
int pyves_basis_size(const pyves_t* h, int64_t* K);

/* ---- walker state: SoA, pos[dims][n], vel[dims][n] ---------------------------------- */
/* context.setPositions / setState (:69-76).  on_host: 0 = device arrays; 1 = host arrays, returns when the copy
 * is done; 2 = PINNED host arrays, copy enqueued on `stream`, the caller synchronises before reusing them
 * (lets the copies of one handle overlap the kernels of another). */
int pyves_set_state(pyves_t* h, const void* pos, const void* vel, int on_host, void* stream);
/* context.getState(getPositions, getVelocities) (:201,208) */
int pyves_get_state(pyves_t* h, void* pos, void* vel, int on_host, void* stream);
/* ves/config_creation.py:56-68 for every walker: up to ntrials uniform draws in
 * box = {xmin,xmax,ymin,ymax}, accept the first with exp(-beta U) >= 0.5; z = 0.
 * n_failed (host) receives the number of walkers that were never accepted. */
int pyves_init_positions_mc(pyves_t* h, const double* box, double beta, int ntrials, int64_t* n_failed, void* stream);
/* context.setVelocitiesToTemperature (:70-74): v = sqrt(kT/m) N(0,1) */
int pyves_init_velocities(pyves_t* h, void* stream);

/* ---- the hot path --------------------------------------------------------------------- */
/* integrator.step(nsteps) (:188) for every walker: potential force + bias force + noise +
 * update, state held in registers for the whole launch.  injected_noise: NULL = Philox, or
 * a device array [nsteps][dims][n] of N(0,1) draws (parity tests). */
int pyves_step(pyves_t* h, int64_t nsteps, const void* injected_noise, void* stream);
/* One iteration between two coefficient updates with HOST buffers, every operation enqueued on `stream` (no synchronisation): the walker
 * state from PINNED host memory (pos_in / vel_in, SoA [dims][n] of the handle's element type), nsteps steps, the basis moments of the
 * handle's walkers into the device buffers sum_w[1] / sum_wf[K], the walker state back to the pinned buffers pos_out / vel_out (which may
 * be the input buffers).  The four calls of set_state(on_host = 2) -> step -> moments -> get_state(on_host = 2) in one: what a caller
 * that keeps its ensemble on the host does every iteration (the reference moves its one particle through
 * context.setPositions / step / getState, ves/langevin_dynamics.py:60-76,188,201). */
int pyves_iterate_pinned(pyves_t* h, const void* pos_in, const void* vel_in, int64_t nsteps, double* sum_w, double* sum_wf,
                         void* pos_out, void* vel_out, void* stream);
/* same, recording positions of walkers [0, n_rec) before every `every`-th step (frame t is
 * written before step t, as :180-188 does) into traj_out (device) [n_frames][dims][n_rec],
 * n_frames = ceil(nsteps / every). */
int pyves_step_record(pyves_t* h, int64_t nsteps, int64_t every, int64_t n_rec, void* traj_out,
                      const void* injected_noise, void* stream);

/* TorchForce evaluation (ves/bias.py:58; LegendreBasis1D.forward ves/basis.py:123-156):
 * bias energy E[n] and force F[n][ncols] (= -dE/dpos) at caller positions pos[n][ncols]
 * (row-major like the reference's [N,3] tensors; ncols 1..3, missing columns read as 0).
 * E or F may be NULL. */
int pyves_eval_bias(pyves_t* h, const void* pos, int64_t n, int ncols, void* E, void* F, int on_host, void* stream);
/* potential energy U + V_bias and kinetic energy per walker of the handle's state
 * (context.getState(getEnergy=True), :227-228) */
int pyves_energies(pyves_t* h, void* PE, void* KE, int on_host, void* stream);

/* The two Python loops of VESBias_SingleParticle_1D.update (ves/bias.py:219-233) and the
 * ensemble averages of textbook VES: sum_w, sum_w f_k, optionally sum_w f_k f_l, with
 * f = dV/dparam (basis functions for Legendre, parameter gradient for the MLP).
 * pos == NULL: the handle's walkers; else caller rows pos[n][ncols].  row_weights == NULL:
 * w = 1.  sum_wff may be NULL.  Outputs are double: sum_w[1], sum_wf[K], sum_wff[K*K]. */
int pyves_moments(pyves_t* h, const void* pos, int64_t n, int ncols, const void* row_weights, double* sum_w,
                  double* sum_wf, double* sum_wff, int on_host, void* stream);

/* ves/visualization.py:475 histogram of positions (pos == NULL: the handle's walkers).
 * ndim 1: axis picks the column; ndim 2: (x, y).  range = {lo0, hi0, lo1, hi1};
 * counts[nbins0 * nbins1] (row-major, x outer) is ACCUMULATED into (zero it first). */
int pyves_histogram(pyves_t* h, const void* pos, int64_t n, int ncols, int ndim, int axis, const double* range,
                    const int* nbins, const void* row_weights, double* counts, int on_host, void* stream);

/* Same as pyves_set_bias_legendre2d with the coefficients read from DEVICE memory (FP64, row-major
 * [deg_x + 1][deg_y + 1]) in stream order: the coefficient update of a device-resident optimiser reaches the
 * step kernels without a host round trip (the per-update TorchScript save / TorchForce reload of
 * ves/bias.py:249-253 + ves/langevin_dynamics.py:153-159 becomes one conversion kernel and, for the static
 * 8 x 8 / 16 x 16 kernels, a device-to-device copy into their __constant__ weights).  w_dev is read when the
 * stream reaches the call; the caller keeps it alive and unchanged until then. */
int pyves_set_bias_legendre2d_device(pyves_t* h, int deg_x, int deg_y, const double* minmax, const double* w_dev, void* stream);

/* The 1D counterpart: degree + 1 FP64 coefficients in device memory.  The packed step kernels read them from
 * their __constant__ array; consumers that take the coefficients by value (pyves_eval_bias, pyves_energies,
 * the general step kernel) refresh a host copy first (one synchronous 8 (degree + 1)-byte read). */
int pyves_set_bias_legendre1d_device(pyves_t* h, int axis, int degree, double min, double max, const double* w_dev, void* stream);

/* h takes the SAME device-resident Legendre coefficients as `leader` (same device and precision; the leader's
 * bias was set by pyves_set_bias_legendre{1d,2d}_device): two device-to-device copies of the leader's converted
 * weights in stream order, and the same content tag, so that handles which evaluate one bias on shards of one
 * ensemble (each on its own stream) share the __constant__ weights instead of re-copying and serialising on
 * them.  Independent handles need no such call: launches that read the __constant__ array and copies into it are
 * ordered across streams with events.  Replaces, like the two calls above, the TorchForce reload of
 * ves/langevin_dynamics.py:153-159 for every replica that shares one bias. */
int pyves_set_bias_like(pyves_t* h, const pyves_t* leader, void* stream);

/* One averaged-SGD update of the coefficients entirely on `device` (all pointers device, FP64): gradient
 * f_target - sum_wf / sum_w, Hessian term beta Cov (alpha - abar) when sum_wff != NULL, optional diagonal
 * preconditioner `scale`, then the EWMA (averaging 0) or running (averaging 1, n_after = updates done incl. this
 * one) average.  The optimiser step that ves/bias.py:236-245 delegates to torch.optim on the host after a
 * Python loop over the trajectory; here it is one kernel between pyves_moments and
 * pyves_set_bias_legendre*_device.  Outputs must not alias the inputs. */
int pyves_averaged_sgd_step(int device, int K, const double* sum_w, const double* sum_wf, const double* sum_wff,
                            const double* f_target, const double* alpha, const double* abar, const double* scale,
                            double mu, double beta, int averaging, double decay, int64_t n_after, double* alpha_out,
                            double* abar_out, void* stream);

/* pyves_set_bias_mlp with the parameters read from DEVICE memory (FP64, torch parameters() order) in stream order:
 * conversion to the element type, the fold of the activation-free pair Linear(1, h0) -> Linear(h0, h1)
 * (ves/basis.py:197-202) and the transposed repack of the third layer run as one kernel, so a device-resident
 * optimiser hands new NN parameters to the step kernels without the TorchScript save / TorchForce reload of
 * ves/bias.py:249-253 + ves/langevin_dynamics.py:153-159 and without any host copy.  n_hidden >= 2. */
int pyves_set_bias_mlp_device(pyves_t* h, int axis, double lo, double hi, int n_hidden, const int* sizes, int act,
                              const double* theta_dev, void* stream);

/* optimiser kinds of pyves_optimizer_step: the three the reference accepts (ves/bias.py:193-200) */
#define PYVES_OPT_SGD 0
#define PYVES_OPT_RMSPROP 1
#define PYVES_OPT_ADAM 2

/* One torch.optim step of P device-resident parameters `theta` (FP64, in place) on the gradient of the VES
 * functional g = f_target - sum_wf / sum_w, i.e. optimizer.step() of ves/bias.py:236-245 without leaving the GPU.
 * grad_average: 0 none, 1 running mean over updates, 2 exponentially weighted mean with grad_decay; grad_state [P]
 * holds the mean and n_average counts the updates averaged so far including this one.  hyper (host, 4 doubles):
 * {lr, momentum, -, -} for SGD; {lr, alpha, -, eps} for RMSprop (state1 = square_avg); {lr, beta1, beta2, eps} for
 * Adam (state1 = exp_avg, state2 = exp_avg_sq, bias corrections from `step`, which counts from 1).  Weight decay,
 * Nesterov, centred RMSprop and AMSGrad are not implemented.  sum_w <= 0 leaves everything unchanged. */
int pyves_optimizer_step(int device, int kind, int64_t P, const double* sum_w, const double* sum_wf, const double* f_target,
                         int grad_average, double grad_decay, int64_t n_average, double* grad_state, double* theta,
                         double* state1, double* state2, const double* hyper, int64_t step, void* stream);

/* Domain of the ensemble averages of pyves_moments (Legendre / path-CV bases; for the NN basis the interval is
 * [min, max] of NNBasis1D, scaled coordinate in [0, 1]).  0 (default): every
 * row counts, rows outside the basis interval enter with their clamped coordinate, exactly what
 * LegendreBasis1D.forward does to them (ves/basis.py:143).  1: rows outside the interval get weight 0, i.e.
 * <f>_V conditional on the box -- the estimator of the VES functional restricted to the support of the
 * target distribution (ves/target.py:24-31 defines p on [-1, 1] only; README.md:10-31). */
int pyves_set_moment_domain(pyves_t* h, int inside_only);

/* ---- multi-GPU: the one exchange step of the path ----------------------------------------------------- */
/* Walkers are independent between coefficient updates, so rank r owns a block of global walker ids
 * (pyves_set_walker_offset) and the only collective is the sum of the moment vector pyves_moments produced,
 * [sum_w, sum_w f (K), sum_w f f (K * K)?], once per update; every rank then applies the identical optimiser step
 * (no broadcast).  The reference has no multi-walker code (ves/langevin_dynamics.py:43-47 adds one particle); this
 * is the all-reduce its README's "ensemble averages" (README.md:10-31) would need.  NCCL is loaded at run time
 * (libnccl.so.2 of the process, e.g. torch's), the communicator is bootstrapped from any out-of-band channel: rank 0
 * calls pyves_comm_unique_id and ships the 128 bytes (torch.distributed broadcast, MPI, a file), every rank calls
 * pyves_comm_create.  pyves_allreduce_moments sums `count` doubles in place on `stream`. */
typedef struct pyves_comm pyves_comm_t;
int pyves_comm_unique_id(void* id128);
int pyves_comm_create(pyves_comm_t** out, int device, int nranks, int rank, const void* id128);
int pyves_comm_destroy(pyves_comm_t* c);
int pyves_allreduce_moments(pyves_comm_t* c, double* moments_dev, int64_t count, void* stream);
int pyves_comm_nccl_version(int* version);
const char* pyves_comm_last_error(void);

/* kernel-variant selector for tuning runs (-1 = default table; 99 = force the general step kernel; 7 / 6 = the 2D Legendre
 * contraction of the step kernel on the tensor cores for every size <= 64 x 64 / on the CUDA cores for every size (default:
 * tensor cores for every FP32 2D Legendre basis up to 64 x 64 except the static 8 x 8 double-well kernel); for ReLU NNBasis1D biases
 * with two / three hidden layers -1 = the piecewise-constant dV/ds table on FP32 handles (and, for ensembles much larger than the table,
 * parameter-gradient moments through its pieces), 8 = the table on any handle and for any number of rows, 7 / 6 = the network itself
 * evaluated every step on the tensor cores / CUDA cores; 98 / 97 = process-wide: covariance on CUDA cores only / tensor cores allowed
 * again; 96 / 95 = process-wide: first moments of 2D bases beyond 16 x 16 with the one-function-per-thread kernel / the register-tiled
 * kernel again; 1000 + k = process-wide: rows the tensor-core covariance accumulates in TMEM between two drains = 2^k) */
int pyves_set_tuning(pyves_t* h, int variant);

/* Walkers that ONE full wave of the handle's step kernel advances on its device (SMs x resident CTAs x walkers per CTA), or 0
 * when the kernel has no such granularity worth respecting (light kernels with many small CTAs).  The tensor-core step kernels
 * keep one large CTA per SM, so a launch costs whole waves: callers that split an ensemble into chunks (pipelined host <-> device
 * copies; the reference has no counterpart, its one particle lives in an openmm.Context, ves/langevin_dynamics.py:60-76) should
 * cut at multiples of this number. */
int pyves_step_wave_walkers(pyves_t* h, int64_t* walkers);

/* FP32 FMA throughput of `device` (TFLOP/s), the roofline denominator of the step kernel */
int pyves_fp32_fma_peak(int device, double* tflops);
/* the same probe issued as packed FP32x2 FMAs (FFMA2, sm_100) */
int pyves_fp32_fma2_peak(int device, double* tflops);
/* kernels launched by this library since load (bench.py's gpu_launches evidence) */
int64_t pyves_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
