/* Oracle (test infrastructure, NOT product code): plain-C restatement of the biased-Langevin
 * step for the CPU baseline of bench.py and for larger parity cases.  FP64, OpenMP over walkers.
 *
 * Follows the same reference call sites as oracle/langevin.py (ves/langevin_dynamics.py:49-51,
 * :188 -> the OpenMM Langevin update), oracle/legendre.py
 * (ves/basis.py:86-156), oracle/mlp.py (ves/basis.py:184-232), oracle/potentials.py
 * (ves/potentials.py:151,183,253-254) and oracle/philox.py.  It is checked against those numpy
 * modules in tests/test_oracle_c.py.  PARITY UNPINNED for the integrator (OpenMM absent).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define POT_DOUBLE_WELL_2D 3
#define POT_SLIP_BOND_2D 4
#define POT_SZABO_BEREZHKOVSKII 6
#define BIAS_NONE 0
#define BIAS_LEGENDRE_1D 1
#define BIAS_LEGENDRE_2D 2
#define BIAS_MLP_1D 3
#define MAXK 64
#define MAXH 128

typedef struct {
  int pot_kind;
  double pot[11];
  int bias_kind;
  int axis, deg_x, deg_y;        /* Legendre: 1D uses deg_x; 2D both */
  double lo_x, hi_x, lo_y, hi_y; /* scaling ranges (MLP: lo_x, hi_x) */
  const double* w;               /* Legendre weights [deg_x+1] or [deg_x+1][deg_y+1] */
  int n_hidden, act;             /* MLP */
  const int* sizes;
  const double* theta;
  double a, b_over_m, c_over_sqrtm, dt;
  uint64_t seed, gid0;
} oracle_cfg;

static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1;
    c[3] = (uint32_t)p0;
    c[0] = n0;
    c[2] = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

static void box_muller(uint32_t r0, uint32_t r1, double* z0, double* z1) {
  const double u1 = ((double)r0 + 0.5) * 2.3283064365386963e-10;
  const double u2 = ((double)r1 + 0.5) * 2.3283064365386963e-10;
  const double rad = sqrt(-2.0 * log(u1)), th = M_PI * (2.0 * u2 - 1.0);
  *z0 = rad * cos(th);
  *z1 = rad * sin(th);
}

static void pot_grad(const oracle_cfg* C, double x, double y, double* gx, double* gy) {
  const double* p = C->pot;
  if (C->pot_kind == POT_DOUBLE_WELL_2D) {
    *gx = 10.0 * (4.0 * x * x * x - 8.0 * x + 0.7);
    *gy = 10.0 * y;
  } else if (C->pot_kind == POT_SLIP_BOND_2D) {
    const double t = (y - p[2]) * (y - p[2]) / p[3] - p[4], u = x - y - p[5];
    *gx = 2.0 * u / p[6] - p[0];
    *gy = 4.0 * t * (y - p[2]) / p[3] - 2.0 * u / p[6] - p[1];
  } else { /* Szabo-Berezhkovskii: x0 omega2 Omega2 Delta */
    const double x0 = p[0], w2 = p[1], W2 = p[2];
    double d;
    if (x < -0.5 * x0) d = w2 * (x + x0);
    else if (x >= 0.5 * x0) d = w2 * (x - x0);
    else d = -w2 * x;
    *gx = d + W2 * (x - y);
    *gy = -W2 * (x - y);
  }
}

static double clamp_scale(double q, double lo, double hi, int* inside) {
  double s = (q - (lo + hi) / 2) / ((hi - lo) / 2);
  *inside = (s >= -1.0) && (s <= 1.0);
  return s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s);
}

static void legendre(double s, int deg, double* P, double* dP) {
  P[0] = 1.0;
  dP[0] = 0.0;
  if (deg >= 1) {
    P[1] = s;
    dP[1] = 1.0;
  }
  for (int n = 1; n < deg; ++n) {
    P[n + 1] = ((2 * n + 1) * s * P[n] - n * P[n - 1]) / (n + 1);
    dP[n + 1] = dP[n - 1] + (2 * n + 1) * P[n];
  }
}

static void bias_grad(const oracle_cfg* C, double x, double y, double* gx, double* gy) {
  *gx = 0.0;
  *gy = 0.0;
  if (C->bias_kind == BIAS_LEGENDRE_1D) {
    double P[MAXK], dP[MAXK];
    int in;
    const double s = clamp_scale(C->axis ? y : x, C->lo_x, C->hi_x, &in);
    legendre(s, C->deg_x, P, dP);
    double g = 0.0;
    for (int n = 0; n <= C->deg_x; ++n) g += C->w[n] * dP[n];
    g = in ? g / ((C->hi_x - C->lo_x) / 2) : 0.0;
    if (C->axis) *gy = g;
    else *gx = g;
  } else if (C->bias_kind == BIAS_LEGENDRE_2D) {
    double Px[MAXK], dPx[MAXK], Py[MAXK], dPy[MAXK];
    int inx, iny;
    const double sx = clamp_scale(x, C->lo_x, C->hi_x, &inx), sy = clamp_scale(y, C->lo_y, C->hi_y, &iny);
    legendre(sx, C->deg_x, Px, dPx);
    legendre(sy, C->deg_y, Py, dPy);
    const int ky = C->deg_y + 1;
    double ax = 0.0, ay = 0.0;
    for (int i = 0; i <= C->deg_x; ++i) {
      const double* wr = C->w + (size_t)i * ky;
      double A = 0.0, B = 0.0;
      for (int j = 0; j < ky; ++j) {
        A += wr[j] * Py[j];
        B += wr[j] * dPy[j];
      }
      ax += dPx[i] * A;
      ay += Px[i] * B;
    }
    *gx = inx ? ax / ((C->hi_x - C->lo_x) / 2) : 0.0;
    *gy = iny ? ay / ((C->hi_y - C->lo_y) / 2) : 0.0;
  } else if (C->bias_kind == BIAS_MLP_1D) {
    double h[MAXH], dh[MAXH], hn[MAXH], dhn[MAXH];
    const double range = C->hi_x - C->lo_x;
    h[0] = ((C->axis ? y : x) - C->lo_x) / range;
    dh[0] = 1.0;
    const double* th = C->theta;
    int nin = 1;
    const int L = C->n_hidden + 1;
    for (int l = 0; l < L; ++l) {
      const int nout = l < C->n_hidden ? C->sizes[l] : 1;
      const double* W = th;
      const double* b = th + (size_t)nin * nout;
      th = b + nout;
      const int activate = (l >= 1) && (l < L - 1);
      for (int j = 0; j < nout; ++j) {
        double z = b[j], dz = 0.0;
        for (int i = 0; i < nin; ++i) {
          z += W[(size_t)j * nin + i] * h[i];
          dz += W[(size_t)j * nin + i] * dh[i];
        }
        if (activate) {
          if (C->act == 0) {
            dz = z > 0.0 ? dz : 0.0;
            z = z > 0.0 ? z : 0.0;
          } else {
            const double t = tanh(z);
            dz *= 1.0 - t * t;
            z = t;
          }
        }
        hn[j] = z;
        dhn[j] = dz;
      }
      memcpy(h, hn, sizeof(double) * nout);
      memcpy(dh, dhn, sizeof(double) * nout);
      nin = nout;
    }
    const double g = dh[0] / range;
    if (C->axis) *gy = g;
    else *gx = g;
  }
}

/* nsteps Langevin steps (2 simulated dimensions, OpenMM scheme, Philox noise) for n walkers;
 * pos / vel are SoA [2][n].  Returns 0. */
int oracle_langevin2d(const oracle_cfg* C, double* pos, double* vel, int64_t n, int64_t step0, int64_t nsteps) {
  const uint32_t k0 = (uint32_t)C->seed, k1 = (uint32_t)(C->seed >> 32);
#pragma omp parallel for schedule(static)
  for (int64_t w = 0; w < n; ++w) {
    double x = pos[w], y = pos[n + w], vx = vel[w], vy = vel[n + w];
    const uint64_t gid = C->gid0 + (uint64_t)w;
    for (int64_t t = step0; t < step0 + nsteps; ++t) {
      const uint64_t call = (uint64_t)(t >> 1);
      uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)call, (uint32_t)(call >> 32) & 0xFFFFFFu};
      philox(c, k0, k1);
      double zx, zy;
      if (t & 1) box_muller(c[2], c[3], &zx, &zy);
      else box_muller(c[0], c[1], &zx, &zy);
      double gx, gy, bx, by;
      pot_grad(C, x, y, &gx, &gy);
      bias_grad(C, x, y, &bx, &by);
      vx = C->a * vx - C->b_over_m * (gx + bx) + C->c_over_sqrtm * zx;
      vy = C->a * vy - C->b_over_m * (gy + by) + C->c_over_sqrtm * zy;
      x += C->dt * vx;
      y += C->dt * vy;
    }
    pos[w] = x;
    pos[n + w] = y;
    vel[w] = vx;
    vel[n + w] = vy;
  }
  return 0;
}

/* first moments sum f_k of the 1D / 2D Legendre basis over the walkers (unit weights) */
int oracle_moments(const oracle_cfg* C, const double* pos, int64_t n, double* sum_f) {
  const int kx = C->deg_x + 1, ky = C->bias_kind == BIAS_LEGENDRE_2D ? C->deg_y + 1 : 1;
  memset(sum_f, 0, sizeof(double) * kx * ky);
  for (int64_t w = 0; w < n; ++w) {
    double Px[MAXK], dP[MAXK], Py[MAXK];
    int in;
    if (C->bias_kind == BIAS_LEGENDRE_2D) {
      legendre(clamp_scale(pos[w], C->lo_x, C->hi_x, &in), C->deg_x, Px, dP);
      legendre(clamp_scale(pos[n + w], C->lo_y, C->hi_y, &in), C->deg_y, Py, dP);
      for (int i = 0; i < kx; ++i)
        for (int j = 0; j < ky; ++j) sum_f[i * ky + j] += Px[i] * Py[j];
    } else {
      legendre(clamp_scale(pos[(C->axis ? n : 0) + w], C->lo_x, C->hi_x, &in), C->deg_x, Px, dP);
      for (int i = 0; i < kx; ++i) sum_f[i] += Px[i];
    }
  }
  return 0;
}

/* sum_w and first moments sum_w f_k with OpenMP (thread-private partials combined in thread order); inside_only: walkers
 * whose scaled coordinate leaves the basis interval get weight 0 (the estimator pyves_set_moment_domain(h, 1) selects) */
int oracle_moments_par(const oracle_cfg* C, const double* pos, int64_t n, int inside_only, double* sum_w, double* sum_f) {
  const int kx = C->deg_x + 1, ky = C->bias_kind == BIAS_LEGENDRE_2D ? C->deg_y + 1 : 1;
  const int K = kx * ky;
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  double* part = (double*)calloc((size_t)nt * (K + 1), sizeof(double));
  if (!part) return -1;
#pragma omp parallel num_threads(nt)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double* acc = part + (size_t)tid * (K + 1);
#pragma omp for schedule(static)
    for (int64_t w = 0; w < n; ++w) {
      double Px[MAXK], dP[MAXK], Py[MAXK];
      int inx = 1, iny = 1;
      if (C->bias_kind == BIAS_LEGENDRE_2D) {
        legendre(clamp_scale(pos[w], C->lo_x, C->hi_x, &inx), C->deg_x, Px, dP);
        legendre(clamp_scale(pos[n + w], C->lo_y, C->hi_y, &iny), C->deg_y, Py, dP);
        if (inside_only && !(inx && iny)) continue;
        for (int i = 0; i < kx; ++i)
          for (int j = 0; j < ky; ++j) acc[i * ky + j] += Px[i] * Py[j];
      } else {
        legendre(clamp_scale(pos[(C->axis ? n : 0) + w], C->lo_x, C->hi_x, &inx), C->deg_x, Px, dP);
        if (inside_only && !inx) continue;
        for (int i = 0; i < kx; ++i) acc[i] += Px[i];
      }
      acc[K] += 1.0;
    }
  }
  memset(sum_f, 0, sizeof(double) * K);
  *sum_w = 0.0;
  for (int t = 0; t < nt; ++t) {
    for (int k = 0; k < K; ++k) sum_f[k] += part[(size_t)t * (K + 1) + k];
    *sum_w += part[(size_t)t * (K + 1) + K];
  }
  free(part);
  return 0;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use all host threads */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
  {
#pragma omp single
    n = omp_get_num_threads();
  }
#endif
  return n;
}
