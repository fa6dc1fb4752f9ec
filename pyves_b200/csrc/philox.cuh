// Philox4x32-10 counter-based noise (Salmon et al., SC'11) and the uniform -> normal mapping.
// Replaces the OpenMM-internal generator behind openmm.LangevinIntegrator
// (ves/langevin_dynamics.py:49-53).  Stream specification: DESIGN.md "Noise";
// CPU twin: oracle/philox.py.
#pragma once
#include "common.cuh"

namespace pv {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

enum : uint32_t { kStreamDyn2 = 0, kStreamDyn3 = 1, kStreamVel = 2, kStreamPos = 3 };

// The ten round keys depend only on the seed: computed on the host, passed by value so the
// key words are constant-bank operands of the LOP3s.
struct RngKeys {
  uint32_t k0[10];
  uint32_t k1[10];
};

inline RngKeys make_keys(uint64_t seed) {
  RngKeys K;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    K.k0[r] = a;
    K.k1[r] = b;
    a += kPhiloxW0;
    b += kPhiloxW1;
  }
  return K;
}

PV_DEV uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const RngKeys& K) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
    const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  return make_uint4(c0, c1, c2, c3);
}

PV_DEV uint4 philox_call(uint64_t gid, uint64_t call, uint32_t stream, const RngKeys& K) {
  return philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)call,
                       ((uint32_t)(call >> 32) & 0xFFFFFFu) | (stream << 24), K);
}

// Consecutive calls of one walker (the dynamics stream: call index = step / 2).  Same output as philox_call,
// with the work that does not change from call to call taken out of the loop: in round 1, M0 * gid_lo is
// constant and M1 * call_lo grows by M1 per call (a 64-bit add); its other output word, and so M1 * c2 of
// round 2, are constant as well.  17 wide multiplies per call instead of 20 (they hold the FMA pipe for
// 4 cycles each and bound the light kernels).  Valid while the upper 32 bits of the call index do not
// change: the host cuts launches at multiples of 2^33 steps (api.cu).
struct PhiloxSeq {
  uint32_t c1;        // gid_hi
  uint32_t lo0;       // low word of M0 * gid_lo  (c3 of round 2)
  uint32_t p1b_lo, p1b_hi;   // M1 * c2 of round 2
  uint64_t p1;        // M1 * call_lo of the next call
  PV_DEV void init(uint64_t gid, uint64_t call, uint32_t stream, const RngKeys& K) {
    const uint64_t p0 = (uint64_t)kPhiloxM0 * (uint32_t)gid;
    const uint32_t c3 = ((uint32_t)(call >> 32) & 0xFFFFFFu) | (stream << 24);
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[0];
    const uint64_t p1b = (uint64_t)kPhiloxM1 * n2;
    c1 = (uint32_t)(gid >> 32);
    lo0 = (uint32_t)p0;
    p1b_lo = (uint32_t)p1b;
    p1b_hi = (uint32_t)(p1b >> 32);
    p1 = (uint64_t)kPhiloxM1 * (uint32_t)call;
  }
  PV_DEV uint4 next(const RngKeys& K) {
    // round 1
    uint32_t c0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[0];
    uint32_t d1 = (uint32_t)p1;
    p1 += kPhiloxM1;
    // round 2: only M0 * c0 is new
    const uint64_t q0 = (uint64_t)kPhiloxM0 * c0;
    uint32_t n0 = p1b_hi ^ d1 ^ K.k0[1];
    uint32_t c2 = (uint32_t)(q0 >> 32) ^ lo0 ^ K.k1[1];
    d1 = p1b_lo;
    uint32_t c3 = (uint32_t)q0;
    c0 = n0;
#pragma unroll
    for (int r = 2; r < 10; ++r) {
      const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
      const uint64_t pp = (uint64_t)kPhiloxM1 * c2;
      n0 = (uint32_t)(pp >> 32) ^ d1 ^ K.k0[r];
      const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r];
      d1 = (uint32_t)pp;
      c3 = (uint32_t)p0;
      c0 = n0;
      c2 = n2;
    }
    return make_uint4(c0, d1, c2, c3);
  }
};

// (r + 0.5) 2^-32 in (0, 1]
PV_DEV float uniform01(uint32_t r, float) { return fmaf(__uint2float_rn(r), 2.3283064365386963e-10f, 1.1641532182693481e-10f); }
PV_DEV double uniform01(uint32_t r, double) { return ((double)r + 0.5) * 2.3283064365386963e-10; }

// Box-Muller: rad = sqrt(-2 ln u1), theta = pi (2 u2 - 1).  FP32 uses the MUFU approximations
// (lg2 / sqrt / sin / cos): 4 MUFU per pair of normals.
PV_DEV void box_muller(uint32_t r0, uint32_t r1, float& z0, float& z1) {
  const float u1 = uniform01(r0, 0.f);
  float l2, rad, s, c;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  const float m2ln = -1.3862943611198906f * l2;   // -2 ln(u1) = -2 ln2 lg2(u1) >= 0
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(m2ln));
  // theta = (r1 + 0.5) 2^-31 pi - pi
  const float theta = fmaf(__uint2float_rn(r1), 1.4629180792671596e-09f, -3.1415926528583340f);
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(theta));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(theta));
  z0 = rad * c;
  z1 = rad * s;
}

PV_DEV void box_muller(uint32_t r0, uint32_t r1, double& z0, double& z1) {
  const double u1 = uniform01(r0, 0.0);
  const double u2 = uniform01(r1, 0.0);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2 - 1.0, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

}  // namespace pv
