// Scalar/vector arithmetic shared by the CUDA kernels and the host-side ray generator.
//
// Everything here is written so that nvcc (--fmad=false, IEEE div/sqrt, no FTZ) and
// g++ (-ffp-contract=off, SSE2 scalar) produce IDENTICAL bits: only + - * / sqrt, integer
// ops and exact conversions are used, and every expression keeps the operation order of
// the reference function it reproduces (cited per function). No libm transcendental is
// called on any path whose result is parity-checked bit-exactly.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rt {

struct V3 {
  float x, y, z;
};

RT_HD V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
RT_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
RT_HD V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
RT_HD V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
// dot/cross/norm keep the left-to-right sums of geometry.hpp:133-145.
RT_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT_HD float norm2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
RT_HD V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// Vec3fa::normalized(): v * (1/sqrt(|v|^2))  (geometry.hpp:139)
RT_HD V3 normalized(V3 a) { return a * (1.f / sqrtf(norm2(a))); }
RT_HD V3 vabs(V3 a) { return mk(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// Vec3fa::maxi() (geometry.hpp:115-121). REFERENCE QUIRK, kept on purpose: despite its name and
// comment the reference's comparison chain selects the index of the signed MINIMUM component
// ((1,2,3) -> 0, (3,2,1) -> 2; probed by compiling the reference header). Every caller
// (RayShearParams, offsetRay, the triangle error bound, evaluateRoulette) inherits that
// behaviour, so bit-exact parity requires the same chain here.
RT_HD int maxi(V3 a) {
  if (a.x < a.y) return a.x < a.z ? 0 : 2;
  return a.y < a.z ? 1 : 2;
}
RT_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
RT_HD float maxc(V3 a) { return comp(a, maxi(a)); }

// ---------------------------------------------------------------------------------------------
// Error-bound constants (include/precision_utils.hpp:19-26), evaluated in fp32 like the reference.
constexpr float kMachineEps = 5.96046448e-08f;  // 2^-24
constexpr float rt_gamma(int i) { return (kMachineEps * i) / (1 - kMachineEps * i); }
constexpr float kRayEpsilon = kMachineEps * 1500.f;
constexpr float kGamma2 = rt_gamma(2);
constexpr float kGamma3 = rt_gamma(3);
constexpr float kGamma5 = rt_gamma(5);
constexpr float kSlabGuard = 1 + 2 * kGamma3;  // 1.00000036f (CompactBVH2Node.hpp:41)

// ---------------------------------------------------------------------------------------------
// Table-driven float-only sincos: degrees conversion, mod 360, nearest-degree lookup and the
// ACC5/ABSERR residual, as ext/math/sincos.cpp:236-355 (flg = 0 path). The table holds
// sin(i deg), i = 0..91, rounded to fp32 (identical to the reference's literals as floats).
#if defined(__CUDACC__)
static __device__ __constant__ float d_sin_deg[92] = {
#include "sin_deg_table.inc"
};
#endif
static const float h_sin_deg[92] = {
#include "sin_deg_table.inc"
};

RT_HD float sin_deg_lookup(int i) {
#if defined(__CUDA_ARCH__)
  return d_sin_deg[i];
#else
  return h_sin_deg[i];
#endif
}

RT_HD void sincos_tbl(float x, float& s, float& c) {
  x = x * 57.2957795130823208768f;  // float(180/pi)
  bool neg = x < 0.f;
  if (neg) x = -x;
  x = x - 360.f * floorf(x / 360.f);
  int ix = (int)(x + .5f);
  const float z = x - (float)ix;  // residual in [-0.5, 0.5] degrees
  bool sneg = false, cneg = false;
  if (ix > 180) { sneg = true; cneg = true; ix -= 180; }
  if (ix > 90) { cneg = !cneg; ix = 180 - ix; }
  float sx = sin_deg_lookup(ix);
  float cx = sin_deg_lookup(90 - ix);
  if (sneg) sx = -sx;
  if (cneg) cx = -cx;
  const float sz = 1.74531263774940077459e-2f * z;
  const float cz = 1.f - 1.52307909153324666207e-4f * z * z;
  float y = sx * cz + cx * sz;
  if (neg) y = -y;
  s = y;
  c = cx * cz - sx * sz;
}

// ---------------------------------------------------------------------------------------------
// xoroshiro128** + splitmix64 (include/xoshiro.hpp:18-80).
RT_HD uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
RT_HD uint64_t splitmix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
struct Rng {
  uint64_t s0, s1;
};
RT_HD void rng_seed(Rng& r, uint64_t seed) {  // xoshiro::seed (xoshiro.hpp:29-32)
  r.s0 = splitmix64(seed);
  r.s1 = splitmix64(r.s0);
}
RT_HD uint64_t rng_next(Rng& r) {  // next128ss (xoshiro.hpp:34-44)
  const uint64_t s0 = r.s0;
  uint64_t s1 = r.s1;
  const uint64_t result = rotl64(s0 * 5, 7) * 9;
  s1 ^= s0;
  r.s0 = rotl64(s0, 24) ^ s1 ^ (s1 << 16);
  r.s1 = rotl64(s1, 37);
  return result;
}
// uniform_0_1 (xoshiro.hpp:67-80): float( double(1.x) - 1.0 ). (x>>12) * 2^-52 is exact in
// double, so a single RNE conversion of the 52-bit integer followed by an exact power-of-two
// scale gives the same float without touching FP64. Can return exactly 1.0f, like the reference.
RT_HD float u01_from_bits(uint64_t bits) {
#if defined(__CUDA_ARCH__)
  return __ull2float_rn(bits >> 12) * 2.220446049250313e-16f;
#else
  return (float)(bits >> 12) * 2.220446049250313e-16f;
#endif
}
RT_HD float rng_uniform(Rng& r) { return u01_from_bits(rng_next(r)); }

// Per-(pixel, sample) stream. BUILD-DEFINED (the reference CPU path uses one global generator
// under `omp critical`, trace.cpp:144-148, and the IPU a hardware RNG; neither is reproducible):
//   key    = splitmix64(rngSeed)                       (host, once)
//   stream = splitmix64( ((pixelIndex << 32) | sample) ^ key )
//   state  = xoshiro::seed(stream)
// pixelIndex = row * imageWidth + col in FULL-image coordinates, so results do not depend on
// crop windows, batch sizes, GPU count or scheduling. The oracle implements the same scheme.
RT_HD void rng_seed_stream(Rng& r, uint64_t key, uint32_t pixelIndex, uint32_t sample) {
  const uint64_t id = ((uint64_t)pixelIndex << 32) | (uint64_t)sample;
  rng_seed(r, splitmix64(id ^ key));
}

// Natural log from basic operations only (Cephes-style reduction to [sqrt(1/2), sqrt(2)) and a
// degree-9 minimax polynomial). Deterministic across host/device; used only for the
// anti-aliasing Gaussian. x must be a positive normal float.
RT_HD float det_logf(float x) {
  union { float f; uint32_t u; } b;
  b.f = x;
  int e = (int)((b.u >> 23) & 0xffu) - 126;           // x = m * 2^e, m in [0.5, 1)
  b.u = (b.u & 0x007fffffu) | 0x3f000000u;
  float m = b.f;
  if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.f; } else { m = m - 1.f; }
  const float z = m * m;
  float p = 7.0376836292e-2f;
  p = p * m - 1.1514610310e-1f;
  p = p * m + 1.1676998740e-1f;
  p = p * m - 1.2420140846e-1f;
  p = p * m + 1.4249322787e-1f;
  p = p * m - 1.6668057665e-1f;
  p = p * m + 2.0000714765e-1f;
  p = p * m - 2.4999993993e-1f;
  p = p * m + 3.3333331174e-1f;
  float y = p * m * z;
  const float fe = (float)e;
  y = y + -2.12194440e-4f * fe;
  y = y + -0.5f * z;
  float r = m + y;
  r = r + 0.693359375f * fe;
  return r;
}

// Two independent N(0,1) variates from two raw 64-bit draws (Box-Muller). BUILD-DEFINED stand-in
// for std::normal_distribution (src/app_utils.cpp:30-40) / __builtin_ipu_f32v2grand
// (codelets/TraceCodelets.cpp:158): u1 in (0,1], u2 in [0,1) are exact 24-bit fractions.
RT_HD void gaussian_pair(uint64_t a, uint64_t b, float& g0, float& g1) {
  const float u1 = (float)((uint32_t)(a >> 40) + 1u) * 5.9604644775390625e-08f;
  const float u2 = (float)((uint32_t)(b >> 40)) * 5.9604644775390625e-08f;
  const float rad = sqrtf(-2.f * det_logf(u1));
  float s, c;
  sincos_tbl(6.28318530717958647692f * u2, s, c);
  g0 = rad * c;
  g1 = rad * s;
}

// pixelToRayDir (include/Render.hpp:74-85). x = column, y = row.
RT_HD V3 pixel_to_ray_dir(float x, float y, float w, float h, float tanTheta) {
  const float aspect = w / h;
  x = (x / w) - .5f;
  y = (y / h) - .5f;
  return normalized(mk(2.f * x * aspect * tanTheta, -2.f * y * tanTheta, -1.f));
}

}  // namespace rt
