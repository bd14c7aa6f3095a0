// device_math.cuh — the arithmetic contract of the transport kernel.
//
// Only IEEE single-precision +,-,*,/, sqrt and the fused multiply-add where __fmaf_rn is written
// (the file is compiled with -fmad=false, so nvcc never contracts a*b+c on its own: an FMA is
// part of the specification, the oracle writes fmaf() at the same places); logarithm, exponential and
// sin/cos are the polynomials below, not the CUDA math library, so that a scalar CPU
// implementation of the same formulas (oracle/transport_oracle.c) reproduces every history
// bit for bit.
#pragma once
#include <cstdint>

#ifndef GB2_PHILOX2_UNROLL
#define GB2_PHILOX2_UNROLL 2   // measured on B200 (C3 / C2 / C1, 1e9 primaries/s): 1: 1.483 / 1.499 / 1.088, 2: 1.506 / 1.538 / 1.115, 5: 1.486 / 1.509 / 1.100, 10: 1.470 / 1.488 / 1.073
#endif
namespace gb2 {
constexpr int kPhilox2Unroll = GB2_PHILOX2_UNROLL;   // rounds of the two-block Philox per loop trip

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }

// natural logarithm: x = m 2^k, m in [sqrt(1/2), sqrt(2)); log(m) = f - f^2/2 + f^3 P(f), f = m - 1,
// P of degree 8 (fitted by tools/fit_log_poly.py, < 0.8 ulp), Horner with fused multiply-adds:
// no division, 22 instructions.
__device__ __forceinline__ float d_logf(float x) {
  const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
  const float c0 = 0.3333331048488617f, c1 = -0.24999986588954926f, c2 = 0.20002226531505585f,
              c3 = -0.16669249534606934f, c4 = 0.14218655228614807f, c5 = -0.12384366989135742f,
              c6 = 0.1187300980091095f, c7 = -0.1178971379995346f, c8 = 0.06915182620286942f;
  uint32_t ix = f2u(x);
  ix += 0x3f800000u - 0x3f3504f3u;
  int k = (int)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  float f = u2f(ix) - 1.0f;
  float p = c8;
  p = __fmaf_rn(p, f, c7); p = __fmaf_rn(p, f, c6); p = __fmaf_rn(p, f, c5); p = __fmaf_rn(p, f, c4);
  p = __fmaf_rn(p, f, c3); p = __fmaf_rn(p, f, c2); p = __fmaf_rn(p, f, c1); p = __fmaf_rn(p, f, c0);
  float z = f * f;
  float dk = (float)k;
  float y = (f * z) * p;
  y = __fmaf_rn(dk, ln2_lo, y);
  y = __fmaf_rn(-0.5f, z, y);
  return __fmaf_rn(dk, ln2_hi, f + y);
}

// exp(x) for moderate |x|: x = n ln2 + r, degree-6 Taylor polynomial in r, scale by 2^n.
__device__ __forceinline__ float d_expf(float x) {
  const float log2e = 1.44269504089f, ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
  float fn = floorf(x * log2e + 0.5f);
  float r = x - fn * ln2_hi;
  r = r - fn * ln2_lo;
  float p = 1.0f / 720.0f;
  p = p * r + 1.0f / 120.0f;
  p = p * r + 1.0f / 24.0f;
  p = p * r + 1.0f / 6.0f;
  p = p * r + 0.5f;
  p = p * r + 1.0f;
  p = p * r + 1.0f;
  int n = (int)fn;
  if (n < -126) return 0.0f;
  return p * u2f((uint32_t)(n + 127) << 23);
}

// sin and cos of 2*pi*t, t in [0,1): quadrant reduction, |a| <= pi/4 polynomials.
__device__ __forceinline__ void d_sincos2pi(float t, float &s, float &c) {
  float y = t * 4.0f;
  float fk = floorf(y + 0.5f);
  float r = y - fk;
  float a = r * 1.57079632679f;
  float a2 = a * a;
  float sp = a2 * 2.75573192e-6f + -1.98412698e-4f;
  sp = sp * a2 + 8.33333333e-3f;
  sp = sp * a2 + -1.66666667e-1f;
  float sn = a + a * a2 * sp;
  float cp = a2 * 2.48015873e-5f + -1.38888889e-3f;
  cp = cp * a2 + 4.16666667e-2f;
  cp = cp * a2 + -0.5f;
  float cs = 1.0f + a2 * cp;
  int q = ((int)fk) & 3;
  if (q == 0) { s = sn; c = cs; }
  else if (q == 1) { s = cs; c = -sn; }
  else if (q == 2) { s = -sn; c = -cs; }
  else { s = -cs; c = sn; }
}

// Philox4x32-10 counter-based generator (Salmon et al., SC'11). key = run seed,
// counter = (history id lo, hi, block number, 0).
struct Rng {
  uint32_t k0, k1, h0, h1, blk;
  uint32_t o0, o1, o2, o3;
  int pos;
};

__device__ __forceinline__ void philox_block(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                             uint32_t c2, uint32_t c3, uint32_t &r0, uint32_t &r1,
                                             uint32_t &r2, uint32_t &r3) {
#pragma unroll 1
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  r0 = c0; r1 = c1; r2 = c2; r3 = c3;
}

// two consecutive blocks (blk, blk+1) of one history in one rolled loop: two independent dependency
// chains per round, 10 short iterations instead of 200 straight-line instructions
__device__ __forceinline__ void philox_block2(uint32_t k0, uint32_t k1, uint32_t h0, uint32_t h1, uint32_t blk,
                                              uint32_t out[8]) {
  uint32_t a0 = h0, a1 = h1, a2 = blk, a3 = 0u;
  uint32_t b0 = h0, b1 = h1, b2 = blk + 1u, b3 = 0u;
#pragma unroll kPhilox2Unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t ah0 = __umulhi(0xD2511F53u, a0), al0 = 0xD2511F53u * a0;
    uint32_t ah1 = __umulhi(0xCD9E8D57u, a2), al1 = 0xCD9E8D57u * a2;
    uint32_t bh0 = __umulhi(0xD2511F53u, b0), bl0 = 0xD2511F53u * b0;
    uint32_t bh1 = __umulhi(0xCD9E8D57u, b2), bl1 = 0xCD9E8D57u * b2;
    uint32_t an0 = ah1 ^ a1 ^ k0, an2 = ah0 ^ a3 ^ k1;
    uint32_t bn0 = bh1 ^ b1 ^ k0, bn2 = bh0 ^ b3 ^ k1;
    a0 = an0; a1 = al1; a2 = an2; a3 = al0;
    b0 = bn0; b1 = bl1; b2 = bn2; b3 = bl0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
  out[4] = b0; out[5] = b1; out[6] = b2; out[7] = b3;
}

__device__ __forceinline__ float u32_to_uniform(uint32_t x) {
  // uniform in (0,1): the top 24 bits, centred in their bin
  // (the product is exact, so the fused form rounds once exactly like multiply-then-add)
  return __fmaf_rn((float)(x >> 8), 5.9604644775390625e-8f, 2.98023223876953125e-8f);
}

__device__ __forceinline__ void rng_init(Rng &r, uint64_t seed, uint64_t hist, uint32_t blk) {
  r.k0 = (uint32_t)seed; r.k1 = (uint32_t)(seed >> 32);
  r.h0 = (uint32_t)hist; r.h1 = (uint32_t)(hist >> 32);
  r.blk = blk; r.pos = 4;
}

// A phase call starts on a fresh block (DESIGN.md "random-number protocol").
__device__ __forceinline__ void rng_begin(Rng &r) { r.pos = 4; }

// one copy of the 10 rounds for all the lazy draws of the rare phases (code size)
__device__ __forceinline__ void rng_refill(Rng &r) {
  philox_block(r.k0, r.k1, r.h0, r.h1, r.blk, 0u, r.o0, r.o1, r.o2, r.o3);
  r.blk += 1u;
  r.pos = 0;
}

__device__ __forceinline__ float rng_uniform(Rng &r) {
  if (r.pos == 4) rng_refill(r);
  uint32_t x = r.o0;
  r.o0 = r.o1; r.o1 = r.o2; r.o2 = r.o3;   // shift instead of a 4-way select on r.pos
  r.pos += 1;
  return u32_to_uniform(x);
}

}  // namespace gb2
