// transport.cuh — device data model + the photon transport kernel (sm_100a).
//
// One thread owns one photon history at a time (persistent threads, refilled from a global
// history counter; Philox stream = history id).  Per history:
//   primary generation  -> world flight (analytic, homogeneous world material)
//   -> Woodcock delta tracking through the label volume against the majorant of the
//      materials present -> Compton (Klein-Nishina x scattering function), Rayleigh (form
//      factor), photoelectric, conversion -> kerma deposit into a per-history coalescing
//      buffer -> at history end one (value, value^2) pair per touched voxel is added to the
//      fp64 grids with warp-aggregated atomics.
// Reference semantics restated (OpenGATE/opengate @ 9b07951): see include/gate_b200.h for the
// file:line of each replaced interface; the oracle (oracle/transport_oracle.c) is the scalar
// statement of the same algorithm.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/gate_b200.h"
#include "device_math.cuh"

namespace gb2 {

constexpr int kMaxDose = 4;
constexpr int kMaxPhsp = 4;
constexpr int kDepBuf = 32;            // per-history coalescing buffer entries
constexpr float kNudge = 1.0e-3f;      // mm pushed across a boundary
constexpr float kBig = 3.0e38f;
constexpr float kMc2 = 0.51099891f;
constexpr float kKPerMeV = 268.17264f; // photon wave number (1/bohr) per MeV
constexpr float kCmmPerNs = 299.792458f;
constexpr int kMaxIter = 1000000;

struct Affine { float m[9]; float t[3]; };

struct DSource {
  int pos_type, dir_type, ene_type, dir_relative;
  float half[3], radius, sigma_x, sigma_y;
  Affine pos;
  float dir_rot[9];
  float cos_tmin, cos_tmax, phi0_turn, dphi_turn;
  float mom[3], focus[3];
  float mono, sigma, emin, emax, weight;
  int n_spec;
  const float *spec_e, *spec_cdf;
  int vox_n[3];
  float vox_sp[3], vox_org[3];
  const float *cdf_z, *cdf_y, *cdf_x;
};

struct DDose {
  int size[3];
  Affine v2i;
  int flags, hit_type;
  double *grid[5];
};

struct DPhsp {
  float half[3];
  Affine w2b;
  int steps;
  unsigned long long capacity;
  float *ke, *pos, *dir, *w, *t;
  long long *ev;
  int *fl;
  unsigned long long *n_rows, *n_dropped;
};

// indices into the device statistics array (unsigned long long)
enum { S_HIST = 0, S_TRACKS, S_STEPS, S_TENT, S_PHOT, S_COMPT, S_RAYL, S_CONV, S_DEP, S_FLUSH,
       S_OVER, S_ESC, S_ABSORBED, S_COUNT };

struct Params {
  unsigned long long seed;
  int nmat, ne, nq, pair_mode;
  float ln_emin, inv_dlne, e_cut;
  float ln_q2min, inv_dlnq2;
  const float *dens, *mu, *cph, *cco, *cra, *maj, *q2n, *ray, *sf;
  float wh[3];
  int world_mat;
  int has_vol;
  int n[3];
  float half[3], inv_sp[3], sp[3];
  Affine w2v;
  const void *labels;
  int ndose;
  DDose dose[kMaxDose];
  int nphsp;
  DPhsp phsp[kMaxPhsp];
  DSource src;
  unsigned long long first, count;
  unsigned long long *next;      // history counter of this launch
  unsigned long long *stats;     // [S_COUNT]
  double *edep_total;
};

// ------------------------------------------------------------------------------ helpers

__device__ __forceinline__ void aff_apply(const Affine &a, const float x[3], float y[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
    y[r] = a.m[3 * r] * x[0] + a.m[3 * r + 1] * x[1] + a.m[3 * r + 2] * x[2] + a.t[r];
}
__device__ __forceinline__ void rot_apply(const float m[9], const float x[3], float y[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) y[r] = m[3 * r] * x[0] + m[3 * r + 1] * x[1] + m[3 * r + 2] * x[2];
}
__device__ __forceinline__ void aff_apply_inv(const Affine &a, const float y[3], float x[3]) {
  float v0 = y[0] - a.t[0], v1 = y[1] - a.t[1], v2 = y[2] - a.t[2];
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = a.m[c] * v0 + a.m[3 + c] * v1 + a.m[6 + c] * v2;
}
__device__ __forceinline__ void rot_apply_inv(const float m[9], const float y[3], float x[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = m[c] * y[0] + m[3 + c] * y[1] + m[6 + c] * y[2];
}

// slab test against an origin-centred box: entry (>=0) and exit distance plus the axes of the
// entry / exit faces (-1 when the start point is inside). A zero-length chord is a miss.
__device__ __forceinline__ bool ray_box_ax(const float p[3], const float d[3], const float h[3],
                                           float &t_in, float &t_out, int &a_in, int &a_out) {
  float tmin = -kBig, tmax = kBig;
  int ain = -1, aout = -1;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (d[a] != 0.0f) {
      float inv = 1.0f / d[a];
      float t1 = (-h[a] - p[a]) * inv;
      float t2 = (h[a] - p[a]) * inv;
      if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
      if (t1 > tmin) { tmin = t1; ain = a; }
      if (t2 < tmax) { tmax = t2; aout = a; }
    } else if (p[a] < -h[a] || p[a] > h[a]) {
      return false;
    }
  }
  float t0 = tmin > 0.0f ? tmin : 0.0f;
  if (tmax <= t0) return false;
  t_in = t0; t_out = tmax;
  a_in = tmin > 0.0f ? ain : -1;
  a_out = aout;
  return true;
}
__device__ __forceinline__ bool ray_box(const float p[3], const float d[3], const float h[3],
                                        float &t_in, float &t_out) {
  int ai, ao;
  return ray_box_ax(p, d, h, t_in, t_out, ai, ao);
}

// Put coordinate `a` just inside / outside the face the ray crosses, so that the slab test and
// the voxel-index test agree about the side even at grazing incidence (a push along the ray can
// be smaller than one ulp of the coordinate).
__device__ __forceinline__ void snap_to_face(float p[3], const float d[3], const float h[3], int a,
                                             bool outward) {
  if (a < 0) return;
  // selects instead of p[a]/d[a]: a run-time index would push the photon state to local memory
  const float ha = a == 0 ? h[0] : (a == 1 ? h[1] : h[2]);
  const float da = a == 0 ? d[0] : (a == 1 ? d[1] : d[2]);
  float face = outward ? (ha * (1.0f + 1.0e-6f) + 1.0e-5f) : (ha * (1.0f - 1.0e-6f) - 1.0e-5f);
  float v;
  if (outward) v = da > 0.0f ? face : -face;
  else v = da > 0.0f ? -face : face;
  if (a == 0) p[0] = v; else if (a == 1) p[1] = v; else p[2] = v;
}

__device__ __forceinline__ bool inside_box(const float p[3], const float h[3]) {
  return p[0] > -h[0] && p[0] < h[0] && p[1] > -h[1] && p[1] < h[1] && p[2] > -h[2] && p[2] < h[2];
}

__device__ __forceinline__ float tab(const float *__restrict__ t, int ne, int m, int ie, float fe) {
  const float *r = t + (size_t)m * ne + ie;
  float a = __ldg(r), b = __ldg(r + 1);
  return a + fe * (b - a);
}

__device__ __forceinline__ int lower_bound_f(const float *__restrict__ a, int n, float p) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < p) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------ per-thread state

struct Photon {
  float p[3], d[3];   // world frame outside the volume, volume frame inside
  float E, w, path, seg, fe;
  int ie;
  int iter;
};

struct History {
  Rng rng;
  unsigned long long id;
  int nbuf;
  int inside_mask;     // bit k: currently inside phase-space box k
  int stored;
  int pend;            // 0 none, 1 pending photon stored in the volume frame, 2 world frame
  float pend_p[3], pend_d[3];
  float edep;          // everything deposited by this history
  // coalescing buffer (local memory)
  int bkey[kDepBuf];
  float be[kDepBuf], bd[kDepBuf];
  int bn[kDepBuf];
};

struct Counters {      // per-thread, flushed at the end of the kernel
  unsigned int hist, tracks, steps, tent, phot, compt, rayl, conv, dep, flush, over, esc, absorbed;
  double edep;
};

// ------------------------------------------------------------------------------ scoring

// Add the coalesced (voxel, history) entries to the fp64 grids.  Lanes of the warp that flush
// at the same time and hit the same voxel are combined first (one atomic per distinct
// address); squares are summed per history, never of the combined value.
__device__ __forceinline__ void flush_buffer(const Params &P, History &h, Counters &c) {
  const unsigned lane = threadIdx.x & 31u;
  for (int j = 0; j < h.nbuf; ++j) {
    const int key = h.bkey[j];
    const int a = (int)((unsigned)key >> 28);
    const int v = key & 0x0fffffff;
    double e = (double)h.be[j], d = (double)h.bd[j];
    double e2 = e * e, d2 = d * d, n = (double)h.bn[j];
    const unsigned act = __activemask();
    const unsigned grp = __match_any_sync(act, key);
    const int leader = __ffs(grp) - 1;
    if (grp != (1u << lane)) {
      // serial accumulation into the leader, every member of the group walks the same list
      double se = 0.0, se2 = 0.0, sd = 0.0, sd2 = 0.0, sn = 0.0;
      for (unsigned rem = grp; rem; rem &= rem - 1) {
        const int src = __ffs(rem) - 1;
        se += __shfl_sync(grp, e, src);
        se2 += __shfl_sync(grp, e2, src);
        sd += __shfl_sync(grp, d, src);
        sd2 += __shfl_sync(grp, d2, src);
        sn += __shfl_sync(grp, n, src);
      }
      e = se; e2 = se2; d = sd; d2 = sd2; n = sn;
    }
    if ((int)lane == leader) {
      const DDose &D = P.dose[a];
      atomicAdd(D.grid[GATE_B200_GRID_EDEP] + v, e);
      if (D.flags & GATE_B200_SCORE_EDEP_SQ) atomicAdd(D.grid[GATE_B200_GRID_EDEP_SQ] + v, e2);
      if (D.flags & GATE_B200_SCORE_DOSE) atomicAdd(D.grid[GATE_B200_GRID_DOSE] + v, d);
      if (D.flags & GATE_B200_SCORE_DOSE_SQ) atomicAdd(D.grid[GATE_B200_GRID_DOSE_SQ] + v, d2);
      if (D.flags & GATE_B200_SCORE_COUNTS) atomicAdd(D.grid[GATE_B200_GRID_COUNTS] + v, n);
    }
    c.flush++;
  }
  h.nbuf = 0;
}

// deposit e (already weighted) at the photon position; in_vol selects whether dose actors see it
__device__ __forceinline__ void deposit(const Params &P, History &h, Photon &g, Counters &c, float e,
                                        int m, bool in_vol, int vix, int viy, int viz) {
  h.edep += e;
  c.edep += (double)e;
  c.dep++;
  if (!in_vol) return;
  for (int a = 0; a < P.ndose; ++a) {
    const DDose &D = P.dose[a];
    float pos[3] = {g.p[0], g.p[1], g.p[2]};
    if (D.hit_type != GATE_B200_HIT_POST) {
      float back = g.seg;
      const int vi[3] = {vix, viy, viz};
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        float dd = g.d[ax];
        if (dd > 0.0f) {
          float lo = -P.half[ax] + (float)vi[ax] * P.sp[ax];
          float tb = (g.p[ax] - lo) / dd;
          if (tb < back) back = tb;
        } else if (dd < 0.0f) {
          float hi = -P.half[ax] + (float)(vi[ax] + 1) * P.sp[ax];
          float tb = (g.p[ax] - hi) / dd;
          if (tb < back) back = tb;
        }
      }
      if (back < 0.0f) back = 0.0f;
      float fr = 1.0f;
      if (D.hit_type == GATE_B200_HIT_RANDOM) fr = 1.0f - rng_uniform(h.rng);
      else if (D.hit_type == GATE_B200_HIT_MIDDLE) fr = 0.5f;
      float bk = back * fr;
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) pos[ax] = g.p[ax] - g.d[ax] * bk;
    }
    float ci[3];
    aff_apply(D.v2i, pos, ci);
    float fx = floorf(ci[0] + 0.5f), fy = floorf(ci[1] + 0.5f), fz = floorf(ci[2] + 0.5f);
    if (fx < 0.0f || fy < 0.0f || fz < 0.0f || fx >= (float)D.size[0] || fy >= (float)D.size[1] ||
        fz >= (float)D.size[2])
      continue;
    int v = (int)fx + D.size[0] * ((int)fy + D.size[1] * (int)fz);
    int key = (int)(((unsigned)a << 28) | (unsigned)v);
    float dd = e / __ldg(P.dens + m);
    int j;
    for (j = 0; j < h.nbuf; ++j)
      if (h.bkey[j] == key) break;
    if (j == h.nbuf) {
      if (h.nbuf == kDepBuf) {
        flush_buffer(P, h, c);
        c.over++;
        j = 0;
      }
      h.bkey[j] = key; h.be[j] = 0.0f; h.bd[j] = 0.0f; h.bn[j] = 0;
      h.nbuf++;
    }
    h.be[j] += e; h.bd[j] += dd; h.bn[j] += 1;
  }
}

// ------------------------------------------------------------------------------ physics

__device__ __forceinline__ void rotate_dir(float d[3], float cost, float sint, float sphi, float cphi) {
  float u = d[0], v = d[1], w = d[2];
  float nu, nv, nw;
  float aw = w < 0.0f ? -w : w;
  if (aw < 0.99999f) {
    float sz = sqrtf(1.0f - w * w);
    float a = sint / sz;
    nu = u * cost + a * (u * w * cphi - v * sphi);
    nv = v * cost + a * (v * w * cphi + u * sphi);
    nw = w * cost - sint * sz * cphi;
  } else {
    float sg = w < 0.0f ? -1.0f : 1.0f;
    nu = sint * cphi; nv = sint * sphi; nw = sg * cost;
  }
  float inv = 1.0f / sqrtf(nu * nu + nv * nv + nw * nw);
  d[0] = nu * inv; d[1] = nv * inv; d[2] = nw * inv;
}

__device__ __forceinline__ int q2_segment(const Params &P, float q2) {
  if (q2 < __ldg(P.q2n)) return -1;
  if (q2 >= __ldg(P.q2n + P.nq - 1)) return P.nq - 1;
  float t = (d_logf(q2) - P.ln_q2min) * P.inv_dlnq2;
  int j = (int)t;
  if (j < 0) j = 0;
  if (j > P.nq - 2) j = P.nq - 2;
  while (j < P.nq - 2 && __ldg(P.q2n + j + 1) <= q2) ++j;
  while (j > 0 && __ldg(P.q2n + j) > q2) --j;
  return j;
}
__device__ __forceinline__ float interp_q2(const Params &P, const float *__restrict__ row, float q2) {
  int j = q2_segment(P, q2);
  if (j < 0) return __ldg(row) * (q2 / __ldg(P.q2n));
  if (j >= P.nq - 1) return __ldg(row + P.nq - 1);
  float q0 = __ldg(P.q2n + j), q1 = __ldg(P.q2n + j + 1);
  float f = (q2 - q0) / (q1 - q0);
  float r0 = __ldg(row + j), r1 = __ldg(row + j + 1);
  return r0 + f * (r1 - r0);
}

__device__ __forceinline__ void energy_bin(const Params &P, Photon &g) {
  float t = (d_logf(g.E) - P.ln_emin) * P.inv_dlne;
  if (t < 0.0f) t = 0.0f;
  int i = (int)t;
  if (i > P.ne - 2) i = P.ne - 2;
  float f = t - (float)i;
  if (f > 1.0f) f = 1.0f;
  g.ie = i; g.fe = f;
}

// one real interaction in material m; returns true if the photon survives
__device__ __forceinline__ bool interact(const Params &P, History &h, Photon &g, Counters &c, int m,
                                         bool in_vol, int vix, int viy, int viz) {
  float u = rng_uniform(h.rng);
  float pph = tab(P.cph, P.ne, m, g.ie, g.fe);
  float pco = tab(P.cco, P.ne, m, g.ie, g.fe);
  float pra = tab(P.cra, P.ne, m, g.ie, g.fe);
  c.steps++;
  if (u < pph) {
    c.phot++;
    deposit(P, h, g, c, g.E * g.w, m, in_vol, vix, viy, viz);
    return false;
  }
  if (u < pco) {
    c.compt++;
    float E = g.E;
    float k = E / kMc2;
    float eps0 = 1.0f / (1.0f + 2.0f * k);
    float eps0sq = eps0 * eps0;
    float alpha1 = -d_logf(eps0);
    float alpha2 = alpha1 + 0.5f * (1.0f - eps0sq);
    float kq = kKPerMeV * E;
    float two_kq2 = 2.0f * kq * kq;
    const float *sfrow = P.sf + (size_t)m * P.nq;
    float eps = 1.0f, onecost = 0.0f, sint2 = 0.0f;
    for (int it = 0; it < 1000; ++it) {
      float u1 = rng_uniform(h.rng), u2 = rng_uniform(h.rng), u3 = rng_uniform(h.rng);
      float epssq;
      if (alpha1 > alpha2 * u1) {
        eps = d_expf(-alpha1 * u2);
        epssq = eps * eps;
      } else {
        epssq = eps0sq + (1.0f - eps0sq) * u2;
        eps = sqrtf(epssq);
      }
      onecost = (1.0f - eps) / (eps * k);
      sint2 = onecost * (2.0f - onecost);
      float grej = 1.0f - eps * sint2 / (1.0f + epssq);
      if (grej < u3) continue;
      float sfv = interp_q2(P, sfrow, two_kq2 * onecost);
      float u4 = rng_uniform(h.rng);
      if (u4 > sfv) continue;
      break;
    }
    float cost = 1.0f - onecost;
    float sint = sqrtf(sint2 > 0.0f ? sint2 : 0.0f);
    float sphi, cphi;
    d_sincos2pi(rng_uniform(h.rng), sphi, cphi);
    float E1 = eps * E;
    deposit(P, h, g, c, (E - E1) * g.w, m, in_vol, vix, viy, viz);
    rotate_dir(g.d, cost, sint, sphi, cphi);
    g.E = E1;
    g.seg = 0.0f;
    if (E1 < P.e_cut) {
      deposit(P, h, g, c, E1 * g.w, m, in_vol, vix, viy, viz);
      return false;
    }
    energy_bin(P, g);
    return true;
  }
  if (u < pra) {
    c.rayl++;
    float kq = kKPerMeV * g.E;
    float kq2 = kq * kq;
    const float *A = P.ray + (size_t)m * P.nq;
    float q2max = 4.0f * kq2;
    int jmax = q2_segment(P, q2max);
    float amax = interp_q2(P, A, q2max);
    float cost = 1.0f;
    const float a0 = __ldg(A);
    const float q20 = __ldg(P.q2n);
    for (int it = 0; it < 1000; ++it) {
      float target = rng_uniform(h.rng) * amax;
      float q2;
      if (target < a0) {
        q2 = q20 * (target / a0);
      } else {
        int lo = 0, hi = jmax < P.nq - 1 ? jmax + 1 : P.nq - 1;
        if (target >= __ldg(A + hi)) lo = hi > 0 ? hi - 1 : 0;
        else
          while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (__ldg(A + mid) <= target) lo = mid; else hi = mid;
          }
        float al = __ldg(A + lo), ah = __ldg(A + lo + 1);
        float den = ah - al;
        float f = den > 0.0f ? (target - al) / den : 0.0f;
        if (f > 1.0f) f = 1.0f;
        float ql = __ldg(P.q2n + lo), qh = __ldg(P.q2n + lo + 1);
        q2 = ql + f * (qh - ql);
      }
      if (q2 > q2max) q2 = q2max;
      cost = 1.0f - q2 / (2.0f * kq2);
      if (cost < -1.0f) cost = -1.0f;
      float u2 = rng_uniform(h.rng);
      if (0.5f * (1.0f + cost * cost) >= u2) break;
    }
    float sint = sqrtf(1.0f - cost * cost);
    float sphi, cphi;
    d_sincos2pi(rng_uniform(h.rng), sphi, cphi);
    rotate_dir(g.d, cost, sint, sphi, cphi);
    g.seg = 0.0f;
    return true;
  }
  c.conv++;
  if (P.pair_mode == 0) {
    deposit(P, h, g, c, g.E * g.w, m, in_vol, vix, viy, viz);
    return false;
  }
  deposit(P, h, g, c, (g.E - 2.0f * kMc2) * g.w, m, in_vol, vix, viy, viz);
  {
    float cost = 2.0f * rng_uniform(h.rng) - 1.0f;
    float sint = sqrtf(1.0f - cost * cost);
    float sphi, cphi;
    d_sincos2pi(rng_uniform(h.rng), sphi, cphi);
    g.d[0] = sint * cphi; g.d[1] = sint * sphi; g.d[2] = cost;
    g.E = kMc2;
    g.seg = 0.0f;
    energy_bin(P, g);
#pragma unroll
    for (int a = 0; a < 3; ++a) { h.pend_p[a] = g.p[a]; h.pend_d[a] = -g.d[a]; }
    h.pend = in_vol ? 1 : 2;
    c.tracks += 2;
  }
  return true;
}

// ------------------------------------------------------------------------------ primaries

__device__ __forceinline__ float gauss(Rng &r) {
  float u1 = rng_uniform(r), u2 = rng_uniform(r);
  float s, c;
  d_sincos2pi(u2, s, c);
  return sqrtf(-2.0f * d_logf(u1)) * c;
}

__device__ __noinline__ void generate_primary(const Params &P, History &h, Photon &g) {
  const DSource &S = P.src;
  Rng &r = h.rng;
  float l[3] = {0.0f, 0.0f, 0.0f};
  switch (S.pos_type) {
  case GATE_B200_POS_POINT: break;
  case GATE_B200_POS_BOX:
    for (int a = 0; a < 3; ++a) l[a] = (2.0f * rng_uniform(r) - 1.0f) * S.half[a];
    break;
  case GATE_B200_POS_SPHERE: {
    float R = S.radius;
    for (;;) {
      for (int a = 0; a < 3; ++a) l[a] = (2.0f * rng_uniform(r) - 1.0f) * R;
      if (l[0] * l[0] + l[1] * l[1] + l[2] * l[2] <= R * R) break;
    }
  } break;
  case GATE_B200_POS_DISC:
  case GATE_B200_POS_CYLINDER: {
    float R = S.radius;
    if (R > 0.0f)
      for (;;) {
        l[0] = (2.0f * rng_uniform(r) - 1.0f) * R;
        l[1] = (2.0f * rng_uniform(r) - 1.0f) * R;
        if (l[0] * l[0] + l[1] * l[1] <= R * R) break;
      }
    if (S.pos_type == GATE_B200_POS_DISC) {
      if (S.sigma_x > 0.0f) l[0] += gauss(r) * S.sigma_x;
      if (S.sigma_y > 0.0f) l[1] += gauss(r) * S.sigma_y;
    } else {
      l[2] = (2.0f * rng_uniform(r) - 1.0f) * S.half[2];
    }
  } break;
  case GATE_B200_POS_VOXELS: {
    const int nx = S.vox_n[0], ny = S.vox_n[1], nz = S.vox_n[2];
    int i, j, k, tries;
    tries = 0;
    do { i = lower_bound_f(S.cdf_z, nz, rng_uniform(r)); } while (i >= nz && ++tries < 1000);
    if (i >= nz) i = nz - 1;
    tries = 0;
    do { j = lower_bound_f(S.cdf_y + (size_t)i * ny, ny, rng_uniform(r)); } while (j >= ny && ++tries < 1000);
    if (j >= ny) j = ny - 1;
    tries = 0;
    do { k = lower_bound_f(S.cdf_x + ((size_t)i * ny + j) * nx, nx, rng_uniform(r)); } while (k >= nx && ++tries < 1000);
    if (k >= nx) k = nx - 1;
    l[0] = S.vox_org[0] + (float)k * S.vox_sp[0];
    l[1] = S.vox_org[1] + (float)j * S.vox_sp[1];
    l[2] = S.vox_org[2] + (float)i * S.vox_sp[2];
    l[0] += (rng_uniform(r) - 0.5f) * S.vox_sp[0];
    l[1] += (rng_uniform(r) - 0.5f) * S.vox_sp[1];
    l[2] += (rng_uniform(r) - 0.5f) * S.vox_sp[2];
  } break;
  }
  aff_apply(S.pos, l, g.p);

  switch (S.dir_type) {
  case GATE_B200_DIR_ISO: {
    float cost = S.cos_tmin - rng_uniform(r) * (S.cos_tmin - S.cos_tmax);
    float s2 = 1.0f - cost * cost;
    float sint = sqrtf(s2 > 0.0f ? s2 : 0.0f);
    float turn = S.phi0_turn + S.dphi_turn * rng_uniform(r);
    turn = turn - floorf(turn);
    float sp, cp;
    d_sincos2pi(turn, sp, cp);
    float v[3] = {-sint * cp, -sint * sp, -cost};
    if (S.dir_relative) rot_apply(S.dir_rot, v, g.d);
    else { g.d[0] = v[0]; g.d[1] = v[1]; g.d[2] = v[2]; }
  } break;
  case GATE_B200_DIR_MOMENTUM:
    g.d[0] = S.mom[0]; g.d[1] = S.mom[1]; g.d[2] = S.mom[2];
    break;
  case GATE_B200_DIR_FOCUSED: {
    float v0 = S.focus[0] - g.p[0], v1 = S.focus[1] - g.p[1], v2 = S.focus[2] - g.p[2];
    float inv = 1.0f / sqrtf(v0 * v0 + v1 * v1 + v2 * v2);
    g.d[0] = v0 * inv; g.d[1] = v1 * inv; g.d[2] = v2 * inv;
  } break;
  }

  float E = 0.0f;
  switch (S.ene_type) {
  case GATE_B200_ENE_MONO: E = S.mono; break;
  case GATE_B200_ENE_GAUSS: E = S.mono + gauss(r) * S.sigma; break;
  case GATE_B200_ENE_RANGE: E = S.emin + rng_uniform(r) * (S.emax - S.emin); break;
  case GATE_B200_ENE_SPECTRUM_DISCRETE: {
    float p = rng_uniform(r);
    int i = 0;
    while (i < S.n_spec - 1 && p >= __ldg(S.spec_cdf + i)) ++i;
    E = __ldg(S.spec_e + i);
  } break;
  case GATE_B200_ENE_SPECTRUM_HISTOGRAM: {
    float p = rng_uniform(r);
    int i = 0;
    while (i < S.n_spec - 1 && p >= __ldg(S.spec_cdf + i)) ++i;
    float e0 = __ldg(S.spec_e + i), e1 = __ldg(S.spec_e + i + 1);
    E = e0 + rng_uniform(r) * (e1 - e0);
  } break;
  case GATE_B200_ENE_SPECTRUM_HISTOGRAM_LINEAR: {
    float p0 = rng_uniform(r);
    int i = 0;
    while (i < S.n_spec - 2 && p0 >= __ldg(S.spec_cdf + i)) ++i;
    float a = (__ldg(S.spec_e + i) + __ldg(S.spec_e + i + 1)) * 0.5f;
    float b = (__ldg(S.spec_e + i + 1) + __ldg(S.spec_e + i + 2)) * 0.5f;
    float dd = __ldg(S.spec_cdf + i + 1) - __ldg(S.spec_cdf + i);
    float ad = dd < 0.0f ? -dd : dd;
    if (ad < 1.1920929e-7f) {
      E = a + rng_uniform(r) * (b - a);
    } else {
      float alpha = dd / (b - a);
      float beta = __ldg(S.spec_cdf + i) - alpha * a;
      float norm = 0.5f * alpha * (b * b - a * a) + beta * (b - a);
      float p = rng_uniform(r);
      float sq = sqrtf((alpha * a + beta) * (alpha * a + beta) + 2.0f * alpha * norm * p);
      float x = (-beta + sq) / alpha;
      E = ((x - a) * (x - b) <= 0.0f) ? x : (-beta - sq) / alpha;
    }
  } break;
  case GATE_B200_ENE_CDF: {
    float x = rng_uniform(r);
    int idx = lower_bound_f(S.spec_cdf, S.n_spec, x) - 1;
    if (idx < 0) idx = 0;
    if (idx > S.n_spec - 2) idx = S.n_spec - 2;
    float c0 = __ldg(S.spec_cdf + idx), c1 = __ldg(S.spec_cdf + idx + 1);
    float ratio = (x - c0) / (c1 - c0);
    float e0 = __ldg(S.spec_e + idx), e1 = __ldg(S.spec_e + idx + 1);
    E = e0 + ratio * (e1 - e0);
  } break;
  }
  g.E = E;
  g.w = S.weight;
}

// ------------------------------------------------------------------------------ phase space

__device__ __forceinline__ void phsp_store(const Params &P, History &h, const Photon &g, int k, int flag) {
  const DPhsp &Q = P.phsp[k];
  h.stored = 1;
  unsigned long long idx = atomicAdd(Q.n_rows, 1ull);
  if (idx >= Q.capacity) {
    atomicAdd(Q.n_dropped, 1ull);
    return;
  }
  Q.ke[idx] = g.E;
#pragma unroll
  for (int a = 0; a < 3; ++a) { Q.pos[3 * idx + a] = g.p[a]; Q.dir[3 * idx + a] = g.d[a]; }
  Q.w[idx] = g.w;
  Q.t[idx] = g.path / kCmmPerNs;
  Q.ev[idx] = (long long)h.id;
  Q.fl[idx] = flag;
}

// ------------------------------------------------------------------------------ regions

__device__ __forceinline__ void to_volume(const Params &P, Photon &g) {
  float pl[3], dl[3];
  aff_apply(P.w2v, g.p, pl);
  rot_apply(P.w2v.m, g.d, dl);
#pragma unroll
  for (int a = 0; a < 3; ++a) { g.p[a] = pl[a]; g.d[a] = dl[a]; }
}
__device__ __forceinline__ void to_world(const Params &P, Photon &g) {
  float pw[3], dw[3];
  aff_apply_inv(P.w2v, g.p, pw);
  rot_apply_inv(P.w2v.m, g.d, dw);
#pragma unroll
  for (int a = 0; a < 3; ++a) { g.p[a] = pw[a]; g.d[a] = dw[a]; }
}

enum { ST_FETCH = 0, ST_WORLD = 1, ST_VOLUME = 2, ST_DEAD = 3, ST_INTERACT_W = 4, ST_INTERACT_V = 5, ST_DONE = 6 };

// place a photon given in the world frame into its region; returns the state
__device__ __forceinline__ int locate(const Params &P, History &h, Photon &g) {
  int st = ST_WORLD;
  if (P.has_vol) {
    float pl[3];
    aff_apply(P.w2v, g.p, pl);
    if (inside_box(pl, P.half)) { to_volume(P, g); st = ST_VOLUME; }
  }
  h.inside_mask = 0;
  for (int k = 0; k < P.nphsp; ++k) {
    float pb[3];
    aff_apply(P.phsp[k].w2b, g.p, pb);
    if (st == ST_WORLD && inside_box(pb, P.phsp[k].half)) h.inside_mask |= (1 << k);
  }
  return st;
}

// one step of a photon travelling in the world material (outside the image volume)
__device__ __forceinline__ int world_step(const Params &P, History &h, Photon &g, Counters &c) {
  float t_exit, t0;
  if (!inside_box(g.p, P.wh) || !ray_box(g.p, g.d, P.wh, t0, t_exit)) {
    c.esc++;
    return ST_DEAD;
  }
  int ev = 0;
  float mu = tab(P.mu, P.ne, P.world_mat, g.ie, g.fe);
  float t_best = -d_logf(rng_uniform(h.rng)) / mu;
  if (t_exit < t_best) { t_best = t_exit; ev = 1; }
  int ax_in = -1;
  if (P.has_vol) {
    float pl[3], dl[3], ti, to;
    int ai, ao;
    aff_apply(P.w2v, g.p, pl);
    rot_apply(P.w2v.m, g.d, dl);
    if (ray_box_ax(pl, dl, P.half, ti, to, ai, ao) && ti < t_best) { t_best = ti; ev = 2; ax_in = ai; }
  }
  for (int k = 0; k < P.nphsp; ++k) {
    float pb[3], db[3], ti, to;
    aff_apply(P.phsp[k].w2b, g.p, pb);
    rot_apply(P.phsp[k].w2b.m, g.d, db);
    if (!ray_box(pb, db, P.phsp[k].half, ti, to)) {
      h.inside_mask &= ~(1 << k);
      continue;
    }
    if (h.inside_mask & (1 << k)) {
      if (to < t_best) { t_best = to; ev = 4 + 2 * k; }
    } else if (ti < t_best) { t_best = ti; ev = 3 + 2 * k; }
  }
  float step = (ev >= 3) ? t_best + kNudge : t_best;
  if (ev == 1) { c.esc++; g.path += t_best; return ST_DEAD; }
#pragma unroll
  for (int a = 0; a < 3; ++a) g.p[a] = g.p[a] + g.d[a] * step;
  g.path += step;
  if (ev == 0) {
    return ST_INTERACT_W;   // real interaction in the world material, done in the interaction phase
  } else if (ev == 2) {
    c.steps++;
    to_volume(P, g);
    snap_to_face(g.p, g.d, P.half, ax_in, false);
    g.seg = 0.0f;
    return ST_VOLUME;
  }
  int k = (ev - 3) >> 1;
  c.steps++;
  if (((ev - 3) & 1) == 0) {
    h.inside_mask |= (1 << k);
    if (P.phsp[k].steps & GATE_B200_PHSP_ENTERING) phsp_store(P, h, g, k, GATE_B200_PHSP_ENTERING);
  } else {
    h.inside_mask &= ~(1 << k);
    if (P.phsp[k].steps & GATE_B200_PHSP_EXITING) phsp_store(P, h, g, k, GATE_B200_PHSP_EXITING);
  }
  return ST_WORLD;
}

// one Woodcock candidate collision inside the image volume
template <typename LabelT>
__device__ __forceinline__ int volume_step(const Params &P, const float *__restrict__ s_maj, History &h,
                                           Photon &g, Counters &c, int &mat) {
  float m0 = s_maj[g.ie], m1 = s_maj[g.ie + 1];
  float maj = m0 + g.fe * (m1 - m0);
  float s = -d_logf(rng_uniform(h.rng)) / maj;
  float q0 = g.p[0] + g.d[0] * s, q1 = g.p[1] + g.d[1] * s, q2 = g.p[2] + g.d[2] * s;
  float fx = (q0 + P.half[0]) * P.inv_sp[0];
  float fy = (q1 + P.half[1]) * P.inv_sp[1];
  float fz = (q2 + P.half[2]) * P.inv_sp[2];
  if (fx < 0.0f || fy < 0.0f || fz < 0.0f || fx >= (float)P.n[0] || fy >= (float)P.n[1] ||
      fz >= (float)P.n[2]) {
    float ti, to;
    int ai, ao;
    if (!ray_box_ax(g.p, g.d, P.half, ti, to, ai, ao)) { to = 0.0f; ao = -1; }
#pragma unroll
    for (int a = 0; a < 3; ++a) g.p[a] = g.p[a] + g.d[a] * to;
    snap_to_face(g.p, g.d, P.half, ao, true);
    g.path += to;
    c.steps++;
    to_world(P, g);
    return ST_WORLD;
  }
  g.p[0] = q0; g.p[1] = q1; g.p[2] = q2;
  g.path += s;
  g.seg += s;
  c.tent++;
  int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  size_t flat = (size_t)ix + (size_t)P.n[0] * ((size_t)iy + (size_t)P.n[1] * (size_t)iz);
  int m = (int)__ldg(reinterpret_cast<const LabelT *>(P.labels) + flat);
  float mu = tab(P.mu, P.ne, m, g.ie, g.fe);
  if (rng_uniform(h.rng) * maj <= mu) {
    mat = m;                // accepted: the interaction itself runs in the interaction phase
    return ST_INTERACT_V;
  }
  return ST_VOLUME;
}

// ------------------------------------------------------------------------------ kernel

// The kernel is a warp-level phase scheduler.  Every lane carries one photon whose state says
// which kind of work it needs next (refill, world flight, Woodcock step, interaction).  In
// each round the warp votes and runs ONE phase, the one most lanes are waiting for, so that
// the long and rare code paths (interaction sampling, primary generation) execute with many
// lanes at once instead of being dragged through on every round by one or two lanes.
template <typename LabelT>
__global__ void __launch_bounds__(256) transport_kernel(const __grid_constant__ Params P) {
  extern __shared__ float s_maj[];
  for (int i = threadIdx.x; i < P.ne; i += blockDim.x) s_maj[i] = P.maj[i];
  __syncthreads();

  History h;
  Photon g;
  Counters c;
  c.hist = c.tracks = c.steps = c.tent = c.phot = c.compt = c.rayl = c.conv = 0;
  c.dep = c.flush = c.over = c.esc = c.absorbed = 0;
  c.edep = 0.0;
  h.nbuf = 0; h.pend = 0; h.stored = 0; h.inside_mask = 0; h.edep = 0.0f; h.id = 0;
  int st = ST_FETCH;
  int mat = 0;            // material of the pending interaction
  bool open = false;      // a history is in flight (must be closed before the next fetch)
  const unsigned lane = threadIdx.x & 31u;
  const unsigned FULL = 0xffffffffu;

  for (;;) {
    const unsigned bV = __ballot_sync(FULL, st == ST_VOLUME);
    const unsigned bI = __ballot_sync(FULL, st == ST_INTERACT_V || st == ST_INTERACT_W);
    const unsigned bW = __ballot_sync(FULL, st == ST_WORLD);
    const unsigned bF = __ballot_sync(FULL, st == ST_FETCH);
    if (!(bV | bI | bW | bF)) break;
    const int nV = __popc(bV), nI = __popc(bI), nW = __popc(bW), nF = __popc(bF);

    if (nV >= nI && nV >= nW && nV >= nF) {
      // ---------------- Woodcock phase
      if (st == ST_VOLUME) {
        st = volume_step<LabelT>(P, s_maj, h, g, c, mat);
        if (++g.iter >= kMaxIter) st = ST_DEAD;
      }
    } else if (nI >= nW && nI >= nF) {
      // ---------------- interaction phase
      if (st == ST_INTERACT_V || st == ST_INTERACT_W) {
        const bool in_vol = st == ST_INTERACT_V;
        int ix = 0, iy = 0, iz = 0;
        if (in_vol) {   // same expressions as in volume_step: same voxel
          ix = (int)((g.p[0] + P.half[0]) * P.inv_sp[0]);
          iy = (int)((g.p[1] + P.half[1]) * P.inv_sp[1]);
          iz = (int)((g.p[2] + P.half[2]) * P.inv_sp[2]);
        }
        const bool alive = interact(P, h, g, c, in_vol ? mat : P.world_mat, in_vol, ix, iy, iz);
        st = alive ? (in_vol ? ST_VOLUME : ST_WORLD) : ST_DEAD;
      }
    } else if (nW >= nF) {
      // ---------------- world phase
      if (st == ST_WORLD) {
        st = world_step(P, h, g, c);
        if (++g.iter >= kMaxIter) st = ST_DEAD;
      }
    } else {
      // ---------------- refill phase: close the finished history, claim the next id
      if (st == ST_FETCH) {
        if (open) {
          flush_buffer(P, h, c);
          if (P.nphsp > 0 && !h.stored) c.absorbed++;
          open = false;
        }
        const unsigned act = __activemask();
        const int leader = __ffs(act) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(P.next, (unsigned long long)__popc(act));
        base = __shfl_sync(act, base, leader);
        const unsigned long long idx = base + (unsigned long long)__popc(act & ((1u << lane) - 1u));
        if (idx >= P.count) {
          st = ST_DONE;
        } else {
          h.id = P.first + idx;
          rng_init(h.rng, P.seed, h.id);
          h.nbuf = 0; h.pend = 0; h.stored = 0; h.edep = 0.0f;
          generate_primary(P, h, g);
          g.path = 0.0f; g.seg = 0.0f; g.iter = 0;
          c.hist++; c.tracks++;
          open = true;
          st = locate(P, h, g);
          if (g.E < P.e_cut) {
            c.edep += (double)(g.E * g.w);
            st = ST_DEAD;
          } else {
            energy_bin(P, g);
          }
        }
      }
    }

    if (st == ST_DEAD) {
      if (h.pend) {
        // second annihilation photon of a conversion (pair_mode 1)
        float pw[3], dw[3];
        if (h.pend == 1) {
          aff_apply_inv(P.w2v, h.pend_p, pw);
          rot_apply_inv(P.w2v.m, h.pend_d, dw);
        } else {
#pragma unroll
          for (int a = 0; a < 3; ++a) { pw[a] = h.pend_p[a]; dw[a] = h.pend_d[a]; }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) { g.p[a] = pw[a]; g.d[a] = dw[a]; }
        g.E = kMc2; g.seg = 0.0f; g.iter = 0;
        h.pend = 0;
        st = locate(P, h, g);
        energy_bin(P, g);
      } else {
        st = ST_FETCH;   // the history is closed (flushed) in the refill phase
      }
    }
  }

  // per-thread counters -> global statistics (warp reduce, one atomic per warp and counter)
  unsigned int vals[S_COUNT] = {c.hist, c.tracks, c.steps, c.tent, c.phot, c.compt, c.rayl, c.conv,
                                c.dep, c.flush, c.over, c.esc, c.absorbed};
  __syncwarp();
#pragma unroll
  for (int i = 0; i < S_COUNT; ++i) {
    unsigned long long v = vals[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) atomicAdd(P.stats + i, v);
  }
  double e = c.edep;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if (lane == 0 && e != 0.0) atomicAdd(P.edep_total, e);
}

}  // namespace gb2
