/* transport_oracle.c — CPU ORACLE (test infrastructure, NOT a product path).
 *
 * A scalar, single-precision restatement of the photon-in-image-volume dose path, written
 * independently of opengate_b200/csrc/ but to the same arithmetic specification
 * (DESIGN.md section "Arithmetic contract"): IEEE float +,-,*,/ ,sqrt and the fused multiply-add
 * exactly where fmaf() is written (never by compiler contraction: compile with
 * -ffp-contract=off), own log/exp/sincos polynomials, Philox4x32-10
 * keyed by (seed, history id).  Under that contract the CUDA engine must reproduce every
 * history bit for bit; the accumulated grids then differ only by the order of the double
 * additions.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this file.
 *
 * What it restates (reference = OpenGATE/opengate @ 9b07951, paths under /root/reference):
 *   - voxel -> material lookup: core/opengate_core/opengate_lib/GateImageNestedParameterisation.cpp:47-75
 *   - dose scoring + event-wise squares: GateDoseActor.cpp:144-243, 359-397
 *   - dose-grid index (itk TransformPhysicalPointToIndex on the attached image):
 *     GateHelpersImage.txx:20-58
 *   - primary generation: GateGenericSource.cpp:312-664, GateSPSEneDistribution.cpp:16-157,
 *     GateSPSVoxelsPosDistribution.cpp:29-72, GateSingleParticleSource.cpp:89-137
 *     (position, then direction, then energy)
 *   - phase-space entering/exiting rows: GatePhaseSpaceActor.cpp:119-201
 *   - phase-space replay source: GatePhaseSpaceSource.cpp:92-178, sources/phspsources.py:119-258
 *   - track-length estimator: GateTLEDoseActor.cpp:185-300 (mu_en tables: GateMaterialMuHandler.cpp:159-263,
 *     GateMuTables.cpp:47-107, restated in opengate_b200/mu_handler.py and checked against the reference's
 *     own GateMuTables.cpp compiled into oracle/_ref)
 *   - G4SPS position / direction samplers (surface_sphere, cosine law, beam2d): Geant4, not in the tree,
 *     restated from the published algorithms
 *   - photon physics: Geant4 11.4.2 is NOT in the reference tree; the
 *     interaction sampling follows the published algorithms (Klein-Nishina composition-
 *     rejection of the Geant4 Physics Reference Manual "Compton scattering", scattering-
 *     function rejection and form-factor sampling as in the Livermore models, Woodcock
 *     delta tracking).  PARITY UNPINNED against Geant4 for this part: no golden vector of
 *     the reference's transport is available in this container (tests/data submodule absent).
 *     What the tree does pin (tests/test_oracle_physics.py, tests/test_golden_reference.py): total
 *     attenuation = NIST (opengate/data/PhotonAttenuation.py), and the energy-transfer coefficient
 *     mu_tr/rho of this physics = mu_en/rho of the reference's EPDL database within 0.8 %.
 */
#include "transport_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ arithmetic contract */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* natural logarithm: x = m 2^k with m in [sqrt(1/2), sqrt(2)); log(m) = f - f^2/2 + f^3 P(f), f = m - 1,
 * P a degree-8 polynomial fitted for this repository (tools/fit_log_poly.py; < 0.8 ulp on the ranges the
 * transport uses), evaluated by Horner's rule with fused multiply-adds: no division. */
static float o_logf(float x) {
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
  p = fmaf(p, f, c7); p = fmaf(p, f, c6); p = fmaf(p, f, c5); p = fmaf(p, f, c4);
  p = fmaf(p, f, c3); p = fmaf(p, f, c2); p = fmaf(p, f, c1); p = fmaf(p, f, c0);
  float z = f * f;
  float dk = (float)k;
  float y = (f * z) * p;
  y = fmaf(dk, ln2_lo, y);
  y = fmaf(-0.5f, z, y);
  return fmaf(dk, ln2_hi, f + y);
}

/* exp(x) for moderate |x| (used on [-6,0]) */
static float o_expf(float x) {
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

/* sin and cos of 2*pi*t for t in [0,1) */
static void o_sincos2pi(float t, float *s, float *c) {
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
  if (q == 0) { *s = sn; *c = cs; }
  else if (q == 1) { *s = cs; *c = -sn; }
  else if (q == 2) { *s = -sn; *c = -cs; }
  else { *s = -cs; *c = sn; }
}

/* ------------------------------------------------------------------ precision switch
 * Default build (liboracle.so): real_t = float and the functions above: the arithmetic contract the CUDA
 * engine reproduces bit for bit.  -DORACLE_FP64 (liboracle64.so): the SAME algorithm and random-number
 * protocol carried out in double precision with libm's log / exp / sin / cos -- the arithmetic of the
 * reference (Geant4 and itk::Image<double,3> are double throughout).  It exists to show that the
 * single-precision transport of the engine stays inside the acceptance tolerances (tests/test_fp64_acceptance.py);
 * the tables are the same float values in both builds. */
#ifdef ORACLE_FP64
typedef double real_t;
#define R_FMA(a, b, c) fma((a), (b), (c))
#define R_SQRT(x) sqrt(x)
#define R_FLOOR(x) floor(x)
#define R_LOG(x) log(x)
#define R_EXP(x) exp(x)
static void R_SINCOS2PI(double t, double *s, double *c) {
  const double a = 6.283185307179586476925286766559 * t;
  *s = sin(a); *c = cos(a);
}
#else
typedef float real_t;
#define R_FMA(a, b, c) fmaf((a), (b), (c))
#define R_SQRT(x) sqrtf(x)
#define R_FLOOR(x) floorf(x)
#define R_LOG(x) o_logf(x)
#define R_EXP(x) o_expf(x)
#define R_SINCOS2PI(t, s, c) o_sincos2pi((t), (s), (c))
#endif

/* ------------------------------------------------------------------ Philox4x32-10 */

typedef struct {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t out[4];
  int pos;
} rng_t;

static void philox_block(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void rng_init(rng_t *r, uint64_t seed, uint64_t history) {
  r->key[0] = (uint32_t)seed; r->key[1] = (uint32_t)(seed >> 32);
  r->ctr[0] = (uint32_t)history; r->ctr[1] = (uint32_t)(history >> 32);
  r->ctr[2] = 0; r->ctr[3] = 0;
  r->pos = 4;
}

/* Every phase call (primary generation, one world step, one Woodcock call, one interaction)
 * starts on a fresh Philox block: whatever is left of the previous block is discarded.  This is
 * part of the algorithm definition (DESIGN.md "random-number protocol"): it lets the device run
 * the phases of one history on different warps without carrying generator output around. */
static void rng_begin(rng_t *r) { r->pos = 4; }

/* uniform in (0,1): 24 random bits, centred in their bin */
static real_t rng_uniform(rng_t *r) {
  if (r->pos == 4) {
    philox_block(r->key, r->ctr, r->out);
    r->ctr[2] += 1u;
    r->pos = 0;
  }
  uint32_t x = r->out[r->pos++];
  return (real_t)(x >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f;
}

/* ------------------------------------------------------------------ model (real_t copies) */

#define ORACLE_MAX_PHSP 4
#define ORACLE_MAX_DOSE 4
#define DEP_BUF 32
#define WOODCOCK_K 4
#define TLE_STEPS 16   /* voxel chords per call of the device's march phase (each call: fresh Philox block) */
#define TLE_H 4096     /* per-history table of (actor, voxel) sums of a TLE actor with squares */
#define TLE_MAX 3072   /* entries at which the table is written out (counted as an overflow) */
#define TLE_EMPTY (-1)
#define TLE_REC_MAX 64 /* scored flights + point deposits of one history a TLE actor with squares keeps apart: one more
                        * writes the table out first (counted as an overflow, like a full table) */
#define MAX_BLOCKS 4000000u
#define NUDGE 1.0e-3f
#define BIG 3.0e38f
#define MC2 0.51099891f
#define K_PER_MEV 268.17264f /* photon wave number (1/bohr) per MeV */
#define C_MM_PER_NS 299.792458f

typedef struct { real_t m[9]; real_t t[3]; } affine_t; /* y = m x + t */

typedef struct {
  gate_b200_source d;
  affine_t pos;      /* local -> world */
  real_t dir_rot[9];
  real_t cos_tmin, cos_tmax, phi0_turn, dphi_turn;
  real_t sin2_tmin, sin2_tmax, dir_sx, dir_sy;
  real_t mom[3], focus[3];
  float *spec_e, *spec_cdf;   /* as handed to the device */
  float *ht_e, *ht_c, *hp_e, *hp_c;   /* user angular histograms: edges and cumulative probabilities */
  affine_t conf;              /* world -> confining box */
  real_t conf_half[3];
} osource_t;

typedef struct {
  int size[3];
  affine_t v2i;      /* phantom-local point -> continuous index */
  int flags, hit_type;
  double *grid[5];
  /* track-length estimator (GateTLEDoseActor): tle = 1, aligned = the grid is the image grid */
  int tle, aligned;
  real_t tle_thr;
  float *muen;        /* [nmat][ne] mu_en/rho, cm2/g */
} odose_t;

typedef struct {
  real_t half[3];
  affine_t w2b;
  int steps;
  uint64_t capacity;
  oracle_phsp_row *rows;
  uint64_t n_rows, n_dropped;
} ophsp_t;

struct oracle_engine {
  uint64_t seed;
  /* tables */
  int nmat, ne, nq, pair_mode;
  real_t ln_emin, inv_dlne, e_cut, e_min, e_max;
  real_t ln_q2min, inv_dlnq2;
  float *dens, *mu, *cph, *cco, *cra, *maj, *q2n, *ray, *sf;   /* tables: the float values the device holds */
  /* geometry */
  real_t wh[3];
  int world_mat;
  int has_vol, n[3], label_bytes;
  real_t half[3], inv_sp[3], sp[3], hs[3];
  affine_t w2v;
  void *labels;
  /* the homogeneous shell: every voxel outside the box (centre tc, half size th, volume frame; the bounding box of
   * the voxels whose label differs from that of voxel 0) carries the label shell_mat.  A Woodcock call that starts
   * out there first crosses the shell analytically with the optical depth of its first flight. */
  int shell, shell_mat;
  real_t tc[3], th[3];
  /* actors, sources */
  int nsrc; osource_t *src;
  int ndose; odose_t dose[ORACLE_MAX_DOSE];
  int ntle;            /* dose actors that are track-length estimators */
  real_t tle_emin;      /* max energy_min over them */
  real_t tle_thr_max;   /* max threshold over them */
  int tle_sq;          /* one of them scores event-wise squares (a per-history table exists) */
  int tle_rec_cap;     /* records a history may hold (TLE_REC_MAX; tests lower it to reach the overflow path) */
  int tle_fold;        /* test switch (index of the TLE actor + 1): score it by the device's record-fold statement, not the table */
  int nphsp; ophsp_t phsp[ORACLE_MAX_PHSP];
  int nreg; struct { real_t half[3]; affine_t w2b; int mat; } reg[4];   /* homogeneous Box regions in the world */
  gate_b200_stats stats;
  char err[256];
};

typedef struct { int key; real_t e; real_t d; int n; } dep_t;
/* one scored flight (point = 0) or one point deposit (point = 1) of a history, as the device keeps it (transport.cuh) */
typedef struct tle_rec {
  int point;
  real_t s[3], d[3], t_end, t_min, kE, fe; int ie; int lo[3], hi[3];   /* flight: start, direction, length, energy bin, voxel range */
  int key; real_t e, dd;                                              /* point deposit */
} tle_rec;

typedef struct {
  oracle_engine *g;
  rng_t rng;
  real_t p[3], d[3];  /* world frame in WORLD, phantom frame in PHANTOM */
  real_t E, w, path;
  int in_phantom;
  int ie; real_t fe;  /* energy bin + fraction */
  real_t seg;         /* path since last step point inside the phantom */
  int inside_plane[ORACLE_MAX_PHSP];
  int stored;
  uint64_t hist;
  dep_t buf[DEP_BUF]; int nbuf;
  dep_t *ttab; uint16_t *tlist; int nlist;   /* TLE table of this history (thread scratch) */
  int nrec;                                  /* records (scored flights, point deposits) behind the table's content */
  struct tle_rec *recs;                      /* the records themselves (only used by the fold statement) */
  real_t fold_tmin;                          /* fold: chords of the current flight before this t are written out already */
  gate_b200_stats st;
  oracle_trace *trace;
} hist_t;

static void aff_from_d(affine_t *a, const double m12[12]) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) a->m[3 * r + c] = (real_t)m12[4 * r + c];
    a->t[r] = (real_t)m12[4 * r + 3];
  }
}
static void aff_apply(const affine_t *a, const real_t x[3], real_t y[3]) {
  for (int r = 0; r < 3; ++r)
    y[r] = a->m[3 * r] * x[0] + a->m[3 * r + 1] * x[1] + a->m[3 * r + 2] * x[2] + a->t[r];
}
static void rot_apply(const real_t m[9], const real_t x[3], real_t y[3]) {
  for (int r = 0; r < 3; ++r) y[r] = m[3 * r] * x[0] + m[3 * r + 1] * x[1] + m[3 * r + 2] * x[2];
}
/* inverse of a rigid transform y = m x + t :  x = m^T (y - t) */
static void aff_apply_inv(const affine_t *a, const real_t y[3], real_t x[3]) {
  real_t v[3] = {y[0] - a->t[0], y[1] - a->t[1], y[2] - a->t[2]};
  for (int c = 0; c < 3; ++c) x[c] = a->m[c] * v[0] + a->m[3 + c] * v[1] + a->m[6 + c] * v[2];
}
static void rot_apply_inv(const real_t m[9], const real_t y[3], real_t x[3]) {
  for (int c = 0; c < 3; ++c) x[c] = m[c] * y[0] + m[3 + c] * y[1] + m[6 + c] * y[2];
}

/* slab test against an origin-centred box; returns 1 on hit with entry (>=0) and exit
 * distances and the axes of the entry / exit faces (-1 when the start point is inside).
 * A zero-length chord (touching a face from outside) is a miss. */
static int ray_box_ax(const real_t p[3], const real_t d[3], const real_t h[3], real_t *t_in, real_t *t_out,
                      int *a_in, int *a_out) {
  real_t tmin = -BIG, tmax = BIG;
  int ain = -1, aout = -1;
  for (int a = 0; a < 3; ++a) {
    if (d[a] != 0.0f) {
      real_t inv = 1.0f / d[a];
      real_t t1 = (-h[a] - p[a]) * inv;
      real_t t2 = (h[a] - p[a]) * inv;
      if (t1 > t2) { real_t s = t1; t1 = t2; t2 = s; }
      if (t1 > tmin) { tmin = t1; ain = a; }
      if (t2 < tmax) { tmax = t2; aout = a; }
    } else if (p[a] < -h[a] || p[a] > h[a]) {
      return 0;
    }
  }
  real_t t0 = tmin > 0.0f ? tmin : 0.0f;
  if (tmax <= t0) return 0;
  *t_in = t0; *t_out = tmax;
  *a_in = tmin > 0.0f ? ain : -1;
  *a_out = aout;
  return 1;
}
static int ray_box(const real_t p[3], const real_t d[3], const real_t h[3], real_t *t_in, real_t *t_out) {
  int ai, ao;
  return ray_box_ax(p, d, h, t_in, t_out, &ai, &ao);
}

/* Put coordinate `a` just inside (sign -1) or just outside (+1) the face the ray crosses, so
 * that the slab test and the voxel-index test agree about which side the point is on even at
 * grazing incidence (a push along the ray can be smaller than one ulp of the coordinate). */
static void snap_to_face(real_t p[3], const real_t d[3], const real_t h[3], int a, int outward) {
  if (a < 0) return;
  real_t face = outward ? (h[a] * (1.0f + 1.0e-6f) + 1.0e-5f) : (h[a] * (1.0f - 1.0e-6f) - 1.0e-5f);
  if (outward) p[a] = d[a] > 0.0f ? face : -face;
  else p[a] = d[a] > 0.0f ? -face : face;
}

static int inside_box(const real_t p[3], const real_t h[3]) {
  return p[0] > -h[0] && p[0] < h[0] && p[1] > -h[1] && p[1] < h[1] && p[2] > -h[2] && p[2] < h[2];
}

/* point against / coordinate snapped to a face of the inner box of the shell (centre tc, half size th) */
static int inside_tbox(const oracle_engine *g, const real_t p[3]) {
  const real_t q[3] = {p[0] - g->tc[0], p[1] - g->tc[1], p[2] - g->tc[2]};
  return inside_box(q, g->th);
}
static void snap_to_tface(const oracle_engine *g, real_t p[3], const real_t d[3], int a, int outward) {
  real_t q[3] = {p[0] - g->tc[0], p[1] - g->tc[1], p[2] - g->tc[2]};
  snap_to_face(q, d, g->th, a, outward);
  for (int k = 0; k < 3; ++k) p[k] = q[k] + g->tc[k];
}

/* ------------------------------------------------------------------ table access */

static void energy_bin(const oracle_engine *g, real_t E, int *ie, real_t *fe) {
  real_t t = (R_LOG(E) - g->ln_emin) * g->inv_dlne;
  if (t < 0.0f) t = 0.0f;
  int i = (int)t;
  if (i > g->ne - 2) i = g->ne - 2;
  real_t f = t - (real_t)i;
  if (f > 1.0f) f = 1.0f;
  *ie = i; *fe = f;
}
static real_t tab(const float *t, int ne, int m, int ie, real_t fe) {
  const float *r = t + (size_t)m * ne + ie;
  return R_FMA(fe, (real_t)r[1] - (real_t)r[0], (real_t)r[0]);
}

/* ------------------------------------------------------------------ scoring */

static void grid_add(oracle_engine *g, const dep_t *b);

static void flush_buffer(hist_t *h) {
  oracle_engine *g = h->g;
  for (int j = 0; j < h->nbuf; ++j) {
    grid_add(g, &h->buf[j]);
    h->st.flushed_entries++;
  }
  h->nbuf = 0;
}

/* one (actor, voxel) sum of one history -> the fp64 grids */
static void grid_add(oracle_engine *g, const dep_t *b) {
  {
    int a = (int)((uint32_t)b->key >> 28);
    int v = b->key & 0x0fffffff;
    odose_t *D = &g->dose[a];
    double e = (double)b->e, d = (double)b->d;
    /* GateDoseActor.cpp:219-229 (sums) and :359-381 (one square per event and voxel) */
#pragma omp atomic
    D->grid[GATE_B200_GRID_EDEP][v] += e;
    if (D->flags & GATE_B200_SCORE_EDEP_SQ) {
#pragma omp atomic
      D->grid[GATE_B200_GRID_EDEP_SQ][v] += e * e;
    }
    if (D->flags & GATE_B200_SCORE_DOSE) {
#pragma omp atomic
      D->grid[GATE_B200_GRID_DOSE][v] += d;
    }
    if (D->flags & GATE_B200_SCORE_DOSE_SQ) {
#pragma omp atomic
      D->grid[GATE_B200_GRID_DOSE_SQ][v] += d * d;
    }
    if (D->flags & GATE_B200_SCORE_COUNTS) {
#pragma omp atomic
      D->grid[GATE_B200_GRID_COUNTS][v] += (double)b->n;
    }
  }
}

/* ---- per-history table of a TLE actor that scores squares: open addressing, Fibonacci hash */
static void tle_flush(hist_t *h) {
  for (int j = 0; j < h->nlist; ++j) {
    dep_t *b = &h->ttab[h->tlist[j]];
    grid_add(h->g, b);
    b->key = TLE_EMPTY;
    h->st.tle_entries++;
  }
  h->nlist = 0;
  h->nrec = 0;
}
static void tle_table_add(hist_t *h, int key, real_t e, real_t d, int n) {
  uint32_t hh = ((uint32_t)key * 2654435761u) >> 20;   /* 32 - log2(TLE_H) */
  for (;;) {
    dep_t *b = &h->ttab[hh];
    if (b->key == key) { b->e += e; b->d += d; b->n += n; return; }
    if (b->key == TLE_EMPTY) {
      if (h->nlist == TLE_MAX) {   /* table full: write it out, start again (squares of this history split) */
        tle_flush(h);
        h->nrec = 1;               /* the flight / deposit that is being scored goes on in the new table */
        h->st.tle_overflows++;
        hh = ((uint32_t)key * 2654435761u) >> 20;
        b = &h->ttab[hh];
      }
      b->key = key; b->e = e; b->d = d; b->n = n;
      h->tlist[h->nlist++] = (uint16_t)hh;
      return;
    }
    hh = (hh + 1u) & (TLE_H - 1u);
  }
}

/* ---- the same sums WITHOUT a table (the device's statement for one TLE actor on the image grid, transport.cuh
 * tle_fold_*; here behind oracle_set_tle_fold as a CPU cross-check of that algorithm against the table above).
 * A straight flight crosses a voxel once, so two contributions of a history to one voxel come from different
 * flights (or point deposits).  The history keeps its scored flights and point deposits as RECORDS; the sum a voxel
 * holds so far is re-evaluated from them: the chord of an earlier flight through a given voxel has a closed form
 * with exactly the bits the voxel walk produced (entry = the latest entering-face distance of the axes that moved,
 * exit = the earliest leaving-face distance, each face distance the same expression of the integer face index).
 * With the previous sum p and the new one p + e of a voxel, the grids get (p + e) - p and (p + e)^2 - p^2: at the end
 * of the history they hold the sum and the square of the event's total, as GateDoseActor.cpp:359-381 asks. */
static real_t tle_rec_chord(const oracle_engine *g, const tle_rec *r, const int iv[3], real_t *t_in_out) {
  real_t t_in = 0.0f, t_out = r->t_end;
  for (int a = 0; a < 3; ++a) {
    if (iv[a] < r->lo[a] || iv[a] > r->hi[a]) return -1.0f;
    if (r->d[a] == 0.0f) continue;
    const real_t inv = 1.0f / r->d[a];
    const int up = r->d[a] > 0.0f;
    const real_t f_in = R_FMA((real_t)(iv[a] + (up ? 0 : 1)), g->sp[a], -g->half[a]);
    const real_t f_out = R_FMA((real_t)(iv[a] + (up ? 1 : 0)), g->sp[a], -g->half[a]);
    if (iv[a] != (up ? r->lo[a] : r->hi[a])) {   /* the walk entered this voxel layer through f_in */
      const real_t tb = (f_in - r->s[a]) * inv;
      if (tb > t_in) t_in = tb;
    }
    const real_t tb = (f_out - r->s[a]) * inv;
    if (tb < t_out) t_out = tb;
  }
  *t_in_out = t_in;
  return t_out - t_in;
}
/* sum of the earlier contributions of this history to voxel v = (iv) of the one TLE actor; m < 0: label not known yet */
static int tle_fold_prev(hist_t *h, int key, const int iv[3], int m, real_t *pe, real_t *pd) {
  oracle_engine *g = h->g;
  const odose_t *D = &g->dose[g->tle_fold - 1];
  int found = 0;
  real_t se = 0.0f, sd = 0.0f, rho = 0.0f;
  const int nst = h->nrec > 0 ? h->nrec - 1 : 0;   /* stored records: all but the one being scored */
  for (int r = 0; r < nst; ++r) {
    const tle_rec *R = &h->recs[r];
    if (R->point) {
      if (R->key == key) { se += R->e; sd += R->dd; found = 1; }
      continue;
    }
    real_t t_in;
    const real_t L = tle_rec_chord(g, R, iv, &t_in);
    if (!(L > 0.0f) || t_in < R->t_min) continue;
    if (m < 0) {
      const size_t flat = (size_t)iv[0] + (size_t)g->n[0] * ((size_t)iv[1] + (size_t)g->n[1] * (size_t)iv[2]);
      m = g->label_bytes == 1 ? (int)((const uint8_t *)g->labels)[flat] : (int)((const uint16_t *)g->labels)[flat];
    }
    rho = g->dens[m];
    const real_t muen = tab(D->muen, g->ne, m, R->ie, R->fe);
    const real_t e = ((R->kE * muen) * L) * rho;
    se += e; sd += e / rho; found = 1;
  }
  *pe = se; *pd = sd;
  return found;
}
/* one contribution (chord or point deposit) of the history to voxel v; tcur: where the current flight stands */
static void tle_fold_add(hist_t *h, int key, const int iv[3], int m, real_t e, real_t dd, int n, real_t tcur) {
  oracle_engine *g = h->g;
  odose_t *D = &g->dose[g->tle_fold - 1];
  real_t pe, pd;
  if (!tle_fold_prev(h, key, iv, m, &pe, &pd)) {
    if (h->nlist == TLE_MAX) {   /* the table this stands for would be full: forget what was scored so far */
      h->nrec = 1;
      h->fold_tmin = tcur;
      h->nlist = 0;
      h->st.tle_overflows++;
    }
    h->nlist++;
    h->st.tle_entries++;
  }
  const real_t ne = pe + e, nd = pd + dd;
  const int v = key & 0x0fffffff;
  const double e0 = (double)pe, e1 = (double)ne, d0 = (double)pd, d1 = (double)nd;
#pragma omp atomic
  D->grid[GATE_B200_GRID_EDEP][v] += e1 - e0;
  if (D->flags & GATE_B200_SCORE_EDEP_SQ) {
#pragma omp atomic
    D->grid[GATE_B200_GRID_EDEP_SQ][v] += e1 * e1 - e0 * e0;
  }
  if (D->flags & GATE_B200_SCORE_DOSE) {
#pragma omp atomic
    D->grid[GATE_B200_GRID_DOSE][v] += d1 - d0;
  }
  if (D->flags & GATE_B200_SCORE_DOSE_SQ) {
#pragma omp atomic
    D->grid[GATE_B200_GRID_DOSE_SQ][v] += d1 * d1 - d0 * d0;
  }
  if ((D->flags & GATE_B200_SCORE_COUNTS) && n) {
#pragma omp atomic
    D->grid[GATE_B200_GRID_COUNTS][v] += (double)n;
  }
}
/* a scored flight or a point deposit is about to add to the history's sums: one more record behind them */
static void tle_record_begin(hist_t *h) {
  if (!h->g->tle_sq) return;
  if (h->nrec >= h->g->tle_rec_cap) {
    if (h->g->tle_fold) { h->nlist = 0; h->nrec = 0; }
    else tle_flush(h);
    h->st.tle_overflows++;
  }
  h->nrec++;
  h->fold_tmin = 0.0f;
}

/* deposit e (already weighted) at the current phantom-frame point, material m.  e_tle: flight energy
 * of the gamma whose secondaries deposit here (a TLE actor ignores them while that gamma scores by
 * track length, GateTLEDoseActor.cpp:236-243); BIG for a gamma absorbed below the cut (always scored). */
static void deposit(hist_t *h, real_t e, int m, int vix, int viy, int viz, real_t e_tle) {
  oracle_engine *g = h->g;
  h->st.edep_total_MeV += (double)e;
  h->st.deposits++;
  if (h->trace && h->trace->n_dep < ORACLE_TRACE_MAX) {
    h->trace->dep_e[h->trace->n_dep] = e;
    h->trace->n_dep++;
  }
  if (!h->in_phantom) return; /* dose actors are attached to the image volume */
  for (int a = 0; a < g->ndose; ++a) {
    odose_t *D = &g->dose[a];
    if (D->tle && e_tle <= D->tle_thr) continue;
    real_t pos[3] = {h->p[0], h->p[1], h->p[2]};
    /* A Geant4 step inside the nested-parameterised image never leaves its voxel: when the dose grid IS the image
     * grid, any point strictly between the pre- and the post-step point (hit_type random, middle) lies in the voxel
     * of the interaction: no emulation, no random number; "pre" keeps it (a pre-step point ON a voxel face is the
     * reference's own ambiguity) */
    if (D->hit_type != GATE_B200_HIT_POST && !(D->aligned && D->hit_type != GATE_B200_HIT_PRE)) {
      /* GateDoseActor.cpp:144-178: pre/random/middle between the Geant4 pre and post step
       * points; the Geant4 step starts at the voxel entry or at the previous interaction */
      real_t back = h->seg;
      const int vi[3] = {vix, viy, viz};
      for (int ax = 0; ax < 3; ++ax) {
        real_t dd = h->d[ax];
        if (dd > 0.0f) {
          real_t lo = -g->half[ax] + (real_t)vi[ax] * g->sp[ax];
          real_t tb = (h->p[ax] - lo) / dd;
          if (tb < back) back = tb;
        } else if (dd < 0.0f) {
          real_t hi = -g->half[ax] + (real_t)(vi[ax] + 1) * g->sp[ax];
          real_t tb = (h->p[ax] - hi) / dd;
          if (tb < back) back = tb;
        }
      }
      if (back < 0.0f) back = 0.0f;
      real_t fr = 1.0f;
      if (D->hit_type == GATE_B200_HIT_RANDOM) fr = 1.0f - rng_uniform(&h->rng);
      else if (D->hit_type == GATE_B200_HIT_MIDDLE) fr = 0.5f;
      real_t bk = back * fr;
      for (int ax = 0; ax < 3; ++ax) pos[ax] = h->p[ax] - h->d[ax] * bk;
    }
    real_t c[3];
    aff_apply(&D->v2i, pos, c);
    /* itk::ImageBase::TransformPhysicalPointToIndex: round half up, then region test */
    real_t fx = R_FLOOR(c[0] + 0.5f), fy = R_FLOOR(c[1] + 0.5f), fz = R_FLOOR(c[2] + 0.5f);
    if (fx < 0.0f || fy < 0.0f || fz < 0.0f || fx >= (real_t)D->size[0] || fy >= (real_t)D->size[1] ||
        fz >= (real_t)D->size[2])
      continue;
    int v = (int)fx + D->size[0] * ((int)fy + D->size[1] * (int)fz);
    int key = (int)(((uint32_t)a << 28) | (uint32_t)v);
    real_t dd = e / g->dens[m];
    if (D->tle && (D->flags & (GATE_B200_SCORE_EDEP_SQ | GATE_B200_SCORE_DOSE_SQ))) {
      tle_record_begin(h);
      if (g->tle_fold) {
        const int iv[3] = {v % D->size[0], (v / D->size[0]) % D->size[1], v / (D->size[0] * D->size[1])};
        tle_fold_add(h, key, iv, -1, e, dd, 1, 0.0f);
        tle_rec *R = &h->recs[h->nrec - 1];
        R->point = 1; R->key = key; R->e = e; R->dd = dd;
      } else {
        tle_table_add(h, key, e, dd, 1);
      }
      continue;
    }
    int j;
    for (j = 0; j < h->nbuf; ++j)
      if (h->buf[j].key == key) break;
    if (j == h->nbuf) {
      if (h->nbuf == DEP_BUF) { /* coalescing buffer full: write it out, start again */
        flush_buffer(h);
        h->st.buffer_overflows++;
        j = 0;
      }
      h->buf[j].key = key; h->buf[j].e = 0.0f; h->buf[j].d = 0.0f; h->buf[j].n = 0;
      h->nbuf++;
    }
    h->buf[j].e += e; h->buf[j].d += dd; h->buf[j].n += 1;
  }
}

/* ------------------------------------------------------------------ direction change */

static void rotate_dir(real_t d[3], real_t cost, real_t sint, real_t sphi, real_t cphi) {
  real_t u = d[0], v = d[1], w = d[2];
  real_t nu, nv, nw;
  real_t aw = w < 0.0f ? -w : w;
  if (aw < 0.99999f) {
    real_t sz = R_SQRT(1.0f - w * w);
    real_t a = sint / sz;
    nu = u * cost + a * (u * w * cphi - v * sphi);
    nv = v * cost + a * (v * w * cphi + u * sphi);
    nw = w * cost - sint * sz * cphi;
  } else {
    real_t sg = w < 0.0f ? -1.0f : 1.0f;
    nu = sint * cphi; nv = sint * sphi; nw = sg * cost;
  }
  real_t inv = 1.0f / R_SQRT(nu * nu + nv * nv + nw * nw);
  d[0] = nu * inv; d[1] = nv * inv; d[2] = nw * inv;
}

/* ------------------------------------------------------------------ momentum-transfer tables */

static int q2_segment(const oracle_engine *g, real_t q2) {
  /* largest j with q2n[j] <= q2, -1 if below the first node, nq-1 if above the last */
  if (q2 < g->q2n[0]) return -1;
  if (q2 >= g->q2n[g->nq - 1]) return g->nq - 1;
  real_t t = (R_LOG(q2) - g->ln_q2min) * g->inv_dlnq2;
  int j = (int)t;
  if (j < 0) j = 0;
  if (j > g->nq - 2) j = g->nq - 2;
  while (j < g->nq - 2 && g->q2n[j + 1] <= q2) ++j;
  while (j > 0 && g->q2n[j] > q2) --j;
  return j;
}
static real_t interp_q2(const oracle_engine *g, const float *row, real_t q2) {
  int j = q2_segment(g, q2);
  if (j < 0) return row[0] * (q2 / g->q2n[0]);
  if (j >= g->nq - 1) return row[g->nq - 1];
  real_t f = (q2 - g->q2n[j]) / (g->q2n[j + 1] - g->q2n[j]);
  return row[j] + f * (row[j + 1] - row[j]);
}

/* ------------------------------------------------------------------ interactions */

static void update_energy(hist_t *h) { energy_bin(h->g, h->E, &h->ie, &h->fe); }

/* Process selection.  v is uniform in (0,1): inside the volume it is the Woodcock acceptance
 * variate rescaled, v = (u * majorant) / mu_total(material) (conditional on acceptance this is
 * uniform, so no extra random number is needed); in the world it is the second uniform of the
 * flight block.  Returns 0 phot, 1 compt, 2 rayl, 3 conv. */
static int select_process(hist_t *h, int m, real_t v) {
  oracle_engine *g = h->g;
  real_t pph = tab(g->cph, g->ne, m, h->ie, h->fe);
  real_t pco = tab(g->cco, g->ne, m, h->ie, h->fe);
  real_t pra = tab(g->cra, g->ne, m, h->ie, h->fe);
  h->st.steps++;
  if (v < pph) return 0;
  if (v < pco) return 1;
  if (v < pra) return 2;
  return 3;
}

/* One real interaction of kind proc in material m; returns 1 if the photon survives. */
static int interact(hist_t *h, int proc, int m, int vix, int viy, int viz, real_t pend_p[3], real_t pend_d[3],
                    int *pend) {
  oracle_engine *g = h->g;
  if (proc == 0) { /* photoelectric: everything stays here (kerma, fluorescence not tracked) */
    h->st.n_phot++;
    rng_begin(&h->rng);
    deposit(h, h->E * h->w, m, vix, viy, viz, h->E);
    return 0;
  }
  if (proc == 1) { /* Compton: Klein-Nishina composition-rejection x scattering function */
    h->st.n_compt++;
    real_t E = h->E;
    real_t k = E / MC2;
    real_t eps0 = 1.0f / (1.0f + 2.0f * k);
    real_t eps0sq = eps0 * eps0;
    real_t alpha1 = -R_LOG(eps0);
    real_t alpha2 = alpha1 + 0.5f * (1.0f - eps0sq);
    real_t kq = K_PER_MEV * E;
    real_t two_kq2 = 2.0f * kq * kq;
    const float *sfrow = g->sf + (size_t)m * g->nq;
    real_t eps = 1.0f, onecost = 0.0f, sint2 = 0.0f;
    for (int it = 0; it < 100000; ++it) {
      /* one trial = one call of the device's Compton phase: it owns two Philox blocks, the
       * first feeds the trial (4 uniforms, all drawn), the second the final state */
      uint32_t b0 = h->rng.ctr[2];
      rng_begin(&h->rng);
      real_t u1 = rng_uniform(&h->rng), u2 = rng_uniform(&h->rng), u3 = rng_uniform(&h->rng);
      real_t u4 = rng_uniform(&h->rng);
      real_t epssq;
      int ok = 1;
      if (alpha1 > alpha2 * u1) {
        eps = R_EXP(-alpha1 * u2);
        epssq = eps * eps;
      } else {
        epssq = eps0sq + (1.0f - eps0sq) * u2;
        eps = R_SQRT(epssq);
      }
      onecost = (1.0f - eps) / (eps * k);
      sint2 = onecost * (2.0f - onecost);
      real_t grej = 1.0f - eps * sint2 / (1.0f + epssq);
      if (grej < u3) ok = 0;
      if (ok) {
        real_t sfv = interp_q2(g, sfrow, two_kq2 * onecost);
        if (u4 > sfv) ok = 0;
      }
      if (ok) break;
      h->rng.ctr[2] = b0 + 2u; /* rejected: the rest of the call's blocks is discarded */
      h->rng.pos = 4;
    }
    real_t cost = 1.0f - onecost;
    real_t sint = R_SQRT(sint2 > 0.0f ? sint2 : 0.0f);
    real_t sphi, cphi;
    R_SINCOS2PI(rng_uniform(&h->rng), &sphi, &cphi);
    real_t E1 = eps * E;
    deposit(h, (E - E1) * h->w, m, vix, viy, viz, E);
    rotate_dir(h->d, cost, sint, sphi, cphi);
    h->E = E1;
    h->seg = 0.0f;
    if (E1 < g->e_cut || (h->in_phantom && E1 <= g->tle_emin)) {
      deposit(h, E1 * h->w, m, vix, viy, viz, BIG);
      return 0;
    }
    update_energy(h);
    return 1;
  }
  if (proc == 2) { /* Rayleigh: q2 from the cumulative squared form factor, (1+cos^2)/2 test */
    h->st.n_rayl++;
    rng_begin(&h->rng);
    real_t kq = K_PER_MEV * h->E;
    real_t kq2 = kq * kq;
    const float *A = g->ray + (size_t)m * g->nq;
    real_t q2max = 4.0f * kq2;
    int jmax = q2_segment(g, q2max);
    real_t amax = interp_q2(g, A, q2max);
    real_t cost = 1.0f;
    for (int it = 0; it < 1000; ++it) {
      real_t target = rng_uniform(&h->rng) * amax;
      real_t q2;
      if (target < A[0]) {
        q2 = g->q2n[0] * (target / A[0]);
      } else {
        int lo = 0, hi = jmax < g->nq - 1 ? jmax + 1 : g->nq - 1; /* A[lo] <= target < A[hi] */
        if (target >= A[hi]) lo = hi > 0 ? hi - 1 : 0;
        else
          while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (A[mid] <= target) lo = mid; else hi = mid;
          }
        real_t den = A[lo + 1] - A[lo];
        real_t f = den > 0.0f ? (target - A[lo]) / den : 0.0f;
        if (f > 1.0f) f = 1.0f;
        q2 = g->q2n[lo] + f * (g->q2n[lo + 1] - g->q2n[lo]);
      }
      if (q2 > q2max) q2 = q2max;
      cost = 1.0f - q2 / (2.0f * kq2);
      if (cost < -1.0f) cost = -1.0f;
      real_t u2 = rng_uniform(&h->rng);
      if (0.5f * (1.0f + cost * cost) >= u2) break;
    }
    real_t sint = R_SQRT(1.0f - cost * cost);
    real_t sphi, cphi;
    R_SINCOS2PI(rng_uniform(&h->rng), &sphi, &cphi);
    rotate_dir(h->d, cost, sint, sphi, cphi);
    h->seg = 0.0f;
    return 1;
  }
  /* gamma conversion */
  h->st.n_conv++;
  rng_begin(&h->rng);
  if (g->pair_mode == 0) {
    /* production cuts >> range with apply_cuts: e-, e+ (and its 2 mc2) stay on the spot */
    deposit(h, h->E * h->w, m, vix, viy, viz, h->E);
    return 0;
  }
  deposit(h, (h->E - 2.0f * MC2) * h->w, m, vix, viy, viz, h->E);
  {
    real_t cost = 2.0f * rng_uniform(&h->rng) - 1.0f;
    real_t sint = R_SQRT(1.0f - cost * cost);
    real_t sphi, cphi;
    R_SINCOS2PI(rng_uniform(&h->rng), &sphi, &cphi);
    h->d[0] = sint * cphi; h->d[1] = sint * sphi; h->d[2] = cost;
    h->E = MC2;
    h->seg = 0.0f;
    update_energy(h);
    for (int a = 0; a < 3; ++a) { pend_p[a] = h->p[a]; pend_d[a] = -h->d[a]; }
    *pend = h->in_phantom ? 1 : 2; /* frame the stored point/direction is expressed in */
    h->st.tracks += 2;
  }
  return 1;
}

/* ------------------------------------------------------------------ primaries */

static real_t gauss(rng_t *r) {
  real_t u1 = rng_uniform(r), u2 = rng_uniform(r);
  real_t s, c;
  R_SINCOS2PI(u2, &s, &c);
  return R_SQRT(-2.0f * R_LOG(u1)) * c;
}

static int lower_bound_f(const float *a, int n, real_t p) { /* first i with a[i] >= p */
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < p) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static void generate_primary(hist_t *h, const osource_t *s) {
  rng_t *r = &h->rng;
  const gate_b200_source *d = &s->d;
  real_t l[3] = {0.0f, 0.0f, 0.0f};
  real_t ed[3] = {0.0f, 0.0f, 1.0f}, eE = 0.0f, ew = 1.0f;   /* phase-space entry */
  /* position.confine (GateGenericSource.cpp:223-227 -> G4SPSPosDistribution::GenerateOne with a confining volume):
   * draw again until the point lies in the confining box, at most 100000 times */
  for (int tries = 0;; ++tries) {
  l[0] = 0.0f; l[1] = 0.0f; l[2] = 0.0f;
  switch (d->pos_type) {
  case GATE_B200_POS_PHSP: {
    /* GatePhaseSpaceSource::GenerateOnePrimary (GatePhaseSpaceSource.cpp:137-166); entries are read
     * sequentially and wrap at the end of the file (sources/phspsources.py:119-146) */
    const uint64_t en = (d->phsp_entry_start % d->phsp_n_entries + h->hist) % d->phsp_n_entries;
    l[0] = d->phsp_pos_x[en]; l[1] = d->phsp_pos_y[en]; l[2] = d->phsp_pos_z[en];
    if (d->dir_type == GATE_B200_DIR_PHSP) { ed[0] = d->phsp_dir_x[en]; ed[1] = d->phsp_dir_y[en]; ed[2] = d->phsp_dir_z[en]; }
    eE = d->phsp_energy[en];
    if (d->phsp_weight) ew = d->phsp_weight[en];
  } break;
  case GATE_B200_POS_POINT: break;
  case GATE_B200_POS_BOX:
    for (int a = 0; a < 3; ++a) l[a] = (2.0f * rng_uniform(r) - 1.0f) * (real_t)d->half_size[a];
    break;
  case GATE_B200_POS_SPHERE: {
    real_t R = (real_t)d->radius;
    for (;;) {
      for (int a = 0; a < 3; ++a) l[a] = (2.0f * rng_uniform(r) - 1.0f) * R;
      if (l[0] * l[0] + l[1] * l[1] + l[2] * l[2] <= R * R) break;
    }
  } break;
  case GATE_B200_POS_DISC:
  case GATE_B200_POS_CYLINDER: {
    real_t R = (real_t)d->radius;
    if (R > 0.0f)
      for (;;) {
        l[0] = (2.0f * rng_uniform(r) - 1.0f) * R;
        l[1] = (2.0f * rng_uniform(r) - 1.0f) * R;
        if (l[0] * l[0] + l[1] * l[1] <= R * R) break;
      }
    if (d->pos_type == GATE_B200_POS_DISC) {
      if (d->sigma_x > 0.0) l[0] += gauss(r) * (real_t)d->sigma_x;
      if (d->sigma_y > 0.0) l[1] += gauss(r) * (real_t)d->sigma_y;
    } else {
      l[2] = (2.0f * rng_uniform(r) - 1.0f) * (real_t)d->half_size[2];
    }
  } break;
  case GATE_B200_POS_SURFACE_SPHERE: {
    /* G4SPSPosDistribution::GeneratePointsOnSurface, Sphere: theta = acos(1 - 2u), phi = 2 pi u' */
    real_t R = (real_t)d->radius;
    real_t cost = 1.0f - 2.0f * rng_uniform(r);
    real_t s2 = 1.0f - cost * cost;
    real_t sint = R_SQRT(s2 > 0.0f ? s2 : 0.0f);
    real_t sp, cp;
    R_SINCOS2PI(rng_uniform(r), &sp, &cp);
    l[0] = R * sint * cp; l[1] = R * sint * sp; l[2] = R * cost;
  } break;
  case GATE_B200_POS_VOXELS: {
    int nx = d->vox_size[0], ny = d->vox_size[1], nz = d->vox_size[2];
    int i, j, k, tries;
    tries = 0;
    do { i = lower_bound_f(d->cdf_z, nz, rng_uniform(r)); } while (i >= nz && ++tries < 1000);
    if (i >= nz) i = nz - 1;
    tries = 0;
    do { j = lower_bound_f(d->cdf_y + (size_t)i * ny, ny, rng_uniform(r)); } while (j >= ny && ++tries < 1000);
    if (j >= ny) j = ny - 1;
    tries = 0;
    do { k = lower_bound_f(d->cdf_x + ((size_t)i * ny + j) * nx, nx, rng_uniform(r)); } while (k >= nx && ++tries < 1000);
    if (k >= nx) k = nx - 1;
    l[0] = (real_t)d->vox_origin[0] + (real_t)k * (real_t)d->vox_spacing[0];
    l[1] = (real_t)d->vox_origin[1] + (real_t)j * (real_t)d->vox_spacing[1];
    l[2] = (real_t)d->vox_origin[2] + (real_t)i * (real_t)d->vox_spacing[2];
    l[0] += (rng_uniform(r) - 0.5f) * (real_t)d->vox_spacing[0];
    l[1] += (rng_uniform(r) - 0.5f) * (real_t)d->vox_spacing[1];
    l[2] += (rng_uniform(r) - 0.5f) * (real_t)d->vox_spacing[2];
  } break;
  }
  aff_apply(&s->pos, l, h->p);
  if (!d->confine || tries >= 99999) break;
  real_t qc[3];
  aff_apply(&s->conf, h->p, qc);
  if (inside_box(qc, s->conf_half)) break;
  }

  switch (d->dir_type) {
  case GATE_B200_DIR_ISO: {
    /* G4SPSAngDistribution::GenerateIsotropicFlux: momentum = -(sin t cos p, sin t sin p, cos t) */
    real_t cost = s->cos_tmin - rng_uniform(r) * (s->cos_tmin - s->cos_tmax);
    real_t s2 = 1.0f - cost * cost;
    real_t sint = R_SQRT(s2 > 0.0f ? s2 : 0.0f);
    real_t turn = s->phi0_turn + s->dphi_turn * rng_uniform(r);
    turn = turn - R_FLOOR(turn);
    real_t sp, cp;
    R_SINCOS2PI(turn, &sp, &cp);
    real_t v[3] = {-sint * cp, -sint * sp, -cost};
    if (d->dir_relative) rot_apply(s->dir_rot, v, h->d);
    else { h->d[0] = v[0]; h->d[1] = v[1]; h->d[2] = v[2]; }
  } break;
  case GATE_B200_DIR_MOMENTUM:
    h->d[0] = s->mom[0]; h->d[1] = s->mom[1]; h->d[2] = s->mom[2];
    break;
  case GATE_B200_DIR_FOCUSED: {
    real_t v[3] = {s->focus[0] - h->p[0], s->focus[1] - h->p[1], s->focus[2] - h->p[2]};
    real_t inv = 1.0f / R_SQRT(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    h->d[0] = v[0] * inv; h->d[1] = v[1] * inv; h->d[2] = v[2] * inv;
  } break;
  case GATE_B200_DIR_COS: {
    /* G4SPSAngDistribution::GenerateCosineLawFlux: sin^2(theta) uniform between sin^2(min) and sin^2(max) */
    real_t st2 = rng_uniform(r) * (s->sin2_tmax - s->sin2_tmin) + s->sin2_tmin;
    real_t sint = R_SQRT(st2 > 0.0f ? st2 : 0.0f);
    real_t c2 = 1.0f - st2;
    real_t cost = R_SQRT(c2 > 0.0f ? c2 : 0.0f);
    real_t turn = s->phi0_turn + s->dphi_turn * rng_uniform(r);
    turn = turn - R_FLOOR(turn);
    real_t sp, cp;
    R_SINCOS2PI(turn, &sp, &cp);
    real_t v[3] = {-sint * cp, -sint * sp, -cost};
    if (d->dir_relative) rot_apply(s->dir_rot, v, h->d);
    else { h->d[0] = v[0]; h->d[1] = v[1]; h->d[2] = v[2]; }
  } break;
  case GATE_B200_DIR_BEAM2D: {
    /* G4SPSAngDistribution::GenerateBeamFlux (beam2d): px, py ~ N(0, sigma), theta = |(px, py)|,
     * phi = atan2(py, px), momentum = -(sin t cos p, sin t sin p, cos t) with cos p = px / theta */
    real_t px = gauss(r) * s->dir_sx;
    real_t py = gauss(r) * s->dir_sy;
    real_t th = R_SQRT(px * px + py * py);
    real_t v[3] = {0.0f, 0.0f, -1.0f};
    if (th != 0.0f) {
      real_t turn = th * 0.15915494309189535f;
      turn = turn - R_FLOOR(turn);
      real_t st, ct;
      R_SINCOS2PI(turn, &st, &ct);
      real_t k = st / th;
      v[0] = -(k * px); v[1] = -(k * py); v[2] = -ct;
    }
    if (d->dir_relative) rot_apply(s->dir_rot, v, h->d);
    else { h->d[0] = v[0]; h->d[1] = v[1]; h->d[2] = v[2]; }
  } break;
  case GATE_B200_DIR_HISTOGRAM: {
    /* GateGenericSource.cpp:435-464 -> G4SPSAngDistribution::GenerateUserDefFlux with theta and phi histograms: the
     * angle is the inverse of the piecewise-linear cumulative at a uniform variate (G4PhysicsFreeVector::GetEnergy),
     * momentum = -(sin t cos p, sin t sin p, cos t) */
    real_t ang[2];
    for (int k = 0; k < 2; ++k) {
      const float *ce = k == 0 ? s->ht_e : s->hp_e;
      const float *cc = k == 0 ? s->ht_c : s->hp_c;
      const int nb = k == 0 ? d->n_dir_theta : d->n_dir_phi;
      const real_t u = rng_uniform(r);
      int b = 0;
      while (b < nb - 1 && u >= cc[b + 1]) ++b;   /* first bin whose upper cumulative exceeds u */
      const real_t c0 = cc[b], c1 = cc[b + 1], e0 = ce[b], e1 = ce[b + 1];
      const real_t den = c1 - c0;
      real_t f = den > 0.0f ? (u - c0) / den : 0.0f;
      if (f > 1.0f) f = 1.0f;
      ang[k] = e0 + f * (e1 - e0);
    }
    real_t tt = ang[0] * 0.15915494309189535f, tp = ang[1] * 0.15915494309189535f;
    tt = tt - R_FLOOR(tt); tp = tp - R_FLOOR(tp);
    real_t st, ct, sp, cp;
    R_SINCOS2PI(tt, &st, &ct);
    R_SINCOS2PI(tp, &sp, &cp);
    real_t v[3] = {-st * cp, -st * sp, -ct};
    if (d->dir_relative) rot_apply(s->dir_rot, v, h->d);
    else { h->d[0] = v[0]; h->d[1] = v[1]; h->d[2] = v[2]; }
  } break;
  case GATE_B200_DIR_PHSP: {
    /* direction / direction.mag(), then the rotation of the attached volume (GatePhaseSpaceSource.cpp:158-162) */
    real_t inv = 1.0f / R_SQRT(ed[0] * ed[0] + ed[1] * ed[1] + ed[2] * ed[2]);
    real_t v[3] = {ed[0] * inv, ed[1] * inv, ed[2] * inv};
    if (d->dir_relative) rot_apply(s->dir_rot, v, h->d);
    else { h->d[0] = v[0]; h->d[1] = v[1]; h->d[2] = v[2]; }
  } break;
  }

  real_t E = 0.0f;
  switch (d->ene_type) {
  case GATE_B200_ENE_PHSP: E = eE; break;
  case GATE_B200_ENE_MONO: E = (real_t)d->mono_MeV; break;
  case GATE_B200_ENE_GAUSS: E = (real_t)d->mono_MeV + gauss(r) * (real_t)d->sigma_MeV; break;
  case GATE_B200_ENE_RANGE:
    E = (real_t)d->e_min_MeV + rng_uniform(r) * ((real_t)d->e_max_MeV - (real_t)d->e_min_MeV);
    break;
  case GATE_B200_ENE_SPECTRUM_DISCRETE: {
    real_t p = rng_uniform(r);
    int i = 0;
    while (i < d->n_spectrum - 1 && p >= s->spec_cdf[i]) ++i; /* IndexForProbability */
    E = s->spec_e[i];
  } break;
  case GATE_B200_ENE_SPECTRUM_HISTOGRAM: {
    real_t p = rng_uniform(r);
    int i = 0;
    while (i < d->n_spectrum - 1 && p >= s->spec_cdf[i]) ++i;
    E = s->spec_e[i] + rng_uniform(r) * (s->spec_e[i + 1] - s->spec_e[i]);
  } break;
  case GATE_B200_ENE_SPECTRUM_HISTOGRAM_LINEAR: {
    /* GateSPSEneDistribution.cpp:124-147 */
    real_t p0 = rng_uniform(r);
    int i = 0;
    while (i < d->n_spectrum - 2 && p0 >= s->spec_cdf[i]) ++i;
    real_t a = (s->spec_e[i] + s->spec_e[i + 1]) * 0.5f;
    real_t b = (s->spec_e[i + 1] + s->spec_e[i + 2]) * 0.5f;
    real_t dd = s->spec_cdf[i + 1] - s->spec_cdf[i];
    real_t ad = dd < 0.0f ? -dd : dd;
    if (ad < 1.1920929e-7f) {
      E = a + rng_uniform(r) * (b - a);
    } else {
      real_t alpha = dd / (b - a);
      real_t beta = s->spec_cdf[i] - alpha * a;
      real_t norm = 0.5f * alpha * (b * b - a * a) + beta * (b - a);
      real_t p = rng_uniform(r);
      real_t sq = R_SQRT((alpha * a + beta) * (alpha * a + beta) + 2.0f * alpha * norm * p);
      real_t x = (-beta + sq) / alpha;
      E = ((x - a) * (x - b) <= 0.0f) ? x : (-beta - sq) / alpha;
    }
  } break;
  case GATE_B200_ENE_CDF: {
    /* GateSPSEneDistribution.cpp:38-49 */
    real_t x = rng_uniform(r);
    int idx = lower_bound_f(s->spec_cdf, d->n_spectrum, x) - 1;
    if (idx < 0) idx = 0;
    if (idx > d->n_spectrum - 2) idx = d->n_spectrum - 2;
    real_t ratio = (x - s->spec_cdf[idx]) / (s->spec_cdf[idx + 1] - s->spec_cdf[idx]);
    E = s->spec_e[idx] + ratio * (s->spec_e[idx + 1] - s->spec_e[idx]);
  } break;
  }
  h->E = E;
  h->w = d->pos_type == GATE_B200_POS_PHSP ? ew * (real_t)d->weight : (real_t)d->weight;
}

/* ------------------------------------------------------------------ phase space */

static void phsp_store(hist_t *h, int k, int flag) {
  ophsp_t *P = &h->g->phsp[k];
  h->stored = 1;
  uint64_t idx;
#pragma omp atomic capture
  idx = P->n_rows++;
  if (idx >= P->capacity) {
#pragma omp atomic
    P->n_dropped++;
    return;
  }
  oracle_phsp_row *r = &P->rows[idx];
  r->kinetic_energy = h->E;
  for (int a = 0; a < 3; ++a) { r->position[a] = h->p[a]; r->direction[a] = h->d[a]; }
  r->weight = h->w;
  r->time_ns = h->path / C_MM_PER_NS;
  r->event_id = (int64_t)h->hist;
  r->flags = flag;
}

/* ------------------------------------------------------------------ track-length estimator */

/* Score the straight flight that ends at h->p (volume frame): it started h->seg before and, when the
 * gamma is leaving the volume, continues for `to` up to the exit face.  One Geant4 step of the
 * reference = the chord of one image voxel (GateImageNestedParameterisation); each chord deposits
 * edep = w * 0.1 * E * mu_en/rho * L * rho at a point chosen by hit_type and dose = edep / rho
 * (GateTLEDoseActor::ScoreTLEDepositStep, GateTLEDoseActor.cpp:252-300).  The voxel walk is an
 * Amanatides-Woo traversal whose face distances are always computed from the integer voxel index.
 * The device runs the walk in calls of TLE_STEPS iterations, each on a fresh Philox block. */
static void tle_march(hist_t *h, real_t to) {
  oracle_engine *g = h->g;
  if (!g->ntle || !(h->E <= g->tle_thr_max)) return;
  const real_t t_end = h->seg + to;
  real_t s[3], inv_d[3];
  int i[3];
  for (int a = 0; a < 3; ++a) {
    s[a] = R_FMA(-h->d[a], h->seg, h->p[a]);
    inv_d[a] = h->d[a] != 0.0f ? 1.0f / h->d[a] : 0.0f;
    int v = (int)R_FLOOR(R_FMA(s[a], g->inv_sp[a], g->hs[a]));
    if (v < 0) v = 0;
    if (v > g->n[a] - 1) v = g->n[a] - 1;
    i[a] = v;
  }
  const real_t kE = (0.1f * h->E) * h->w;
  real_t tcur = 0.0f;
  const int i0[3] = {i[0], i[1], i[2]};
  tle_record_begin(h);
  for (int it = 0;; ++it) {
    if (it % TLE_STEPS == 0) rng_begin(&h->rng);
    real_t tn = BIG;
    int ax = -1;
    for (int a = 0; a < 3; ++a) {
      real_t face;
      if (h->d[a] > 0.0f) face = R_FMA((real_t)(i[a] + 1), g->sp[a], -g->half[a]);
      else if (h->d[a] < 0.0f) face = R_FMA((real_t)i[a], g->sp[a], -g->half[a]);
      else continue;
      real_t tb = (face - s[a]) * inv_d[a];
      if (tb < tn) { tn = tb; ax = a; }
    }
    const int last = !(tn < t_end);
    const real_t tnext = last ? t_end : tn;
    const real_t L = tnext - tcur;
    if (L > 0.0f) {
      const size_t flat = (size_t)i[0] + (size_t)g->n[0] * ((size_t)i[1] + (size_t)g->n[1] * (size_t)i[2]);
      const int m = g->label_bytes == 1 ? (int)((const uint8_t *)g->labels)[flat] : (int)((const uint16_t *)g->labels)[flat];
      const real_t rho = g->dens[m];
      h->st.tle_steps++;
      for (int a = 0; a < g->ndose; ++a) {
        odose_t *D = &g->dose[a];
        if (!D->tle || !(h->E <= D->tle_thr)) continue;
        const real_t muen = tab(D->muen, g->ne, m, h->ie, h->fe);
        const real_t e = ((kE * muen) * L) * rho;
        const real_t dd = e / rho;
        int v;
        if (D->aligned) {
          v = (int)flat;
        } else {
          real_t fr = 1.0f;
          if (D->hit_type == GATE_B200_HIT_RANDOM) fr = rng_uniform(&h->rng);
          else if (D->hit_type == GATE_B200_HIT_MIDDLE) fr = 0.5f;
          else if (D->hit_type == GATE_B200_HIT_PRE) fr = 0.0f;
          const real_t tp = R_FMA(fr, L, tcur);
          real_t pos[3] = {R_FMA(h->d[0], tp, s[0]), R_FMA(h->d[1], tp, s[1]), R_FMA(h->d[2], tp, s[2])};
          real_t c[3];
          aff_apply(&D->v2i, pos, c);
          real_t fx = R_FLOOR(c[0] + 0.5f), fy = R_FLOOR(c[1] + 0.5f), fz = R_FLOOR(c[2] + 0.5f);
          if (fx < 0.0f || fy < 0.0f || fz < 0.0f || fx >= (real_t)D->size[0] || fy >= (real_t)D->size[1] ||
              fz >= (real_t)D->size[2])
            continue;
          v = (int)fx + D->size[0] * ((int)fy + D->size[1] * (int)fz);
        }
        const int key = (int)(((uint32_t)a << 28) | (uint32_t)v);
        if (D->flags & (GATE_B200_SCORE_EDEP_SQ | GATE_B200_SCORE_DOSE_SQ)) {
          if (g->tle_fold) tle_fold_add(h, key, i, m, e, dd, 0, tcur);
          else tle_table_add(h, key, e, dd, 0);
        } else {
          dep_t b; b.key = key; b.e = e; b.d = dd; b.n = 0;
          grid_add(g, &b);
          h->st.tle_entries++;
        }
      }
    }
    if (last) break;
    if (tnext > tcur) tcur = tnext;
    const int ni = i[ax] + (h->d[ax] > 0.0f ? 1 : -1);
    if (ni < 0 || ni >= g->n[ax]) break;
    i[ax] = ni;
  }
  if (g->tle_fold && g->tle_sq) {
    /* the flight becomes a record: start, direction, length, energy, the range of voxels it walked */
    tle_rec *R = &h->recs[h->nrec - 1];
    R->point = 0;
    for (int a = 0; a < 3; ++a) {
      R->s[a] = s[a]; R->d[a] = h->d[a];
      R->lo[a] = i0[a] < i[a] ? i0[a] : i[a];
      R->hi[a] = i0[a] < i[a] ? i[a] : i0[a];
    }
    R->t_end = t_end; R->t_min = h->fold_tmin; R->kE = kE; R->ie = h->ie; R->fe = h->fe;
  }
}

/* ------------------------------------------------------------------ one history */

static void to_phantom(hist_t *h) {
  real_t pl[3], dl[3];
  aff_apply(&h->g->w2v, h->p, pl);
  rot_apply(h->g->w2v.m, h->d, dl);
  for (int a = 0; a < 3; ++a) { h->p[a] = pl[a]; h->d[a] = dl[a]; }
}
static void to_world(hist_t *h) {
  real_t pw[3], dw[3];
  aff_apply_inv(&h->g->w2v, h->p, pw);
  rot_apply_inv(h->g->w2v.m, h->d, dw);
  for (int a = 0; a < 3; ++a) { h->p[a] = pw[a]; h->d[a] = dw[a]; }
}

static void run_photon(hist_t *h, real_t pend_p[3], real_t pend_d[3], int *pend) {
  oracle_engine *g = h->g;
  int alive = 1;
  if (h->E < g->e_cut) { /* below the table: absorbed where it starts */
    h->st.edep_total_MeV += (double)(h->E * h->w);
    return;
  }
  update_energy(h);
  for (int iter = 0; alive; ++iter) {
    if (h->rng.ctr[2] > MAX_BLOCKS) break; /* runaway guard: Philox blocks one history may consume */
    if (h->trace) h->trace->n_iter++;
#ifdef ORACLE_DEBUG
    if (h->trace && h->trace->n_iter < 40)
      fprintf(stderr, "it %d reg %d p %.6f %.6f %.6f d %.6f %.6f %.6f E %.5f\n", iter, h->in_phantom,
              h->p[0], h->p[1], h->p[2], h->d[0], h->d[1], h->d[2], h->E);
#endif
    if (!h->in_phantom) {
      real_t t_exit, t0;
      rng_begin(&h->rng);
      if (!inside_box(h->p, g->wh) || !ray_box(h->p, h->d, g->wh, &t0, &t_exit)) {
        h->st.escaped++;
        return;
      }
      int ev = 0; /* 0 interact, 1 leave world, 2 enter phantom, 3+2k enter plane k, 4+2k leave plane k, 11 cross the face of a Box region */
      /* the medium of this flight: the Box region the photon is in (BoxVolume with its own material), else the world */
      int cur = -1, mat_here = g->world_mat;
      for (int k = 0; k < g->nreg; ++k) {
        real_t pb[3];
        aff_apply(&g->reg[k].w2b, h->p, pb);
        if (cur < 0 && inside_box(pb, g->reg[k].half)) { cur = k; mat_here = g->reg[k].mat; }
      }
      real_t mu = tab(g->mu, g->ne, mat_here, h->ie, h->fe);
      real_t t_best = -R_LOG(rng_uniform(&h->rng)) / mu;
      if (t_exit < t_best) { t_best = t_exit; ev = 1; }

      int ax_in = -1;
      if (g->has_vol) {
        real_t pl[3], dl[3], ti, to;
        int ai, ao;
        aff_apply(&g->w2v, h->p, pl);
        rot_apply(g->w2v.m, h->d, dl);
        if (ray_box_ax(pl, dl, g->half, &ti, &to, &ai, &ao) && ti < t_best) { t_best = ti; ev = 2; ax_in = ai; }
      }
      for (int k = 0; k < g->nphsp; ++k) {
        real_t pb[3], db[3], ti, to;
        aff_apply(&g->phsp[k].w2b, h->p, pb);
        rot_apply(g->phsp[k].w2b.m, h->d, db);
        if (!ray_box(pb, db, g->phsp[k].half, &ti, &to)) {
          h->inside_plane[k] = 0;
          continue;
        }
        if (h->inside_plane[k]) {
          if (to < t_best) { t_best = to; ev = 4 + 2 * k; }
        } else if (ti < t_best) { t_best = ti; ev = 3 + 2 * k; }
      }
      /* faces of the Box regions, after the planes (a Box that is both reports the crossing as the plane's event) */
      for (int k = 0; k < g->nreg; ++k) {
        real_t pb[3], db[3], ti, to;
        aff_apply(&g->reg[k].w2b, h->p, pb);
        rot_apply(g->reg[k].w2b.m, h->d, db);
        if (!ray_box(pb, db, g->reg[k].half, &ti, &to)) continue;
        if (k == cur) {
          if (to < t_best) { t_best = to; ev = 11; }
        } else if (ti < t_best) { t_best = ti; ev = 11; }
      }
      real_t step = (ev >= 3) ? t_best + NUDGE : t_best;
      if (ev == 1) { h->st.escaped++; h->path += t_best; return; }
      for (int a = 0; a < 3; ++a) h->p[a] = h->p[a] + h->d[a] * step;
      h->path += step;
      if (ev == 0) {
        int proc = select_process(h, mat_here, rng_uniform(&h->rng));
        alive = interact(h, proc, mat_here, 0, 0, 0, pend_p, pend_d, pend);
      } else if (ev == 11) {
        h->st.steps++;   /* the next flight starts on the other side of the face, in the other medium */
      } else if (ev == 2) {
        h->st.steps++;
        to_phantom(h);
        snap_to_face(h->p, h->d, g->half, ax_in, 0);
        h->in_phantom = 1;
        h->seg = 0.0f;
      } else {
        int k = (ev - 3) >> 1;
        h->st.steps++;
        if (((ev - 3) & 1) == 0) {
          h->inside_plane[k] = 1;
          if (g->phsp[k].steps & GATE_B200_PHSP_ENTERING) phsp_store(h, k, GATE_B200_PHSP_ENTERING);
        } else {
          h->inside_plane[k] = 0;
          if (g->phsp[k].steps & GATE_B200_PHSP_EXITING) phsp_store(h, k, GATE_B200_PHSP_EXITING);
        }
      }
    } else {
      /* Woodcock delta tracking against the majorant of the materials in the volume.
       * One call = up to WOODCOCK_K tentative collisions fed by 2 Philox blocks: u[2k] gives
       * the flight distance of iteration k, u[2k+1] its acceptance test. */
      real_t maj = R_FMA(h->fe, g->maj[h->ie + 1] - g->maj[h->ie], g->maj[h->ie]);
      real_t inv_maj = 1.0f / maj;
      real_t u[2 * WOODCOCK_K];
      rng_begin(&h->rng);
      for (int k = 0; k < 2 * WOODCOCK_K; ++k) u[k] = rng_uniform(&h->rng);
      real_t tau0 = -R_LOG(u[0]); /* optical depth to the first event of the call */
      int shell_done = 0;
      if (g->shell && !inside_tbox(g, h->p)) {
        /* The call starts in the homogeneous shell (outside the inner box).  The optical depth of the first flight
         * first runs down at the attenuation of the shell material along the way to the box (or, if the ray misses
         * it, to the face of the volume): either it is used up there -- a real interaction in the shell, no
         * rejection needed, the acceptance variate of the flight selects the process --, or the photon reaches the
         * box with the remainder, or it leaves the volume.  Slab tests against the box (entry) and the volume (exit)
         * with one reciprocal per axis. */
        const real_t mu_s = tab(g->mu, g->ne, g->shell_mat, h->ie, h->fe);
        real_t tmin = -BIG, tmax = BIG, tout = BIG;
        int ain = -1, miss = 0;
        for (int a = 0; a < 3; ++a) {
          const real_t q = h->p[a] - g->tc[a];
          if (h->d[a] != 0.0f) {
            const real_t inv = 1.0f / h->d[a];
            real_t t1 = (-g->th[a] - q) * inv, t2 = (g->th[a] - q) * inv;
            if (t1 > t2) { const real_t w = t1; t1 = t2; t2 = w; }
            if (t1 > tmin) { tmin = t1; ain = a; }
            if (t2 < tmax) tmax = t2;
            const real_t v1 = (-g->half[a] - h->p[a]) * inv, v2 = (g->half[a] - h->p[a]) * inv;
            const real_t vm = v1 > v2 ? v1 : v2;
            if (vm < tout) tout = vm;
          } else if (q < -g->th[a] || q > g->th[a]) {
            miss = 1;
          }
        }
        const real_t t0 = tmin > 0.0f ? tmin : 0.0f;
        const int hit_box = !miss && tmax > t0;
        const int ax_in = tmin > 0.0f ? ain : -1;
        const real_t t_lim = hit_box ? t0 : (tout > 0.0f ? tout : 0.0f);
        if (tau0 < mu_s * t_lim) {
          const real_t ts = tau0 / mu_s;
          for (int a = 0; a < 3; ++a) h->p[a] = h->p[a] + h->d[a] * ts;
          h->path = h->path + ts;
          h->seg = h->seg + ts;
          int proc = select_process(h, g->shell_mat, u[1]);
          tle_march(h, 0.0f);
          int ix = (int)R_FMA(h->p[0], g->inv_sp[0], g->hs[0]), iy = (int)R_FMA(h->p[1], g->inv_sp[1], g->hs[1]),
              iz = (int)R_FMA(h->p[2], g->inv_sp[2], g->hs[2]);
          alive = interact(h, proc, g->shell_mat, ix, iy, iz, pend_p, pend_d, pend);
          shell_done = 1;
        } else if (!hit_box) {
          /* leaves the volume through the shell: restart from the exit point in the world frame */
          real_t ti, to;
          int ai, ao;
          if (!ray_box_ax(h->p, h->d, g->half, &ti, &to, &ai, &ao)) { to = 0.0f; ao = -1; }
          tle_march(h, to);
          for (int a = 0; a < 3; ++a) h->p[a] = h->p[a] + h->d[a] * to;
          snap_to_face(h->p, h->d, g->half, ao, 1);
          h->path += to;
          h->st.steps++;
          to_world(h);
          h->in_phantom = 0;
          shell_done = 1;
        } else {
          real_t q[3] = {h->p[0] + h->d[0] * t0, h->p[1] + h->d[1] * t0, h->p[2] + h->d[2] * t0};
          snap_to_tface(g, q, h->d, ax_in, 0);
          for (int a = 0; a < 3; ++a) h->p[a] = q[a];
          h->path = h->path + t0;
          h->seg = h->seg + t0;
          tau0 = R_FMA(-mu_s, t0, tau0);
        }
      }
      for (int k = 0; k < WOODCOCK_K && !shell_done; ++k) {
        real_t s = (k == 0 ? tau0 : -R_LOG(u[2 * k])) * inv_maj;
        real_t q[3] = {R_FMA(h->d[0], s, h->p[0]), R_FMA(h->d[1], s, h->p[1]), R_FMA(h->d[2], s, h->p[2])};
        /* continuous voxel coordinate: q / spacing + n / 2 in one fused operation */
        real_t fx = R_FMA(q[0], g->inv_sp[0], g->hs[0]);
        real_t fy = R_FMA(q[1], g->inv_sp[1], g->hs[1]);
        real_t fz = R_FMA(q[2], g->inv_sp[2], g->hs[2]);
        if (fx < 0.0f || fy < 0.0f || fz < 0.0f || fx >= (real_t)g->n[0] || fy >= (real_t)g->n[1] ||
            fz >= (real_t)g->n[2]) {
          /* left the volume: restart from the exit point in the world frame */
          real_t ti, to;
          int ai, ao;
          if (!ray_box_ax(h->p, h->d, g->half, &ti, &to, &ai, &ao)) { to = 0.0f; ao = -1; }
          tle_march(h, to);
          for (int a = 0; a < 3; ++a) h->p[a] = h->p[a] + h->d[a] * to;
          snap_to_face(h->p, h->d, g->half, ao, 1);
          h->path += to;
          h->st.steps++;
          to_world(h);
          h->in_phantom = 0;
          break;
        }
        h->p[0] = q[0]; h->p[1] = q[1]; h->p[2] = q[2];
        h->path += s;
        h->seg += s;
        h->st.tentative_collisions++;
        int ix = (int)fx, iy = (int)fy, iz = (int)fz;
        size_t flat = (size_t)ix + (size_t)g->n[0] * ((size_t)iy + (size_t)g->n[1] * (size_t)iz);
        int m = g->label_bytes == 1 ? (int)((const uint8_t *)g->labels)[flat]
                                    : (int)((const uint16_t *)g->labels)[flat];
        real_t mu = tab(g->mu, g->ne, m, h->ie, h->fe);
        real_t t = u[2 * k + 1] * maj;
        if (t <= mu) {
          int proc = select_process(h, m, t / mu);
          tle_march(h, 0.0f);
          alive = interact(h, proc, m, ix, iy, iz, pend_p, pend_d, pend);
          break;
        }
      }
    }
  }
}

typedef struct { dep_t *tab; uint16_t *list; tle_rec *recs; } tle_scratch;

static void run_history(oracle_engine *g, const osource_t *s, uint64_t hist, gate_b200_stats *acc,
                        oracle_trace *trace, tle_scratch *ts) {
  hist_t h;
  memset(&h, 0, sizeof h);
  h.g = g; h.hist = hist; h.trace = trace;
  h.ttab = ts->tab; h.tlist = ts->list; h.nlist = 0; h.recs = ts->recs; h.nrec = 0;
  if (trace) { trace->n_dep = 0; trace->n_iter = 0; }
  rng_init(&h.rng, g->seed, hist);
  generate_primary(&h, s);
  h.st.histories = 1; h.st.tracks = 1;
  if (trace) {
    trace->E0 = h.E;
    for (int a = 0; a < 3; ++a) { trace->p0[a] = h.p[a]; trace->d0[a] = h.d[a]; }
  }
  real_t pend_p[3], pend_d[3];
  int pend = 0;
  /* region of the vertex */
  {
    real_t pl[3];
    h.in_phantom = 0;
    if (g->has_vol) {
      aff_apply(&g->w2v, h.p, pl);
      if (inside_box(pl, g->half)) { to_phantom(&h); h.in_phantom = 1; }
    }
    for (int k = 0; k < g->nphsp; ++k) {
      real_t pb[3];
      aff_apply(&g->phsp[k].w2b, h.p, pb);
      h.inside_plane[k] = (!h.in_phantom) && inside_box(pb, g->phsp[k].half);
    }
  }
  run_photon(&h, pend_p, pend_d, &pend);
  if (pend) { /* second annihilation photon (pair_mode 1) */
    /* pend_* are expressed in the frame current at creation (1 = phantom, 2 = world): go back
     * to the world frame and find the region again */
    real_t pw[3], dw[3];
    if (pend == 1) {
      aff_apply_inv(&g->w2v, pend_p, pw);
      rot_apply_inv(g->w2v.m, pend_d, dw);
    } else {
      for (int a = 0; a < 3; ++a) { pw[a] = pend_p[a]; dw[a] = pend_d[a]; }
    }
    for (int a = 0; a < 3; ++a) { h.p[a] = pw[a]; h.d[a] = dw[a]; }
    h.E = MC2; h.seg = 0.0f; h.in_phantom = 0;
    if (g->has_vol) {
      real_t pl[3];
      aff_apply(&g->w2v, h.p, pl);
      if (inside_box(pl, g->half)) { to_phantom(&h); h.in_phantom = 1; }
    }
    for (int k = 0; k < g->nphsp; ++k) {
      real_t pb[3];
      aff_apply(&g->phsp[k].w2b, h.p, pb);
      h.inside_plane[k] = (!h.in_phantom) && inside_box(pb, g->phsp[k].half);
    }
    int pend2 = 0;
    run_photon(&h, pend_p, pend_d, &pend2);
  }
  flush_buffer(&h);
  if (!g->tle_fold) tle_flush(&h);
  if (g->nphsp > 0 && !h.stored) h.st.absorbed_events++;
  /* accumulate */
  acc->histories += h.st.histories; acc->tracks += h.st.tracks; acc->steps += h.st.steps;
  acc->tentative_collisions += h.st.tentative_collisions;
  acc->n_phot += h.st.n_phot; acc->n_compt += h.st.n_compt; acc->n_rayl += h.st.n_rayl;
  acc->n_conv += h.st.n_conv; acc->deposits += h.st.deposits;
  acc->flushed_entries += h.st.flushed_entries; acc->buffer_overflows += h.st.buffer_overflows;
  acc->escaped += h.st.escaped; acc->absorbed_events += h.st.absorbed_events;
  acc->edep_total_MeV += h.st.edep_total_MeV;
  acc->tle_steps += h.st.tle_steps; acc->tle_entries += h.st.tle_entries; acc->tle_overflows += h.st.tle_overflows;
  if (trace) {
    trace->E_end = h.E;
    trace->edep = (real_t)h.st.edep_total_MeV;
    trace->n_tentative = (int)h.st.tentative_collisions;
    trace->n_interactions = (int)(h.st.n_phot + h.st.n_compt + h.st.n_rayl + h.st.n_conv);
  }
}

/* ------------------------------------------------------------------ API */

static float *dup_f(const float *src, size_t n) {
  float *p = (float *)malloc(n * sizeof(float));
  memcpy(p, src, n * sizeof(float));
  return p;
}

oracle_engine *oracle_create(uint64_t seed) {
  oracle_engine *g = (oracle_engine *)calloc(1, sizeof *g);
  g->seed = seed;
  g->wh[0] = g->wh[1] = g->wh[2] = 1.0e6f;
  g->tle_rec_cap = TLE_REC_MAX;
  return g;
}

void oracle_destroy(oracle_engine *g) {
  if (!g) return;
  free(g->dens); free(g->mu); free(g->cph); free(g->cco); free(g->cra); free(g->maj);
  free(g->q2n); free(g->ray); free(g->sf); free(g->labels);
  for (int i = 0; i < g->nsrc; ++i) { free(g->src[i].spec_e); free(g->src[i].spec_cdf); }
  free(g->src);
  for (int a = 0; a < g->ndose; ++a)
    for (int k = 0; k < 5; ++k) free(g->dose[a].grid[k]);
  for (int a = 0; a < g->ndose; ++a) free(g->dose[a].muen);
  for (int k = 0; k < g->nphsp; ++k) free(g->phsp[k].rows);
  free(g);
}

int oracle_set_tables(oracle_engine *g, const gate_b200_tables *t) {
  g->nmat = t->n_materials; g->ne = t->n_energy; g->nq = t->n_q; g->pair_mode = t->pair_mode;
  size_t n = (size_t)g->nmat * g->ne, nqq = (size_t)g->nmat * g->nq;
  g->dens = dup_f(t->density_g_cm3, g->nmat);
  g->mu = dup_f(t->mu_total_per_mm, n);
  g->cph = dup_f(t->cum_phot, n); g->cco = dup_f(t->cum_compt, n); g->cra = dup_f(t->cum_rayl, n);
  g->ray = dup_f(t->rayl_cum_ff2, nqq); g->sf = dup_f(t->compt_sf, nqq);
  double le0 = log(t->e_min_MeV), le1 = log(t->e_max_MeV);
  g->ln_emin = (real_t)le0;
  g->inv_dlne = (real_t)((double)(g->ne - 1) / (le1 - le0));
  g->e_cut = (real_t)t->e_cut_MeV; g->e_min = (real_t)t->e_min_MeV; g->e_max = (real_t)t->e_max_MeV;
  double lq0 = log(t->q2_min), lq1 = log(t->q2_max);
  g->ln_q2min = (real_t)lq0;
  g->inv_dlnq2 = (real_t)((double)(g->nq - 1) / (lq1 - lq0));
  g->q2n = (float *)malloc(sizeof(float) * g->nq);
  for (int j = 0; j < g->nq; ++j)
    g->q2n[j] = (real_t)exp(lq0 + (lq1 - lq0) * (double)j / (double)(g->nq - 1));
  g->maj = (float *)calloc(g->ne, sizeof(float));
  return 0;
}

static void build_majorant(oracle_engine *g) {
  /* majorant over the materials actually present in the volume */
  char *present = (char *)calloc(g->nmat, 1);
  size_t nv = (size_t)g->n[0] * g->n[1] * g->n[2];
  for (size_t i = 0; i < nv; ++i) {
    int m = g->label_bytes == 1 ? ((uint8_t *)g->labels)[i] : ((uint16_t *)g->labels)[i];
    if (m < g->nmat) present[m] = 1;
  }
  for (int i = 0; i < g->ne; ++i) {
    real_t mx = 0.0f;
    for (int m = 0; m < g->nmat; ++m)
      if (present[m] && g->mu[(size_t)m * g->ne + i] > mx) mx = g->mu[(size_t)m * g->ne + i];
    g->maj[i] = mx;
  }
  free(present);
}

int oracle_set_world(oracle_engine *g, const double half[3], int material) {
  for (int a = 0; a < 3; ++a) g->wh[a] = (real_t)half[a];
  g->world_mat = material;
  return 0;
}

int oracle_set_image_volume(oracle_engine *g, const void *labels, int label_bytes, const int size[3],
                            const double spacing[3], const double w2v[12]) {
  size_t nv = (size_t)size[0] * size[1] * size[2];
  free(g->labels);
  g->labels = malloc(nv * label_bytes);
  memcpy(g->labels, labels, nv * label_bytes);
  g->label_bytes = label_bytes;
  for (int a = 0; a < 3; ++a) {
    g->n[a] = size[a];
    g->sp[a] = (real_t)spacing[a];
    g->inv_sp[a] = (real_t)(1.0 / spacing[a]);
    g->hs[a] = (real_t)(0.5 * size[a]);
    g->half[a] = (real_t)(0.5 * size[a] * spacing[a]);
  }
  aff_from_d(&g->w2v, w2v);
  g->has_vol = 1;
  build_majorant(g);
  {
    int lo[3] = {size[0], size[1], size[2]}, hi[3] = {-1, -1, -1};
    const int l0 = label_bytes == 1 ? ((const uint8_t *)g->labels)[0] : ((const uint16_t *)g->labels)[0];
    size_t i = 0;
    for (int z = 0; z < size[2]; ++z)
      for (int y = 0; y < size[1]; ++y)
        for (int x = 0; x < size[0]; ++x, ++i) {
          const int m = label_bytes == 1 ? ((const uint8_t *)g->labels)[i] : ((const uint16_t *)g->labels)[i];
          if (m == l0) continue;
          if (x < lo[0]) lo[0] = x;
          if (x > hi[0]) hi[0] = x;
          if (y < lo[1]) lo[1] = y;
          if (y > hi[1]) hi[1] = y;
          if (z < lo[2]) lo[2] = z;
          if (z > hi[2]) hi[2] = z;
        }
    /* the shell exists when that box is neither empty (homogeneous volume) nor the whole volume */
    g->shell = hi[0] >= 0 && !(lo[0] == 0 && hi[0] == size[0] - 1 && lo[1] == 0 && hi[1] == size[1] - 1 && lo[2] == 0 && hi[2] == size[2] - 1);
    /* only a shell of gas (the air around a patient) is crossed this way (device: finish_image_volume) */
    if (l0 >= g->nmat || !(g->dens[l0] < 0.05f)) g->shell = 0;
    g->shell_mat = l0;
    for (int a = 0; a < 3; ++a) {
      const double l = -0.5 * size[a] * spacing[a] + lo[a] * spacing[a];
      const double u = -0.5 * size[a] * spacing[a] + (hi[a] + 1) * spacing[a];
      g->tc[a] = g->shell ? (real_t)(0.5 * (l + u)) : 0.0f;
      g->th[a] = g->shell ? (real_t)(0.5 * (u - l)) : 0.0f;
    }
  }
  return 0;
}

int oracle_add_source(oracle_engine *g, const gate_b200_source *s) {
  g->src = (osource_t *)realloc(g->src, sizeof(osource_t) * (g->nsrc + 1));
  osource_t *o = &g->src[g->nsrc];
  memset(o, 0, sizeof *o);
  o->d = *s;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      o->pos.m[3 * r + c] = (real_t)s->pos_rotation[3 * r + c];
      o->dir_rot[3 * r + c] = (real_t)s->dir_rotation[3 * r + c];
    }
    o->pos.t[r] = (real_t)s->pos_translation[r];
    o->mom[r] = (real_t)s->momentum[r];
    o->focus[r] = (real_t)s->focus_point[r];
  }
  o->cos_tmin = (real_t)cos(s->theta[0]); o->cos_tmax = (real_t)cos(s->theta[1]);
  o->sin2_tmin = (real_t)(sin(s->theta[0]) * sin(s->theta[0])); o->sin2_tmax = (real_t)(sin(s->theta[1]) * sin(s->theta[1]));
  o->dir_sx = (real_t)s->dir_sigma[0]; o->dir_sy = (real_t)s->dir_sigma[1];
  o->phi0_turn = (real_t)(s->phi[0] / 6.283185307179586);
  o->dphi_turn = (real_t)((s->phi[1] - s->phi[0]) / 6.283185307179586);
  int ne = 0;
  if (s->ene_type == GATE_B200_ENE_SPECTRUM_DISCRETE || s->ene_type == GATE_B200_ENE_CDF) ne = s->n_spectrum;
  if (s->ene_type == GATE_B200_ENE_SPECTRUM_HISTOGRAM || s->ene_type == GATE_B200_ENE_SPECTRUM_HISTOGRAM_LINEAR)
    ne = s->n_spectrum + 1;
  if (ne > 0) {
    o->spec_e = (float *)malloc(sizeof(float) * ne);
    o->spec_cdf = (float *)malloc(sizeof(float) * s->n_spectrum);
    for (int i = 0; i < ne; ++i) o->spec_e[i] = (real_t)s->spectrum_energy[i];
    for (int i = 0; i < s->n_spectrum; ++i) o->spec_cdf[i] = (real_t)s->spectrum_cdf[i];
  }
  if (s->dir_type == GATE_B200_DIR_HISTOGRAM) {
    const double *src4[4] = {s->dir_theta_edges, s->dir_theta_cdf, s->dir_phi_edges, s->dir_phi_cdf};
    float **dst4[4] = {&o->ht_e, &o->ht_c, &o->hp_e, &o->hp_c};
    for (int k = 0; k < 4; ++k) {
      const int n = (k < 2 ? s->n_dir_theta : s->n_dir_phi) + 1;
      *dst4[k] = (float *)malloc(sizeof(float) * n);
      for (int i = 0; i < n; ++i) (*dst4[k])[i] = (float)src4[k][i];
    }
  }
  if (s->confine) {
    aff_from_d(&o->conf, s->world_to_confine);
    for (int a = 0; a < 3; ++a) o->conf_half[a] = (real_t)s->confine_half_size[a];
  }
  return g->nsrc++;
}

static int grid_is_image_grid(const oracle_engine *g, const odose_t *D);
int oracle_add_dose_actor(oracle_engine *g, const gate_b200_dose_desc *d) {
  if (g->ndose >= ORACLE_MAX_DOSE) return -1;
  odose_t *D = &g->dose[g->ndose];
  memset(D, 0, sizeof *D);
  for (int a = 0; a < 3; ++a) D->size[a] = d->size[a];
  D->flags = d->score_flags | GATE_B200_SCORE_EDEP;
  D->hit_type = d->hit_type;
  /* volume -> index = (world -> index) o (volume -> world), composed in double */
  double w2v[12];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) w2v[4 * r + c] = (double)g->w2v.m[3 * r + c];
    w2v[4 * r + 3] = (double)g->w2v.t[r];
  }
  double v2w[12]; /* inverse rigid */
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) v2w[4 * r + c] = w2v[4 * c + r];
    v2w[4 * r + 3] = -(w2v[0 * 4 + r] * w2v[3] + w2v[1 * 4 + r] * w2v[7] + w2v[2 * 4 + r] * w2v[11]);
  }
  double comp[12];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 4; ++c) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += d->world_to_index[4 * r + k] * v2w[4 * k + c];
      if (c == 3) acc += d->world_to_index[4 * r + 3];
      comp[4 * r + c] = acc;
    }
  }
  aff_from_d(&D->v2i, comp);
  D->aligned = grid_is_image_grid(g, D);
  size_t nv = (size_t)D->size[0] * D->size[1] * D->size[2];
  const int want[5] = {GATE_B200_SCORE_EDEP, GATE_B200_SCORE_EDEP_SQ, GATE_B200_SCORE_DOSE,
                       GATE_B200_SCORE_DOSE_SQ, GATE_B200_SCORE_COUNTS};
  for (int k = 0; k < 5; ++k)
    D->grid[k] = (D->flags & want[k]) ? (double *)calloc(nv, sizeof(double)) : NULL;
  return g->ndose++;
}

/* the grid is the image grid when volume point -> index is p / spacing + n / 2 - 1 / 2 on the image's
 * own lattice: then the voxel of a chord is the image voxel and hit_type cannot matter */
static int grid_is_image_grid(const oracle_engine *g, const odose_t *D) {
  for (int a = 0; a < 3; ++a) {
    if (D->size[a] != g->n[a]) return 0;
    for (int c = 0; c < 3; ++c) {
      double want = a == c ? 1.0 / (double)g->sp[a] : 0.0;
      if (fabs((double)D->v2i.m[3 * a + c] - want) > 1.0e-6 * (1.0 / (double)g->sp[a])) return 0;
    }
    if (fabs((double)D->v2i.t[a] - (0.5 * (double)g->n[a] - 0.5)) > 1.0e-4) return 0;
  }
  return 1;
}

int oracle_add_tle_dose_actor(oracle_engine *g, const gate_b200_dose_desc *d, const gate_b200_tle_desc *t) {
  int id = oracle_add_dose_actor(g, d);
  if (id < 0) return id;
  odose_t *D = &g->dose[id];
  D->tle = 1;
  D->tle_thr = t->energy_threshold_MeV > 3.0e38 ? BIG : (real_t)t->energy_threshold_MeV;
  D->muen = dup_f(t->mu_en_over_rho, (size_t)g->nmat * g->ne);
  D->aligned = grid_is_image_grid(g, D);
  real_t emin = (real_t)t->energy_min_MeV;
  if (g->ntle == 0 || emin > g->tle_emin) g->tle_emin = emin;
  if (g->ntle == 0 || D->tle_thr > g->tle_thr_max) g->tle_thr_max = D->tle_thr;
  g->ntle++;
  if (D->flags & (GATE_B200_SCORE_EDEP_SQ | GATE_B200_SCORE_DOSE_SQ)) g->tle_sq = 1;
  return id;
}

/* Test switch: score the TLE actor by the device's record-fold statement (above) instead of the per-history table.
 * Only for the configuration the device uses it for: ONE dose actor, a TLE actor with squares on the image grid. */
int oracle_set_tle_rec_cap(oracle_engine *g, int cap) {
  if (cap < 1 || cap > TLE_REC_MAX) return -1;
  g->tle_rec_cap = cap;
  return 0;
}
int oracle_set_tle_fold(oracle_engine *g, int on) {
  g->tle_fold = 0;
  if (!on) return 0;
  if (g->ntle != 1 || !g->tle_sq) return -1;
  for (int a = 0; a < g->ndose; ++a)
    if (g->dose[a].tle && g->dose[a].aligned) g->tle_fold = a + 1;   /* that actor's index + 1 */
  return g->tle_fold ? 0 : -1;
}

int oracle_add_box_region(oracle_engine *g, const gate_b200_region_desc *d) {
  if (g->nreg >= 4 || d->material < 0 || d->material >= g->nmat) return -1;
  for (int a = 0; a < 3; ++a) g->reg[g->nreg].half[a] = (real_t)d->half_size[a];
  aff_from_d(&g->reg[g->nreg].w2b, d->world_to_box);
  g->reg[g->nreg].mat = d->material;
  return g->nreg++;
}

int oracle_add_phsp_actor(oracle_engine *g, const gate_b200_phsp_desc *d) {
  if (g->nphsp >= ORACLE_MAX_PHSP) return -1;
  ophsp_t *P = &g->phsp[g->nphsp];
  memset(P, 0, sizeof *P);
  for (int a = 0; a < 3; ++a) P->half[a] = (real_t)d->half_size[a];
  aff_from_d(&P->w2b, d->world_to_box);
  P->steps = d->steps_to_store;
  P->capacity = d->capacity;
  P->rows = (oracle_phsp_row *)calloc(d->capacity ? d->capacity : 1, sizeof(oracle_phsp_row));
  return g->nphsp++;
}

void oracle_begin_run(oracle_engine *g) {
  memset(&g->stats, 0, sizeof g->stats);
  for (int a = 0; a < g->ndose; ++a) {
    size_t nv = (size_t)g->dose[a].size[0] * g->dose[a].size[1] * g->dose[a].size[2];
    for (int k = 0; k < 5; ++k)
      if (g->dose[a].grid[k]) memset(g->dose[a].grid[k], 0, nv * sizeof(double));
  }
  for (int k = 0; k < g->nphsp; ++k) { g->phsp[k].n_rows = 0; g->phsp[k].n_dropped = 0; }
}

static void tle_scratch_init(tle_scratch *ts) {
  ts->tab = (dep_t *)malloc(TLE_H * sizeof(dep_t));
  ts->list = (uint16_t *)malloc(TLE_MAX * sizeof(uint16_t));
  ts->recs = (tle_rec *)malloc(TLE_REC_MAX * sizeof(tle_rec));
  for (int i = 0; i < TLE_H; ++i) ts->tab[i].key = TLE_EMPTY;
}
static void tle_scratch_free(tle_scratch *ts) { free(ts->tab); free(ts->list); free(ts->recs); }

int oracle_run_source(oracle_engine *g, int source_id, uint64_t first, uint64_t n, int n_threads) {
  const osource_t *s = &g->src[source_id];
  gate_b200_stats total;
  memset(&total, 0, sizeof total);
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
  {
    gate_b200_stats acc;
    memset(&acc, 0, sizeof acc);
    tle_scratch ts;
    tle_scratch_init(&ts);
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; ++i) run_history(g, s, first + (uint64_t)i, &acc, NULL, &ts);
    tle_scratch_free(&ts);
#pragma omp critical
    {
      total.histories += acc.histories; total.tracks += acc.tracks; total.steps += acc.steps;
      total.tentative_collisions += acc.tentative_collisions;
      total.n_phot += acc.n_phot; total.n_compt += acc.n_compt; total.n_rayl += acc.n_rayl;
      total.n_conv += acc.n_conv; total.deposits += acc.deposits;
      total.flushed_entries += acc.flushed_entries; total.buffer_overflows += acc.buffer_overflows;
      total.escaped += acc.escaped; total.absorbed_events += acc.absorbed_events;
      total.edep_total_MeV += acc.edep_total_MeV;
      total.tle_steps += acc.tle_steps; total.tle_entries += acc.tle_entries; total.tle_overflows += acc.tle_overflows;
    }
  }
#else
  (void)n_threads;
  {
    tle_scratch ts;
    tle_scratch_init(&ts);
    for (uint64_t i = 0; i < n; ++i) run_history(g, s, first + i, &total, NULL, &ts);
    tle_scratch_free(&ts);
  }
#endif
  gate_b200_stats *S = &g->stats;
  S->histories += total.histories; S->tracks += total.tracks; S->steps += total.steps;
  S->tentative_collisions += total.tentative_collisions;
  S->n_phot += total.n_phot; S->n_compt += total.n_compt; S->n_rayl += total.n_rayl;
  S->n_conv += total.n_conv; S->deposits += total.deposits;
  S->flushed_entries += total.flushed_entries; S->buffer_overflows += total.buffer_overflows;
  S->escaped += total.escaped; S->absorbed_events += total.absorbed_events;
  S->edep_total_MeV += total.edep_total_MeV;
  S->tle_steps += total.tle_steps; S->tle_entries += total.tle_entries; S->tle_overflows += total.tle_overflows;
  S->phsp_rows = 0; S->phsp_dropped = 0;
  for (int k = 0; k < g->nphsp; ++k) {
    uint64_t stored = g->phsp[k].n_rows < g->phsp[k].capacity ? g->phsp[k].n_rows : g->phsp[k].capacity;
    S->phsp_rows += stored; S->phsp_dropped += g->phsp[k].n_dropped;
  }
  return 0;
}

int oracle_trace_history(oracle_engine *g, int source_id, uint64_t hist, oracle_trace *out) {
  gate_b200_stats acc;
  memset(&acc, 0, sizeof acc);
  tle_scratch ts;
  tle_scratch_init(&ts);
  run_history(g, &g->src[source_id], hist, &acc, out, &ts);
  tle_scratch_free(&ts);
  return 0;
}

const double *oracle_dose_grid(oracle_engine *g, int actor, int which) { return g->dose[actor].grid[which]; }
void oracle_get_stats(oracle_engine *g, gate_b200_stats *out) { *out = g->stats; }
uint64_t oracle_phsp_rows(oracle_engine *g, int actor, const oracle_phsp_row **rows) {
  *rows = g->phsp[actor].rows;
  return g->phsp[actor].n_rows < g->phsp[actor].capacity ? g->phsp[actor].n_rows : g->phsp[actor].capacity;
}

/* exported for unit tests of the arithmetic contract */
float oracle_logf(float x) { return o_logf(x); }
float oracle_expf(float x) { return o_expf(x); }
void oracle_sincos2pi(float t, float *s, float *c) { o_sincos2pi(t, s, c); }
void oracle_philox(uint64_t seed, uint64_t hist, uint32_t block, uint32_t out[4]) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t ctr[4] = {(uint32_t)hist, (uint32_t)(hist >> 32), block, 0};
  philox_block(key, ctr, out);
}
void oracle_uniforms(uint64_t seed, uint64_t hist, int n, float *out) {
  rng_t r;
  rng_init(&r, seed, hist);
  for (int i = 0; i < n; ++i) out[i] = rng_uniform(&r);
}
