// transport.cuh — device data model + the photon transport kernel (sm_100a).
//
// Execution model (DESIGN.md "kernel"): one persistent CTA per SM owns kSlots photon slots whose
// state lives in shared memory.  A photon is always waiting for exactly one kind of work:
//   Q_V  Woodcock flight through the label volume (kWoodcockK speculative tentative collisions)
//   Q_C  one Compton trial (Klein-Nishina x scattering function) and, if accepted, its final state
//   Q_R  the rare interactions: Rayleigh, conversion with annihilation photons
//   Q_W  one flight segment in the homogeneous world (entry, exit, phase-space planes)
//   Q_E  end of a photon: second annihilation photon, or close the history (flush its coalesced
//        deposits with warp-aggregated fp64 atomics) and start the next primary
//   Q_T  track-length estimator: walk kTleSteps voxel chords of the flight that just ended (or
//        write kTleSteps entries of the history's TLE table to the grids), re-queued until done
// and sits in the shared-memory ring queue of that kind.  Every warp of the CTA runs the same loop:
// claim up to 32 slots from the fullest queue, run that phase for all of them in lock-step, push each
// slot to the queue of its next phase.  This is stream compaction inside a persistent kernel: warps
// always execute one code path with (nearly) all lanes, whatever the mix of photon states.
// The history -> random-number mapping does not depend on the scheduling: every phase call starts on
// a fresh Philox block of that history (counter = history id, block number).
// Reference semantics restated (OpenGATE/opengate @ 9b07951): see include/gate_b200.h for the
// file:line of each replaced interface; the oracle (oracle/transport_oracle.c) is the scalar
// statement of the same algorithm and is matched bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/gate_b200.h"
#include "device_math.cuh"

#ifndef GB2_THREADS
// 28 warps x 72 registers.  Measured on B200 with everything inlined (profiles/experiments/README.md):
// 1024 threads (64 registers, 216 B of spills) C3 1.40e9 / C2 1.25e9 / C1 1.11e9 primaries/s,
// 896 (72 registers) 1.42 / 1.30 / 1.10, 768 (80 registers) 1.42 / 1.33 / 1.08; two CTAs of 640 threads
// per SM (48 registers, 40 warps) 0.81 / 0.89 / 0.75: more warps do not help, spills and a second
// phase's code in the instruction cache hurt.
#define GB2_THREADS 896
#endif
#ifndef GB2_MIN_CTAS
#define GB2_MIN_CTAS 1                     // resident CTAs per SM the register budget is compiled for
#endif
#ifndef GB2_SLOTS
// 64 B of shared memory per slot (52 B photon state + 6 queue cells).  The slot count sets the shared
// memory carve-out of the SM and with it what is left as L1 for the table gathers (measured on B200,
// profiles/experiments/README.md): 3072 slots + 6 queues fall into the 228 KB carve-out (C3 6.8e8/s),
// 3040 fit the 196 KB one (1.05e9/s), 2528 the 164 KB one (1.12e9/s), 2016 the 132 KB one (1.06e9/s).
#define GB2_SLOTS 2528
#endif

namespace gb2 {

constexpr int kThreads = GB2_THREADS;
#ifndef GB2_TLE_THREADS
#define GB2_TLE_THREADS 640              // 96 registers: the march state stays in registers (C3 TLE +5..9 % over 896 threads)
#endif
// the kernel with the track-length-estimator phase runs with fewer threads and more registers each
// (the voxel march keeps a long walk state live; at 72 registers it spills 700 bytes)
constexpr int kTleThreads = GB2_TLE_THREADS;
template <bool kIsTle> constexpr int threads_of() { return kIsTle ? kTleThreads : kThreads; }
constexpr int kSlots = GB2_SLOTS;          // photon slots per CTA (< 65535: queue cells are 16 bit)
#ifndef GB2_TLE_SLOTS
// measured on B200 (C3, TLE with event-wise squares / without, histories/s): 2528 slots 1.795e7 / 6.49e7,
// 1280 1.827e7 / 6.53e7, 640 1.780e7 / 6.37e7, 320 1.200e7 / 2.82e7: the phase is bound by the dependent chain of a
// chord, not by the number of histories in flight -- 640 slots (one per thread) need a quarter of the table memory
#define GB2_TLE_SLOTS 640
#endif
// slots a CTA of the track-length-estimator kernel puts to work (the rest of the kSlots stay empty): every history
// in flight owns a 64 KB table of per-voxel sums there, and the live part of those tables is what that kernel's
// DRAM traffic is made of
constexpr int kTleSlots = GB2_TLE_SLOTS < GB2_SLOTS ? GB2_TLE_SLOTS : GB2_SLOTS;
// kernel variants (template parameter kMode): 1 and 5 carry the track-length-estimator phase (5: the sums of a history per
// voxel re-evaluated from its flight records instead of kept in a table), 2-4 are the single-dose-actor variants
__host__ __device__ constexpr bool is_tle(int m) { return m == 1 || m == 5; }
__host__ __device__ constexpr bool is_one(int m) { return m >= 2 && m <= 4; }
template <int kMode> constexpr int kActiveSlots = is_tle(kMode) ? kTleSlots : kSlots;
constexpr int kWoodcockK = 4;              // tentative collisions per Woodcock call (2 Philox blocks)
#ifndef GB2_COMPTON_LAZY
#define GB2_COMPTON_LAZY 0
#endif
#ifndef GB2_CLOSE_FUSED
#define GB2_CLOSE_FUSED 1
#endif
#ifndef GB2_PREFETCH_KEYS
#define GB2_PREFETCH_KEYS 1
#endif
#ifndef GB2_PREFETCH_PROB
#define GB2_PREFETCH_PROB 0                // first n tentative points of a call whose process probabilities are prefetched
#endif
#ifndef GB2_SPLIT
#define GB2_SPLIT 0                        // > 0: that many of the four schedulers' warps prefer the Woodcock queue
#endif
#ifndef GB2_VWIDTH
#define GB2_VWIDTH 1
#endif
#ifndef GB2_CLASSVOL
#define GB2_CLASSVOL 0                    // 1: Woodcock phase with early rejection by a 2-bit material-class volume (Params::cls); measured: -7 %
#endif
#ifndef GB2_CHAIN
// Phase chaining: when at least GB2_CHAIN_MIN lanes of a warp leave a single-slot phase for the world phase or the
// end-of-photon phase, the warp runs that phase for them at once instead of sending them through the queues (up to
// GB2_CHAIN times in a row).  The typical chain of a history that leaves the volume -- world (exit, escape) -> end
// (close the history, next primary) -> world (first flight to the volume) -- then costs one visit of the scheduler
// instead of three.  Every phase call still starts on a fresh Philox block: the histories do not change.
#define GB2_CHAIN 0                       // measured (profiles/r2x6_ab_phase_chaining.log): 2 or 3 links cost 2.6 ... 5.6 %
#endif
#ifndef GB2_CHAIN_MIN
#define GB2_CHAIN_MIN 16
#endif
#ifndef GB2_TLE_PREFETCH
#define GB2_TLE_PREFETCH 0                // 1 / 2: next-voxel label prefetch to L2 / L1; measured: no effect (the march is bound by the DRAM rate of random sectors)
#endif
#ifndef GB2_SPLIT_SHELL
// Woodcock calls that START in the homogeneous shell of the image volume wait in a queue of their own (the cells of
// Q_T, which the kernels without the track-length estimator do not use): the shell crossing at the head of such a
// call (about 300 instructions) then runs with full warps, and the calls that start inside the box carry none of it
#define GB2_SPLIT_SHELL 1
#endif
constexpr int kVWidth = GB2_VWIDTH;        // slots one lane carries through a Woodcock call
constexpr int kMaxDose = 4;
constexpr int kMaxPhsp = 4;
constexpr int kMaxReg = 4;                // homogeneous Box regions of their own material in the world
constexpr int kDepBuf = 32;                // per-history coalescing buffer entries
constexpr int kTleSteps = 16;              // voxel chords (or table entries) per call of the TLE phase
constexpr int kTleLog2 = 12;
constexpr int kTleH = 1 << kTleLog2;       // per-history table of (actor, voxel) sums (TLE actor with squares)
constexpr int kTleMax = 3072;              // entries at which the table is written out (an overflow)
constexpr int kTleEmpty = -1;
constexpr int kTleRecMax = 64;             // scored flights + point deposits of one history kept apart (oracle TLE_REC_MAX)
constexpr int kTleRecW = 4;                // float4 words of one record
constexpr float kNudge = 1.0e-3f;          // mm pushed across a boundary
constexpr float kBig = 3.0e38f;
constexpr float kMc2 = 0.51099891f;
constexpr float kKPerMeV = 268.17264f;     // photon wave number (1/bohr) per MeV
constexpr float kCmmPerNs = 299.792458f;
constexpr unsigned kMaxBlocks = 4000000u; // Philox blocks one history may consume (runaway guard)

enum { Q_V = 0, Q_C = 1, Q_R = 2, Q_W = 3, Q_E = 4, Q_T = 5, Q_COUNT = 6, Q_RETIRED = -1 };

struct Affine { float m[9]; float t[3]; };

struct DSource {
  int pos_type, dir_type, ene_type, dir_relative;
  float half[3], radius, sigma_x, sigma_y;
  Affine pos;
  float dir_rot[9];
  float cos_tmin, cos_tmax, phi0_turn, dphi_turn;
  float sin2_tmin, sin2_tmax, dir_sx, dir_sy;   // cosine law, beam2d
  float mom[3], focus[3];
  float mono, sigma, emin, emax, weight;
  int n_spec;
  const float *spec_e, *spec_cdf;
  int vox_n[3];
  float vox_sp[3], vox_org[3];
  const float *cdf_z, *cdf_y, *cdf_x;
  // phase-space replay: entry e = (ph_start + history id) % ph_n, ph_a[e] = {x, y, z, E}, ph_b[e] = {dx, dy, dz, w}
  const float4 *ph_a, *ph_b;
  unsigned long long ph_n, ph_start;
  // user angular histograms (dir_type HISTOGRAM): edges and cumulative probabilities, [n + 1] each
  int n_ht, n_hp;
  const float *ht_e, *ht_c, *hp_e, *hp_c;
  // position.confine: the position is drawn again until it lies in this box (world -> box frame, half sizes)
  int confine;
  Affine conf;
  float conf_half[3];
};

struct DDose {
  int size[3];
  Affine v2i;
  int flags, hit_type;
  // accumulators: the scored ones of edep, edep2, dose, dose2 are the doubles of ONE record per voxel
  // (1, 2 or 4 doubles: all the adds of a deposit land in one 32-byte DRAM sector); grid[k] points at
  // component k of voxel 0 and consecutive voxels are gstride doubles apart.  counts (grid[4]) is a
  // plain array.
  double *grid[5];
  int gstride;
  // track-length estimator (GateTLEDoseActor): tle = 1; aligned = the grid is the image grid
  int tle, aligned;
  float tle_thr;
  const float2 *muen2;   // [ne][nmat] {mu_en/rho(ie), mu_en/rho(ie+1)}, cm2/g
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

struct DRegion {
  float half[3];
  Affine w2b;
  int mat;
};

struct DepEntry { int key; float e; float d; int n; };   // one (voxel, history) partial sum

// indices into the statistics array (unsigned long long)
enum { S_HIST = 0, S_TRACKS, S_STEPS, S_TENT, S_PHOT, S_COMPT, S_RAYL, S_CONV, S_DEP, S_FLUSH,
       S_OVER, S_ESC, S_ABSORBED, S_TLE_STEPS, S_TLE_ENT, S_TLE_OVER, S_COUNT };

struct Params {
  unsigned long long seed;
  int nmat, ne, nq, pair_mode;
  float ln_emin, inv_dlne, e_cut;
  float ln_q2min, inv_dlnq2;
  // energy-major tables: a photon keeps its energy between interactions, so all the Woodcock
  // look-ups of one flight hit the same two rows
  const float2 *mu2;     // [ne][nmat] {mu(ie), mu(ie+1)} total attenuation, 1/mm
  const float4 *prob;    // [ne][nmat] {P(phot), P(phot)+P(compt), ...+P(rayl), 0}
  const float *dens, *maj, *q2n, *ray, *sf;
  float wh[3];
  int world_mat;
  int has_vol;
  int n[3];
  float half[3], inv_sp[3], sp[3], hs[3];   // hs = n / 2: voxel coordinate = p * inv_sp + hs
  Affine w2v;
  const void *labels;
  // Material CLASS volume: 2 bits per voxel (4 voxels per byte, x fastest), the materials present in the volume sorted
  // into four classes by density; cmax[ie] = the largest attenuation coefficient of each class at energy node ie.
  // A tentative collision whose acceptance variate exceeds the bound of its voxel's class is rejected whatever the
  // material is -- the same decision the exact test takes -- and needs neither the label nor the attenuation
  // look-up; only the others (every accepted collision among them) gather the label.  With at most four materials
  // present the class IS the material (cls_exact) and the label volume is not read by the Woodcock phase at all.
  const unsigned char *cls;
  const float4 *cmax;            // [ne]
  int cls_exact;
  int cls_mat[4];
  // The homogeneous shell of the image volume: every voxel outside the box (centre tc, half size th, volume frame)
  // carries the label shell_mat (the air around a patient).  A Woodcock call that starts out there first crosses
  // the shell analytically with the optical depth of its first flight (see volume_gather_labels).
  int shell, shell_mat;
  float tc[3], th[3];
  int ndose;
  DDose dose[kMaxDose];
  int nphsp;
  DPhsp phsp[kMaxPhsp];
  int nreg;                      // Box regions (general kernel variants only: 0, 1, 5)
  DRegion reg[kMaxReg];
  DSource src;
  unsigned long long first, count;
  unsigned long long *next;      // history counter of this launch
  unsigned long long *stats;     // [S_COUNT]
  double *edep_total;
  DepEntry *depbuf;              // [gridDim.x * kSlots][kDepBuf] coalescing buffers (global scratch)
  float *pend;                   // [gridDim.x * kSlots][6] second annihilation photon (pair_mode 1)
  float *slot_w;                 // [gridDim.x * kSlots] weight of the history in the slot (phase-space sources)
  // track-length estimators (dose actors with tle = 1)
  int ntle;
  float tle_emin, tle_thr_max;
  DepEntry *tle_tab;             // [gridDim.x * kTleSlots][kTleH] per-history (actor, voxel) sums, only with squares
  unsigned short *tle_list;      // [slots][kTleMax] occupied table cells in insertion order
  uint4 *tle_state;              // [slots] {t along the flight (-1 new flight, -2 writing the table out),
                                 //          ix | iy << 16 (or next list entry), iz, entries in the table}
  uint2 *tle_aux;                // [slots] {t before which the current flight is written out already, records of the history}
  // kernel variant 5 (ONE TLE actor with squares, on the image grid; other dose actors analog): no table, the sum a
  // voxel holds is re-evaluated from the records of the history's scored flights and point deposits
  int tle_a;                     // index of that actor
  int tle_rec_cap;               // records a history may hold (<= kTleRecMax; lowered by tests to reach the overflow path)
  int label_bytes;
  float4 *tle_rec;               // [slots][kTleRecMax][kTleRecW]
};

// Shared memory of one CTA: photon slots (structure of arrays), the four ring queues, statistics.
struct Smem {
  float px[kSlots], py[kSlots], pz[kSlots];   // world frame outside the volume, volume frame inside
  float dx[kSlots], dy[kSlots], dz[kSlots];
  float E[kSlots], tbin[kSlots], path[kSlots], seg[kSlots];
  unsigned hidx[kSlots];                      // history index inside this launch (id = first + hidx)
  unsigned blk[kSlots];                       // next Philox block of the history
  unsigned fm[kSlots];                        // flags (low 20 bits) | material of the pending interaction << 20
  unsigned short cell[Q_COUNT][kSlots];       // ring queues of slot numbers, kEmpty = free cell
  int head[Q_COUNT], tail[Q_COUNT], avail[Q_COUNT];
  int cur;                                    // phase the CTA is currently draining
  int cur2;                                   // GB2_SPLIT: phase the second warp group is draining
  int retired;
  unsigned stats_lo[S_COUNT], stats_hi[S_COUNT];   // 64-bit counters as two words: shared memory has no native 64-bit add
  double edep;
};
constexpr unsigned kEmpty = 0xffffu;

// bits of Smem::fm (flags in the low 20 bits, material label in the high 12)
constexpr unsigned F_INSIDE = 0xfu;        // bit k: currently inside phase-space box k
constexpr unsigned F_STORED = 1u << 4;     // the history left at least one phase-space row
constexpr unsigned F_PEND_SHIFT = 5;       // 0 none, 1 pending photon in the volume frame, 2 world frame
constexpr unsigned F_PEND = 3u << F_PEND_SHIFT;
constexpr unsigned F_INVOL = 1u << 7;      // the pending interaction is inside the image volume
constexpr unsigned F_OPEN = 1u << 8;       // a history is in flight in this slot
constexpr unsigned F_NBUF_SHIFT = 9;       // entries in the coalescing buffer (0..kDepBuf)
constexpr unsigned F_NBUF = 0x3fu << F_NBUF_SHIFT;
constexpr unsigned F_ABSORB = 1u << 15;    // Q_E: the photon is absorbed on the spot first (phot / conv kerma)
constexpr unsigned F_EXIT = 1u << 16;      // Q_W: the photon is leaving the image volume (still in its frame)
constexpr unsigned F_CONV = 1u << 17;      // the pending interaction is a conversion (else Rayleigh / phot)
constexpr unsigned F_NEXT_SHIFT = 18;      // Q_T: queue to continue with, minus one (Q_C..Q_E)
constexpr unsigned F_NEXT = 3u << F_NEXT_SHIFT;
constexpr unsigned F_MAT_SHIFT = 20;
constexpr unsigned F_MASK = (1u << F_MAT_SHIFT) - 1u;
constexpr int kMaxMaterials = 1 << 12;

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

// ray / point against the inner box of the shell (centre P.tc, half size P.th, volume frame)
__device__ __forceinline__ bool ray_tbox_ax(const Params &P, const float p[3], const float d[3], float &t_in, float &t_out, int &a_in, int &a_out) {
  const float q[3] = {p[0] - P.tc[0], p[1] - P.tc[1], p[2] - P.tc[2]};
  return ray_box_ax(q, d, P.th, t_in, t_out, a_in, a_out);
}
__device__ __forceinline__ bool inside_tbox(const Params &P, const float p[3]) {
  const float q[3] = {p[0] - P.tc[0], p[1] - P.tc[1], p[2] - P.tc[2]};
  return inside_box(q, P.th);
}
__device__ __forceinline__ void snap_to_tface(const Params &P, float p[3], const float d[3], int a, bool outward) {
  float q[3] = {p[0] - P.tc[0], p[1] - P.tc[1], p[2] - P.tc[2]};
  snap_to_face(q, d, P.th, a, outward);
#pragma unroll
  for (int k = 0; k < 3; ++k) p[k] = q[k] + P.tc[k];
}

__device__ __forceinline__ int lower_bound_f(const float *__restrict__ a, int n, float p) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < p) lo = mid + 1; else hi = mid;
  }
  return lo;
}


__device__ __forceinline__ int ld_vol(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_vol(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }
__device__ __forceinline__ unsigned ld_cell(const unsigned short *p) { return *reinterpret_cast<const volatile unsigned short *>(p); }
__device__ __forceinline__ void st_cell(unsigned short *p, unsigned v) { *reinterpret_cast<volatile unsigned short *>(p) = (unsigned short)v; }

// The coalescing buffers and the pending-photon scratch live in global memory and are handed from
// warp to warp with the slot: access them at L2 (ld.cg / st.cg), never through the read-only path.
#ifndef GB2_DEP_L1
#define GB2_DEP_L1 1
#endif
#ifndef GB2_TLE_L1
#define GB2_TLE_L1 GB2_DEP_L1
#endif
// Per-slot scratch in global memory (coalescing buffers, pending photons, history weights, TLE state and tables)
// is only ever touched by the CTA that owns the slot, i.e. by ONE SM, and a slot changes hands between warps
// through the shared-memory queues with CTA-scope fences on both sides: ordinary (L1-cached) loads and stores
// are coherent for that.  Round 1 used ld.cg / st.cg here: every search of the coalescing buffer and every flush
// then paid an L2 round trip (the hottest stall of the deposit and end-of-history code); through L1 the entries
// a history wrote a few microseconds ago are hits (measured: C3 +17 %, C2 +42 %, C1 +39 %, C4 +47 %).
template <typename T>
__device__ __forceinline__ T sc_ld(const T *p) {
#if GB2_DEP_L1
  return *p;
#else
  return __ldcg(p);
#endif
}
template <typename T>
__device__ __forceinline__ void sc_st(T *p, T v) {
#if GB2_DEP_L1
  *p = v;
#else
  __stcg(p, v);
#endif
}
template <typename T>
__device__ __forceinline__ T tl_ld(const T *p) {
#if GB2_TLE_L1
  return *p;
#else
  return __ldcg(p);
#endif
}
template <typename T>
__device__ __forceinline__ void tl_st(T *p, T v) {
#if GB2_TLE_L1
  *p = v;
#else
  __stcg(p, v);
#endif
}
__device__ __forceinline__ DepEntry ld_dep(const DepEntry *p) {
#if GB2_DEP_L1
  const int4 v = *reinterpret_cast<const int4 *>(p);
#else
  const int4 v = __ldcg(reinterpret_cast<const int4 *>(p));
#endif
  DepEntry e; e.key = v.x; e.e = __int_as_float(v.y); e.d = __int_as_float(v.z); e.n = v.w;
  return e;
}
// the per-history tables of the track-length estimator (64 KB per slot, random access)
__device__ __forceinline__ DepEntry ld_tab(const DepEntry *p) {
#if GB2_TLE_L1
  const int4 v = *reinterpret_cast<const int4 *>(p);
#else
  const int4 v = __ldcg(reinterpret_cast<const int4 *>(p));
#endif
  DepEntry e; e.key = v.x; e.e = __int_as_float(v.y); e.d = __int_as_float(v.z); e.n = v.w;
  return e;
}
__device__ __forceinline__ void st_tab(DepEntry *p, const DepEntry &e) {
#if GB2_TLE_L1
  *reinterpret_cast<int4 *>(p) = make_int4(e.key, __float_as_int(e.e), __float_as_int(e.d), e.n);
#else
  __stcg(reinterpret_cast<int4 *>(p), make_int4(e.key, __float_as_int(e.e), __float_as_int(e.d), e.n));
#endif
}
__device__ __forceinline__ void st_dep(DepEntry *p, const DepEntry &e) {
#if GB2_DEP_L1
  *reinterpret_cast<int4 *>(p) = make_int4(e.key, __float_as_int(e.e), __float_as_int(e.d), e.n);
#else
  __stcg(reinterpret_cast<int4 *>(p), make_int4(e.key, __float_as_int(e.e), __float_as_int(e.d), e.n));
#endif
}

// statistical weight of the history in a slot: per entry for a phase-space source (written by
// generate_primary), the source's constant otherwise
template <int kMode>
__device__ __forceinline__ float hist_weight(const Params &P, int gslot) {
  if (is_one(kMode)) return P.src.weight;
  return P.src.ph_b ? sc_ld(P.slot_w + gslot) : P.src.weight;
}

__device__ __forceinline__ float tab_mu(const Params &P, int m, int ie, float fe) {
  const float2 v = __ldg(P.mu2 + (size_t)ie * P.nmat + m);
  return __fmaf_rn(fe, v.y - v.x, v.x);
}

// 64-bit statistics counter in shared memory: a native 32-bit add plus a carry (a 64-bit atomicAdd on
// shared memory is a compare-and-swap loop)
__device__ __forceinline__ void stat_add64(Smem &S, int idx, unsigned v) {
  const unsigned old = atomicAdd(&S.stats_lo[idx], v);
  if (old + v < old) atomicAdd(&S.stats_hi[idx], 1u);
}
// The event counters of one CTA and one launch stay below 2^32 (the host gives a launch at most 2^24 histories per
// CTA, the per-history counts are a few units): a plain 32-bit add whose result nobody waits for (a fire-and-forget
// reduction) instead of add + carry.  The counters that can exceed 2^32 per CTA -- tentative collisions, chords
// of the track-length estimator -- keep the carry.
__device__ __forceinline__ void stat_add32(Smem &S, int idx, unsigned v) { atomicAdd(&S.stats_lo[idx], v); }
// opportunistic warp aggregation of a statistics increment (the lanes that arrive together)
__device__ __forceinline__ void stat_add(Smem &S, int idx, unsigned v) {
  const unsigned act = __activemask();
  const unsigned sum = __reduce_add_sync(act, v);
  if ((int)(threadIdx.x & 31u) == __ffs(act) - 1 && sum) stat_add64(S, idx, sum);
}
// Several small per-lane event counts of one phase call packed into bit fields of one word: ONE warp
// reduction for all of them (a field of `bits` bits holds the sum over 32 lanes, so every per-lane
// count must stay below 2^bits / 32).  Returns the packed sums on the leader lane, 0 elsewhere.
__device__ __forceinline__ unsigned stat_reduce_packed(unsigned packed) {
  const unsigned act = __activemask();
  const unsigned sum = __reduce_add_sync(act, packed);
  return (int)(threadIdx.x & 31u) == __ffs(act) - 1 ? sum : 0u;
}
__device__ __forceinline__ void stat_add_small(Smem &S, int idx, unsigned v) {
  const unsigned act = __activemask();
  const unsigned sum = __reduce_add_sync(act, v);
  if ((int)(threadIdx.x & 31u) == __ffs(act) - 1 && sum) stat_add32(S, idx, sum);
}
__device__ __forceinline__ void stat_bump(Smem &S, int idx, unsigned v) {
  if (v) stat_add32(S, idx, v);
}
// the rare part of a deposit's bookkeeping: coalescing-buffer overflows
__device__ __forceinline__ void stat_overflow(Smem &S, unsigned over, unsigned flush) {
  if (__any_sync(__activemask(), over != 0u)) {
    stat_add(S, S_OVER, over);
    stat_add(S, S_FLUSH, flush);
  }
}
__device__ __forceinline__ void stat_tle_overflow(Smem &S, unsigned over, unsigned ent) {
  if (__any_sync(__activemask(), over != 0u)) {
    stat_add(S, S_TLE_OVER, over);
    stat_add(S, S_TLE_ENT, ent);
  }
}
// the TLE bookkeeping of a phase's point deposits: variant 1 counts table entries when a table is written out (an
// overflow), variant 5 when a voxel is touched for the first time by its history
template <int kMode>
__device__ __forceinline__ void stat_tle_deposits(Smem &S, unsigned over, unsigned ent) {
  if (kMode == 5) {
    if (__any_sync(__activemask(), (over | ent) != 0u)) {
      stat_add(S, S_TLE_OVER, over);
      stat_add(S, S_TLE_ENT, ent);
    }
  } else {
    stat_tle_overflow(S, over, ent);
  }
}

__device__ __forceinline__ void energy_bin_t(const Params &P, float E, float &t) {
  t = (d_logf(E) - P.ln_emin) * P.inv_dlne;
  if (t < 0.0f) t = 0.0f;
}
__device__ __forceinline__ void bin_split(const Params &P, float t, int &ie, float &fe) {
  int i = (int)t;
  if (i > P.ne - 2) i = P.ne - 2;
  float f = t - (float)i;
  if (f > 1.0f) f = 1.0f;
  ie = i; fe = f;
}

// ------------------------------------------------------------------------------ photon in registers

struct Ph {
  float p[3], d[3];
  float E, path, seg, fe;
  int ie, mat;
  unsigned flags;
  Rng rng;
};

__device__ __forceinline__ void load_ph(const Params &P, const Smem &S, int s, Ph &g) {
  g.p[0] = S.px[s]; g.p[1] = S.py[s]; g.p[2] = S.pz[s];
  g.d[0] = S.dx[s]; g.d[1] = S.dy[s]; g.d[2] = S.dz[s];
  g.E = S.E[s]; g.path = S.path[s]; g.seg = S.seg[s];
  bin_split(P, S.tbin[s], g.ie, g.fe);
  const unsigned fm = S.fm[s];
  g.flags = fm & F_MASK;
  g.mat = (int)(fm >> F_MAT_SHIFT);
  const unsigned long long id = P.first + S.hidx[s];
  g.rng.k0 = (uint32_t)P.seed; g.rng.k1 = (uint32_t)(P.seed >> 32);
  g.rng.h0 = (uint32_t)id; g.rng.h1 = (uint32_t)(id >> 32); g.rng.blk = S.blk[s]; g.rng.pos = 4;
}
__device__ __forceinline__ void store_ph(Smem &S, int s, const Ph &g) {
  S.px[s] = g.p[0]; S.py[s] = g.p[1]; S.pz[s] = g.p[2];
  S.dx[s] = g.d[0]; S.dy[s] = g.d[1]; S.dz[s] = g.d[2];
  S.E[s] = g.E; S.path[s] = g.path; S.seg[s] = g.seg;
  S.fm[s] = (g.flags & F_MASK) | ((unsigned)g.mat << F_MAT_SHIFT);
  S.blk[s] = g.rng.blk;
}
__device__ __forceinline__ void set_energy(const Params &P, Smem &S, int s, Ph &g) {
  float t;
  energy_bin_t(P, g.E, t);
  S.tbin[s] = t;
  bin_split(P, t, g.ie, g.fe);
}

// ------------------------------------------------------------------------------ scoring

// Fire-and-forget fp64 add on a GLOBAL address.  atomicAdd(double*) on the generic pointers of Params
// compiles to ATOM.E.ADD.F64 with a predicate result the warp waits for (then a QSPC test and a CAS
// loop for the shared-memory case): one blocking L2/DRAM round trip per grid.  red.global has no
// result: the adds of the three or four grids of a deposit leave back to back.
__device__ __forceinline__ void red_add_f64(double *p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}

// Add one (voxel, history) sum to the fp64 grids.  Lanes that arrive together and hit the same
// voxel are combined first (one atomic per distinct address); squares are summed per history, never
// of the combined value (GateDoseActor.cpp:359-381).
template <int kMode = 0>
__device__ __forceinline__ void grid_add(const Params &P, const DepEntry &en) {
  const unsigned lane = threadIdx.x & 31u;
  const int key = en.key;
  const int a = (int)((unsigned)key >> 28);
  const int v = key & 0x0fffffff;
  double e = (double)en.e, d = (double)en.d;
  double e2 = e * e, d2 = d * d, n = (double)en.n;
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
    const DDose &D = P.dose[is_one(kMode) ? 0 : a];   // variants 2-4: the one dose actor, fields as immediates
    const size_t o = (size_t)v * (size_t)D.gstride;
    red_add_f64(D.grid[GATE_B200_GRID_EDEP] + o, e);
    if (D.flags & GATE_B200_SCORE_EDEP_SQ) red_add_f64(D.grid[GATE_B200_GRID_EDEP_SQ] + o, e2);
    if (D.flags & GATE_B200_SCORE_DOSE) red_add_f64(D.grid[GATE_B200_GRID_DOSE] + o, d);
    if (D.flags & GATE_B200_SCORE_DOSE_SQ) red_add_f64(D.grid[GATE_B200_GRID_DOSE_SQ] + o, d2);
    if ((D.flags & GATE_B200_SCORE_COUNTS) && n != 0.0) red_add_f64(D.grid[GATE_B200_GRID_COUNTS] + v, n);
  }
}

// The same for increments given as doubles (kernel variant 5: a voxel's event sum grows from p to p + e and the
// grids get (p + e) - p and (p + e)^2 - p^2).
__device__ __forceinline__ void grid_add_vals(const DDose &D, int v, double e, double e2, double d, double d2, double n) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned act = __activemask();
  const unsigned grp = __match_any_sync(act, v);
  const int leader = __ffs(grp) - 1;
  if (grp != (1u << lane)) {
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
    const size_t o = (size_t)v * (size_t)D.gstride;
    red_add_f64(D.grid[GATE_B200_GRID_EDEP] + o, e);
    if (D.flags & GATE_B200_SCORE_EDEP_SQ) red_add_f64(D.grid[GATE_B200_GRID_EDEP_SQ] + o, e2);
    if (D.flags & GATE_B200_SCORE_DOSE) red_add_f64(D.grid[GATE_B200_GRID_DOSE] + o, d);
    if (D.flags & GATE_B200_SCORE_DOSE_SQ) red_add_f64(D.grid[GATE_B200_GRID_DOSE_SQ] + o, d2);
    if ((D.flags & GATE_B200_SCORE_COUNTS) && n != 0.0) red_add_f64(D.grid[GATE_B200_GRID_COUNTS] + v, n);
  }
}

// Add the coalesced (voxel, history) entries of a closed history to the fp64 grids.
template <int kMode = 0>
__device__ __forceinline__ void flush_buffer(const Params &P, const DepEntry *buf, int nbuf) {
  for (int j = 0; j < nbuf; ++j) grid_add<kMode>(P, ld_dep(buf + j));
}
// the same when entry `kj` is held in registers (the deposit that closed the history)
template <int kMode>
__device__ __forceinline__ void flush_buffer_keep(const Params &P, const DepEntry *buf, int nbuf, const DepEntry &ken, int kj) {
  for (int j = 0; j < nbuf; ++j) grid_add<kMode>(P, j == kj ? ken : ld_dep(buf + j));
}

// ---- per-history table of the TLE actors that score squares: open addressing, Fibonacci hash.
// A history touches hundreds of voxels along its flights; the event-wise squares need its sum per
// voxel (GateTLEDoseActor.cpp:281-299 -> ScoreSquaredValue), so the sums are kept here until the
// history ends.  The table of a slot is only ever touched by the lane that holds the slot.
__device__ __forceinline__ void tle_flush_all(const Params &P, int gslot, unsigned nlist) {
  DepEntry *tab = P.tle_tab + (size_t)gslot * kTleH;
  const unsigned short *list = P.tle_list + (size_t)gslot * kTleMax;
  for (unsigned j = 0; j < nlist; ++j) {
    const unsigned hh = tl_ld(list + j);
    grid_add(P, ld_tab(tab + hh));
    tl_st(&tab[hh].key, (int)kTleEmpty);
  }
}
__device__ __forceinline__ void tle_table_add(const Params &P, int gslot, unsigned &nlist, int key, float e, float d,
                                           int n, unsigned &n_ent, unsigned &n_over) {
  DepEntry *tab = P.tle_tab + (size_t)gslot * kTleH;
  unsigned hh = ((unsigned)key * 2654435761u) >> (32 - kTleLog2);
  for (;;) {
    DepEntry en = ld_tab(tab + hh);
    if (en.key == key) {
      en.e += e; en.d += d; en.n += n;
      st_tab(tab + hh, en);
      return;
    }
    if (en.key == kTleEmpty) {
      if (nlist == (unsigned)kTleMax) {   // table full: write it out, start again
        tle_flush_all(P, gslot, nlist);
        n_ent += nlist;
        n_over++;
        nlist = 0;
        hh = ((unsigned)key * 2654435761u) >> (32 - kTleLog2);
      }
      DepEntry z; z.key = key; z.e = e; z.d = d; z.n = n;
      st_tab(tab + hh, z);
      tl_st(P.tle_list + (size_t)gslot * kTleMax + nlist, (unsigned short)hh);
      nlist++;
      return;
    }
    hh = (hh + 1u) & (unsigned)(kTleH - 1);
  }
}

// a scored flight or a point deposit is about to add to the history's table: one more record behind its content;
// the table is written out first when the history has kTleRecMax of them already (oracle tle_record_begin)
__device__ __forceinline__ void tle_record_begin_tab(const Params &P, int gslot, unsigned &nlist, unsigned &nrec,
                                                     unsigned &n_ent, unsigned &n_over) {
  if (nrec >= (unsigned)P.tle_rec_cap) {
    tle_flush_all(P, gslot, nlist);
    n_ent += nlist;
    n_over++;
    nlist = 0;
    nrec = 0;
  }
  nrec++;
}

// ---- kernel variant 5: the per-voxel sums of a history WITHOUT a table (oracle tle_fold_*).
// A straight flight crosses a voxel once, so two contributions of a history to one voxel come from different flights
// (or point deposits).  The history keeps its scored flights and point deposits as records (64 bytes each, slot
// scratch); what a voxel holds so far is re-evaluated from them: the chord of an earlier flight through a given
// voxel has a closed form with exactly the bits the voxel walk produced -- entry = the latest entering-face distance
// of the axes that moved, exit = the earliest leaving-face distance, every face distance the walk's own expression of
// the integer face index.  With the previous sum p and the new one p + e the grids get (p + e) - p and
// (p + e)^2 - p^2: at the end of the history they hold the event's sum and its square (GateDoseActor.cpp:359-381),
// one 32-byte voxel record touched per chord as in the run without squares, and no table traffic at all
// (the table variant moved 162 KB of DRAM traffic per history on config 3, this one 27 KB).
// record: [0] = {start x, y, z, t_end} (point deposit: {e, dose, key, 0}), [1] = {direction x, y, z, 0.1 E w} (point: 0),
//         [2] = {lo.x | lo.y << 16, lo.z | hi.x << 16, hi.y | hi.z << 16, energy-bin coordinate}: the voxel range walked,
//         [3] = {t before which the flight's chords are not part of the current sums (table overflow), -, -, -}
__device__ __forceinline__ float tle_rec_chord(const Params &P, const float4 &A, const float4 &B, const int iv[3],
                                               const int lo[3], const int hi[3], float &t_in_out) {
  float t_in = 0.0f, t_out = A.w;
  const float s[3] = {A.x, A.y, A.z}, d[3] = {B.x, B.y, B.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (d[a] == 0.0f) continue;
    const float inv = 1.0f / d[a];
    const bool up = d[a] > 0.0f;
    const float f_in = __fmaf_rn((float)(iv[a] + (up ? 0 : 1)), P.sp[a], -P.half[a]);
    const float f_out = __fmaf_rn((float)(iv[a] + (up ? 1 : 0)), P.sp[a], -P.half[a]);
    if (iv[a] != (up ? lo[a] : hi[a])) {   // the walk entered this voxel layer through f_in
      const float tb = (f_in - s[a]) * inv;
      if (tb > t_in) t_in = tb;
    }
    const float tb = (f_out - s[a]) * inv;
    if (tb < t_out) t_out = tb;
  }
  t_in_out = t_in;
  return t_out - t_in;
}
// sum of the earlier contributions of the history to voxel iv of the TLE actor D; m < 0: label not gathered yet
// cand: the stored records that can hold the voxel at all (bit r = record r; in record order = chronological order)
__device__ __forceinline__ bool tle_fold_prev(const Params &P, const DDose &D, int gslot, unsigned long long cand, int key,
                                              const int iv[3], int m, float &pe, float &pd) {
  const float4 *rec = P.tle_rec + (size_t)gslot * (kTleRecMax * kTleRecW);
  bool found = false;
  float se = 0.0f, sd = 0.0f;
  for (; cand; cand &= cand - 1ull) {
    const unsigned r = (unsigned)__ffsll((long long)cand) - 1u;
    const float4 Cf = tl_ld(rec + kTleRecW * r + 2);
    const unsigned cx = __float_as_uint(Cf.x), cy = __float_as_uint(Cf.y), cz = __float_as_uint(Cf.z);
    const int lo[3] = {(int)(cx & 0xffffu), (int)(cx >> 16), (int)(cy & 0xffffu)};
    const int hi[3] = {(int)(cy >> 16), (int)(cz & 0xffffu), (int)(cz >> 16)};
    if (iv[0] < lo[0] || iv[0] > hi[0] || iv[1] < lo[1] || iv[1] > hi[1] || iv[2] < lo[2] || iv[2] > hi[2]) continue;
    const float4 A = tl_ld(rec + kTleRecW * r), B = tl_ld(rec + kTleRecW * r + 1);
    if (B.x == 0.0f && B.y == 0.0f && B.z == 0.0f) {   // a point deposit
      if (__float_as_int(A.z) == key) { se += A.x; sd += A.y; found = true; }
      continue;
    }
    float t_in;
    const float L = tle_rec_chord(P, A, B, iv, lo, hi, t_in);
    if (!(L > 0.0f) || t_in < tl_ld(rec + kTleRecW * r + 3).x) continue;
    if (m < 0) {
      const size_t flat = (size_t)iv[0] + (size_t)P.n[0] * ((size_t)iv[1] + (size_t)P.n[1] * (size_t)iv[2]);
      m = P.label_bytes == 1 ? (int)__ldg(reinterpret_cast<const uint8_t *>(P.labels) + flat)
                             : (int)__ldg(reinterpret_cast<const uint16_t *>(P.labels) + flat);
    }
    const float rho = __ldg(P.dens + m);
    int ie; float fe;
    bin_split(P, Cf.w, ie, fe);
    const float2 mv = __ldg(D.muen2 + (size_t)ie * P.nmat + m);
    const float muen = __fmaf_rn(fe, mv.y - mv.x, mv.x);
    const float e = ((B.w * muen) * L) * rho;
    se += e; sd += e / rho; found = true;
  }
  pe = se; pd = sd;
  return found;
}
// one contribution (chord or point deposit) of the history to voxel iv; tcur: where the current flight stands
__device__ __forceinline__ void tle_fold_add(const Params &P, const DDose &D, int gslot, unsigned &nrec, unsigned &nlist,
                                             float &tmin, unsigned long long &cand, int key, const int iv[3], int m, float e,
                                             float dd, int n, float tcur, unsigned &n_ent, unsigned &n_over) {
  float pe, pd;
  if (!tle_fold_prev(P, D, gslot, cand, key, iv, m, pe, pd)) {
    if (nlist == (unsigned)kTleMax) {   // the table this stands for would be full: forget what was scored so far
      nrec = 1;
      cand = 0ull;
      tmin = tcur;
      nlist = 0;
      n_over++;
    }
    nlist++;
    n_ent++;
  }
  const float ne = pe + e, nd = pd + dd;
  const double e0 = (double)pe, e1 = (double)ne, d0 = (double)pd, d1 = (double)nd;
  grid_add_vals(D, key & 0x0fffffff, e1 - e0, e1 * e1 - e0 * e0, d1 - d0, d1 * d1 - d0 * d0, (double)n);
}

struct DepStat { unsigned dep, over, flush, tle_ent, tle_over; double edep; };
struct DepKeys;
struct DepKeep;

// deposit e (already weighted) at the photon position; in_vol selects whether dose actors see it.
// e_tle: flight energy of the gamma whose secondaries deposit here (a TLE actor ignores them while that
// gamma scores by track length, GateTLEDoseActor.cpp:236-243); kBig for a gamma absorbed below the cut.
// Optional extras of the single-actor kernels (is_one(kMode)):
//  * pk: the keys of the first two buffer entries, loaded by the caller ahead of its own arithmetic (the search of
//    the buffer is then a register compare for the common histories with at most two voxels);
//  * keep: the entry this deposit creates or updates is handed back in registers (keep->n = its index) instead of
//    being stored -- the caller closes the history in the same call and adds it to the grids directly.
struct DepKeys { int k0, k1; };
struct DepKeep { DepEntry en; int j; };
template <int kMode, bool kPre = false, bool kKeep = false>
__device__ __forceinline__ void deposit(const Params &P, Ph &g, DepEntry *buf, DepStat &ds, float e,
                                     int m, bool in_vol, float e_tle, int gslot, const DepKeys &pk, DepKeep &keep) {
  ds.edep += (double)e;
  ds.dep++;
  if (!in_vol) return;
  int nbuf = (int)((g.flags & F_NBUF) >> F_NBUF_SHIFT);
  const int ndose = is_one(kMode) ? 1 : P.ndose;
  for (int a = 0; a < ndose; ++a) {
    const DDose &D = P.dose[is_one(kMode) ? 0 : a];
    if (is_tle(kMode) && D.tle && e_tle <= D.tle_thr) continue;
    float pos[3] = {g.p[0], g.p[1], g.p[2]};
    // A Geant4 step inside the nested-parameterised image never leaves its voxel: when the dose grid IS the image grid,
    // any point strictly between the pre- and the post-step point (hit_type random, middle) lies in the voxel of the
    // interaction, so the emulation below -- and the random number of "random" -- is not needed; "pre" keeps it (a
    // pre-step point ON a voxel face is the reference's own ambiguity)
    if (kMode != 2 && D.hit_type != GATE_B200_HIT_POST && !(D.aligned && D.hit_type != GATE_B200_HIT_PRE)) {
      // pre/random/middle between the Geant4 pre and post step points; the Geant4 step starts at
      // the voxel entry or at the previous interaction (GateDoseActor.cpp:144-178)
      // the image voxel of the interaction point, by the SAME expression that gathered its label in the Woodcock
      // call (a differently rounded index disagreed with the oracle for one deposit in 3.5e6 on the full CT)
      const int vi[3] = {__float2int_rd(__fmaf_rn(g.p[0], P.inv_sp[0], P.hs[0])), __float2int_rd(__fmaf_rn(g.p[1], P.inv_sp[1], P.hs[1])),
                         __float2int_rd(__fmaf_rn(g.p[2], P.inv_sp[2], P.hs[2]))};
      float back = g.seg;
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
      if (D.hit_type == GATE_B200_HIT_RANDOM) fr = 1.0f - rng_uniform(g.rng);
      else if (D.hit_type == GATE_B200_HIT_MIDDLE) fr = 0.5f;
      float bk = back * fr;
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) pos[ax] = g.p[ax] - g.d[ax] * bk;
    }
    float ci[3];
    aff_apply(D.v2i, pos, ci);
    // itk::ImageBase::TransformPhysicalPointToIndex: round half up, then region test
    float fx = floorf(ci[0] + 0.5f), fy = floorf(ci[1] + 0.5f), fz = floorf(ci[2] + 0.5f);
    if (fx < 0.0f || fy < 0.0f || fz < 0.0f || fx >= (float)D.size[0] || fy >= (float)D.size[1] ||
        fz >= (float)D.size[2])
      continue;
    int v = (int)fx + D.size[0] * ((int)fy + D.size[1] * (int)fz);
    int key = (int)(((unsigned)a << 28) | (unsigned)v);
    float dd = e / __ldg(P.dens + m);
    if (is_tle(kMode) && D.tle && (D.flags & (GATE_B200_SCORE_EDEP_SQ | GATE_B200_SCORE_DOSE_SQ))) {
      unsigned nl = tl_ld(&P.tle_state[gslot].w), nr = tl_ld(&P.tle_aux[gslot].y);
      if (kMode == 5) {
        // a point deposit is a record of its own: bounding box = its voxel, direction 0
        if (nr >= (unsigned)P.tle_rec_cap) { nl = 0; nr = 0; ds.tle_over++; }
        nr++;
        const int iv[3] = {v % D.size[0], (v / D.size[0]) % D.size[1], v / (D.size[0] * D.size[1])};
        float tmin = 0.0f;
        unsigned long long cand = (1ull << (nr - 1u)) - 1ull;   // every stored record
        tle_fold_add(P, D, gslot, nr, nl, tmin, cand, key, iv, -1, e, dd, 1, 0.0f, ds.tle_ent, ds.tle_over);
        float4 *rec = P.tle_rec + ((size_t)gslot * kTleRecMax + (nr - 1u)) * kTleRecW;
        tl_st(rec, make_float4(e, dd, __int_as_float(key), 0.0f));
        tl_st(rec + 1, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        tl_st(rec + 2, make_float4(__uint_as_float((unsigned)iv[0] | ((unsigned)iv[1] << 16)), __uint_as_float((unsigned)iv[2] | ((unsigned)iv[0] << 16)),
                                   __uint_as_float((unsigned)iv[1] | ((unsigned)iv[2] << 16)), 0.0f));
        tl_st(&P.tle_state[gslot].w, nl);
        tl_st(P.tle_aux + gslot, make_uint2(0u, nr));
      } else {
        const unsigned nl0 = nl, nr0 = nr, over0 = ds.tle_over;
        tle_record_begin_tab(P, gslot, nl, nr, ds.tle_ent, ds.tle_over);
        const unsigned over1 = ds.tle_over;
        tle_table_add(P, gslot, nl, key, e, dd, 1, ds.tle_ent, ds.tle_over);
        if (ds.tle_over != over1) nr = 1;   // the table was written out while this deposit was scored
        if (nl != nl0 || ds.tle_over != over0) tl_st(&P.tle_state[gslot].w, nl);
        if (nr != nr0) tl_st(&P.tle_aux[gslot].y, nr);
      }
      continue;
    }
    int j = 0;
    if (is_one(kMode) && kPre) {
      // entries 0 and 1 by their prefetched keys, the rest (rare) by the loop
      if (nbuf > 0 && pk.k0 == key) j = 0;
      else if (nbuf > 1 && pk.k1 == key) j = 1;
      else {
        for (j = nbuf < 2 ? nbuf : 2; j < nbuf; ++j)
          if (sc_ld(&buf[j].key) == key) break;
      }
    } else {
      for (j = 0; j < nbuf; ++j)
        if (sc_ld(&buf[j].key) == key) break;
    }
    DepEntry en;
    if (j == nbuf) {
      if (nbuf == kDepBuf) {   // coalescing buffer full: write it out, start again
        flush_buffer<kMode>(P, buf, nbuf);
        ds.flush += (unsigned)nbuf;
        ds.over++;
        nbuf = 0;
        j = 0;
      }
      en.key = key; en.e = e; en.d = dd; en.n = 1;
      nbuf++;
    } else {
      en = ld_dep(buf + j);
      en.e += e; en.d += dd; en.n += 1;
    }
    if (is_one(kMode) && kKeep) { keep.en = en; keep.j = j; }
    else st_dep(buf + j, en);
  }
  g.flags = (g.flags & ~F_NBUF) | ((unsigned)nbuf << F_NBUF_SHIFT);
}

template <int kMode>
__device__ __forceinline__ void deposit(const Params &P, Ph &g, DepEntry *buf, DepStat &ds, float e,
                                     int m, bool in_vol, float e_tle, int gslot) {
  DepKeys pk; pk.k0 = -1; pk.k1 = -1;
  DepKeep keep;
  deposit<kMode, false, false>(P, g, buf, ds, e, m, in_vol, e_tle, gslot, pk, keep);
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

// Process selection from a variate v uniform in (0,1) (oracle select_process): inside the volume
// v is the rescaled Woodcock acceptance variate, in the world the second uniform of the flight
// block.  Sets the routing flags and returns the queue of the phase that carries the process out.
__device__ __forceinline__ int route_interaction(const Params &P, unsigned &flags, int m, int ie, float fe, float v) {
  const float4 pa = __ldg(P.prob + (size_t)ie * P.nmat + m);
  const float4 pb = __ldg(P.prob + (size_t)(ie + 1) * P.nmat + m);
  const float pph = __fmaf_rn(fe, pb.x - pa.x, pa.x);
  const float pco = __fmaf_rn(fe, pb.y - pa.y, pa.y);
  const float pra = __fmaf_rn(fe, pb.z - pa.z, pa.z);
  flags &= ~(F_ABSORB | F_CONV);
  if (v < pph) { flags |= F_ABSORB; return Q_E; }                  // photoelectric: kerma, dies
  if (v < pco) return Q_C;
  if (v < pra) return Q_R;
  flags |= F_CONV;
  if (P.pair_mode == 0) { flags |= F_ABSORB; return Q_E; }         // conversion, everything local
  return Q_R;
}

// ------------------------------------------------------------------------------ primaries

__device__ __forceinline__ float gauss(Rng &r) {
  float u1 = rng_uniform(r), u2 = rng_uniform(r);
  float s, c;
  d_sincos2pi(u2, s, c);
  return sqrtf(-2.0f * d_logf(u1)) * c;
}

// first i in [0, last] with p < cdf[i], `last` if there is none: the index the reference's linear
// scan returns (GateSPSEneDistribution::IndexForProbability) for a non-decreasing cdf
__device__ __forceinline__ int cdf_index(const float *__restrict__ cdf, int last, float p) {
  int lo = 0, hi = last;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (p >= __ldg(cdf + mid)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// position of a primary in the frame of its source (the position samplers of GateGenericSource::InitializePosition /
// G4SPSPosDistribution, the voxel sampler, the entry of a phase-space file)
__device__ __forceinline__ void sample_position(const Params &P, const DSource &S, Rng &r, float l[3], float4 &ea, float4 &eb, int gslot) {
  l[0] = 0.0f; l[1] = 0.0f; l[2] = 0.0f;
  switch (S.pos_type) {
  case GATE_B200_POS_PHSP: {
    // GatePhaseSpaceSource::GenerateOnePrimary (GatePhaseSpaceSource.cpp:137-166): sequential entries,
    // wrapping at the end of the file (phspsources.py:119-146)
    const unsigned long long id = ((unsigned long long)r.h1 << 32) | r.h0;
    const unsigned long long en = (S.ph_start + id) % S.ph_n;
    ea = __ldg(S.ph_a + en);
    eb = __ldg(S.ph_b + en);
    l[0] = ea.x; l[1] = ea.y; l[2] = ea.z;
    sc_st(P.slot_w + gslot, eb.w * S.weight);
  } break;
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
  case GATE_B200_POS_SURFACE_SPHERE: {
    // G4SPSPosDistribution::GeneratePointsOnSurface, Sphere: theta = acos(1 - 2u), phi = 2 pi u'
    const float cost = 1.0f - 2.0f * rng_uniform(r);
    const float s2 = 1.0f - cost * cost;
    const float sint = sqrtf(s2 > 0.0f ? s2 : 0.0f);
    float sp, cp;
    d_sincos2pi(rng_uniform(r), sp, cp);
    l[0] = S.radius * sint * cp; l[1] = S.radius * sint * sp; l[2] = S.radius * cost;
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
}

// kMode: the single-dose-actor kernel variants (2-4) are launched for sources without position.confine and without
// angular histograms only and carry neither piece of code (measured: the presence of the confine loop alone cost
// config 3 0.8 %)
template <int kMode>
__device__ __forceinline__ void generate_primary(const Params &P, Ph &g, int gslot) {
  const DSource &S = P.src;
  Rng &r = g.rng;
  rng_begin(r);
  float l[3];
  float4 ea = make_float4(0.0f, 0.0f, 0.0f, 0.0f), eb = ea;
  sample_position(P, S, r, l, ea, eb, gslot);
  aff_apply(S.pos, l, g.p);
  if (!is_one(kMode) && S.confine) {
    // position.confine (G4SPSPosDistribution::GenerateOne with ConfineSourceToVolume): draw again until the point
    // lies in the confining box, at most 100000 draws in all
    for (int tries = 1; tries < 100000; ++tries) {
      float qc[3];
      aff_apply(S.conf, g.p, qc);
      if (inside_box(qc, S.conf_half)) break;
      sample_position(P, S, r, l, ea, eb, gslot);
      aff_apply(S.pos, l, g.p);
    }
  }

  switch (S.dir_type) {
  case GATE_B200_DIR_ISO: {
    // G4SPSAngDistribution::GenerateIsotropicFlux: momentum = -(sin t cos p, sin t sin p, cos t)
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
  case GATE_B200_DIR_COS: {
    // G4SPSAngDistribution::GenerateCosineLawFlux: sin^2(theta) uniform between sin^2(min) and sin^2(max)
    const float st2 = rng_uniform(r) * (S.sin2_tmax - S.sin2_tmin) + S.sin2_tmin;
    const float sint = sqrtf(st2 > 0.0f ? st2 : 0.0f);
    const float c2 = 1.0f - st2;
    const float cost = sqrtf(c2 > 0.0f ? c2 : 0.0f);
    float turn = S.phi0_turn + S.dphi_turn * rng_uniform(r);
    turn = turn - floorf(turn);
    float sp, cp;
    d_sincos2pi(turn, sp, cp);
    float v[3] = {-sint * cp, -sint * sp, -cost};
    if (S.dir_relative) rot_apply(S.dir_rot, v, g.d);
    else { g.d[0] = v[0]; g.d[1] = v[1]; g.d[2] = v[2]; }
  } break;
  case GATE_B200_DIR_BEAM2D: {
    // G4SPSAngDistribution::GenerateBeamFlux (beam2d): px, py ~ N(0, sigma_x/y), theta = |(px, py)|,
    // phi = atan2(py, px); momentum = -(sin t cos p, sin t sin p, cos t).  cos p = px / theta and
    // sin p = py / theta, so no inverse trigonometric function is needed.
    const float px = gauss(r) * S.dir_sx;
    const float py = gauss(r) * S.dir_sy;
    const float th = sqrtf(px * px + py * py);
    float v[3] = {0.0f, 0.0f, -1.0f};
    if (th != 0.0f) {
      float turn = th * 0.15915494309189535f;
      turn = turn - floorf(turn);
      float st, ct;
      d_sincos2pi(turn, st, ct);
      const float k = st / th;
      v[0] = -(k * px); v[1] = -(k * py); v[2] = -ct;
    }
    if (S.dir_relative) rot_apply(S.dir_rot, v, g.d);
    else { g.d[0] = v[0]; g.d[1] = v[1]; g.d[2] = v[2]; }
  } break;
  case GATE_B200_DIR_HISTOGRAM: if (!is_one(kMode)) {
    // G4SPSAngDistribution::GenerateUserDefFlux, theta and phi from user histograms: the angle is the inverse of the
    // piecewise-linear cumulative at a uniform variate (uniform inside its bin), momentum = -(sin t cos p, sin t sin p, cos t)
    float ang[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float *ce = k == 0 ? S.ht_e : S.hp_e;
      const float *cc = k == 0 ? S.ht_c : S.hp_c;
      const int nb = k == 0 ? S.n_ht : S.n_hp;
      const float u = rng_uniform(r);
      const int b = cdf_index(cc + 1, nb - 1, u);
      const float c0 = __ldg(cc + b), c1 = __ldg(cc + b + 1), e0 = __ldg(ce + b), e1 = __ldg(ce + b + 1);
      const float den = c1 - c0;
      float f = den > 0.0f ? (u - c0) / den : 0.0f;
      if (f > 1.0f) f = 1.0f;
      ang[k] = e0 + f * (e1 - e0);
    }
    float tt = ang[0] * 0.15915494309189535f, tp = ang[1] * 0.15915494309189535f;
    tt = tt - floorf(tt); tp = tp - floorf(tp);
    float st, ct, sp, cp;
    d_sincos2pi(tt, st, ct);
    d_sincos2pi(tp, sp, cp);
    float v[3] = {-st * cp, -st * sp, -ct};
    if (S.dir_relative) rot_apply(S.dir_rot, v, g.d);
    else { g.d[0] = v[0]; g.d[1] = v[1]; g.d[2] = v[2]; }
  } break;
  case GATE_B200_DIR_PHSP: {
    // direction / direction.mag(), then the rotation of the attached volume (GatePhaseSpaceSource.cpp:158-162)
    float inv = 1.0f / sqrtf(eb.x * eb.x + eb.y * eb.y + eb.z * eb.z);
    float v[3] = {eb.x * inv, eb.y * inv, eb.z * inv};
    if (S.dir_relative) rot_apply(S.dir_rot, v, g.d);
    else { g.d[0] = v[0]; g.d[1] = v[1]; g.d[2] = v[2]; }
  } break;
  }

  float E = 0.0f;
  switch (S.ene_type) {
  case GATE_B200_ENE_PHSP: E = ea.w; break;
  case GATE_B200_ENE_MONO: E = S.mono; break;
  case GATE_B200_ENE_GAUSS: E = S.mono + gauss(r) * S.sigma; break;
  case GATE_B200_ENE_RANGE: E = S.emin + rng_uniform(r) * (S.emax - S.emin); break;
  case GATE_B200_ENE_SPECTRUM_DISCRETE: {
    int i = cdf_index(S.spec_cdf, S.n_spec - 1, rng_uniform(r));
    E = __ldg(S.spec_e + i);
  } break;
  case GATE_B200_ENE_SPECTRUM_HISTOGRAM: {
    int i = cdf_index(S.spec_cdf, S.n_spec - 1, rng_uniform(r));
    float e0 = __ldg(S.spec_e + i), e1 = __ldg(S.spec_e + i + 1);
    E = e0 + rng_uniform(r) * (e1 - e0);
  } break;
  case GATE_B200_ENE_SPECTRUM_HISTOGRAM_LINEAR: {
    // GateSPSEneDistribution.cpp:124-147
    int i = cdf_index(S.spec_cdf, S.n_spec - 2, rng_uniform(r));
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
    // GateSPSEneDistribution.cpp:38-49
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
}

// ------------------------------------------------------------------------------ phase space

template <int kMode>
__device__ __forceinline__ void phsp_store(const Params &P, Ph &g, unsigned long long hist, int k, int flag, int gslot) {
  const DPhsp &Q = P.phsp[k];
  g.flags |= F_STORED;
  unsigned long long idx = atomicAdd(Q.n_rows, 1ull);
  if (idx >= Q.capacity) {
    atomicAdd(Q.n_dropped, 1ull);
    return;
  }
  Q.ke[idx] = g.E;
#pragma unroll
  for (int a = 0; a < 3; ++a) { Q.pos[3 * idx + a] = g.p[a]; Q.dir[3 * idx + a] = g.d[a]; }
  Q.w[idx] = hist_weight<kMode>(P, gslot);
  Q.t[idx] = g.path / kCmmPerNs;
  Q.ev[idx] = (long long)hist;
  Q.fl[idx] = flag;
}

// ------------------------------------------------------------------------------ regions

__device__ __forceinline__ void to_volume(const Params &P, Ph &g) {
  float pl[3], dl[3];
  aff_apply(P.w2v, g.p, pl);
  rot_apply(P.w2v.m, g.d, dl);
#pragma unroll
  for (int a = 0; a < 3; ++a) { g.p[a] = pl[a]; g.d[a] = dl[a]; }
}
__device__ __forceinline__ void to_world(const Params &P, Ph &g) {
  float pw[3], dw[3];
  aff_apply_inv(P.w2v, g.p, pw);
  rot_apply_inv(P.w2v.m, g.d, dw);
#pragma unroll
  for (int a = 0; a < 3; ++a) { g.p[a] = pw[a]; g.d[a] = dw[a]; }
}

// place a photon given in the world frame into its region; returns the queue of its next phase
template <int kMode>
__device__ __forceinline__ int locate(const Params &P, Ph &g) {
  int st = Q_W;
  if (P.has_vol) {
    float pl[3];
    aff_apply(P.w2v, g.p, pl);
    if (inside_box(pl, P.half)) { to_volume(P, g); st = Q_V; }
  }
  g.flags &= ~F_INSIDE;
  for (int k = 0; k < ((kMode == 2 || kMode == 3) ? 0 : P.nphsp); ++k) {
    float pb[3];
    aff_apply(P.phsp[k].w2b, g.p, pb);
    if (st == Q_W && inside_box(pb, P.phsp[k].half)) g.flags |= (1u << k);
  }
  return st;
}

// ------------------------------------------------------------------------------ phases

// Q_W: one flight segment of a photon travelling in the world material (outside the image volume);
// a photon that has just left the volume is first taken to its exit point and to the world frame.
template <int kMode>
__device__ __forceinline__ int phase_world(const Params &P, Smem &S, int slot, int gslot) {
  Ph g;
  load_ph(P, S, slot, g);
  int next;
  unsigned n_esc = 0, n_steps = 0;
  if (g.flags & F_EXIT) {
    float ti, to;
    int ai, ao;
    if (!ray_box_ax(g.p, g.d, P.half, ti, to, ai, ao)) { to = 0.0f; ao = -1; }
#pragma unroll
    for (int a = 0; a < 3; ++a) g.p[a] = g.p[a] + g.d[a] * to;
    snap_to_face(g.p, g.d, P.half, ao, true);
    g.path += to;
    n_steps++;
    to_world(P, g);
    g.flags &= ~F_EXIT;
  }
  float t_exit, t0;
  if (g.rng.blk > kMaxBlocks) {
    next = Q_E;
  } else if (!inside_box(g.p, P.wh) || !ray_box(g.p, g.d, P.wh, t0, t_exit)) {
    n_esc = 1;
    next = Q_E;
  } else {
    int ev = 0;   // 0 interact, 1 leave world, 2 enter volume, 3+2k enter plane k, 4+2k leave plane k, 11 cross the face of a Box region
    // the medium of this flight: the Box region the photon is in, else the world material
    const int nreg = is_one(kMode) ? 0 : P.nreg;
    int cur = -1, mat_here = P.world_mat;
    for (int k = 0; k < nreg; ++k) {
      float pb[3];
      aff_apply(P.reg[k].w2b, g.p, pb);
      if (cur < 0 && inside_box(pb, P.reg[k].half)) { cur = k; mat_here = P.reg[k].mat; }
    }
    float mu = tab_mu(P, mat_here, g.ie, g.fe);
    float t_best = -d_logf(rng_uniform(g.rng)) / mu;
    if (t_exit < t_best) { t_best = t_exit; ev = 1; }

    int ax_in = -1;
    if (P.has_vol) {
      float pl[3], dl[3], ti, to;
      int ai, ao;
      aff_apply(P.w2v, g.p, pl);
      rot_apply(P.w2v.m, g.d, dl);
      if (ray_box_ax(pl, dl, P.half, ti, to, ai, ao) && ti < t_best) { t_best = ti; ev = 2; ax_in = ai; }
    }
    for (int k = 0; k < ((kMode == 2 || kMode == 3) ? 0 : P.nphsp); ++k) {
      float pb[3], db[3], ti, to;
      aff_apply(P.phsp[k].w2b, g.p, pb);
      rot_apply(P.phsp[k].w2b.m, g.d, db);
      if (!ray_box(pb, db, P.phsp[k].half, ti, to)) {
        g.flags &= ~(1u << k);
        continue;
      }
      if (g.flags & (1u << k)) {
        if (to < t_best) { t_best = to; ev = 4 + 2 * k; }
      } else if (ti < t_best) { t_best = ti; ev = 3 + 2 * k; }
    }
    // faces of the Box regions, after the planes: a Box that is both (a detector slab with a phase-space actor) reports
    // the crossing as the plane's event, the medium follows from the position at the next call
    for (int k = 0; k < nreg; ++k) {
      float pb[3], db[3], ti, to;
      aff_apply(P.reg[k].w2b, g.p, pb);
      rot_apply(P.reg[k].w2b.m, g.d, db);
      if (!ray_box(pb, db, P.reg[k].half, ti, to)) continue;
      if (k == cur) {
        if (to < t_best) { t_best = to; ev = 11; }
      } else if (ti < t_best) { t_best = ti; ev = 11; }
    }
    float step = (ev >= 3) ? t_best + kNudge : t_best;
    if (ev == 1) {
      n_esc = 1;
      g.path += t_best;
      next = Q_E;
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) g.p[a] = g.p[a] + g.d[a] * step;
      g.path += step;
      if (ev == 0) {
        // real interaction in the world material (or in the material of the Box region the photon is in)
        g.flags &= ~F_INVOL;
        g.mat = mat_here;
        next = route_interaction(P, g.flags, g.mat, g.ie, g.fe, rng_uniform(g.rng));
      } else if (ev == 11) {
        n_steps++;
        next = Q_W;       // the next flight starts on the other side of the face, in the other medium
      } else if (ev == 2) {
        n_steps++;
        to_volume(P, g);
        snap_to_face(g.p, g.d, P.half, ax_in, false);
        g.seg = 0.0f;
        next = Q_V;
      } else {
        const int k = (ev - 3) >> 1;
        const unsigned long long hist = ((unsigned long long)g.rng.h1 << 32) | g.rng.h0;
        n_steps++;
        if (((ev - 3) & 1) == 0) {
          g.flags |= (1u << k);
          if (P.phsp[k].steps & GATE_B200_PHSP_ENTERING) phsp_store<kMode>(P, g, hist, k, GATE_B200_PHSP_ENTERING, gslot);
        } else {
          g.flags &= ~(1u << k);
          if (P.phsp[k].steps & GATE_B200_PHSP_EXITING) phsp_store<kMode>(P, g, hist, k, GATE_B200_PHSP_EXITING, gslot);
        }
        next = Q_W;
      }
    }
  }
  store_ph(S, slot, g);
  {
    const unsigned t = stat_reduce_packed(n_esc | (n_steps << 8));
    stat_bump(S, S_ESC, t & 0xffu);
    stat_bump(S, S_STEPS, t >> 8);
  }
  return next;
}

// Q_V: kWoodcockK tentative collisions of Woodcock delta tracking against the majorant of the
// materials present in the volume.  The flight distances, positions, label gathers and attenuation
// look-ups of all kWoodcockK iterations are issued up front (they only depend on the random numbers,
// not on the outcome of the earlier iterations), so the dependent global loads of the iterations
// overlap; the accept / leave decisions are then taken in order.  Arithmetic and order of the
// floating-point operations are those of the sequential statement in the oracle.
// The call is split into three stages (gather labels / gather attenuation / decide) so that one lane
// can run them for kVWidth slots side by side: kVWidth * kWoodcockK independent gathers in flight per
// thread, which is what hides the DRAM/L2 latency at 32 warps per SM.
constexpr unsigned kShellHit = 1u << 8;    // VCall::inb: the call ended with an interaction in the shell
constexpr unsigned kShellLeft = 1u << 9;   //             the photon leaves the volume through the shell
// a slot moves along its ray by t (shell crossing at the start of a Woodcock call)
__device__ __forceinline__ void S_put_point(Smem &S, int slot, float x, float y, float z, float t) {
  S.px[slot] = x; S.py[slot] = y; S.pz[slot] = z;
  S.path[slot] = S.path[slot] + t;
  S.seg[slot] = S.seg[slot] + t;
}
struct VCall {
  float s[kWoodcockK];          // flight lengths
  uint32_t racc[kWoodcockK];    // acceptance variates (raw Philox words)
  int lab[kWoodcockK];          // labels at the tentative collision points
  float mu[kWoodcockK];
  unsigned inb;                 // bit k: point k is inside the volume
  unsigned blk;
  int ie;
  float fe, maj;
};

// kShell: 0 = the call starts inside the inner box (no shell code at all), 1 = it starts in the shell,
//         2 = decided at run time (kernels without the separate queue)
template <typename LabelT, int kShell = 2>
__device__ __forceinline__ void volume_gather_labels(const Params &P, Smem &S, const float *__restrict__ s_maj,
                                                     int slot, VCall &c) {
  c.blk = S.blk[slot];
  float a0 = S.px[slot], a1 = S.py[slot], a2 = S.pz[slot];
  const float d0 = S.dx[slot], d1 = S.dy[slot], d2 = S.dz[slot];
  bin_split(P, S.tbin[slot], c.ie, c.fe);
  const float m0 = s_maj[c.ie], m1 = s_maj[c.ie + 1];
  c.maj = __fmaf_rn(c.fe, m1 - m0, m0);
  const float inv_maj = 1.0f / c.maj;
  c.inb = 0u;
  // ---- a call that starts in the homogeneous shell (outside the inner box): geometry of the way to the box (or,
  // if the ray misses it, to the face of the volume); evaluated before the random numbers exist to keep it cheap
  float sh_mu = 0.0f, sh_tb = 0.0f, sh_lim = -1.0f;   // sh_lim < 0: not in the shell
  int sh_ax = -1;
  bool sh_hit = false;
  if (kShell != 0 && (kShell == 1 || P.shell)) {
    const float ps[3] = {a0, a1, a2}, dd[3] = {d0, d1, d2};
    if (!inside_tbox(P, ps)) {
      sh_mu = tab_mu(P, P.shell_mat, c.ie, c.fe);
      // slab tests against the inner box (entry) and the volume (exit) with one reciprocal per axis
      float tmin = -kBig, tmax = kBig, tout = kBig;
      int ain = -1;
      bool miss = false;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float q = ps[a] - P.tc[a];
        if (dd[a] != 0.0f) {
          const float inv = 1.0f / dd[a];
          float t1 = (-P.th[a] - q) * inv, t2 = (P.th[a] - q) * inv;
          if (t1 > t2) { const float w = t1; t1 = t2; t2 = w; }
          if (t1 > tmin) { tmin = t1; ain = a; }
          if (t2 < tmax) tmax = t2;
          const float v1 = (-P.half[a] - ps[a]) * inv, v2 = (P.half[a] - ps[a]) * inv;
          const float vm = v1 > v2 ? v1 : v2;
          if (vm < tout) tout = vm;
        } else if (q < -P.th[a] || q > P.th[a]) {
          miss = true;
        }
      }
      const float t0 = tmin > 0.0f ? tmin : 0.0f;
      sh_hit = !miss && tmax > t0;
      sh_tb = t0;
      sh_ax = tmin > 0.0f ? ain : -1;
      sh_lim = sh_hit ? t0 : (tout > 0.0f ? tout : 0.0f);
    }
  }
  uint32_t r[2 * kWoodcockK];
  {
    const unsigned long long id = P.first + S.hidx[slot];
    philox_block2((uint32_t)P.seed, (uint32_t)(P.seed >> 32), (uint32_t)id, (uint32_t)(id >> 32), c.blk, r);
  }
  const LabelT *__restrict__ labels = reinterpret_cast<const LabelT *>(P.labels);
  float tau0 = -d_logf(u32_to_uniform(r[0]));   // optical depth to the first event of the call
  if (kShell != 0 && sh_lim >= 0.0f) {
    // The same optical depth first runs down at the attenuation of the shell material: either it is used up before
    // the box / the volume face -- a real interaction in the shell, no rejection needed --, or the photon reaches
    // the box with the remainder, or it leaves the volume.
    if (tau0 < sh_mu * sh_lim) {
      const float ts = tau0 / sh_mu;
      S_put_point(S, slot, a0 + d0 * ts, a1 + d1 * ts, a2 + d2 * ts, ts);
      c.inb |= kShellHit;
    } else if (!sh_hit) {
      c.inb |= kShellLeft;
    } else {
      float q[3] = {a0 + d0 * sh_tb, a1 + d1 * sh_tb, a2 + d2 * sh_tb};
      const float dd[3] = {d0, d1, d2};
      snap_to_tface(P, q, dd, sh_ax, false);
      S_put_point(S, slot, q[0], q[1], q[2], sh_tb);
      a0 = q[0]; a1 = q[1]; a2 = q[2];
      tau0 = __fmaf_rn(-sh_mu, sh_tb, tau0);
    }
  }
#pragma unroll
  for (int k = 0; k < kWoodcockK; ++k) {
    c.s[k] = (k == 0 ? tau0 : -d_logf(u32_to_uniform(r[2 * k]))) * inv_maj;
    c.racc[k] = r[2 * k + 1];
    a0 = __fmaf_rn(d0, c.s[k], a0); a1 = __fmaf_rn(d1, c.s[k], a1); a2 = __fmaf_rn(d2, c.s[k], a2);
    // continuous voxel coordinate: a / spacing + n / 2 in one fused operation
    const float fx = __fmaf_rn(a0, P.inv_sp[0], P.hs[0]);
    const float fy = __fmaf_rn(a1, P.inv_sp[1], P.hs[1]);
    const float fz = __fmaf_rn(a2, P.inv_sp[2], P.hs[2]);
    // 0 <= f < n  <=>  0 <= floor(f) < n; inside, floor(f) is the voxel index (int)f of the oracle
    const int ix = __float2int_rd(fx), iy = __float2int_rd(fy), iz = __float2int_rd(fz);
    const bool in = (unsigned)ix < (unsigned)P.n[0] && (unsigned)iy < (unsigned)P.n[1] && (unsigned)iz < (unsigned)P.n[2];
    unsigned flat = 0;   // the ABI bounds the volume to < 2^31 voxels
    if (in) {
      flat = (unsigned)ix + (unsigned)P.n[0] * ((unsigned)iy + (unsigned)P.n[1] * (unsigned)iz);
      c.inb |= 1u << k;
    }
#if GB2_CLASSVOL
    c.lab[k] = (int)__ldg(P.cls + (flat >> 2));   // the byte of four class codes
    c.mu[k] = __uint_as_float(flat);              // parked until volume_gather_mu needs it
#else
    c.lab[k] = (int)__ldg(labels + flat);
#endif
  }
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l1(const void *p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
template <typename LabelT>
__device__ __forceinline__ void volume_gather_mu(const Params &P, VCall &c) {
  const float2 *__restrict__ murow = P.mu2 + (size_t)c.ie * P.nmat;
#if GB2_CLASSVOL
  {
    // class bounds at the photon's energy, a hair above the interpolated maxima (the interpolation of the maxima
    // bounds the interpolated coefficient of every material of the class up to rounding: 1e-6 covers it)
    const float4 q0 = __ldg(P.cmax + c.ie), q1 = __ldg(P.cmax + c.ie + 1);
    const float cm0 = __fmaf_rn(c.fe, q1.x - q0.x, q0.x) * 1.000001f, cm1 = __fmaf_rn(c.fe, q1.y - q0.y, q0.y) * 1.000001f;
    const float cm2 = __fmaf_rn(c.fe, q1.z - q0.z, q0.z) * 1.000001f, cm3 = __fmaf_rn(c.fe, q1.w - q0.w, q0.w) * 1.000001f;
    const LabelT *__restrict__ labels = reinterpret_cast<const LabelT *>(P.labels);
#pragma unroll
    for (int k = 0; k < kWoodcockK; ++k) {
      const unsigned flat = __float_as_uint(c.mu[k]);
      const int cl = (c.lab[k] >> ((flat & 3u) * 2u)) & 3;
      const float bound = cl == 0 ? cm0 : (cl == 1 ? cm1 : (cl == 2 ? cm2 : cm3));
      const float t = u32_to_uniform(c.racc[k]) * c.maj;          // the acceptance variate of volume_decide
      const bool need = ((c.inb >> k) & 1u) && t <= bound;
      int lab = -1;
      if (need) {
        if (P.cls_exact) lab = cl == 0 ? P.cls_mat[0] : (cl == 1 ? P.cls_mat[1] : (cl == 2 ? P.cls_mat[2] : P.cls_mat[3]));
        else lab = (int)__ldg(labels + flat);
      }
      c.lab[k] = lab;
    }
#pragma unroll
    for (int k = 0; k < kWoodcockK; ++k) {
      float mu = 0.0f;   // a sure rejection: t <= 0 never holds
      if (c.lab[k] >= 0) {
        const float2 v = __ldg(murow + c.lab[k]);
        mu = __fmaf_rn(c.fe, v.y - v.x, v.x);
      }
      c.mu[k] = mu;
    }
  }
#else
#pragma unroll
  for (int k = 0; k < kWoodcockK; ++k) {
    const float2 v = __ldg(murow + c.lab[k]);
#if GB2_PREFETCH_PROB
    // the process probabilities of this material (needed only if the collision is accepted: a third dependent
    // look-up after label and attenuation) are asked into L1 now, without a register or a wait
    if (k < GB2_PREFETCH_PROB) {
      prefetch_l1(P.prob + (size_t)c.ie * P.nmat + c.lab[k]);
      prefetch_l1(P.prob + (size_t)(c.ie + 1) * P.nmat + c.lab[k]);
    }
#endif
    c.mu[k] = __fmaf_rn(c.fe, v.y - v.x, v.x);
  }
#endif
}

template <int kMode>
__device__ __forceinline__ int volume_decide(const Params &P, Smem &S, int slot, int gslot, const VCall &c, unsigned &c_tent) {
  if (c.blk > kMaxBlocks) return Q_E;
  S.blk[slot] = c.blk + 2u;
  float p0 = S.px[slot], p1 = S.py[slot], p2 = S.pz[slot];
  const float d0 = S.dx[slot], d1 = S.dy[slot], d2 = S.dz[slot];
  float path = S.path[slot], seg = S.seg[slot];
  int mat = 0;
  float vsel = 0.0f;
  unsigned tent = 0;
  bool left = (c.inb & kShellLeft) != 0, hit = (c.inb & kShellHit) != 0;
  if (hit) {                  // interaction in the shell material: the acceptance variate of the first flight selects the process
    mat = P.shell_mat;
    vsel = u32_to_uniform(c.racc[0]);
  }
#pragma unroll
  for (int k = 0; k < kWoodcockK; ++k) {
    if (hit || left) continue;
    if (!(c.inb & (1u << k))) { left = true; continue; }
    // the same sequence of operations that produced the gather address
    p0 = __fmaf_rn(d0, c.s[k], p0); p1 = __fmaf_rn(d1, c.s[k], p1); p2 = __fmaf_rn(d2, c.s[k], p2);
    path += c.s[k];
    seg += c.s[k];
    tent++;
    const float t = u32_to_uniform(c.racc[k]) * c.maj;
    if (t <= c.mu[k]) {       // accepted: a real interaction in material lab[k]
      mat = c.lab[k];
      vsel = t / c.mu[k];
      hit = true;
    }
  }
  c_tent += tent;
  S.px[slot] = p0; S.py[slot] = p1; S.pz[slot] = p2;
  S.path[slot] = path; S.seg[slot] = seg;
  int next = Q_V;
  if (left) {
    S.fm[slot] |= F_EXIT;   // taken to the exit point and to the world frame by the world phase
    next = Q_W;
  } else if (hit) {
    unsigned flags = (S.fm[slot] & F_MASK) | F_INVOL;
    next = route_interaction(P, flags, mat, c.ie, c.fe, vsel);
    S.fm[slot] = flags | ((unsigned)mat << F_MAT_SHIFT);
  }
  if (is_tle(kMode) && next != Q_V && S.E[slot] <= P.tle_thr_max) {
    // the flight has ended (interaction or exit): score its voxel chords first, then carry on
    S.fm[slot] |= (unsigned)(next - 1) << F_NEXT_SHIFT;
    tl_st(&P.tle_state[gslot].x, __float_as_uint(-1.0f));
    next = Q_T;
  }
  return next;
}

// Q_C: one Compton trial (Klein-Nishina composition-rejection x scattering function).  A rejected
// trial simply stays in the queue: trials are independent, each call owns two fresh Philox blocks
// (trial, final state), so the lanes of a warp never wait for each other's rejection loops.
template <int kMode>
__device__ __forceinline__ int phase_compton(const Params &P, Smem &S, int slot, int gslot, double &c_edep) {
  const unsigned b0 = S.blk[slot];
  S.blk[slot] = b0 + 2u;
  const float E = S.E[slot];
  const unsigned fm = S.fm[slot];
  const int m = (int)(fm >> F_MAT_SHIFT);
#if GB2_PREFETCH_KEYS
  DepKeys pk;
  {
    const DepEntry *pb = P.depbuf + (size_t)gslot * kDepBuf;
    const int nb = (int)((fm & F_NBUF) >> F_NBUF_SHIFT);
    pk.k0 = nb > 0 ? sc_ld(&pb[0].key) : -1;
    pk.k1 = nb > 1 ? sc_ld(&pb[1].key) : -1;
  }
#endif
  uint32_t r[8];
  const unsigned long long id = P.first + S.hidx[slot];
#if GB2_COMPTON_LAZY
  // the trial needs the first block only; the second one (azimuth, hit-point draws) is generated when the trial is
  // accepted (same blocks, same numbers: the protocol does not change)
  philox_block((uint32_t)P.seed, (uint32_t)(P.seed >> 32), (uint32_t)id, (uint32_t)(id >> 32), b0, 0u, r[0], r[1], r[2], r[3]);
#else
  philox_block2((uint32_t)P.seed, (uint32_t)(P.seed >> 32), (uint32_t)id, (uint32_t)(id >> 32), b0, r);
#endif
  const float k = E / kMc2;
  const float eps0 = 1.0f / (1.0f + 2.0f * k);
  const float eps0sq = eps0 * eps0;
  const float alpha1 = -d_logf(eps0);
  const float alpha2 = alpha1 + 0.5f * (1.0f - eps0sq);
  const float kq = kKPerMeV * E;
  const float two_kq2 = 2.0f * kq * kq;
  const float u1 = u32_to_uniform(r[0]), u2 = u32_to_uniform(r[1]), u3 = u32_to_uniform(r[2]), u4 = u32_to_uniform(r[3]);
  float eps, epssq;
  if (alpha1 > alpha2 * u1) {
    eps = d_expf(-alpha1 * u2);
    epssq = eps * eps;
  } else {
    epssq = eps0sq + (1.0f - eps0sq) * u2;
    eps = sqrtf(epssq);
  }
  const float onecost = (1.0f - eps) / (eps * k);
  const float sint2 = onecost * (2.0f - onecost);
  const float grej = 1.0f - eps * sint2 / (1.0f + epssq);
  bool ok = !(grej < u3);
  const float sfv = interp_q2(P, P.sf + (size_t)m * P.nq, two_kq2 * onecost);
  if (u4 > sfv) ok = false;
  int next = Q_C;
  unsigned n_compt = 0;
  DepStat ds; ds.dep = 0; ds.over = 0; ds.flush = 0; ds.tle_ent = 0; ds.tle_over = 0; ds.edep = 0.0;
  if (ok) {
    Ph g;
    load_ph(P, S, slot, g);
#if GB2_COMPTON_LAZY
    philox_block((uint32_t)P.seed, (uint32_t)(P.seed >> 32), (uint32_t)id, (uint32_t)(id >> 32), b0 + 1u, 0u, r[4], r[5], r[6], r[7]);
#endif
    // the rest of the call's random numbers: r[4] is the azimuth, deposits continue from r[5]
    g.rng.blk = b0 + 2u; g.rng.pos = 1;
    g.rng.o0 = r[5]; g.rng.o1 = r[6]; g.rng.o2 = r[7];
    const bool in_vol = (g.flags & F_INVOL) != 0;
    DepEntry *buf = P.depbuf + (size_t)gslot * kDepBuf;
    n_compt = 1;
    const float cost = 1.0f - onecost;
    const float sint = sqrtf(sint2 > 0.0f ? sint2 : 0.0f);
    float sphi, cphi;
    d_sincos2pi(u32_to_uniform(r[4]), sphi, cphi);
    const float E1 = eps * E;
    const float wgt = hist_weight<kMode>(P, gslot);
#if GB2_PREFETCH_KEYS
    {
      DepKeep nokeep;
      deposit<kMode, true, false>(P, g, buf, ds, (E - E1) * wgt, m, in_vol, E, gslot, pk, nokeep);
    }
#else
    deposit<kMode>(P, g, buf, ds, (E - E1) * wgt, m, in_vol, E, gslot);
#endif
    rotate_dir(g.d, cost, sint, sphi, cphi);
    g.E = E1;
    g.seg = 0.0f;
    if (E1 < P.e_cut || (is_tle(kMode) && in_vol && E1 <= P.tle_emin)) {
      deposit<kMode>(P, g, buf, ds, E1 * wgt, m, in_vol, kBig, gslot);
      next = Q_E;
    } else {
      set_energy(P, S, slot, g);
      next = in_vol ? Q_V : Q_W;
    }
    store_ph(S, slot, g);
    c_edep += ds.edep;
  }
  {
    const unsigned t = stat_reduce_packed(n_compt | (ds.dep << 8));
    stat_bump(S, S_COMPT, t & 0xffu);
    stat_bump(S, S_STEPS, t & 0xffu);
    stat_bump(S, S_DEP, t >> 8);
  }
  stat_overflow(S, ds.over, ds.flush);
  if (is_tle(kMode)) stat_tle_deposits<kMode>(S, ds.tle_over, ds.tle_ent);
  return next;
}

// Q_R: the rare interactions.  Rayleigh: q2 from the cumulative squared form factor, (1+cos^2)/2
// test.  Conversion with pair_mode 1: deposit E - 2 mc2, emit two back-to-back 511 keV photons.
template <int kMode>
__device__ __forceinline__ int phase_rare(const Params &P, Smem &S, int slot, int gslot, double &c_edep) {
  Ph g;
  load_ph(P, S, slot, g);
  const bool in_vol = (g.flags & F_INVOL) != 0;
  const int m = g.mat;
  unsigned n_rayl = 0, n_conv = 0;
  DepStat ds; ds.dep = 0; ds.over = 0; ds.flush = 0; ds.tle_ent = 0; ds.tle_over = 0; ds.edep = 0.0;
  rng_begin(g.rng);
  if (!(g.flags & F_CONV)) {
    n_rayl = 1;
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
      float target = rng_uniform(g.rng) * amax;
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
      float u2 = rng_uniform(g.rng);
      if (0.5f * (1.0f + cost * cost) >= u2) break;
    }
    float sint = sqrtf(1.0f - cost * cost);
    float sphi, cphi;
    d_sincos2pi(rng_uniform(g.rng), sphi, cphi);
    rotate_dir(g.d, cost, sint, sphi, cphi);
    g.seg = 0.0f;
  } else {
    n_conv = 1;
    deposit<kMode>(P, g, P.depbuf + (size_t)gslot * kDepBuf, ds, (g.E - 2.0f * kMc2) * hist_weight<kMode>(P, gslot), m, in_vol, g.E, gslot);
    float cost = 2.0f * rng_uniform(g.rng) - 1.0f;
    float sint = sqrtf(1.0f - cost * cost);
    float sphi, cphi;
    d_sincos2pi(rng_uniform(g.rng), sphi, cphi);
    g.d[0] = sint * cphi; g.d[1] = sint * sphi; g.d[2] = cost;
    g.E = kMc2;
    g.seg = 0.0f;
    set_energy(P, S, slot, g);
    float *pend = P.pend + (size_t)gslot * 6;
#pragma unroll
    for (int a = 0; a < 3; ++a) { sc_st(pend + a, g.p[a]); sc_st(pend + 3 + a, -g.d[a]); }
    g.flags = (g.flags & ~(F_PEND | F_CONV)) | ((in_vol ? 1u : 2u) << F_PEND_SHIFT);
  }
  store_ph(S, slot, g);
  c_edep += ds.edep;
  {
    const unsigned t = stat_reduce_packed(n_rayl | (n_conv << 8) | (ds.dep << 16));
    const unsigned tr = t & 0xffu, tc = (t >> 8) & 0xffu;
    stat_bump(S, S_RAYL, tr);
    stat_bump(S, S_CONV, tc);
    stat_bump(S, S_TRACKS, 2u * tc);
    stat_bump(S, S_STEPS, tr + tc);
    stat_bump(S, S_DEP, t >> 16);
  }
  stat_overflow(S, ds.over, ds.flush);
  if (is_tle(kMode)) stat_tle_deposits<kMode>(S, ds.tle_over, ds.tle_ent);
  return in_vol ? Q_V : Q_W;
}

// Q_E: the photon of this slot is gone: absorbed on the spot (photoelectric effect, conversion in
// the kerma configuration), below the energy cut, or out of the world.  Either the history still
// owes its second annihilation photon, or it is closed (coalesced deposits -> grids) and the slot
// claims the next history.
template <int kMode>
__device__ __forceinline__ int phase_end(const Params &P, Smem &S, int slot, int gslot, double &c_edep) {
  Ph g;
  load_ph(P, S, slot, g);
  int next;
  unsigned n_flush = 0, n_abs = 0, n_hist = 0, n_phot = 0, n_conv = 0;
  DepStat ds; ds.dep = 0; ds.over = 0; ds.flush = 0; ds.tle_ent = 0; ds.tle_over = 0; ds.edep = 0.0;
  DepKeep keep;
  keep.j = -1;
  if (g.flags & F_ABSORB) {
    if (g.flags & F_CONV) n_conv = 1; else n_phot = 1;
    rng_begin(g.rng);
    // the single-actor kernels close the history in this very call unless a second annihilation photon is
    // pending: the entry of this last deposit goes to the grids from registers, not through the buffer
    DepKeys nopk; nopk.k0 = -1; nopk.k1 = -1;
    if (GB2_CLOSE_FUSED && is_one(kMode) && !(g.flags & F_PEND))
      deposit<kMode, false, true>(P, g, P.depbuf + (size_t)gslot * kDepBuf, ds, g.E * hist_weight<kMode>(P, gslot), g.mat, (g.flags & F_INVOL) != 0, g.E, gslot, nopk, keep);
    else
      deposit<kMode>(P, g, P.depbuf + (size_t)gslot * kDepBuf, ds, g.E * hist_weight<kMode>(P, gslot), g.mat, (g.flags & F_INVOL) != 0, g.E, gslot);
    g.flags &= ~(F_ABSORB | F_CONV);
    c_edep += ds.edep;
  }
  if (g.flags & F_PEND) {
    // second annihilation photon of a conversion (pair_mode 1): stored in the frame current at
    // creation (1 = volume, 2 = world); go back to the world frame and find the region again
    const float *pe = P.pend + (size_t)gslot * 6;
    float pp[3] = {sc_ld(pe), sc_ld(pe + 1), sc_ld(pe + 2)}, pd[3] = {sc_ld(pe + 3), sc_ld(pe + 4), sc_ld(pe + 5)};
    if (((g.flags & F_PEND) >> F_PEND_SHIFT) == 1u) {
      aff_apply_inv(P.w2v, pp, g.p);
      rot_apply_inv(P.w2v.m, pd, g.d);
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) { g.p[a] = pp[a]; g.d[a] = pd[a]; }
    }
    g.E = kMc2; g.seg = 0.0f;
    g.flags &= ~(F_PEND | F_EXIT);
    next = locate<kMode>(P, g);
    set_energy(P, S, slot, g);
    store_ph(S, slot, g);
  } else if (kMode == 1 && (g.flags & F_OPEN) && P.tle_tab && tl_ld(&P.tle_state[gslot].w) != 0u) {
    // the history's TLE table goes to the grids first: chunked, in the TLE phase; then back here
    tl_st(&P.tle_state[gslot].x, __float_as_uint(-2.0f));
    tl_st(&P.tle_state[gslot].y, 0u);
    g.flags = (g.flags & ~F_NEXT) | ((unsigned)(Q_E - 1) << F_NEXT_SHIFT);
    store_ph(S, slot, g);
    next = Q_T;
  } else {
    if (g.flags & F_OPEN) {
      if (kMode == 5) {   // the history's records are forgotten: its sums are in the grids already
        tl_st(&P.tle_state[gslot].w, 0u);
        tl_st(P.tle_aux + gslot, make_uint2(0u, 0u));
      }
      const int nbuf = (int)((g.flags & F_NBUF) >> F_NBUF_SHIFT);
      if (nbuf) {
        if (GB2_CLOSE_FUSED && is_one(kMode)) flush_buffer_keep<kMode>(P, P.depbuf + (size_t)gslot * kDepBuf, nbuf, keep.en, keep.j);
        else flush_buffer<kMode>(P, P.depbuf + (size_t)gslot * kDepBuf, nbuf);
      }
      n_flush = (unsigned)nbuf;
      if (kMode != 2 && kMode != 3 && P.nphsp > 0 && !(g.flags & F_STORED)) n_abs = 1;
    }
    // claim the next history id (one atomic per group of lanes that arrive together)
    const unsigned lane = threadIdx.x & 31u;
    const unsigned act = __activemask();
    const int leader = __ffs(act) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(P.next, (unsigned long long)__popc(act));
    base = __shfl_sync(act, base, leader);
    const unsigned long long idx = base + (unsigned long long)__popc(act & ((1u << lane) - 1u));
    if (idx >= P.count) {
      S.fm[slot] = 0;
      next = Q_RETIRED;
    } else {
      S.hidx[slot] = (unsigned)idx;
      rng_init(g.rng, P.seed, P.first + idx, 0u);
      g.flags = F_OPEN;
      g.mat = 0;
      generate_primary<kMode>(P, g, gslot);
      g.path = 0.0f; g.seg = 0.0f;
      n_hist = 1;
      next = locate<kMode>(P, g);
      if (g.E < P.e_cut) {
        c_edep += (double)(g.E * hist_weight<kMode>(P, gslot));   // below the table: absorbed where it starts
        next = Q_E;
        S.tbin[slot] = 0.0f;
      } else {
        set_energy(P, S, slot, g);
      }
      store_ph(S, slot, g);
    }
  }
  stat_add_small(S, S_FLUSH, n_flush);
  {
    // five 6-bit fields: every count is 0 or 1 per lane
    const unsigned t = stat_reduce_packed(n_abs | (n_hist << 6) | (n_phot << 12) | (n_conv << 18) | (ds.dep << 24));
    const unsigned th = (t >> 6) & 63u, tp = (t >> 12) & 63u, tc = (t >> 18) & 63u;
    stat_bump(S, S_ABSORBED, t & 63u);
    stat_bump(S, S_HIST, th);
    stat_bump(S, S_TRACKS, th);
    stat_bump(S, S_PHOT, tp);
    stat_bump(S, S_CONV, tc);
    stat_bump(S, S_STEPS, tp + tc);
    stat_bump(S, S_DEP, t >> 24);
  }
  stat_overflow(S, ds.over, ds.flush);
  if (is_tle(kMode)) stat_tle_deposits<kMode>(S, ds.tle_over, ds.tle_ent);
  return next;
}

// Q_T: the track-length estimator (GateTLEDoseActor::ScoreTLEDepositStep, GateTLEDoseActor.cpp:252-300).
// The flight that just ended at the photon position started `seg` before it and, for a photon that is
// leaving the volume, runs on to the exit face.  One Geant4 step of the reference is the chord of one
// image voxel; every chord deposits edep = w * 0.1 * E * mu_en/rho * L * rho (dose = edep / rho) at a
// point of the chord chosen by hit_type.  The voxel walk is an Amanatides-Woo traversal whose face
// distances are always computed from the integer voxel index; a call walks kTleSteps chords and the
// slot is queued again until the flight is done (lanes never wait for each other's long flights).
// The same phase writes the per-history table out at the end of a history (tle_state.x == -2).
template <typename LabelT, int kMode>
__device__ __forceinline__ int phase_tle(const Params &P, Smem &S, int slot, int gslot) {
  const unsigned fm = S.fm[slot];
  uint4 st = tl_ld(P.tle_state + gslot);
  float tcur = __uint_as_float(st.x);
  unsigned nlist = st.w;
  unsigned n_steps = 0, n_ent = 0, n_over = 0;
  bool done = false;
  if (kMode != 5 && tcur == -2.0f) {
    DepEntry *tab = P.tle_tab + (size_t)gslot * kTleH;
    const unsigned short *list = P.tle_list + (size_t)gslot * kTleMax;
    unsigned j = st.y;
    const unsigned jend = nlist < j + (unsigned)kTleSteps ? nlist : j + (unsigned)kTleSteps;
    for (; j < jend; ++j) {
      const unsigned hh = tl_ld(list + j);
      grid_add(P, ld_tab(tab + hh));
      tl_st(&tab[hh].key, (int)kTleEmpty);
      n_ent++;
    }
    done = j >= nlist;
    st.y = j;
    if (done) {
      st.x = __float_as_uint(-1.0f); st.y = 0u; st.w = 0u;
      tl_st(&P.tle_aux[gslot].y, 0u);
    }
  } else {
    const float p[3] = {S.px[slot], S.py[slot], S.pz[slot]};
    const float d[3] = {S.dx[slot], S.dy[slot], S.dz[slot]};
    const float seg = S.seg[slot], E = S.E[slot];
    int ie; float fe;
    bin_split(P, S.tbin[slot], ie, fe);
    float to = 0.0f;
    if (fm & F_EXIT) {
      float ti; int ai, ao;
      if (!ray_box_ax(p, d, P.half, ti, to, ai, ao)) to = 0.0f;
    }
    const float t_end = seg + to;
    float s[3], inv_d[3];
    int i[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      s[a] = __fmaf_rn(-d[a], seg, p[a]);
      inv_d[a] = d[a] != 0.0f ? 1.0f / d[a] : 0.0f;
    }
    // records behind the sums of the history (only where event-wise squares are scored)
    const bool sq = kMode == 5 || P.tle_tab != nullptr;
    uint2 aux = make_uint2(0u, 0u);
    if (sq) aux = tl_ld(P.tle_aux + gslot);
    unsigned nrec = aux.y;
    const unsigned nrec0 = nrec;
    float tmin = __uint_as_float(aux.x);
    if (tcur < 0.0f) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        int v = (int)floorf(__fmaf_rn(s[a], P.inv_sp[a], P.hs[a]));
        if (v < 0) v = 0;
        if (v > P.n[a] - 1) v = P.n[a] - 1;
        i[a] = v;
      }
      tcur = 0.0f;
      // a new scored flight = one more record of the history
      if (kMode == 5) {
        if (nrec >= (unsigned)P.tle_rec_cap) { nlist = 0; nrec = 0; n_over++; }
        nrec++;
        tmin = 0.0f;
      } else if (sq) {
        tle_record_begin_tab(P, gslot, nlist, nrec, n_ent, n_over);
      }
    } else {
      i[0] = (int)(st.y & 0xffffu); i[1] = (int)(st.y >> 16); i[2] = (int)st.z;
    }
    const float kE = (0.1f * E) * hist_weight<1>(P, gslot);
    Rng r;
    {
      const unsigned long long id = P.first + S.hidx[slot];
      r.k0 = (uint32_t)P.seed; r.k1 = (uint32_t)(P.seed >> 32);
      r.h0 = (uint32_t)id; r.h1 = (uint32_t)(id >> 32); r.blk = S.blk[slot]; r.pos = 4;
    }
    const LabelT *__restrict__ labels = reinterpret_cast<const LabelT *>(P.labels);
    // variant 5: the stored records whose voxel range meets the voxels this call can reach (kTleSteps steps from the
    // current voxel, each axis in its direction of travel): the others are not looked at again chord by chord
    unsigned long long cand = 0ull;
    if (kMode == 5 && nrec > 1u) {
      int blo[3], bhi[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        blo[a] = d[a] < 0.0f ? i[a] - kTleSteps : i[a];
        bhi[a] = d[a] > 0.0f ? i[a] + kTleSteps : i[a];
      }
      const float4 *rec = P.tle_rec + (size_t)gslot * (kTleRecMax * kTleRecW);
      for (unsigned rr = 0; rr + 1u < nrec; ++rr) {
        const float4 Cf = tl_ld(rec + kTleRecW * rr + 2);
        const unsigned cx = __float_as_uint(Cf.x), cy = __float_as_uint(Cf.y), cz = __float_as_uint(Cf.z);
        if ((int)(cx & 0xffffu) <= bhi[0] && (int)(cy >> 16) >= blo[0] && (int)(cx >> 16) <= bhi[1] && (int)(cz & 0xffffu) >= blo[1] &&
            (int)(cy & 0xffffu) <= bhi[2] && (int)(cz >> 16) >= blo[2]) {
          cand |= 1ull << rr;
        }
      }
    }
    for (int it = 0; it < kTleSteps; ++it) {
      float tn = kBig;
      int ax = -1;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float face;
        if (d[a] > 0.0f) face = __fmaf_rn((float)(i[a] + 1), P.sp[a], -P.half[a]);
        else if (d[a] < 0.0f) face = __fmaf_rn((float)i[a], P.sp[a], -P.half[a]);
        else continue;
        const float tb = (face - s[a]) * inv_d[a];
        if (tb < tn) { tn = tb; ax = a; }
      }
      const bool last = !(tn < t_end);
      const float tnext = last ? t_end : tn;
      const float L = tnext - tcur;
#if GB2_TLE_PREFETCH
      // the label of the NEXT voxel of the walk is asked into L2 before this voxel's label is waited for: the chain
      // label -> mu_en -> score of a chord then overlaps the DRAM access of its successor
      if (!last) {
        const int sx = ax == 0 ? (d[0] > 0.0f ? 1 : -1) : 0, sy = ax == 1 ? (d[1] > 0.0f ? 1 : -1) : 0, sz = ax == 2 ? (d[2] > 0.0f ? 1 : -1) : 0;
        const int jx = i[0] + sx, jy = i[1] + sy, jz = i[2] + sz;
        if ((unsigned)jx < (unsigned)P.n[0] && (unsigned)jy < (unsigned)P.n[1] && (unsigned)jz < (unsigned)P.n[2])
#if GB2_TLE_PREFETCH == 2
          prefetch_l1(labels + ((unsigned)jx + (unsigned)P.n[0] * ((unsigned)jy + (unsigned)P.n[1] * (unsigned)jz)));
#else
          prefetch_l2(labels + ((unsigned)jx + (unsigned)P.n[0] * ((unsigned)jy + (unsigned)P.n[1] * (unsigned)jz)));
#endif
      }
#endif
      if (L > 0.0f) {
        const unsigned flat = (unsigned)i[0] + (unsigned)P.n[0] * ((unsigned)i[1] + (unsigned)P.n[1] * (unsigned)i[2]);
        const int m = (int)__ldg(labels + flat);
        const float rho = __ldg(P.dens + m);
        n_steps++;
        if (kMode == 5) {
          // the one TLE actor, on the image grid: the voxel of the chord is the image voxel
          const DDose &D = P.dose[P.tle_a];
          const float2 mv = __ldg(D.muen2 + (size_t)ie * P.nmat + m);
          const float muen = __fmaf_rn(fe, mv.y - mv.x, mv.x);
          const float e = ((kE * muen) * L) * rho;
          const float dd = e / rho;
          const int key = (int)(((unsigned)P.tle_a << 28) | flat);
          tle_fold_add(P, D, gslot, nrec, nlist, tmin, cand, key, i, m, e, dd, 0, tcur, n_ent, n_over);
        } else
        for (int a = 0; a < P.ndose; ++a) {
          const DDose &D = P.dose[a];
          if (!D.tle || !(E <= D.tle_thr)) continue;
          const float2 mv = __ldg(D.muen2 + (size_t)ie * P.nmat + m);
          const float muen = __fmaf_rn(fe, mv.y - mv.x, mv.x);
          const float e = ((kE * muen) * L) * rho;
          const float dd = e / rho;
          int v;
          if (D.aligned) {
            v = (int)flat;
          } else {
            float fr = 1.0f;
            if (D.hit_type == GATE_B200_HIT_RANDOM) fr = rng_uniform(r);
            else if (D.hit_type == GATE_B200_HIT_MIDDLE) fr = 0.5f;
            else if (D.hit_type == GATE_B200_HIT_PRE) fr = 0.0f;
            const float tp = __fmaf_rn(fr, L, tcur);
            const float pos[3] = {__fmaf_rn(d[0], tp, s[0]), __fmaf_rn(d[1], tp, s[1]), __fmaf_rn(d[2], tp, s[2])};
            float ci[3];
            aff_apply(D.v2i, pos, ci);
            const float fx = floorf(ci[0] + 0.5f), fy = floorf(ci[1] + 0.5f), fz = floorf(ci[2] + 0.5f);
            if (fx < 0.0f || fy < 0.0f || fz < 0.0f || fx >= (float)D.size[0] || fy >= (float)D.size[1] ||
                fz >= (float)D.size[2])
              continue;
            v = (int)fx + D.size[0] * ((int)fy + D.size[1] * (int)fz);
          }
          const int key = (int)(((unsigned)a << 28) | (unsigned)v);
          if (D.flags & (GATE_B200_SCORE_EDEP_SQ | GATE_B200_SCORE_DOSE_SQ)) {
            const unsigned over1 = n_over;
            tle_table_add(P, gslot, nlist, key, e, dd, 0, n_ent, n_over);
            if (n_over != over1) nrec = 1;   // the table was written out while this flight is being scored
          } else {
            DepEntry z; z.key = key; z.e = e; z.d = dd; z.n = 0;
            grid_add(P, z);
            n_ent++;
          }
        }
      }
      if (last) { done = true; break; }
      if (tnext > tcur) tcur = tnext;
      // selects instead of i[ax]/d[ax]: a run-time index would push the walk state to local memory
      const float dax = ax == 0 ? d[0] : (ax == 1 ? d[1] : d[2]);
      const int stp = dax > 0.0f ? 1 : -1;
      const int iv = (ax == 0 ? i[0] : (ax == 1 ? i[1] : i[2])) + stp;
      const int nax = ax == 0 ? P.n[0] : (ax == 1 ? P.n[1] : P.n[2]);
      if (iv < 0 || iv >= nax) { done = true; break; }
      if (ax == 0) i[0] = iv; else if (ax == 1) i[1] = iv; else i[2] = iv;
    }
    S.blk[slot] = r.blk;
    if (kMode == 5 && done) {
      // the flight becomes a record: start, direction, length, energy, the range of voxels it walked
      int lo[3], hi[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        int v = (int)floorf(__fmaf_rn(s[a], P.inv_sp[a], P.hs[a]));
        if (v < 0) v = 0;
        if (v > P.n[a] - 1) v = P.n[a] - 1;
        lo[a] = v < i[a] ? v : i[a];
        hi[a] = v < i[a] ? i[a] : v;
      }
      float4 *rec = P.tle_rec + ((size_t)gslot * kTleRecMax + (nrec - 1u)) * kTleRecW;
      tl_st(rec, make_float4(s[0], s[1], s[2], t_end));
      tl_st(rec + 1, make_float4(d[0], d[1], d[2], kE));
      tl_st(rec + 2, make_float4(__uint_as_float((unsigned)lo[0] | ((unsigned)lo[1] << 16)), __uint_as_float((unsigned)lo[2] | ((unsigned)hi[0] << 16)),
                                 __uint_as_float((unsigned)hi[1] | ((unsigned)hi[2] << 16)), S.tbin[slot]));
      tl_st(rec + 3, make_float4(tmin, 0.0f, 0.0f, 0.0f));
    }
    if (sq && (nrec != nrec0 || __float_as_uint(tmin) != aux.x)) tl_st(P.tle_aux + gslot, make_uint2(__float_as_uint(tmin), nrec));
    st.x = __float_as_uint(done ? -1.0f : tcur);
    st.y = (unsigned)i[0] | ((unsigned)i[1] << 16);
    st.z = (unsigned)i[2];
    st.w = nlist;
  }
  tl_st(P.tle_state + gslot, st);
  stat_add(S, S_TLE_STEPS, n_steps);
  stat_add(S, S_TLE_ENT, n_ent);
  stat_tle_overflow(S, n_over, 0u);
  if (!done) return Q_T;
  S.fm[slot] = fm & ~F_NEXT;
  return (int)((fm & F_NEXT) >> F_NEXT_SHIFT) + 1;
}

// ------------------------------------------------------------------------------ kernel

// kMode 1: the engine has track-length-estimator actors (the plain kernels carry none of that code);
// kMode 2 and 3: the common dose run -- ONE dose actor, no phase-space actor, no phase-space source -- whose
// actor loop, plane loops and weight look-ups are resolved at compile time (the descriptor fields become
// immediate constant-bank operands); 2 also compiles hit_type post in, 3 keeps the hit-point emulation (the
// reference's default hit_type is random); kMode 4: ONE dose actor with phase-space actors (their loops stay);
// kMode 0: everything else
template <typename LabelT, int kMode>
__global__ void __launch_bounds__(threads_of<is_tle(kMode)>(), GB2_MIN_CTAS) transport_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);
  float *s_maj = reinterpret_cast<float *>(smem_raw + sizeof(Smem));
  const unsigned lane = threadIdx.x & 31u;
  const unsigned FULL = 0xffffffffu;

  for (int i = threadIdx.x; i < P.ne; i += blockDim.x) s_maj[i] = P.maj[i];
  for (int i = threadIdx.x; i < kSlots; i += blockDim.x) {
    S.fm[i] = 0; S.blk[i] = 0; S.hidx[i] = 0; S.tbin[i] = 0.0f;
#pragma unroll
    for (int t = 0; t < Q_COUNT; ++t) S.cell[t][i] = (unsigned short)((t == Q_E && i < kActiveSlots<kMode>) ? i : kEmpty);
  }
  if (threadIdx.x < Q_COUNT) {
    S.head[threadIdx.x] = 0;
    S.tail[threadIdx.x] = threadIdx.x == Q_E ? kActiveSlots<kMode> : 0;
    S.avail[threadIdx.x] = threadIdx.x == Q_E ? kActiveSlots<kMode> : 0;   // every slot starts by asking for a history
  }
  if (threadIdx.x < S_COUNT) { S.stats_lo[threadIdx.x] = 0u; S.stats_hi[threadIdx.x] = 0u; }
  if (threadIdx.x == 0) { S.retired = 0; S.edep = 0.0; S.cur = Q_E; S.cur2 = Q_E; }
  __syncthreads();

  unsigned c_tent = 0;
  double c_edep = 0.0;
  int polls = 0;
  const int slot_base = blockIdx.x * kActiveSlots<kMode>;   // index of slot 0 of this CTA in the per-slot global scratch

  for (;;) {
    // ---- claim slots that wait for the same phase (lane 0 negotiates with the queues): up to 32,
    // or up to 32 * kVWidth for the Woodcock phase, where every lane carries kVWidth slots.
    // The CTA drains ONE queue at a time (S.cur): all warps run the same code, which keeps the
    // instruction working set of the SM at one phase instead of the whole kernel.  When the current
    // queue cannot fill a warp any more, the warp that notices moves the CTA to the fullest queue.
    int q = -1, got = 0, base = 0;
    if (lane == 0) {
      for (;;) {
#if GB2_SPLIT
        // Two warp groups by scheduler (warp id & 3): the first prefers the Woodcock queue, the second every
        // other queue; a group whose own queues cannot fill a warp helps with the fullest queue of all.  The
        // Woodcock code then stays resident in the instruction caches of "its" schedulers instead of being
        // refetched after every visit of another phase.
        const bool vgroup = ((threadIdx.x >> 5) & 3u) < (unsigned)GB2_SPLIT;
        int *curp = vgroup ? &S.cur : &S.cur2;
        int bq = ld_vol(curp);
        int best = ld_vol(&S.avail[bq]);
        if (best < 32) {
          if (vgroup) {
            bq = Q_V; best = ld_vol(&S.avail[Q_V]);
          } else {
            best = -1;
#pragma unroll
            for (int t = 0; t < Q_COUNT; ++t) {
              if (t == Q_V) continue;
              const int a = ld_vol(&S.avail[t]);
              if (a > best) { best = a; bq = t; }
            }
          }
          if (best < 32) {
#pragma unroll
            for (int t = 0; t < Q_COUNT; ++t) {
              const int a = ld_vol(&S.avail[t]);
              if (a > best) { best = a; bq = t; }
            }
          }
        }
#else
        int *curp = &S.cur;
        int bq = ld_vol(&S.cur);
        int best = ld_vol(&S.avail[bq]);
        if (best < 32) {
#pragma unroll
          for (int t = 0; t < Q_COUNT; ++t) {
            const int a = ld_vol(&S.avail[t]);
            if (a > best) { best = a; bq = t; }
          }
        }
#endif
        if (best >= 32 || (best > 0 && polls >= 2)) {
          const int cap = bq == Q_V ? 32 * kVWidth : 32;
          const int want = best < cap ? best : cap;
          const int old = atomicSub(&S.avail[bq], want);
          const int g = old < want ? (old > 0 ? old : 0) : want;
          if (g < want) atomicAdd(&S.avail[bq], want - g);
          if (g > 0) {
            q = bq; got = g; base = atomicAdd(&S.head[bq], g);
            st_vol(curp, bq);
            break;
          }
          continue;
        }
        if (best <= 0 && ld_vol(&S.retired) >= kActiveSlots<kMode>) { q = -2; break; }
        ++polls;
        __nanosleep(polls < 8 ? 64 : 512);
      }
    }
    q = __shfl_sync(FULL, q, 0);
    if (q == -2) break;
    got = __shfl_sync(FULL, got, 0);
    base = __shfl_sync(FULL, base, 0);
    polls = 0;
    int slot[kVWidth], next[kVWidth];    // next = -3: this lane holds no slot in that column
#pragma unroll
    for (int w = 0; w < kVWidth; ++w) {
      slot[w] = -1;
      next[w] = -3;
      if (32 * w + (int)lane < got) {
        unsigned short *c = &S.cell[q][(unsigned)(base + 32 * w + (int)lane) % (unsigned)kSlots];
        unsigned v;
        while ((v = ld_cell(c)) == kEmpty) { }
        st_cell(c, kEmpty);
        slot[w] = (int)v;
      }
    }
    __threadfence_block();

    // ---- run the phase for all claimed slots
    constexpr bool kHasVS = GB2_SPLIT_SHELL && kMode == 2;       // measured: +1.3 ... 1.6 % for variant 2, -1 ... -2 % for 3 and 4   // Q_T serves as the queue of the shell-start calls
    if (q == Q_V) {
      VCall vc[kVWidth];
#pragma unroll
      for (int w = 0; w < kVWidth; ++w)
        if (slot[w] >= 0) volume_gather_labels<LabelT, kHasVS ? 0 : 2>(P, S, s_maj, slot[w], vc[w]);
#pragma unroll
      for (int w = 0; w < kVWidth; ++w)
        if (slot[w] >= 0) volume_gather_mu<LabelT>(P, vc[w]);
#pragma unroll
      for (int w = 0; w < kVWidth; ++w)
        if (slot[w] >= 0) next[w] = volume_decide<kMode>(P, S, slot[w], slot_base + slot[w], vc[w], c_tent);
    } else if (kHasVS && q == Q_T) {
      if (slot[0] >= 0) {
        VCall vc;
        volume_gather_labels<LabelT, 1>(P, S, s_maj, slot[0], vc);
        volume_gather_mu<LabelT>(P, vc);
        next[0] = volume_decide<kMode>(P, S, slot[0], slot_base + slot[0], vc, c_tent);
      }
    } else {
      // single-slot phases, with chaining (GB2_CHAIN): each phase function is inlined ONCE, the chain loops back here
      int qq = q;
      bool active = slot[0] >= 0;
      for (int chain = 0;; ++chain) {
        if (active) {
          if (qq == Q_C) next[0] = phase_compton<kMode>(P, S, slot[0], slot_base + slot[0], c_edep);
          else if (qq == Q_E) next[0] = phase_end<kMode>(P, S, slot[0], slot_base + slot[0], c_edep);
          else if (qq == Q_W) next[0] = phase_world<kMode>(P, S, slot[0], slot_base + slot[0]);
          else if (is_tle(kMode) && qq == Q_T) next[0] = phase_tle<LabelT, kMode>(P, S, slot[0], slot_base + slot[0]);
          else next[0] = phase_rare<kMode>(P, S, slot[0], slot_base + slot[0], c_edep);
        }
        if (chain >= GB2_CHAIN) break;
        __syncwarp();
        const int nx = slot[0] >= 0 ? next[0] : -3;
        const int nE = __popc(__ballot_sync(FULL, nx == Q_E)), nW = __popc(__ballot_sync(FULL, nx == Q_W));
        int qn = -1;
        if (nE >= GB2_CHAIN_MIN && nE >= nW) qn = Q_E;
        else if (nW >= GB2_CHAIN_MIN) qn = Q_W;
        if (qn < 0) break;                      // warp-uniform
        qq = qn;
        active = nx == qn;
      }
    }
    if (kHasVS && P.shell) {
      // a photon whose next Woodcock call starts outside the inner box waits in the queue of the shell-start calls
#pragma unroll
      for (int w = 0; w < kVWidth; ++w)
        if (next[w] == Q_V) {
          const float ps[3] = {S.px[slot[w]], S.py[slot[w]], S.pz[slot[w]]};
          if (!inside_tbox(P, ps)) next[w] = Q_T;
        }
    }
    __syncwarp();
    __threadfence_block();   // slot state visible before the slot is published again

    // ---- push every slot to the queue of its next phase: one pass, one leader lane per target queue
#pragma unroll
    for (int w = 0; w < kVWidth; ++w) {
      if (w > 0 && (q != Q_V || got <= 32 * w)) break;   // warp-uniform
      const int nx = next[w];
      const unsigned mine = __match_any_sync(FULL, nx);   // the lanes that go to the same queue
      const unsigned mret = __ballot_sync(FULL, nx == Q_RETIRED);
      const bool pushing = nx >= 0;
      const int leader = pushing ? __ffs(mine) - 1 : (int)lane;
      const int n = __popc(mine);
      int tk = 0;
      if (pushing && (int)lane == leader) tk = atomicAdd(&S.tail[nx], n);
      tk = __shfl_sync(FULL, tk, leader);
      if (pushing) {
        unsigned short *c = &S.cell[nx][(unsigned)(tk + __popc(mine & ((1u << lane) - 1u))) % (unsigned)kSlots];
        while (ld_cell(c) != kEmpty) { }   // the previous occupant is claimed, its reader clears it at once
        st_cell(c, (unsigned)slot[w]);
      }
      __syncwarp();
      if (pushing && (int)lane == leader) atomicAdd(&S.avail[nx], n);
      if (mret && lane == 0) atomicAdd(&S.retired, __popc(mret));
    }
  }

  // ---- statistics: per-thread -> warp -> CTA -> global
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c_tent += __shfl_xor_sync(FULL, c_tent, o);
    c_edep += __shfl_xor_sync(FULL, c_edep, o);
  }
  if (lane == 0) {
    if (c_tent) stat_add64(S, S_TENT, c_tent);
    if (c_edep != 0.0) atomicAdd(&S.edep, c_edep);
  }
  __syncthreads();
  if (threadIdx.x < S_COUNT) {
    const unsigned long long v = ((unsigned long long)S.stats_hi[threadIdx.x] << 32) | S.stats_lo[threadIdx.x];
    if (v) atomicAdd(P.stats + threadIdx.x, v);
  }
  if (threadIdx.x == 0 && S.edep != 0.0) atomicAdd(P.edep_total, S.edep);
}

}  // namespace gb2
