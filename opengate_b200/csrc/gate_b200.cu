// gate_b200.cu — host side of the C-ABI declared in include/gate_b200.h.
//
// Owns all device memory of one engine (tables, label volume, accumulator grids, phase-space
// columns, statistics), converts the double-precision descriptors of the ABI to the float
// device model exactly once (same expressions as the oracle so both sides see identical
// constants), and launches gb2::transport_kernel as a persistent grid sized from the SM
// count.  No torch, no pybind: plain CUDA runtime.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "transport.cuh"

using namespace gb2;

namespace {

struct DeviceBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

struct DoseActor {
  gate_b200_dose_desc desc;
  DDose dev;
  size_t nvox = 0;
  double *rec = nullptr;      // [nvox][gstride] voxel records: the scored ones of {edep, edep2, dose, dose2}
};

struct PhspActor {
  gate_b200_phsp_desc desc;
  DPhsp dev;
};

struct SourceRec {
  DSource dev;
};

}  // namespace

struct gate_b200_engine {
  int device = 0;
  uint64_t seed = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  bool have_tables = false, have_volume = false;
  Params P;
  std::vector<void *> allocs;           // everything to cudaFree
  std::vector<float> h_mu, h_dens;      // host copies for the majorant / the shell rule
  std::vector<DoseActor> dose;
  std::vector<PhspActor> phsp;
  std::vector<SourceRec> src;
  float *d_maj = nullptr;
  void *d_labels = nullptr;
  unsigned long long *d_next = nullptr;
  unsigned long long *d_stats = nullptr;
  double *d_edep = nullptr;
  double *d_stage = nullptr;            // dense staging array of get/set_dose_grid
  size_t stage_elems = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_events;
  double kernel_ms = 0.0;
  std::string kernel_name;               // instantiation of the most recent launch
  int sm_count = 148;
  int max_ctas = 148;                   // persistent CTAs (one per SM); scratch is sized for it
  int label_bytes = 0;
  size_t label_total_bytes = 0;
  size_t l2_persist_max = 0, l2_window_max = 0;
  bool l2_persist = false;               // measured on B200: a persisting window for the labels is slower (C3 6.0e8 vs 1.0e9/s)
  size_t l2_limit_set = 0;               // cudaLimitPersistingL2CacheSize this engine applied to ITS device
  // bounding box (voxel indices, inclusive) of the voxels whose label differs from that of voxel 0 (shell_mat);
  // has_shell: the box is neither empty nor the whole volume
  int has_shell = 0, shell_mat = 0;
  int box_lo[3] = {0, 0, 0}, box_hi[3] = {0, 0, 0};
  double vol_spacing[3] = {1, 1, 1};
  unsigned long long *d_unc_u = nullptr; // scratch of gate_b200_mean_uncertainty: {max bits, count}
  double *d_unc_s = nullptr;             //                                         sum
};

namespace {

int fail(gate_b200_engine *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

#define CU(call)                                                                             \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail(h, GATE_B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                       \
  } while (0)

// Device memory comes from the device's default stream-ordered pool (cudaMallocAsync) with the
// release threshold lifted: an engine's gigabytes of accumulator grids and scratch go back to the
// pool when it is destroyed and the next engine of the process gets them without a trip to the
// driver (measured on B200, C3: cudaFree of one engine's buffers 185 ms, cudaMalloc 60 ms per run).
// Everything is allocated, filled and used in the order of h->stream.
template <typename T>
int dev_alloc(gate_b200_engine *h, T **out, size_t n, bool zero = false) {
  void *p = nullptr;
  CU(cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), h->stream));
  h->allocs.push_back(p);
  if (zero) CU(cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
  *out = static_cast<T *>(p);
  return 0;
}

// host -> device in stream order; returns when the host buffer may be reused (pageable memory is
// staged by the runtime before cudaMemcpyAsync returns, pinned memory needs the synchronize)
template <typename T>
int dev_upload(gate_b200_engine *h, const T **out, const T *src, size_t n) {
  T *p = nullptr;
  int rc = dev_alloc(h, &p, n);
  if (rc) return rc;
  CU(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *out = p;
  return 0;
}

// class volume of the Woodcock phase: 2 bits per voxel, four voxels per byte (x fastest, flat index >> 2)
template <typename LabelT>
__global__ void __launch_bounds__(256) class_pack_kernel(const LabelT *__restrict__ labels, unsigned long long nv,
                                                         const unsigned char *__restrict__ cls_of, unsigned char *__restrict__ out) {
  const unsigned long long nb = (nv + 3ull) / 4ull;
  for (unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; b < nb; b += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned v = 0;
    for (int j = 0; j < 4; ++j) {
      const unsigned long long i = 4ull * b + j;
      if (i < nv) v |= (unsigned)cls_of[labels[i]] << (2 * j);
    }
    out[b] = (unsigned char)v;
  }
}

// Which labels occur in the volume (bit set, one bit per material), the first voxel whose label has no material,
// and the bounding box (voxel indices) of the voxels whose label differs from that of voxel 0 (the inner box of
// the homogeneous shell): one pass over the device copy of the label volume.
template <typename LabelT>
__global__ void __launch_bounds__(256) label_scan_kernel(const LabelT *__restrict__ labels, unsigned long long nv, int nmat,
                                                         int nx, int ny, unsigned *__restrict__ present,
                                                         unsigned long long *first_bad, int *__restrict__ bbox) {
  extern __shared__ unsigned s_present[];
  __shared__ int s_box[6];
  const int nw = (nmat + 31) >> 5;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) s_present[i] = 0u;
  if (threadIdx.x < 6) s_box[threadIdx.x] = (threadIdx.x & 1) ? -1 : 0x7fffffff;   // {min, max} per axis
  __syncthreads();
  const unsigned l0 = (unsigned)labels[0];
  constexpr int kPer = 16 / (int)sizeof(LabelT);          // labels per 16-byte load
  const unsigned long long nvec = nv / kPer;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  int bx0 = 0x7fffffff, bx1 = -1, by0 = 0x7fffffff, by1 = -1, bz0 = 0x7fffffff, bz1 = -1;
  auto see = [&](unsigned l, unsigned long long idx) {
    if (l >= (unsigned)nmat) { atomicMin(first_bad, idx); return; }
    if (!((s_present[l >> 5] >> (l & 31u)) & 1u)) atomicOr(&s_present[l >> 5], 1u << (l & 31u));
    if (l != l0) {
      const int x = (int)(idx % (unsigned long long)nx);
      const unsigned long long r = idx / (unsigned long long)nx;
      const int y = (int)(r % (unsigned long long)ny), z = (int)(r / (unsigned long long)ny);
      bx0 = min(bx0, x); bx1 = max(bx1, x); by0 = min(by0, y); by1 = max(by1, y); bz0 = min(bz0, z); bz1 = max(bz1, z);
    }
  };
  const uint4 *v4 = reinterpret_cast<const uint4 *>(labels);   // cudaMalloc'ed: 256-byte aligned
  const unsigned pat = sizeof(LabelT) == 1 ? l0 * 0x01010101u : l0 * 0x00010001u;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 v = __ldg(v4 + i);
    if (v.x == pat && v.y == pat && v.z == pat && v.w == pat) continue;   // sixteen bytes of the shell label (seen at voxel 0)
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (sizeof(LabelT) == 1) {
#pragma unroll
        for (int b = 0; b < 4; ++b) see((w[k] >> (8 * b)) & 0xffu, i * kPer + 4 * k + b);
      } else {
        see(w[k] & 0xffffu, i * kPer + 2 * k);
        see(w[k] >> 16, i * kPer + 2 * k + 1);
      }
    }
  }
  for (unsigned long long i = nvec * kPer + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride)
    see((unsigned)labels[i], i);
  if (blockIdx.x == 0 && threadIdx.x == 0) see(l0, 0);          // the skipped vectors: their label is l0
  if (bx1 >= 0) {
    atomicMin(&s_box[0], bx0); atomicMax(&s_box[1], bx1); atomicMin(&s_box[2], by0); atomicMax(&s_box[3], by1);
    atomicMin(&s_box[4], bz0); atomicMax(&s_box[5], bz1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nw; i += blockDim.x)
    if (s_present[i]) atomicOr(&present[i], s_present[i]);
  if (threadIdx.x < 6 && s_box[1] >= 0) {
    if (threadIdx.x & 1) atomicMax(&bbox[threadIdx.x], s_box[threadIdx.x]);
    else atomicMin(&bbox[threadIdx.x], s_box[threadIdx.x]);
  }
}

// HU -> label on the device: label = bin_labels[np.digitize(hu, bins)] with sorted bins, i.e. the number
// of bins <= hu (an upper bound search in double: exactly what numpy does after promoting the HU type;
// reference geometry/volumes.py:1200-1243).  bins and labels are staged in shared memory.
constexpr int kMaxHuBins = 4096;
template <typename HuT, typename LabT>
__global__ void __launch_bounds__(256) hu_to_label_kernel(const HuT *__restrict__ hu, size_t n, int n_bins,
                                                          const double *__restrict__ bins, const uint16_t *__restrict__ bin_labels,
                                                          LabT *__restrict__ out) {
  extern __shared__ __align__(8) unsigned char hu_smem[];
  double *s_bins = reinterpret_cast<double *>(hu_smem);
  uint16_t *s_lab = reinterpret_cast<uint16_t *>(s_bins + n_bins);
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x) s_bins[i] = bins[i];
  for (int i = threadIdx.x; i <= n_bins; i += blockDim.x) s_lab[i] = bin_labels[i];
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double x = (double)hu[i];
    int lo = 0, hi = n_bins;            // first index with bins[idx] > x
    if (x != x) lo = n_bins;            // numpy sorts NaN last
    else
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_bins[mid] <= x) lo = mid + 1; else hi = mid;
      }
    out[i] = (LabT)s_lab[lo];
  }
}

// component k of every voxel record <-> a dense array (the layout the ABI hands to the caller)
__global__ void __launch_bounds__(256) grid_extract_kernel(const double *__restrict__ comp, size_t n, size_t stride,
                                                           double scale, double *__restrict__ dense) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dense[i] = comp[i * stride] * scale;
}
// relative uncertainty of the mean from sum(x) and sum(x^2) over n events (utility.py:551-560, dataitems.py:1956-1993);
// IEEE double division and square root, no contraction: the same bits as the numpy expressions
__global__ void __launch_bounds__(256) grid_uncertainty_kernel(const double *__restrict__ val, const double *__restrict__ sq, size_t n,
                                                               size_t stride, double nev, double *__restrict__ dense) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = val[i * stride];
    const double mean = v / nev;
    double var = (sq[i * stride] / nev - mean * mean) / (nev - 1.0);
    if (var < 0.0) var = 0.0;
    dense[i] = v != 0.0 ? sqrt(var) / mean : 0.0;
  }
}
__global__ void __launch_bounds__(256) grid_scatter_kernel(double *__restrict__ comp, size_t n, size_t stride,
                                                           const double *__restrict__ dense) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    comp[i * stride] = dense[i];
}

void affine_from_double(Affine &a, const double m12[12]) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) a.m[3 * r + c] = (float)m12[4 * r + c];
    a.t[r] = (float)m12[4 * r + 3];
  }
}

int collect_kernel_time(gate_b200_engine *h) {
  for (auto &pr : h->pending_events) {
    float ms = 0.f;
    CU(cudaEventSynchronize(pr.second));
    CU(cudaEventElapsedTime(&ms, pr.first, pr.second));
    h->kernel_ms += ms;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  h->pending_events.clear();
  return 0;
}

}  // namespace

extern "C" {

int gate_b200_abi_version(void) { return GATE_B200_ABI_VERSION; }

int gate_b200_create(int device_id, uint64_t seed, gate_b200_engine **out) {
  if (!out) return GATE_B200_EINVAL;
  *out = nullptr;
  gate_b200_engine *h = new (std::nothrow) gate_b200_engine();
  if (!h) return GATE_B200_ENOMEM;
  *out = h;  // returned even on failure so that the caller can read last_error
  h->device = device_id;
  h->seed = seed;
  memset(&h->P, 0, sizeof h->P);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(h, GATE_B200_ECUDA, "no CUDA device available: %s (the engine has no CPU fallback)",
                cudaGetErrorString(e));
  if (device_id < 0 || device_id >= ndev)
    return fail(h, GATE_B200_EINVAL, "device %d out of range (%d devices)", device_id, ndev);
  const bool trace = getenv("GATE_B200_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[gate_b200_create] %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  CU(cudaSetDevice(device_id));
  lap("set_device");
  // three attributes instead of cudaGetDeviceProperties (2.4 ms per call on B200)
  int sm_count = 0, l2_persist_max = 0, l2_window_max = 0;
  CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device_id));
  CU(cudaDeviceGetAttribute(&l2_persist_max, cudaDevAttrMaxPersistingL2CacheSize, device_id));
  CU(cudaDeviceGetAttribute(&l2_window_max, cudaDevAttrMaxAccessPolicyWindowSize, device_id));
  h->sm_count = sm_count;
  lap("device_attributes");
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  if (const char *g = getenv("GATE_B200_L2_FETCH")) {   // experiment: L2 fetch granularity (32 / 64 / 128 bytes) for the random gathers
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
  }
  h->own_stream = true;
  lap("stream_create");
  {
    // keep freed blocks in the pool for the next engine of this process (GATE_B200_POOL_RETAIN=0:
    // give them back to the driver at the next synchronisation, the CUDA default)
    const char *s = getenv("GATE_B200_POOL_RETAIN");
    if (!s || atoi(s) != 0) {
      cudaMemPool_t pool;
      CU(cudaDeviceGetDefaultMemPool(&pool, device_id));
      uint64_t thr = UINT64_MAX;
      CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
  }
  if (int rc = dev_alloc(h, &h->d_next, 1, true)) return rc;
  if (int rc = dev_alloc(h, &h->d_stats, (size_t)S_COUNT, true)) return rc;
  if (int rc = dev_alloc(h, &h->d_edep, 1, true)) return rc;
  h->P.seed = seed;
  h->P.next = h->d_next;
  h->P.stats = h->d_stats;
  h->P.edep_total = h->d_edep;
  h->P.wh[0] = h->P.wh[1] = h->P.wh[2] = 1.0e6f;
  h->max_ctas = h->sm_count;
  if (const char *s = getenv("GATE_B200_CTAS")) h->max_ctas = std::max(1, atoi(s));
  h->l2_persist_max = (size_t)l2_persist_max;
  h->l2_window_max = (size_t)l2_window_max;
  if (const char *s = getenv("GATE_B200_L2_PERSIST")) h->l2_persist = atoi(s) != 0;
  // per-slot global scratch: coalescing buffers of the histories in flight, pending photons
  const size_t nslots = (size_t)h->max_ctas * kSlots;
  lap("pool + counters");
  if (int rc = dev_alloc(h, &h->P.depbuf, nslots * kDepBuf)) return rc;
  if (int rc = dev_alloc(h, &h->P.pend, nslots * 6)) return rc;
  if (int rc = dev_alloc(h, &h->P.slot_w, nslots)) return rc;
  lap("slot scratch");
  return GATE_B200_OK;
}

void gate_b200_destroy(gate_b200_engine *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto &pr : h->pending_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (void *p : h->allocs) cudaFreeAsync(p, h->stream);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int gate_b200_release_cached_memory(int device_id) {
  cudaMemPool_t pool;
  if (cudaSetDevice(device_id) != cudaSuccess) return GATE_B200_ECUDA;
  if (cudaDeviceSynchronize() != cudaSuccess) return GATE_B200_ECUDA;
  if (cudaDeviceGetDefaultMemPool(&pool, device_id) != cudaSuccess) return GATE_B200_ECUDA;
  return cudaMemPoolTrimTo(pool, 0) == cudaSuccess ? GATE_B200_OK : GATE_B200_ECUDA;
}

const char *gate_b200_last_error(const gate_b200_engine *h) { return h ? h->err.c_str() : "null engine"; }
const char *gate_b200_kernel_name(const gate_b200_engine *h) { return h ? h->kernel_name.c_str() : ""; }

int gate_b200_set_stream(gate_b200_engine *h, void *cuda_stream) {
  if (!h) return GATE_B200_EINVAL;
  CU(cudaSetDevice(h->device));
  if (h->stream) CU(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (cuda_stream == nullptr) {
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  } else {
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    h->own_stream = false;
  }
  return GATE_B200_OK;
}

int gate_b200_set_tables(gate_b200_engine *h, const gate_b200_tables *t) {
  if (!h || !t) return GATE_B200_EINVAL;
  if (t->n_materials < 1 || t->n_energy < 2 || t->n_q < 2)
    return fail(h, GATE_B200_EINVAL, "tables: need >=1 material, >=2 energy and q nodes");
  if (t->n_materials > kMaxMaterials)
    return fail(h, GATE_B200_EINVAL, "tables: at most %d materials (got %d)", kMaxMaterials, t->n_materials);
  if (!(t->e_min_MeV > 0) || !(t->e_max_MeV > t->e_min_MeV) || !(t->q2_max > t->q2_min) || !(t->q2_min > 0))
    return fail(h, GATE_B200_EINVAL, "tables: bad grid limits");
  CU(cudaSetDevice(h->device));
  Params &P = h->P;
  P.nmat = t->n_materials; P.ne = t->n_energy; P.nq = t->n_q; P.pair_mode = t->pair_mode;
  const size_t n = (size_t)P.nmat * P.ne, nqq = (size_t)P.nmat * P.nq;
  if (int rc = dev_upload(h, &P.dens, t->density_g_cm3, (size_t)P.nmat)) return rc;
  {
    // energy-major re-packing of the material-major ABI arrays (same float values)
    std::vector<float2> mu2(n);
    std::vector<float4> prob(n);
    for (int ie = 0; ie < P.ne; ++ie) {
      const int ie1 = ie + 1 < P.ne ? ie + 1 : ie;
      for (int m = 0; m < P.nmat; ++m) {
        const size_t a = (size_t)m * P.ne + ie, b = (size_t)m * P.ne + ie1, o = (size_t)ie * P.nmat + m;
        mu2[o] = make_float2(t->mu_total_per_mm[a], t->mu_total_per_mm[b]);
        prob[o] = make_float4(t->cum_phot[a], t->cum_compt[a], t->cum_rayl[a], 0.0f);
      }
    }
    if (int rc = dev_upload(h, &P.mu2, mu2.data(), n)) return rc;
    if (int rc = dev_upload(h, &P.prob, prob.data(), n)) return rc;
  }
  if (int rc = dev_upload(h, &P.ray, t->rayl_cum_ff2, nqq)) return rc;
  if (int rc = dev_upload(h, &P.sf, t->compt_sf, nqq)) return rc;
  h->h_mu.assign(t->mu_total_per_mm, t->mu_total_per_mm + n);
  h->h_dens.assign(t->density_g_cm3, t->density_g_cm3 + P.nmat);
  const double le0 = log(t->e_min_MeV), le1 = log(t->e_max_MeV);
  P.ln_emin = (float)le0;
  P.inv_dlne = (float)((double)(P.ne - 1) / (le1 - le0));
  P.e_cut = (float)t->e_cut_MeV;
  const double lq0 = log(t->q2_min), lq1 = log(t->q2_max);
  P.ln_q2min = (float)lq0;
  P.inv_dlnq2 = (float)((double)(P.nq - 1) / (lq1 - lq0));
  std::vector<float> q2n(P.nq);
  for (int j = 0; j < P.nq; ++j)
    q2n[j] = (float)exp(lq0 + (lq1 - lq0) * (double)j / (double)(P.nq - 1));
  if (int rc = dev_upload(h, &P.q2n, q2n.data(), q2n.size())) return rc;
  if (int rc = dev_alloc(h, &h->d_maj, (size_t)P.ne, true)) return rc;
  P.maj = h->d_maj;
  h->have_tables = true;
  return GATE_B200_OK;
}

int gate_b200_set_world(gate_b200_engine *h, const double half_size[3], int32_t material) {
  if (!h || !half_size) return GATE_B200_EINVAL;
  if (!h->have_tables) return fail(h, GATE_B200_ESTATE, "set_world before set_tables");
  if (material < 0 || material >= h->P.nmat) return fail(h, GATE_B200_EINVAL, "world material %d out of range", material);
  for (int a = 0; a < 3; ++a) h->P.wh[a] = (float)half_size[a];
  h->P.world_mat = material;
  return GATE_B200_OK;
}

// Common tail of the two ways to hand over the image volume: `dl` is the device copy of the label volume.
// Scans it for the materials present (-> majorant, same rule as the oracle's build_majorant) and for labels
// without a material (the scan is inside the end-to-end path of every run, hence on the device), then
// records the geometry.  `host_labels` (may be null) only serves the error message.
static int finish_image_volume(gate_b200_engine *h, uint8_t *dl, const void *host_labels, int32_t label_bytes,
                               const int32_t size[3], const double spacing[3], const double world_to_volume[12]) {
  Params &P = h->P;
  const size_t nv = (size_t)size[0] * size[1] * size[2];
  std::vector<char> present(P.nmat, 0);
  {
    const int nw = (P.nmat + 31) / 32;
    unsigned *d_present = nullptr;
    unsigned long long *d_bad = nullptr;
    int *d_box = nullptr;
    if (int rc = dev_alloc(h, &d_present, (size_t)nw, true)) return rc;
    if (int rc = dev_alloc(h, &d_bad, 1)) return rc;
    if (int rc = dev_alloc(h, &d_box, 6)) return rc;
    CU(cudaMemsetAsync(d_bad, 0xff, sizeof(unsigned long long), h->stream));
    int box[6] = {0x7fffffff, -1, 0x7fffffff, -1, 0x7fffffff, -1};
    CU(cudaMemcpyAsync(d_box, box, sizeof box, cudaMemcpyHostToDevice, h->stream));
    const int blocks = (int)std::min<size_t>((size_t)h->sm_count * 8, (nv + 4095) / 4096);
    if (label_bytes == 1)
      label_scan_kernel<uint8_t><<<blocks, 256, nw * sizeof(unsigned), h->stream>>>(dl, nv, P.nmat, size[0], size[1], d_present, d_bad, d_box);
    else
      label_scan_kernel<uint16_t><<<blocks, 256, nw * sizeof(unsigned), h->stream>>>(reinterpret_cast<const uint16_t *>(dl), nv, P.nmat, size[0], size[1], d_present, d_bad, d_box);
    CU(cudaGetLastError());
    std::vector<unsigned> words(nw, 0u);
    unsigned long long bad = ~0ull;
    uint16_t lab0 = 0;
    CU(cudaMemcpyAsync(words.data(), d_present, nw * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(box, d_box, sizeof box, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&lab0, dl, (size_t)label_bytes, cudaMemcpyDeviceToHost, h->stream));   // little endian
    CU(cudaStreamSynchronize(h->stream));
    if (bad != ~0ull) {
      const size_t i = (size_t)bad;
      uint16_t v16 = 0;
      if (host_labels) {
        v16 = label_bytes == 1 ? (uint16_t)static_cast<const uint8_t *>(host_labels)[i] : static_cast<const uint16_t *>(host_labels)[i];
      } else {
        CU(cudaMemcpy(&v16, dl + i * (size_t)label_bytes, (size_t)label_bytes, cudaMemcpyDeviceToHost));   // little endian
      }
      return fail(h, GATE_B200_EINVAL, "label %d at voxel %zu has no material", (int)v16, i);
    }
    // the shell: the voxels outside the bounding box of the labels that differ from the label of voxel 0 (all of
    // one material); none when the volume is homogeneous or the box is the whole volume
    h->has_shell = 0;
    h->shell_mat = (int)lab0;
    if (box[1] >= 0 && !(box[0] == 0 && box[1] == size[0] - 1 && box[2] == 0 && box[3] == size[1] - 1 && box[4] == 0 && box[5] == size[2] - 1)) {
      h->has_shell = 1;
      for (int a = 0; a < 3; ++a) { h->box_lo[a] = box[2 * a]; h->box_hi[a] = box[2 * a + 1]; }
    }
    for (int a = 0; a < 3; ++a) h->vol_spacing[a] = spacing[a];
    // the inner box of the shell in the volume frame (same double expressions as the oracle)
    // only a shell of gas (density < 0.05 g/cm3: the air around a patient) is crossed analytically: in a dense
    // shell (the water slabs of config 2) the Woodcock steps it replaces are accepted half of the time anyway, and
    // the extra work at the start of every call that begins out there was measured to cost 11 %
    P.shell = (h->has_shell && h->h_dens[h->shell_mat] < 0.05f && !getenv("GATE_B200_NO_SHELL")) ? 1 : 0;
    P.shell_mat = h->shell_mat;
    for (int a = 0; a < 3; ++a) {
      const double lo = -0.5 * size[a] * spacing[a] + h->box_lo[a] * spacing[a];
      const double hi = -0.5 * size[a] * spacing[a] + (h->box_hi[a] + 1) * spacing[a];
      P.tc[a] = P.shell ? (float)(0.5 * (lo + hi)) : 0.0f;
      P.th[a] = P.shell ? (float)(0.5 * (hi - lo)) : 0.0f;
    }
    for (int m = 0; m < P.nmat; ++m) present[m] = (char)((words[m >> 5] >> (m & 31)) & 1u);
  }
  std::vector<float> maj(P.ne, 0.f);
  for (int i = 0; i < P.ne; ++i) {
    float mx = 0.f;
    for (int m = 0; m < P.nmat; ++m)
      if (present[m] && h->h_mu[(size_t)m * P.ne + i] > mx) mx = h->h_mu[(size_t)m * P.ne + i];
    maj[i] = mx;
  }
  CU(cudaMemcpyAsync(h->d_maj, maj.data(), maj.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  // ---- the class volume (Params::cls): the materials present in four classes; at most four present -> class = material
  P.cls = nullptr;
  P.cmax = nullptr;
  if (GB2_CLASSVOL) {
    std::vector<unsigned char> cls_of(P.nmat, 0);
    std::vector<int> pres;
    for (int m = 0; m < P.nmat; ++m) if (present[m]) pres.push_back(m);
    P.cls_exact = pres.size() <= 4 ? 1 : 0;
    for (int k = 0; k < 4; ++k) P.cls_mat[k] = pres.empty() ? 0 : pres[std::min<size_t>(k, pres.size() - 1)];
    if (P.cls_exact) {
      for (size_t k = 0; k < pres.size(); ++k) cls_of[pres[k]] = (unsigned char)k;
    } else {
      // by density: gas, lung-like, soft tissue, denser (Schneider materials of a CT: a class spans a few percent of
      // attenuation except the last one)
      for (int m : pres) {
        const float rho = h->h_dens[m];
        cls_of[m] = rho < 0.05f ? 0 : (rho < 0.6f ? 1 : (rho < 1.12f ? 2 : 3));
      }
    }
    std::vector<float4> cmax(P.ne, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int i = 0; i < P.ne; ++i) {
      float mx[4] = {0.f, 0.f, 0.f, 0.f};
      for (int m : pres) mx[cls_of[m]] = std::max(mx[cls_of[m]], h->h_mu[(size_t)m * P.ne + i]);
      cmax[i] = make_float4(mx[0], mx[1], mx[2], mx[3]);
    }
    const unsigned char *d_cls_of = nullptr;
    unsigned char *d_cls = nullptr;
    const float4 *d_cmax = nullptr;
    if (int rc = dev_upload(h, &d_cls_of, cls_of.data(), cls_of.size())) return rc;
    if (int rc = dev_upload(h, &d_cmax, cmax.data(), cmax.size())) return rc;
    if (int rc = dev_alloc(h, &d_cls, (nv + 3) / 4)) return rc;
    const int blocks = (int)std::min<size_t>((size_t)h->sm_count * 8, ((nv + 3) / 4 + 255) / 256);
    if (label_bytes == 1)
      class_pack_kernel<uint8_t><<<blocks, 256, 0, h->stream>>>(dl, nv, d_cls_of, d_cls);
    else
      class_pack_kernel<uint16_t><<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const uint16_t *>(dl), nv, d_cls_of, d_cls);
    CU(cudaGetLastError());
    P.cls = d_cls;
    P.cmax = d_cmax;
  }
  CU(cudaStreamSynchronize(h->stream));
  h->d_labels = dl;
  P.labels = dl;
  for (int a = 0; a < 3; ++a) {
    P.n[a] = size[a];
    P.sp[a] = (float)spacing[a];
    P.inv_sp[a] = (float)(1.0 / spacing[a]);
    P.hs[a] = (float)(0.5 * size[a]);
    P.half[a] = (float)(0.5 * size[a] * spacing[a]);
  }
  affine_from_double(P.w2v, world_to_volume);
  P.has_vol = 1;
  h->have_volume = true;
  h->label_bytes = label_bytes;
  h->label_total_bytes = nv * (size_t)label_bytes;
  return GATE_B200_OK;
}

static int check_volume_args(gate_b200_engine *h, int32_t label_bytes, const int32_t size[3], size_t *nv) {
  if (!h->have_tables) return fail(h, GATE_B200_ESTATE, "set_image_volume before set_tables");
  if (label_bytes != 1 && label_bytes != 2) return fail(h, GATE_B200_EINVAL, "label_bytes must be 1 or 2");
  if (!h->dose.empty()) return fail(h, GATE_B200_ESTATE, "set_image_volume after add_dose_actor");
  *nv = (size_t)size[0] * size[1] * size[2];
  if (*nv == 0) return fail(h, GATE_B200_EINVAL, "empty image volume");
  if (*nv >= (1ull << 31)) return fail(h, GATE_B200_EINVAL, "image volume must have fewer than 2^31 voxels");
  return 0;
}

int gate_b200_set_image_volume(gate_b200_engine *h, const void *labels, int32_t label_bytes,
                               const int32_t size[3], const double spacing[3],
                               const double world_to_volume[12]) {
  if (!h || !labels || !size || !spacing || !world_to_volume) return GATE_B200_EINVAL;
  size_t nv = 0;
  if (int rc = check_volume_args(h, label_bytes, size, &nv)) return rc;
  CU(cudaSetDevice(h->device));
  // the label volume goes to the device first (stream order; the caller's buffer may be pinned)
  uint8_t *dl = nullptr;
  if (int rc = dev_alloc(h, &dl, nv * label_bytes)) return rc;
  CU(cudaMemcpyAsync(dl, labels, nv * label_bytes, cudaMemcpyHostToDevice, h->stream));
  return finish_image_volume(h, dl, labels, label_bytes, size, spacing, world_to_volume);
}

int gate_b200_set_image_volume_hu(gate_b200_engine *h, const void *hu, int32_t hu_type, const int32_t size[3],
                                  const double spacing[3], const double world_to_volume[12], int32_t n_bins,
                                  const double *bins, const uint16_t *bin_labels, int32_t label_bytes,
                                  void *labels_out) {
  if (!h || !hu || !size || !spacing || !world_to_volume || !bins || !bin_labels) return GATE_B200_EINVAL;
  if (hu_type < GATE_B200_HU_INT16 || hu_type > GATE_B200_HU_FLOAT64) return fail(h, GATE_B200_EINVAL, "bad hu_type %d", hu_type);
  if (n_bins < 1 || n_bins > kMaxHuBins) return fail(h, GATE_B200_EINVAL, "n_bins must be 1..%d", kMaxHuBins);
  for (int i = 1; i < n_bins; ++i)
    if (!(bins[i] >= bins[i - 1])) return fail(h, GATE_B200_EINVAL, "bins must be sorted");
  size_t nv = 0;
  if (int rc = check_volume_args(h, label_bytes, size, &nv)) return rc;
  for (int i = 0; i <= n_bins; ++i)
    if (label_bytes == 1 && bin_labels[i] > 255) return fail(h, GATE_B200_EINVAL, "label %d does not fit one byte", (int)bin_labels[i]);
  CU(cudaSetDevice(h->device));
  const size_t hu_bytes = hu_type == GATE_B200_HU_INT16 ? 2 : (hu_type == GATE_B200_HU_FLOAT32 ? 4 : 8);
  uint8_t *d_hu = nullptr, *dl = nullptr;
  double *d_bins = nullptr;
  uint16_t *d_bl = nullptr;
  CU(cudaMallocAsync((void **)&d_hu, nv * hu_bytes, h->stream));          // temporary: freed below in stream order
  CU(cudaMemcpyAsync(d_hu, hu, nv * hu_bytes, cudaMemcpyHostToDevice, h->stream));
  if (int rc = dev_alloc(h, &d_bins, (size_t)n_bins)) return rc;
  if (int rc = dev_alloc(h, &d_bl, (size_t)n_bins + 1)) return rc;
  CU(cudaMemcpyAsync(d_bins, bins, n_bins * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_bl, bin_labels, ((size_t)n_bins + 1) * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
  if (int rc = dev_alloc(h, &dl, nv * label_bytes)) return rc;
  const int blocks = (int)std::min<size_t>((size_t)h->sm_count * 8, (nv + 255) / 256);
  const size_t smem = (size_t)n_bins * sizeof(double) + ((size_t)n_bins + 1) * sizeof(uint16_t);
#define GB2_LAUNCH_HU(HuT, LabT)                                                                                   \
  hu_to_label_kernel<HuT, LabT><<<blocks, 256, smem, h->stream>>>(reinterpret_cast<const HuT *>(d_hu), nv, n_bins, \
                                                                  d_bins, d_bl, reinterpret_cast<LabT *>(dl))
  if (label_bytes == 1) {
    if (hu_type == GATE_B200_HU_INT16) GB2_LAUNCH_HU(int16_t, uint8_t);
    else if (hu_type == GATE_B200_HU_FLOAT32) GB2_LAUNCH_HU(float, uint8_t);
    else GB2_LAUNCH_HU(double, uint8_t);
  } else {
    if (hu_type == GATE_B200_HU_INT16) GB2_LAUNCH_HU(int16_t, uint16_t);
    else if (hu_type == GATE_B200_HU_FLOAT32) GB2_LAUNCH_HU(float, uint16_t);
    else GB2_LAUNCH_HU(double, uint16_t);
  }
#undef GB2_LAUNCH_HU
  CU(cudaGetLastError());
  CU(cudaFreeAsync(d_hu, h->stream));
  if (labels_out) CU(cudaMemcpyAsync(labels_out, dl, nv * label_bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return finish_image_volume(h, dl, nullptr, label_bytes, size, spacing, world_to_volume);
}

int gate_b200_add_source(gate_b200_engine *h, const gate_b200_source *s, int32_t *source_id) {
  if (!h || !s) return GATE_B200_EINVAL;
  CU(cudaSetDevice(h->device));
  SourceRec rec;
  DSource &d = rec.dev;
  memset(&d, 0, sizeof d);
  d.pos_type = s->pos_type; d.dir_type = s->dir_type; d.ene_type = s->ene_type;
  d.dir_relative = s->dir_relative;
  if (d.pos_type < 0 || d.pos_type > GATE_B200_POS_SURFACE_SPHERE) return fail(h, GATE_B200_EINVAL, "bad pos_type %d", d.pos_type);
  if (d.dir_type < 0 || d.dir_type > GATE_B200_DIR_HISTOGRAM) return fail(h, GATE_B200_EINVAL, "bad dir_type %d", d.dir_type);
  if (d.ene_type < 0 || d.ene_type > GATE_B200_ENE_PHSP) return fail(h, GATE_B200_EINVAL, "bad ene_type %d", d.ene_type);
  const bool phsp = d.pos_type == GATE_B200_POS_PHSP;
  if (phsp != (d.ene_type == GATE_B200_ENE_PHSP) || (!phsp && d.dir_type == GATE_B200_DIR_PHSP) ||
      (phsp && d.dir_type != GATE_B200_DIR_PHSP && d.dir_type != GATE_B200_DIR_ISO))
    return fail(h, GATE_B200_EINVAL, "phase-space source: pos_type PHSP goes with ene_type PHSP and dir_type PHSP or ISO");
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      d.pos.m[3 * r + c] = (float)s->pos_rotation[3 * r + c];
      d.dir_rot[3 * r + c] = (float)s->dir_rotation[3 * r + c];
    }
    d.pos.t[r] = (float)s->pos_translation[r];
    d.mom[r] = (float)s->momentum[r];
    d.focus[r] = (float)s->focus_point[r];
    d.half[r] = (float)s->half_size[r];
    d.vox_n[r] = s->vox_size[r];
    d.vox_sp[r] = (float)s->vox_spacing[r];
    d.vox_org[r] = (float)s->vox_origin[r];
  }
  d.radius = (float)s->radius; d.sigma_x = (float)s->sigma_x; d.sigma_y = (float)s->sigma_y;
  d.cos_tmin = (float)cos(s->theta[0]); d.cos_tmax = (float)cos(s->theta[1]);
  d.sin2_tmin = (float)(sin(s->theta[0]) * sin(s->theta[0])); d.sin2_tmax = (float)(sin(s->theta[1]) * sin(s->theta[1]));
  d.dir_sx = (float)s->dir_sigma[0]; d.dir_sy = (float)s->dir_sigma[1];
  d.phi0_turn = (float)(s->phi[0] / 6.283185307179586);
  d.dphi_turn = (float)((s->phi[1] - s->phi[0]) / 6.283185307179586);
  d.mono = (float)s->mono_MeV; d.sigma = (float)s->sigma_MeV;
  d.emin = (float)s->e_min_MeV; d.emax = (float)s->e_max_MeV;
  d.weight = (float)s->weight;
  d.n_spec = s->n_spectrum;
  int ne = 0;
  if (s->ene_type == GATE_B200_ENE_SPECTRUM_DISCRETE || s->ene_type == GATE_B200_ENE_CDF) ne = s->n_spectrum;
  if (s->ene_type == GATE_B200_ENE_SPECTRUM_HISTOGRAM || s->ene_type == GATE_B200_ENE_SPECTRUM_HISTOGRAM_LINEAR)
    ne = s->n_spectrum + 1;
  if (ne > 0) {
    if (!s->spectrum_energy || !s->spectrum_cdf || s->n_spectrum < 1)
      return fail(h, GATE_B200_EINVAL, "spectrum arrays missing");
    std::vector<float> e(ne), c(s->n_spectrum);
    for (int i = 0; i < ne; ++i) e[i] = (float)s->spectrum_energy[i];
    for (int i = 0; i < s->n_spectrum; ++i) c[i] = (float)s->spectrum_cdf[i];
    if (int rc = dev_upload(h, &d.spec_e, e.data(), e.size())) return rc;
    if (int rc = dev_upload(h, &d.spec_cdf, c.data(), c.size())) return rc;
  }
  if (s->pos_type == GATE_B200_POS_VOXELS) {
    const size_t nx = s->vox_size[0], ny = s->vox_size[1], nz = s->vox_size[2];
    if (!s->cdf_x || !s->cdf_y || !s->cdf_z || nx * ny * nz == 0)
      return fail(h, GATE_B200_EINVAL, "voxel source needs cdf_x/y/z");
    if (int rc = dev_upload(h, &d.cdf_z, s->cdf_z, nz)) return rc;
    if (int rc = dev_upload(h, &d.cdf_y, s->cdf_y, nz * ny)) return rc;
    if (int rc = dev_upload(h, &d.cdf_x, s->cdf_x, nz * ny * nx)) return rc;
  }
  if (d.dir_type == GATE_B200_DIR_HISTOGRAM) {
    if (s->n_dir_theta < 1 || s->n_dir_phi < 1 || !s->dir_theta_edges || !s->dir_theta_cdf || !s->dir_phi_edges || !s->dir_phi_cdf)
      return fail(h, GATE_B200_EINVAL, "direction histogram: theta and phi edges + cdf are required");
    const double *src4[4] = {s->dir_theta_edges, s->dir_theta_cdf, s->dir_phi_edges, s->dir_phi_cdf};
    const float **dst4[4] = {&d.ht_e, &d.ht_c, &d.hp_e, &d.hp_c};
    for (int k = 0; k < 4; ++k) {
      const int n = (k < 2 ? s->n_dir_theta : s->n_dir_phi) + 1;
      std::vector<float> v(n);
      for (int i = 0; i < n; ++i) v[i] = (float)src4[k][i];
      if (int rc = dev_upload(h, dst4[k], v.data(), v.size())) return rc;
    }
    d.n_ht = s->n_dir_theta; d.n_hp = s->n_dir_phi;
  }
  if (s->confine) {
    if (phsp) return fail(h, GATE_B200_EINVAL, "a phase-space source cannot be confined");
    d.confine = 1;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) d.conf.m[3 * r + c] = (float)s->world_to_confine[4 * r + c];
      d.conf.t[r] = (float)s->world_to_confine[4 * r + 3];
      d.conf_half[r] = (float)s->confine_half_size[r];
    }
  }
  if (phsp) {
    const uint64_t n = s->phsp_n_entries;
    const bool need_dir = d.dir_type == GATE_B200_DIR_PHSP;
    if (n == 0 || !s->phsp_pos_x || !s->phsp_pos_y || !s->phsp_pos_z || !s->phsp_energy ||
        (need_dir && (!s->phsp_dir_x || !s->phsp_dir_y || !s->phsp_dir_z)))
      return fail(h, GATE_B200_EINVAL, "phase-space source: position, energy%s columns and n_entries > 0 are required",
                  need_dir ? ", direction" : "");
    // two 16-byte records per entry: one coalesced load each for consecutive events
    std::vector<float4> a(n), b(n);
    for (uint64_t i = 0; i < n; ++i) {
      a[i] = make_float4(s->phsp_pos_x[i], s->phsp_pos_y[i], s->phsp_pos_z[i], s->phsp_energy[i]);
      b[i] = make_float4(need_dir ? s->phsp_dir_x[i] : 0.0f, need_dir ? s->phsp_dir_y[i] : 0.0f,
                         need_dir ? s->phsp_dir_z[i] : 1.0f, s->phsp_weight ? s->phsp_weight[i] : 1.0f);
      if (need_dir && !(b[i].x * b[i].x + b[i].y * b[i].y + b[i].z * b[i].z > 0.0f))
        return fail(h, GATE_B200_EINVAL, "phase-space source: entry %llu has a null direction", (unsigned long long)i);
    }
    if (int rc = dev_upload(h, &d.ph_a, a.data(), (size_t)n)) return rc;
    if (int rc = dev_upload(h, &d.ph_b, b.data(), (size_t)n)) return rc;
    d.ph_n = n;
    d.ph_start = s->phsp_entry_start % n;
  }
  h->src.push_back(rec);
  if (source_id) *source_id = (int32_t)h->src.size() - 1;
  return GATE_B200_OK;
}

int gate_b200_add_dose_actor(gate_b200_engine *h, const gate_b200_dose_desc *d, int32_t *actor_id) {
  if (!h || !d) return GATE_B200_EINVAL;
  if (!h->have_volume) return fail(h, GATE_B200_ESTATE, "add_dose_actor needs the image volume first");
  if ((int)h->dose.size() >= kMaxDose) return fail(h, GATE_B200_EINVAL, "at most %d dose actors", kMaxDose);
  const size_t nv = (size_t)d->size[0] * d->size[1] * d->size[2];
  if (nv == 0 || nv >= (1u << 28)) return fail(h, GATE_B200_EINVAL, "dose grid must have 1..2^28-1 voxels");
  if (d->hit_type < 0 || d->hit_type > GATE_B200_HIT_MIDDLE) return fail(h, GATE_B200_EINVAL, "bad hit_type");
  CU(cudaSetDevice(h->device));
  DoseActor A;
  A.desc = *d;
  A.nvox = nv;
  DDose &D = A.dev;
  memset(&D, 0, sizeof D);
  for (int a = 0; a < 3; ++a) D.size[a] = d->size[a];
  D.flags = d->score_flags | GATE_B200_SCORE_EDEP;
  D.hit_type = d->hit_type;
  // volume -> index = (world -> index) o (volume -> world); composed in double from the float
  // world->volume transform (identical to the oracle's oracle_add_dose_actor)
  const Affine &w = h->P.w2v;
  double w2v[12];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) w2v[4 * r + c] = (double)w.m[3 * r + c];
    w2v[4 * r + 3] = (double)w.t[r];
  }
  double v2w[12];
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
  affine_from_double(D.v2i, comp);
  // the grid is the image grid when volume point -> index is p / spacing + n / 2 - 1 / 2 on the image's own
  // lattice (same test as the oracle's grid_is_image_grid): the voxel of a TLE chord is then the image voxel, and
  // the point a hit type picks on a Geant4 step (which never leaves its voxel) cannot change the voxel scored
  {
    const Params &P = h->P;
    D.aligned = 1;
    for (int a = 0; a < 3; ++a) {
      if (D.size[a] != P.n[a]) D.aligned = 0;
      for (int c = 0; c < 3; ++c) {
        const double wnt = a == c ? 1.0 / (double)P.sp[a] : 0.0;
        if (fabs((double)D.v2i.m[3 * a + c] - wnt) > 1.0e-6 * (1.0 / (double)P.sp[a])) D.aligned = 0;
      }
      if (fabs((double)D.v2i.t[a] - (0.5 * (double)P.n[a] - 0.5)) > 1.0e-4) D.aligned = 0;
    }
  }
  const int want[5] = {GATE_B200_SCORE_EDEP, GATE_B200_SCORE_EDEP_SQ, GATE_B200_SCORE_DOSE,
                       GATE_B200_SCORE_DOSE_SQ, GATE_B200_SCORE_COUNTS};
  // one 32-byte record per voxel for the four sums (a deposit touches one DRAM sector instead of
  // one per grid); the counts image, rarely asked for, is a plain array
  int n_active = 0;
  for (int k = 0; k < 4; ++k) n_active += (D.flags & want[k]) ? 1 : 0;
  D.gstride = n_active <= 1 ? 1 : (n_active == 2 ? 2 : 4);      // a record never straddles a sector
  if (int rc = dev_alloc(h, &A.rec, nv * (size_t)D.gstride, true)) return rc;
  for (int k = 0, slot = 0; k < 4; ++k)
    if (D.flags & want[k]) D.grid[k] = A.rec + slot++;
  if (D.flags & want[4]) {
    if (int rc = dev_alloc(h, &D.grid[4], nv, true)) return rc;
  }
  h->dose.push_back(A);
  h->P.dose[h->dose.size() - 1] = D;
  h->P.ndose = (int)h->dose.size();
  if (actor_id) *actor_id = (int32_t)h->dose.size() - 1;
  return GATE_B200_OK;
}

int gate_b200_add_tle_dose_actor(gate_b200_engine *h, const gate_b200_dose_desc *d, const gate_b200_tle_desc *t,
                                 int32_t *actor_id) {
  if (!h || !d || !t) return GATE_B200_EINVAL;
  if (!t->mu_en_over_rho) return fail(h, GATE_B200_EINVAL, "tle: mu_en_over_rho table missing");
  if (!h->have_volume) return fail(h, GATE_B200_ESTATE, "add_tle_dose_actor needs the image volume first");
  Params &P = h->P;
  if (P.n[0] > 65535 || P.n[1] > 65535) return fail(h, GATE_B200_EINVAL, "tle: image volume wider than 65535 voxels");
  int32_t id = -1;
  if (int rc = gate_b200_add_dose_actor(h, d, &id)) return rc;
  DDose &D = h->dose[id].dev;
  D.tle = 1;
  D.tle_thr = t->energy_threshold_MeV > 3.0e38 ? kBig : (float)t->energy_threshold_MeV;
  {
    // energy-major pairs, like the attenuation table (same float values as the caller's array)
    const size_t n = (size_t)P.nmat * P.ne;
    std::vector<float2> m2(n);
    for (int ie = 0; ie < P.ne; ++ie) {
      const int ie1 = ie + 1 < P.ne ? ie + 1 : ie;
      for (int m = 0; m < P.nmat; ++m)
        m2[(size_t)ie * P.nmat + m] = make_float2(t->mu_en_over_rho[(size_t)m * P.ne + ie], t->mu_en_over_rho[(size_t)m * P.ne + ie1]);
    }
    if (int rc = dev_upload(h, &D.muen2, m2.data(), n)) return rc;
  }
  const float emin = (float)t->energy_min_MeV;
  if (P.ntle == 0 || emin > P.tle_emin) P.tle_emin = emin;
  if (P.ntle == 0 || D.tle_thr > P.tle_thr_max) P.tle_thr_max = D.tle_thr;
  P.ntle++;
  // the per-slot scratch of the TLE phase is allocated by the first launch, when the kernel variant is known (tle_prepare)
  P.dose[id] = D;
  if (actor_id) *actor_id = id;
  return GATE_B200_OK;
}

int gate_b200_add_phsp_actor(gate_b200_engine *h, const gate_b200_phsp_desc *d, int32_t *actor_id) {
  if (!h || !d) return GATE_B200_EINVAL;
  if ((int)h->phsp.size() >= kMaxPhsp) return fail(h, GATE_B200_EINVAL, "at most %d phase-space actors", kMaxPhsp);
  CU(cudaSetDevice(h->device));
  PhspActor A;
  A.desc = *d;
  DPhsp &Q = A.dev;
  memset(&Q, 0, sizeof Q);
  for (int a = 0; a < 3; ++a) Q.half[a] = (float)d->half_size[a];
  affine_from_double(Q.w2b, d->world_to_box);
  Q.steps = d->steps_to_store;
  Q.capacity = d->capacity;
  const size_t cap = (size_t)d->capacity;
  if (int rc = dev_alloc(h, &Q.ke, cap)) return rc;
  if (int rc = dev_alloc(h, &Q.pos, 3 * cap)) return rc;
  if (int rc = dev_alloc(h, &Q.dir, 3 * cap)) return rc;
  if (int rc = dev_alloc(h, &Q.w, cap)) return rc;
  if (int rc = dev_alloc(h, &Q.t, cap)) return rc;
  if (int rc = dev_alloc(h, &Q.ev, cap)) return rc;
  if (int rc = dev_alloc(h, &Q.fl, cap)) return rc;
  if (int rc = dev_alloc(h, &Q.n_rows, 1, true)) return rc;
  if (int rc = dev_alloc(h, &Q.n_dropped, 1, true)) return rc;
  h->phsp.push_back(A);
  h->P.phsp[h->phsp.size() - 1] = Q;
  h->P.nphsp = (int)h->phsp.size();
  if (actor_id) *actor_id = (int32_t)h->phsp.size() - 1;
  return GATE_B200_OK;
}

int gate_b200_add_box_region(gate_b200_engine *h, const gate_b200_region_desc *d, int32_t *region_id) {
  if (!h || !d) return GATE_B200_EINVAL;
  Params &P = h->P;
  if (P.nreg >= kMaxReg) return fail(h, GATE_B200_EINVAL, "at most %d box regions", kMaxReg);
  if (!h->have_tables) return fail(h, GATE_B200_ESTATE, "add_box_region before set_tables");
  if (d->material < 0 || d->material >= P.nmat) return fail(h, GATE_B200_EINVAL, "box region: material %d is not in the tables", d->material);
  for (int a = 0; a < 3; ++a)
    if (!(d->half_size[a] > 0.0)) return fail(h, GATE_B200_EINVAL, "box region: half sizes must be positive");
  DRegion &R = P.reg[P.nreg];
  for (int a = 0; a < 3; ++a) R.half[a] = (float)d->half_size[a];
  affine_from_double(R.w2b, d->world_to_box);
  R.mat = d->material;
  if (region_id) *region_id = P.nreg;
  P.nreg++;
  return GATE_B200_OK;
}

int gate_b200_begin_run(gate_b200_engine *h, int32_t /*run_id*/) {
  if (!h) return GATE_B200_EINVAL;
  CU(cudaSetDevice(h->device));
  for (auto &A : h->dose) {
    CU(cudaMemsetAsync(A.rec, 0, A.nvox * (size_t)A.dev.gstride * sizeof(double), h->stream));
    if (A.dev.grid[GATE_B200_GRID_COUNTS])
      CU(cudaMemsetAsync(A.dev.grid[GATE_B200_GRID_COUNTS], 0, A.nvox * sizeof(double), h->stream));
  }
  for (auto &A : h->phsp) {
    CU(cudaMemsetAsync(A.dev.n_rows, 0, sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(A.dev.n_dropped, 0, sizeof(unsigned long long), h->stream));
  }
  CU(cudaMemsetAsync(h->d_stats, 0, S_COUNT * sizeof(unsigned long long), h->stream));
  CU(cudaMemsetAsync(h->d_edep, 0, sizeof(double), h->stream));
  if (int rc = collect_kernel_time(h)) return rc;
  h->kernel_ms = 0.0;
  return GATE_B200_OK;
}

// Kernel variant of the track-length estimator and its per-slot scratch.  Variant 5 (no per-history tables: the sums of
// a history per voxel are re-evaluated from its flight records) serves ONE TLE actor that scores event-wise squares on
// the image grid -- next to any number of analog dose actors; every other combination runs variant 1, whose actors
// with squares keep a 64 KB table per photon slot.
static int tle_mode(const Params &P) {
  if (P.ntle != 1 || getenv("GATE_B200_TLE_TABLE")) return 1;
  for (int a = 0; a < P.ndose; ++a) {
    const DDose &D = P.dose[a];
    if (D.tle) return (D.aligned && (D.flags & (GATE_B200_SCORE_EDEP_SQ | GATE_B200_SCORE_DOSE_SQ)) && P.n[2] <= 65535) ? 5 : 1;
  }
  return 1;
}
static int tle_prepare(gate_b200_engine *h) {
  Params &P = h->P;
  if (P.ntle == 0) return GATE_B200_OK;
  const size_t nslots = (size_t)h->max_ctas * kTleSlots;   // the TLE kernels work with kTleSlots slots per CTA
  if (!P.tle_state) {
    if (int rc = dev_alloc(h, &P.tle_state, nslots, true)) return rc;
    if (int rc = dev_alloc(h, &P.tle_aux, nslots, true)) return rc;
  }
  bool sq = false;
  for (int a = 0; a < P.ndose; ++a)
    if (P.dose[a].tle && (P.dose[a].flags & (GATE_B200_SCORE_EDEP_SQ | GATE_B200_SCORE_DOSE_SQ))) { sq = true; P.tle_a = a; }
  if (tle_mode(P) == 5) {
    if (!P.tle_rec) {
      if (int rc = dev_alloc(h, &P.tle_rec, nslots * kTleRecMax * kTleRecW)) return rc;
    }
  } else if (sq && !P.tle_tab) {
    // per-history tables (event-wise squares): 64 KB per photon slot, empty = key -1
    if (int rc = dev_alloc(h, &P.tle_tab, nslots * kTleH)) return rc;
    CU(cudaMemsetAsync(P.tle_tab, 0xff, nslots * kTleH * sizeof(DepEntry), h->stream));
    if (int rc = dev_alloc(h, &P.tle_list, nslots * kTleMax)) return rc;
  }
  return GATE_B200_OK;
}

static int launch_histories(gate_b200_engine *h, int32_t source_id, uint64_t first_history, uint64_t n_histories) {
  if (int rc = tle_prepare(h)) return rc;
  Params P = h->P;
  P.label_bytes = h->label_bytes;
  P.tle_rec_cap = kTleRecMax;
  if (const char *c = getenv("GATE_B200_TLE_REC_CAP")) {   // tests: reach the record-overflow path with short histories
    const int v = atoi(c);
    if (v >= 1 && v <= kTleRecMax) P.tle_rec_cap = v;
  }
  const int label_bytes = h->label_bytes;
  P.src = h->src[source_id].dev;
  P.first = first_history;
  P.count = n_histories;
  CU(cudaMemsetAsync(h->d_next, 0, sizeof(unsigned long long), h->stream));
  const size_t smem = sizeof(Smem) + (size_t)P.ne * sizeof(float);
  // kernel variant: 1 = with the TLE phase; 2 / 3 = the common dose run (one dose actor, no phase-space actor
  // or source) with those choices compiled in, 2 for hit_type post, 3 for the other hit types; 4 = one dose actor
  // with phase-space actors; 0 = everything else
  const bool one_dose = P.ntle == 0 && P.ndose == 1 && P.nreg == 0 && P.src.ph_b == nullptr && !P.src.confine && P.src.dir_type != GATE_B200_DIR_HISTOGRAM &&
                        !getenv("GATE_B200_GENERAL_KERNEL");
  // hit_type random / middle on the image's own grid score the voxel of the interaction, like post (hit_point_is_post)
  const bool post = P.ndose > 0 && (P.dose[0].hit_type == GATE_B200_HIT_POST ||
                                    (P.dose[0].aligned && P.dose[0].hit_type != GATE_B200_HIT_PRE));
  const int mode = P.ntle > 0 ? tle_mode(P) : !one_dose ? 0 : P.nphsp > 0 ? 4 : (post ? 2 : 3);
  auto kern = transport_kernel<uint8_t, 0>;
  if (label_bytes == 2) {
    kern = mode == 1 ? transport_kernel<uint16_t, 1> : mode == 2 ? transport_kernel<uint16_t, 2>
         : mode == 3 ? transport_kernel<uint16_t, 3> : mode == 4 ? transport_kernel<uint16_t, 4>
         : mode == 5 ? transport_kernel<uint16_t, 5> : transport_kernel<uint16_t, 0>;
  } else {
    kern = mode == 1 ? transport_kernel<uint8_t, 1> : mode == 2 ? transport_kernel<uint8_t, 2>
         : mode == 3 ? transport_kernel<uint8_t, 3> : mode == 4 ? transport_kernel<uint8_t, 4>
         : mode == 5 ? transport_kernel<uint8_t, 5> : transport_kernel<uint8_t, 0>;
  }
  {
    char nm[96];
    snprintf(nm, sizeof nm, "gb2::transport_kernel<unsigned %s, %d>", label_bytes == 2 ? "short" : "char", mode);
    h->kernel_name = nm;
  }
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent CTAs: one per SM, fewer only when there is not enough work to fill the slots
  unsigned long long want = (n_histories + 255) / 256;
  unsigned long long grid = std::min<unsigned long long>((unsigned long long)h->max_ctas, want);
  if (grid < 1) grid = 1;
  // Keep the label volume resident in L2 (persisting access window): the random 1-byte gathers of
  // the Woodcock steps then hit L2 while the fp64 grid atomics stream through the rest of it.
  cudaLaunchAttribute attr[1];
  int nattr = 0;
  if (h->l2_persist && h->d_labels && h->l2_persist_max > 0 && h->l2_window_max > 0 && h->label_total_bytes > (1u << 20)) {
    const size_t win = std::min(h->label_total_bytes, h->l2_window_max);
    const size_t persist = std::min(win, h->l2_persist_max);
    if (h->l2_limit_set != persist) {   // the limit is per DEVICE: remembered in the handle, not per thread
      CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
      h->l2_limit_set = persist;
    }
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = h->d_labels;
    attr[0].val.accessPolicyWindow.num_bytes = win;
    attr[0].val.accessPolicyWindow.hitRatio = (float)((double)persist / (double)win);
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    nattr = 1;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(P.ntle > 0 ? kTleThreads : kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, h->stream));
  CU(cudaLaunchKernelEx(&cfg, kern, P));
  CU(cudaGetLastError());
  CU(cudaEventRecord(e1, h->stream));
  h->pending_events.emplace_back(e0, e1);
  return GATE_B200_OK;
}

int gate_b200_run_source(gate_b200_engine *h, int32_t source_id, uint64_t first_history, uint64_t n_histories) {
  if (!h) return GATE_B200_EINVAL;
  if (!h->have_tables) return fail(h, GATE_B200_ESTATE, "run_source before set_tables");
  if (source_id < 0 || source_id >= (int)h->src.size()) return fail(h, GATE_B200_EINVAL, "unknown source %d", source_id);
  if (n_histories == 0) return GATE_B200_OK;
  CU(cudaSetDevice(h->device));
  // a slot stores the history index inside its launch in 32 bits: longer requests become several launches
  // (and at most 2^24 per persistent CTA: the 32-bit per-CTA event counters of the kernel cannot overflow then)
  const uint64_t kMaxPerLaunch = std::min<uint64_t>(1ull << 31, (uint64_t)std::max(h->max_ctas, 1) << 24);
  for (uint64_t done = 0; done < n_histories;) {
    const uint64_t k = std::min<uint64_t>(kMaxPerLaunch, n_histories - done);
    if (int rc = launch_histories(h, source_id, first_history + done, k)) return rc;
    done += k;
  }
  return GATE_B200_OK;
}

int gate_b200_sync(gate_b200_engine *h) {
  if (!h) return GATE_B200_EINVAL;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  return collect_kernel_time(h);
}

int gate_b200_timed_run(gate_b200_engine *h, int32_t source_id, uint64_t first_history,
                        uint64_t n_histories, double *device_ms) {
  if (!h) return GATE_B200_EINVAL;
  if (int rc = gate_b200_sync(h)) return rc;
  const double before = h->kernel_ms;
  if (int rc = gate_b200_run_source(h, source_id, first_history, n_histories)) return rc;
  if (int rc = gate_b200_sync(h)) return rc;
  if (device_ms) *device_ms = h->kernel_ms - before;
  return GATE_B200_OK;
}

static int check_grid(gate_b200_engine *h, int32_t actor_id, int32_t which, double **p, size_t *n) {
  if (actor_id < 0 || actor_id >= (int)h->dose.size()) return fail(h, GATE_B200_EINVAL, "unknown dose actor %d", actor_id);
  if (which < 0 || which > 4) return fail(h, GATE_B200_EINVAL, "unknown grid %d", which);
  DoseActor &A = h->dose[actor_id];
  if (!A.dev.grid[which]) return fail(h, GATE_B200_EINVAL, "grid %d of actor %d is not scored", which, actor_id);
  *p = A.dev.grid[which];
  *n = A.nvox;
  return 0;
}

// dense staging array for get/set_dose_grid (kept for the life of the engine, grown on demand)
static int staging(gate_b200_engine *h, size_t n, double **out) {
  if (h->stage_elems < n) {
    if (h->d_stage) {   // the outgrown one goes back to the pool in stream order
      h->allocs.erase(std::remove(h->allocs.begin(), h->allocs.end(), (void *)h->d_stage), h->allocs.end());
      CU(cudaFreeAsync(h->d_stage, h->stream));
      h->d_stage = nullptr;
      h->stage_elems = 0;
    }
    if (int rc = dev_alloc(h, &h->d_stage, n)) return rc;
    h->stage_elems = n;
  }
  *out = h->d_stage;
  return 0;
}

int gate_b200_get_dose_grid_range(gate_b200_engine *h, int32_t actor_id, int32_t which, double scale, uint64_t first,
                                  uint64_t count, double *out) {
  if (!h || !out) return GATE_B200_EINVAL;
  double *p; size_t n;
  if (int rc = check_grid(h, actor_id, which, &p, &n)) return rc;
  if (first > n || count > n - first) return fail(h, GATE_B200_EINVAL, "grid range [%llu, +%llu) exceeds %zu voxels", (unsigned long long)first, (unsigned long long)count, n);
  if (count == 0) return GATE_B200_OK;
  CU(cudaSetDevice(h->device));
  const size_t stride = which == GATE_B200_GRID_COUNTS ? 1 : (size_t)h->dose[actor_id].dev.gstride;
  p += (size_t)first * stride;
  if (stride > 1 || scale != 1.0) {
    double *dense;
    if (int rc = staging(h, (size_t)count, &dense)) return rc;
    grid_extract_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(p, (size_t)count, stride, scale, dense);
    CU(cudaGetLastError());
    p = dense;
  }
  CU(cudaMemcpyAsync(out, p, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return GATE_B200_OK;
}

int gate_b200_get_dose_grid_scaled(gate_b200_engine *h, int32_t actor_id, int32_t which, double scale, double *out) {
  if (!h) return GATE_B200_EINVAL;
  if (actor_id < 0 || actor_id >= (int)h->dose.size()) return fail(h, GATE_B200_EINVAL, "unknown dose actor %d", actor_id);
  return gate_b200_get_dose_grid_range(h, actor_id, which, scale, 0, (uint64_t)h->dose[actor_id].nvox, out);
}

int gate_b200_get_dose_grid(gate_b200_engine *h, int32_t actor_id, int32_t which, double *out) {
  return gate_b200_get_dose_grid_scaled(h, actor_id, which, 1.0, out);
}

int gate_b200_get_uncertainty(gate_b200_engine *h, int32_t actor_id, int32_t which, uint64_t n_events, double *out) {
  if (!h || !out) return GATE_B200_EINVAL;
  if (which != GATE_B200_GRID_EDEP && which != GATE_B200_GRID_DOSE) return fail(h, GATE_B200_EINVAL, "uncertainty of grid %d: edep or dose only", which);
  if (n_events < 2) return fail(h, GATE_B200_EINVAL, "uncertainty needs more than one event");
  double *v, *sq; size_t n;
  if (int rc = check_grid(h, actor_id, which, &v, &n)) return rc;
  if (int rc = check_grid(h, actor_id, which + 1, &sq, &n)) return rc;   // the squares follow their values in the enum
  CU(cudaSetDevice(h->device));
  double *dense;
  if (int rc = staging(h, n, &dense)) return rc;
  grid_uncertainty_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(v, sq, n, (size_t)h->dose[actor_id].dev.gstride, (double)n_events, dense);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, dense, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return GATE_B200_OK;
}

// Page-lock the caller's array.  Most of the time of cudaHostRegister on a freshly allocated array is the first touch
// of its pages (one fault and one zeroed page per 4 KB, serial inside the driver call: 0.3 ... 0.9 s for the 0.6 GB of
// one config-3 image, depending on the state of the box).  The pages are therefore touched first, from several
// threads; the contents are left as they are.  (No MADV_HUGEPAGE hint: with defrag = madvise the fault path may
// compact memory synchronously.)
int gate_b200_host_register(void *ptr, uint64_t bytes) {
  if (!ptr || !bytes) return GATE_B200_EINVAL;
  if (bytes >= (8u << 20)) {
    const uintptr_t a0 = (uintptr_t)ptr, a1 = a0 + (uintptr_t)bytes;
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
    const uintptr_t chunk = (((a1 - a0) / nt) + 4095) & ~(uintptr_t)4095;
    std::vector<std::thread> th;
    for (unsigned k = 0; k < nt; ++k) {
      const uintptr_t b0 = a0 + k * chunk, b1 = std::min(a1, b0 + chunk);
      if (b0 >= b1) break;
      th.emplace_back([b0, b1]() {
        for (uintptr_t p = b0; p < b1; p += 4096) {
          volatile char *q = (volatile char *)p;
          const char c = *q;
          *q = c;                                   // a write fault maps the page; the byte keeps its value
        }
      });
    }
    for (auto &x : th) x.join();
  }
  return cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable) == cudaSuccess ? GATE_B200_OK : (cudaGetLastError(), GATE_B200_ECUDA);
}
int gate_b200_host_unregister(void *ptr) {
  if (!ptr) return GATE_B200_EINVAL;
  return cudaHostUnregister(ptr) == cudaSuccess ? GATE_B200_OK : (cudaGetLastError(), GATE_B200_ECUDA);
}

int gate_b200_set_dose_grid(gate_b200_engine *h, int32_t actor_id, int32_t which, const double *in) {
  if (!h || !in) return GATE_B200_EINVAL;
  double *p; size_t n;
  if (int rc = check_grid(h, actor_id, which, &p, &n)) return rc;
  CU(cudaSetDevice(h->device));
  const size_t stride = (size_t)h->dose[actor_id].dev.gstride;
  if (which == GATE_B200_GRID_COUNTS || stride == 1) {
    CU(cudaMemcpyAsync(p, in, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  } else {
    double *dense;
    if (int rc = staging(h, n, &dense)) return rc;
    CU(cudaMemcpyAsync(dense, in, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    grid_scatter_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(p, n, stride, dense);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(h->stream));
  return GATE_B200_OK;
}

int gate_b200_dose_buffers(gate_b200_engine *h, int32_t actor_id, int32_t max_buffers, void **ptrs,
                           uint64_t *n_elements, int32_t *n_buffers) {
  if (!h || !ptrs || !n_elements || !n_buffers) return GATE_B200_EINVAL;
  if (actor_id < 0 || actor_id >= (int)h->dose.size()) return fail(h, GATE_B200_EINVAL, "unknown dose actor %d", actor_id);
  DoseActor &A = h->dose[actor_id];
  int nb = 0;
  if (nb < max_buffers) { ptrs[nb] = A.rec; n_elements[nb] = (uint64_t)A.nvox * (uint64_t)A.dev.gstride; ++nb; }
  if (A.dev.grid[GATE_B200_GRID_COUNTS] && nb < max_buffers) {
    ptrs[nb] = A.dev.grid[GATE_B200_GRID_COUNTS]; n_elements[nb] = (uint64_t)A.nvox; ++nb;
  }
  *n_buffers = nb;
  return GATE_B200_OK;
}

int gate_b200_phsp_count(gate_b200_engine *h, int32_t actor_id, uint64_t *n_rows) {
  if (!h || !n_rows) return GATE_B200_EINVAL;
  if (actor_id < 0 || actor_id >= (int)h->phsp.size()) return fail(h, GATE_B200_EINVAL, "unknown phsp actor %d", actor_id);
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  unsigned long long n = 0;
  CU(cudaMemcpy(&n, h->phsp[actor_id].dev.n_rows, sizeof n, cudaMemcpyDeviceToHost));
  *n_rows = std::min<unsigned long long>(n, h->phsp[actor_id].dev.capacity);
  return GATE_B200_OK;
}

int gate_b200_get_phsp(gate_b200_engine *h, int32_t actor_id, uint64_t max_rows,
                       const gate_b200_phsp_columns *out, uint64_t *n_rows) {
  if (!h || !out) return GATE_B200_EINVAL;
  uint64_t n = 0;
  if (int rc = gate_b200_phsp_count(h, actor_id, &n)) return rc;
  n = std::min<uint64_t>(n, max_rows);
  const DPhsp &Q = h->phsp[actor_id].dev;
  if (n) {
    if (out->kinetic_energy) CU(cudaMemcpy(out->kinetic_energy, Q.ke, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (out->position) CU(cudaMemcpy(out->position, Q.pos, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
    if (out->direction) CU(cudaMemcpy(out->direction, Q.dir, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
    if (out->weight) CU(cudaMemcpy(out->weight, Q.w, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (out->time_ns) CU(cudaMemcpy(out->time_ns, Q.t, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (out->event_id) CU(cudaMemcpy(out->event_id, Q.ev, n * sizeof(long long), cudaMemcpyDeviceToHost));
    if (out->flags) CU(cudaMemcpy(out->flags, Q.fl, n * sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (n_rows) *n_rows = n;
  return GATE_B200_OK;
}

int gate_b200_get_stats(gate_b200_engine *h, gate_b200_stats *out) {
  if (!h || !out) return GATE_B200_EINVAL;
  if (int rc = gate_b200_sync(h)) return rc;
  unsigned long long s[S_COUNT];
  CU(cudaMemcpy(s, h->d_stats, sizeof s, cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof *out);
  out->histories = s[S_HIST]; out->tracks = s[S_TRACKS]; out->steps = s[S_STEPS];
  out->tentative_collisions = s[S_TENT];
  out->n_phot = s[S_PHOT]; out->n_compt = s[S_COMPT]; out->n_rayl = s[S_RAYL]; out->n_conv = s[S_CONV];
  out->deposits = s[S_DEP]; out->flushed_entries = s[S_FLUSH]; out->buffer_overflows = s[S_OVER];
  out->escaped = s[S_ESC]; out->absorbed_events = s[S_ABSORBED];
  out->tle_steps = s[S_TLE_STEPS]; out->tle_entries = s[S_TLE_ENT]; out->tle_overflows = s[S_TLE_OVER];
  CU(cudaMemcpy(&out->edep_total_MeV, h->d_edep, sizeof(double), cudaMemcpyDeviceToHost));
  for (auto &A : h->phsp) {
    unsigned long long n = 0, d = 0;
    CU(cudaMemcpy(&n, A.dev.n_rows, sizeof n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&d, A.dev.n_dropped, sizeof d, cudaMemcpyDeviceToHost));
    out->phsp_rows += std::min<unsigned long long>(n, A.dev.capacity);
    out->phsp_dropped += d;
  }
  out->kernel_ms = h->kernel_ms;
  return GATE_B200_OK;
}

int gate_b200_mean_uncertainty(gate_b200_engine *h, int32_t actor_id, uint64_t n_events,
                               int32_t top_voxels, double threshold, double *out);

}  // extern "C"

#include "uncertainty.cuh"
