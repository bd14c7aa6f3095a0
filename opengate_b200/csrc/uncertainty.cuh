// uncertainty.cuh — device version of GateDoseActor::ComputeMeanUncertainty
// (reference core/opengate_core/opengate_lib/GateDoseActor.cpp:280-330 and
// GetMeanOfHighestNValues :401-430): mean relative uncertainty over the voxels whose edep
// exceeds threshold * mean(top-N edep).  The top-N mean is found by N "largest value below
// the previous one" reductions (edep >= 0, so the bit pattern of a double orders like the
// value and atomicMax on the 64-bit pattern is exact).
#pragma once

namespace gb2 {

__global__ void max_below_kernel(const double *__restrict__ g, size_t n, size_t stride, double bound,
                                 unsigned long long *out_bits) {
  unsigned long long best = 0ull;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double v = g[i * stride];
    if (v < bound && v > 0.0) {
      unsigned long long b = (unsigned long long)__double_as_longlong(v);
      if (b > best) best = b;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    if (t > best) best = t;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(out_bits, best);
}

__global__ void count_equal_kernel(const double *__restrict__ g, size_t n, size_t stride, double value,
                                   unsigned long long *out_count) {
  unsigned long long c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (g[i * stride] == value) ++c;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out_count, c);
}

__global__ void mean_unc_kernel(const double *__restrict__ e, const double *__restrict__ e2, size_t n, size_t stride,
                                double thr, double nev, double *out_sum, unsigned long long *out_cnt) {
  double s = 0.0;
  unsigned long long c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double val = e[i * stride];
    if (val > thr) {
      val /= nev;
      double sq = e2[i * stride] / nev;
      double u = (1.0 / (nev - 1.0)) * (sq - val * val);
      if (u < 0.0) u = 0.0;
      u = sqrt(u) / val;
      if (u > 1.0) u = 1.0;
      s += u;
      ++c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0 && c) { atomicAdd(out_sum, s); atomicAdd(out_cnt, c); }
}

}  // namespace gb2

extern "C" int gate_b200_mean_uncertainty(gate_b200_engine *h, int32_t actor_id, uint64_t n_events,
                                          int32_t top_voxels, double threshold, double *out) {
  if (!h || !out) return GATE_B200_EINVAL;
  if (actor_id < 0 || actor_id >= (int)h->dose.size()) return fail(h, GATE_B200_EINVAL, "unknown dose actor %d", actor_id);
  DoseActor &A = h->dose[actor_id];
  const double *e = A.dev.grid[GATE_B200_GRID_EDEP], *e2 = A.dev.grid[GATE_B200_GRID_EDEP_SQ];
  if (!e || !e2) return fail(h, GATE_B200_EINVAL, "mean uncertainty needs edep and edep_squared");
  CU(cudaSetDevice(h->device));
  // scratch from the engine's stream-ordered pool, allocated once and freed with the engine (an early return
  // on a CUDA error leaks nothing)
  if (!h->d_unc_u) {
    if (int rc = dev_alloc(h, &h->d_unc_u, 2)) return rc;
    if (int rc = dev_alloc(h, &h->d_unc_s, 1)) return rc;
  }
  unsigned long long *d_u = h->d_unc_u;
  double *d_s = h->d_unc_s;
  const int blocks = h->sm_count * 8, threads = 256;
  double nev = (double)n_events;
  if (nev < 2.0) nev = 2.0;
  // mean of the N highest values
  double bound = INFINITY, sum = 0.0;
  long long taken = 0;
  while (taken < top_voxels) {
    CU(cudaMemsetAsync(d_u, 0, 2 * sizeof(unsigned long long), h->stream));
    gb2::max_below_kernel<<<blocks, threads, 0, h->stream>>>(e, A.nvox, (size_t)A.dev.gstride, bound, d_u);
    unsigned long long bits = 0;
    CU(cudaMemcpyAsync(&bits, d_u, sizeof bits, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (bits == 0) break;  // only zeros left: they add nothing to the sum
    double v;
    memcpy(&v, &bits, sizeof v);
    gb2::count_equal_kernel<<<blocks, threads, 0, h->stream>>>(e, A.nvox, (size_t)A.dev.gstride, v, d_u + 1);
    unsigned long long cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d_u + 1, sizeof cnt, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    long long take = std::min<long long>((long long)cnt, top_voxels - taken);
    sum += v * (double)take;
    taken += take;
    bound = v;
  }
  long long denom = std::min<long long>(top_voxels, (long long)A.nvox);
  const double mean_max = denom > 0 ? sum / (double)denom : 0.0;
  CU(cudaMemsetAsync(d_u, 0, 2 * sizeof(unsigned long long), h->stream));
  CU(cudaMemsetAsync(d_s, 0, sizeof(double), h->stream));
  gb2::mean_unc_kernel<<<blocks, threads, 0, h->stream>>>(e, e2, A.nvox, (size_t)A.dev.gstride, mean_max * threshold, nev, d_s, d_u);
  double s = 0.0;
  unsigned long long c = 0;
  CU(cudaMemcpyAsync(&s, d_s, sizeof s, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(&c, d_u, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *out = (c > 0 && s > 0.0) ? s / (double)c : 1.0;
  return GATE_B200_OK;
}
