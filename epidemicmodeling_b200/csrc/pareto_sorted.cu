// pareto_sorted.cu -- Pareto front for LARGE point sets (BASELINE config 5: 100 000
// schedules per region), sm_100a.
//
// Reference semantics (Tools/TrainPredictPrescribeNPI.m:624-627):
//     on_front(i) = ( count_j [ J0_j < J0_i  &  J1_j < J1_i ] == 0 )      (strict in both)
// The O(n^2) count of pareto_kernel is exact but too slow beyond ~1e4 points.  The same
// predicate, evaluated in O(n log n):  sort by J0;  i is dominated  <=>  the minimum J1 over
// all points with STRICTLY smaller J0 is < J1_i.  Ties in J0 are handled by taking the prefix
// minimum up to the first element of the tie group (lower bound); NaN never compares true, so
// NaN points neither dominate (J1 = NaN -> +inf in the running minimum; J0 = NaN sorts last)
// nor are dominated -- exactly what the reference's `<` comparisons do.
// The segmented sort is library code (cub::DeviceSegmentedSort); the scan / lower-bound /
// scatter kernels are ours.
#include <cub/cub.cuh>

#include "epi_device.cuh"
#include "epi_internal.h"

namespace epi {

// sort keys: J0 with NaN mapped to +inf (a NaN key would break the sort's ordering; the
// front kernel re-checks the ORIGINAL value, so a NaN point is still never dominated, and
// as the largest key it never counts as "strictly smaller" than anything)
__global__ void iota_segments_kernel(const double *__restrict__ J0, double *__restrict__ keys, int *offsets,
                                     int *idx, int n_sets, int n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q <= (size_t)n_sets) offsets[q] = (int)(q * (size_t)n);
  if (q < (size_t)n_sets * n) {
    idx[q] = (int)(q % (size_t)n);
    const double v = J0[q];
    keys[q] = (v != v) ? __longlong_as_double(0x7ff0000000000000ll) : v;
  }
}

// one CTA per set: inclusive running minimum of J1 in J0-sorted order (NaN -> +inf)
constexpr int kScanBlock = 256;
__global__ void __launch_bounds__(kScanBlock) prefix_min_kernel(const double *__restrict__ J1,
                                                                const int *__restrict__ sorted_idx, int n,
                                                                double *__restrict__ run_min) {
  using Scan = cub::BlockScan<double, kScanBlock>;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ double carry_s;
  const int set = blockIdx.x;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double *j1 = J1 + (size_t)set * n;
  const int *si = sorted_idx + (size_t)set * n;
  double *rm = run_min + (size_t)set * n;
  double carry = inf;
  for (int base = 0; base < n; base += kScanBlock) {
    const int i = base + threadIdx.x;
    double v = inf;
    if (i < n) {
      v = j1[si[i]];
      if (v != v) v = inf;
    }
    double out;
    Scan(tmp).InclusiveScan(v, out, cub::Min());
    out = (carry < out) ? carry : out;
    if (i < n) rm[i] = out;
    if (threadIdx.x == kScanBlock - 1) carry_s = out;
    __syncthreads();
    carry = carry_s;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) front_from_sorted_kernel(const double *__restrict__ J0,
                                                                const double *__restrict__ J1,
                                                                const double *__restrict__ sorted_key,
                                                                const int *__restrict__ sorted_idx,
                                                                const double *__restrict__ run_min, int n_sets,
                                                                int n, unsigned char *__restrict__ on_front) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)n_sets * n) return;
  const int set = (int)(q / n), i = (int)(q % n);
  const double *ks = sorted_key + (size_t)set * n;
  const double key = ks[i];
  // lower bound: first position s with !(ks[s] < key)
  int lo = 0, hi = i;  // ks[i] itself is not < key, so the answer is <= i
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ks[mid] < key) lo = mid + 1; else hi = mid;
  }
  const int orig = sorted_idx[(size_t)set * n + i];
  const double x1 = J1[(size_t)set * n + orig];
  const double x0 = J0[(size_t)set * n + orig];
  const bool dominated = (x0 == x0) && (lo > 0) && (run_min[(size_t)set * n + lo - 1] < x1);
  on_front[(size_t)set * n + orig] = dominated ? 0 : 1;
}

// knee point for sets that do not fit the shared-memory path: same reductions as
// pareto_kernel (:633; MATLAB max/min skip NaN, first minimum wins), reading global memory
__global__ void __launch_bounds__(256) knee_kernel(const double *__restrict__ J0g, const double *__restrict__ J1g,
                                                   int n, int *__restrict__ I_opt) {
  const int set = blockIdx.x;
  const double *J0 = J0g + (size_t)set * n, *J1 = J1g + (size_t)set * n;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  __shared__ double r0[8], r1[8];
  __shared__ int ri[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double m0 = nan, m1 = nan;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { m0 = mmax(m0, J0[i]); m1 = mmax(m1, J1[i]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 = mmax(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = mmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
  }
  if (lane == 0) { r0[wid] = m0; r1[wid] = m1; }
  __syncthreads();
  m0 = r0[0]; m1 = r1[0];
  for (int w = 1; w < 8; ++w) { m0 = mmax(m0, r0[w]); m1 = mmax(m1, r1[w]); }
  __syncthreads();
  double bv = nan;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double q0 = J0[i] / m0, q1 = J1[i] / m1;
    const double v = q0 * q0 + q1 * q1;
    if (v == v && (bv != bv || v < bv)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if ((ov == ov) && (bv != bv || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
  }
  if (lane == 0) { r0[wid] = bv; ri[wid] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      const double ov = r0[w];
      const int oi = ri[w];
      if ((ov == ov) && (bv != bv || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    I_opt[set] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

size_t pareto_sorted_scratch_bytes(int n_sets, int n) {
  const size_t tot = (size_t)n_sets * n;
  size_t cub_bytes = 0;
  cub::DeviceSegmentedSort::SortPairs(nullptr, cub_bytes, (const double *)nullptr, (double *)nullptr,
                                      (const int *)nullptr, (int *)nullptr, (int)tot, n_sets, (const int *)nullptr,
                                      (const int *)nullptr);
  // keys in/out + running min (double), idx in/out + offsets (int), cub temp
  return tot * 24 + tot * 8 + ((size_t)n_sets + 1) * 4 + cub_bytes + 2048;
}

int launch_pareto_sorted(const ParetoParams &p, void *scratch, size_t scratch_bytes, cudaStream_t st) {
  const size_t tot = (size_t)p.n_sets * p.n;
  int launches = 0;
  if (p.on_front) {
    char *w = (char *)scratch;
    double *keys_in = (double *)w; w += tot * 8;
    double *keys_out = (double *)w; w += tot * 8;
    double *run_min = (double *)w; w += tot * 8;
    int *idx_in = (int *)w; w += tot * 4;
    int *idx_out = (int *)w; w += tot * 4;
    int *offsets = (int *)w; w += ((size_t)p.n_sets + 1) * 4;
    w = (char *)(((size_t)w + 255) & ~(size_t)255);  // cub temp storage: 256-byte aligned
    size_t cub_bytes = scratch_bytes - (size_t)(w - (char *)scratch);
    iota_segments_kernel<<<(unsigned)((tot + 1 + 255) / 256), 256, 0, st>>>(p.J0, keys_in, offsets, idx_in,
                                                                            p.n_sets, p.n);
    cub::DeviceSegmentedSort::SortPairs(w, cub_bytes, keys_in, keys_out, idx_in, idx_out, (int)tot, p.n_sets, offsets,
                                        offsets + 1, st);
    prefix_min_kernel<<<p.n_sets, kScanBlock, 0, st>>>(p.J1, idx_out, p.n, run_min);
    front_from_sorted_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p.J0, p.J1, keys_out, idx_out, run_min,
                                                                            p.n_sets, p.n, p.on_front);
    launches += 4;
  }
  if (p.I_opt) {
    knee_kernel<<<p.n_sets, 256, 0, st>>>(p.J0, p.J1, p.n, p.I_opt);
    launches += 1;
  }
  return launches;
}

}  // namespace epi
