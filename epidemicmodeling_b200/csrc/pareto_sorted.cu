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
//
// Pruning before the sort (round 2): of 100 000 random schedules ~3 % are on the front, and sorting whole 100k-point
// segments took 1.6 of the step's 2.6 ms (30 sets).  The J0 range of a set is cut into kBuckets buckets (a monotone
// map, see bucket_of: a point in a LOWER bucket has a STRICTLY smaller J0), every bucket keeps its minimum J1, and a point
// whose J1 exceeds the minimum over all lower buckets is dominated for certain.  Only the survivors are sorted and
// tested exactly; a survivor can only be dominated by another survivor (if y dominates x and z, from a lower bucket
// than y, dominates y, then z prunes x as well), so the mask is the reference's, bit for bit.
#include <cstdlib>
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

// ---- pruning ------------------------------------------------------------------------------------
constexpr int kBuckets = 1024;
constexpr int kPruneBlock = 256;
constexpr int kPruneChunk = 8192;  // points of one set per CTA

// order-preserving map double -> uint64 (for atomicMin / atomicMax); NaN is never passed in
__device__ __forceinline__ unsigned long long ord64(double v) {
  const long long b = __double_as_longlong(v + 0.0);  // -0 and +0 compare equal: one pattern for both
  return (b < 0) ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord64_inv(unsigned long long u) {
  return __longlong_as_double((long long)((u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u));
}
__device__ __forceinline__ bool finite64(double v) { return fabs(v) <= 1.79769313486231570815e308; }

struct PruneBufs {
  unsigned long long *range;  // [n_sets][2]: ordered min / max of the finite J0
  unsigned long long *bmin;   // [n_sets][kBuckets]: ordered minimum J1 of the bucket (all ones = empty)
  double *pmin;               // [n_sets][kBuckets]: minimum J1 over all LOWER buckets (+inf = none)
  int *count;                 // [n_sets]: survivors
  int *seg_begin, *seg_end;   // [n_sets]: the survivors of set s live at s*n + [0, count)
};

// Bucket of a point: monotone non-decreasing in x0.  The buckets are equal steps of the ORDER-PRESERVING BIT PATTERN of
// J0 between the set's finite minimum and maximum, i.e. logarithmic in the value: the infection cost of random schedules
// spans five decades, and equal-width buckets left 15 % of the points (83 % in the worst region) to the sort where
// these leave 2.9 % -- the front itself is 2.8 % (measured on config 5's costs, 30 regions x 100 000 schedules).
struct BucketMap {
  unsigned long long lo;  // ordered bit pattern of the smallest finite J0
  int shift;              // (ord - lo) >> shift < 2^20
  double inv;             // kBuckets / (((hi - lo) >> shift) + 1)
};
__device__ __forceinline__ int bucket_of(double x0, const BucketMap &m) {
  const unsigned long long o = ord64(x0);
  if (o <= m.lo) return 0;
  const double t = (double)((o - m.lo) >> m.shift) * m.inv;
  return (t >= (double)(kBuckets - 1)) ? kBuckets - 1 : (int)t;
}

__global__ void prune_init_kernel(PruneBufs B, int n_sets, int n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < (size_t)n_sets * kBuckets) B.bmin[q] = ~0ull;
  if (q < (size_t)n_sets) {
    B.range[2 * q] = ~0ull; B.range[2 * q + 1] = 0ull;
    B.count[q] = 0;
    B.seg_begin[q] = (int)(q * (size_t)n);
  }
}

__global__ void __launch_bounds__(kPruneBlock) prune_range_kernel(const double *__restrict__ J0, int n, PruneBufs B) {
  const int set = blockIdx.y, base = blockIdx.x * kPruneChunk;
  const double *j0 = J0 + (size_t)set * n;
  unsigned long long lo = ~0ull, hi = 0ull;
  for (int i = base + threadIdx.x; i < n && i < base + kPruneChunk; i += kPruneBlock) {
    const double v = j0[i];
    if (finite64(v)) { const unsigned long long o = ord64(v); lo = o < lo ? o : lo; hi = o > hi ? o : hi; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    if (lo != ~0ull) atomicMin(B.range + 2 * set, lo);
    if (hi != 0ull) atomicMax(B.range + 2 * set + 1, hi);
  }
}

// the bucket map of a set (no finite J0: everything lands in bucket 0 or the last one, nothing is pruned wrongly)
__device__ __forceinline__ BucketMap bucket_map(const PruneBufs &B, int set) {
  const unsigned long long l = B.range[2 * set], h = B.range[2 * set + 1];
  BucketMap m;
  m.lo = l; m.shift = 0; m.inv = (double)kBuckets;
  if (l <= h && l != ~0ull) {
    const unsigned long long range = h - l;
    const int bits = 64 - __clzll((long long)range);  // 0 for range == 0
    m.shift = bits > 20 ? bits - 20 : 0;
    m.inv = (double)kBuckets / ((double)(range >> m.shift) + 1.0);
  }
  return m;
}

__global__ void __launch_bounds__(kPruneBlock) prune_bucket_kernel(const double *__restrict__ J0,
                                                                    const double *__restrict__ J1, int n, PruneBufs B) {
  __shared__ unsigned long long sm[kBuckets];
  const int set = blockIdx.y, base = blockIdx.x * kPruneChunk;
  for (int b = threadIdx.x; b < kBuckets; b += kPruneBlock) sm[b] = ~0ull;
  __syncthreads();
  const BucketMap bm = bucket_map(B, set);
  const double *j0 = J0 + (size_t)set * n, *j1 = J1 + (size_t)set * n;
  for (int i = base + threadIdx.x; i < n && i < base + kPruneChunk; i += kPruneBlock) {
    const double x0 = j0[i], x1 = j1[i];
    if (x0 == x0 && x1 == x1) atomicMin(&sm[bucket_of(x0, bm)], ord64(x1));  // a NaN never dominates
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kBuckets; b += kPruneBlock)
    if (sm[b] != ~0ull) atomicMin(B.bmin + (size_t)set * kBuckets + b, sm[b]);
}

// one CTA of kBuckets threads per set: pmin[b] = min over buckets < b
__global__ void __launch_bounds__(kBuckets) prune_scan_kernel(PruneBufs B) {
  using Scan = cub::BlockScan<double, kBuckets>;
  __shared__ typename Scan::TempStorage tmp;
  const int set = blockIdx.x, b = threadIdx.x;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const unsigned long long o = B.bmin[(size_t)set * kBuckets + b];
  const double v = (o == ~0ull) ? inf : ord64_inv(o);
  double out;
  Scan(tmp).ExclusiveScan(v, out, inf, cub::Min());
  B.pmin[(size_t)set * kBuckets + b] = out;
}

// survivors -> keys / idx at the head of the set's own segment (order within the segment is irrelevant: it is sorted next)
__global__ void __launch_bounds__(kPruneBlock) prune_compact_kernel(const double *__restrict__ J0,
                                                                     const double *__restrict__ J1, int n, PruneBufs B,
                                                                     double *__restrict__ keys, int *__restrict__ idx) {
  const int set = blockIdx.y, base = blockIdx.x * kPruneChunk;
  const BucketMap bm = bucket_map(B, set);
  const double *j0 = J0 + (size_t)set * n, *j1 = J1 + (size_t)set * n;
  const double *pm = B.pmin + (size_t)set * kBuckets;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  for (int i0 = base; i0 < n && i0 < base + kPruneChunk; i0 += kPruneBlock) {
    const int i = i0 + threadIdx.x;
    bool keep = false;
    double x0 = 0.0;
    if (i < n && i < base + kPruneChunk) {
      x0 = j0[i];
      const double x1 = j1[i];
      keep = !(x0 == x0) || !(pm[bucket_of(x0, bm)] < x1);   // NaN J0 is never dominated; NaN J1 compares false
    }
    // warp-aggregated append
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      int pos = 0;
      if (lane == leader) pos = atomicAdd(B.count + set, __popc(m));
      pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(m & ((1u << lane) - 1));
      if (keep) {
        keys[(size_t)set * n + pos] = (x0 != x0) ? inf : x0;
        idx[(size_t)set * n + pos] = i;
      }
    }
  }
}

__global__ void prune_ends_kernel(PruneBufs B, int n_sets) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n_sets) B.seg_end[q] = B.seg_begin[q] + B.count[q];
}

// one CTA per set: inclusive running minimum of J1 in J0-sorted order (NaN -> +inf)
constexpr int kScanBlock = 256;
__global__ void __launch_bounds__(kScanBlock) prefix_min_kernel(const double *__restrict__ J1,
                                                                const int *__restrict__ sorted_idx, int n,
                                                                double *__restrict__ run_min,
                                                                const int *__restrict__ count) {
  using Scan = cub::BlockScan<double, kScanBlock>;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ double carry_s;
  const int set = blockIdx.x;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double *j1 = J1 + (size_t)set * n;
  const int *si = sorted_idx + (size_t)set * n;
  double *rm = run_min + (size_t)set * n;
  double carry = inf;
  const int m = count ? count[set] : n;  // points of this set that were sorted (all of them without pruning)
  for (int base = 0; base < m; base += kScanBlock) {
    const int i = base + threadIdx.x;
    double v = inf;
    if (i < m) {
      v = j1[si[i]];
      if (v != v) v = inf;
    }
    double out;
    Scan(tmp).InclusiveScan(v, out, cub::Min());
    out = (carry < out) ? carry : out;
    if (i < m) rm[i] = out;
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
                                                                int n, unsigned char *__restrict__ on_front,
                                                                const int *__restrict__ count) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)n_sets * n) return;
  const int set = (int)(q / n), i = (int)(q % n);
  if (count && i >= count[set]) return;  // pruned points keep the 0 the mask was cleared to
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
constexpr int kKneeBlock = 1024;  // one CTA per set walks its n points twice: 0.39 -> 0.12 ms for 30 x 100k against 256 threads
constexpr int kKneeWarps = kKneeBlock / 32;
__global__ void __launch_bounds__(kKneeBlock) knee_kernel(const double *__restrict__ J0g, const double *__restrict__ J1g,
                                                          int n, int *__restrict__ I_opt) {
  const int set = blockIdx.x;
  const double *J0 = J0g + (size_t)set * n, *J1 = J1g + (size_t)set * n;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  __shared__ double r0[kKneeWarps], r1[kKneeWarps];
  __shared__ int ri[kKneeWarps];
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
  for (int w = 1; w < kKneeWarps; ++w) { m0 = mmax(m0, r0[w]); m1 = mmax(m1, r1[w]); }
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
    for (int w = 1; w < kKneeWarps; ++w) {
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
  // keys in/out + running min (double), idx in/out + offsets (int), pruning tables, cub temp
  return tot * 24 + tot * 8 + ((size_t)n_sets + 1) * 4 + (size_t)n_sets * (16 + 16 * kBuckets + 12 + 64) + cub_bytes + 4096;
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
    w = (char *)(((size_t)w + 15) & ~(size_t)15);
    PruneBufs B;
    B.range = (unsigned long long *)w; w += (size_t)p.n_sets * 16;
    B.bmin = (unsigned long long *)w; w += (size_t)p.n_sets * kBuckets * 8;
    B.pmin = (double *)w; w += (size_t)p.n_sets * kBuckets * 8;
    B.count = (int *)w; w += (size_t)p.n_sets * 4;
    B.seg_begin = (int *)w; w += (size_t)p.n_sets * 4;
    B.seg_end = (int *)w; w += (size_t)p.n_sets * 4;
    w = (char *)(((size_t)w + 255) & ~(size_t)255);  // cub temp storage: 256-byte aligned
    size_t cub_bytes = scratch_bytes - (size_t)(w - (char *)scratch);
    int prune = 1;
    if (const char *e = getenv("EPI_PARETO_PRUNE")) prune = atoi(e);  // 0: sort whole sets (the parity tests compare both)
    if (prune) {
      const dim3 grid((unsigned)((p.n + kPruneChunk - 1) / kPruneChunk), (unsigned)p.n_sets);
      const size_t ninit = (size_t)p.n_sets * kBuckets;
      cudaMemsetAsync(p.on_front, 0, tot, st);
      prune_init_kernel<<<(unsigned)((ninit + 255) / 256), 256, 0, st>>>(B, p.n_sets, p.n);
      prune_range_kernel<<<grid, kPruneBlock, 0, st>>>(p.J0, p.n, B);
      prune_bucket_kernel<<<grid, kPruneBlock, 0, st>>>(p.J0, p.J1, p.n, B);
      prune_scan_kernel<<<p.n_sets, kBuckets, 0, st>>>(B);
      prune_compact_kernel<<<grid, kPruneBlock, 0, st>>>(p.J0, p.J1, p.n, B, keys_in, idx_in);
      prune_ends_kernel<<<(unsigned)((p.n_sets + 255) / 256), 256, 0, st>>>(B, p.n_sets);
      cub::DeviceSegmentedSort::SortPairs(w, cub_bytes, keys_in, keys_out, idx_in, idx_out, (int)tot, p.n_sets, B.seg_begin,
                                          B.seg_end, st);
      prefix_min_kernel<<<p.n_sets, kScanBlock, 0, st>>>(p.J1, idx_out, p.n, run_min, B.count);
      front_from_sorted_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p.J0, p.J1, keys_out, idx_out, run_min,
                                                                              p.n_sets, p.n, p.on_front, B.count);
      launches += 9;
    } else {
      iota_segments_kernel<<<(unsigned)((tot + 1 + 255) / 256), 256, 0, st>>>(p.J0, keys_in, offsets, idx_in,
                                                                              p.n_sets, p.n);
      cub::DeviceSegmentedSort::SortPairs(w, cub_bytes, keys_in, keys_out, idx_in, idx_out, (int)tot, p.n_sets, offsets,
                                          offsets + 1, st);
      prefix_min_kernel<<<p.n_sets, kScanBlock, 0, st>>>(p.J1, idx_out, p.n, run_min, nullptr);
      front_from_sorted_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p.J0, p.J1, keys_out, idx_out, run_min,
                                                                              p.n_sets, p.n, p.on_front, nullptr);
      launches += 4;
    }
  }
  if (p.I_opt) {
    knee_kernel<<<p.n_sets, kKneeBlock, 0, st>>>(p.J0, p.J1, p.n, p.I_opt);
    launches += 1;
  }
  return launches;
}

}  // namespace epi
