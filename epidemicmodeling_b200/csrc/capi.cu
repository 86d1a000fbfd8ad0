// capi.cu -- the C ABI of libepi_b200.so (include/epi_b200.h): argument
// validation, host<->device staging, wave scheduling over the scratch budget,
// kernel sequencing and per-phase device timing.  No exception crosses the
// boundary; there is no CPU fallback (epi_create fails without a CUDA device).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <cstdlib>
#include <chrono>
#include <vector>
#include <thread>
#include <algorithm>

#include "epi_device.cuh"
#include "epi_internal.h"

using namespace epi;

namespace {

struct EpiError {
  int code;
  std::string msg;
};

thread_local std::string g_create_err;

struct Phase {
  const char *name;
  cudaEvent_t a, b;
};

}  // namespace

struct epi_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  std::string err;
  size_t scratch_limit = 0;
  long long launches = 0;
  std::vector<Phase> phases;       // of the most recent batched call
  std::vector<cudaEvent_t> ev_pool;
  // context-owned device memory: staging/scratch blocks are cached here between calls, so a
  // repeated call (a sweep per step, a MEX call per region) never pays cudaMalloc again.  All
  // work of a context is enqueued on ONE stream, so recycling a block is stream-ordered.
  std::multimap<size_t, void *> free_blocks;
  size_t cached_bytes = 0;
  // (B, bytes per trajectory) shapes that ran as a single wave: while their blocks are still cached the
  // budget query (cudaMemGetInfo: 0.1-7 ms of driver time, measured) is skipped on a repeated call
  std::set<std::pair<long long, size_t>> fits;
  // side stream for host-memory inputs that the first kernels of a call do not read (the sweep's cost
  // weights: 12.7 MB that upload beside the forward pass instead of in front of it)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_ev[2] = {nullptr, nullptr};
  // EPI_MEM_DEVICE calls cannot read their epi_model_params on the host without draining the stream: a
  // one-CTA kernel checks them on the device and records the first violation in this mapped host word
  // (EPI_ERR_* code); it is reported by the next epi_sync or the next call on the context.
  int *deferred_err = nullptr;      // pinned, device-mapped
  int *deferred_err_dev = nullptr;
  // piped schedule of small sweeps (forward || gains): second stream, fork/join events, the driver's stream wait on a
  // device word (cuStreamWaitValue32, fetched through the runtime so the library needs no link against libcuda)
  cudaStream_t pipe_stream = nullptr;
  cudaEvent_t pipe_ev[2] = {nullptr, nullptr};
  void *wait_value32 = nullptr;
  int pipe_state = 0;               // 0 untried, 1 ready, -1 unavailable
  int n_sms = 0;
};

namespace {

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      char buf_[512];                                                                  \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
               __FILE__, __LINE__);                                                    \
      throw EpiError{e_ == cudaErrorMemoryAllocation ? EPI_ERR_NOMEM : EPI_ERR_CUDA, buf_}; \
    }                                                                                  \
  } while (0)

[[noreturn]] void bad_arg(const std::string &m, int code = EPI_ERR_ARG) { throw EpiError{code, m}; }
constexpr int kMaxDynSmem = 232448;  // 227 KB: the opt-in dynamic shared memory limit of one CTA on sm_100

cudaEvent_t get_event(epi_ctx *c) {
  if (!c->ev_pool.empty()) {
    cudaEvent_t e = c->ev_pool.back();
    c->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  CK(cudaEventCreate(&e));
  return e;
}
void reset_phases(epi_ctx *c) {
  for (auto &p : c->phases) { c->ev_pool.push_back(p.a); c->ev_pool.push_back(p.b); }
  c->phases.clear();
}
struct PhaseScope {
  epi_ctx *c;
  Phase ph;
  PhaseScope(epi_ctx *ctx, const char *name) : c(ctx) {
    ph.name = name; ph.a = get_event(c); ph.b = get_event(c);
    CK(cudaEventRecord(ph.a, c->stream));
  }
  void end() {
    CK(cudaEventRecord(ph.b, c->stream));
    c->phases.push_back(ph);
  }
};

// cuStreamWaitValue32(stream, addr, value, flags): flags 1 = CU_STREAM_WAIT_VALUE_GEQ
typedef int (*WaitValue32Fn)(cudaStream_t, unsigned long long, unsigned, unsigned);
bool pipe_ready(epi_ctx *c) {
  if (c->pipe_state == 0) {
    c->pipe_state = -1;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess && fn &&
        cudaStreamCreateWithFlags(&c->pipe_stream, cudaStreamNonBlocking) == cudaSuccess) {
      cudaEventCreateWithFlags(&c->pipe_ev[0], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&c->pipe_ev[1], cudaEventDisableTiming);
      cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, c->device);
      c->wait_value32 = fn;
      c->pipe_state = 1;
    } else {
      cudaGetLastError();
    }
  }
  return c->pipe_state == 1;
}

void trim_cache(epi_ctx *c) {
  if (c->free_blocks.empty()) return;
  cudaStreamSynchronize(c->stream);
  for (auto &kv : c->free_blocks) cudaFree(kv.second);
  c->free_blocks.clear();
  c->cached_bytes = 0;
  c->fits.clear();
}
void *ctx_alloc(epi_ctx *c, size_t *bytes_io) {
  size_t bytes = (*bytes_io + 511) & ~(size_t)511;
  if (bytes == 0) bytes = 512;
  *bytes_io = bytes;
  auto it = c->free_blocks.lower_bound(bytes);
  if (it != c->free_blocks.end() && it->first <= bytes + bytes / 4 + 65536) {
    void *p = it->second;
    *bytes_io = it->first;
    c->cached_bytes -= it->first;
    c->free_blocks.erase(it);
    return p;
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give the cache back and retry once
    cudaGetLastError();
    trim_cache(c);
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    char buf[256];
    snprintf(buf, sizeof buf, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    throw EpiError{e == cudaErrorMemoryAllocation ? EPI_ERR_NOMEM : EPI_ERR_CUDA, buf};
  }
  return p;
}
void ctx_free(epi_ctx *c, void *p, size_t bytes) {
  c->free_blocks.emplace(bytes, p);
  c->cached_bytes += bytes;
}

// one batched call (or one wave of it): owns the staging buffers it allocated
// and the list of device->host copies to issue once the kernels are enqueued.
struct Call {
  epi_ctx *c;
  int mem;
  std::vector<std::pair<void *, size_t>> dev;
  struct Copy { void *dst; const void *src; size_t dpitch, spitch, width, height; };
  std::vector<Copy> pend;

  Call(epi_ctx *ctx, int m) : c(ctx), mem(m) {}
  ~Call() { release(); }
  Call(const Call &) = delete;

  void *dalloc(size_t bytes) {
    void *p = ctx_alloc(c, &bytes);
    dev.emplace_back(p, bytes);
    return p;
  }
  void release() {
    for (auto &pb : dev) ctx_free(c, pb.first, pb.second);
    dev.clear();
  }
  // whole-array input of n elements (per-group tables etc.)
  template <class T>
  const T *in(const T *p, size_t n) {
    if (!p || mem == EPI_MEM_DEVICE) return p;
    T *d = (T *)dalloc(n * sizeof(T));
    CK(cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return d;
  }
  // whole-array output of n elements
  template <class T>
  T *out(T *p, size_t n) {
    if (!p || mem == EPI_MEM_DEVICE) return p;
    T *d = (T *)dalloc(n * sizeof(T));
    pend.push_back({p, d, n * sizeof(T), n * sizeof(T), n * sizeof(T), 1});
    return d;
  }
  // per-trajectory input [rows][B]: the slice [b0, b0+Bw) of every row
  template <class T>
  const T *traj_in_raw(const T *p, size_t rows, long long B, long long b0, long long Bw,
                       long long *stride, long long *off) {
    if (!p) { *stride = 0; *off = 0; return nullptr; }
    if (mem == EPI_MEM_DEVICE) { *stride = B; *off = b0; return p; }
    T *d = (T *)dalloc(rows * (size_t)Bw * sizeof(T));
    CK(cudaMemcpy2DAsync(d, (size_t)Bw * sizeof(T), p + b0, (size_t)B * sizeof(T),
                         (size_t)Bw * sizeof(T), rows, cudaMemcpyHostToDevice, c->stream));
    *stride = Bw; *off = 0;
    return d;
  }
  CArr traj_in(const double *p, size_t rows, long long B, long long b0, long long Bw) {
    CArr a;
    a.p = traj_in_raw<double>(p, rows, B, b0, Bw, &a.stride, &a.off);
    return a;
  }
  // per-trajectory output [rows][B]
  TArr traj_out(double *p, size_t rows, long long B, long long b0, long long Bw) {
    if (!p) return TArr{nullptr, 0, 0};
    if (mem == EPI_MEM_DEVICE) return TArr{p, B, b0};
    double *d = (double *)dalloc(rows * (size_t)Bw * sizeof(double));
    pend.push_back({p + b0, d, (size_t)B * sizeof(double), (size_t)Bw * sizeof(double),
                    (size_t)Bw * sizeof(double), rows});
    return TArr{d, Bw, 0};
  }
  // library scratch [rows][Bw] in the caller-style strided layout
  TArr scratch(size_t rows, long long Bw) {
    return TArr{(double *)dalloc(rows * (size_t)Bw * sizeof(double)), Bw, 0};
  }
  // library scratch in the tile layout [ceil(Bw/32)][rows][32] (ekf_common.cuh)
  TArr scratch_tiled(size_t rows, long long Bw) {
    const size_t tiles = (size_t)((Bw + 31) / 32);
    return TArr{(double *)dalloc(tiles * rows * 32 * sizeof(double)), 32, 0};
  }
  void flush() {
    for (auto &k : pend)
      CK(cudaMemcpy2DAsync(k.dst, k.dpitch, k.src, k.spitch, k.width, k.height,
                           cudaMemcpyDeviceToHost, c->stream));
    pend.clear();
  }
};

void check_launch(epi_ctx *c, int n) {
  c->launches += n;
  CK(cudaGetLastError());
}

size_t scratch_budget(epi_ctx *c) {
  if (c->scratch_limit) return c->scratch_limit;
  size_t fr = 0, tot = 0;
  CK(cudaMemGetInfo(&fr, &tot));
  return (size_t)(0.6 * (double)(fr + c->cached_bytes));  // blocks cached by this context are reusable
}

long long wave_size(long long B, size_t bytes_per_traj, size_t budget) {
  if (bytes_per_traj == 0) return B;
  long long w = (long long)(budget / bytes_per_traj);
  if (w >= B) return B;
  if (w >= 4096) w -= w % 1024;
  else if (w >= 64) w -= w % 32;
  if (w < 1) bad_arg("scratch limit too small for a single trajectory", EPI_ERR_NOMEM);
  return w;
}

// wave size of a batch of B trajectories needing `per` bytes of device memory each
long long plan_wave(epi_ctx *c, long long B, size_t per) {
  const auto key = std::make_pair(B, per);
  if (!c->scratch_limit && c->fits.count(key) && c->cached_bytes >= (size_t)B * per) return B;
  const long long w = wave_size(B, per, scratch_budget(c));
  if (w >= B && !c->scratch_limit) c->fits.insert(key);
  return w;
}

void take_deferred_error(epi_ctx *c) {
  if (!c->deferred_err) return;
  const int code = *(volatile int *)c->deferred_err;
  if (!code) return;
  *c->deferred_err = 0;
  throw EpiError{code, code == EPI_ERR_OBS_TYPE
                           ? "unknown observation type (found on the device in an earlier EPI_MEM_DEVICE call)"  // SIAlphaModelEKF.m:57
                           : "epi_model_params.L does not match args.L (found on the device in an earlier EPI_MEM_DEVICE call)"};
}

template <class F>
int guarded(epi_ctx *ctx, F &&f) {
  if (!ctx) return EPI_ERR_ARG;
  // the calling thread's current device is the host framework's business (torch tracks it per thread):
  // switch to the context's GPU for the call and put the caller's back afterwards
  struct DeviceScope {
    int prev = -1, mine;
    explicit DeviceScope(int dev) : mine(dev) {
      if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
      if (prev != mine) CK(cudaSetDevice(mine));
    }
    ~DeviceScope() {
      if (prev >= 0 && prev != mine) cudaSetDevice(prev);
    }
  };
  try {
    DeviceScope scope(ctx->device);
    take_deferred_error(ctx);   // a violation a previous EPI_MEM_DEVICE call found on the device
    f();
    ctx->err.clear();
    return EPI_OK;
  } catch (const EpiError &e) {
    ctx->err = e.msg;
    cudaGetLastError();
    return e.code;
  } catch (const std::exception &e) {
    ctx->err = e.what();
    return EPI_ERR_ARG;
  }
}

void check_mem(int mem) {
  if (mem != EPI_MEM_HOST && mem != EPI_MEM_DEVICE) bad_arg("mem must be EPI_MEM_HOST or EPI_MEM_DEVICE");
}

void finish(epi_ctx *c, int mem) {
  if (mem == EPI_MEM_HOST) CK(cudaStreamSynchronize(c->stream));
}

// EPI_TRACE_HOST=1: host-side timeline of a call on stderr (where a blocking host-memory call spends
// its wall time: argument checks, copy/launch enqueue, waiting for the stream)
struct HostTrace {
  bool on;
  std::chrono::steady_clock::time_point t0, last;
  std::string line;
  explicit HostTrace(const char *what) : on(getenv("EPI_TRACE_HOST") != nullptr) {
    if (on) { t0 = last = std::chrono::steady_clock::now(); line = what; }
  }
  void mark(const char *name) {
    if (!on) return;
    auto t = std::chrono::steady_clock::now();
    char b[64];
    snprintf(b, sizeof b, " %s %.3f", name, std::chrono::duration<double, std::milli>(t - last).count());
    line += b;
    last = t;
  }
  ~HostTrace() {
    if (!on) return;
    fprintf(stderr, "[epi host trace] %s | total %.3f ms\n", line.c_str(),
            std::chrono::duration<double, std::milli>(last - t0).count());
  }
};

}  // namespace

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" int epi_create(int device, epi_ctx **out) {
  if (!out) return EPI_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_create_err = std::string("no CUDA device available (") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                   "): libepi_b200 has no CPU fallback";
    cudaGetLastError();
    return EPI_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    g_create_err = "device index out of range";
    return EPI_ERR_ARG;
  }
  epi_ctx *c = new epi_ctx;
  c->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_err = std::string("CUDA init failed: ") + cudaGetErrorString(e);
    delete c;
    return EPI_ERR_CUDA;
  }
  c->stream = c->own_stream;
  *out = c;
  g_create_err.clear();
  return EPI_OK;
}

extern "C" void epi_destroy(epi_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  reset_phases(c);
  trim_cache(c);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
    cudaEventDestroy(c->copy_ev[0]);
    cudaEventDestroy(c->copy_ev[1]);
  }
  if (c->pipe_stream) {
    cudaStreamSynchronize(c->pipe_stream);
    cudaStreamDestroy(c->pipe_stream);
    cudaEventDestroy(c->pipe_ev[0]);
    cudaEventDestroy(c->pipe_ev[1]);
  }
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->deferred_err) cudaFreeHost(c->deferred_err);
  delete c;
}

extern "C" const char *epi_last_error(const epi_ctx *c) { return c ? c->err.c_str() : g_create_err.c_str(); }

extern "C" int epi_set_stream(epi_ctx *c, void *s) {
  if (!c) return EPI_ERR_ARG;
  cudaStream_t ns = s ? (cudaStream_t)s : c->own_stream;
  if (ns != c->stream) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);  // cached blocks are recycled in stream order: drain the old stream
    c->stream = ns;
  }
  return EPI_OK;
}

extern "C" int epi_sync(epi_ctx *c) {
  return guarded(c, [&] {
    CK(cudaStreamSynchronize(c->stream));
    take_deferred_error(c);
  });
}

extern "C" int epi_set_scratch_limit(epi_ctx *c, size_t bytes) {
  if (!c) return EPI_ERR_ARG;
  c->scratch_limit = bytes;
  return EPI_OK;
}

extern "C" int epi_release_cache(epi_ctx *c) {
  return guarded(c, [&] { trim_cache(c); });
}

extern "C" long long epi_launch_count(const epi_ctx *c) { return c ? c->launches : 0; }

extern "C" int epi_last_kernel_times(epi_ctx *c, float *ms, const char **names, int max) {
  if (!c) return 0;
  int n = 0;
  for (auto &p : c->phases) {
    if (cudaEventSynchronize(p.b) != cudaSuccess) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p.a, p.b) != cudaSuccess) continue;
    int k = 0;
    for (; k < n; ++k)
      if (!strcmp(names[k], p.name)) break;
    if (k == n) {
      if (n >= max) continue;
      names[n] = p.name; ms[n] = 0.f; ++n;
    }
    ms[k] += t;
  }
  return n;
}

extern "C" int epi_fp64_probe(epi_ctx *c, int iters, double *tflops) {
  return guarded(c, [&] {
    if (iters < 1 || !tflops) bad_arg("epi_fp64_probe: bad arguments");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    Call k(c, EPI_MEM_DEVICE);
    double *out = (double *)k.dalloc((size_t)blocks * threads * sizeof(double));
    launch_fp64_probe(out, blocks, threads, 16, c->stream);  // warm-up
    check_launch(c, 1);
    cudaEvent_t a = get_event(c), b = get_event(c);
    CK(cudaEventRecord(a, c->stream));
    launch_fp64_probe(out, blocks, threads, iters, c->stream);
    check_launch(c, 1);
    CK(cudaEventRecord(b, c->stream));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    c->ev_pool.push_back(a); c->ev_pool.push_back(b);
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / ((double)ms * 1e-3) / 1e12;
  });
}

// ---------------------------------------------------------------------------
// SEIRP
// ---------------------------------------------------------------------------
extern "C" int epi_seirp_batch(epi_ctx *c, const epi_seirp_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->B < 0 || a->K < 0) bad_arg("epi_seirp_batch: B and K must be >= 0");
    if (a->rate_mode < EPI_RATES_CONST || a->rate_mode > EPI_RATES_SERIES) bad_arg("epi_seirp_batch: bad rate_mode");
    if (a->out_mode != EPI_SEIRP_OUT_FULL && a->out_mode != EPI_SEIRP_OUT_FINAL) bad_arg("epi_seirp_batch: bad out_mode");
    if (a->B == 0 || a->K == 0) return;
    if (!a->rates || !a->ic || !a->out) bad_arg("epi_seirp_batch: rates, ic and out are required");
    reset_phases(c);
    const long long B = a->B;
    const int K = a->K;
    const size_t rate_rows = a->rate_mode == EPI_RATES_SERIES ? (size_t)7 * K : (a->rate_mode == EPI_RATES_CONST ? 7 : 0);
    const size_t out_rows = a->out_mode == EPI_SEIRP_OUT_FULL ? (size_t)5 * K : 5;
    Call shared(c, a->mem);
    const double *rates_shared =
        a->rate_mode == EPI_RATES_SHARED_SERIES ? shared.in(a->rates, (size_t)7 * K) : nullptr;
    const long long Bw = a->mem == EPI_MEM_HOST
                             ? plan_wave(c, B, (rate_rows + 5 + out_rows) * sizeof(double))
                             : B;
    for (long long b0 = 0; b0 < B; b0 += Bw) {
      const long long nb = (B - b0 < Bw) ? (B - b0) : Bw;
      Call w(c, a->mem);
      SeirpParams p{};
      p.B = (int)nb; p.K = K; p.rate_mode = a->rate_mode; p.saturated = a->saturated; p.out_mode = a->out_mode;
      p.dt = a->dt; p.beta_0 = a->beta_0; p.beta_s = a->beta_s; p.mu_0 = a->mu_0; p.mu_s = a->mu_s;
      p.sigma = a->sigma; p.i_0 = a->i_0;
      p.rates = rate_rows ? w.traj_in(a->rates, rate_rows, B, b0, nb) : CArr{nullptr, 0, 0};
      p.rates_shared = rates_shared;
      p.ic = w.traj_in(a->ic, 5, B, b0, nb);
      p.out = w.traj_out(a->out, out_rows, B, b0, nb);
      PhaseScope ph(c, "seirp");
      launch_seirp(p, c->stream);
      check_launch(c, 1);
      ph.end();
      w.flush();
    }
    finish(c, a->mem);
  });
}

// ---------------------------------------------------------------------------
// SIalpha_Controlled + NPICost
// ---------------------------------------------------------------------------
// EPI_U_PHILOX draws integer levels lo..hi per NPI: both bounds must be integers in 0..255
static void validate_levels_host(const epi_model_params *prm, long long n_groups, int L, const char *who) {
  for (long long g = 0; g < n_groups; ++g)
    for (int j = 0; j < L; ++j) {
      const double lo = prm[g].u_min[j], hi = prm[g].u_max[j];
      if (!(lo >= 0.0 && hi >= lo && hi <= 255.0) || lo != std::floor(lo) || hi != std::floor(hi))
        bad_arg(std::string(who) + ": prm.u_min / prm.u_max must be integers with 0 <= u_min <= u_max <= 255");
    }
}

extern "C" int epi_rollout_cost_batch(epi_ctx *c, const epi_rollout_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->B < 0 || a->K < 0 || a->G < 1) bad_arg("epi_rollout_cost_batch: bad B/K/G");
    if (a->L < 1 || a->L > EPI_LMAX) bad_arg("epi_rollout_cost_batch: L must be in 1..12");
    if (a->u_kind != EPI_U_F64 && a->u_kind != EPI_U_U8 && a->u_kind != EPI_U_PHILOX)
      bad_arg("epi_rollout_cost_batch: bad u_kind");
    const bool gen = a->u_kind == EPI_U_PHILOX;
    if (gen && (a->first < 0 || a->first % a->G != 0))
      bad_arg("epi_rollout_cost_batch: first must be a non-negative multiple of G");
    if ((a->J0 == nullptr) != (a->J1 == nullptr)) bad_arg("epi_rollout_cost_batch: J0 and J1 go together");
    if (a->B == 0) return;
    if (!a->prm || !a->x0 || (a->K > 0 && !a->u && !gen)) bad_arg("epi_rollout_cost_batch: prm, x0 and u are required");
    if (a->J0 && a->T_total < 1) bad_arg("epi_rollout_cost_batch: T_total must be >= 1 when costs are requested");
    reset_phases(c);
    const long long B = a->B;
    const int K = a->K, L = a->L;
    const long long n_groups = (B + a->G - 1) / a->G;
    if (gen && a->mem == EPI_MEM_HOST) validate_levels_host(a->prm, n_groups, L, "epi_rollout_cost_batch");
    Call shared(c, a->mem);
    const epi_model_params *prm = shared.in(a->prm, (size_t)n_groups);
    const double *x0 = shared.in(a->x0, (size_t)3 * n_groups);
    const double *nstd = shared.in(a->noise_std, (size_t)3 * n_groups);
    const double *j0p = shared.in(a->j0_prefix, (size_t)n_groups);
    const double *j1p = shared.in(a->j1_prefix, (size_t)n_groups);
    const double *wts = shared.in(a->w, (size_t)n_groups * K * L);
    const size_t usz = a->u_kind == EPI_U_F64 ? 8 : gen ? 0 : 1;
    size_t per = (size_t)K * L * usz + (a->noise ? (size_t)K * 24 : 0) +
                 ((a->s ? 1 : 0) + (a->i ? 1 : 0) + (a->alpha ? 1 : 0)) * (size_t)K * 8 + 16;
    const long long Bw = a->mem == EPI_MEM_HOST ? plan_wave(c, B, per) : B;
    for (long long b0 = 0; b0 < B; b0 += Bw) {
      const long long nb = (B - b0 < Bw) ? (B - b0) : Bw;
      Call w(c, a->mem);
      RolloutParams p{};
      p.B = (int)nb; p.K = K; p.L = L; p.G = a->G; p.b0 = b0;
      p.prm = prm; p.x0 = x0; p.noise_std = nstd;
      p.u_kind = a->u_kind;
      p.seed = a->seed;
      p.first = a->first;
      if (gen) {
        // nothing to stage: the kernel draws the schedule from (seed, first + b0 + b)
      } else if (a->u_kind == EPI_U_F64)
        p.u = w.traj_in_raw<double>((const double *)a->u, (size_t)K * L, B, b0, nb, &p.u_stride, &p.u_off);
      else
        p.u = w.traj_in_raw<unsigned char>((const unsigned char *)a->u, (size_t)K * L, B, b0, nb, &p.u_stride, &p.u_off);
      p.noise = w.traj_in(a->noise, (size_t)K * 3, B, b0, nb);
      p.s = w.traj_out(a->s, K, B, b0, nb);
      p.i = w.traj_out(a->i, K, B, b0, nb);
      p.alpha = w.traj_out(a->alpha, K, B, b0, nb);
      p.T_total = a->T_total; p.T_hist = 0;
      p.j0_prefix = j0p; p.j1_prefix = j1p; p.w = wts;
      p.J0 = w.traj_out(a->J0, 1, B, b0, nb);
      p.J1 = w.traj_out(a->J1, 1, B, b0, nb);
      PhaseScope ph(c, "rollout_cost");
      launch_rollout(p, c->stream);
      check_launch(c, 1);
      ph.end();
      w.flush();
    }
    finish(c, a->mem);
  });
}

extern "C" int epi_random_schedules(epi_ctx *c, const epi_schedules_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->B < 0 || a->K < 0 || a->G < 1) bad_arg("epi_random_schedules: bad B/K/G");
    if (a->L < 1 || a->L > EPI_LMAX) bad_arg("epi_random_schedules: L must be in 1..12");
    if (a->first < 0 || a->first % a->G != 0) bad_arg("epi_random_schedules: first must be a non-negative multiple of G");
    if (a->B == 0 || a->K == 0) return;
    if (!a->prm || !a->u) bad_arg("epi_random_schedules: prm and u are required");
    reset_phases(c);
    const long long B = a->B;
    const long long n_groups = (B + a->G - 1) / a->G;
    if (a->mem == EPI_MEM_HOST) validate_levels_host(a->prm, n_groups, a->L, "epi_random_schedules");
    Call shared(c, a->mem);
    const epi_model_params *prm = shared.in(a->prm, (size_t)n_groups);
    unsigned char *u = shared.out(a->u, (size_t)a->K * a->L * B);
    PhaseScope ph(c, "random_schedules");
    launch_random_schedules(prm, a->seed, a->first, (int)B, a->K, a->L, a->G, u, B, 0, c->stream);
    check_launch(c, 1);
    ph.end();
    shared.flush();
    finish(c, a->mem);
  });
}

extern "C" int epi_npicost_batch(epi_ctx *c, const epi_npicost_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->B < 0 || a->T < 1 || a->G < 1) bad_arg("epi_npicost_batch: bad B/T/G");
    if (a->L < 1 || a->L > EPI_LMAX) bad_arg("epi_npicost_batch: L must be in 1..12");
    if (a->B == 0) return;
    if (!a->newcases || !a->inputs || !a->weights || !a->J0 || !a->J1) bad_arg("epi_npicost_batch: null array");
    reset_phases(c);
    const long long B = a->B;
    const long long n_groups = (B + a->G - 1) / a->G;
    Call shared(c, a->mem);
    const double *wts = shared.in(a->weights, (size_t)n_groups * a->T * a->L);
    const long long Bw = a->mem == EPI_MEM_HOST
                             ? plan_wave(c, B, ((size_t)a->T * (a->L + 1) + 2) * 8) : B;
    for (long long b0 = 0; b0 < B; b0 += Bw) {
      const long long nb = (B - b0 < Bw) ? (B - b0) : Bw;
      Call w(c, a->mem);
      CostParams p{};
      p.B = (int)nb; p.T = a->T; p.L = a->L; p.G = a->G; p.b0 = b0;
      p.newcases = w.traj_in(a->newcases, a->T, B, b0, nb);
      p.inputs = w.traj_in(a->inputs, (size_t)a->T * a->L, B, b0, nb);
      p.weights = wts;
      p.J0 = w.traj_out(a->J0, 1, B, b0, nb);
      p.J1 = w.traj_out(a->J1, 1, B, b0, nb);
      PhaseScope ph(c, "npicost");
      launch_npicost(p, c->stream);
      check_launch(c, 1);
      ph.end();
      w.flush();
    }
    finish(c, a->mem);
  });
}

extern "C" int epi_si_controlled_batch(epi_ctx *c, const epi_si_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->B < 0 || a->K < 0) bad_arg("epi_si_controlled_batch: bad B/K");
    if (a->B == 0 || a->K == 0) return;
    if (!a->alpha || !a->beta || !a->s0 || !a->i0 || !a->s || !a->i) bad_arg("epi_si_controlled_batch: null array");
    reset_phases(c);
    const long long B = a->B;
    const long long Bw = a->mem == EPI_MEM_HOST ? plan_wave(c, B, ((size_t)3 * a->K + 3) * 8) : B;
    for (long long b0 = 0; b0 < B; b0 += Bw) {
      const long long nb = (B - b0 < Bw) ? (B - b0) : Bw;
      Call w(c, a->mem);
      SiParams p{};
      p.B = (int)nb; p.K = a->K; p.dt = a->dt;
      p.alpha = w.traj_in(a->alpha, a->K, B, b0, nb);
      p.beta = w.traj_in(a->beta, 1, B, b0, nb);
      p.s0 = w.traj_in(a->s0, 1, B, b0, nb);
      p.i0 = w.traj_in(a->i0, 1, B, b0, nb);
      p.s = w.traj_out(a->s, a->K, B, b0, nb);
      p.i = w.traj_out(a->i, a->K, B, b0, nb);
      PhaseScope ph(c, "si_controlled");
      launch_si(p, c->stream);
      check_launch(c, 1);
      ph.end();
      w.flush();
    }
    finish(c, a->mem);
  });
}

extern "C" int epi_preprocess_batch(epi_ctx *c, const epi_preprocess_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->B < 0 || a->L < 0 || a->L > EPI_LMAX || a->n_first < 0) bad_arg("epi_preprocess_batch: bad B/L/n_first");
    if (a->W < 1 || a->W > 32) bad_arg("epi_preprocess_batch: W must be in 1..32");
    if (a->T < 2) bad_arg("epi_preprocess_batch: Insufficient data (T < 2)");                    // :166
    const int Wh = (a->W + 1) / 2, nfact = (3 * (Wh - 1) > 1) ? 3 * (Wh - 1) : 1;
    if (a->T <= nfact) bad_arg("epi_preprocess_batch: Data length must be larger than 3 times the filter order");
    if (a->B == 0) return;
    if (!a->cc || !a->population || (a->L > 0 && !a->ip)) bad_arg("epi_preprocess_batch: cc, population and ip are required");
    reset_phases(c);
    const size_t B = (size_t)a->B, T = (size_t)a->T, L = (size_t)a->L;
    Call w(c, a->mem);
    PreprocParams p{};
    p.B = a->B; p.T = a->T; p.L = a->L; p.W = a->W; p.n_first = a->n_first; p.min_cases = a->min_cases;
    p.cc = w.in(a->cc, T * B);
    p.population = w.in(a->population, B);
    p.ip_in = w.in(a->ip, T * L * B);
    auto out_or_scratch = [&](double *o, size_t n) { return o ? w.out(o, n) : (double *)w.dalloc((n ? n : 1) * sizeof(double)); };
    p.ip_out = out_or_scratch(a->ip_filled, T * L * B);
    p.refined = out_or_scratch(a->refined, T * B);
    p.smoothed = out_or_scratch(a->smoothed, T * B);
    p.zerolag = out_or_scratch(a->zerolag, T * B);
    p.normalized = out_or_scratch(a->normalized, T * B);
    p.confirmed_norm = out_or_scratch(a->confirmed_norm, T * B);
    p.R_v = out_or_scratch(a->R_v, T * B);
    p.I0 = out_or_scratch(a->I0, B);
    p.scratch = (double *)w.dalloc(2 * (T + 2 * (size_t)nfact) * B * sizeof(double));
    PhaseScope ph(c, "preprocess");
    launch_preprocess(p, c->stream);
    check_launch(c, 1);
    ph.end();
    w.flush();
    finish(c, a->mem);
  });
}

extern "C" int epi_nnls_affine_batch(epi_ctx *c, const epi_nnls_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->B < 0 || a->n < 1 || a->max_alt < 0) bad_arg("epi_nnls_affine_batch: bad B/n/max_alt");
    if (a->p < 1 || a->p > EPI_LMAX) bad_arg("epi_nnls_affine_batch: p must be in 1..12");
    if (a->B == 0) return;
    if (!a->X || !a->y || !a->a || !a->b) bad_arg("epi_nnls_affine_batch: X, y, a and b are required");
    reset_phases(c);
    const size_t B = (size_t)a->B;
    Call w(c, a->mem);
    NnlsParams q{};
    q.B = a->B; q.n = a->n; q.p = a->p; q.max_alt = a->max_alt;
    q.X = w.in(a->X, (size_t)a->n * a->p * B);
    q.y = w.in(a->y, (size_t)a->n * B);
    q.a = w.out(a->a, (size_t)a->p * B);
    q.b = w.out(a->b, B);
    q.n_alt = w.out(a->n_alt, B);
    PhaseScope ph(c, "nnls_affine");
    launch_nnls_affine(q, c->stream);
    check_launch(c, 1);
    ph.end();
    w.flush();
    finish(c, a->mem);
  });
}

extern "C" int epi_rt_expfit_batch(epi_ctx *c, const epi_rt_expfit_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->B < 0 || a->T < 0 || a->G < 1 || a->W < 1) bad_arg("epi_rt_expfit_batch: bad B/T/G/W");
    if (a->order != 1 && a->order != 2) bad_arg("Undefined order", EPI_ERR_ORDER);  // Rt_ExpFitEKF.m:47,75
    if (a->W > kMaxDynSmem / (3 * 8 * 64))
      bad_arg("inv_monitor_len = " + std::to_string(a->W) + " exceeds the window the device kernel holds in shared memory (max " +
              std::to_string(kMaxDynSmem / (3 * 8 * 64)) + ")");
    if (a->B == 0 || a->T == 0) return;
    if (!a->x || !a->s_init || !a->params || !a->w_bar || !a->Ps_init || !a->Q || !a->R)
      bad_arg("epi_rt_expfit_batch: a required array is null");
    reset_phases(c);
    const long long B = a->B;
    const int T = a->T;
    const long long n_groups = (B + a->G - 1) / a->G;
    Call shared(c, a->mem);
    const double *params = shared.in(a->params, (size_t)3 * n_groups);
    const double *w_bar = shared.in(a->w_bar, (size_t)2 * n_groups);
    const double *Ps_init = shared.in(a->Ps_init, (size_t)4 * n_groups);
    const double *Q = shared.in(a->Q, (size_t)4 * n_groups);
    const double *R = shared.in(a->R, (size_t)n_groups);
    const size_t per = (size_t)T * (1 + 2 + 2 + 4 + 4 + 2 + 2 + 4 + 1 + 1) * 8 + 16;
    const long long Bw = a->mem == EPI_MEM_HOST ? plan_wave(c, B, per) : B;
    for (long long b0 = 0; b0 < B; b0 += Bw) {
      const long long nb = (B - b0 < Bw) ? (B - b0) : Bw;
      Call w(c, a->mem);
      RtParams p{};
      p.B = (int)nb; p.T = T; p.G = a->G; p.W = a->W; p.order = a->order; p.b0 = b0;
      p.x = w.traj_in(a->x, T, B, b0, nb);
      p.s_init = w.traj_in(a->s_init, 2, B, b0, nb);
      p.params = params; p.w_bar = w_bar; p.Ps_init = Ps_init; p.Q = Q; p.R = R;
      p.v_bar = a->v_bar; p.beta = a->beta; p.gamma = a->gamma;
      auto tape = [&](double *out, size_t rows) {
        return out ? w.traj_out(out, rows, B, b0, nb) : w.scratch(rows, nb);
      };
      p.S_MINUS = tape(a->S_MINUS, (size_t)T * 2); p.S_PLUS = tape(a->S_PLUS, (size_t)T * 2);
      p.P_MINUS = tape(a->P_MINUS, (size_t)T * 4); p.P_PLUS = tape(a->P_PLUS, (size_t)T * 4);
      p.K_GAIN = w.traj_out(a->K_GAIN, (size_t)T * 2, B, b0, nb);
      p.S_SMOOTH = w.traj_out(a->S_SMOOTH, (size_t)T * 2, B, b0, nb);
      p.P_SMOOTH = w.traj_out(a->P_SMOOTH, (size_t)T * 4, B, b0, nb);
      p.innov = w.traj_out(a->innovations, T, B, b0, nb);
      p.rho = w.traj_out(a->rho, T, B, b0, nb);
      // P_SMOOTH without S_SMOOTH still needs the backward pass
      if (p.P_SMOOTH.p && !p.S_SMOOTH.p) p.S_SMOOTH = w.scratch((size_t)T * 2, nb);
      PhaseScope ph(c, "rt_expfit");
      launch_rt_expfit(p, c->stream);
      check_launch(c, 1);
      ph.end();
      w.flush();
    }
    finish(c, a->mem);
  });
}

// ---------------------------------------------------------------------------
// EKF + smoother
// ---------------------------------------------------------------------------
namespace {

__global__ void validate_params_kernel(const epi_model_params *prm, long long n, int L, int model, int *flag) {
  for (long long g = threadIdx.x; g < n; g += blockDim.x) {
    int code = 0;
    if (prm[g].L != L) code = EPI_ERR_ARG;
    else if (model != EPI_MODEL_LEGACY_CODEGEN && prm[g].obs_type != EPI_OBS_NEWCASES && prm[g].obs_type != EPI_OBS_TOTALCASES)
      code = EPI_ERR_OBS_TYPE;
    if (code) atomicCAS(flag, 0, code);
  }
}

// the device-memory form of validate_params_host below: no host read, no stream drain
void validate_params_device(epi_ctx *c, const epi_model_params *prm, long long n, int L, int model) {
  if (n <= 0) return;
  if (!c->deferred_err) {
    CK(cudaHostAlloc((void **)&c->deferred_err, sizeof(int), cudaHostAllocMapped));
    *c->deferred_err = 0;
    CK(cudaHostGetDevicePointer((void **)&c->deferred_err_dev, c->deferred_err, 0));
  }
  validate_params_kernel<<<1, 128, 0, c->stream>>>(prm, n, L, model, c->deferred_err_dev);
  check_launch(c, 1);
}

void validate_params_host(const epi_model_params *prm, long long n, int L, int model) {
  for (long long g = 0; g < n; ++g) {
    if (prm[g].L != L) bad_arg("epi_model_params.L does not match args.L");
    if (model != EPI_MODEL_LEGACY_CODEGEN && prm[g].obs_type != EPI_OBS_NEWCASES &&
        prm[g].obs_type != EPI_OBS_TOTALCASES)
      bad_arg("unknown observation type", EPI_ERR_OBS_TYPE);  // SIAlphaModelEKF.m:57,87
  }
}

}  // namespace

extern "C" int epi_ekf_eks_batch(epi_ctx *c, const epi_ekf_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->model < EPI_MODEL_SIALPHA || a->model > EPI_MODEL_LEGACY_CODEGEN) bad_arg("epi_ekf_eks_batch: unknown model");
    if (a->order != 1 && a->order != 2) bad_arg("Undefined order", EPI_ERR_ORDER);  // GenericEKF.m:111,151
    if (a->B < 0 || a->T < 1 || a->G < 1 || a->W < 1) bad_arg("epi_ekf_eks_batch: bad B/T/G/W");
    if (a->L < 1 || a->L > EPI_LMAX) bad_arg("epi_ekf_eks_batch: L must be in 1..12");
    if (a->q_mode < EPI_Q_CONST || a->q_mode > EPI_Q_PERDAY_FULL)
      bad_arg("Process noise covariance noise mismatch", EPI_ERR_QR_SHAPE);  // :75
    if (a->r_mode != EPI_R_CONST && a->r_mode != EPI_R_PERDAY)
      bad_arg("Observation noise covariance noise mismatch", EPI_ERR_QR_SHAPE);  // :90
    // the innovation-monitor window (GenericExtendedKalmanFilter.m:170-181) is a shared-memory ring of
    // 3 * W doubles per thread (csrc/ekf_forward.cu: 64-thread CTAs for m = 3, 32 for m = 6; 227 KB per CTA)
    {
      const int w_max = kMaxDynSmem / (3 * 8 * (model_dim(a->model) == 6 ? 32 : 64));
      if (a->W > w_max)
        bad_arg("inv_monitor_len = " + std::to_string(a->W) + " exceeds the window the device kernels hold in shared memory (max " +
                std::to_string(w_max) + " for this model)");
    }
    const bool legacy = model_legacy(a->model);
    if (legacy && (a->q_mode != EPI_Q_CONST || a->r_mode != EPI_R_CONST))
      bad_arg("the legacy estimator takes a constant Q and a scalar R", EPI_ERR_QR_SHAPE);
    if (a->B == 0) return;
    if (!a->prm || !a->u || !a->x || !a->R || !a->Q || !a->s_init || !a->Ps_init || !a->s_final || !a->Ps_final)
      bad_arg("epi_ekf_eks_batch: a required input array is null");
    reset_phases(c);
    const int M = model_dim(a->model), MM = M * M, T = a->T, L = a->L;
    const long long B = a->B;
    const long long n_groups = (B + a->G - 1) / a->G;
    if (a->mem == EPI_MEM_HOST) validate_params_host(a->prm, n_groups, L, a->model);
    else validate_params_device(c, a->prm, n_groups, L, a->model);

    Call shared(c, a->mem);
    const epi_model_params *prm = shared.in(a->prm, (size_t)n_groups);
    const double *u_grp = a->u_per_traj ? nullptr : shared.in(a->u, (size_t)n_groups * T * L);
    const double *x_grp = a->x_per_traj ? nullptr : shared.in(a->x, (size_t)n_groups * T);
    const double *R_grp = a->r_per_traj ? nullptr
                                        : shared.in(a->R, (size_t)n_groups * (a->r_mode == EPI_R_CONST ? 1 : T));
    const size_t q_per = a->q_mode == EPI_Q_CONST ? (size_t)MM : a->q_mode == EPI_Q_PERDAY_SCALAR ? (size_t)T : (size_t)T * MM;
    const double *Q = shared.in(a->Q, (size_t)n_groups * q_per);
    const double *s_init_g = nullptr, *Ps_init_g = nullptr, *s_final_g = nullptr, *Ps_final_g = nullptr;
    if (!a->init_per_traj) {
      s_init_g = shared.in(a->s_init, (size_t)n_groups * M);
      Ps_init_g = shared.in(a->Ps_init, (size_t)n_groups * MM);
      s_final_g = shared.in(a->s_final, (size_t)n_groups * M);
      Ps_final_g = shared.in(a->Ps_final, (size_t)n_groups * MM);
    }

    // scratch / staging bytes per trajectory
    const bool host = a->mem == EPI_MEM_HOST;
    // tile layout (+ packed symmetric P pages) when the whole tape is library scratch
    const bool tiled = !a->S_MINUS && !a->S_PLUS && !a->P_MINUS && !a->P_PLUS;
    const int PF = (tiled && !legacy) ? M * (M + 1) / 2 : MM;
    // per-(group, day) input pre-pass (only meaningful when u is shared by the group)
    double *dot_grp = nullptr;
    if (!a->u_per_traj) {
      dot_grp = (double *)shared.dalloc((size_t)n_groups * T * 8);
      PhaseScope ph(c, "group_day");
      launch_group_day(prm, u_grp, nullptr, (int)n_groups, T, L, dot_grp, nullptr, c->stream);
      check_launch(c, 1);
      ph.end();
    }
    size_t per = (size_t)(T - 1) * MM * 8;  // J
    if (!a->S_MINUS || host) per += (size_t)T * M * 8;
    if (!a->S_PLUS || host) per += (size_t)T * M * 8;
    if (!a->P_MINUS || host) per += (size_t)T * PF * 8;
    if (!a->P_PLUS || host) per += (size_t)T * PF * 8;
    if (host) {
      if (a->u_per_traj) per += (size_t)T * L * 8;
      if (a->x_per_traj) per += (size_t)T * 8;
      if (a->r_per_traj) per += (size_t)(a->r_mode == EPI_R_CONST ? 1 : T) * 8;
      if (a->init_per_traj) per += (size_t)(2 * M + 2 * MM) * 8;
      per += ((a->u_opt ? 1 : 0) + (a->u_opt_smooth ? 1 : 0)) * (size_t)T * L * 8;
      per += ((a->S_SMOOTH ? 1 : 0) + (a->K_GAIN ? 1 : 0)) * (size_t)T * M * 8;
      per += (a->P_SMOOTH ? (size_t)T * MM * 8 : 0);
      per += ((a->innovations ? 1 : 0) + (a->rho ? 1 : 0)) * (size_t)T * 8 + 16;
    }
    const long long Bw = plan_wave(c, B, per);

    for (long long b0 = 0; b0 < B; b0 += Bw) {
      const long long nb = (B - b0 < Bw) ? (B - b0) : Bw;
      Call w(c, a->mem);
      EkfParams p{};
      p.model = a->model; p.B = (int)nb; p.T = T; p.L = L; p.G = a->G; p.W = a->W; p.b0 = b0;
      p.prm = prm;
      p.epsilon = w.traj_in(a->epsilon, 1, B, b0, nb);
      p.eps_grid = nullptr; p.eps_mod = 1;
      p.u_grp = u_grp; p.x_grp = x_grp; p.R_grp = R_grp;
      p.u_trj = a->u_per_traj ? w.traj_in(a->u, (size_t)T * L, B, b0, nb) : CArr{nullptr, 0, 0};
      p.x_trj = a->x_per_traj ? w.traj_in(a->x, T, B, b0, nb) : CArr{nullptr, 0, 0};
      p.R_trj = a->r_per_traj ? w.traj_in(a->R, a->r_mode == EPI_R_CONST ? 1 : T, B, b0, nb) : CArr{nullptr, 0, 0};
      p.r_mode = a->r_mode; p.fixed_R = a->fixed_R; p.q_mode = a->q_mode; p.Q = Q;
      p.init_per_traj = a->init_per_traj;
      p.s_init_g = s_init_g; p.Ps_init_g = Ps_init_g; p.s_final_g = s_final_g; p.Ps_final_g = Ps_final_g;
      if (a->init_per_traj) {
        p.s_init_t = w.traj_in(a->s_init, M, B, b0, nb);
        p.Ps_init_t = w.traj_in(a->Ps_init, MM, B, b0, nb);
        p.s_final_t = w.traj_in(a->s_final, M, B, b0, nb);
        p.Ps_final_t = w.traj_in(a->Ps_final, MM, B, b0, nb);
      }
      if (model_flipped(a->model)) {
        // SIAlphaModelBackwardEKF.m:21-24: the time-reversed filter starts from
        // (s_final, Ps_final) and its smoother ends on (s_init, Ps_init)
        std::swap(p.s_init_g, p.s_final_g); std::swap(p.Ps_init_g, p.Ps_final_g);
        std::swap(p.s_init_t, p.s_final_t); std::swap(p.Ps_init_t, p.Ps_final_t);
      }
      p.v_bar = a->v_bar; p.beta = a->beta; p.gamma = a->gamma;
      p.tiled = tiled ? 1 : 0;
      p.dot_grp = dot_grp; p.cost_grp = nullptr;
      if (tiled) {
        p.S_MINUS = w.scratch_tiled((size_t)T * M, nb);
        p.S_PLUS = w.scratch_tiled((size_t)T * M, nb);
        p.P_MINUS = w.scratch_tiled((size_t)T * PF, nb);
        p.P_PLUS = w.scratch_tiled((size_t)T * PF, nb);
      } else {
        p.S_MINUS = a->S_MINUS ? w.traj_out(a->S_MINUS, (size_t)T * M, B, b0, nb) : w.scratch((size_t)T * M, nb);
        p.S_PLUS = a->S_PLUS ? w.traj_out(a->S_PLUS, (size_t)T * M, B, b0, nb) : w.scratch((size_t)T * M, nb);
        p.P_MINUS = a->P_MINUS ? w.traj_out(a->P_MINUS, (size_t)T * MM, B, b0, nb) : w.scratch((size_t)T * MM, nb);
        p.P_PLUS = a->P_PLUS ? w.traj_out(a->P_PLUS, (size_t)T * MM, B, b0, nb) : w.scratch((size_t)T * MM, nb);
      }
      p.J = w.scratch_tiled((size_t)(T > 1 ? T - 1 : 1) * MM, nb);
      p.u_opt = w.traj_out(a->u_opt, (size_t)T * L, B, b0, nb);
      p.u_opt_smooth = legacy ? TArr{nullptr, 0, 0} : w.traj_out(a->u_opt_smooth, (size_t)T * L, B, b0, nb);
      p.S_SMOOTH = w.traj_out(a->S_SMOOTH, (size_t)T * M, B, b0, nb);
      p.P_SMOOTH = w.traj_out(a->P_SMOOTH, (size_t)T * MM, B, b0, nb);
      p.K_GAIN = w.traj_out(a->K_GAIN, (size_t)T * M, B, b0, nb);
      p.innov = w.traj_out(a->innovations, T, B, b0, nb);
      p.rho = w.traj_out(a->rho, T, B, b0, nb);
      p.status = nullptr;
      if (a->status) {
        if (host) {
          p.status = (int *)w.dalloc((size_t)nb * sizeof(int));
          w.pend.push_back({a->status + b0, p.status, (size_t)nb * 4, (size_t)nb * 4, (size_t)nb * 4, 1});
        } else {
          p.status = a->status + b0;
        }
        CK(cudaMemsetAsync(p.status, 0, (size_t)nb * sizeof(int), c->stream));
      }
      p.T_hist = 0;
      {
        PhaseScope ph(c, "ekf_forward");
        launch_ekf_forward(p, c->stream);
        check_launch(c, 1);
        ph.end();
      }
      if (T > 1) {
        PhaseScope ph(c, "eks_gain");
        launch_eks_gain(p, c->stream);
        check_launch(c, 1);
        ph.end();
      }
      {
        PhaseScope ph(c, "eks_backward");
        launch_eks_backward(p, c->stream);
        check_launch(c, 1);
        ph.end();
      }
      w.flush();
    }
    finish(c, a->mem);
  });
}

// ---------------------------------------------------------------------------
// Pareto
// ---------------------------------------------------------------------------
extern "C" int epi_pareto_batch(epi_ctx *c, const epi_pareto_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->n_sets < 0 || a->n < 0) bad_arg("epi_pareto_batch: bad sizes");
    if (a->n_sets == 0 || a->n == 0) return;
    if (!a->J0 || !a->J1) bad_arg("epi_pareto_batch: J0 and J1 are required");
    if ((size_t)a->n_sets * (size_t)a->n > 0x7fffffffull) bad_arg("epi_pareto_batch: more than 2^31 points");
    reset_phases(c);
    Call k(c, a->mem);
    const size_t tot = (size_t)a->n_sets * a->n;
    ParetoParams p{};
    p.n_sets = a->n_sets; p.n = a->n;
    p.J0 = k.in(a->J0, tot); p.J1 = k.in(a->J1, tot);
    p.on_front = k.out(a->on_front, tot);
    p.I_opt = k.out(a->I_opt, (size_t)a->n_sets);
    PhaseScope ph(c, "pareto");
    if (a->n <= kParetoBruteMax) {
      launch_pareto(p, c->stream);
      check_launch(c, 1);
    } else {
      const size_t sb = pareto_sorted_scratch_bytes(a->n_sets, a->n);
      void *scratch = k.dalloc(sb);
      check_launch(c, launch_pareto_sorted(p, scratch, sb, c->stream));
    }
    ph.end();
    k.flush();
    finish(c, a->mem);
  });
}

// ---------------------------------------------------------------------------
// fused optimal-NPI Pareto sweep
// ---------------------------------------------------------------------------
extern "C" int epi_sweep(epi_ctx *c, const epi_sweep_args *a) {
  return guarded(c, [&] {
    if (!a) bad_arg("null args");
    check_mem(a->mem);
    if (a->n_regions < 0 || a->n_eps < 1 || a->T < 1 || a->T_hist < 0 || a->T_hist > a->T || a->W < 1)
      bad_arg("epi_sweep: bad sizes");
    if (a->L < 1 || a->L > EPI_LMAX) bad_arg("epi_sweep: L must be in 1..12");
    if (a->n_regions == 0) return;
    if (!a->prm || !a->eps || !a->u || !a->x || !a->R || !a->s_init || !a->Ps_init || !a->s_final ||
        !a->Ps_final || !a->Q || !a->x0 || !a->weights || !a->J0 || !a->J1 || (a->T_hist > 0 && !a->newcases_hist))
      bad_arg("epi_sweep: a required array is null");
    if (a->u_knee && !a->I_opt) bad_arg("epi_sweep: u_knee needs I_opt");
    if (a->lean && a->P_first) bad_arg("epi_sweep: P_first needs the full smoother (lean = 0)");
    reset_phases(c);
    HostTrace tr("epi_sweep");
    const int M = 6, MM = 36, PF = 21, T = a->T, L = a->L, Tf = a->T - a->T_hist;
    const long long nR = a->n_regions, B = nR * a->n_eps;
    const bool host = a->mem == EPI_MEM_HOST;
    if (host) validate_params_host(a->prm, nR, L, EPI_MODEL_OPTCTRL);
    else validate_params_device(c, a->prm, nR, L, EPI_MODEL_OPTCTRL);
    tr.mark("validate");

    Call shared(c, a->mem);
    const epi_model_params *prm = shared.in(a->prm, (size_t)nR);
    const double *eps = shared.in(a->eps, (size_t)a->n_eps);
    const double *u = shared.in(a->u, (size_t)nR * T * L);
    const double *x = shared.in(a->x, (size_t)nR * T);
    const double *R = shared.in(a->R, (size_t)nR * T);
    const double *s_init = shared.in(a->s_init, (size_t)nR * M);
    const double *Ps_init = shared.in(a->Ps_init, (size_t)nR * MM);
    const double *s_final = shared.in(a->s_final, (size_t)nR * M);
    const double *Ps_final = shared.in(a->Ps_final, (size_t)nR * MM);
    const double *Q = shared.in(a->Q, (size_t)nR * MM);
    const double *x0 = shared.in(a->x0, (size_t)nR * 3);
    const double *nch = shared.in(a->newcases_hist, (size_t)nR * a->T_hist);
    // The cost weights are first read by the backward pass: in host mode they are uploaded on the side
    // stream, beside the forward pass, and the launching stream picks them up through an event.
    const double *wts = nullptr;
    bool wts_deferred = false;
    struct CopyGuard {  // never leave a copy in flight into a block that is about to be recycled
      cudaStream_t s = nullptr;
      ~CopyGuard() { if (s) cudaStreamSynchronize(s); }
    } copy_guard;
    if (host && !getenv("EPI_NO_SIDE_COPY")) {
      if (!c->copy_stream) {
        CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->copy_ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->copy_ev[1], cudaEventDisableTiming));
      }
      double *d = (double *)shared.dalloc((size_t)nR * T * L * sizeof(double));
      CK(cudaEventRecord(c->copy_ev[0], c->stream));  // the block may still be read by work queued earlier
      CK(cudaStreamWaitEvent(c->copy_stream, c->copy_ev[0], 0));
      CK(cudaMemcpyAsync(d, a->weights, (size_t)nR * T * L * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
      CK(cudaEventRecord(c->copy_ev[1], c->copy_stream));
      copy_guard.s = c->copy_stream;
      wts = d;
      wts_deferred = true;
    } else {
      wts = shared.in(a->weights, (size_t)nR * T * L);
    }
    const double *nstd = shared.in(a->noise_std, (size_t)nR * 3);
    // whole-batch outputs (the Pareto step needs every epsilon of a region)
    double *J0 = shared.out(a->J0, (size_t)B);
    double *J1 = shared.out(a->J1, (size_t)B);
    unsigned char *on_front = shared.out(a->on_front, (size_t)B);
    int *I_opt = shared.out(a->I_opt, (size_t)nR);
    double *u_knee = shared.out(a->u_knee, (size_t)nR * Tf * L);
    // the knee gather needs every smoothed schedule on the device
    double *u_fore_dev = nullptr;
    if (a->u_fore && !host) u_fore_dev = a->u_fore;
    else if (a->u_fore || a->u_knee) { u_fore_dev = (double *)shared.dalloc((size_t)Tf * L * B * 8); }
    if (a->u_fore && host)
      shared.pend.push_back({a->u_fore, u_fore_dev, (size_t)Tf * L * B * 8, (size_t)Tf * L * B * 8, (size_t)Tf * L * B * 8, 1});

    tr.mark("h2d_enqueue");
    // per-(region, day) input pre-pass: days whose NPIs are all given share their input term and cost
    double *dot_grp = (double *)shared.dalloc((size_t)nR * T * 8);
    double *cost_grp = (double *)shared.dalloc((size_t)nR * T * 8);
    {
      PhaseScope ph(c, "group_day");
      // (deferred weights: the input terms now, the per-day costs once the weights have arrived)
      launch_group_day(prm, u, wts_deferred ? nullptr : wts, (int)nR, T, L, dot_grp, wts_deferred ? nullptr : cost_grp,
                       c->stream);
      check_launch(c, 1);
      ph.end();
    }
    // lean: the smoother is needed from the first day to optimise on (k0 = T_hist); keep at least
    // the last day so the terminal condition has its tape page
    const int k0 = a->lean ? (a->T_hist < T - 1 ? a->T_hist : T - 1) : 0;
    const int Tn = T - k0;  // days of tape kept
    size_t per = (size_t)(Tn > 1 ? Tn - 1 : 1) * MM * 8 + (size_t)2 * Tn * M * 8 + (size_t)2 * Tn * PF * 8 + (size_t)2 * T * 8;
    if (host && a->noise) per += (size_t)Tf * 24;
    if (host && a->P_first) per += (size_t)MM * 8;
    const long long Bw = plan_wave(c, B, per);
    tr.mark("budget");

    double *hist_pre0 = nullptr, *hist_pre1 = nullptr;  // per-region NPICost sums over the history (launch_hist_prefix)
    for (long long b0 = 0; b0 < B; b0 += Bw) {
      const long long nb = (B - b0 < Bw) ? (B - b0) : Bw;
      Call w(c, a->mem);
      EkfParams p{};
      p.model = EPI_MODEL_OPTCTRL; p.B = (int)nb; p.T = T; p.L = L; p.G = a->n_eps; p.W = a->W; p.b0 = b0;
      p.prm = prm;
      p.epsilon = CArr{nullptr, 0, 0};
      p.eps_grid = eps; p.eps_mod = a->n_eps;
      p.u_grp = u; p.x_grp = x; p.R_grp = R;
      p.r_mode = EPI_R_PERDAY; p.fixed_R = 0; p.q_mode = EPI_Q_CONST; p.Q = Q;
      p.init_per_traj = 0;
      p.s_init_g = s_init; p.Ps_init_g = Ps_init; p.s_final_g = s_final; p.Ps_final_g = Ps_final;
      p.v_bar = 0.0; p.beta = a->beta_ekf; p.gamma = a->gamma_ekf;
      p.tiled = 1;
      p.k0 = k0;
      p.dot_grp = dot_grp; p.cost_grp = cost_grp;
      p.S_MINUS = w.scratch_tiled((size_t)Tn * M, nb);
      p.S_PLUS = w.scratch_tiled((size_t)Tn * M, nb);
      p.P_MINUS = w.scratch_tiled((size_t)Tn * PF, nb);
      p.P_PLUS = w.scratch_tiled((size_t)Tn * PF, nb);
      p.J = w.scratch_tiled((size_t)(Tn > 1 ? Tn - 1 : 1) * MM, nb);
      p.dot_day = w.scratch_tiled(T, nb);
      p.cost_day = w.scratch_tiled(T, nb);
      p.weights = wts;
      p.T_hist = a->T_hist;
      if (u_fore_dev) p.u_fore = TArr{u_fore_dev, B, b0};
      p.P_first = w.traj_out(a->P_first, MM, B, b0, nb);
      {
        // few-wave batches: time-segmented persistent forward launch (no idle tail)
        static int slots6 = 0;
        if (!slots6) slots6 = forward_resident_slots6();
        const long long tiles = (nb + 31) / 32;
        p.fwd_segments = a->beta_ekf == 1.0 ? forward_segments(tiles, slots6) : 1;
        if (const char *e = getenv("EPI_FWD_SEGMENTS")) p.fwd_segments = atoi(e);  // tuning experiments
        p.bwd_prefetch = 3;   // measured at 7.5k / 14.7k / 29.5k / 59k trajectories: 0.57 / 0.63 / 1.07 / 2.15 ms against 1.02 / 1.09 / 1.38 / 2.31
        if (const char *e = getenv("EPI_BWD_PREFETCH")) p.bwd_prefetch = atoi(e);
        if (p.fwd_segments > 1) {
          p.fwd_sync = (int *)w.dalloc((size_t)(tiles + 1) * sizeof(int));
          CK(cudaMemsetAsync(p.fwd_sync, 0, (size_t)(tiles + 1) * sizeof(int), c->stream));
        }
      }
      // small shard (a region block of the strong-scaling sweep): forward || gains, see csrc/ekf_forward.cu
      bool piped = false;
      int want = 1;  // EPI_PIPE: 0 = one-stream schedule, 2 = the piped kernels one after the other (tuning experiments)
      {
        if (const char *e = getenv("EPI_PIPE")) want = atoi(e);
        piped = want != 0 && Tn > 1 && pipe_ready(c) && forward_piped_ok(p, c->n_sms);
      }
      if (piped) {
        p.pipe_chunks = 8;
        if (const char *e = getenv("EPI_PIPE_CHUNKS")) p.pipe_chunks = atoi(e) > 0 ? atoi(e) : 8;
        if (p.pipe_chunks > 64) p.pipe_chunks = 64;
        p.pipe_sync = (unsigned *)w.dalloc((size_t)p.pipe_chunks * sizeof(unsigned));
        const unsigned tiles = (unsigned)((nb + 31) / 32);
        CK(cudaMemsetAsync(p.pipe_sync, 0, (size_t)p.pipe_chunks * sizeof(unsigned), c->stream));
        CK(cudaEventRecord(c->pipe_ev[0], c->stream));        // fork: the counters are zero, the inputs in place
        CK(cudaStreamWaitEvent(c->pipe_stream, c->pipe_ev[0], 0));
        {
          PhaseScope ph(c, "ekf_forward_piped");  // (beside the gains of the chunks it has finished)
          launch_ekf_forward_piped(p, c->stream);
          check_launch(c, 1);
          ph.end();
        }
        PhaseScope ph(c, "eks_gain_tail");  // on the launching stream: what is left of the gains when the forward pass ends
        try {
        for (int ch = 0; ch < p.pipe_chunks; ++ch) {
          int kb = 0, ke = 0;
          pipe_chunk_days(p, ch, kb, ke);
          EkfParams g = p;
          g.gk_lo = kb < k0 ? k0 : kb;
          g.gk_hi = ke < T - 1 ? ke : T - 1;
          if (g.gk_hi <= g.gk_lo) continue;
          // all tiles have left chunk ch: the tape holds P-(k + 1), P+(k), S+(k) for every k < ke
          if (want == 2) { launch_eks_gain(g, c->stream); check_launch(c, 1); continue; }
          if (((WaitValue32Fn)c->wait_value32)(c->pipe_stream, (unsigned long long)(uintptr_t)(p.pipe_sync + ch), tiles, 1u) != 0)
            throw EpiError{EPI_ERR_CUDA, "cuStreamWaitValue32 failed"};
          launch_eks_gain(g, c->pipe_stream);
          check_launch(c, 1);
        }
        } catch (...) {  // never leave the second stream parked on a counter: release every queued wait
          cudaMemsetAsync(p.pipe_sync, 0xff, (size_t)p.pipe_chunks * sizeof(unsigned), c->stream);
          throw;
        }
        CK(cudaEventRecord(c->pipe_ev[1], c->pipe_stream));   // join
        CK(cudaStreamWaitEvent(c->stream, c->pipe_ev[1], 0));
        ph.end();
      } else {
      {
        PhaseScope ph(c, "ekf_forward");
        launch_ekf_forward(p, c->stream);
        check_launch(c, 1);
        ph.end();
      }
      if (Tn > 1) {
        PhaseScope ph(c, "eks_gain");
        launch_eks_gain(p, c->stream);
        check_launch(c, 1);
        ph.end();
      }
      }
      if (wts_deferred) {  // first consumer of the weights and of cost_grp
        CK(cudaStreamWaitEvent(c->stream, c->copy_ev[1], 0));
        launch_group_day(prm, u, wts, (int)nR, T, L, dot_grp, cost_grp, c->stream);  // same dot_grp bits again + costs
        check_launch(c, 1);
        wts_deferred = false;
      }
      {
        PhaseScope ph(c, "eks_backward");
        launch_eks_backward(p, c->stream);
        check_launch(c, 1);
        ph.end();
      }
      RolloutParams r{};
      // the cost of a history day whose NPIs are all given is the same for every trajectory of the region: read it per
      // group (L1) instead of 441 per-trajectory tape lines; days with missing NPIs keep the per-trajectory value
      r.hist_cost_grp = cost_grp;
      r.hist_cost_per_traj = a->lean ? 0 : 1;
      if (!hist_pre0) {  // once per call, when the per-group day costs are complete (the weights may have arrived late)
        hist_pre0 = (double *)shared.dalloc((size_t)nR * 8);
        hist_pre1 = (double *)shared.dalloc((size_t)nR * 8);
        launch_hist_prefix(nch, cost_grp, (int)nR, a->T_hist, T, hist_pre0, hist_pre1, c->stream);
        check_launch(c, 1);
      }
      r.j0_prefix = hist_pre0; r.j1_prefix = hist_pre1;
      r.B = (int)nb; r.K = Tf; r.L = L; r.G = a->n_eps; r.b0 = b0;
      r.prm = prm; r.x0 = x0; r.noise_std = nstd;
      r.u_kind = 2;
      r.noise = w.traj_in(a->noise, (size_t)Tf * 3, B, b0, nb);
      r.T_total = T; r.T_hist = a->T_hist;
      r.newcases_hist = nch;
      r.dot_day = p.dot_day.p;
      r.cost_day = p.cost_day.p;
      r.J0 = TArr{J0, B, b0};
      r.J1 = TArr{J1, B, b0};
      {
        PhaseScope ph(c, "rollout_cost");
        launch_rollout(r, c->stream);
        check_launch(c, 1);
        ph.end();
      }
      w.flush();
    }
    if (on_front || I_opt) {
      ParetoParams pp{};
      pp.n_sets = (int)nR; pp.n = a->n_eps; pp.J0 = J0; pp.J1 = J1; pp.on_front = on_front; pp.I_opt = I_opt;
      PhaseScope ph(c, "pareto");
      if (a->n_eps <= kParetoBruteMax) {
        launch_pareto(pp, c->stream);
        check_launch(c, 1);
      } else {
        const size_t sb = pareto_sorted_scratch_bytes((int)nR, a->n_eps);
        void *scratch = shared.dalloc(sb);
        check_launch(c, launch_pareto_sorted(pp, scratch, sb, c->stream));
      }
      if (u_knee) {
        launch_gather_knee(u_fore_dev, I_opt, u_knee, (int)nR, a->n_eps, Tf, L, c->stream);
        check_launch(c, 1);
      }
      ph.end();
    }
    shared.flush();
    tr.mark("launch_enqueue");
    finish(c, a->mem);
    tr.mark("sync");
  });
}

// -------------------------------------------------------------------------------------------------
// epi_sweep_multi: the region-sharded sweep on several GPUs from ONE blocking host call
// (Tools/TrainPredictPrescribeNPI.m:93 is the region loop, :421 the epsilon loop).  Regions are independent
// (SURVEY 8e), so shard i -- a contiguous block of regions -- is one epi_sweep on context i, driven by its
// own host thread; every per-region input of epi_sweep_args is region-major, so a shard's arguments are the
// caller's pointers advanced to its first region, and J0/J1/front/knee land directly in the caller's rows:
// the "gather" of a single-process host is the D2H copy of each shard.  The three trajectory-minor arrays
// (noise in, u_fore / P_first out: [..][B]) go through per-shard staging buffers.
namespace {
struct ShardPlan { int lo, hi; };
ShardPlan shard_of(int n_regions, int n_shards, int i) {
  const int per = (n_regions + n_shards - 1) / n_shards;
  ShardPlan s;
  s.lo = std::min(n_regions, i * per);
  s.hi = std::min(n_regions, s.lo + per);
  return s;
}
}  // namespace

extern "C" int epi_sweep_multi(epi_ctx *const *ctxs, int n_ctx, const epi_sweep_args *a) {
  if (!ctxs || n_ctx < 1 || !a) return EPI_ERR_ARG;
  for (int i = 0; i < n_ctx; ++i)
    if (!ctxs[i]) return EPI_ERR_ARG;
  if (n_ctx == 1) return epi_sweep(ctxs[0], a);
  if (a->mem != EPI_MEM_HOST) {
    ctxs[0]->err = "epi_sweep_multi: EPI_MEM_HOST only (device pointers belong to one GPU; shard on the caller's side "
                   "and call epi_sweep per context instead)";
    return EPI_ERR_ARG;
  }
  if (a->n_regions < 0 || a->n_eps < 1 || a->T < 1 || a->T_hist < 0 || a->T_hist > a->T || a->L < 1) {
    ctxs[0]->err = "epi_sweep_multi: bad sizes";
    return EPI_ERR_ARG;
  }
  const size_t nE = (size_t)a->n_eps, T = (size_t)a->T, L = (size_t)a->L, Th = (size_t)a->T_hist, Tf = T - Th;
  const size_t B = (size_t)a->n_regions * nE;
  std::vector<int> rc((size_t)n_ctx, EPI_OK);
  std::vector<std::thread> th;
  th.reserve((size_t)n_ctx);
  for (int i = 0; i < n_ctx; ++i) {
    const ShardPlan sp = shard_of(a->n_regions, n_ctx, i);
    if (sp.hi <= sp.lo) continue;
    th.emplace_back([=, &rc]() {
      const size_t r0 = (size_t)sp.lo, nr = (size_t)(sp.hi - sp.lo), b0 = r0 * nE, nb = nr * nE;
      epi_sweep_args s = *a;
      s.n_regions = (int)nr;
      auto adv = [](auto *p, size_t n) { return p ? p + n : p; };
      s.prm = adv(a->prm, r0);
      s.u = adv(a->u, r0 * T * L);
      s.x = adv(a->x, r0 * T);
      s.R = adv(a->R, r0 * T);
      s.s_init = adv(a->s_init, r0 * 6); s.s_final = adv(a->s_final, r0 * 6);
      s.Ps_init = adv(a->Ps_init, r0 * 36); s.Ps_final = adv(a->Ps_final, r0 * 36); s.Q = adv(a->Q, r0 * 36);
      s.x0 = adv(a->x0, r0 * 3);
      s.newcases_hist = adv(a->newcases_hist, r0 * Th);
      s.weights = adv(a->weights, r0 * T * L);
      s.noise_std = adv(a->noise_std, r0 * 3);
      s.J0 = adv(a->J0, b0); s.J1 = adv(a->J1, b0);
      s.on_front = adv(a->on_front, b0);
      s.I_opt = adv(a->I_opt, r0);
      s.u_knee = adv(a->u_knee, r0 * Tf * L);
      // trajectory-minor arrays: rows of B values, this shard owns columns [b0, b0 + nb)
      std::vector<double> noise, u_fore, P_first;
      auto cut = [&](const double *src, size_t rows, std::vector<double> &dst) {
        dst.resize(rows * nb);
        for (size_t r = 0; r < rows; ++r) memcpy(dst.data() + r * nb, src + r * B + b0, nb * sizeof(double));
      };
      auto paste = [&](const std::vector<double> &src, size_t rows, double *dst) {
        for (size_t r = 0; r < rows; ++r) memcpy(dst + r * B + b0, src.data() + r * nb, nb * sizeof(double));
      };
      if (a->noise) { cut(a->noise, Tf * 3, noise); s.noise = noise.data(); }
      if (a->u_fore) { u_fore.resize(Tf * L * nb); s.u_fore = u_fore.data(); }
      if (a->P_first) { P_first.resize(36 * nb); s.P_first = P_first.data(); }
      rc[(size_t)i] = epi_sweep(ctxs[i], &s);
      if (rc[(size_t)i] == EPI_OK) {
        if (a->u_fore) paste(u_fore, Tf * L, a->u_fore);
        if (a->P_first) paste(P_first, 36, a->P_first);
      }
    });
  }
  for (auto &t : th) t.join();
  for (int i = 0; i < n_ctx; ++i)
    if (rc[(size_t)i] != EPI_OK) {
      if (i != 0) ctxs[0]->err = "shard " + std::to_string(i) + ": " + ctxs[i]->err;
      return rc[(size_t)i];
    }
  return EPI_OK;
}
