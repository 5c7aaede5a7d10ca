// pk_engine.cu -- host side of the C ABI (include/pockit_b200.h): device memory, one stream,
// NVRTC compilation of the generated per-node programs for sm_100a, and the launch sequences
// of the five callbacks.
//
// Launch sequence of a mode (all on the engine stream):
//   node programs (one per phase, NVRTC)  ->  pk_reduce_rows  ->  system program (NVRTC)
//   ->  pk_defects | pk_generic_jobs + pk_expand_blocks | pk_grad_range + pk_grad_scalar
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "pk_kernels.cuh"

static thread_local std::string g_err;

static int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +         \
                  std::to_string(__LINE__) + ")");                                                 \
  } while (0)

// timing event that is destroyed on every return path
struct ScopedEvent {
  cudaEvent_t ev = nullptr;
  cudaError_t create() { return cudaEventCreate(&ev); }
  ~ScopedEvent() {
    if (ev) cudaEventDestroy(ev);
  }
  operator cudaEvent_t() const { return ev; }
};

// one launch of the block expansion: jobs [first, first + count) of the mode's EXPAND stage
struct ExpandGroup {
  long long first = 0, count = 0;
  bool lam = false;
  // persistent column kernel (any mix of orders)
  long long* prefix = nullptr; long long blocks = 0; int uniform = 0; size_t smem = 0;
  // parameter-driven column kernel (same-order meshes, few jobs / lists)
  bool cols = false; PkXcParams xc; unsigned xc_gx = 0; size_t xc_smem = 0;
  // bulk-store variant of it (opt-in): whole intervals per block, image handed to the TMA engine
  bool bulk = false; int xb_per_block = 0; unsigned xb_gx = 0; size_t xb_smem = 0;
  // batches of small problems: parameter-driven, (instance, pair) space flattened over the grid
  bool batch = false; unsigned xbt_gx = 0; int batch_lists = PK_XM_LISTS, batch_rows = PK_XM_ROWS; size_t batch_smem = 0; unsigned n_groups = 0;
  bool slots = false; unsigned xsl_gx = 0; int slot_lists = PK_XS_LISTS;  // pk_expand_slots (the default batch kernel)
};

struct ModeState {
  bool loaded = false;
  cudaLibrary_t lib = nullptr;
  std::vector<cudaKernel_t> node_kernels;
  std::vector<pk_node_program> node_programs;
  cudaKernel_t sys_kernel = nullptr;
  long long n_scalar = 0, n_out = 0;
  double* OUT = nullptr;  // [B][n_out], owned by the mode so that a full evaluation set stays resident
  double* OUT_alloc = nullptr;  // the allocation; OUT starts 0..3 doubles into it (see load_mode)
  double *S = nullptr, *W = nullptr;  // scalar / node tables of this mode (modes may run concurrently)
  cudaStream_t stream = nullptr;      // used when a whole evaluation set is launched at once
  cudaEvent_t done = nullptr;
  cudaEvent_t node_done = nullptr;    // recorded behind the per-node programs when a set is staggered (run_set)
  cudaEvent_t node_fork = nullptr, node_join[2] = {nullptr, nullptr};  // phases' per-node programs side by side
  cudaStream_t side = nullptr;        // the small slot runs overlap the block expansion
  cudaEvent_t side_fork = nullptr, side_join = nullptr;
  // two more branches of the small-kernel DAG (defects / gradient gather run beside the reduction chain)
  cudaStream_t aux1 = nullptr, aux2 = nullptr;
  cudaEvent_t sys_done = nullptr, aux1_join = nullptr, aux2_join = nullptr;
  std::vector<char> cubin;
  pk_job* jobs[PK_N_STAGES] = {};
  long long n_jobs[PK_N_STAGES] = {};
  // block -> (job, chunk) maps of the two slot-streaming kernels
  int* gen_job = nullptr; int* gen_chunk = nullptr; long long gen_blocks = 0;
  std::vector<ExpandGroup> exp;  // block expansion, one launch per group
  long long max_defect_rows = 0, max_grad_count = 0, max_reduce_len = 0;
  double* red_partial = nullptr; unsigned* red_ticket = nullptr; int red_parts = 0;
  bool def_fast = false, def_table = false;
  struct DefectGeo { bool fast; long long rows, n_x, rb; };
  std::vector<DefectGeo> def_geo;  // launch geometry of every defect job (same-order fast path)
  bool idx32 = true;  // every flattened (instance, slot) space fits 32-bit index math
  long long grad_off = 0, grad_cnt = 0;
  long long sub_off[PK_N_CALLBACKS] = {}, sub_cnt[PK_N_CALLBACKS] = {};
  bool in_set = false;  // callback mode: the latest values live in the set pipeline's combined output
  // mesh sharding: the (offset, count) runs of the output this engine computes; empty = all of it
  std::vector<long long> dl_runs;
  // de-duplicated pattern: OUTC[u] = sum of OUT[perm[ptr[u] .. ptr[u+1])]
  long long n_compact = 0;
  unsigned *cp_ptr = nullptr, *cp_perm = nullptr;
  double* OUTC = nullptr;
};

struct pk_engine {
  pk_dims dims;
  int device = 0;
  cudaStream_t stream = nullptr;
  double *X = nullptr, *LAM = nullptr, *SIG = nullptr, *FIX = nullptr;
  cudaEvent_t fork = nullptr, x_done = nullptr, lam_done = nullptr;
  cudaGraphExec_t set_graph = nullptr;  // captured evaluation set (all requested modes, one stream each)
  std::vector<int> set_modes;
  double* dpool = nullptr; long long* ipool = nullptr;
  std::vector<long long> h_ipool;  // host copy (list tables are folded into kernel parameters)
  double *hX = nullptr, *hLAM = nullptr, *hSIG = nullptr;  // pinned staging
  double* flush = nullptr; long long n_flush = 0;
  long long n_out_max = 0;
  ModeState mode[PK_N_MODES];
  struct AugPhase {
    cudaKernel_t prep = nullptr, node = nullptr;
    long long x_offset = 0, L = 0, L_xu = 0, L_x_all = 0, n_x = 0, n_u = 0, Lm_aug = 0, rows = 0;
    double *tm = nullptr, *XS = nullptr, *XU = nullptr, *WA = nullptr, *TX = nullptr, *IF = nullptr;
    long long *Vp = nullptr, *Vi = nullptr, *Tp = nullptr, *Ti = nullptr, *Ip = nullptr, *Ii = nullptr;
    double *Vv = nullptr, *Tv = nullptr, *Iv = nullptr;
  };
  std::vector<AugPhase> aug;  // continuous error estimate (pk_engine_load_error_estimate)
  cudaLibrary_t aug_lib = nullptr;
  std::vector<char> aug_cubin;
  long long launches = 0, set_launches = 0, x_uploads = 0;
  bool x_resident = false;
  // pk_timeline: timing events around every launch of a set
  struct Mark { cudaEvent_t ev; int mode, tag, edge; };
  std::vector<Mark>* trace = nullptr;
  // set launches: the HBM-bound block expansions of different modes are chained one after the other
  // (each then streams at full bandwidth while the next mode's latency-bound prologue overlaps it)
  cudaEvent_t chain_ev[PK_N_MODES] = {};
  int chain_last = -1;
  bool chain = false;
  bool stagger = false;  // set capture: record node_done behind every mode's per-node programs
  // modes started by pk_eval_set_async whose completion (kernels + device-to-host copy) the engine stream
  // does not yet depend on: joined before anything that could disturb them, and by pk_sync
  std::vector<int> pending;
};

// tag: job stage 0..5, 6 node programs, 7 system program, 8 compaction; edge 0 = before, 1 = after
static void tr(pk_engine* e, int mode, int tag, int edge, cudaStream_t s) {
  if (!e->trace) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, s);
  e->trace->push_back({ev, mode, tag, edge});
}

extern "C" int pk_abi_version(void) { return PK_ABI_VERSION; }
extern "C" const char* pk_last_error(void) { return g_err.c_str(); }

extern "C" int pk_device_count(int* count) {
  CK(cudaGetDeviceCount(count));
  return 0;
}

extern "C" void* pk_alloc_host(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void pk_free_host(void* p) {
  if (p) cudaFreeHost(p);
}

// Page-lock a caller-owned host range (e.g. a shared-memory mapping that several ranks copy
// their shares of a result into) so device-to-host copies into it run at full PCIe rate.
extern "C" int pk_host_register(void* p, size_t bytes) {
  if (!p || !bytes) return fail("pk_host_register: bad argument");
  CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return 0;
}
extern "C" int pk_host_unregister(void* p) {
  if (!p) return 0;
  CK(cudaHostUnregister(p));
  return 0;
}

static int engine_allocate(pk_engine* e);
extern "C" int pk_engine_destroy(pk_engine* e);

extern "C" int pk_engine_create(const pk_dims* d, int device, pk_engine** out) {
  if (!d || !out) return fail("pk_engine_create: null argument");
  if (d->abi_version != PK_ABI_VERSION) return fail("pk_engine_create: ABI version mismatch");
  if (d->batch < 1) return fail("pk_engine_create: batch must be >= 1");
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail("pk_engine_create: no such CUDA device");
  CK(cudaSetDevice(device));
  pk_engine* e = new pk_engine();
  e->dims = *d;
  e->device = device;
  if (engine_allocate(e)) {  // release whatever was allocated before the failure
    const std::string why = g_err;
    pk_engine_destroy(e);
    g_err = why;
    return 1;
  }
  *out = e;
  return 0;
}

static int engine_allocate(pk_engine* e) {
  const pk_dims* d = &e->dims;
  const long long B = d->batch;
  long long n_out = d->L;
  if (d->m > n_out) n_out = d->m;
  if (d->nnz_jac > n_out) n_out = d->nnz_jac;
  if (d->nnz_hess > n_out) n_out = d->nnz_hess;
  if (n_out < 1) n_out = 1;
  e->n_out_max = 1 + d->L + d->m + d->nnz_jac + d->nnz_hess;  // the set pipeline holds all five
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  auto dev = [&](double** p, long long count) { return cudaMalloc((void**)p, sizeof(double) * (size_t)(count > 0 ? count : 1)); };
  CK(dev(&e->X, B * d->L));
  CK(dev(&e->LAM, B * d->m));
  CK(dev(&e->SIG, B));
  CK(dev(&e->FIX, B * d->n_fixed));
  CK(cudaMemsetAsync(e->LAM, 0, sizeof(double) * (size_t)(B * d->m > 0 ? B * d->m : 1), e->stream));
  CK(cudaMemsetAsync(e->SIG, 0, sizeof(double) * (size_t)B, e->stream));
  CK(cudaEventCreateWithFlags(&e->fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->x_done, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->lam_done, cudaEventDisableTiming));
  for (auto& ev : e->chain_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CK(cudaMallocHost((void**)&e->hX, sizeof(double) * (size_t)(B * d->L > 0 ? B * d->L : 1)));
  CK(cudaMallocHost((void**)&e->hLAM, sizeof(double) * (size_t)(B * d->m > 0 ? B * d->m : 1)));
  CK(cudaMallocHost((void**)&e->hSIG, sizeof(double) * (size_t)B));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

static void free_mode(ModeState& ms) {
  for (int s = 0; s < PK_N_STAGES; ++s) {
    if (ms.jobs[s]) cudaFree(ms.jobs[s]);
    ms.jobs[s] = nullptr;
  }
  if (ms.gen_job) cudaFree(ms.gen_job);
  if (ms.gen_chunk) cudaFree(ms.gen_chunk);
  for (auto& g : ms.exp)
    if (g.prefix) cudaFree(g.prefix);
  if (ms.lib) cudaLibraryUnload(ms.lib);
  if (ms.OUT_alloc) cudaFree(ms.OUT_alloc);
  if (ms.S) cudaFree(ms.S);
  if (ms.W) cudaFree(ms.W);
  if (ms.stream) cudaStreamDestroy(ms.stream);
  if (ms.done) cudaEventDestroy(ms.done);
  if (ms.node_done) cudaEventDestroy(ms.node_done);
  if (ms.node_fork) cudaEventDestroy(ms.node_fork);
  for (auto& ev : ms.node_join)
    if (ev) cudaEventDestroy(ev);
  if (ms.side) cudaStreamDestroy(ms.side);
  if (ms.side_fork) cudaEventDestroy(ms.side_fork);
  if (ms.side_join) cudaEventDestroy(ms.side_join);
  if (ms.aux1) cudaStreamDestroy(ms.aux1);
  if (ms.aux2) cudaStreamDestroy(ms.aux2);
  if (ms.sys_done) cudaEventDestroy(ms.sys_done);
  if (ms.aux1_join) cudaEventDestroy(ms.aux1_join);
  if (ms.aux2_join) cudaEventDestroy(ms.aux2_join);
  if (ms.red_partial) cudaFree(ms.red_partial);
  if (ms.red_ticket) cudaFree(ms.red_ticket);
  if (ms.cp_ptr) cudaFree(ms.cp_ptr);
  if (ms.cp_perm) cudaFree(ms.cp_perm);
  if (ms.OUTC) cudaFree(ms.OUTC);
  ms = ModeState();
}

static void free_aug(pk_engine* e) {
  for (auto& a : e->aug) {
    void* ptrs[] = {a.tm, a.XS, a.XU, a.WA, a.TX, a.IF, a.Vp, a.Vi, a.Tp, a.Ti, a.Ip, a.Ii, a.Vv, a.Tv, a.Iv};
    for (void* q : ptrs)
      if (q) cudaFree(q);
  }
  e->aug.clear();
  if (e->aug_lib) cudaLibraryUnload(e->aug_lib);
  e->aug_lib = nullptr;
}

extern "C" int pk_engine_destroy(pk_engine* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  free_aug(e);
  for (auto& ms : e->mode) free_mode(ms);
  cudaFree(e->X); cudaFree(e->LAM); cudaFree(e->SIG);
  if (e->set_graph) cudaGraphExecDestroy(e->set_graph);
  if (e->fork) cudaEventDestroy(e->fork);
  if (e->x_done) cudaEventDestroy(e->x_done);
  if (e->lam_done) cudaEventDestroy(e->lam_done);
  for (auto& ev : e->chain_ev)
    if (ev) cudaEventDestroy(ev);
  cudaFree(e->FIX); cudaFree(e->dpool); cudaFree(e->ipool); cudaFree(e->flush);
  cudaFreeHost(e->hX); cudaFreeHost(e->hLAM); cudaFreeHost(e->hSIG);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return 0;
}

extern "C" int pk_engine_set_pools(pk_engine* e, const double* dp, int64_t nd, const int64_t* ip, int64_t ni) {
  if (!e) return fail("pk_engine_set_pools: null engine");
  CK(cudaSetDevice(e->device));
  cudaFree(e->dpool); cudaFree(e->ipool);
  e->dpool = nullptr; e->ipool = nullptr;
  CK(cudaMalloc((void**)&e->dpool, sizeof(double) * (size_t)(nd > 0 ? nd : 1)));
  CK(cudaMalloc((void**)&e->ipool, sizeof(long long) * (size_t)(ni > 0 ? ni : 1)));
  if (nd > 0) CK(cudaMemcpy(e->dpool, dp, sizeof(double) * (size_t)nd, cudaMemcpyHostToDevice));
  if (ni > 0) CK(cudaMemcpy(e->ipool, ip, sizeof(long long) * (size_t)ni, cudaMemcpyHostToDevice));
  e->h_ipool.assign(ip, ip + (ni > 0 ? ni : 0));
  return 0;
}

extern "C" int pk_engine_set_fixed(pk_engine* e, const double* v) {
  if (!e) return fail("pk_engine_set_fixed: null engine");
  CK(cudaSetDevice(e->device));
  const size_t n = (size_t)e->dims.batch * (size_t)e->dims.n_fixed;
  if (n) CK(cudaMemcpy(e->FIX, v, sizeof(double) * n, cudaMemcpyHostToDevice));
  return 0;
}

static int compile_source(const char* source, const char* const* extra, int n_extra, int device, std::vector<char>& cubin);
static int compile(const pk_mode_desc* d, int device, std::vector<char>& cubin) {
  return compile_source(d->cuda_source, d->nvrtc_options, d->n_nvrtc_options, device, cubin);
}

// Process-wide cubin cache keyed by (architecture, options, source).  The generated programs read
// every size and offset from a __constant__ table, so re-meshing a model (set_discretization ->
// System.update(): a new engine) produces the same source and costs no NVRTC compilation.
static std::mutex g_cache_mutex;
static std::map<std::string, std::vector<char>> g_cubin_cache;
static long long g_cache_hits = 0, g_cache_misses = 0;

extern "C" int pk_cubin_cache_stats(int64_t* hits, int64_t* misses) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  if (hits) *hits = g_cache_hits;
  if (misses) *misses = g_cache_misses;
  return 0;
}

static int compile_uncached(const char* source, const std::vector<const char*>& opts, std::vector<char>& cubin);

static int compile_source(const char* source, const char* const* extra, int n_extra, int device, std::vector<char>& cubin) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  // B200 is sm_100: compile for the arch-specific target; anything else gets its own real arch
  std::string arch = "--gpu-architecture=sm_" + std::to_string(prop.major * 10 + prop.minor) +
                     ((prop.major == 10 && prop.minor == 0) ? "a" : "");
  std::vector<const char*> opts = {arch.c_str(), "--std=c++17", "-lineinfo"};
  bool fmad_given = false;  // strict IEEE products by default; a fastmath model passes --fmad=true
  for (int i = 0; i < n_extra; ++i) {
    if (strncmp(extra[i], "--fmad", 6) == 0 || strncmp(extra[i], "-fmad", 5) == 0) fmad_given = true;
    opts.push_back(extra[i]);
  }
  if (!fmad_given) opts.push_back("--fmad=false");
  std::string key;
  for (const char* o : opts) { key += o; key += '\n'; }
  key += '\0';
  key += source;
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    auto it = g_cubin_cache.find(key);
    if (it != g_cubin_cache.end()) {
      cubin = it->second;
      ++g_cache_hits;
      return 0;
    }
  }
  if (compile_uncached(source, opts, cubin)) return 1;
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  ++g_cache_misses;
  if (g_cubin_cache.size() >= 64) g_cubin_cache.clear();  // bounded: a long sweep over models must not grow without limit
  g_cubin_cache[key] = cubin;
  return 0;
}

static int compile_uncached(const char* source, const std::vector<const char*>& opts, std::vector<char>& cubin) {
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, source, "pockit_b200_generated.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
    return fail("nvrtcCreateProgram failed");
  nvrtcResult r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
  if (r != NVRTC_SUCCESS) {
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    nvrtcGetProgramLog(prog, &log[0]);
    nvrtcDestroyProgram(&prog);
    return fail("NVRTC compilation failed:\n" + log);
  }
  size_t n = 0;
  if (nvrtcGetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) {
    nvrtcDestroyProgram(&prog);
    return fail("NVRTC produced no cubin");
  }
  cubin.resize(n);
  nvrtcGetCUBIN(prog, cubin.data());
  nvrtcDestroyProgram(&prog);
  return 0;
}

// m = floor(2^64 / d) + 1: floor(x / d) == umul64hi(x, m) for all x, d < 2^32; 0 encodes d == 1
static unsigned long long div_multiplier(unsigned long long d) {
  if (d <= 1) return 0;
  return (unsigned long long)((((unsigned __int128)1) << 64) / d) + 1ull;
}

static int build_block_map(const pk_job* jobs, long long n, int field, int per, long long batch, int** d_job, int** d_chunk, long long* n_blocks) {
  std::vector<int> bj, bc;
  for (long long j = 0; j < n; ++j) {
    long long chunks = (jobs[j].i[field] * batch + per - 1) / per;
    for (long long c = 0; c < chunks; ++c) {
      bj.push_back((int)j);
      bc.push_back((int)c);
    }
  }
  *n_blocks = (long long)bj.size();
  if (bj.empty()) return 0;
  CK(cudaMalloc((void**)d_job, sizeof(int) * bj.size()));
  CK(cudaMalloc((void**)d_chunk, sizeof(int) * bc.size()));
  CK(cudaMemcpy(*d_job, bj.data(), sizeof(int) * bj.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(*d_chunk, bc.data(), sizeof(int) * bc.size(), cudaMemcpyHostToDevice));
  return 0;
}

// pk_expand_batch is instantiated for a few unrolled row counts (the smallest one that holds the block rows
// is "exact"), with and without streaming stores
typedef void (*BatchKernel)(PkCtx, const PkXcParams, unsigned);
template <bool LAM, bool STREAM>
static BatchKernel batch_kernel_of(int rows) {
  if (rows <= 5) return pk_expand_batch<LAM, 5, STREAM>;
  if (rows <= 8) return pk_expand_batch<LAM, 8, STREAM>;
  if (rows <= 12) return pk_expand_batch<LAM, 12, STREAM>;
  return pk_expand_batch<LAM, PK_XM_ROWS, STREAM>;
}
static BatchKernel batch_kernel(bool lam, int rows, bool stream) {
  if (lam) return stream ? batch_kernel_of<true, true>(rows) : batch_kernel_of<true, false>(rows);
  return stream ? batch_kernel_of<false, true>(rows) : batch_kernel_of<false, false>(rows);
}
static BatchKernel slots_kernel(bool lam, bool stream) {
  if (lam) return stream ? pk_expand_slots<true, true> : pk_expand_slots<true, false>;
  return stream ? pk_expand_slots<false, true> : pk_expand_slots<false, false>;
}

static int setup_expand_group(pk_engine* e, const pk_job* ej, long long first, long long count, ExpandGroup& g) {
  g.first = first;
  g.count = count;
  g.lam = (ej[0].flags & PK_F_LAM) != 0;
  std::vector<long long> prefix(1, 0);
  g.uniform = 1;
  long long n_lists = 0, max_pairs = 0;
  for (long long j = 0; j < count; ++j) {
    const pk_job& jb = ej[j];
    if (jb.i[11] >= (1LL << 31) || jb.i[1] * jb.i[11] >= (1LL << 32)) return fail("expand job too large for 32-bit unit indices");
    prefix.push_back(prefix.back() + jb.i[1] * jb.i[11]);
    const size_t sm = sizeof(double) * (size_t)(jb.i[3] * jb.i[4]);
    if (sm > g.smem) g.smem = sm;
    if (jb.i[7] != ej[0].i[7] || jb.i[3] != ej[0].i[3] || jb.i[4] != ej[0].i[4] || jb.f[0] != ej[0].f[0]) g.uniform = 0;
    n_lists += jb.i[1];
    if (jb.i[11] > max_pairs) max_pairs = jb.i[11];
  }
  if (!g.uniform) g.smem = 0;
  if (g.smem > 200 * 1024) { g.uniform = 0; g.smem = 0; }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
  // parameter-driven column walk: large same-order meshes.  POCKIT_B200_EXPAND=columns forces the
  // persistent kernel, =params insists on the parameter-driven one.
  const long long n0 = ej[0].i[3], r0 = ej[0].i[4];
  const size_t xsm = sizeof(double) * (size_t)(n0 * r0 + (PK_XC_THREADS / n0 + 2) * r0);
  bool cols = g.uniform && count <= PK_XC_JOBS && n_lists <= PK_XC_LISTS && max_pairs >= 8 * PK_XC_THREADS &&
              e->dims.batch <= 65535 && xsm <= 200 * 1024;
  // batches of small problems (too few pairs per instance for the kernel above): the same parameter block
  // drives pk_expand_batch -- one thread per (instance, interval, column) of a job, the block column of
  // every list of the job written from registers.  POCKIT_B200_EXPAND=columns forces the persistent kernel.
  bool batch = !cols && g.uniform && count <= PK_XC_JOBS && n_lists <= PK_XC_LISTS && e->dims.batch > 1 && r0 <= PK_XM_ROWS &&
               max_pairs * e->dims.batch >= 8 * PK_XC_THREADS && max_pairs * e->dims.batch < (1LL << 32);
  if (const char* env = getenv("POCKIT_B200_EXPAND")) {
    if (!strcmp(env, "params") && !cols) return fail("POCKIT_B200_EXPAND=params: jobs do not fit the parameter-driven kernel");
    if (!strcmp(env, "batch") && !batch) return fail("POCKIT_B200_EXPAND=batch: jobs do not fit the batch kernel");
    if (!strcmp(env, "slots") && !(batch && max_pairs * r0 * e->dims.batch < (1LL << 32)))
      return fail("POCKIT_B200_EXPAND=slots: jobs do not fit the slot-order batch kernel");
    if (!strcmp(env, "columns")) cols = batch = false;
    if (!strcmp(env, "bulk") && !cols) return fail("POCKIT_B200_EXPAND=bulk: jobs do not fit the parameter-driven kernel");
  }
  if (cols || batch) {
    g.cols = cols;
    g.batch = batch;
    g.xbt_gx = (unsigned)((max_pairs * e->dims.batch + PK_XC_THREADS - 1) / PK_XC_THREADS);
    if (batch) {
      // Two thread mappings.  Slot order (pk_expand_slots: a thread per output slot, warps write consecutive
      // doubles, all lists of a job per thread) for groups whose jobs have three or more lists: it amortises its
      // per-slot index arithmetic over the lists of a job (quadrotor Jacobian, one job of 5 lists: 48.2 us against
      // 68.5).  The column mapping (pk_expand_batch) for the others (quadrotor Hessian, 2 jobs of 1-2 lists: 32.1 us
      // against 37.3).  configs[4] set 264 -> 243 us (profiles/r02_call29_batch_loads_ahead.log,
      // r02_final_stage_times_quadrotor.log, r02_call30_batch_mapping_per_group.log).
      // POCKIT_B200_EXPAND=batch | slots forces one of them for every group.
      const char* env = getenv("POCKIT_B200_EXPAND");
      long long most_lists = 0;
      for (long long j = 0; j < count; ++j) most_lists = ej[j].i[1] > most_lists ? ej[j].i[1] : most_lists;
      g.slots = env ? !strcmp(env, "slots") : (most_lists >= 3 && max_pairs * r0 * e->dims.batch < (1LL << 32));
      g.xsl_gx = (unsigned)((max_pairs * r0 * e->dims.batch + PK_XC_THREADS - 1) / PK_XC_THREADS);
      // experiment knobs: unrolled row count (exact = smallest instantiation that holds the rows; default 16) and
      // an unused dynamic shared-memory request that caps the resident blocks per SM
      if (const char* pr = getenv("POCKIT_B200_BATCH_ROWS"))
        if (!strcmp(pr, "exact")) g.batch_rows = (int)r0;
      if (const char* psm = getenv("POCKIT_B200_BATCH_SMEM")) {
        const long long v = atoll(psm);
        if (v > 0 && v <= 200 * 1024) {
          g.batch_smem = (size_t)v;
          if (v > 48 * 1024)
            for (int sw = 0; sw < 2; ++sw)
              CK(cudaFuncSetAttribute((const void*)batch_kernel(g.lam, g.batch_rows, sw != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v));
        }
      }
      if (const char* pl = getenv("POCKIT_B200_BATCH_LISTS")) {
        const int v = atoi(pl);
        if (v >= 1 && v <= PK_XC_LISTS) g.batch_lists = v;
      }
      if (const char* pl = getenv("POCKIT_B200_SLOT_LISTS")) {
        const int v = atoi(pl);
        if (v >= 1 && v <= PK_XC_LISTS) g.slot_lists = v;
      }
    }
    g.xc_smem = xsm;
    PkXcParams& q = g.xc;
    memset(&q, 0, sizeof(q));
    q.n = (int)n0; q.rows = (int)r0; q.n_jobs = (int)count; q.unit = ej[0].i[7]; q.sign = ej[0].f[0];
    q.m_n = div_multiplier((unsigned long long)n0);
    q.m_bn = div_multiplier((unsigned long long)(n0 * r0));
    int li = 0;
    for (int j = 0; j < q.n_jobs; ++j) {
      const pk_job& jb = ej[j];
      q.job[j].lam0 = jb.i[2]; q.job[j].node0 = jb.i[6]; q.job[j].width = jb.i[8]; q.job[j].Lm = jb.i[10];
      q.job[j].step = (int)jb.i[5]; q.job[j].pairs = (unsigned)jb.i[11];
      q.job[j].m_pairs = div_multiplier((unsigned long long)jb.i[11]);
      q.job[j].m_run = div_multiplier((unsigned long long)(jb.i[11] * r0));
      q.job[j].list0 = li; q.job[j].n_lists = (int)jb.i[1];
      for (long long l = 0; l < jb.i[1]; ++l, ++li) {
        const size_t at = (size_t)(jb.i[0] + 2 * l);
        if (at + 1 >= e->h_ipool.size()) return fail("expand job: list table outside the integer pool");
        q.list[li].dst = e->h_ipool[at];
        q.list[li].wbase = e->h_ipool[at + 1];
        q.list[li].job = j;
      }
    }
    q.n_lists = li;
    if (batch) {
      // blockIdx.y of the batch kernels -> (job, first list, count)
      const int per = g.slots ? g.slot_lists : g.batch_lists;
      g.n_groups = 0;
      for (int j = 0; j < q.n_jobs; ++j)
        for (int l = 0; l < q.job[j].n_lists; l += per) {
          const int cnt = q.job[j].n_lists - l < per ? q.job[j].n_lists - l : per;
          q.grp[g.n_groups++] = (unsigned)j | ((unsigned)(q.job[j].list0 + l) << 8) | ((unsigned)cnt << 20);
        }
    }
    g.xc_gx = (unsigned)((max_pairs + PK_XC_THREADS - 1) / PK_XC_THREADS);
    auto kern = g.lam ? (const void*)pk_expand_cols<true> : (const void*)pk_expand_cols<false>;
    if (xsm > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsm));
    // TMA bulk-store variant (same parameters): chosen when the interval blocks are NOT a whole number of
    // 32-byte sectors (LGL n = 10: 90 slots), where the column walk's 8-byte stores straddle sectors whatever
    // the base alignment; the bulk copy of a shared-memory image writes whole sectors.  Measured on B200
    // (round 2, profiles/r02_call15_bulk_lgl.log): humanoid set 128.7 -> 121.4 us, rocket 56.5 -> 55.5; on the
    // sector-aligned 20 x 20 blocks of robot_arm it is slower (56.7 -> 62.8), so the column walk stays there.
    // POCKIT_B200_EXPAND=bulk forces it, =params forbids it.
    {
      const char* env = getenv("POCKIT_B200_EXPAND");
      const bool want_bulk = env ? !strcmp(env, "bulk") : (((n0 * r0) & 3) != 0 && e->dims.batch == 1);
      if (cols && want_bulk && n0 <= PK_XB_THREADS) {
        const long long per = PK_XB_THREADS / n0, bn0 = n0 * r0;
        const size_t bsm = sizeof(double) * (size_t)(((bn0 + 1) & ~1LL) + ((per * r0 + 1) & ~1LL) + per * bn0 + 2);
        if (bsm <= 200 * 1024) {
          g.bulk = true;
          g.xb_per_block = (int)per;
          g.xb_smem = bsm;
          g.xb_gx = (unsigned)((max_pairs / n0 + per - 1) / per);
          auto bk = g.lam ? (const void*)pk_expand_bulk<true> : (const void*)pk_expand_bulk<false>;
          if (bsm > 48 * 1024) CK(cudaFuncSetAttribute(bk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
        }
      }
    }
    return 0;
  }
  if (g.smem > 48 * 1024)
    CK(cudaFuncSetAttribute(pk_expand_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  CK(cudaMalloc((void**)&g.prefix, sizeof(long long) * prefix.size()));
  CK(cudaMemcpy(g.prefix, prefix.data(), sizeof(long long) * prefix.size(), cudaMemcpyHostToDevice));
  // one resident wave: SMs x blocks/SM, shared between the instances of a batch
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk_expand_blocks, PK_THREADS, g.smem));
  const long long wave = (long long)(per_sm > 0 ? per_sm : 1) * sms;
  const long long need = (prefix.back() + PK_THREADS - 1) / PK_THREADS;
  long long gx = (wave + e->dims.batch - 1) / e->dims.batch;
  if (gx < 1) gx = 1;
  g.blocks = need < gx ? need : gx;
  return 0;
}

extern "C" int pk_engine_load_mode(pk_engine* e, int mode, const pk_mode_desc* d) {
  if (!e || !d) return fail("pk_engine_load_mode: null argument");
  if (mode < 0 || mode >= PK_N_MODES) return fail("pk_engine_load_mode: bad mode");
  CK(cudaSetDevice(e->device));
  ModeState& ms = e->mode[mode];
  free_mode(ms);
  if (d->n_scalar > e->dims.n_scalar) return fail("pk_engine_load_mode: n_scalar exceeds pk_dims.n_scalar");
  if (d->n_out > e->n_out_max) return fail("pk_engine_load_mode: n_out exceeds the engine's output buffer");
  if (compile(d, e->device, ms.cubin)) return 1;
  CK(cudaLibraryLoadData(&ms.lib, ms.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  for (int i = 0; i < d->n_node_programs; ++i) {
    cudaKernel_t k;
    CK(cudaLibraryGetKernel(&k, ms.lib, d->node_programs[i].kernel));
    ms.node_kernels.push_back(k);
    ms.node_programs.push_back(d->node_programs[i]);
    ms.node_programs.back().kernel = nullptr;
  }
  if (d->system_kernel) CK(cudaLibraryGetKernel(&ms.sys_kernel, ms.lib, d->system_kernel));
  if (d->n_table_entries > 0) {
    void* dptr = nullptr;
    size_t bytes = 0;
    CK(cudaLibraryGetGlobal(&dptr, &bytes, ms.lib, d->table_symbol));
    if (bytes < sizeof(long long) * (size_t)d->n_table_entries) return fail("constant table smaller than its contents");
    CK(cudaMemcpy(dptr, d->table, sizeof(long long) * (size_t)d->n_table_entries, cudaMemcpyHostToDevice));
  }
  ms.n_scalar = d->n_scalar;
  ms.n_out = d->n_out;
  ms.grad_off = d->grad_offset;
  ms.grad_cnt = d->grad_count;
  if (ms.grad_off < 0 || ms.grad_cnt < 0 || ms.grad_off + ms.grad_cnt > d->n_out) return fail("pk_engine_load_mode: gradient range outside the output");
  for (int k = 0; k < PK_N_CALLBACKS; ++k) {
    ms.sub_off[k] = d->sub_offset[k];
    ms.sub_cnt[k] = d->sub_count[k];
    if (ms.sub_off[k] < 0 || ms.sub_cnt[k] < 0 || ms.sub_off[k] + ms.sub_cnt[k] > d->n_out) return fail("pk_engine_load_mode: callback range outside the output");
  }
  for (auto& other : e->mode) other.in_set = false;
  {
    // The reference's slot offsets are only 8-byte aligned, but the block expansion writes most of
    // the bytes: start the output 0..3 doubles into its allocation so that the offset shared by most
    // expanded lists falls on a 32-byte sector boundary (robot_arm LGR 2000x20: all 15 Jacobian lists
    // start at 2 mod 4 -> 10.4 sectors per store request become 8).
    long long votes[4] = {0, 0, 0, 0};
    const pk_job* ej = d->jobs[PK_STAGE_EXPAND];
    for (long long j = 0; j < d->n_jobs[PK_STAGE_EXPAND]; ++j)
      for (long long l = 0; l < ej[j].i[1]; ++l) {
        const size_t at = (size_t)(ej[j].i[0] + 2 * l);
        if (at < e->h_ipool.size()) votes[e->h_ipool[at] & 3] += ej[j].i[11] * ej[j].i[4];
      }
    int best = 0;
    for (int k = 1; k < 4; ++k)
      if (votes[k] > votes[best]) best = k;
    const char* env = getenv("POCKIT_B200_ALIGN");
    const int shift = (env && env[0] == '0') ? 0 : (4 - best) & 3;
    CK(cudaMalloc((void**)&ms.OUT_alloc, sizeof(double) * ((size_t)e->dims.batch * (size_t)(d->n_out > 0 ? d->n_out : 1) + 4)));
    ms.OUT = ms.OUT_alloc + shift;
  }
  {
    const size_t ns = sizeof(double) * (size_t)e->dims.batch * (size_t)(d->n_scalar > 0 ? d->n_scalar : 1);
    const size_t nw = sizeof(double) * (size_t)(e->dims.n_table > 0 ? e->dims.n_table : 1);
    CK(cudaMalloc((void**)&ms.S, ns));
    CK(cudaMalloc((void**)&ms.W, nw));
    CK(cudaMemset(ms.S, 0, ns));
    // the latency-bound chains (small callbacks, and the side chain of the large ones) get the higher
    // priority: their few blocks must not queue behind the thousands of blocks of a block expansion
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const char* pe = getenv("POCKIT_B200_PRIORITY");
    const bool use_prio = !(pe && pe[0] == '0');
    const bool big = mode == PK_MODE_JACOBIAN || mode == PK_MODE_HESSIAN || mode == PK_MODE_SET;
    CK(cudaStreamCreateWithPriority(&ms.stream, cudaStreamNonBlocking, use_prio && !big ? prio_hi : prio_lo));
    CK(cudaEventCreateWithFlags(&ms.done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ms.node_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ms.node_fork, cudaEventDisableTiming));
    for (auto& ev : ms.node_join) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaStreamCreateWithPriority(&ms.side, cudaStreamNonBlocking, use_prio ? prio_hi : prio_lo));
    CK(cudaEventCreateWithFlags(&ms.side_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ms.side_join, cudaEventDisableTiming));
    CK(cudaStreamCreateWithPriority(&ms.aux1, cudaStreamNonBlocking, use_prio ? prio_hi : prio_lo));
    CK(cudaStreamCreateWithPriority(&ms.aux2, cudaStreamNonBlocking, use_prio ? prio_hi : prio_lo));
    CK(cudaEventCreateWithFlags(&ms.sys_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ms.aux1_join, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ms.aux2_join, cudaEventDisableTiming));
  }
  if (e->set_graph) {  // a re-loaded mode invalidates the captured set
    cudaGraphExecDestroy(e->set_graph);
    e->set_graph = nullptr;
    e->set_modes.clear();
  }
  for (int s = 0; s < PK_N_STAGES; ++s) {
    ms.n_jobs[s] = d->n_jobs[s];
    if (d->n_jobs[s] > 0) {
      std::vector<pk_job> host(d->jobs[s], d->jobs[s] + d->n_jobs[s]);
      // division multipliers the kernels use instead of runtime integer divisions (pk_div)
      if (s == PK_STAGE_GENERIC)
        for (pk_job& jb : host) jb.i[15] = (int64_t)div_multiplier((unsigned long long)(jb.i[1] > 0 ? jb.i[1] : 1));
      if (s == PK_STAGE_EXPAND)
        for (pk_job& jb : host) {
          jb.i[12] = (int64_t)div_multiplier((unsigned long long)(jb.i[11] > 0 ? jb.i[11] : 1));
          jb.i[13] = (int64_t)div_multiplier((unsigned long long)(jb.i[3] > 0 ? jb.i[3] : 1));
        }
      CK(cudaMalloc((void**)&ms.jobs[s], sizeof(pk_job) * host.size()));
      CK(cudaMemcpy(ms.jobs[s], host.data(), sizeof(pk_job) * host.size(), cudaMemcpyHostToDevice));
    }
  }
  if (build_block_map(d->jobs[PK_STAGE_GENERIC], d->n_jobs[PK_STAGE_GENERIC], 1, PK_CHUNK, e->dims.batch, &ms.gen_job, &ms.gen_chunk, &ms.gen_blocks)) return 1;
  // block expansion: consecutive jobs with the same multiplier use form a group (the set pipeline
  // holds the Jacobian's and the Hessian's jobs), each launched with the best kernel for it
  {
    const pk_job* ej = d->jobs[PK_STAGE_EXPAND];
    const long long nj = d->n_jobs[PK_STAGE_EXPAND];
    for (long long a = 0; a < nj;) {
      long long b = a + 1;
      while (b < nj && (ej[b].flags & PK_F_LAM) == (ej[a].flags & PK_F_LAM)) ++b;
      ms.exp.emplace_back();
      if (setup_expand_group(e, ej + a, a, b - a, ms.exp.back())) return 1;
      a = b;
    }
  }
  for (long long j = 0; j < d->n_jobs[PK_STAGE_DEFECT]; ++j) {
    const pk_job& jb = d->jobs[PK_STAGE_DEFECT][j];
    const long long r = jb.i[4] * jb.i[3];
    if (r > ms.max_defect_rows) ms.max_defect_rows = r;
    if (jb.i[13]) {
      ms.def_fast = true;
      if (jb.i[4] >= (1LL << 31) || jb.i[1] >= (1LL << 31)) return fail("defect job: phase too large for 32-bit row indices");
    } else {
      ms.def_table = true;
    }
    ms.def_geo.push_back({jb.i[13] != 0, jb.i[4], jb.i[3], (long long)(jb.flags > 0 ? jb.flags : 1)});
  }
  for (long long j = 0; j < d->n_jobs[PK_STAGE_REDUCE]; ++j) {
    const long long len = d->jobs[PK_STAGE_REDUCE][j].i[3] - d->jobs[PK_STAGE_REDUCE][j].i[2];
    if (len > ms.max_reduce_len) ms.max_reduce_len = len;
  }
  if (ms.max_reduce_len >= 2048) {
    ms.red_parts = (int)((ms.max_reduce_len + PK_REDUCE_SPAN - 1) / PK_REDUCE_SPAN);
    const size_t n = (size_t)d->n_jobs[PK_STAGE_REDUCE] * (size_t)e->dims.batch;
    CK(cudaMalloc((void**)&ms.red_partial, sizeof(double) * n * (size_t)ms.red_parts));
    CK(cudaMalloc((void**)&ms.red_ticket, sizeof(unsigned) * n));
    CK(cudaMemset(ms.red_ticket, 0, sizeof(unsigned) * n));
  }
  for (long long j = 0; j < d->n_jobs[PK_STAGE_GRAD_RANGE]; ++j)
    if (d->jobs[PK_STAGE_GRAD_RANGE][j].i[1] > ms.max_grad_count) ms.max_grad_count = d->jobs[PK_STAGE_GRAD_RANGE][j].i[1];
  {
    const long long lim = 1LL << 31, B = e->dims.batch;
    for (long long j = 0; j < d->n_jobs[PK_STAGE_GENERIC]; ++j)
      if (d->jobs[PK_STAGE_GENERIC][j].i[1] * B >= lim) ms.idx32 = false;
    if (ms.max_defect_rows * B >= lim || ms.max_grad_count * B >= lim) ms.idx32 = false;
  }
  ms.loaded = true;
  return 0;
}

extern "C" int pk_engine_get_cubin(pk_engine* e, int mode, const void** data, size_t* size) {
  if (!e || mode < 0 || mode >= PK_N_MODES) return fail("pk_engine_get_cubin: bad argument");
  *data = e->mode[mode].cubin.data();
  *size = e->mode[mode].cubin.size();
  return 0;
}

static PkCtx make_ctx(pk_engine* e, const ModeState& ms) {
  PkCtx cx;
  cx.X = e->X; cx.LAM = e->LAM; cx.SIG = e->SIG; cx.S = ms.S; cx.W = ms.W; cx.OUT = ms.OUT;
  cx.dpool = e->dpool; cx.ipool = e->ipool;
  cx.L = e->dims.L; cx.m = e->dims.m; cx.n_scalar = ms.n_scalar; cx.n_out = ms.n_out;
  // streaming result stores unless the compaction pass re-reads the values (POCKIT_B200_STREAM=0: plain stores)
  static const bool stream_env = !(getenv("POCKIT_B200_STREAM") && getenv("POCKIT_B200_STREAM")[0] == '0');
  cx.stream = (stream_env && !ms.n_compact) ? 1 : 0;
  return cx;
}

static unsigned use_pipeline(pk_engine* e, const int* modes, int n_modes);

static inline unsigned blocks_for(long long n, int per) { return (unsigned)((n + per - 1) / per); }

static void launch_expand(const ModeState& ms, const PkCtx& cx, int B, cudaStream_t st) {
  for (const ExpandGroup& g : ms.exp) {
    const pk_job* jobs = ms.jobs[PK_STAGE_EXPAND] + g.first;
    if (g.bulk) {
      const dim3 grid(g.xb_gx, (unsigned)g.xc.n_lists, B);
      if (g.lam)
        pk_expand_bulk<true><<<grid, PK_XB_THREADS, g.xb_smem, st>>>(cx, g.xc, g.xb_per_block);
      else
        pk_expand_bulk<false><<<grid, PK_XB_THREADS, g.xb_smem, st>>>(cx, g.xc, g.xb_per_block);
    } else if (g.batch && g.slots) {
      const dim3 grid(g.xsl_gx, g.n_groups);
      slots_kernel(g.lam, cx.stream != 0)<<<grid, PK_XC_THREADS, g.batch_smem <= 48 * 1024 ? g.batch_smem : 0, st>>>(cx, g.xc, (unsigned)B);
    } else if (g.batch) {
      const dim3 grid(g.xbt_gx, g.n_groups);
      batch_kernel(g.lam, g.batch_rows, cx.stream != 0)<<<grid, PK_XC_THREADS, g.batch_smem, st>>>(cx, g.xc, (unsigned)B);
    } else if (g.cols) {
      const dim3 grid(g.xc_gx, (unsigned)g.xc.n_lists, B);
      if (g.lam)
        pk_expand_cols<true><<<grid, PK_XC_THREADS, g.xc_smem, st>>>(cx, g.xc);
      else
        pk_expand_cols<false><<<grid, PK_XC_THREADS, g.xc_smem, st>>>(cx, g.xc);
    } else {
      pk_expand_blocks<<<dim3((unsigned)g.blocks, B), PK_THREADS, g.smem, st>>>(cx, jobs, (int)g.count, g.prefix, g.uniform);
    }
  }
}

// stage_mask selects which parts run (bit s = job stage s, bit PK_N_STAGES = node programs,
// bit PK_N_STAGES + 1 = system program); used by pk_time to attribute time.
static int launch_mode(pk_engine* e, int mode, unsigned stage_mask, cudaStream_t st) {
  ModeState& ms = e->mode[mode];
  if (!ms.loaded) return fail("mode not loaded");
  const int B = e->dims.batch;
  if (mode == PK_MODE_SET) {
    for (int k = 0; k < PK_N_CALLBACKS; ++k)
      if (ms.sub_cnt[k] > 0) e->mode[k].in_set = true;  // the callbacks this pipeline covers
  } else {
    ms.in_set = false;
  }
  PkCtx cx = make_ctx(e, ms);
  auto on = [&](int stage) { return (stage_mask & (1u << stage)) != 0; };
  if (on(PK_N_STAGES)) {
    tr(e, mode, 6, 0, st);
    // The per-node programs of the phases of a multi-phase system are independent (they write disjoint
    // table rows / scalar slots): phases 1, 2, ... run beside phase 0 on the auxiliary streams and are joined
    // before anything reads their results (two-stage rocket: 2 x 50 k nodes, each program latency-bound).
    const bool spread = stage_mask == ~0u && ms.node_kernels.size() > 1 && !e->trace;
    if (spread) {
      CK(cudaEventRecord(ms.node_fork, st));
      CK(cudaStreamWaitEvent(ms.aux1, ms.node_fork, 0));
      if (ms.node_kernels.size() > 2) CK(cudaStreamWaitEvent(ms.aux2, ms.node_fork, 0));
    }
    for (size_t p = 0; p < ms.node_kernels.size(); ++p) {
      const pk_node_program& np_ = ms.node_programs[p];
      const double* tm = e->dpool + np_.tm_offset;
      const double* wm = e->dpool + np_.wm_offset;
      int Bi = B;
      void* args[] = {&e->X, &e->LAM, &e->FIX, &tm, &wm, &ms.S, &ms.W, &ms.OUT, &Bi, &e->dpool, &e->ipool};
      const long long threads = (long long)B * np_.n_nodes;
      cudaStream_t sp = (!spread || p == 0) ? st : (p % 2 == 1 ? ms.aux1 : ms.aux2);
      CK(cudaLaunchKernel((void*)ms.node_kernels[p], dim3(blocks_for(threads, 128)), dim3(128), args, 0, sp));
      ++e->launches;
    }
    if (spread) {
      CK(cudaEventRecord(ms.node_join[0], ms.aux1));
      CK(cudaStreamWaitEvent(st, ms.node_join[0], 0));
      if (ms.node_kernels.size() > 2) {
        CK(cudaEventRecord(ms.node_join[1], ms.aux2));
        CK(cudaStreamWaitEvent(st, ms.node_join[1], 0));
      }
    }
    if (e->stagger) CK(cudaEventRecord(ms.node_done, st));
    tr(e, mode, 6, 1, st);
  }
  // Dependencies after the per-node programs (N):
  //   block expansion          <- N                      (only reads the node table)       stream st
  //   reductions -> system program (SYS) -> small slot runs   <- N                         stream ss
  //   defects                  <- N                                                        stream a1
  //   gradient: zero fill <- nothing; range gather <- SYS, zero fill                       stream a2
  //             scalar gather <- SYS                                                        stream a1
  // With an expansion, ss is the side stream and everything latency-bound hides behind it; a
  // whole-mode launch spreads the three small branches over ss / a1 / a2, a stage-masked one
  // (pk_time) keeps them on st in this order.
  const bool whole = stage_mask == ~0u;
  const bool run_exp = on(PK_STAGE_EXPAND) && !ms.exp.empty();
  const bool run_red = on(PK_STAGE_REDUCE) && ms.n_jobs[PK_STAGE_REDUCE];
  const bool run_sys = on(PK_N_STAGES + 1) && ms.sys_kernel;
  const bool run_def = on(PK_STAGE_DEFECT) && ms.n_jobs[PK_STAGE_DEFECT];
  const bool run_gen = on(PK_STAGE_GENERIC) && ms.gen_blocks;
  const bool run_grad = ms.grad_cnt && (on(PK_STAGE_GRAD_RANGE) || on(PK_STAGE_GRAD_SCALAR));
  const bool small = run_red || run_sys || run_def || run_gen || run_grad;
  const bool fork = run_exp && small;
  cudaStream_t ss = fork ? ms.side : st;
  const bool use_a1 = whole && (run_def || (run_grad && ms.n_jobs[PK_STAGE_GRAD_SCALAR])) && (run_red || run_sys || run_gen || fork);
  const bool use_a2 = whole && run_grad && (run_red || run_sys);
  cudaStream_t a1 = use_a1 ? ms.aux1 : ss, a2 = use_a2 ? ms.aux2 : ss;
  if (fork || use_a1 || use_a2) CK(cudaEventRecord(ms.side_fork, st));
  if (fork) CK(cudaStreamWaitEvent(ms.side, ms.side_fork, 0));
  if (use_a1) CK(cudaStreamWaitEvent(ms.aux1, ms.side_fork, 0));
  if (use_a2) CK(cudaStreamWaitEvent(ms.aux2, ms.side_fork, 0));
  if (run_exp) {
    tr(e, mode, PK_STAGE_EXPAND, 0, st);
    if (e->chain && e->chain_last >= 0) CK(cudaStreamWaitEvent(st, e->chain_ev[e->chain_last], 0));
    launch_expand(ms, cx, B, st);
    if (e->chain) { CK(cudaEventRecord(e->chain_ev[mode], st)); e->chain_last = mode; }
    e->launches += (long long)ms.exp.size();
    tr(e, mode, PK_STAGE_EXPAND, 1, st);
  }
  if (run_grad) {  // zero fill first: it depends on nothing
    CK(cudaMemset2DAsync(ms.OUT + ms.grad_off, sizeof(double) * (size_t)ms.n_out, 0, sizeof(double) * (size_t)ms.grad_cnt, (size_t)B, a2));
    if (a1 != a2 && ms.n_jobs[PK_STAGE_GRAD_SCALAR]) CK(cudaEventRecord(ms.aux2_join, a2));  // the scalar gather writes filled slots
  }
  if (run_def) {
    dim3 grid(blocks_for(ms.max_defect_rows * B, PK_THREADS), (unsigned)ms.n_jobs[PK_STAGE_DEFECT]);
    if (ms.def_table) {
      pk_defects<<<grid, PK_THREADS, 0, a1>>>(cx, ms.jobs[PK_STAGE_DEFECT], (int)ms.n_jobs[PK_STAGE_DEFECT], B);
      ++e->launches;
    }
    if (ms.def_fast) {
      // all phases' defect jobs in one launch (blockIdx.y), PK_DEF_JOBS at a time
      for (size_t j0 = 0; j0 < ms.def_geo.size(); j0 += PK_DEF_JOBS) {
        const size_t nj = ms.def_geo.size() - j0 < PK_DEF_JOBS ? ms.def_geo.size() - j0 : PK_DEF_JOBS;
        PkDefectDivs dvs;
        memset(&dvs, 0, sizeof(dvs));
        long long most = 0;
        bool any = false;
        for (size_t j = 0; j < nj; ++j) {
          const ModeState::DefectGeo& dg = ms.def_geo[j0 + j];
          if (!dg.fast) continue;
          any = true;
          dvs.d[j] = {div_multiplier((unsigned long long)dg.rows), div_multiplier((unsigned long long)dg.n_x),
                      div_multiplier((unsigned long long)dg.rb)};
          const long long total = dg.rows * dg.n_x * (long long)B;
          if (total > most) most = total;
        }
        if (!any) continue;
        const dim3 dgrid(blocks_for(most, PK_THREADS), (unsigned)nj);
        if (most < (1LL << 32))
          pk_defects_blocks<true><<<dgrid, PK_THREADS, 0, a1>>>(cx, ms.jobs[PK_STAGE_DEFECT], (int)j0, B, dvs);
        else
          pk_defects_blocks<false><<<dgrid, PK_THREADS, 0, a1>>>(cx, ms.jobs[PK_STAGE_DEFECT], (int)j0, B, dvs);
        ++e->launches;
      }
    }
    tr(e, mode, PK_STAGE_DEFECT, 1, a1);
  }
  if (run_red) {
    const long long warps = ms.n_jobs[PK_STAGE_REDUCE] * (long long)B;
    if (ms.red_parts)
      pk_reduce_rows_block<<<(unsigned)(warps * ms.red_parts), PK_THREADS, 0, ss>>>(cx, ms.jobs[PK_STAGE_REDUCE], (int)ms.n_jobs[PK_STAGE_REDUCE], B,
                                                                                   ms.red_parts, ms.red_partial, ms.red_ticket);
    else
      pk_reduce_rows<<<blocks_for(warps * 32, PK_THREADS), PK_THREADS, 0, ss>>>(cx, ms.jobs[PK_STAGE_REDUCE], (int)ms.n_jobs[PK_STAGE_REDUCE], B);
    ++e->launches;
    tr(e, mode, PK_STAGE_REDUCE, 1, ss);
  }
  if (run_sys) {
    int Bi = B;
    void* args[] = {&e->X, &ms.S, &ms.OUT, &Bi};
    CK(cudaLaunchKernel((void*)ms.sys_kernel, dim3(blocks_for(B, 64)), dim3(64), args, 0, ss));
    ++e->launches;
    tr(e, mode, 7, 1, ss);
  }
  if ((use_a1 || use_a2) && (run_red || run_sys)) {
    CK(cudaEventRecord(ms.sys_done, ss));
    if (use_a1) CK(cudaStreamWaitEvent(ms.aux1, ms.sys_done, 0));
    if (use_a2) CK(cudaStreamWaitEvent(ms.aux2, ms.sys_done, 0));
  }
  if (run_gen) {
    if (ms.idx32)
      pk_generic_jobs<unsigned><<<(unsigned)ms.gen_blocks, PK_THREADS, 0, ss>>>(cx, ms.jobs[PK_STAGE_GENERIC], ms.gen_job, ms.gen_chunk, B);
    else
      pk_generic_jobs<unsigned long long><<<(unsigned)ms.gen_blocks, PK_THREADS, 0, ss>>>(cx, ms.jobs[PK_STAGE_GENERIC], ms.gen_job, ms.gen_chunk, B);
    ++e->launches;
    tr(e, mode, PK_STAGE_GENERIC, 1, ss);
  }
  if (run_grad) {
    if (ms.n_jobs[PK_STAGE_GRAD_RANGE]) {
      dim3 grid(blocks_for(ms.max_grad_count * B, PK_CHUNK), (unsigned)ms.n_jobs[PK_STAGE_GRAD_RANGE]);
      if (ms.idx32)
        pk_grad_range<unsigned><<<grid, PK_THREADS, 0, a2>>>(cx, ms.jobs[PK_STAGE_GRAD_RANGE], B);
      else
        pk_grad_range<unsigned long long><<<grid, PK_THREADS, 0, a2>>>(cx, ms.jobs[PK_STAGE_GRAD_RANGE], B);
      ++e->launches;
    }
    if (ms.n_jobs[PK_STAGE_GRAD_SCALAR]) {
      if (a1 != a2) CK(cudaStreamWaitEvent(a1, ms.aux2_join, 0));  // behind the zero fill
      const long long n = ms.n_jobs[PK_STAGE_GRAD_SCALAR] * (long long)B;
      pk_grad_scalar<<<blocks_for(n, PK_THREADS), PK_THREADS, 0, a1>>>(cx, ms.jobs[PK_STAGE_GRAD_SCALAR], (int)ms.n_jobs[PK_STAGE_GRAD_SCALAR], B);
      ++e->launches;
    }
    tr(e, mode, PK_STAGE_GRAD_RANGE, 1, a2);
  }
  if (use_a1) {
    CK(cudaEventRecord(ms.aux1_join, ms.aux1));
    CK(cudaStreamWaitEvent(st, ms.aux1_join, 0));
  }
  if (use_a2) {
    CK(cudaEventRecord(ms.aux2_join, ms.aux2));
    CK(cudaStreamWaitEvent(st, ms.aux2_join, 0));
  }
  if (fork) {
    CK(cudaEventRecord(ms.side_join, ms.side));
    CK(cudaStreamWaitEvent(st, ms.side_join, 0));
  }
  if (ms.n_compact && stage_mask == ~0u) {
    pk_compact<<<blocks_for(ms.n_compact * B, PK_THREADS), PK_THREADS, 0, st>>>(ms.OUT, ms.OUTC, ms.cp_ptr, ms.cp_perm, ms.n_out,
                                                                               ms.n_compact, B);
    ++e->launches;
    tr(e, mode, 8, 1, st);
  }
  CK(cudaGetLastError());
  return 0;
}

// Make the engine stream wait (on the device) for the asynchronously started modes: all of them, or only
// those that read the multipliers.
static int drain(pk_engine* e, bool only_multiplier_readers = false) {
  std::vector<int> keep;
  for (int m : e->pending) {
    if (only_multiplier_readers && m != PK_MODE_HESSIAN && m != PK_MODE_SET) {
      keep.push_back(m);
      continue;
    }
    CK(cudaStreamWaitEvent(e->stream, e->mode[m].done, 0));
  }
  e->pending.swap(keep);
  return 0;
}

extern "C" int pk_upload_x(pk_engine* e, const double* x) {
  if (!e || !x) return fail("pk_upload_x: null argument");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;  // running modes still read the resident x
  const size_t n = sizeof(double) * (size_t)e->dims.batch * (size_t)e->dims.L;
  // x already in page-locked memory (pk_alloc_host, or a range registered with pk_host_register such as
  // the mapping the ranks of a sharded mesh share): copy straight from it.  The caller must leave it
  // alone until the evaluation that uses it returns (every pk_eval_* synchronises before returning).
  cudaPointerAttributes at;
  const bool locked = cudaPointerGetAttributes(&at, x) == cudaSuccess && at.type == cudaMemoryTypeHost;
  if (!locked) (void)cudaGetLastError();
  if (locked) {
    CK(cudaMemcpyAsync(e->X, x, n, cudaMemcpyHostToDevice, e->stream));
    CK(cudaEventRecord(e->x_done, e->stream));
    e->x_resident = true;
    ++e->x_uploads;
    return 0;
  }
  CK(cudaEventSynchronize(e->x_done));  // the previous copy may still be reading the staging buffer
  memcpy(e->hX, x, n);
  CK(cudaMemcpyAsync(e->X, e->hX, n, cudaMemcpyHostToDevice, e->stream));
  CK(cudaEventRecord(e->x_done, e->stream));
  e->x_resident = true;
  ++e->x_uploads;
  return 0;
}

extern "C" int pk_upload_multipliers(pk_engine* e, const double* lambda, const double* sigma) {
  if (!e) return fail("pk_upload_multipliers: null engine");
  CK(cudaSetDevice(e->device));
  if (drain(e, true)) return 1;
  const size_t B = (size_t)e->dims.batch;
  CK(cudaEventSynchronize(e->lam_done));
  if (lambda && e->dims.m > 0) {
    const size_t n = sizeof(double) * B * (size_t)e->dims.m;
    memcpy(e->hLAM, lambda, n);
    CK(cudaMemcpyAsync(e->LAM, e->hLAM, n, cudaMemcpyHostToDevice, e->stream));
  }
  if (sigma) {
    memcpy(e->hSIG, sigma, sizeof(double) * B);
    CK(cudaMemcpyAsync(e->SIG, e->hSIG, sizeof(double) * B, cudaMemcpyHostToDevice, e->stream));
  }
  CK(cudaEventRecord(e->lam_done, e->stream));
  return 0;
}

extern "C" int pk_run(pk_engine* e, int mode) {
  if (!e || mode < 0 || mode >= PK_N_MODES) return fail("pk_run: bad argument");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  return launch_mode(e, mode, ~0u, e->stream);
}

extern "C" int pk_sync(pk_engine* e) {
  if (!e) return fail("pk_sync: null engine");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// Enqueue the device-to-host copy of a mode's result on `st`.  `out` may be pinned (pk_alloc_host)
// or pageable; the runtime handles both.  De-duplicated pattern: the compacted values; mesh
// sharding: only the runs this engine computed, to the same offsets of `out`.
static int download_async(pk_engine* e, int mode, double* out, cudaStream_t st) {
  ModeState& ms = e->mode[mode];
  const size_t B = (size_t)e->dims.batch;
  if (mode < PK_N_CALLBACKS && ms.in_set) {  // the set pipeline produced it: a slice of the combined output
    const ModeState& set = e->mode[PK_MODE_SET];
    CK(cudaMemcpy2DAsync(out, sizeof(double) * (size_t)set.sub_cnt[mode], set.OUT + set.sub_off[mode], sizeof(double) * (size_t)set.n_out,
                         sizeof(double) * (size_t)set.sub_cnt[mode], B, cudaMemcpyDeviceToHost, st));
    return 0;
  }
  if (ms.n_compact) {
    CK(cudaMemcpyAsync(out, ms.OUTC, sizeof(double) * B * (size_t)ms.n_compact, cudaMemcpyDeviceToHost, st));
  } else if (!ms.dl_runs.empty()) {
    for (size_t b = 0; b < B; ++b)
      for (size_t r = 0; r + 1 < ms.dl_runs.size(); r += 2) {
        const size_t off = b * (size_t)ms.n_out + (size_t)ms.dl_runs[r];
        CK(cudaMemcpyAsync(out + off, ms.OUT + off, sizeof(double) * (size_t)ms.dl_runs[r + 1], cudaMemcpyDeviceToHost, st));
      }
  } else {
    CK(cudaMemcpyAsync(out, ms.OUT, sizeof(double) * B * (size_t)ms.n_out, cudaMemcpyDeviceToHost, st));
  }
  return 0;
}

extern "C" int pk_download(pk_engine* e, int mode, double* out) {
  if (!e || !out || mode < 0 || mode >= PK_N_MODES) return fail("pk_download: bad argument");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  if (download_async(e, mode, out, e->stream)) return 1;
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// Part of a mode's result: `count` values from slot `offset` of every instance, packed [B][count].
// hessian_o / hessian_c (systembase.py:735, 786) are the head / tail of the Hessian values: the engine
// evaluates the mode once and only the requested part crosses PCIe.
extern "C" int pk_download_range(pk_engine* e, int mode, int64_t offset, int64_t count, double* out) {
  if (!e || !out || mode < 0 || mode >= PK_N_MODES) return fail("pk_download_range: bad argument");
  ModeState& ms = e->mode[mode];
  if (!ms.loaded) return fail("pk_download_range: mode not loaded");
  if (ms.n_compact || !ms.dl_runs.empty() || ms.in_set) return fail("pk_download_range: not available with shaped outputs");
  if (offset < 0 || count < 0 || offset + count > ms.n_out) return fail("pk_download_range: range outside the output");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  if (count)
    CK(cudaMemcpy2DAsync(out, sizeof(double) * (size_t)count, ms.OUT + offset, sizeof(double) * (size_t)ms.n_out,
                         sizeof(double) * (size_t)count, (size_t)e->dims.batch, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// Device address of a mode's latest result: `count` values per instance, instances `stride` doubles
// apart (stride > count when the values are a slice of the set pipeline's combined output).  Lets a
// caller hand the values to a collective (NCCL all-gather of sharded batches) without a host round trip.
extern "C" int pk_out_device_pointer(pk_engine* e, int mode, void** ptr, int64_t* count, int64_t* stride) {
  if (!e || !ptr || mode < 0 || mode >= PK_N_MODES) return fail("pk_out_device_pointer: bad argument");
  ModeState& ms = e->mode[mode];
  if (!ms.loaded) return fail("pk_out_device_pointer: mode not loaded");
  if (mode < PK_N_CALLBACKS && ms.in_set) {  // a slice of the set pipeline's combined output
    const ModeState& set = e->mode[PK_MODE_SET];
    *ptr = (void*)(set.OUT + set.sub_off[mode]);
    if (count) *count = set.sub_cnt[mode];
    if (stride) *stride = set.n_out;
    return 0;
  }
  *ptr = ms.n_compact ? (void*)ms.OUTC : (void*)ms.OUT;
  const int64_t n = ms.n_compact ? ms.n_compact : ms.n_out;
  if (count) *count = n;
  if (stride) *stride = n;
  return 0;
}

extern "C" int pk_engine_set_output_runs(pk_engine* e, int mode, const int64_t* runs, int64_t n_runs) {
  if (!e || mode < 0 || mode >= PK_N_MODES || n_runs < 0 || (n_runs && !runs)) return fail("pk_engine_set_output_runs: bad argument");
  ModeState& ms = e->mode[mode];
  if (!ms.loaded) return fail("pk_engine_set_output_runs: mode not loaded");
  if (ms.n_compact && n_runs) return fail("pk_engine_set_output_runs: not available with a de-duplicated pattern");
  for (int64_t r = 0; r < n_runs; ++r)
    if (runs[2 * r] < 0 || runs[2 * r + 1] < 0 || runs[2 * r] + runs[2 * r + 1] > ms.n_out)
      return fail("pk_engine_set_output_runs: run outside the output");
  ms.dl_runs.assign(runs, runs + 2 * n_runs);
  ms.in_set = false;
  return 0;
}

extern "C" int pk_engine_set_compaction(pk_engine* e, int mode, int64_t n_unique, const int64_t* seg_ptr, const int64_t* perm) {
  if (!e || mode < 0 || mode >= PK_N_MODES || n_unique < 0) return fail("pk_engine_set_compaction: bad argument");
  CK(cudaSetDevice(e->device));
  ModeState& ms = e->mode[mode];
  if (!ms.loaded) return fail("pk_engine_set_compaction: mode not loaded");
  if (!ms.dl_runs.empty()) return fail("pk_engine_set_compaction: not available on a mesh shard");
  if (ms.cp_ptr) cudaFree(ms.cp_ptr);
  if (ms.cp_perm) cudaFree(ms.cp_perm);
  if (ms.OUTC) cudaFree(ms.OUTC);
  ms.cp_ptr = ms.cp_perm = nullptr;
  ms.OUTC = nullptr;
  ms.n_compact = 0;
  ms.in_set = false;
  if (e->set_graph) {  // the captured set no longer matches
    cudaGraphExecDestroy(e->set_graph);
    e->set_graph = nullptr;
    e->set_modes.clear();
  }
  if (n_unique == 0) return 0;
  if (!seg_ptr || !perm) return fail("pk_engine_set_compaction: null table");
  if (ms.n_out >= (1LL << 32)) return fail("pk_engine_set_compaction: pattern too large for 32-bit slot indices");
  if (seg_ptr[0] != 0 || seg_ptr[n_unique] != ms.n_out) return fail("pk_engine_set_compaction: segments must cover every slot once");
  std::vector<unsigned> p32((size_t)n_unique + 1), q32((size_t)ms.n_out);
  for (int64_t u = 0; u <= n_unique; ++u) {
    if (u && seg_ptr[u] < seg_ptr[u - 1]) return fail("pk_engine_set_compaction: segment table not monotone");
    p32[(size_t)u] = (unsigned)seg_ptr[u];
  }
  for (int64_t k = 0; k < ms.n_out; ++k) {
    if (perm[k] < 0 || perm[k] >= ms.n_out) return fail("pk_engine_set_compaction: slot index out of range");
    q32[(size_t)k] = (unsigned)perm[k];
  }
  CK(cudaMalloc((void**)&ms.cp_ptr, sizeof(unsigned) * p32.size()));
  CK(cudaMalloc((void**)&ms.cp_perm, sizeof(unsigned) * (q32.size() ? q32.size() : 1)));
  CK(cudaMalloc((void**)&ms.OUTC, sizeof(double) * (size_t)e->dims.batch * (size_t)n_unique));
  CK(cudaMemcpy(ms.cp_ptr, p32.data(), sizeof(unsigned) * p32.size(), cudaMemcpyHostToDevice));
  if (!q32.empty()) CK(cudaMemcpy(ms.cp_perm, q32.data(), sizeof(unsigned) * q32.size(), cudaMemcpyHostToDevice));
  ms.n_compact = n_unique;
  return 0;
}

extern "C" int pk_out_size(pk_engine* e, int mode, int64_t* n) {
  if (!e || !n || mode < 0 || mode >= PK_N_MODES) return fail("pk_out_size: bad argument");
  const ModeState& ms = e->mode[mode];
  *n = ms.n_compact ? ms.n_compact : ms.n_out;
  return 0;
}

static int eval(pk_engine* e, int mode, const double* x, const double* lam, const double* sig, double* out) {
  if (pk_upload_x(e, x)) return 1;
  if (mode == PK_MODE_HESSIAN && pk_upload_multipliers(e, lam, sig)) return 1;
  if (launch_mode(e, mode, ~0u, e->stream)) return 1;
  return pk_download(e, mode, out);
}

extern "C" int pk_eval_objective(pk_engine* e, const double* x, double* f) { return eval(e, PK_MODE_OBJECTIVE, x, nullptr, nullptr, f); }
extern "C" int pk_eval_gradient(pk_engine* e, const double* x, double* g) { return eval(e, PK_MODE_GRADIENT, x, nullptr, nullptr, g); }
extern "C" int pk_eval_constraints(pk_engine* e, const double* x, double* c) { return eval(e, PK_MODE_CONSTRAINTS, x, nullptr, nullptr, c); }
extern "C" int pk_eval_jacobian(pk_engine* e, const double* x, double* v) { return eval(e, PK_MODE_JACOBIAN, x, nullptr, nullptr, v); }
extern "C" int pk_eval_hessian(pk_engine* e, const double* x, const double* lam, const double* sig, double* v) {
  if (!lam || !sig) return fail("pk_eval_hessian: multipliers required");
  return eval(e, PK_MODE_HESSIAN, x, lam, sig, v);
}

// Host-to-host evaluation of several callbacks at one x: x (and the multipliers) cross PCIe once,
// every mode runs on its own stream and its device-to-host copy is queued right behind its kernels,
// so the copies of the large Jacobian / Hessian value arrays overlap the other modes' compute and
// the host blocks once.  This is the entry point of an x-keyed evaluation cache in a solver adapter
// (Ipopt asks for f, grad f, g, J and H at the same x; ipopt.py:41-53).
static int eval_set(pk_engine* e, const double* x, const double* lam, const double* sig, const int* modes, int n_modes,
                    double* const* outs, bool wait);

extern "C" int pk_eval_set(pk_engine* e, const double* x, const double* lam, const double* sig, const int* modes,
                           int n_modes, double* const* outs) {
  return eval_set(e, x, lam, sig, modes, n_modes, outs, true);
}

// Same, but returns as soon as everything is enqueued: the results are complete after pk_sync.  A caller
// that receives its inputs in stages (a mesh-shard worker gets x first and the multipliers a moment later)
// starts the Jacobian and its copy while the rest is still on its way.  Host buffers passed as x / lambda
// must stay untouched until pk_sync when they are page-locked (they are read by the copy engine directly).
extern "C" int pk_eval_set_async(pk_engine* e, const double* x, const double* lam, const double* sig, const int* modes,
                                 int n_modes, double* const* outs) {
  return eval_set(e, x, lam, sig, modes, n_modes, outs, false);
}

static int eval_set(pk_engine* e, const double* x, const double* lam, const double* sig, const int* modes, int n_modes,
                    double* const* outs, bool wait) {
  if (!e || !modes || !outs || n_modes < 1) return fail("pk_eval_set: bad argument");
  if (!x && !e->x_resident) return fail("pk_eval_set: x = NULL reuses the resident point, but none was uploaded yet");
  CK(cudaSetDevice(e->device));
  bool hess = false;
  for (int k = 0; k < n_modes; ++k) {
    if (modes[k] < 0 || modes[k] >= PK_N_MODES || !e->mode[modes[k]].loaded) return fail("pk_eval_set: mode not loaded");
    if (!outs[k]) return fail("pk_eval_set: null output");
    for (int q = 0; q < k; ++q)
      if (modes[q] == modes[k]) return fail("pk_eval_set: a mode may appear only once");
    hess = hess || modes[k] == PK_MODE_HESSIAN;
  }
  if (hess && (!lam || !sig)) return fail("pk_eval_set: multipliers required for the Hessian");
  if (wait && drain(e)) return 1;
  for (int k = 0; k < n_modes; ++k)  // a mode started again while pending: its own stream keeps the order
    for (size_t q = 0; q < e->pending.size(); ++q)
      if (e->pending[q] == modes[k]) { e->pending.erase(e->pending.begin() + q); break; }
  if (x && pk_upload_x(e, x)) return 1;
  unsigned asked = 0;
  for (int k = 0; k < n_modes; ++k) asked |= 1u << modes[k];
  if (use_pipeline(e, modes, n_modes) == asked) {
    if (hess && pk_upload_multipliers(e, lam, sig)) return 1;
    // one pipeline for everything that was asked for; the copies of its slices follow on the same stream, largest first
    CK(cudaEventRecord(e->fork, e->stream));
    ModeState& ps = e->mode[PK_MODE_SET];
    CK(cudaStreamWaitEvent(ps.stream, e->fork, 0));
    if (launch_mode(e, PK_MODE_SET, ~0u, ps.stream)) return 1;
    std::vector<int> ord(n_modes);
    for (int k = 0; k < n_modes; ++k) ord[k] = k;
    for (int a = 1; a < n_modes; ++a)
      for (int b = a; b > 0 && ps.sub_cnt[modes[ord[b]]] > ps.sub_cnt[modes[ord[b - 1]]]; --b) std::swap(ord[b], ord[b - 1]);
    for (int q = 0; q < n_modes; ++q)
      if (download_async(e, modes[ord[q]], outs[ord[q]], ps.stream)) return 1;
    CK(cudaEventRecord(ps.done, ps.stream));
    if (!wait) {
      e->pending.push_back(PK_MODE_SET);
      return 0;
    }
    CK(cudaStreamWaitEvent(e->stream, ps.done, 0));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
  }
  CK(cudaEventRecord(e->fork, e->stream));
  // largest outputs first: their copies keep the copy engine busy while the small modes compute
  std::vector<int> order(n_modes);
  for (int k = 0; k < n_modes; ++k) order[k] = k;
  for (int a = 1; a < n_modes; ++a)
    for (int b = a; b > 0 && e->mode[modes[order[b]]].n_out > e->mode[modes[order[b - 1]]].n_out; --b) {
      const int t = order[b]; order[b] = order[b - 1]; order[b - 1] = t;
    }
  // Two passes: everything that does not read the multipliers starts right behind the upload of x -- the
  // Jacobian is expanded and its copy is on the link while the host still stages the multipliers; the
  // Hessian follows their upload.  The engine stream is joined to the modes only at the very end, so the
  // upload in between does not wait for the first pass (device-to-host copies included).
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      if (hess && pk_upload_multipliers(e, lam, sig)) return 1;
      CK(cudaEventRecord(e->fork, e->stream));
    }
    for (int q = 0; q < n_modes; ++q) {
      const int k = order[q];
      if ((modes[k] == PK_MODE_HESSIAN || modes[k] == PK_MODE_SET) != (pass == 1)) continue;
      ModeState& ms = e->mode[modes[k]];
      CK(cudaStreamWaitEvent(ms.stream, e->fork, 0));
      if (launch_mode(e, modes[k], ~0u, ms.stream)) return 1;
      if (download_async(e, modes[k], outs[k], ms.stream)) return 1;
      CK(cudaEventRecord(ms.done, ms.stream));
    }
  }
  for (int k = 0; k < n_modes; ++k) {
    if (wait)
      CK(cudaStreamWaitEvent(e->stream, e->mode[modes[k]].done, 0));
    else
      e->pending.push_back(modes[k]);  // the engine stream stays free for the next stage's uploads
  }
  if (wait) CK(cudaStreamSynchronize(e->stream));
  return 0;
}

extern "C" int pk_time(pk_engine* e, int mode, int iters, float* ms_total, float* ms_stage) {
  if (!e || mode < 0 || mode >= PK_N_MODES || iters < 1) return fail("pk_time: bad argument");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  ScopedEvent a, b;
  CK(a.create());
  CK(b.create());
  auto timed = [&](unsigned mask, float* out) -> int {
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventRecord(a, e->stream));
    for (int i = 0; i < iters; ++i)
      if (launch_mode(e, mode, mask, e->stream)) return 1;
    CK(cudaEventRecord(b, e->stream));
    CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(out, a, b));
    return 0;
  };
  if (ms_total && timed(~0u, ms_total)) return 1;
  if (ms_stage)
    for (int s = 0; s < PK_N_STAGES + 2; ++s) {
      unsigned mask = 1u << s;
      if (s == PK_STAGE_GRAD_RANGE) mask |= 1u << PK_STAGE_GRAD_SCALAR;
      if (s == PK_STAGE_GRAD_SCALAR) { ms_stage[s] = 0.f; continue; }
      if (timed(mask, &ms_stage[s])) return 1;
    }
  return 0;
}

// One stage of a mode timed launch by launch with the L2 flushed (untimed) before every launch: the
// roofline figure of the dominant kernel is taken under the same cache conditions as the whole-set
// throughput.  stage_mask as in launch_mode; ms_each[i] = CUDA-event time of launch i on the engine stream.
extern "C" int pk_time_stage(pk_engine* e, int mode, unsigned stage_mask, int iters, int flush_l2, float* ms_each) {
  if (!e || mode < 0 || mode >= PK_N_MODES || iters < 1 || !ms_each) return fail("pk_time_stage: bad argument");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  ScopedEvent a, b;
  CK(a.create());
  CK(b.create());
  for (int i = 0; i < iters; ++i) {
    if (flush_l2 && pk_flush_l2(e)) return 1;
    CK(cudaEventRecord(a, e->stream));
    if (launch_mode(e, mode, stage_mask, e->stream)) return 1;
    CK(cudaEventRecord(b, e->stream));
    CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(&ms_each[i], a, b));
  }
  return 0;
}

// The same stage of several modes launched in turn, `rounds` times, back to back on the engine stream
// (one pair of CUDA events around everything): the Jacobian and the Hessian expansion alternate exactly as
// they follow each other inside a set, and their outputs together (robot_arm: 207 MB) exceed the 126 MB
// L2, so every byte has to drain to HBM -- the sustained-streaming figure of the dominant kernel.
extern "C" int pk_time_stage_alternating(pk_engine* e, const int* modes, int n_modes, unsigned stage_mask, int rounds, float* ms_total) {
  if (!e || !modes || n_modes < 1 || rounds < 1 || !ms_total) return fail("pk_time_stage_alternating: bad argument");
  for (int k = 0; k < n_modes; ++k)
    if (modes[k] < 0 || modes[k] >= PK_N_MODES || !e->mode[modes[k]].loaded) return fail("pk_time_stage_alternating: mode not loaded");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  ScopedEvent a, b;
  CK(a.create());
  CK(b.create());
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaEventRecord(a, e->stream));
  for (int r = 0; r < rounds; ++r)
    for (int k = 0; k < n_modes; ++k)
      if (launch_mode(e, modes[k], stage_mask, e->stream)) return 1;
  CK(cudaEventRecord(b, e->stream));
  CK(cudaEventSynchronize(b));
  CK(cudaEventElapsedTime(ms_total, a, b));
  return 0;
}

// Launch several callbacks at the same x as ONE graph: every mode runs on its own stream (its
// tables and output buffer are private), forked from and joined to the engine stream, so the
// latency-bound small callbacks overlap the HBM-bound expansions and the host pays one launch.
// The callbacks the loaded set pipeline covers (all five, or the three small ones) are evaluated as
// PK_MODE_SET when all of them are requested and no per-callback output shaping (de-duplicated
// pattern / mesh shard) is in the way; use_pipeline returns their bit mask (0: no pipeline).
static unsigned use_pipeline(pk_engine* e, const int* modes, int n_modes) {
  const ModeState& set = e->mode[PK_MODE_SET];
  if (!set.loaded) return 0;
  const char* env = getenv("POCKIT_B200_SET");
  if (env && env[0] == '0') return 0;
  unsigned covered = 0, seen = 0;
  for (int k = 0; k < PK_N_CALLBACKS; ++k)
    if (set.sub_cnt[k] > 0) covered |= 1u << k;
  for (int k = 0; k < n_modes; ++k) {
    if (modes[k] < 0 || modes[k] >= PK_N_CALLBACKS) return 0;
    seen |= 1u << modes[k];
  }
  if (!covered || (seen & covered) != covered) return 0;  // every covered callback must be asked for
  for (int k = 0; k < PK_N_CALLBACKS; ++k)
    if ((covered >> k & 1u) && (e->mode[k].n_compact || !e->mode[k].dl_runs.empty())) return 0;
  return covered;
}

static int run_set(pk_engine* e, const int* modes, int n_modes) {
  std::vector<int> want;
  const unsigned covered = use_pipeline(e, modes, n_modes);
  const bool pipeline = covered != 0;
  for (int k = 0; k < n_modes; ++k)
    if (!(covered >> modes[k] & 1u)) want.push_back(modes[k]);
  if (pipeline) want.push_back(PK_MODE_SET);
  n_modes = (int)want.size();
  modes = want.data();
  if (!e->set_graph || want != e->set_modes) {
    for (int k = 0; k < n_modes; ++k)
      if (modes[k] < 0 || modes[k] >= PK_N_MODES || !e->mode[modes[k]].loaded) return fail("pk_run_set: mode not loaded");
    if (e->set_graph) cudaGraphExecDestroy(e->set_graph);
    e->set_graph = nullptr;
    cudaGraph_t graph = nullptr;
    const long long before = e->launches;
    CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    CK(cudaEventRecord(e->fork, e->stream));
    int rc = 0;
    // Schedule of a set.  Opt-in alternative (POCKIT_B200_STAGGER=1|2): the critical path -- node
    // program -> expansion of the modes that have one (Jacobian, Hessian), expansions chained -- gets
    // the machine first:
    //   1. the per-node program of the first expansion mode runs alone, its expansion starts right behind;
    //   2. the next expansion mode's per-node program is released when the previous one has finished
    //      (it overlaps the running, HBM-bound expansion) -- its expansion follows the chain;
    //   3. the small callbacks (objective, gradient, constraints) are released behind the last of those
    //      per-node programs (POCKIT_B200_STAGGER=1: behind the first) and hide under the expansions.
    // POCKIT_B200_STAGGER=0 (default): everything is released at once, largest first.
    // Measured on B200 (round 2, profiles/r02_call3_stagger_ab.log) the staggered schedules are SLOWER
    // -- robot_arm 60.1 -> 59.0 (=1) / 64.3 us (=2), humanoid 143.1 -> 145.6 / 156.6, rocket 59.9 -> 67.9 /
    // 73.3 -- the kernels released together do overlap usefully; holding the small callbacks back
    // only moves them to the tail.  Kept as an opt-in switch.
    const char* sg = getenv("POCKIT_B200_STAGGER");
    const int stagger = sg ? atoi(sg) : 0;
    std::vector<int> order;
    if (stagger > 0) {
      std::vector<int> big, small;
      for (int m : want) (e->mode[m].exp.empty() ? small : big).push_back(m);
      for (size_t a = 1; a < big.size(); ++a)  // lighter mode first: its expansion heads the chain
        for (size_t b = a; b > 0 && e->mode[big[b]].n_out < e->mode[big[b - 1]].n_out; --b) std::swap(big[b], big[b - 1]);
      for (size_t a = 1; a < small.size(); ++a)
        for (size_t b = a; b > 0 && e->mode[small[b]].n_out > e->mode[small[b - 1]].n_out; --b) std::swap(small[b], small[b - 1]);
      order = big;
      order.insert(order.end(), small.begin(), small.end());
    } else {
      order = want;  // largest outputs first: their expansions head the chain
      for (int a = 1; a < n_modes; ++a)
        for (int b = a; b > 0 && e->mode[order[b]].n_out > e->mode[order[b - 1]].n_out; --b) std::swap(order[b], order[b - 1]);
    }
    const char* ch = getenv("POCKIT_B200_CHAIN");
    e->chain = !(ch && ch[0] == '0');
    e->chain_last = -1;
    e->stagger = stagger > 0;
    cudaEvent_t gate = nullptr, first_gate = nullptr;
    for (int k = 0; k < n_modes && !rc; ++k) {
      ModeState& ms = e->mode[order[k]];
      const bool big = !ms.exp.empty();
      if (cudaStreamWaitEvent(ms.stream, e->fork, 0) != cudaSuccess) rc = 1;
      cudaEvent_t wait_for = (stagger == 1 && !big) ? first_gate : gate;
      if (!rc && stagger > 0 && wait_for && cudaStreamWaitEvent(ms.stream, wait_for, 0) != cudaSuccess) rc = 1;
      if (!rc) rc = launch_mode(e, order[k], ~0u, ms.stream);
      if (big && !ms.node_kernels.empty()) {
        gate = ms.node_done;
        if (!first_gate) first_gate = gate;
      }
      if (!rc && cudaEventRecord(ms.done, ms.stream) != cudaSuccess) rc = 1;
      if (!rc && cudaStreamWaitEvent(e->stream, ms.done, 0) != cudaSuccess) rc = 1;
    }
    e->chain = false;
    e->stagger = false;
    cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    if (rc || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      return rc ? 1 : fail("pk_run_set: stream capture failed");
    }
    e->set_launches = e->launches - before;
    e->launches = before;
    // Priorities inside the graph: every kernel except the block expansions is latency-bound and
    // small; give those nodes the highest priority so that their few blocks are dispatched as soon as
    // an expansion block retires instead of queueing behind the thousands still pending.
    {
      const char* pe = getenv("POCKIT_B200_GRAPH_PRIORITY");
      int prio_lo = 0, prio_hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      if (pe && pe[0] == '1' && prio_hi != prio_lo) {
        size_t n_nodes = 0;
        CK(cudaGraphGetNodes(graph, nullptr, &n_nodes));
        std::vector<cudaGraphNode_t> nodes(n_nodes);
        if (n_nodes) CK(cudaGraphGetNodes(graph, nodes.data(), &n_nodes));
        const void* big[] = {(const void*)pk_expand_cols<true>, (const void*)pk_expand_cols<false>, (const void*)pk_expand_blocks,
                             (const void*)slots_kernel(true, true), (const void*)slots_kernel(true, false),
                             (const void*)slots_kernel(false, true), (const void*)slots_kernel(false, false)};
        for (cudaGraphNode_t nd : nodes) {
          cudaGraphNodeType ty;
          if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
          cudaKernelNodeParams kp;
          bool is_big = false;
          if (cudaGraphKernelNodeGetParams(nd, &kp) == cudaSuccess) {
            for (const void* f : big) is_big = is_big || kp.func == f;
            for (int rr = 1; rr <= PK_XM_ROWS && !is_big; ++rr)
              for (int v4 = 0; v4 < 4; ++v4) is_big = is_big || kp.func == (const void*)batch_kernel((v4 & 1) != 0, rr, (v4 & 2) != 0);
          } else {
            (void)cudaGetLastError();  // library (NVRTC) kernels: not a host function pointer -- small by construction
          }
          cudaLaunchAttributeValue v;
          memset(&v, 0, sizeof(v));
          v.priority = is_big ? prio_lo : prio_hi;
          if (cudaGraphKernelNodeSetAttribute(nd, cudaLaunchAttributePriority, &v) != cudaSuccess) (void)cudaGetLastError();
        }
      }
    }
    {
      // the per-node priorities set above only take effect when the executable graph is instantiated with
      // cudaGraphInstantiateFlagUseNodePriority (round 1 instantiated with flags = 0 and measured "no effect")
      const char* pe = getenv("POCKIT_B200_GRAPH_PRIORITY");
      // Opt-in (POCKIT_B200_GRAPH_PRIORITY=1): measured again WITH the flag in round 2
      // (profiles/r02_call17_graph_priority.log) -- humanoid 122.2 vs 121.5 us, rocket 55.4 vs 55.5, robot_arm
      // 61.5 vs 58.1, quadrotor batch 271 vs 264: the kernels of a set share bandwidth and issue slots, not a
      // dispatch queue, so priorities do not help.
      const unsigned long long flags = (pe && pe[0] == '1') ? (unsigned long long)cudaGraphInstantiateFlagUseNodePriority : 0ull;
      CK(cudaGraphInstantiateWithFlags(&e->set_graph, graph, flags));
    }
    cudaGraphDestroy(graph);
    e->set_modes = want;
  }
  CK(cudaGraphLaunch(e->set_graph, e->stream));
  e->launches += e->set_launches;
  // replays do not pass through launch_mode: keep track of where the latest values live
  for (int k = 0; k < n_modes; ++k)
    if (want[k] < PK_N_CALLBACKS) e->mode[want[k]].in_set = false;
  for (int k = 0; k < PK_N_CALLBACKS; ++k)
    if (covered >> k & 1u) e->mode[k].in_set = true;
  return 0;
}

extern "C" int pk_run_set(pk_engine* e, const int* modes, int n_modes) {
  if (!e || !modes || n_modes < 1) return fail("pk_run_set: bad argument");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  return run_set(e, modes, n_modes);
}

extern "C" int pk_time_steps(pk_engine* e, const int* modes, int n_modes, int steps, int flush_l2, float* ms_steps) {
  if (!e || !modes || n_modes < 1 || steps < 1 || !ms_steps) return fail("pk_time_steps: bad argument");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  ScopedEvent a, b;
  CK(a.create());
  CK(b.create());
  for (int s = 0; s < steps; ++s) {
    if (flush_l2 && pk_flush_l2(e)) return 1;
    CK(cudaEventRecord(a, e->stream));
    if (run_set(e, modes, n_modes)) return 1;
    CK(cudaEventRecord(b, e->stream));
    CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(&ms_steps[s], a, b));
  }
  return 0;
}

// Device-side timeline of one evaluation set launched on the per-mode streams (no graph): the
// engine stream is first kept busy with a few L2-flush fills so that the host gets ahead and the
// marks show dependency resolution on the device, not host launch latency.  rows = (mode, tag,
// edge, microseconds since the set was released); tags as in tr().
extern "C" int pk_timeline(pk_engine* e, const int* modes, int n_modes, double* rows, int max_rows, int* n_rows) {
  if (!e || !modes || n_modes < 1 || !rows || !n_rows) return fail("pk_timeline: bad argument");
  CK(cudaSetDevice(e->device));
  if (drain(e)) return 1;
  for (int k = 0; k < n_modes; ++k)
    if (modes[k] < 0 || modes[k] >= PK_N_MODES || !e->mode[modes[k]].loaded) return fail("pk_timeline: mode not loaded");
  struct Marks {  // the trace events are released on every return path
    std::vector<pk_engine::Mark> v;
    ~Marks() {
      for (auto& mk : v)
        if (mk.ev) cudaEventDestroy(mk.ev);
    }
  } owned;
  std::vector<pk_engine::Mark>& marks = owned.v;
  ScopedEvent base, end;
  CK(base.create());
  CK(end.create());
  CK(cudaStreamSynchronize(e->stream));
  for (int k = 0; k < 4; ++k)
    if (pk_flush_l2(e)) return 1;
  CK(cudaEventRecord(base, e->stream));
  CK(cudaEventRecord(e->fork, e->stream));
  e->trace = &marks;
  int rc = 0;
  for (int k = 0; k < n_modes && !rc; ++k) {
    ModeState& ms = e->mode[modes[k]];
    if (cudaStreamWaitEvent(ms.stream, e->fork, 0) != cudaSuccess) rc = 1;
    if (!rc) rc = launch_mode(e, modes[k], ~0u, ms.stream);
    if (!rc && cudaEventRecord(ms.done, ms.stream) != cudaSuccess) rc = 1;
    if (!rc && cudaStreamWaitEvent(e->stream, ms.done, 0) != cudaSuccess) rc = 1;
  }
  e->trace = nullptr;
  if (rc) return rc;
  CK(cudaEventRecord(end, e->stream));
  CK(cudaEventSynchronize(end));
  int n = 0;
  for (auto& mk : marks) {
    float ms_ = 0.f;
    cudaEventElapsedTime(&ms_, base, mk.ev);
    if (n < max_rows) {
      rows[4 * n] = mk.mode; rows[4 * n + 1] = mk.tag; rows[4 * n + 2] = mk.edge; rows[4 * n + 3] = 1000.0 * ms_;
      ++n;
    }
  }
  float tot = 0.f;
  cudaEventElapsedTime(&tot, base, end);
  if (n < max_rows) {
    rows[4 * n] = -1; rows[4 * n + 1] = -1; rows[4 * n + 2] = 1; rows[4 * n + 3] = 1000.0 * tot;
    ++n;
  }
  *n_rows = n;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// continuous error-estimate data (phasebase.py:1339-1366)
template <typename T>
static int to_device(T** dst, const T* src, size_t n) {
  CK(cudaMalloc((void**)dst, sizeof(T) * (n ? n : 1)));
  if (n) CK(cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int pk_engine_load_error_estimate(pk_engine* e, const char* source, const char* const* opts, int n_opts,
                                             const pk_aug_phase* phases, int n_phases) {
  if (!e || !source || !phases || n_phases < 1) return fail("pk_engine_load_error_estimate: bad argument");
  if (e->dims.batch != 1) return fail("pk_engine_load_error_estimate: engines with batch == 1 only");
  CK(cudaSetDevice(e->device));
  free_aug(e);
  if (compile_source(source, opts, n_opts, e->device, e->aug_cubin)) return 1;
  CK(cudaLibraryLoadData(&e->aug_lib, e->aug_cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  e->aug.resize(n_phases);
  for (int i = 0; i < n_phases; ++i) {
    const pk_aug_phase& p = phases[i];
    pk_engine::AugPhase& a = e->aug[i];
    if (p.x_offset < 0 || p.x_offset + p.L > e->dims.L || p.L_xu > p.L || p.L_x_all > p.L_xu)
      return fail("pk_engine_load_error_estimate: phase layout outside x");
    CK(cudaLibraryGetKernel(&a.prep, e->aug_lib, p.prep_kernel));
    CK(cudaLibraryGetKernel(&a.node, e->aug_lib, p.node_kernel));
    a.x_offset = p.x_offset; a.L = p.L; a.L_xu = p.L_xu; a.L_x_all = p.L_x_all;
    a.n_x = p.n_x; a.n_u = p.n_u; a.Lm_aug = p.Lm_aug; a.rows = p.rows;
    const size_t nv = (size_t)((p.n_x + p.n_u) * p.Lm_aug), nt = (size_t)(p.n_x * p.rows), ni = (size_t)p.rows;
    if (to_device(&a.tm, p.tm_aug, (size_t)p.Lm_aug)) return 1;
    if (to_device(&a.Vp, (const long long*)p.V_ptr, nv + 1) || to_device(&a.Vi, (const long long*)p.V_idx, (size_t)p.V_ptr[nv]) ||
        to_device(&a.Vv, p.V_val, (size_t)p.V_ptr[nv])) return 1;
    if (to_device(&a.Tp, (const long long*)p.T_ptr, nt + 1) || to_device(&a.Ti, (const long long*)p.T_idx, (size_t)p.T_ptr[nt]) ||
        to_device(&a.Tv, p.T_val, (size_t)p.T_ptr[nt])) return 1;
    if (to_device(&a.Ip, (const long long*)p.I_ptr, ni + 1) || to_device(&a.Ii, (const long long*)p.I_idx, (size_t)p.I_ptr[ni]) ||
        to_device(&a.Iv, p.I_val, (size_t)p.I_ptr[ni])) return 1;
    CK(cudaMalloc((void**)&a.XS, sizeof(double) * (size_t)(p.L > 0 ? p.L : 1)));
    CK(cudaMalloc((void**)&a.XU, sizeof(double) * (nv ? nv : 1)));
    CK(cudaMalloc((void**)&a.WA, sizeof(double) * (size_t)(p.n_x * p.Lm_aug > 0 ? p.n_x * p.Lm_aug : 1)));
    CK(cudaMalloc((void**)&a.TX, sizeof(double) * (nt ? nt : 1)));
    CK(cudaMalloc((void**)&a.IF, sizeof(double) * (nt ? nt : 1)));
  }
  return 0;
}

extern "C" int pk_eval_error_data(pk_engine* e, const double* x, double* t_x, double* i_f) {
  if (!e || !x || !t_x || !i_f) return fail("pk_eval_error_data: null argument");
  if (e->aug.empty()) return fail("pk_eval_error_data: pk_engine_load_error_estimate first");
  CK(cudaSetDevice(e->device));
  if (pk_upload_x(e, x)) return 1;
  cudaStream_t st = e->stream;
  size_t off = 0;
  for (auto& a : e->aug) {
    int B = 1;
    {
      void* args[] = {&e->X, &e->FIX, &a.XS, &B};
      CK(cudaLaunchKernel((void*)a.prep, dim3(blocks_for(a.L, 128)), dim3(128), args, 0, st));
    }
    const long long nv = (a.n_x + a.n_u) * a.Lm_aug, nt = a.n_x * a.rows;
    if (nv)
      pk_csr_matvec<<<blocks_for(nv, PK_THREADS), PK_THREADS, 0, st>>>(a.Vp, a.Vi, a.Vv, a.XS, a.XU, nv, 0, 0, 1, nullptr, nullptr);
    {
      void* args[] = {&e->X, &a.XS, &a.XU, &a.tm, &a.WA, &B};
      CK(cudaLaunchKernel((void*)a.node, dim3(blocks_for(a.Lm_aug, 128)), dim3(128), args, 0, st));
    }
    if (nt) {
      pk_csr_matvec<<<blocks_for(nt, PK_THREADS), PK_THREADS, 0, st>>>(a.Tp, a.Ti, a.Tv, a.XS, a.TX, nt, 0, 0, 1, nullptr, nullptr);
      pk_csr_matvec<<<blocks_for(nt, PK_THREADS), PK_THREADS, 0, st>>>(a.Ip, a.Ii, a.Iv, a.WA, a.IF, a.rows, a.Lm_aug, a.rows, (int)a.n_x,
                                                                      a.XS + a.L - 1, a.XS + a.L - 2);
      CK(cudaMemcpyAsync(t_x + off, a.TX, sizeof(double) * (size_t)nt, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(i_f + off, a.IF, sizeof(double) * (size_t)nt, cudaMemcpyDeviceToHost, st));
    }
    e->launches += 5;
    off += (size_t)nt;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int pk_x_uploads(pk_engine* e, int64_t* count) {
  if (!e || !count) return fail("pk_x_uploads: null argument");
  *count = e->x_uploads;
  return 0;
}

extern "C" int pk_expand_variant(pk_engine* e, int mode, int* variant) {
  if (!e || !variant || mode < 0 || mode >= PK_N_MODES) return fail("pk_expand_variant: bad argument");
  const ModeState& ms = e->mode[mode];
  *variant = ms.exp.empty() ? 0 : (ms.exp.back().bulk ? 3 : (ms.exp.back().cols ? 2 : (ms.exp.back().batch ? (ms.exp.back().slots ? 5 : 4) : 1)));
  return 0;
}

extern "C" int pk_kernel_launches(pk_engine* e, int64_t* count) {
  if (!e || !count) return fail("pk_kernel_launches: null argument");
  *count = e->launches;
  return 0;
}

extern "C" int pk_flush_l2(pk_engine* e) {
  if (!e) return fail("pk_flush_l2: null engine");
  CK(cudaSetDevice(e->device));
  if (!e->flush) {
    e->n_flush = (256LL << 20) / 8;  // 256 MiB > 126 MB of L2
    CK(cudaMalloc((void**)&e->flush, sizeof(double) * (size_t)e->n_flush));
  }
  pk_fill<<<148 * 8, 256, 0, e->stream>>>(e->flush, e->n_flush, 1.0);
  CK(cudaGetLastError());
  return 0;
}
