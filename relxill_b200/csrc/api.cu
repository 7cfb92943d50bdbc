// api.cu — runtime + C ABI of librelxill_b200.so (see include/relxill_b200.h).
//
// Runtime: process-wide switches (the reference's environment variables, the state-cache switch, ...) and the list
// of engines.  Engine: one per CUDA device; owns that device's HBM-resident tables, a scratch arena sized for one
// chunk of parameter vectors, its streams and staging buffers, and runs the kernel sequence.  Batches larger than the
// chunk capacity stream through the arena chunk by chunk; a host-buffer call (relxill_batch_eval) is sharded over the
// engines — one host thread per device, every device writes its rows straight into the caller's array — and inside a
// device it is cut into pieces whose device->host copies overlap the kernels of the next piece.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/relxill_b200.h"
#include "common.h"
#include "kernels.h"
#include "models.h"
#include "tables.h"

using namespace rx;

struct relxill_b200_batch;

namespace {

thread_local std::string g_err;
void set_err(const std::string &s) {
  g_err = s;
  if (!s.empty()) fprintf(stderr, " *** relxill_b200 error: %s\n", s.c_str());
}

#define CK(call)                                                            \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      set_err(std::string(#call) + ": " + cudaGetErrorString(e_));          \
      return -2;                                                            \
    }                                                                       \
  } while (0)

enum KFam { KF_SYSPAR, KF_ZONE, KF_FINE, KF_DIST, KF_LINE, KF_XILL, KF_CONV, KF_FINISH, KF_NTH, KF_PRIMNTH, KF_XILLVER, KF_COUNT };
const char *KF_NAMES[KF_COUNT] = {"k_syspar", "k_zone", "k_fine", "k_dist", "k_line", "k_xill", "k_conv", "k_linefinish", "k_nth", "k_prim_nth", "k_xillver"};

struct Engine {
  std::mutex mu;              // serialises the calls that use this device's arena
  bool inited = false;
  int device = 0;
  Tables *tables = nullptr;
  Scratch S{};
  bool S_nth = false;         // the nthcomp work arrays of the arena are allocated
  bool S_fine = false;        // ... and the fine transfer-function / emission-angle arrays
  std::vector<void *> scratch_allocs;
  double *d_io = nullptr;     // output staging of host-buffer calls
  size_t d_io_cap = 0;
  double *h_io = nullptr;     // pinned staging for callers whose flux array is pageable memory
  size_t h_io_cap = 0;
  cudaStream_t stream_c = nullptr, stream_d = nullptr;   // compute / copy streams of host-buffer calls
  cudaEvent_t arena_busy = nullptr;                      // recorded behind the last kernel that touches the arena
  bool arena_busy_set = false;
  cudaStream_t arena_stream = nullptr;                   // ... and the stream it was recorded on
  // device / pinned buffers recycled between batches (cudaMalloc, cudaFree and cudaMallocHost synchronise and cost
  // milliseconds).  A buffer only goes back to the pool once the work that uses it has been waited for (wait_arena).
  std::vector<std::pair<size_t, void *>> pool, pool_pinned;
  // Device-resident state cache.  The scratch arena keeps the intermediates of the batch that
  // ran last and fitted the arena (arena_owner = that batch's uid); when the same batch runs again with updated
  // parameters, vectors whose relativistic half / whole parameter set is unchanged re-use their rows.
  unsigned long arena_owner = 0;
  relxill_b200_batch *retained = nullptr;   // the batch of the last host-buffer call (what an XSPEC fit re-evaluates)
};

struct Runtime {
  std::mutex mu;              // guards the engine list and the switches below
  std::vector<std::unique_ptr<Engine>> engines;
  std::string table_dir;
  HostConfig cfg;             // env-derived fields are refreshed per call (read_call_env)
  long max_chunk = 4096;
  long pipe_piece = 2072;     // vectors per pipelined piece of a host-buffer call (RELXILL_B200_PIPE)
  long pipe_last = 888;       // ... and of the last piece, whose device->host copy nothing overlaps (RELXILL_B200_PIPE_LAST)
  bool env_read = false;
  bool profiling = false;
  bool keep_intermediates = false;   // store what only the test probes read (emission-angle tables)
  // k_xill blends the convolution-grid copy of the table (rows rebinned once at load, xill.cu) and k_conv reads the zone
  // spectra with plain coalesced loads; off (RELXILL_B200_XILL_GRID=table): no such copy is built, zone spectra on the
  // table grid, rebinned per zone in k_conv
  bool xill_conv_grid = true;
  bool cache_on = true;
  int interleave = 0;         // multi-device sharding of host-buffer calls: 0 contiguous blocks, 1 round-robin rows
  unsigned long next_uid = 1;
};
Runtime g_rt;

// what one call needs of the process-wide switches, taken once under the runtime mutex
struct CallCfg {
  HostConfig cfg;
  long max_chunk, pipe_piece, pipe_last;
  bool profiling, keep_intermediates, xill_conv_grid, cache_on;
};

// The reference's environment switches, read per call like the reference does (a pyxspec session or the reference's own
// e2e tests flip them between evaluations): get_num_zones (src/relutility.c:506-544), get_returnrad_switch / is_env_set
// (src/ModelDefinition.cpp:123-149), do_not_normalize_relline (src/relutility.c:386-396), constantDiskDensity (:372-382),
// do_renorm_relxill (src/Relxill.cpp:241-247).  Their values flow into every VPar (nz, return_rad, renorm,
// const_density), so the state cache sees a change as a parameter change.
void read_call_env_locked() {
  auto is_one = [](const char *name) {
    const char *env = getenv(name);
    return (env && (int) strtod(env, nullptr) == 1) ? 1 : 0;
  };
  HostConfig &c = g_rt.cfg;
  const char *nz = getenv("RELXILL_NUM_RZONES");
  c.env_num_zones = nz ? (int) atof(nz) : 0;
  c.env_returnrad = getenv("RELXILL_RETURNRAD_SWITCH") ? is_one("RELXILL_RETURNRAD_SWITCH") : -1;
  c.env_phys_norm = is_one("RELLINE_PHYSICAL_NORM");
  c.env_const_density = is_one("RELXILL_CONSTANT_DENSITY");
  c.env_renorm_relxill = is_one("RELXILL_RENORMALIZE");
  if (!g_rt.env_read) {   // this library's own tuning knobs are read once
    g_rt.env_read = true;
    // vectors per chunk / piece: the vector index rides in gridDim.y of several kernels (limit 65535)
    if (const char *env = getenv("RELXILL_B200_CHUNK")) g_rt.max_chunk = std::min(65535L, std::max(1L, atol(env)));
    if (const char *env = getenv("RELXILL_B200_PIPE")) g_rt.pipe_piece = std::min(65535L, std::max(1L, atol(env)));
    if (const char *env = getenv("RELXILL_B200_PIPE_LAST")) g_rt.pipe_last = std::max(0L, atol(env));
    if (const char *env = getenv("RELXILL_B200_XILL_GRID")) g_rt.xill_conv_grid = std::string(env) != "table";
    if (const char *env = getenv("RELXILL_B200_INTERLEAVE")) g_rt.interleave = atoi(env) != 0;
  }
}

CallCfg call_cfg(bool refresh_env) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  if (refresh_env || !g_rt.env_read) read_call_env_locked();
  return CallCfg{g_rt.cfg, g_rt.max_chunk, g_rt.pipe_piece, g_rt.pipe_last, g_rt.profiling, g_rt.keep_intermediates,
                 g_rt.xill_conv_grid, g_rt.cache_on};
}

// ---------------------------------------------------------------------------------- buffers
void *pool_get(std::vector<std::pair<size_t, void *>> &pool, size_t bytes, bool pinned) {
  for (size_t i = 0; i < pool.size(); i++) {
    if (pool[i].first >= bytes && pool[i].first <= 2 * bytes + 4096) {
      void *p = pool[i].second;
      pool.erase(pool.begin() + i);
      return p;
    }
  }
  void *d = nullptr;
  const cudaError_t e = pinned ? cudaMallocHost(&d, bytes) : cudaMalloc(&d, bytes);
  if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return d;
}
void pool_put(std::vector<std::pair<size_t, void *>> &pool, void *p, size_t bytes, bool pinned) {
  if (!p) return;
  if (pool.size() >= 8) {
    if (pinned) cudaFreeHost(pool.front().second); else cudaFree(pool.front().second);
    pool.erase(pool.begin());
  }
  pool.emplace_back(bytes, p);
}

// Everything that was enqueued on this engine's arena (kernels of the last run, the status copy behind them) has
// finished.  Called before host code overwrites what those kernels read (parameter upload, recycled buffers).
void wait_arena(Engine &E) {
  if (E.arena_busy_set) {
    cudaEventSynchronize(E.arena_busy);
    E.arena_busy_set = false;
  }
}

void free_scratch(Engine &E) {
  wait_arena(E);
  E.arena_owner = 0;   // whatever state the arena held is gone
  for (void *p : E.scratch_allocs) cudaFree(p);
  E.scratch_allocs.clear();
  E.S = Scratch{};
  E.S_nth = false;
  E.S_fine = false;
}

int ensure_scratch(Engine &E, long cap, int nz_cap, int ne_cap, int nex_stride, bool nth, bool fine) {
  Scratch &S = E.S;
  if (S.cap >= cap && S.nz_cap >= nz_cap && S.ne_line_cap >= ne_cap && S.nex_stride >= nex_stride && (!nth || E.S_nth) &&
      (!fine || E.S_fine)) return 0;
  nth = nth || E.S_nth;
  fine = fine || E.S_fine;
  cap = std::max(cap, S.cap);
  nz_cap = std::max(nz_cap, S.nz_cap);
  ne_cap = std::max(ne_cap, S.ne_line_cap);
  nex_stride = std::max(nex_stride, S.nex_stride);
  free_scratch(E);
  bool ok = true;
  const size_t c = (size_t) cap;
  const size_t nzc = (size_t) nz_cap, nec = (size_t) ne_cap, nxs = (size_t) std::max(nex_stride, 1);
  auto alloc = [&](void **p, size_t bytes) {
    if (!ok) return;
    void *d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaGetLastError(); ok = false; return; }
    E.scratch_allocs.push_back(d);
    *p = d;
  };
#define RX_ALLOC(type, name, count) alloc((void **) &S.name, c * (size_t) (count) * sizeof(type));
#define RX_ALLOC_NTH(type, name, count) if (nth) alloc((void **) &S.name, c * (size_t) (count) * sizeof(type));
#define RX_ALLOC_FINE(type, name, count) if (fine) alloc((void **) &S.name, c * (size_t) (count) * sizeof(type));
  SCRATCH_FIELDS(RX_ALLOC, RX_ALLOC_NTH, RX_ALLOC_FINE)
#undef RX_ALLOC_FINE
#undef RX_ALLOC
#undef RX_ALLOC_NTH
  (void) nzc; (void) nec; (void) nxs;
  if (!ok) {
    free_scratch(E);
    set_err("out of device memory for the scratch arena");
    return -2;
  }
  S.cap = cap; S.nz_cap = nz_cap; S.ne_line_cap = ne_cap; S.nex_stride = nex_stride;
  E.S_nth = nth;
  E.S_fine = fine;
  return 0;
}

// The part of the arena that belongs to the vectors [c0, c0 + ...) of the batch that owns it
Scratch scratch_slice(const Scratch &S, long c0, bool nth) {
  if (c0 == 0) return S;
  Scratch R = S;
  const size_t o = (size_t) c0;
  const size_t nzc = (size_t) S.nz_cap, nec = (size_t) S.ne_line_cap, nxs = (size_t) std::max(S.nex_stride, 1);
#define RX_OFF(type, name, count) R.name = S.name + o * (size_t) (count);
#define RX_OFF_NTH(type, name, count) if (nth) R.name = S.name + o * (size_t) (count);
#define RX_OFF_FINE(type, name, count) if (S.name) R.name = S.name + o * (size_t) (count);
  SCRATCH_FIELDS(RX_OFF, RX_OFF_NTH, RX_OFF_FINE)
#undef RX_OFF_FINE
#undef RX_OFF
#undef RX_OFF_NTH
  (void) nzc; (void) nec; (void) nxs;
  if (S.reuse) R.reuse = S.reuse + o;
  R.cap = S.cap - c0;
  return R;
}

int engine_init(Engine &E, const std::string &dir, int device, bool conv_grid) {
  if (E.inited) return 0;
  CK(cudaSetDevice(device));
  E.device = device;
  if (!E.stream_c) CK(cudaStreamCreateWithFlags(&E.stream_c, cudaStreamNonBlocking));
  if (!E.stream_d) CK(cudaStreamCreateWithFlags(&E.stream_d, cudaStreamNonBlocking));
  if (!E.arena_busy) CK(cudaEventCreateWithFlags(&E.arena_busy, cudaEventDisableTiming));
  if (kernels_init() != 0) {
    set_err("kernel attribute setup failed (is this an sm_100a device?)");
    return -1;
  }
  E.tables = new Tables();
  E.tables->set_conv_grid_copy(conv_grid);
  const std::string err = E.tables->load(dir);
  if (!err.empty()) {
    set_err(err);
    delete E.tables;
    E.tables = nullptr;
    return -1;
  }
  E.inited = true;
  return 0;
}

std::string resolve_table_dir(const char *dir) {
  if (dir && *dir) return dir;
  if (const char *env = getenv("RELXILL_TABLE_PATH")) return env;   // src/relutility.c:320-328
  return "./";
}

// Engines for devices first .. first + n - 1 (n < 1: all visible devices from `first` on; first < 0: the current
// device).  Idempotent when the set is unchanged.
int runtime_init(const char *table_dir, int first, int n) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  if (!g_rt.env_read) read_call_env_locked();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_err("no CUDA device available (this library has no CPU fallback)");
    return -1;
  }
  if (first < 0) {
    if (cudaGetDevice(&first) != cudaSuccess) first = 0;
  }
  if (n < 1) n = ndev - first;
  if (n < 1 || first + n > ndev) {
    set_err("requested devices " + std::to_string(first) + ".." + std::to_string(first + n - 1) + ", " + std::to_string(ndev) + " visible");
    return -1;
  }
  if (!g_rt.engines.empty()) {
    bool same = (int) g_rt.engines.size() == n;
    for (int i = 0; same && i < n; i++) same = g_rt.engines[i]->device == first + i;
    if (same) return 0;
    set_err("already initialised on another device set: call relxill_b200_shutdown() first");
    return -1;
  }
  g_rt.table_dir = resolve_table_dir(table_dir);
  for (int i = 0; i < n; i++) {
    std::unique_ptr<Engine> e(new Engine());
    if (engine_init(*e, g_rt.table_dir, first + i, g_rt.xill_conv_grid)) {
      g_rt.engines.clear();
      return -1;
    }
    g_rt.engines.push_back(std::move(e));
  }
  cudaSetDevice(g_rt.engines[0]->device);
  return 0;
}

Engine *engine_at(int idx) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  if (idx < 0 || idx >= (int) g_rt.engines.size()) return nullptr;
  return g_rt.engines[idx].get();
}
int num_engines() {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  return (int) g_rt.engines.size();
}
// lazily: the first evaluation initialises one engine on the current device (mirrors the reference's lazy table load),
// or on the devices RELXILL_B200_DEVICES names ("all" or a count)
int ensure_runtime() {
  if (num_engines() > 0) return 0;
  int n = 1, first = -1;
  if (const char *env = getenv("RELXILL_B200_DEVICES")) {
    first = 0;
    n = (std::string(env) == "all") ? 0 : std::max(1, atoi(env));
  }
  return runtime_init(nullptr, first, n);
}

}  // namespace

// The interpreted parameter vectors of a batch in page-locked host memory (from the engine's pinned pool): a 4096-vector
// batch uploads 2.8 MB on every call, and out of pageable memory that copy cost as much as the interpretation.
// interpret_params clears every VPar it fills, so the buffer is handed over as it comes.
struct HostVPars {
  VPar *p = nullptr;
  long n = 0;
  size_t bytes = 0;
  bool pinned = false;
  VPar &operator[](long i) { return p[i]; }
  const VPar &operator[](long i) const { return p[i]; }
  VPar *data() { return p; }
  const VPar *data() const { return p; }
  long size() const { return n; }
};

struct relxill_b200_batch {
  Engine *eng = nullptr;
  const ModelDef *m = nullptr;
  long n = 0;
  int n_flux = 0;
  int nz_max = 1, nz_min = 1;
  bool any_corr = false;
  bool any_limb = false;      // some vector uses a limb law: k_fine must keep the emission angles for k_line
  int renorm3 = 0;            // RELXILL_RENORMALIZE as read when the parameters were interpreted
  HostVPars vps;             // interpreted vectors, page-locked: the upload is one DMA
  int *status = nullptr;      // [n] pinned: the status copy of a run is asynchronous
  VPar *d_vps = nullptr;
  double *d_energy = nullptr;
  size_t vps_bytes = 0, energy_bytes = 0;
  long launches = 0;
  long last_chunk0 = 0, last_chunk_n = 0;
  long arena_c0 = 0;          // arena slot of the first vector of the last chunk (0 unless the batch keeps its state)
  double kt_ms[KF_COUNT] = {0};
  long kt_n[KF_COUNT] = {0};
  // state cache: identity, the parameters the arena rows were computed from, per-vector re-use flags of the last run
  unsigned long uid = 0;
  std::vector<VPar> state_vps;
  bool state_valid = false;
  std::vector<unsigned char> reuse;
  unsigned char *d_reuse = nullptr;
  long n_reuse_rel = 0, n_reuse_all = 0;
  std::vector<double> energy;   // host copy of the grid (retained batches compare it)
  int xill_conv_grid = 0;       // the last run filed its zone spectra on the convolution grid
};

namespace {

struct Timer {
  bool on;
  relxill_b200_batch *b;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  Timer(bool on_, relxill_b200_batch *b_, cudaStream_t s) : on(on_), b(b_), st(s) {
    if (on) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
  }
  ~Timer() {
    if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
  }
  void begin() { if (on) cudaEventRecord(e0, st); }
  void end(int fam, int kernels = 1) {
    b->launches += kernels;
    b->kt_n[fam]++;
    if (on) {
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      b->kt_ms[fam] += ms;
    }
  }
};

// Host-buffer calls overlap the device->host copy of the spectra with the kernels of the next piece of the
// batch: pieces of `pipe_piece` vectors (a whole number of waves of the one-CTA-per-vector kernels), the odd
// remainder first, so that only the copy of the last piece is exposed.  A pageable destination goes through the
// engine's pinned staging buffer (cudaMemcpyAsync into pageable memory blocks the host until the copy is done, which
// would serialise the pieces): the device->host copies land there and the host moves each piece on while the GPU works
// on the next ones.
struct PipeOut {
  double *h_flux;       // the caller's array
  double *h_stage;      // pinned staging (null: h_flux is pinned or registered memory, copy straight into it)
  cudaStream_t copy_stream;
};

int run_batch(Engine &E, relxill_b200_batch *b, double *d_flux, cudaStream_t st, const CallCfg &cc, const PipeOut *pipe = nullptr) {
  CK(cudaSetDevice(E.device));
  const ModelDef &m = *b->m;
  const DevTables &T = E.tables->dev();
  const int xtab = model_xtab(m);
  const int which = xtab < 0 ? 0 : xtab;
  const bool relxill = (m.type == T_RELXILL);
  const int ne_line = (m.type == T_LINE) ? b->n_flux : NCONV;
  const int nex_stride = (relxill || m.type == T_XILL) ? std::max(E.tables->xill_host(xtab).stride, E.tables->xill_host(xtab).xc_stride) : 1;
  const int cgrid = (relxill && cc.xill_conv_grid && E.tables->xill_host(xtab).has_conv_copy) ? 1 : 0;
  b->xill_conv_grid = cgrid;
  const int n_incl = relxill ? E.tables->xill_host(xtab).n_incl : 0;
  const bool xillver = (m.type == T_XILL);
  const bool nth = (relxill || xillver) && m.prim == PRIM_NTHCOMP;
  const long cap = std::min(b->n, cc.max_chunk);
  // the previous run on this arena may still be in flight on another stream
  if (E.arena_busy_set && st != E.arena_stream) CK(cudaStreamWaitEvent(st, E.arena_busy, 0));
  const int nz_line_min = relxill ? b->nz_min : 1, nz_line_max = relxill ? b->nz_max : 1;
  const int line_nk = line_launches(nz_line_min, nz_line_max);
  const bool fine = !xillver && (b->any_limb || cc.keep_intermediates);   // the fine arrays are filed: probes, limb darkening
  if (ensure_scratch(E, cap, xillver ? b->nz_max : std::max(b->nz_max, line_rows(nz_line_min, nz_line_max)), ne_line, nex_stride, nth, fine)) return -2;
  const Scratch S0 = E.S;
  b->launches = 0;
  for (int k = 0; k < KF_COUNT; k++) { b->kt_ms[k] = 0; b->kt_n[k] = 0; }
  Timer tm(cc.profiling, b, st);
  const std::vector<double> &econv = E.tables->econv();
  // A batch that fits the arena keeps one arena slot per vector, whatever pieces it is cut into: its state survives the
  // run.  A larger batch streams through the arena chunk by chunk (every chunk starts at slot 0) and leaves no state.
  const bool resident = b->n <= S0.cap;
  // Host-buffer call of a relxill flavour that fits the arena: everything up to the zone spectra runs on the WHOLE batch
  // (full grids, no per-piece tails) and only the convolution, the last kernel, is cut into pieces whose device->host
  // copies overlap the next piece's convolution.
  const bool split_tail = pipe && resident && relxill && b->n >= cc.pipe_piece + cc.pipe_last;
  std::vector<long> piece_n;
  {
    long piece = S0.cap, last = 0;
    if (pipe && b->n >= cc.pipe_piece + cc.pipe_last) {
      piece = std::min(S0.cap, split_tail ? cc.pipe_piece / 2 : cc.pipe_piece);
      last = std::min(split_tail ? cc.pipe_last / 2 : cc.pipe_last, piece);   // only the copy of the last piece is exposed: keep that piece short
    }
    const long body = b->n - last;
    long first = body % piece;
    if (first == 0) first = std::min(piece, body);
    for (long c0 = 0, nc = first; c0 < body; c0 += nc, nc = std::min(piece, body - c0)) piece_n.push_back(nc);
    if (last > 0) piece_n.push_back(last);
  }
  // ---- state cache: which vectors can keep the rows the arena still holds for them
  b->n_reuse_rel = b->n_reuse_all = 0;
  const unsigned char *d_reuse = nullptr;
  if (cc.cache_on && resident && b->state_valid && E.arena_owner == b->uid && b->d_reuse && !cc.keep_intermediates &&
      (m.type == T_RELXILL || m.type == T_CONV || m.type == T_LINE)) {
    b->reuse.resize(b->n);
    long any = 0;
    for (long i = 0; i < b->n; i++) {
      int f = reusable_state(b->state_vps[i], b->vps[i]);
      if (m.type != T_RELXILL) f &= REUSE_REL;
      b->reuse[i] = (unsigned char) f;
      if (f & REUSE_ALL) b->n_reuse_all++; else if (f & REUSE_REL) b->n_reuse_rel++;
      any += f != 0;
    }
    if (any) {
      CK(cudaMemcpyAsync(b->d_reuse, b->reuse.data(), b->n, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));   // b->reuse is pageable: the copy has left it when the host goes on
      d_reuse = b->d_reuse;
    }
  }
  b->state_valid = false;   // until this run has been enqueued completely
  E.arena_owner = 0;
  std::vector<cudaEvent_t> ev(pipe ? piece_n.size() : 0);
  for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  std::vector<cudaEvent_t> ev_copied((pipe && pipe->h_stage) ? piece_n.size() : 0);
  for (auto &e : ev_copied) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  std::vector<long> piece_c0(piece_n.size());
  auto copy_piece = [&](size_t k) {
    const long c0 = piece_c0[k], nc = piece_n[k];
    double *dst = (pipe->h_stage ? pipe->h_stage : pipe->h_flux) + (size_t) c0 * b->n_flux;
    CK(cudaStreamWaitEvent(pipe->copy_stream, ev[k], 0));
    CK(cudaMemcpyAsync(dst, d_flux + (size_t) c0 * b->n_flux, (size_t) nc * b->n_flux * sizeof(double), cudaMemcpyDeviceToHost,
                       pipe->copy_stream));
    if (pipe->h_stage) CK(cudaEventRecord(ev_copied[k], pipe->copy_stream));
    return 0;
  };
  long c0 = 0;
  for (size_t ip = 0; ip < piece_n.size(); c0 += piece_n[ip], ip++) {
    const long nc = piece_n[ip];
    piece_c0[ip] = c0;
    if (pipe && ip > 0) {   // the previous piece's kernels are enqueued behind it: its copy overlaps this piece
      if (copy_piece(ip - 1)) return -2;
    }
    struct AtEnd {   // record the piece's completion event however the body is left
      std::vector<cudaEvent_t> &ev; size_t ip; cudaStream_t st; bool on;
      ~AtEnd() { if (on) cudaEventRecord(ev[ip], st); }
    } at_end{ev, ip, st, pipe != nullptr};
    const long slot0 = resident ? c0 : 0;
    Scratch S = scratch_slice(S0, slot0, nth);
    S.reuse = d_reuse ? d_reuse + c0 : nullptr;
    const VPar *vps = b->d_vps + c0;
    double *out = d_flux + (size_t) c0 * b->n_flux;
    b->last_chunk0 = split_tail ? 0 : c0;
    b->last_chunk_n = split_tail ? b->n : nc;
    b->arena_c0 = split_tail ? 0 : slot0;
    if (split_tail) {
      if (ip == 0) {   // the stages before the convolution, once, on all vectors
        Scratch Sa = S0;
        Sa.reuse = d_reuse;
        const VPar *va = b->d_vps;
        const long na = b->n;
        tm.begin(); launch_syspar(va, T, Sa, na, 1, st); tm.end(KF_SYSPAR);
        tm.begin(); launch_zone(va, T, Sa, na, st); tm.end(KF_ZONE);
        if (nth) { tm.begin(); launch_nth(va, T, Sa, na, b->nz_max, st); tm.end(KF_NTH, 2); }
        if (b->any_corr) { tm.begin(); launch_syspar(va, T, Sa, na, 2, st); tm.end(KF_SYSPAR); }
        tm.begin();
        launch_fine(va, T, Sa, na, n_incl, econv[0], econv[NCONV], (b->any_limb || cc.keep_intermediates) ? 1 : 0,
                    cc.keep_intermediates ? 1 : 0, st);
        tm.end(KF_FINE, 2);
        tm.begin(); launch_dist(va, T, Sa, na, n_incl, st); tm.end(KF_DIST);
        tm.begin(); launch_line(va, T, Sa, na, T.econv, NCONV, 0, nz_line_min, nz_line_max, st); tm.end(KF_LINE, line_nk);
        tm.begin(); launch_xill(va, T, Sa, na, which, cgrid, st); tm.end(KF_XILL);
      }
      // the zone spectra were filed by a launch over the whole batch: the kernels stride them by the row length in use
      // (k_xill / k_conv: nz_cap x row stride per vector), not by the arena's allocation stride that scratch_slice applies
      S.xillz = S0.xillz + (size_t) c0 * S0.nz_cap * (size_t) (cgrid ? E.tables->xill_host(xtab).xc_stride : E.tables->xill_host(xtab).stride);
      tm.begin(); launch_conv(vps, T, S, nc, b->d_energy, b->n_flux, out, S.total, which, 0, cgrid, b->renorm3, st); tm.end(KF_CONV);
      if (nth) {
        tm.begin(); launch_prim_nth(vps, T, S, nc, S.total, b->d_energy, b->n_flux, out, b->renorm3, st); tm.end(KF_PRIMNTH);
      }
      CK(cudaMemcpyAsync(b->status + c0, S.status, nc * sizeof(int), cudaMemcpyDeviceToHost, st));
      continue;
    }
    if (xillver) {
      tm.begin(); launch_xillver(vps, T, S, nc, which, b->d_energy, b->n_flux, out, nex_stride, st); tm.end(KF_XILLVER);
      if (nth) {
        tm.begin(); launch_nth(vps, T, S, nc, 0, st); tm.end(KF_NTH, 2);
        tm.begin(); launch_xillver_prim_nth(vps, T, S, nc, b->d_energy, b->n_flux, out, st); tm.end(KF_PRIMNTH);
      }
      CK(cudaMemcpyAsync(b->status + c0, S.status, nc * sizeof(int), cudaMemcpyDeviceToHost, st));
      continue;
    }
    tm.begin(); launch_syspar(vps, T, S, nc, 1, st); tm.end(KF_SYSPAR);
    if (relxill) {
      tm.begin(); launch_zone(vps, T, S, nc, st); tm.end(KF_ZONE);
      if (nth) { tm.begin(); launch_nth(vps, T, S, nc, b->nz_max, st); tm.end(KF_NTH, 2); }
      if (b->any_corr) { tm.begin(); launch_syspar(vps, T, S, nc, 2, st); tm.end(KF_SYSPAR); }
    }
    tm.begin();
    launch_fine(vps, T, S, nc, relxill ? n_incl : 0, econv[0], econv[NCONV], (b->any_limb || cc.keep_intermediates) ? 1 : 0,
                cc.keep_intermediates ? 1 : 0, st);
    tm.end(KF_FINE, 2);   // two kernels (k_rows, k_fine) timed as one family
    if (relxill) {
      tm.begin(); launch_dist(vps, T, S, nc, n_incl, st); tm.end(KF_DIST);
    }
    if (m.type == T_LINE) {
      tm.begin(); launch_line(vps, T, S, nc, b->d_energy, b->n_flux, 1, 1, 1, st); tm.end(KF_LINE, line_nk);
      tm.begin(); launch_linefinish(vps, S, nc, b->n_flux, out, st); tm.end(KF_FINISH);
    } else {
      tm.begin(); launch_line(vps, T, S, nc, T.econv, NCONV, 0, nz_line_min, nz_line_max, st); tm.end(KF_LINE, line_nk);
      if (relxill) {
        tm.begin(); launch_xill(vps, T, S, nc, which, cgrid, st); tm.end(KF_XILL);
        tm.begin(); launch_conv(vps, T, S, nc, b->d_energy, b->n_flux, out, S.total, which, 0, cgrid, b->renorm3, st); tm.end(KF_CONV);
        if (nth) {
          tm.begin(); launch_prim_nth(vps, T, S, nc, S.total, b->d_energy, b->n_flux, out, b->renorm3, st); tm.end(KF_PRIMNTH);
        }
      } else {
        tm.begin(); launch_conv(vps, T, S, nc, b->d_energy, b->n_flux, out, nullptr, 0, 1, 0, 0, st); tm.end(KF_CONV);
      }
    }
    CK(cudaMemcpyAsync(b->status + c0, S.status, nc * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaEventRecord(E.arena_busy, st));
  E.arena_busy_set = true;
  E.arena_stream = st;
  if (pipe) {
    int rc = 0;
    if (copy_piece(piece_n.size() - 1)) return -2;
    if (pipe->h_stage) {   // pageable destination: move every piece on as soon as it has landed in the staging buffer
      for (size_t k = 0; k < piece_n.size(); k++) {
        if (cudaEventSynchronize(ev_copied[k]) != cudaSuccess) { rc = -2; break; }
        const size_t off = (size_t) piece_c0[k] * b->n_flux;
        memcpy(pipe->h_flux + off, pipe->h_stage + off, (size_t) piece_n[k] * b->n_flux * sizeof(double));
      }
    }
    if (cudaStreamSynchronize(pipe->copy_stream) != cudaSuccess) rc = -2;
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = -2;
    for (auto &e : ev) cudaEventDestroy(e);
    for (auto &e : ev_copied) cudaEventDestroy(e);
    if (rc) { set_err(std::string("pipelined run: ") + cudaGetErrorString(cudaGetLastError())); return rc; }
  }
  CK(cudaGetLastError());
  if (resident && !xillver) {   // the arena now holds this batch's state
    b->state_vps.assign(b->vps.data(), b->vps.data() + b->vps.size());
    b->state_valid = true;
    E.arena_owner = b->uid;
  }
  return 0;
}

// A few resident worker threads for the host-side interpretation of large batches: creating sixteen threads per call cost
// more (0.5 ms) than the interpretation they shared (0.75 ms on one core for 4096 vectors).  One job at a time; a caller
// that finds the crew busy (another device's shard is being interpreted) does its own work inline.
class HostCrew {
 public:
  explicit HostCrew(int n) : pid_(getpid()) {
    for (int k = 0; k < n; k++) th_.emplace_back([this, k] { loop(k); });
  }
  ~HostCrew() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  int size() const { return (int) th_.size(); }
  // runs work(part, parts) for part = 0 .. parts-1 (parts = size() + 1: the caller takes the last part); false if busy
  bool run(const std::function<void(int, int)> &work) {
    if (getpid() != pid_) return false;   // a forked child has the object but not the threads
    std::unique_lock<std::mutex> job(job_mu_, std::try_to_lock);
    if (!job.owns_lock()) return false;
    const int parts = size() + 1;
    {
      std::lock_guard<std::mutex> lk(mu_);
      work_ = &work;
      pending_ = size();
      gen_++;
    }
    cv_.notify_all();
    work(parts - 1, parts);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    work_ = nullptr;
    return true;
  }

 private:
  void loop(int k) {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(int, int)> *w;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
        if (quit_) return;
        seen = gen_;
        w = work_;
      }
      (*w)(k, size() + 1);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_, job_mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int, int)> *work_ = nullptr;
  unsigned long gen_ = 0;
  int pending_ = 0;
  bool quit_ = false;
  pid_t pid_;
};
HostCrew &host_crew() {
  static HostCrew crew((int) std::max(1u, std::min(7u, std::thread::hardware_concurrency() > 1 ? std::thread::hardware_concurrency() - 1 : 1u)));
  return crew;
}

// host-side interpretation of the raw parameter vectors (spread over a few threads for large batches) and the
// batch-level switches derived from it
void interpret_all(Engine &E, relxill_b200_batch *b, const double *params, const CallCfg &cc) {
  const ModelDef *m = b->m;
  const long n_vec = b->n;
  const std::vector<double> &sp = E.tables->rr_spins();
  auto work = [&](long lo, long hi) {
    for (long i = lo; i < hi; i++)
      interpret_params(*m, params + (size_t) i * m->npar, cc.cfg, sp.empty() ? nullptr : sp.data(), (int) sp.size(), b->vps[i]);
  };
  bool shared = false;
  if (n_vec >= 1024) {
    const std::function<void(int, int)> part = [&](int k, int parts) { work(n_vec * k / parts, n_vec * (k + 1) / parts); };
    shared = host_crew().run(part);
  }
  if (!shared) work(0, n_vec);
  b->nz_max = 1;
  b->nz_min = NZMAX;
  b->any_corr = b->any_limb = false;
  b->renorm3 = cc.cfg.env_renorm_relxill;
  for (long i = 0; i < n_vec; i++) {
    if (b->vps[i].status == ST_OK) {
      b->nz_max = std::max(b->nz_max, b->vps[i].nz);
      b->nz_min = std::min(b->nz_min, b->vps[i].nz);
      if (b->vps[i].do_corr) b->any_corr = true;
      if (b->vps[i].limb != 0) b->any_limb = true;
    }
  }
}

void free_batch_locked(Engine &E, relxill_b200_batch *b) {
  if (!b) return;
  cudaSetDevice(E.device);
  wait_arena(E);   // nothing in flight reads the buffers that go back to the pool
  if (E.arena_owner == b->uid) E.arena_owner = 0;
  if (b->vps.pinned) pool_put(E.pool_pinned, b->vps.p, b->vps.bytes, true); else free(b->vps.p);
  b->vps.p = nullptr;
  pool_put(E.pool, b->d_vps, b->vps_bytes, false);
  pool_put(E.pool, b->d_energy, b->energy_bytes, false);
  pool_put(E.pool, b->d_reuse, (size_t) b->n, false);
  pool_put(E.pool_pinned, b->status, (size_t) b->n * sizeof(int), true);
  delete b;
}

relxill_b200_batch *prepare_on(Engine &E, const char *model, const double *energy, int n_flux, const double *params,
                               long n_vec, const CallCfg &cc) {
  std::lock_guard<std::mutex> lk(E.mu);
  g_err.clear();
  if (cudaSetDevice(E.device) != cudaSuccess) { set_err("cannot select the engine's device"); return nullptr; }
  const ModelDef *m = model ? find_model(model) : nullptr;
  if (!m) { set_err(std::string("unknown model ") + (model ? model : "(null)")); return nullptr; }
  if (n_vec < 1 || n_flux < 1 || !energy || !params) { set_err("empty batch or energy grid"); return nullptr; }
  if (m->type == T_LINE && n_flux > line_max_bins()) {
    set_err("line models: energy grids above " + std::to_string(line_max_bins()) + " bins are not supported yet");
    return nullptr;
  }
  // tables this flavour needs; returning radiation can be switched per vector -> load if the table exists
  bool want_rr = (m->irrad == EMIS_LP) || cc.cfg.env_returnrad == 1;
  std::string err = (m->type == T_XILL)
                        ? E.tables->require_xill_only(model_xtab(*m))
                        : E.tables->require(m->irrad == EMIS_LP, false, model_xtab(*m));
  if (!err.empty()) { set_err(err); return nullptr; }
  if (want_rr) {
    err = E.tables->require(false, true, XT_NONE);
    // missing table is only an error for the vectors that switch returning radiation on
  }
  auto *b = new relxill_b200_batch();
  b->eng = &E;
  b->m = m;
  b->n = n_vec;
  b->n_flux = n_flux;
  b->vps.bytes = (size_t) n_vec * sizeof(VPar);
  b->vps.p = (VPar *) pool_get(E.pool_pinned, b->vps.bytes, true);
  b->vps.pinned = b->vps.p != nullptr;
  if (!b->vps.p) b->vps.p = (VPar *) malloc(b->vps.bytes);   // no page-locked memory to be had: pageable will do
  if (!b->vps.p) { set_err("out of host memory (batch)"); delete b; return nullptr; }
  b->vps.n = n_vec;
  {
    std::lock_guard<std::mutex> lr(g_rt.mu);
    b->uid = g_rt.next_uid++;
  }
  b->energy.assign(energy, energy + n_flux + 1);
  static const bool dbg_t = getenv("RELXILL_B200_TIMING") != nullptr;
  const auto tp0 = std::chrono::steady_clock::now();
  interpret_all(E, b, params, cc);
  const auto tp1 = std::chrono::steady_clock::now();
  b->vps_bytes = n_vec * sizeof(VPar);
  b->energy_bytes = (n_flux + 1) * sizeof(double);
  b->d_vps = (VPar *) pool_get(E.pool, b->vps_bytes, false);
  b->d_energy = (double *) pool_get(E.pool, b->energy_bytes, false);
  b->d_reuse = (unsigned char *) pool_get(E.pool, (size_t) n_vec, false);
  b->status = (int *) pool_get(E.pool_pinned, (size_t) n_vec * sizeof(int), true);
  if (!b->d_vps || !b->d_energy || !b->d_reuse || !b->status) {
    set_err("out of device memory (batch)");
    free_batch_locked(E, b);
    return nullptr;
  }
  for (long i = 0; i < n_vec; i++) b->status[i] = b->vps[i].status;
  if (cudaMemcpy(b->d_vps, b->vps.data(), n_vec * sizeof(VPar), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(b->d_energy, energy, (n_flux + 1) * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_err(std::string("upload of the batch failed: ") + cudaGetErrorString(cudaGetLastError()));
    free_batch_locked(E, b);
    return nullptr;
  }
  if (dbg_t && n_vec >= 1024) {
    const auto tp2 = std::chrono::steady_clock::now();
    fprintf(stderr, "prepare_on: interpret %.2f ms, buffers + upload %.2f ms\n", std::chrono::duration<double, std::milli>(tp1 - tp0).count(),
            std::chrono::duration<double, std::milli>(tp2 - tp1).count());
  }
  return b;
}

int update_params_locked(Engine &E, relxill_b200_batch *b, const double *params, const CallCfg &cc) {
  CK(cudaSetDevice(E.device));
  wait_arena(E);   // the kernels of the last run read d_vps
  interpret_all(E, b, params, cc);
  CK(cudaMemcpy(b->d_vps, b->vps.data(), b->n * sizeof(VPar), cudaMemcpyHostToDevice));
  return 0;
}

int update_energy_locked(Engine &E, relxill_b200_batch *b, const double *energy, int n_flux) {
  CK(cudaSetDevice(E.device));
  wait_arena(E);
  if (b->m->type == T_LINE) {
    if (n_flux > line_max_bins()) { set_err("line models: energy grid too long"); return -1; }
    b->state_valid = false;   // the line models integrate on the caller's grid: nothing survives a new grid
  }
  const size_t bytes = (size_t) (n_flux + 1) * sizeof(double);
  if (bytes > b->energy_bytes) {
    double *d = (double *) pool_get(E.pool, bytes, false);
    if (!d) { set_err("out of device memory (energy grid)"); return -2; }
    pool_put(E.pool, b->d_energy, b->energy_bytes, false);
    b->d_energy = d;
    b->energy_bytes = bytes;
  }
  b->n_flux = n_flux;
  b->energy.assign(energy, energy + n_flux + 1);
  CK(cudaMemcpy(b->d_energy, energy, bytes, cudaMemcpyHostToDevice));
  return 0;
}

// One device's share of a host-buffer call: params [n_vec][npar], flux [n_vec][n_flux] and status [n_vec] are host
// arrays of this shard only.
int eval_on_engine(Engine &E, const char *model, const double *energy, int n_flux, const double *params, long n_vec,
                   double *flux, int *status, const CallCfg &cc) {
  static const bool dbg = getenv("RELXILL_B200_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  auto fail = [&](int rc) {
    if (flux && n_vec > 0 && n_flux > 0) memset(flux, 0, sizeof(double) * (size_t) n_vec * n_flux);
    if (status) for (long i = 0; i < n_vec; i++) status[i] = ST_BAD_PARAM;
    return rc;
  };
  // An XSPEC fit (or any caller that re-evaluates the same vectors with a few parameters changed) comes back with the
  // same model and batch size: the batch of the previous call was kept, so that the vectors whose relativistic half
  // or whole parameter set is unchanged re-use the state still resident in the arena (run_batch).
  relxill_b200_batch *b = nullptr;
  {
    std::lock_guard<std::mutex> lk(E.mu);
    relxill_b200_batch *r = E.retained;
    E.retained = nullptr;
    if (r && cc.cache_on && r->m == find_model(model) && r->n == n_vec && n_flux >= 1 && energy && params) {
      int rc = 0;
      if (n_flux != r->n_flux || memcmp(energy, r->energy.data(), sizeof(double) * (size_t) (n_flux + 1)) != 0)
        rc = update_energy_locked(E, r, energy, n_flux);
      if (rc == 0) rc = update_params_locked(E, r, params, cc);
      if (rc == 0) b = r; else free_batch_locked(E, r);
    } else {
      free_batch_locked(E, r);
    }
  }
  if (!b) b = prepare_on(E, model, energy, n_flux, params, n_vec, cc);
  const auto t1 = now();
  if (!b) return fail(-1);
  std::lock_guard<std::mutex> lk(E.mu);
  if (cudaSetDevice(E.device) != cudaSuccess) { free_batch_locked(E, b); return fail(-2); }
  const size_t need = (size_t) n_vec * n_flux;
  if (E.d_io_cap < need) {
    wait_arena(E);
    if (E.d_io) cudaFree(E.d_io);
    E.d_io = nullptr;
    E.d_io_cap = 0;
    if (cudaMalloc((void **) &E.d_io, need * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      set_err("out of device memory (output staging)");
      free_batch_locked(E, b);
      return fail(-2);
    }
    E.d_io_cap = need;
  }
  // is the caller's array page-locked?  (cudaMemcpyAsync into pageable memory is synchronous: stage those)
  bool pageable = true;
  {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, flux) == cudaSuccess) pageable = (at.type == cudaMemoryTypeUnregistered);
    else cudaGetLastError();
  }
  const bool staged = pageable && n_vec >= cc.pipe_piece + cc.pipe_last;   // small batches run in one piece: nothing to overlap
  if (staged && E.h_io_cap < need) {
    if (E.h_io) cudaFreeHost(E.h_io);
    E.h_io = nullptr;
    E.h_io_cap = 0;
    if (cudaMallocHost((void **) &E.h_io, need * sizeof(double)) == cudaSuccess) E.h_io_cap = need;
    else cudaGetLastError();   // no pinned memory to be had: fall back to the direct (serialising) copy
  }
  if (b->m->type == T_CONV) {
    // convolution models: flux is the input spectrum; a non-positive total is rejected (src/LocalModel.cpp:84-86)
    bool changed = false;
    for (long i = 0; i < n_vec; i++) {
      double s = 0.0;
      for (int j = 0; j < n_flux; j++) s += flux[(size_t) i * n_flux + j];
      if (s <= 0.0 && b->vps[i].status == ST_OK) { b->vps[i].status = ST_CONV_INPUT; changed = true; }
    }
    wait_arena(E);
    cudaError_t e = cudaSuccess;
    if (changed) e = cudaMemcpy(b->d_vps, b->vps.data(), n_vec * sizeof(VPar), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(E.d_io, flux, need * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_err(std::string("upload of the input spectra failed: ") + cudaGetErrorString(e));
      free_batch_locked(E, b);
      return fail(-2);
    }
  }
  const auto t2 = now();
  PipeOut pipe{flux, (staged && E.h_io_cap >= need) ? E.h_io : nullptr, E.stream_d};
  int rc = run_batch(E, b, E.d_io, E.stream_c, cc, &pipe);   // kernels + pipelined D2H; returns with both streams drained
  if (rc != 0) { cudaStreamSynchronize(E.stream_c); cudaStreamSynchronize(E.stream_d); }
  const auto t3 = now();
  if (status) for (long i = 0; i < n_vec; i++) status[i] = b->status[i];
  if (rc == 0 && b->state_valid && cc.cache_on) E.retained = b;
  else free_batch_locked(E, b);
  if (rc != 0) return fail(rc);
  if (dbg)
    fprintf(stderr, "relxill_batch_eval[dev %d] timing: prepare %.2f ms, staging %.2f ms, run + D2H (pipelined%s) %.2f ms, tail %.2f ms\n",
            E.device, ms(t0, t1), ms(t1, t2), pipe.h_stage ? ", pageable destination staged" : "", ms(t2, t3), ms(t3, now()));
  return 0;
}

}  // namespace

// =================================================================================== C ABI
extern "C" {

const char *relxill_b200_last_error(void) { return g_err.c_str(); }

int relxill_b200_init(const char *table_dir, int device) { return runtime_init(table_dir, device, 1); }

int relxill_b200_init_devices(const char *table_dir, int n_devices) { return runtime_init(table_dir, 0, n_devices); }

int relxill_b200_num_devices(void) { return num_engines(); }

void relxill_b200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  for (auto &ep : g_rt.engines) {
    Engine &E = *ep;
    std::lock_guard<std::mutex> le(E.mu);
    cudaSetDevice(E.device);
    free_batch_locked(E, E.retained);
    E.retained = nullptr;
    free_scratch(E);
    if (E.d_io) cudaFree(E.d_io);
    if (E.h_io) cudaFreeHost(E.h_io);
    for (auto &pr : E.pool) cudaFree(pr.second);
    for (auto &pr : E.pool_pinned) cudaFreeHost(pr.second);
    if (E.stream_c) cudaStreamDestroy(E.stream_c);
    if (E.stream_d) cudaStreamDestroy(E.stream_d);
    if (E.arena_busy) cudaEventDestroy(E.arena_busy);
    delete E.tables;
  }
  g_rt.engines.clear();
}

// the global switches are taken under the runtime mutex: a call in flight on another thread sees either the old or the new value
void relxill_b200_set_num_zones(int n) { std::lock_guard<std::mutex> lk(g_rt.mu); g_rt.cfg.override_num_zones = n > 0 ? n : 0; }
void relxill_b200_set_profiling(int on) { std::lock_guard<std::mutex> lk(g_rt.mu); g_rt.profiling = on != 0; }
void relxill_b200_keep_intermediates(int on) { std::lock_guard<std::mutex> lk(g_rt.mu); g_rt.keep_intermediates = on != 0; }
void relxill_b200_set_cache(int on) { std::lock_guard<std::mutex> lk(g_rt.mu); g_rt.cache_on = on != 0; }
void relxill_b200_set_xill_grid(int conv_grid) { std::lock_guard<std::mutex> lk(g_rt.mu); g_rt.xill_conv_grid = conv_grid != 0; }
int relxill_b200_get_xill_grid(void) { std::lock_guard<std::mutex> lk(g_rt.mu); return g_rt.xill_conv_grid ? 1 : 0; }
void relxill_b200_set_sharding(int interleave) { std::lock_guard<std::mutex> lk(g_rt.mu); g_rt.interleave = interleave != 0; }
void relxill_b200_set_xill_generic(int on) { xill_force_generic(on); }

int relxill_b200_num_params(const char *model) {
  const ModelDef *m = find_model(model);
  return m ? m->npar : -1;
}
int relxill_b200_default_params(const char *model, double *out) {
  const ModelDef *m = find_model(model);
  if (!m) return -1;
  for (int i = 0; i < m->npar; i++) out[i] = m->def[i];
  return m->npar;
}

relxill_b200_batch *relxill_b200_prepare_on(int device_index, const char *model, const double *energy, int n_flux,
                                            const double *params, long n_vec) {
  g_err.clear();
  if (ensure_runtime()) return nullptr;
  Engine *E = engine_at(device_index);
  if (!E) { set_err("prepare: no engine with index " + std::to_string(device_index)); return nullptr; }
  const CallCfg cc = call_cfg(true);
  return prepare_on(*E, model, energy, n_flux, params, n_vec, cc);
}

relxill_b200_batch *relxill_b200_prepare(const char *model, const double *energy, int n_flux, const double *params,
                                         long n_vec) {
  return relxill_b200_prepare_on(0, model, energy, n_flux, params, n_vec);
}

void relxill_b200_free_batch(relxill_b200_batch *b) {
  if (!b) return;
  Engine &E = *b->eng;
  std::lock_guard<std::mutex> lk(E.mu);
  free_batch_locked(E, b);
}

int relxill_b200_update_params(relxill_b200_batch *b, const double *params) {
  if (!b || !params) { set_err("update_params: null argument"); return -1; }
  const CallCfg cc = call_cfg(true);
  Engine &E = *b->eng;
  std::lock_guard<std::mutex> lk(E.mu);
  return update_params_locked(E, b, params, cc);
}

int relxill_b200_update_energy(relxill_b200_batch *b, const double *energy, int n_flux) {
  if (!b || !energy || n_flux < 1) { set_err("update_energy: bad argument"); return -1; }
  Engine &E = *b->eng;
  std::lock_guard<std::mutex> lk(E.mu);
  return update_energy_locked(E, b, energy, n_flux);
}

int relxill_b200_reuse_counts(relxill_b200_batch *b, long *out3) {
  if (!b || !out3) return -1;
  out3[0] = b->n - b->n_reuse_rel - b->n_reuse_all;
  out3[1] = b->n_reuse_rel;
  out3[2] = b->n_reuse_all;
  return 0;
}

int relxill_b200_last_eval_reuse(long *out3) {
  if (!out3) return -1;
  out3[0] = out3[1] = out3[2] = 0;
  std::lock_guard<std::mutex> lk(g_rt.mu);
  for (auto &ep : g_rt.engines) {
    std::lock_guard<std::mutex> le(ep->mu);
    const relxill_b200_batch *r = ep->retained;
    if (!r) continue;
    out3[0] += r->n - r->n_reuse_rel - r->n_reuse_all;
    out3[1] += r->n_reuse_rel;
    out3[2] += r->n_reuse_all;
  }
  return 0;
}

int relxill_b200_run(relxill_b200_batch *b, double *d_flux, void *stream) {
  if (!b) { set_err("run: null batch"); return -1; }
  const CallCfg cc = call_cfg(false);
  Engine &E = *b->eng;
  std::lock_guard<std::mutex> lk(E.mu);
  return run_batch(E, b, d_flux, (cudaStream_t) stream, cc);
}

int relxill_b200_batch_status(relxill_b200_batch *b, int *status) {
  if (!b) return -1;
  Engine &E = *b->eng;
  std::lock_guard<std::mutex> lk(E.mu);
  cudaSetDevice(E.device);
  wait_arena(E);
  for (long i = 0; i < b->n; i++) status[i] = b->status[i];
  return 0;
}

long relxill_b200_last_launches(relxill_b200_batch *b) { return b ? b->launches : 0; }

int relxill_b200_kernel_times(relxill_b200_batch *b, const char **names, double *ms, long *launches, int max) {
  if (!b) return 0;
  int n = 0;
  for (int k = 0; k < KF_COUNT && n < max; k++) {
    if (b->kt_n[k] == 0) continue;
    names[n] = KF_NAMES[k];
    ms[n] = b->kt_ms[k];
    launches[n] = b->kt_n[k];
    n++;
  }
  return n;
}

int relxill_batch_eval_device(const char *model, const double *energy, int n_flux, const double *params, long n_vec,
                              double *d_flux, int *status, void *stream) {
  relxill_b200_batch *b = relxill_b200_prepare(model, energy, n_flux, params, n_vec);
  if (!b) return -1;
  int rc = relxill_b200_run(b, d_flux, stream);
  if (cudaStreamSynchronize((cudaStream_t) stream) != cudaSuccess && rc == 0) rc = -2;
  if (status) for (long i = 0; i < n_vec; i++) status[i] = b->status[i];
  relxill_b200_free_batch(b);
  return rc;
}

int relxill_batch_eval(const char *model, const double *energy, int n_flux, const double *params, long n_vec,
                       double *flux, int *status) {
  g_err.clear();
  auto fail = [&](int rc) {
    if (flux && n_vec > 0 && n_flux > 0) memset(flux, 0, sizeof(double) * (size_t) n_vec * n_flux);
    if (status) for (long i = 0; i < n_vec; i++) status[i] = ST_BAD_PARAM;
    return rc;
  };
  if (ensure_runtime()) return fail(-1);
  const CallCfg cc = call_cfg(true);
  const ModelDef *m = model ? find_model(model) : nullptr;
  if (!m) { set_err(std::string("unknown model ") + (model ? model : "(null)")); return fail(-1); }
  if (n_vec < 1 || n_flux < 1 || !energy || !params || !flux) { set_err("empty batch or energy grid"); return fail(-1); }
  const int ndev = num_engines();
  // ---- one device, or a batch too small to be worth sharding
  if (ndev == 1 || n_vec < 2L * ndev) return eval_on_engine(*engine_at(0), model, energy, n_flux, params, n_vec, flux, status, cc);
  // ---- the batch sharded over the devices of this process: contiguous blocks (each device writes its rows of the
  // caller's arrays directly) or, for structured batches such as parameter-grid sweeps whose cost varies along the
  // batch, round-robin rows.  No collective: the vectors are independent.
  int interleave;
  {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    interleave = g_rt.interleave;
  }
  const int npar = m->npar;
  std::vector<int> rcs(ndev, 0);
  std::vector<std::string> errs(ndev);
  std::vector<std::thread> th;
  for (int d = 0; d < ndev; d++) {
    th.emplace_back([&, d] {
      Engine &E = *engine_at(d);
      if (!interleave) {
        const long base = n_vec / ndev, rem = n_vec % ndev;
        const long lo = d * base + std::min<long>(d, rem), cnt = base + (d < rem ? 1 : 0);
        rcs[d] = eval_on_engine(E, model, energy, n_flux, params + (size_t) lo * npar, cnt, flux + (size_t) lo * n_flux,
                                status ? status + lo : nullptr, cc);
      } else {
        const long cnt = (n_vec - d + ndev - 1) / ndev;
        std::vector<double> p((size_t) cnt * npar), f((size_t) cnt * n_flux);
        std::vector<int> s(cnt);
        for (long k = 0; k < cnt; k++) {
          const long row = d + k * ndev;
          memcpy(&p[(size_t) k * npar], params + (size_t) row * npar, npar * sizeof(double));
          if (m->type == T_CONV) memcpy(&f[(size_t) k * n_flux], flux + (size_t) row * n_flux, n_flux * sizeof(double));
        }
        rcs[d] = eval_on_engine(E, model, energy, n_flux, p.data(), cnt, f.data(), s.data(), cc);
        for (long k = 0; k < cnt; k++) {
          const long row = d + k * ndev;
          memcpy(flux + (size_t) row * n_flux, &f[(size_t) k * n_flux], n_flux * sizeof(double));
          if (status) status[row] = s[k];
        }
      }
      errs[d] = g_err;
    });
  }
  for (auto &x : th) x.join();
  for (int d = 0; d < ndev; d++)
    if (rcs[d] != 0) { g_err = errs[d]; return rcs[d]; }
  return 0;
}

int relxill_b200_algorithmic_bytes(relxill_b200_batch *b, double *out8) {
  double *out4 = out8;
  if (!b) return -1;
  Engine &E = *b->eng;
  std::lock_guard<std::mutex> lk(E.mu);
  cudaSetDevice(E.device);
  cudaDeviceSynchronize();
  const ModelDef &m = *b->m;
  const long nc = b->last_chunk_n;
  const Scratch S = scratch_slice(E.S, b->arena_c0, false);
  double sumU = 0, bound = 0;
  double xbytes = 0, xbytes_union = 0;
  if (m.type == T_RELXILL && nc > 0) {
    const XillHost &xh = E.tables->xill_host(model_xtab(m));
    const int ncorn = (xh.npar == 6) ? 32 : 16;
    std::vector<int> rows((size_t) nc * NZMAX * 32);
    if (cudaMemcpy(rows.data(), S.xrow, rows.size() * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    const double row_bytes = (double) xh.n_incl * xh.n_ener * 4.0;
    std::set<int> all;   // rows any vector of the chunk reads: what has to come out of DRAM at least once per launch
    for (long i = 0; i < nc; i++) {
      const VPar &vp = b->vps[b->last_chunk0 + i];
      if (b->status[b->last_chunk0 + i] != ST_OK) continue;
      std::set<int> u;
      for (int z = 0; z < vp.nz; z++)
        for (int c = 0; c < ncorn; c++) u.insert(rows[((size_t) i * NZMAX + z) * 32 + c]);
      sumU += (double) u.size();
      all.insert(u.begin(), u.end());
      bound += (double) vp.nz * ncorn * row_bytes;
    }
    // the rows k_xill actually reads: the fp64 convolution-grid copy (xc_n bins) or the f32 table rows
    xbytes_union = (double) all.size() * (double) xh.n_incl * (b->xill_conv_grid ? xh.xc_n * 8.0 : xh.n_ener * 4.0)
                   * ((double) b->n / (double) nc > 1.0 ? (double) b->n / (double) nc : 1.0);
    const double scale = (double) b->n / (double) nc;  // chunks beyond the last are assumed alike
    sumU *= scale;
    bound *= scale;
    xbytes = sumU * row_bytes;
  }
  // rel table 4 corners + lp + rrad + params in + spectrum out (+ shared grid once)
  double per_vec = 4.0 * (3 * 100 + 4 * 100 * 40) * 4.0 + m.npar * 8.0 + b->n_flux * 8.0;
  if (m.irrad == EMIS_LP) per_vec += 2 * 2 * 3 * 100 * 4.0;
  double rr = 0;
  for (long i = 0; i < b->n; i++)
    if (b->vps[i].status == ST_OK && b->vps[i].return_rad != 0) rr += (3 * 50 * 50 + 50 * 50 * 20) * 8.0;
  out4[0] = xbytes + per_vec * (double) b->n + rr + (b->n_flux + 1) * 8.0;
  out4[1] = sumU;
  out4[2] = xbytes;
  out4[3] = bound;
  // bytes of the per-zone line profiles that exist (only the bins between a zone's first and last non-zero bin
  // are written by k_line and read by k_conv)
  double prof = 0;
  if (nc > 0 && (m.type == T_RELXILL || m.type == T_LINE || m.type == T_CONV)) {
    std::vector<int> zr((size_t) nc * NZMAX * 2);
    if (cudaMemcpy(zr.data(), S.zrange, zr.size() * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    for (long i = 0; i < nc; i++) {
      const VPar &vp = b->vps[b->last_chunk0 + i];
      if (b->status[b->last_chunk0 + i] != ST_OK) continue;
      for (int z = 0; z < vp.nz; z++) {
        const int lo = zr[((size_t) i * NZMAX + z) * 2], hi = zr[((size_t) i * NZMAX + z) * 2 + 1];
        if (hi >= lo) prof += (hi - lo + 1) * 8.0;
      }
    }
    prof *= (double) b->n / (double) nc;
  }
  out8[4] = prof;
  out8[5] = out8[6] = out8[7] = 0.0;
  if (m.type == T_RELXILL) {   // values per zone spectrum as k_xill files them and k_conv reads them
    const XillHost &xh = E.tables->xill_host(model_xtab(m));
    out8[5] = b->xill_conv_grid ? xh.xc_n : xh.n_ener;
  }
  out8[6] = xbytes_union;
  return 0;
}

int relxill_b200_probe(relxill_b200_batch *b, long iv, const char *what, double *out, long max_len) {
  if (!b) return -1;
  Engine &E = *b->eng;
  std::lock_guard<std::mutex> lk(E.mu);
  cudaSetDevice(E.device);
  cudaDeviceSynchronize();
  if (iv < b->last_chunk0 || iv >= b->last_chunk0 + b->last_chunk_n) { set_err("probe: vector not in the last chunk"); return -1; }
  const size_t v = (size_t) (iv - b->last_chunk0);
  const Scratch S = scratch_slice(E.S, b->arena_c0, false);
  const VPar &vp = b->vps[iv];
  const std::string w = what;
  const double *src = nullptr;
  size_t n = 0;
  std::vector<double> tmp;
  if (w == "re") { src = S.re + v * NR; n = NR; }
  else if (w == "gmin") { src = S.gmin + v * NR; n = NR; }
  else if (w == "gmax") { src = S.gmax + v * NR; n = NR; }
  else if (w == "emis") { src = S.emis + v * NR; n = NR; }
  else if (w == "del_emit") { src = S.del_emit + v * NR; n = NR; }
  else if (w == "del_inc") { src = S.del_inc + v * NR; n = NR; }
  else if (w == "trff" || w == "cosne") {
    const double *base = (w == "trff") ? S.trff : S.cosne;
    if (!base) { set_err("probe: the fine transfer functions are only filed with relxill_b200_keep_intermediates(1)"); return -1; }
    src = base + v * NR * NG * 2; n = (size_t) NR * NG * 2;
  }
  else if (w == "reflfrac") { src = S.reflfrac + v * 8; n = 5; }
  else if (w == "lxi") { src = S.zlxi + v * NZMAX; n = vp.nz; }
  else if (w == "dens") { src = S.zdens + v * NZMAX; n = vp.nz; }
  else if (w == "ect") { src = S.zect + v * NZMAX; n = vp.nz; }
  else if (w == "eshift") { src = S.eshift + v * NZMAX; n = vp.nz; }
  else if (w == "normch") { src = S.normch + v * NZMAX; n = vp.nz; }
  else if (w == "corr_flux") { src = S.corr_flux + v * NZMAX; n = vp.nz; }
  else if (w == "corr_gshift") { src = S.corr_gshift + v * NZMAX; n = vp.nz; }
  else if (w == "total") { src = S.total + v * NCONV; n = NCONV; }
  else if (w == "xill_ener") {   // bin edges of the xillver table of this model
    const std::vector<double> &xe = E.tables->xill_host(model_xtab(*b->m)).ener;
    n = xe.size();
    if ((long) n > max_len) return -1;
    for (size_t i = 0; i < n; i++) out[i] = xe[i];
    return (int) n;
  } else if (w == "xillc") {       // zone spectra on the convolution grid (zero outside the table's range)
    const XillHost &xh = E.tables->xill_host(model_xtab(*b->m));
    if (!b->xill_conv_grid) { set_err("probe: zone spectra are on the table grid (probe \"xill\")"); return -1; }
    const int nz = vp.nz;
    if ((long) nz * NCONV > max_len) return -1;
    for (long i = 0; i < (long) nz * NCONV; i++) out[i] = 0.0;
    for (int z = 0; z < nz; z++)
      if (cudaMemcpy(out + (size_t) z * NCONV + xh.xc_first, S.xillz + (v * S.nz_cap + z) * xh.xc_stride,
                     xh.xc_n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    return nz * NCONV;
  } else if (w == "zone") {
    n = vp.nz + 1;
    if ((long) n > max_len) return -1;
    for (size_t i = 0; i < n; i++) out[i] = vp.zone[i];
    return (int) n;
  } else if (w == "relflux" || w == "xill" || w == "dist") {
    if (w == "xill" && b->xill_conv_grid) { set_err("probe: zone spectra are on the convolution grid (probe \"xillc\")"); return -1; }
    const int nz = vp.nz;
    const size_t len = (w == "relflux") ? (size_t) ((b->m->type == T_LINE) ? b->n_flux : NCONV)
                       : (w == "xill")  ? (size_t) E.tables->xill_host(model_xtab(*b->m)).n_ener
                                        : (size_t) E.tables->xill_host(model_xtab(*b->m)).n_incl;
    if ((long) (len * nz) > max_len) return -1;
    for (int z = 0; z < nz; z++) {
      const double *p = (w == "relflux") ? S.relflux + (v * S.nz_cap + z) * S.ne_line_cap
                        : (w == "xill")  ? S.xillz + (v * S.nz_cap + z) * E.tables->xill_host(model_xtab(*b->m)).stride
                                         : S.dist + (v * NZMAX + z) * MAX_INCL;
      if (cudaMemcpy(out + z * len, p, len * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
      if (w == "relflux") {  // rows are only written inside the zone's bin range
        int rg[2];
        if (cudaMemcpy(rg, S.zrange + (v * NZMAX + z) * 2, sizeof(rg), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
        for (long i = 0; i < (long) len; i++) if (i < rg[0] || i > rg[1]) out[z * len + i] = 0.0;
      }
    }
    return (int) (len * nz);
  } else {
    set_err("probe: unknown quantity " + w);
    return -1;
  }
  if ((long) n > max_len) return -1;
  if (cudaMemcpy(out, src, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  return (int) n;
}

// ---------------------------------------------------------------- XSPEC local-model entry points
// The reference prints its error and leaves flux unspecified when an evaluation fails (src/LocalModel.cpp:149-158); here
// flux is zeroed and EVERY failed call says so on stderr (a fit that silently received zeros would converge on nonsense).
static void lmod_call(const char *name, const double *energy, int Nflux, const double *parameter, double *flux) {
  int st = 0;
  const int rc = relxill_batch_eval(name, energy, Nflux, parameter, 1, flux, &st);
  if (rc != 0 || st != 0) {
    for (int i = 0; i < Nflux; i++) flux[i] = 0.0;
    const std::string why = rc != 0 ? g_err : (st == ST_BAD_PARAM ? std::string("parameters outside the allowed range")
                                                                  : "status " + std::to_string(st));
    fprintf(stderr, " *** relxill_b200: evaluation of %s failed (%s); returning zeros\n", name, why.c_str());
  }
}

#define DEF_LMOD(sym, name)                                                                              \
  void sym(const double *energy, int Nflux, const double *parameter, int spectrum, double *flux,        \
           double *fluxError, const char *init) {                                                        \
    (void) spectrum; (void) fluxError; (void) init;                                                      \
    lmod_call(name, energy, Nflux, parameter, flux);                                                     \
  }

DEF_LMOD(lmodrelline, "relline")
DEF_LMOD(lmodrelconv, "relconv")
DEF_LMOD(lmodrellinelp, "relline_lp")
DEF_LMOD(lmodrelconvlp, "relconv_lp")
DEF_LMOD(lmodrelxill, "relxill")
DEF_LMOD(lmodrelxilllp, "relxilllp")
DEF_LMOD(lmodrelxilldensnthcomp, "relxillCp")
DEF_LMOD(lmodrelxilllpdensnthcomp, "relxilllpCp")
DEF_LMOD(lmodxillver, "xillver")
DEF_LMOD(lmodxillverdensnthcomp, "xillverCp")
DEF_LMOD(lmodxillverns, "xillverNS")
DEF_LMOD(lmodrelxillns, "relxillNS")
DEF_LMOD(lmodxillverco, "xillverCO")
DEF_LMOD(lmodrelxillco, "relxillCO")

}  // extern "C"
