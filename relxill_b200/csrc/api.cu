// api.cu — runtime + C ABI of librelxill_b200.so (see include/relxill_b200.h).
//
// Engine: one per process (one process per GPU); owns the HBM-resident tables, a scratch arena
// sized for one chunk of parameter vectors, and the kernel sequence.  Batches larger than the
// chunk capacity are streamed through the arena chunk by chunk on the caller's stream.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
  std::mutex mu;
  bool inited = false;
  int device = 0;
  Tables *tables = nullptr;
  HostConfig cfg;
  Scratch S{};
  std::vector<void *> scratch_allocs;
  double *d_total = nullptr;  // [cap][NCONV]
  double *d_io = nullptr;     // staging for host-buffer calls
  size_t d_io_cap = 0;
  long max_chunk = 4096;
  long pipe_piece = 2072;     // vectors per pipelined piece of a host-buffer call (RELXILL_B200_PIPE)
  long pipe_last = 888;       // ... and of the last piece, whose device->host copy nothing overlaps (RELXILL_B200_PIPE_LAST)
  cudaStream_t stream_c = nullptr, stream_d = nullptr;   // compute / copy streams of host-buffer calls
  bool profiling = false;
  bool keep_intermediates = false;   // store what only the test probes read (emission-angle tables)
  // k_xill blends the convolution-grid copy of the table (rows rebinned once at load, xill.cu) and k_conv reads the zone
  // spectra with plain coalesced loads; off (RELXILL_B200_XILL_GRID=table): no such copy is built, zone spectra on the
  // table grid, rebinned per zone in k_conv
  bool xill_conv_grid = true;
  // device buffers recycled between batches (cudaMalloc/cudaFree synchronise and cost milliseconds)
  std::vector<std::pair<size_t, void *>> pool;
  // Device-resident state cache.  The scratch arena keeps the intermediates of the batch that
  // ran last in one piece (arena_owner = that batch's uid); when the same batch runs again with updated parameters,
  // vectors whose relativistic half / whole parameter set is unchanged re-use their rows instead of recomputing them.
  bool cache_on = true;
  unsigned long arena_owner = 0, next_uid = 1;
  relxill_b200_batch *retained = nullptr;   // the batch of the last host-buffer call (what an XSPEC fit re-evaluates)
};
Engine g_eng;

void *pool_get(Engine &E, size_t bytes) {
  for (size_t i = 0; i < E.pool.size(); i++) {
    if (E.pool[i].first >= bytes && E.pool[i].first <= 2 * bytes + 4096) {
      void *p = E.pool[i].second;
      E.pool.erase(E.pool.begin() + i);
      return p;
    }
  }
  void *d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
  return d;
}
void pool_put(Engine &E, void *p, size_t bytes) {
  if (!p) return;
  if (E.pool.size() >= 8) {
    cudaFree(E.pool.front().second);
    E.pool.erase(E.pool.begin());
  }
  E.pool.emplace_back(bytes, p);
}

void free_scratch(Engine &E) {
  E.arena_owner = 0;   // whatever state the arena held is gone
  for (void *p : E.scratch_allocs) cudaFree(p);
  E.scratch_allocs.clear();
  E.S = Scratch{};
  E.d_total = nullptr;
}

template <class T> bool salloc(Engine &E, T *&p, size_t n) {
  void *d = nullptr;
  if (cudaMalloc(&d, n * sizeof(T)) != cudaSuccess) return false;
  E.scratch_allocs.push_back(d);
  p = (T *) d;
  return true;
}

int ensure_scratch(Engine &E, long cap, int nz_cap, int ne_cap, int nex_stride, bool nth) {
  Scratch &S = E.S;
  if (S.cap >= cap && S.nz_cap >= nz_cap && S.ne_line_cap >= ne_cap && S.nex_stride >= nex_stride && (!nth || S.nth_spt)) return 0;
  nth = nth || S.nth_spt != nullptr;
  cap = std::max(cap, S.cap);
  nz_cap = std::max(nz_cap, S.nz_cap);
  ne_cap = std::max(ne_cap, S.ne_line_cap);
  nex_stride = std::max(nex_stride, S.nex_stride);
  free_scratch(E);
  bool ok = true;
  const size_t c = (size_t) cap;
  ok &= salloc(E, S.re, c * NR) && salloc(E, S.gmin, c * NR) && salloc(E, S.gmax, c * NR) && salloc(E, S.emis, c * NR);
  ok &= salloc(E, S.del_emit, c * NR) && salloc(E, S.del_inc, c * NR) && salloc(E, S.fr, c * NR);
  ok &= salloc(E, S.zfirst, c * (NZMAX + 1)) && salloc(E, S.brk_i, c * 2) && salloc(E, S.brk_f, c * 2);
  ok &= salloc(E, S.it, c * NR) && salloc(E, S.izone, c * NR) && salloc(E, S.glim, c * 2) && salloc(E, S.reflfrac, c * 8);
  ok &= salloc(E, S.trff, c * NR * NG * 2) && salloc(E, S.cosne, c * NR * NG * 2);
  ok &= salloc(E, S.relrow, c * REL_NRT * NG * 4);
  ok &= salloc(E, S.eshift, c * NZMAX) && salloc(E, S.zlxi, c * NZMAX) && salloc(E, S.zdens, c * NZMAX);
  ok &= salloc(E, S.zect, c * NZMAX) && salloc(E, S.normch, c * NZMAX) && salloc(E, S.corr_flux, c * NZMAX);
  ok &= salloc(E, S.corr_gshift, c * NZMAX) && salloc(E, S.nsrc, c);
  ok &= salloc(E, S.xrow, c * NZMAX * 32) && salloc(E, S.xw, c * NZMAX * 32);
  ok &= salloc(E, S.xkey, c * NZMAX * 32) && salloc(E, S.xwsort, c * NZMAX * 32) && salloc(E, S.xn, c);
  ok &= salloc(E, S.xga_off, c * 4) && salloc(E, S.xga_w, c * 4) && salloc(E, S.zrange, c * NZMAX * 2);
  ok &= salloc(E, S.relflux, c * nz_cap * ne_cap) && salloc(E, S.dist, c * NZMAX * MAX_INCL) && salloc(E, S.distpart, c * NR * 10);
  ok &= salloc(E, S.xillz, c * nz_cap * (size_t) std::max(nex_stride, 1)) && salloc(E, S.status, c);
  ok &= salloc(E, E.d_total, c * NCONV);
  if (nth) {
    ok &= salloc(E, S.nth_gam, c * NTH_MAX * NTH_SOL) && salloc(E, S.nth_g, c * NTH_MAX * NTH_SOL);
    ok &= salloc(E, S.nth_spt, c * NTH_MAX * NTH_SOL) && salloc(E, S.nth_jmax, c * NTH_SOL);
  }
  if (!ok) {
    free_scratch(E);
    set_err("out of device memory for the scratch arena");
    return -2;
  }
  S.cap = cap; S.nz_cap = nz_cap; S.ne_line_cap = ne_cap; S.nex_stride = nex_stride;
  return 0;
}

int engine_init(Engine &E, const char *dir, int device) {
  if (E.inited) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_err("no CUDA device available (this library has no CPU fallback)");
    return -1;
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  CK(cudaSetDevice(device));
  E.device = device;
  std::string d;
  if (dir && *dir) d = dir;
  else if (const char *env = getenv("RELXILL_TABLE_PATH")) d = env;  // src/relutility.c:320-328
  else d = "./";
  if (const char *env = getenv("RELXILL_NUM_RZONES")) E.cfg.env_num_zones = (int) atof(env);
  if (const char *env = getenv("RELXILL_RETURNRAD_SWITCH")) E.cfg.env_returnrad = (int) atof(env);
  if (const char *env = getenv("RELLINE_PHYSICAL_NORM")) E.cfg.env_phys_norm = ((int) strtod(env, nullptr) == 1) ? 1 : 0;
  if (const char *env = getenv("RELXILL_CONSTANT_DENSITY")) E.cfg.env_const_density = ((int) strtod(env, nullptr) == 1) ? 1 : 0;
  if (const char *env = getenv("RELXILL_B200_CHUNK")) E.max_chunk = std::max(1L, atol(env));
  if (const char *env = getenv("RELXILL_B200_PIPE")) E.pipe_piece = std::max(1L, atol(env));
  if (const char *env = getenv("RELXILL_B200_PIPE_LAST")) E.pipe_last = std::max(0L, atol(env));
  if (const char *env = getenv("RELXILL_B200_XILL_GRID")) E.xill_conv_grid = std::string(env) != "table";
  if (!E.stream_c) cudaStreamCreateWithFlags(&E.stream_c, cudaStreamNonBlocking);
  if (!E.stream_d) cudaStreamCreateWithFlags(&E.stream_d, cudaStreamNonBlocking);
  if (kernels_init() != 0) {
    set_err("kernel attribute setup failed (is this an sm_100a device?)");
    return -1;
  }
  E.tables = new Tables();
  E.tables->set_conv_grid_copy(E.xill_conv_grid);
  const std::string err = E.tables->load(d);
  if (!err.empty()) {
    set_err(err);
    delete E.tables;
    E.tables = nullptr;
    return -1;
  }
  E.inited = true;
  return 0;
}

}  // namespace

struct relxill_b200_batch {
  const ModelDef *m = nullptr;
  long n = 0;
  int n_flux = 0;
  int nz_max = 1;
  bool any_corr = false;
  bool any_limb = false;      // some vector uses a limb law: k_fine must keep the emission angles for k_line
  std::vector<VPar> vps;
  std::vector<int> status;
  VPar *d_vps = nullptr;
  double *d_energy = nullptr;
  size_t vps_bytes = 0, energy_bytes = 0;
  long launches = 0;
  long last_chunk0 = 0, last_chunk_n = 0;
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
  Engine &E;
  relxill_b200_batch *b;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  Timer(Engine &E_, relxill_b200_batch *b_, cudaStream_t s) : E(E_), b(b_), st(s) {
    if (E.profiling) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
  }
  ~Timer() {
    if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
  }
  void begin() { if (E.profiling) cudaEventRecord(e0, st); }
  void end(int fam, bool is_kernel = true) {
    if (is_kernel) b->launches++;
    b->kt_n[fam]++;
    if (E.profiling) {
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
// remainder first, so that only the copy of the last piece is exposed.
struct PipeOut {
  double *h_flux;
  cudaStream_t copy_stream;
};

int run_batch(Engine &E, relxill_b200_batch *b, double *d_flux, cudaStream_t st, const PipeOut *pipe = nullptr) {
  const ModelDef &m = *b->m;
  const DevTables &T = E.tables->dev();
  const int xtab = model_xtab(m);
  const int which = xtab < 0 ? 0 : xtab;
  const bool relxill = (m.type == T_RELXILL);
  const int ne_line = (m.type == T_LINE) ? b->n_flux : NCONV;
  const int nex_stride = (relxill || m.type == T_XILL) ? std::max(E.tables->xill_host(xtab).stride, E.tables->xill_host(xtab).xc_stride) : 1;
  const int cgrid = (relxill && E.xill_conv_grid && E.tables->xill_host(xtab).has_conv_copy) ? 1 : 0;
  b->xill_conv_grid = cgrid;
  const int n_incl = relxill ? E.tables->xill_host(xtab).n_incl : 0;
  const bool xillver = (m.type == T_XILL);
  const bool nth = (relxill || xillver) && m.prim == PRIM_NTHCOMP;
  // the Kompaneets work arrays take 1.4 MB per vector: smaller chunks for the Cp models
  const long cap = std::min(b->n, nth ? std::min<long>(E.max_chunk, 2048) : E.max_chunk);
  if (ensure_scratch(E, cap, b->nz_max, ne_line, nex_stride, nth)) return -2;
  Scratch S = E.S;
  b->launches = 0;
  for (int k = 0; k < KF_COUNT; k++) { b->kt_ms[k] = 0; b->kt_n[k] = 0; }
  Timer tm(E, b, st);
  const std::vector<double> &econv = E.tables->econv();
  std::vector<long> piece_n;
  {
    long piece = S.cap, last = 0;
    if (pipe && b->n >= E.pipe_piece + E.pipe_last) {
      piece = std::min(S.cap, E.pipe_piece);
      last = std::min(E.pipe_last, piece);   // only the copy of the last piece is exposed: keep that piece short
    }
    const long body = b->n - last;
    long first = body % piece;
    if (first == 0) first = std::min(piece, body);
    for (long c0 = 0, nc = first; c0 < body; c0 += nc, nc = std::min(piece, body - c0)) piece_n.push_back(nc);
    if (last > 0) piece_n.push_back(last);
  }
  // ---- state cache: which vectors can keep the rows the arena still holds for them
  const bool one_piece = piece_n.size() == 1;
  b->n_reuse_rel = b->n_reuse_all = 0;
  S.reuse = nullptr;
  if (E.cache_on && one_piece && b->state_valid && E.arena_owner == b->uid && b->d_reuse && !E.keep_intermediates &&
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
      S.reuse = b->d_reuse;
    }
  }
  b->state_valid = false;   // until this run has been enqueued completely
  E.arena_owner = 0;
  std::vector<cudaEvent_t> ev(pipe ? piece_n.size() : 0);
  for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  auto copy_piece = [&](size_t k, long c0, long nc) {
    CK(cudaStreamWaitEvent(pipe->copy_stream, ev[k], 0));
    CK(cudaMemcpyAsync(pipe->h_flux + (size_t) c0 * b->n_flux, d_flux + (size_t) c0 * b->n_flux,
                       (size_t) nc * b->n_flux * sizeof(double), cudaMemcpyDeviceToHost, pipe->copy_stream));
    return 0;
  };
  long c0 = 0, prev_c0 = 0;
  for (size_t ip = 0; ip < piece_n.size(); c0 += piece_n[ip], ip++) {
    const long nc = piece_n[ip];
    if (pipe && ip > 0) {   // the previous piece's kernels are enqueued behind it: its copy overlaps this piece
      if (copy_piece(ip - 1, prev_c0, piece_n[ip - 1])) return -2;
    }
    struct AtEnd {   // record the piece's completion event however the body is left
      std::vector<cudaEvent_t> &ev; size_t ip; cudaStream_t st; bool on; long &prev, c0;
      ~AtEnd() { if (on) cudaEventRecord(ev[ip], st); prev = c0; }
    } at_end{ev, ip, st, pipe != nullptr, prev_c0, c0};
    const VPar *vps = b->d_vps + c0;
    double *out = d_flux + (size_t) c0 * b->n_flux;
    if (xillver) {
      tm.begin(); launch_xillver(vps, T, S, nc, which, b->d_energy, b->n_flux, out, nex_stride, st); tm.end(KF_XILLVER);
      if (nth) {
        tm.begin(); launch_nth(vps, T, S, nc, st); tm.end(KF_NTH);
        tm.begin(); launch_xillver_prim_nth(vps, T, S, nc, b->d_energy, b->n_flux, out, st); tm.end(KF_PRIMNTH);
      }
      CK(cudaMemcpyAsync(b->status.data() + c0, S.status, nc * sizeof(int), cudaMemcpyDeviceToHost, st));
      b->last_chunk0 = c0;
      b->last_chunk_n = nc;
      continue;
    }
    tm.begin(); launch_syspar(vps, T, S, nc, 1, st); tm.end(KF_SYSPAR);
    if (relxill) {
      tm.begin(); launch_zone(vps, T, S, nc, st); tm.end(KF_ZONE);
      if (nth) { tm.begin(); launch_nth(vps, T, S, nc, st); tm.end(KF_NTH); }
      if (b->any_corr) { tm.begin(); launch_syspar(vps, T, S, nc, 2, st); tm.end(KF_SYSPAR); }
    }
    tm.begin();
    launch_fine(vps, T, S, nc, relxill ? n_incl : 0, econv[0], econv[NCONV], (b->any_limb || E.keep_intermediates) ? 1 : 0, st);
    tm.end(KF_FINE);
    b->launches++;   // launch_fine is two kernels (k_rows, k_fine) timed as one family
    if (relxill) {
      tm.begin(); launch_dist(vps, T, S, nc, n_incl, st); tm.end(KF_DIST);
    }
    if (m.type == T_LINE) {
      tm.begin(); launch_line(vps, T, S, nc, b->d_energy, b->n_flux, 1, 1, st); tm.end(KF_LINE);
      tm.begin(); launch_linefinish(vps, S, nc, b->n_flux, out, st); tm.end(KF_FINISH);
    } else {
      tm.begin(); launch_line(vps, T, S, nc, T.econv, NCONV, 0, relxill ? b->nz_max : 1, st); tm.end(KF_LINE);
      if (relxill) {
        tm.begin(); launch_xill(vps, T, S, nc, which, cgrid, st); tm.end(KF_XILL);
        tm.begin(); launch_conv(vps, T, S, nc, b->d_energy, b->n_flux, out, E.d_total, which, 0, cgrid, st); tm.end(KF_CONV);
        if (nth) {
          tm.begin(); launch_prim_nth(vps, T, S, nc, E.d_total, b->d_energy, b->n_flux, out, st); tm.end(KF_PRIMNTH);
        }
      } else {
        tm.begin(); launch_conv(vps, T, S, nc, b->d_energy, b->n_flux, out, nullptr, 0, 1, 0, st); tm.end(KF_CONV);
      }
    }
    CK(cudaMemcpyAsync(b->status.data() + c0, S.status, nc * sizeof(int), cudaMemcpyDeviceToHost, st));
    b->last_chunk0 = c0;
    b->last_chunk_n = nc;
  }
  if (pipe) {
    if (copy_piece(piece_n.size() - 1, prev_c0, piece_n.back())) return -2;
    CK(cudaStreamSynchronize(pipe->copy_stream));
    CK(cudaStreamSynchronize(st));
    for (auto &e : ev) cudaEventDestroy(e);
  }
  CK(cudaGetLastError());
  if (one_piece && !xillver) {   // the arena now holds this batch's state
    b->state_vps = b->vps;
    b->state_valid = true;
    E.arena_owner = b->uid;
  }
  return 0;
}

}  // namespace

namespace {

// host-side interpretation of the raw parameter vectors (spread over a few threads for large batches) and the
// batch-level switches derived from it
void interpret_all(Engine &E, relxill_b200_batch *b, const double *params) {
  const ModelDef *m = b->m;
  const long n_vec = b->n;
  const std::vector<double> &sp = E.tables->rr_spins();
  const int nthr = (int) std::max<long>(1, std::min<long>({(long) std::thread::hardware_concurrency(), 16L, n_vec / 256}));
  auto work = [&](long lo, long hi) {
    for (long i = lo; i < hi; i++)
      interpret_params(*m, params + (size_t) i * m->npar, E.cfg, sp.empty() ? nullptr : sp.data(), (int) sp.size(), b->vps[i]);
  };
  if (nthr <= 1) {
    work(0, n_vec);
  } else {
    std::vector<std::thread> th;
    for (int k = 0; k < nthr; k++) th.emplace_back(work, n_vec * k / nthr, n_vec * (k + 1) / nthr);
    for (auto &x : th) x.join();
  }
  b->nz_max = 1;
  b->any_corr = b->any_limb = false;
  for (long i = 0; i < n_vec; i++) {
    if (b->vps[i].status == ST_OK) {
      b->nz_max = std::max(b->nz_max, b->vps[i].nz);
      if (b->vps[i].do_corr) b->any_corr = true;
      if (b->vps[i].limb != 0) b->any_limb = true;
    }
  }
}

void read_call_env(Engine &E) {   // read per call, like the reference's constantDiskDensity() (src/relutility.c:372-382)
  const char *env = getenv("RELXILL_CONSTANT_DENSITY");
  E.cfg.env_const_density = (env && (int) strtod(env, nullptr) == 1) ? 1 : 0;
}

void free_batch_locked(Engine &E, relxill_b200_batch *b) {
  if (!b) return;
  if (E.arena_owner == b->uid) E.arena_owner = 0;
  pool_put(E, b->d_vps, b->vps_bytes);
  pool_put(E, b->d_energy, b->energy_bytes);
  pool_put(E, b->d_reuse, (size_t) b->n);
  delete b;
}

}  // namespace

// =================================================================================== C ABI
extern "C" {

const char *relxill_b200_last_error(void) { return g_err.c_str(); }

int relxill_b200_init(const char *table_dir, int device) {
  std::lock_guard<std::mutex> lk(g_eng.mu);
  return engine_init(g_eng, table_dir, device);
}

void relxill_b200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_eng.mu);
  free_scratch(g_eng);
  if (g_eng.d_io) cudaFree(g_eng.d_io);
  g_eng.d_io = nullptr;
  g_eng.d_io_cap = 0;
  free_batch_locked(g_eng, g_eng.retained);
  g_eng.retained = nullptr;
  for (auto &pr : g_eng.pool) cudaFree(pr.second);
  g_eng.pool.clear();
  delete g_eng.tables;
  g_eng.tables = nullptr;
  g_eng.inited = false;
}

void relxill_b200_set_num_zones(int n) { g_eng.cfg.env_num_zones = n; }
void relxill_b200_set_profiling(int on) { g_eng.profiling = on != 0; }
void relxill_b200_keep_intermediates(int on) { g_eng.keep_intermediates = on != 0; }
void relxill_b200_set_cache(int on) { g_eng.cache_on = on != 0; }
void relxill_b200_set_xill_grid(int conv_grid) { g_eng.xill_conv_grid = conv_grid != 0; }
int relxill_b200_get_xill_grid(void) { return g_eng.xill_conv_grid ? 1 : 0; }
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

relxill_b200_batch *relxill_b200_prepare(const char *model, const double *energy, int n_flux, const double *params,
                                         long n_vec) {
  Engine &E = g_eng;
  std::lock_guard<std::mutex> lk(E.mu);
  g_err.clear();
  if (engine_init(E, nullptr, -1)) return nullptr;
  const ModelDef *m = find_model(model);
  if (!m) { set_err(std::string("unknown model ") + model); return nullptr; }
  if (n_vec < 1 || n_flux < 1) { set_err("empty batch or energy grid"); return nullptr; }
  if (m->type == T_LINE && n_flux > line_max_bins()) {
    set_err("line models: energy grids above " + std::to_string(line_max_bins()) + " bins are not supported yet");
    return nullptr;
  }
  // tables this flavour needs; returning radiation can be switched per vector -> load if the table exists
  read_call_env(E);
  bool want_rr = (m->irrad == EMIS_LP) || E.cfg.env_returnrad == 1;
  std::string err = (m->type == T_XILL)
                        ? E.tables->require_xill_only(model_xtab(*m))
                        : E.tables->require(m->irrad == EMIS_LP, false, model_xtab(*m));
  if (!err.empty()) { set_err(err); return nullptr; }
  if (want_rr) {
    err = E.tables->require(false, true, XT_NONE);
    // missing table is only an error for the vectors that switch returning radiation on
  }
  auto *b = new relxill_b200_batch();
  b->m = m;
  b->n = n_vec;
  b->n_flux = n_flux;
  b->vps.resize(n_vec);
  b->status.assign(n_vec, 0);
  b->uid = E.next_uid++;
  b->energy.assign(energy, energy + n_flux + 1);
  interpret_all(E, b, params);
  b->vps_bytes = n_vec * sizeof(VPar);
  b->energy_bytes = (n_flux + 1) * sizeof(double);
  b->d_vps = (VPar *) pool_get(E, b->vps_bytes);
  b->d_energy = (double *) pool_get(E, b->energy_bytes);
  b->d_reuse = (unsigned char *) pool_get(E, (size_t) n_vec);
  if (!b->d_vps || !b->d_energy || !b->d_reuse) {
    set_err("out of device memory (batch)");
    free_batch_locked(E, b);
    return nullptr;
  }
  cudaMemcpy(b->d_vps, b->vps.data(), n_vec * sizeof(VPar), cudaMemcpyHostToDevice);
  cudaMemcpy(b->d_energy, energy, (n_flux + 1) * sizeof(double), cudaMemcpyHostToDevice);
  return b;
}

void relxill_b200_free_batch(relxill_b200_batch *b) {
  if (!b) return;
  std::lock_guard<std::mutex> lk(g_eng.mu);
  free_batch_locked(g_eng, b);
}

int relxill_b200_update_params(relxill_b200_batch *b, const double *params) {
  Engine &E = g_eng;
  std::lock_guard<std::mutex> lk(E.mu);
  if (!b || !E.inited || !params) { set_err("update_params: library not initialised or null argument"); return -1; }
  read_call_env(E);
  interpret_all(E, b, params);
  CK(cudaMemcpy(b->d_vps, b->vps.data(), b->n * sizeof(VPar), cudaMemcpyHostToDevice));
  return 0;
}

int relxill_b200_update_energy(relxill_b200_batch *b, const double *energy, int n_flux) {
  Engine &E = g_eng;
  std::lock_guard<std::mutex> lk(E.mu);
  if (!b || !E.inited || !energy || n_flux < 1) { set_err("update_energy: library not initialised or bad argument"); return -1; }
  if (b->m->type == T_LINE) {
    if (n_flux > line_max_bins()) { set_err("line models: energy grid too long"); return -1; }
    b->state_valid = false;   // the line models integrate on the caller's grid: nothing survives a new grid
  }
  const size_t bytes = (size_t) (n_flux + 1) * sizeof(double);
  if (bytes > b->energy_bytes) {
    double *d = (double *) pool_get(E, bytes);
    if (!d) { set_err("out of device memory (energy grid)"); return -2; }
    pool_put(E, b->d_energy, b->energy_bytes);
    b->d_energy = d;
    b->energy_bytes = bytes;
  }
  b->n_flux = n_flux;
  b->energy.assign(energy, energy + n_flux + 1);
  CK(cudaMemcpy(b->d_energy, energy, bytes, cudaMemcpyHostToDevice));
  return 0;
}

int relxill_b200_reuse_counts(relxill_b200_batch *b, long *out3) {
  if (!b || !out3) return -1;
  out3[0] = b->n - b->n_reuse_rel - b->n_reuse_all;
  out3[1] = b->n_reuse_rel;
  out3[2] = b->n_reuse_all;
  return 0;
}

int relxill_b200_run(relxill_b200_batch *b, double *d_flux, void *stream) {
  Engine &E = g_eng;
  std::lock_guard<std::mutex> lk(E.mu);
  if (!b || !E.inited) { set_err("run: library not initialised or null batch"); return -1; }
  return run_batch(E, b, d_flux, (cudaStream_t) stream);
}

int relxill_b200_batch_status(relxill_b200_batch *b, int *status) {
  if (!b) return -1;
  cudaDeviceSynchronize();
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
  cudaStreamSynchronize((cudaStream_t) stream);
  if (status) for (long i = 0; i < n_vec; i++) status[i] = b->status[i];
  relxill_b200_free_batch(b);
  return rc;
}

int relxill_batch_eval(const char *model, const double *energy, int n_flux, const double *params, long n_vec,
                       double *flux, int *status) {
  static const bool dbg = getenv("RELXILL_B200_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  // An XSPEC fit (or any caller that re-evaluates the same vectors with a few parameters changed) comes back with the
  // same model and batch size: the batch of the previous call was kept, so that the vectors whose relativistic half
  // or whole parameter set is unchanged re-use the state still resident in the arena (run_batch).
  relxill_b200_batch *b = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_eng.mu);
    relxill_b200_batch *r = g_eng.retained;
    g_eng.retained = nullptr;
    if (r && g_eng.cache_on && g_eng.inited && r->m == find_model(model) && r->n == n_vec && n_flux >= 1 && energy && params) b = r;
    else free_batch_locked(g_eng, r);
  }
  if (b) {
    int rc = 0;
    if (n_flux != b->n_flux || memcmp(energy, b->energy.data(), sizeof(double) * (size_t) (n_flux + 1)) != 0)
      rc = relxill_b200_update_energy(b, energy, n_flux);
    if (rc == 0) rc = relxill_b200_update_params(b, params);
    if (rc != 0) { relxill_b200_free_batch(b); b = nullptr; }
  }
  if (!b) b = relxill_b200_prepare(model, energy, n_flux, params, n_vec);
  const auto t1 = now();
  if (!b) {
    if (flux && n_vec > 0 && n_flux > 0) memset(flux, 0, sizeof(double) * (size_t) n_vec * n_flux);
    if (status) for (long i = 0; i < n_vec; i++) status[i] = ST_BAD_PARAM;
    return -1;
  }
  Engine &E = g_eng;
  const size_t need = (size_t) n_vec * n_flux;
  bool io_ok = true;
  {
    std::lock_guard<std::mutex> lk(E.mu);
    if (E.d_io_cap < need) {
      if (E.d_io) cudaFree(E.d_io);
      E.d_io = nullptr;
      E.d_io_cap = 0;
      if (cudaMalloc((void **) &E.d_io, need * sizeof(double)) != cudaSuccess) io_ok = false;
      else E.d_io_cap = need;
    }
  }
  if (!io_ok) {
    set_err("out of device memory (output staging)");
    relxill_b200_free_batch(b);
    return -2;
  }
  if (b->m->type == T_CONV) {
    // convolution models: flux is the input spectrum; a non-positive total is rejected (src/LocalModel.cpp:84-86)
    for (long i = 0; i < n_vec; i++) {
      double s = 0.0;
      for (int j = 0; j < n_flux; j++) s += flux[(size_t) i * n_flux + j];
      if (s <= 0.0 && b->vps[i].status == ST_OK) b->vps[i].status = ST_CONV_INPUT;
    }
    cudaMemcpy(b->d_vps, b->vps.data(), n_vec * sizeof(VPar), cudaMemcpyHostToDevice);
    cudaMemcpy(E.d_io, flux, need * sizeof(double), cudaMemcpyHostToDevice);
  }
  const auto t2 = now();
  int rc;
  {
    std::lock_guard<std::mutex> lk(E.mu);
    PipeOut pipe{flux, E.stream_d};
    rc = run_batch(E, b, E.d_io, E.stream_c, &pipe);   // kernels + pipelined D2H; returns with both streams drained
    if (rc != 0) { cudaStreamSynchronize(E.stream_c); cudaStreamSynchronize(E.stream_d); }
  }
  const auto t3 = now();
  const auto t4 = now();
  if (status) for (long i = 0; i < n_vec; i++) status[i] = b->status[i];
  if (rc == 0 && b->state_valid && E.cache_on) {
    std::lock_guard<std::mutex> lk(E.mu);
    E.retained = b;
  } else {
    relxill_b200_free_batch(b);
  }
  if (dbg)
    fprintf(stderr, "relxill_batch_eval timing: prepare %.2f ms, staging %.2f ms, run + D2H (pipelined) %.2f ms (+%.2f), free %.2f ms\n",
            ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, now()));
  return rc;
}

int relxill_b200_algorithmic_bytes(relxill_b200_batch *b, double *out8) {
  double *out4 = out8;
  Engine &E = g_eng;
  if (!b || !E.inited) return -1;
  cudaDeviceSynchronize();
  const ModelDef &m = *b->m;
  const long nc = b->last_chunk_n;
  double sumU = 0, bound = 0;
  double xbytes = 0;
  if (m.type == T_RELXILL && nc > 0) {
    const XillHost &xh = E.tables->xill_host(model_xtab(m));
    const int ncorn = (xh.npar == 6) ? 32 : 16;
    std::vector<int> rows((size_t) nc * NZMAX * 32);
    cudaMemcpy(rows.data(), E.S.xrow, rows.size() * sizeof(int), cudaMemcpyDeviceToHost);
    const double row_bytes = (double) xh.n_incl * xh.n_ener * 4.0;
    for (long i = 0; i < nc; i++) {
      const VPar &vp = b->vps[b->last_chunk0 + i];
      if (b->status[b->last_chunk0 + i] != ST_OK) continue;
      std::set<int> u;
      for (int z = 0; z < vp.nz; z++)
        for (int c = 0; c < ncorn; c++) u.insert(rows[((size_t) i * NZMAX + z) * 32 + c]);
      sumU += (double) u.size();
      bound += (double) vp.nz * ncorn * row_bytes;
    }
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
    cudaMemcpy(zr.data(), E.S.zrange, zr.size() * sizeof(int), cudaMemcpyDeviceToHost);
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
  return 0;
}

int relxill_b200_probe(relxill_b200_batch *b, long iv, const char *what, double *out, long max_len) {
  Engine &E = g_eng;
  if (!b || !E.inited) return -1;
  cudaDeviceSynchronize();
  if (iv < b->last_chunk0 || iv >= b->last_chunk0 + b->last_chunk_n) { set_err("probe: vector not in the last chunk"); return -1; }
  const size_t v = (size_t) (iv - b->last_chunk0);
  const Scratch &S = E.S;
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
  else if (w == "trff") { src = S.trff + v * NR * NG * 2; n = (size_t) NR * NG * 2; }
  else if (w == "cosne") { src = S.cosne + v * NR * NG * 2; n = (size_t) NR * NG * 2; }
  else if (w == "reflfrac") { src = S.reflfrac + v * 8; n = 5; }
  else if (w == "lxi") { src = S.zlxi + v * NZMAX; n = vp.nz; }
  else if (w == "dens") { src = S.zdens + v * NZMAX; n = vp.nz; }
  else if (w == "ect") { src = S.zect + v * NZMAX; n = vp.nz; }
  else if (w == "eshift") { src = S.eshift + v * NZMAX; n = vp.nz; }
  else if (w == "normch") { src = S.normch + v * NZMAX; n = vp.nz; }
  else if (w == "corr_flux") { src = S.corr_flux + v * NZMAX; n = vp.nz; }
  else if (w == "corr_gshift") { src = S.corr_gshift + v * NZMAX; n = vp.nz; }
  else if (w == "total") { src = E.d_total + v * NCONV; n = NCONV; }
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
      cudaMemcpy(out + (size_t) z * NCONV + xh.xc_first, S.xillz + (v * S.nz_cap + z) * xh.xc_stride,
                 xh.xc_n * sizeof(double), cudaMemcpyDeviceToHost);
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
      cudaMemcpy(out + z * len, p, len * sizeof(double), cudaMemcpyDeviceToHost);
      if (w == "relflux") {  // rows are only written inside the zone's bin range
        int rg[2];
        cudaMemcpy(rg, S.zrange + (v * NZMAX + z) * 2, sizeof(rg), cudaMemcpyDeviceToHost);
        for (long i = 0; i < (long) len; i++) if (i < rg[0] || i > rg[1]) out[z * len + i] = 0.0;
      }
    }
    return (int) (len * nz);
  } else {
    set_err("probe: unknown quantity " + w);
    return -1;
  }
  if ((long) n > max_len) return -1;
  cudaMemcpy(out, src, n * sizeof(double), cudaMemcpyDeviceToHost);
  return (int) n;
}

// ---------------------------------------------------------------- XSPEC local-model entry points
static void lmod_call(const char *name, const double *energy, int Nflux, const double *parameter, double *flux) {
  static bool warned = false;
  int st = 0;
  const int rc = relxill_batch_eval(name, energy, Nflux, parameter, 1, flux, &st);
  if (rc != 0 || st != 0) {
    for (int i = 0; i < Nflux; i++) flux[i] = 0.0;
    if (!warned) {
      fprintf(stderr, " *** relxill_b200: evaluation of %s failed (rc=%d, status=%d); returning zeros\n", name, rc, st);
      warned = true;
    }
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
