// capi.cxx -- plain-C entry points over the C++ host layer (ppxh_*), so that tests/ and bench.py can drive the same
// drivers a C++ caller links against (alsCP_DT, alsCP_PP, CPD<>::als, hosvd, alsTucker_*) through ctypes.
// Handles are opaque pointers to World / Tensor<> / Matrix<> / CPD objects; data crosses as host double arrays in
// the reference's global (first-index-fastest) order.  Errors: every call returns 0 or -1 (ppxh_last_error()).
#include <string>
#include "als_CP.h"
#include "als_Tucker.h"
#include "src/CP.h"
#include "src/Tucker.h"
#include "src/optimizer/cp_dt_optimizer.h"
#include "src/optimizer/cp_msdt_optimizer.h"
#include "src/optimizer/cp_dt_lr_optimizer.h"
#include "src/optimizer/cp_msdt_lr_optimizer.h"
#include "src/optimizer/cp_simple_optimizer.h"

namespace {
std::string g_err;
TraceSink g_sink;

template <typename F>
int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// the drivers take contiguous Matrix<> arrays; handles are separate objects: move in, run, move back
struct MatArray {
  std::vector<Matrix<>> v;
  void **h;
  int n;
  MatArray(void **handles, int n_) : h(handles), n(handles ? n_ : 0) {
    v.resize(n);
    for (int i = 0; i < n; i++) v[i] = std::move(*(Matrix<> *)h[i]);
  }
  ~MatArray() {
    for (int i = 0; i < n; i++) *(Matrix<> *)h[i] = std::move(v[i]);
  }
  Matrix<> *ptr() { return n ? v.data() : nullptr; }
};

struct CpdBase {
  virtual ~CpdBase() {}
  virtual void init(Tensor<> *V, Matrix<> *W, double lambda, uint64_t grad_seed) = 0;
  virtual double step() = 0;
  virtual bool als(double tol, double timelimit, int maxsweep, int resprint, ofstream &f, bool bench) = 0;
  virtual Matrix<> *W() = 0;
  virtual Matrix<> *grad() = 0;
};
template <class Opt>
struct CpdHolder : CpdBase {
  CPD<double, Opt> cpd;
  CpdHolder(int order, int size, int r, World &dw) : cpd(order, size, r, dw) {}
  CpdHolder(int order, int size, int r, int update_rank, int randomsvd, World &dw)
      : cpd(order, size, r, update_rank, randomsvd, dw) {}
  void init(Tensor<> *V, Matrix<> *W, double lambda, uint64_t grad_seed) override {
    cpd.Init(V, W, lambda);
    for (int i = 0; i < cpd.order; i++) cpd.grad_W[i].fill_random(0, 1, grad_seed, (uint64_t)i);
  }
  double step() override { return cpd.optimizer->step(); }
  bool als(double tol, double timelimit, int maxsweep, int resprint, ofstream &f, bool bench) override {
    return cpd.als(tol, timelimit, maxsweep, resprint, f, bench);
  }
  Matrix<> *W() override { return cpd.W; }
  Matrix<> *grad() override { return cpd.grad_W; }
};
}  // namespace

extern "C" {

const char *ppxh_last_error() { return g_err.c_str(); }

void *ppxh_world_create(int device, int solver, int use_graph, size_t workspace_bytes) {
  World *w = nullptr;
  if (guarded([&] {
        w = new World(device, workspace_bytes ? workspace_bytes : ((size_t)1 << 30));
        w->solver = solver;
        w->use_graph = use_graph != 0;
      }))
    return nullptr;
  return w;
}
void ppxh_world_destroy(void *w) { delete (World *)w; }
void *ppxh_world_ctx(void *w) { return ((World *)w)->ctx; }
int ppxh_world_set(void *w, int solver, int use_graph) {
  ((World *)w)->solver = solver;
  ((World *)w)->use_graph = use_graph != 0;
  return 0;
}
int ppxh_world_set_fast_residual(void *w, int on) {
  ((World *)w)->fast_residual = on != 0;
  return 0;
}
// multi-GPU: NCCL communicator + the shard layout of mode `shard_mode`
int ppxh_world_comm_init(void *w_, const void *id128, int nranks, int rank, int shard_mode, int64_t shard_global,
                         int64_t row_begin, int64_t row_end) {
  World *w = (World *)w_;
  return guarded([&] {
    PPXCK(*w, ppx_comm_init(w->ctx, id128, nranks, rank));
    w->np = nranks;
    w->rank = rank;
    w->shard_mode = shard_mode;
    w->shard_global = shard_global;
    w->row_begin = row_begin;
    w->row_end = row_end;
  });
}
int ppxh_world_set_shard(void *w_, int shard_mode, int64_t shard_global, int64_t row_begin, int64_t row_end) {
  World *w = (World *)w_;
  w->shard_mode = shard_mode;
  w->shard_global = shard_global;
  w->row_begin = row_begin;
  w->row_end = row_end;
  return 0;
}
void ppxh_world_trim(void *w) { ((World *)w)->trim(); }

void *ppxh_tensor_create(void *w, int order, const int64_t *lens) {
  Tensor<> *t = nullptr;
  if (guarded([&] { t = new Tensor<>(order, lens, *(World *)w); })) return nullptr;
  return t;
}
void *ppxh_matrix_create(void *w, int64_t nrow, int64_t ncol) {
  Matrix<> *m = nullptr;
  if (guarded([&] { m = new Matrix<>(nrow, ncol, *(World *)w); })) return nullptr;
  return m;
}
void ppxh_tensor_destroy(void *t) { delete (Tensor<> *)t; }
int ppxh_tensor_write(void *t, const double *host) {
  return guarded([&] { ((Tensor<> *)t)->write_all(host); });
}
int ppxh_tensor_read(void *t, double *host) {
  return guarded([&] { ((Tensor<> *)t)->read_all(host); });
}
// the same without the stream synchronise: `host` must be pinned and stay valid until ppxh_world_sync (stream order
// makes a write visible to every later kernel, and a read complete after the sync)
int ppxh_tensor_write_async(void *t_, const double *host) {
  Tensor<> *t = (Tensor<> *)t_;
  return guarded([&] { PPXCK(*t->wrld, ppx_memcpy_h2d(t->wrld->ctx, t->data, host, sizeof(double) * t->size)); });
}
int ppxh_tensor_read_async(void *t_, double *host) {
  Tensor<> *t = (Tensor<> *)t_;
  return guarded([&] { PPXCK(*t->wrld, ppx_memcpy_d2h(t->wrld->ctx, host, t->data, sizeof(double) * t->size)); });
}
int ppxh_world_sync(void *w) {
  return guarded([&] { ((World *)w)->sync(); });
}
void *ppxh_tensor_data(void *t) { return ((Tensor<> *)t)->data; }
int64_t ppxh_tensor_size(void *t) { return ((Tensor<> *)t)->size; }
int ppxh_tensor_order(void *t) { return ((Tensor<> *)t)->order; }
int64_t ppxh_tensor_len(void *t, int i) { return ((Tensor<> *)t)->lens[i]; }
double ppxh_tensor_norm2(void *t) {
  double v = -1;
  guarded([&] { v = ((Tensor<> *)t)->norm2(); });
  return v;
}
// out[i] = lo + (hi-lo) u(seed, id, start + i)
int ppxh_tensor_fill(void *t_, uint64_t seed, uint64_t id, double lo, double hi, int64_t start) {
  Tensor<> *t = (Tensor<> *)t_;
  return guarded([&] { PPXCK(*t->wrld, ppx_fill_uniform(t->wrld->ctx, t->data, t->size, seed, id, start, lo, hi)); });
}
// V = [[W_0..W_{N-1}]]  (tensor 'r', test_ALS.cxx:275-286)
int ppxh_build_V(void *V, void **W, int N, void *w) {
  return guarded([&] {
    MatArray Wa(W, N);
    build_V(*(Tensor<> *)V, Wa.ptr(), N, *(World *)w);
  });
}
double ppxh_cp_residual(void *V, void **W, int N, void *w) {
  double v = -1;
  guarded([&] {
    MatArray Wa(W, N);
    v = cp_residual_norm(*(Tensor<> *)V, Wa.ptr(), N, *(World *)w);
  });
  return v;
}

// ---- trace ---------------------------------------------------------------------------------------------------
void ppxh_trace_begin(int quiet, int skip_residual) {
  g_sink = TraceSink();
  g_sink.quiet = quiet != 0;
  g_sink.skip_residual = skip_residual != 0;
  trace_sink() = &g_sink;
}
void ppxh_trace_end() { trace_sink() = nullptr; }
int ppxh_trace_rows(double *out, int max_rows) {
  int n = (int)g_sink.rows.size();
  for (int i = 0; i < n && i < max_rows; i++) {
    out[5 * i + 0] = g_sink.rows[i].iter;
    out[5 * i + 1] = g_sink.rows[i].gradnorm;
    out[5 * i + 2] = g_sink.rows[i].pp_update;
    out[5 * i + 3] = g_sink.rows[i].diffV;
    out[5 * i + 4] = g_sink.rows[i].dtime;
  }
  return n;
}
int ppxh_trace_events(int *out, int max_n) {
  int n = (int)g_sink.events.size();
  for (int i = 0; i < n && i < max_n; i++) {
    out[2 * i] = g_sink.events[i].first;
    out[2 * i + 1] = g_sink.events[i].second;
  }
  return n;
}
int ppxh_trace_sweeps(int *out, int max_n) {
  int n = (int)g_sink.sweeps.size();
  for (int i = 0; i < n && i < max_n; i++) {
    out[2 * i] = g_sink.sweeps[i].first;
    out[2 * i + 1] = g_sink.sweeps[i].second;
  }
  return n;
}
int ppxh_trace_bench(double *out, int max_n) {
  int n = (int)g_sink.bench_times.size();
  for (int i = 0; i < n && i < max_n; i++) out[i] = g_sink.bench_times[i];
  return n;
}

// ---- CP drivers (als_CP.h) -------------------------------------------------------------------------------------
// *stopped = the bool the reference returns ("stopped before maxiter+1")
int ppxh_alsCP(void *V, void **W, void **grad_W, void **F, int N, double tol, double timelimit, int maxiter, void *w,
               int *stopped) {
  return guarded([&] {
    MatArray Wa(W, N), Ga(grad_W, N), Fa(F, N);
    *stopped = alsCP(*(Tensor<> *)V, Wa.ptr(), Ga.ptr(), Fa.ptr(), tol, timelimit, maxiter, *(World *)w);
  });
}
int ppxh_alsCP_DT(void *V, void **W, void **grad_W, void **F, int N, double tol, double timelimit, int maxiter,
                  double lambda, const char *csv, int resprint, int bench, void *w, int *stopped) {
  return guarded([&] {
    MatArray Wa(W, N), Ga(grad_W, N), Fa(F, N);
    ofstream f;
    if (csv && csv[0]) f.open(csv, bench ? std::ios::app : std::ios::out);
    *stopped = alsCP_DT(*(Tensor<> *)V, Wa.ptr(), Ga.ptr(), Fa.ptr(), tol, timelimit, maxiter, lambda, f, resprint,
                        bench != 0, *(World *)w);
  });
}
int ppxh_alsCP_PP(void *V, void **W, void **grad_W, void **F, int N, double tol, double tol_init, double timelimit,
                  int maxiter, double lambda, double ratio_step, const char *csv, int resprint, int bench, void *w,
                  int *stopped) {
  return guarded([&] {
    MatArray Wa(W, N), Ga(grad_W, N), Fa(F, N);
    ofstream f;
    if (csv && csv[0]) f.open(csv, bench ? std::ios::app : std::ios::out);
    *stopped = alsCP_PP(*(Tensor<> *)V, Wa.ptr(), Ga.ptr(), Fa.ptr(), tol, tol_init, timelimit, maxiter, lambda,
                        ratio_step, f, resprint, bench != 0, *(World *)w);
  });
}
int ppxh_alsCP_PP_partupdate(void *V, void **W, void **grad_W, void **F, int N, double tol, double tol_init,
                             double timelimit, int maxiter, double lambda, double ratio_step,
                             double update_percentage, const char *csv, int resprint, int bench, void *w,
                             int *stopped) {
  return guarded([&] {
    MatArray Wa(W, N), Ga(grad_W, N), Fa(F, N);
    ofstream f;
    if (csv && csv[0]) f.open(csv, bench ? std::ios::app : std::ios::out);
    *stopped = alsCP_PP_partupdate(*(Tensor<> *)V, Wa.ptr(), Ga.ptr(), Fa.ptr(), tol, tol_init, timelimit, maxiter,
                                   lambda, ratio_step, update_percentage, f, resprint, bench != 0, *(World *)w);
  });
}

// ---- timing helpers (bench.py) ---------------------------------------------------------------------------------
int ppxh_cp_dt_sweeps(void *V, void **W, void **grad_W, int N, int n_sweeps, double lambda, void *w) {
  return guarded([&] {
    MatArray Wa(W, N), Ga(grad_W, N);
    alsCP_DT_sweeps(*(Tensor<> *)V, Wa.ptr(), Ga.ptr(), n_sweeps, lambda, *(World *)w);
  });
}
int ppxh_cp_pp_phase_timed(void *V, void **W, void **grad_W, int N, int n_sweeps, double lambda, double ratio_step,
                           void *w, float *ms_build, float *ms_sweeps) {
  return guarded([&] {
    MatArray Wa(W, N), Ga(grad_W, N);
    alsCP_PP_phase_timed(*(Tensor<> *)V, Wa.ptr(), Ga.ptr(), n_sweeps, lambda, ratio_step, *(World *)w, ms_build,
                         ms_sweeps);
  });
}

// ---- probes at fixed factors (full-size parity checks) -----------------------------------------------------------
int ppxh_cp_dt_mttkrps(void *V, void **W, void **M, int N, void *w) {
  return guarded([&] {
    MatArray Wa(W, N), Ma(M, N);
    alsCP_DT_mttkrps(*(Tensor<> *)V, Wa.ptr(), Ma.ptr(), *(World *)w);
  });
}
// builds every PP operator at W; returns a handle to the map (NULL on error)
void *ppxh_cp_pp_ops_build(void *V, void **W, int N, void *w) {
  map<string, Tensor<>> *ops = new map<string, Tensor<>>();
  if (guarded([&] {
        MatArray Wa(W, N);
        alsCP_PP_operators(*(Tensor<> *)V, Wa.ptr(), *ops, *(World *)w);
      })) {
    delete ops;
    return nullptr;
  }
  return ops;
}
int64_t ppxh_cp_pp_ops_size(void *h, const char *key) {
  auto *ops = (map<string, Tensor<>> *)h;
  auto it = ops->find(key);
  return it == ops->end() ? -1 : it->second.size;
}
int ppxh_cp_pp_ops_read(void *h, const char *key, double *host) {
  return guarded([&] {
    auto *ops = (map<string, Tensor<>> *)h;
    auto it = ops->find(key);
    if (it == ops->end()) throw std::runtime_error(std::string("no PP operator with key ") + key);
    it->second.read_all(host);
  });
}
void ppxh_cp_pp_ops_free(void *h) { delete (map<string, Tensor<>> *)h; }

// ---- OO path (src/CP.h + src/optimizer) ---------------------------------------------------------------------
// kind: 3 = CPDTLROptimizer, 4 = CPMSDTLROptimizer  (run.cxx -pp 2 / 3)
void *ppxh_cpd_create_lr(int kind, int order, int size, int r, int update_rank, int randomsvd, void *w) {
  CpdBase *c = nullptr;
  if (guarded([&] {
        World &dw = *(World *)w;
        if (kind == 3) c = new CpdHolder<CPDTLROptimizer<double>>(order, size, r, update_rank, randomsvd, dw);
        else if (kind == 4) c = new CpdHolder<CPMSDTLROptimizer<double>>(order, size, r, update_rank, randomsvd, dw);
        else throw std::runtime_error("unknown low-rank optimizer kind");
      }))
    return nullptr;
  return c;
}
// kind: 0 = CPSimpleOptimizer, 1 = CPDTOptimizer, 2 = CPMSDTOptimizer  (run.cxx -pp 4 / 0 / 1)
void *ppxh_cpd_create(int kind, int order, int size, int r, void *w) {
  CpdBase *c = nullptr;
  if (guarded([&] {
        World &dw = *(World *)w;
        if (kind == 0) c = new CpdHolder<CPSimpleOptimizer<double>>(order, size, r, dw);
        else if (kind == 1) c = new CpdHolder<CPDTOptimizer<double>>(order, size, r, dw);
        else if (kind == 2) c = new CpdHolder<CPMSDTOptimizer<double>>(order, size, r, dw);
        else throw std::runtime_error("unknown optimizer kind");
      }))
    return nullptr;
  return c;
}
void ppxh_cpd_destroy(void *c) { delete (CpdBase *)c; }
// V stays owned by the caller (the reference adopts it and never frees it); W handles are COPIED into an array the
// CPD owns (Decomposition::Init adopts and delete[]s it)
int ppxh_cpd_init(void *c, void *V, void **W, int N, double lambda, uint64_t grad_seed) {
  return guarded([&] {
    Matrix<> *arr = new Matrix<>[N];
    for (int i = 0; i < N; i++) arr[i] = *(Matrix<> *)W[i];
    ((CpdBase *)c)->init((Tensor<> *)V, arr, lambda, grad_seed);
  });
}
int ppxh_cpd_step(void *c, double *sweep_fraction) {
  return guarded([&] { *sweep_fraction = ((CpdBase *)c)->step(); });
}
int ppxh_cpd_als(void *c, double tol, double timelimit, int maxsweep, int resprint, const char *csv, int bench,
                 int *stopped) {
  return guarded([&] {
    ofstream f;
    if (csv && csv[0]) f.open(csv);
    *stopped = ((CpdBase *)c)->als(tol, timelimit, maxsweep, resprint, f, bench != 0);
  });
}
int ppxh_cpd_read_W(void *c, int i, double *host) {
  return guarded([&] { ((CpdBase *)c)->W()[i].read_all(host); });
}
int ppxh_cpd_read_grad(void *c, int i, double *host) {
  return guarded([&] { ((CpdBase *)c)->grad()[i].read_all(host); });
}

// ---- Tucker (als_Tucker.h) -----------------------------------------------------------------------------------
int ppxh_hosvd(void *V, void *core, void **W, int N, const int *ranks, void *w) {
  return guarded([&] {
    MatArray Wa(W, N);
    std::vector<int> r(ranks, ranks + N);
    hosvd(*(Tensor<> *)V, *(Tensor<> *)core, Wa.ptr(), r.data(), *(World *)w);
  });
}
int ppxh_alsTucker_DT(void *V, void *core, void **W, int N, double tol, double timelimit, int maxiter, const char *csv,
                      int resprint, int bench, void *w, int *stopped) {
  return guarded([&] {
    MatArray Wa(W, N);
    ofstream f;
    if (csv && csv[0]) f.open(csv, bench ? std::ios::app : std::ios::out);
    *stopped = alsTucker_DT(*(Tensor<> *)V, *(Tensor<> *)core, Wa.ptr(), tol, timelimit, maxiter, f, resprint,
                            bench != 0, *(World *)w);
  });
}
int ppxh_alsTucker_PP(void *V, void *core, void **W, int N, double tol, double tol_init, double timelimit,
                      int maxiter, const char *csv, int resprint, int bench, void *w, int *stopped) {
  return guarded([&] {
    MatArray Wa(W, N);
    ofstream f;
    if (csv && csv[0]) f.open(csv, bench ? std::ios::app : std::ios::out);
    *stopped = alsTucker_PP(*(Tensor<> *)V, *(Tensor<> *)core, Wa.ptr(), tol, tol_init, timelimit, maxiter, f,
                            resprint, bench != 0, *(World *)w);
  });
}
int ppxh_alsTucker(void *V, void *core, void **W, int N, double tol, double timelimit, int maxiter, void *w,
                   int *stopped) {
  return guarded([&] {
    MatArray Wa(W, N);
    *stopped = alsTucker(*(Tensor<> *)V, *(Tensor<> *)core, Wa.ptr(), tol, timelimit, maxiter, *(World *)w);
  });
}

// ---- src/Tucker.h ------------------------------------------------------------------------------------------------
void *ppxh_tucker_create(int order, const int *size, const int *ranks, void *w) {
  Tucker<double> *t = nullptr;
  if (guarded([&] {
        std::vector<int> sz(size, size + order), rk(ranks, ranks + order);
        t = new Tucker<double>(order, sz.data(), rk.data(), *(World *)w);
      }))
    return nullptr;
  return t;
}
void ppxh_tucker_destroy(void *t) { delete (Tucker<double> *)t; }
// V stays owned by the caller; the factor array is created here and adopted by the object
int ppxh_tucker_init(void *t_, void *V) {
  return guarded([&] {
    Tucker<double> *t = (Tucker<double> *)t_;
    Matrix<> *arr = new Matrix<>[t->order];
    for (int i = 0; i < t->order; i++) arr[i] = Matrix<>(t->size[i], t->rank[i], *t->world);
    t->Init((Tensor<> *)V, arr);
  });
}
int ppxh_tucker_als(void *t, int pp, double tol, double tol_init, double timelimit, int maxiter, int resprint,
                    const char *csv, int *stopped) {
  return guarded([&] {
    ofstream f;
    if (csv && csv[0]) f.open(csv);
    Tucker<double> *T = (Tucker<double> *)t;
    *stopped = pp ? T->als_pp(tol, tol_init, timelimit, maxiter, resprint, f) : T->als(tol, timelimit, maxiter, resprint, f);
  });
}
int ppxh_tucker_read_W(void *t, int i, double *host) {
  return guarded([&] { ((Tucker<double> *)t)->W[i].read_all(host); });
}
int ppxh_tucker_read_core(void *t, double *host) {
  return guarded([&] { ((Tucker<double> *)t)->core.read_all(host); });
}

}  // extern "C"
