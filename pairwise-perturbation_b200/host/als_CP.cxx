// als_CP.cxx -- CP decomposition drivers on the ppx C ABI.  Control flow, printed lines and CSV columns follow
// /root/reference/als_CP.cxx (cited per block); the arithmetic is one ppx_* call where the reference has a CTF
// expression.  Differences that are deliberate are listed in DESIGN.md ("Reference quirks").
#include "als_CP.h"
#include <algorithm>
#include <cmath>
#include <numeric>

namespace {

const char *kCsvHeader = "[dim],[iter],[gradnorm],[tol],[pp_update],[diffV],[dtime]";

double synced_time(World &dw) {
  dw.sync();
  return wall_time();
}

string all_modes(int N) {
  string s;
  for (int i = 0; i < N; i++) s.push_back((char)('a' + i));
  return s;
}
string without(const string &s, int i, int j = -1) {
  string o;
  for (int k = 0; k < (int)s.size(); k++)
    if (k != i && k != j) o.push_back(s[k]);
  return o;
}

// one "[dim]= .. [iter]= .." line + CSV row (als_CP.cxx:193-201, 484-493, 724-733)
void log_row(Tensor<> &V, int iter, double gradnorm, double tol, int pp_update, double diffV, double dtime,
             ofstream &Plot_File, World &dw) {
  if (trace_sink()) trace_sink()->rows.push_back({(double)iter, gradnorm, pp_update, diffV, dtime});
  if (dw.rank != 0) return;
  // the reference prints the GLOBAL leading dimension; with a sharded leading mode report the global size
  const int64_t dim0 = (dw.np > 1 && dw.shard_mode == 0) ? dw.shard_global : V.lens[0];
  if (!trace_quiet())
    cout << "  [dim]=  " << dim0 << "  [iter]=  " << iter << "  [gradnorm]  " << gradnorm << "  [tol]  " << tol
         << "  [pp_update]  " << pp_update << "  [diffV]  " << diffV << "  [dtime]  " << dtime << "\n";
  if (Plot_File.is_open()) {
    Plot_File << dim0 << "," << iter << "," << gradnorm << "," << tol << "," << pp_update << "," << diffV << ","
              << dtime << "\n";
    if (iter % 100 == 0 && iter != 0) Plot_File << endl;
  }
}

double residual_or_skip(Tensor<> &V, Matrix<> *W, World &dw) {
  if (trace_sink() && trace_sink()->skip_residual) return -1.0;
  return cp_residual_norm(V, W, V.order, dw);
}

// local sums of squares of N matrices -> host; entry `shard_mode` is summed over ranks
void sqnorms_global(Matrix<> *const *mats, int n, double *host, World &dw, const int *mode_of = nullptr) {
  const double *xs[32];
  int64_t ns[32];
  for (int i = 0; i < n; i++) {
    xs[i] = mats[i]->data;
    ns[i] = mats[i]->size;
  }
  for (int b = 0; b < n; b += 16) {
    const int c = std::min(16, n - b);
    PPXCK(dw, ppx_sqnorms(dw.ctx, xs + b, ns + b, c, dw.scal_dev + b));
  }
  if (dw.np > 1)
    for (int i = 0; i < n; i++) {
      const int mode = mode_of ? mode_of[i] : i;
      if (mode == dw.shard_mode) dw.allreduce(dw.scal_dev + i, 1);
    }
  dw.fetch(dw.scal_dev, host, n);
}

double gradnorm_global(Matrix<> *grad_W, int N, World &dw) {
  Matrix<> *ptrs[16];
  for (int i = 0; i < N; i++) ptrs[i] = &grad_W[i];
  double h[16];
  sqnorms_global(ptrs, N, h, dw);
  double acc = 0;
  for (int i = 0; i < N; i++) acc += h[i];
  return std::sqrt(acc);
}

void normalize_with_grams(Matrix<> *W, int N, GramCache &gc, World &dw) {
  double *wp[16], *gp[16];
  int64_t s[16];
  for (int i = 0; i < N; i++) {
    wp[i] = W[i].data;
    gp[i] = gc.G[i].data;
    s[i] = W[i].nrow;
  }
  if (dw.np > 1)
    PPXCK(dw, ppx_normalize_g(dw.ctx, wp, s, N, gc.R, gp));  // norms from trace(G): G of the sharded mode is global
  else
    PPXCK(dw, ppx_normalize(dw.ctx, wp, s, N, gc.R, gp));
}

// MTTKRP of leaf mode i from the dimension tree (als_CP.cxx:236-284 / 520-569).
// Order-3 extension: a leaf whose parent is the root is contracted straight from V by the first-level rule (the
// reference recurses forever there -- als_CP.cxx:125 "V.order should be >=4").
Matrix<> leaf_mttkrp(map<string, Tensor<>> &mttkrp_map, map<string, string> &parent, map<string, string> &sibling,
                     Tensor<> &V, Matrix<> *W, int i, World &dw) {
  const string a(1, (char)('a' + i));
  const string par = parent[a];
  Matrix<> M(W[i].nrow, W[i].ncol, dw, false);
  if ((int)par.size() == V.order) {
    map<string, Tensor<>> tmp;
    mttkrp_map_DT(tmp, parent, sibling, V, W, a, dw);
    PPXCK(dw, ppx_memcpy_d2d(dw.ctx, M.data, tmp[a].data, sizeof(double) * M.size));
  } else {
    if (mttkrp_map.find(par) == mttkrp_map.end()) mttkrp_map_DT(mttkrp_map, parent, sibling, V, W, par, dw);
    Tensor<> &T = mttkrp_map[par];
    const int R = (int)W[i].ncol;
    const int pos = (int)par.find(a[0]);
    if (par.size() == 2) {  // als_CP.cxx:243-259
      const int xs = 1 - pos;
      Matrix<> &Wx = W[par[xs] - 'a'];
      PPXCK(dw, ppx_mttv(dw.ctx, T.data, T.lens, 2, xs, Wx.data, Wx.nrow, R, M.data));
    } else {  // als_CP.cxx:260-283: two factors at once
      int x1 = -1, x2 = -1;
      for (int k = 0; k < 3; k++)
        if (k != pos) (x1 < 0 ? x1 : x2) = k;
      Matrix<> &W1 = W[par[x1] - 'a'], &W2 = W[par[x2] - 'a'];
      PPXCK(dw, ppx_mttv2(dw.ctx, T.data, T.lens, 3, x1, W1.data, W1.nrow, x2, W2.data, W2.nrow, R, M.data));
    }
  }
  if (dw.np > 1 && i != dw.shard_mode) dw.allreduce(M.data, M.size);  // partial sums over the sharded mode
  return M;
}

// one exact ALS sweep over all modes with the dimension tree
// `fit_terms` (device, 3 doubles, optional): <M_N, W_N>, <S_N, G_N>, <F_N, W_N> of the LAST mode update, from which
// the residual of the sweep's result follows without a pass over V (World::fast_residual).
void dt_sweep(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *F, double lambda, bool always_regul,
              map<string, string> &parent, map<string, string> &sibling, GramCache &gc, Matrix<> &S, World &dw,
              double *fit_terms = nullptr) {
  const int N = V.order;
  map<string, Tensor<>> mttkrp_map;  // cleared every sweep (als_CP.cxx:215)
  Matrix<> Sinv((int64_t)gc.R, (int64_t)gc.R, dw, false);
  for (int i = 0; i < N; i++) {
    // S = Hadamard of the cached Grams (+ lambda I) (:288-292 / :573-579) and its inverse depend only on the OTHER
    // factors: they run on the side stream while the main stream forms the MTTKRP (35 us per mode that the sharded
    // sweep, 6.7 ms on 8 GPUs, would otherwise wait for); then gradient + solve (:296-297)
    const double lam_i = (always_regul || lambda != 0) ? lambda : 0.0;
    const double *gp[16];
    for (int j = 0; j < N; j++) gp[j] = gc.G[j].data;
    PPXCK(dw, ppx_side_begin(dw.ctx));
    PPXCK(dw, ppx_spd_inverse_g(dw.ctx, gp, N, i, lam_i, gc.R, dw.solver, S.data, Sinv.data));
    PPXCK(dw, ppx_side_end(dw.ctx));
    Matrix<> M = leaf_mttkrp(mttkrp_map, parent, sibling, V, W, i, dw);
    if (F) PPXCK(dw, ppx_axpby(dw.ctx, 1.0, F[i].data, 1.0, M.data, M.size));     // :294
    PPXCK(dw, ppx_side_join(dw.ctx));
    PPXCK(dw, ppx_solve_apply(dw.ctx, M.data, S.data, Sinv.data, W[i].data, W[i].nrow, gc.R, nullptr, 1.0,
                              grad_W[i].data, nullptr));
    gc.refresh(W, i, dw);
    if (fit_terms && i == N - 1) {
      gc.hadamard(i, 0.0, S, dw);  // without lambda I: <S, G_N> = ||[[W]]||^2
      const double *xs[3] = {M.data, S.data, F ? F[i].data : M.data};
      const double *ys[3] = {W[i].data, gc.G[i].data, W[i].data};
      int64_t ns[3] = {M.size, S.size, F ? M.size : 0};
      PPXCK(dw, ppx_dots(dw.ctx, xs, ys, ns, 3, fit_terms));
    }
  }
  normalize_with_grams(W, N, gc, dw);  // :303
}

// The exact sweep as a CUDA graph.  A sweep enqueues a fixed sequence of kernels on fixed buffers (the tree
// intermediates come out of the World's pool at the same addresses every sweep, the workspace arena is reset per call,
// nothing is read back to the host), so the FIRST sweep of a driver call runs eagerly -- it also fills the pool --, the
// SECOND is captured and launched, and every later one is a single graph launch.  At BASELINE configs[0] (order 3,
// s = 200, R = 10) a sweep is ~30 launches of a few microseconds of work each: launch-bound when enqueued one by one.
// With several GPUs the NCCL all-reduces of the sweep are captured with it (NCCL collectives are capturable; the first,
// eager sweep has already made NCCL set up its channels); every rank issues the same sequence of collectives whether it
// replays its graph or, after a failed capture, runs eagerly, so the ranks need not agree on that.  PPX_NO_MG_GRAPH
// keeps multi-GPU sweeps eager.  Any allocation that misses the pool during the capture abandons it.
struct DtSweepGraph {
  World &dw;
  void *graph = nullptr;
  bool disabled = false;
  int eager_done = 0;
  explicit DtSweepGraph(World &w) : dw(w) {
    disabled = !w.use_graph || getenv("PPX_NO_DT_GRAPH") || (w.np > 1 && getenv("PPX_NO_MG_GRAPH"));
  }
  DtSweepGraph(const DtSweepGraph &) = delete;
  ~DtSweepGraph() {
    if (graph) ppx_graph_destroy(dw.ctx, graph);
  }
  // `remaining`: sweeps still to come after this one (a capture pays off from the second replay on)
  template <typename Sweep>
  void run(Sweep &&sweep, int remaining) {
    if (graph) {
      PPXCK(dw, ppx_graph_launch(dw.ctx, graph));
      return;
    }
    if (!disabled && eager_done >= 1 && remaining >= 2) {
      bool ok = ppx_graph_begin(dw.ctx) == PPX_OK;
      if (ok) {
        dw.capturing = true;
        try {
          sweep();
        } catch (const World::CaptureMiss &) {
          ok = false;
        } catch (...) {
          dw.capturing = false;
          void *g = nullptr;
          ppx_graph_end(dw.ctx, &g);
          if (g) ppx_graph_destroy(dw.ctx, g);
          throw;
        }
        dw.capturing = false;
        void *g = nullptr;
        const int rc = ppx_graph_end(dw.ctx, &g);  // always ends the capture
        if (ok && rc == PPX_OK && g) {
          graph = g;
          PPXCK(dw, ppx_graph_launch(dw.ctx, graph));
          return;
        }
        if (g) ppx_graph_destroy(dw.ctx, g);
      }
      disabled = true;  // nothing was executed: run this sweep, and all later ones, eagerly
      if (dw.rank == 0)
        fprintf(stderr, "ppx: the exact sweep could not be captured as a CUDA graph (%s); running it kernel by kernel\n",
                ok ? ppx_last_error(dw.ctx) : "an allocation missed the pool or the capture could not start");
    }
    sweep();
    eager_done++;
  }
};

}  // namespace

vector<int> sort_indexes(const vector<double> &v) {
  vector<int> idx(v.size());
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&v](int i1, int i2) { return v[i1] > v[i2]; });
  return idx;
}

bool alsCP(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *F, double tol, double timelimit, int maxiter,
           World &dw) {
  // als_CP.cxx:20-115: one full MTTKRP per mode (KhatriRao_contract), gradient from scratch every 100 iterations
  const int N = V.order;
  double st_time = synced_time(dw);
  int iter;
  double projnorm = 0, Fnorm = 0;
  GramCache gc;
  gc.init(W, N, dw);
  Matrix<> S((int64_t)W[0].ncol, (int64_t)W[0].ncol, dw);
  for (iter = 0; iter <= maxiter; iter++) {
    if (iter % 100 == 0 || iter == maxiter) {
      gradient_CP(V, W, grad_W, dw);
      double h[32];
      Matrix<> *ptrs[32];
      vector<Matrix<>> proj;
      for (int i = 0; i < N; i++) {
        proj.emplace_back(grad_W[i]);
        PPXCK(dw, ppx_axpby(dw.ctx, -1.0, F[i].data, 1.0, proj[i].data, proj[i].size));  // grad - F  (:48)
      }
      int modes[32];
      for (int i = 0; i < N; i++) {
        ptrs[i] = &proj[i];
        ptrs[N + i] = &F[i];
        modes[i] = modes[N + i] = i;
      }
      sqnorms_global(ptrs, 2 * N, h, dw, modes);
      projnorm = 0;
      Fnorm = 0;
      for (int i = 0; i < N; i++) {
        projnorm += h[i];
        Fnorm += std::sqrt(h[N + i]);
      }
      projnorm = std::sqrt(projnorm);
      if (dw.rank == 0 && !trace_quiet())
        cout << "  [dim]=  " << V.lens[0] << "  [iter]=  " << iter << "  [projnorm]  " << projnorm << "  [tol]  "
             << tol << "  [Fnorm]  " << Fnorm << "\n";
      if (trace_sink()) trace_sink()->rows.push_back({(double)iter, projnorm, 0, Fnorm, 0.0});
      if (projnorm < tol || synced_time(dw) - st_time > timelimit) break;
    }
    for (int i = 0; i < N; i++) {
      int index[16], lens_H[16];
      for (int j = 0; j < N; j++) index[j] = j;
      std::swap(index[i], index[N - 1]);  // :66-79
      Matrix<> M(W[i].nrow, W[i].ncol, dw);
      KhatriRao_contract(M, V, W, index, lens_H, dw);
      if (dw.np > 1 && i != dw.shard_mode) dw.allreduce(M.data, M.size);
      gc.hadamard(i, 0.0, S, dw);
      PPXCK(dw, ppx_axpby(dw.ctx, 1.0, F[i].data, 1.0, M.data, M.size));
      SVD_solve(M, W[i], S);
      gc.refresh(W, i, dw);
    }
    if (Fnorm == 0) normalize_with_grams(W, N, gc, dw);
    if (iter % 10 == 0 && dw.rank == 0 && !trace_quiet()) printf(".");
  }
  if (dw.rank == 0 && !trace_quiet()) {
    printf("\nIter = %d Final proj-grad norm %E \n", iter, projnorm);
    printf("tf took %lf seconds\n", synced_time(dw) - st_time);
  }
  return iter != maxiter + 1;
}

bool alsCP_DT(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *F, double tol, double timelimit, int maxiter,
              double lambda, ofstream &Plot_File, int resprint, bool bench, World &dw) {
  // als_CP.cxx:127-320
  cout.precision(13);
  const int N = V.order;
  if (!bench && dw.rank == 0 && Plot_File.is_open()) Plot_File << kCsvHeader << "\n";
  double st_time = synced_time(dw);
  int iter;
  double projnorm = 0, diffnorm_V = 1000;
  Matrix<> S((int64_t)W[0].ncol, (int64_t)W[0].ncol, dw);
  map<string, string> parent, sibling;
  Construct_Dimension_Tree(parent, sibling, 0, N - 1);
  GramCache gc;
  gc.init(W, N, dw);
  // fast residual (World::fast_residual): ||V||^2 once, then three inner products per sweep instead of a pass over V.
  // Needs the last mode replicated (it is: only mode 0 is ever sharded); the first print point (no sweep yet) is exact.
  const bool fast = dw.fast_residual && N >= 2 && !(trace_sink() && trace_sink()->skip_residual);
  double vnorm_sq = 0;
  bool have_terms = false;
  double *fit_terms = fast ? dw.scal_dev + 40 : nullptr;
  if (fast) {
    const double *xs[1] = {V.data};
    int64_t ns[1] = {V.size};
    PPXCK(dw, ppx_sqnorms(dw.ctx, xs, ns, 1, dw.scal_dev));
    dw.allreduce(dw.scal_dev, 1);
    dw.fetch(dw.scal_dev, &vnorm_sq, 1);
  }
  DtSweepGraph sweep_graph(dw);
  for (iter = 0; iter <= maxiter; iter++) {
    if (iter % resprint == 0 || iter == maxiter) {  // :166-213
      const double st_time1 = synced_time(dw);
      projnorm = gradnorm_global(grad_W, N, dw);
      if (fast && have_terms) {
        double h[3];
        dw.fetch(fit_terms, h, 3);
        // M includes F (:294): <MTTKRP, W> = <M, W> - <F, W>
        const double r2 = vnorm_sq - 2.0 * (h[0] - (F ? h[2] : 0.0)) + h[1];
        diffnorm_V = std::sqrt(r2 > 0 ? r2 : 0.0);
      } else {
        diffnorm_V = residual_or_skip(V, W, dw);
      }
      st_time += synced_time(dw) - st_time1;  // the residual evaluation is taken off the clock (:189)
      const double dtime = wall_time() - st_time;
      if (!bench) {
        log_row(V, iter, projnorm, tol, 0, diffnorm_V, dtime, Plot_File, dw);
      } else if (iter != 0) {
        if (trace_sink()) trace_sink()->bench_times.push_back(dtime);
        if (dw.rank == 0) {
          if (!trace_quiet()) cout << "  [dimension tree step time]  " << dtime << "\n";
          if (Plot_File.is_open()) Plot_File << "[DTtime]" << "," << dtime << "\n";
        }
      }
      if (projnorm < tol || wall_time() - st_time > timelimit) break;
    }
    sweep_graph.run([&]() { dt_sweep(V, W, grad_W, F, lambda, true, parent, sibling, gc, S, dw, fit_terms); },
                    maxiter - iter);
    have_terms = fast;
    if (trace_sink()) trace_sink()->sweeps.push_back({0, iter});
    if (iter % 10 == 0 && dw.rank == 0 && !trace_quiet()) printf(".");
  }
  if (dw.rank == 0 && !trace_quiet()) {
    printf("\nIter = %d Final proj-grad norm %E \n", iter, projnorm);
    printf("tf took %lf seconds\n", synced_time(dw) - st_time);
  }
  if (!bench && Plot_File.is_open()) Plot_File.close();
  return iter != maxiter + 1;
}

void stringbuilder_mttkrp(const char *seq, char *seq_return, int N, World &dw) {
  // als_CP.cxx:323-350: contracted modes -> remaining modes followed by '*'; "0" -> all modes
  (void)dw;
  const string s(seq);
  string out;
  if (s == "0") {
    out = all_modes(N);
  } else {
    for (int i = 0; i < N; i++)
      if (s.find((char)('a' + i)) == string::npos) out.push_back((char)('a' + i));
    out.push_back('*');
  }
  memcpy(seq_return, out.c_str(), out.size() + 1);
}

void Build_mttkrp_map(map<string, Tensor<>> &mttkrp_map, Tensor<> &V, Matrix<> *W, const char *seq_c, World &dw) {
  // als_CP.cxx:352-409
  const string seq(seq_c);
  const int N = V.order;
  if (seq.size() == 1) {  // level 1: the first contraction with V (:360-380)
    mttkrp_map[seq] = contract_mode(V, all_modes(N), false, seq[0], W[seq[0] - 'a'], dw);
    return;
  }
  const string prefix = seq.substr(0, seq.size() - 1);  // :385-390
  if (mttkrp_map.find(prefix) == mttkrp_map.end()) Build_mttkrp_map(mttkrp_map, V, W, prefix.c_str(), dw);
  string kept;
  for (int i = 0; i < N; i++)
    if (prefix.find((char)('a' + i)) == string::npos) kept.push_back((char)('a' + i));
  const char x = seq.back();
  mttkrp_map[seq] = contract_mode(mttkrp_map[prefix], kept, true, x, W[x - 'a'], dw);  // :407-408
}

namespace {

// All pair operators, then all singles (als_CP.cxx:676-694); afterwards only the operators the PP sweep reads are
// kept (the reference leaves the level-1 tensors -- 32 GB at N=4, s=300, R=50 -- in the map until the next clear).
//
// Keys and results are the reference's (key = contracted modes in increasing order, als_CP.cxx:352-409); the ORDER in
// which the modes of a key are contracted is not: Build_mttkrp_map contracts them in increasing order (prefix = key
// minus its last letter), here the highest mode goes first (prefix = key minus its FIRST letter).  The operators are
// the same sums; what changes is the shape of the work:
//   * the tensor-sized first contraction always removes a trailing mode, so its row extent L = prod of the earlier
//     extents stays long.  With the leading mode sharded over GPUs the reference order contracts mode 1 of a slab with
//     37 rows (row tiles 71 % idle: the operator build reached only 4.3x on 8 GPUs in round 1);
//   * the sharded mode 0 is contracted LAST in every chain, so the only partial sums that cross GPUs are the finished
//     s x s x R operators and the s x R singles, never a level-1 tensor;
//   * a chain nobody else shares whose modes are adjacent ("ab" for N = 4, "abcd" for N = 6) is ONE fused GEMM against the
//     Khatri-Rao rows (ppx_ttm_multi): its level-1 tensor is never written.
struct PPBuildPlan {
  map<string, int> uses;  // how many target chains pass through a node (keyed like the map)
  static string suffix(const string &key, size_t from) { return key.substr(from); }
  explicit PPBuildPlan(const vector<string> &targets) {
    for (const string &k : targets)
      for (size_t f = 0; f < k.size(); f++) uses[suffix(k, f)]++;
  }
  bool private_adjacent_chain(const string &key) const {
    for (size_t j = 1; j < key.size(); j++)
      if (key[j] != key[j - 1] + 1) return false;
    auto it = uses.find(string(1, key.back()));
    return key.size() >= 2 && it != uses.end() && it->second == 1;
  }
};

void build_key_last_first(map<string, Tensor<>> &m, Tensor<> &V, Matrix<> *W, const string &key, const PPBuildPlan &plan,
                          World &dw) {
  if (m.find(key) != m.end()) return;
  const int N = V.order;
  const int R = (int)W[0].ncol;
  if (plan.private_adjacent_chain(key) && key.size() <= 8) {
    const int x_first = key[0] - 'a', n = (int)key.size();
    int64_t out_lens[17];
    int k = 0;
    for (int i = 0; i < N; i++)
      if (i < x_first || i >= x_first + n) out_lens[k++] = V.lens[i];
    out_lens[k++] = R;
    Tensor<> out(k, out_lens, dw, false);
    const double *wp[8];
    int64_t ld[8];
    for (int j = 0; j < n; j++) {
      wp[j] = W[x_first + j].data;
      ld[j] = W[x_first + j].nrow;
    }
    PPXCK(dw, ppx_ttm_multi(dw.ctx, V.data, V.lens, N, x_first, n, wp, ld, R, out.data));
    m[key] = std::move(out);
    return;
  }
  if (key.size() == 1) {  // level 1: the first contraction with V (als_CP.cxx:360-380)
    m[key] = contract_mode(V, all_modes(N), false, key[0], W[key[0] - 'a'], dw);
    return;
  }
  const string prefix = key.substr(1);
  build_key_last_first(m, V, W, prefix, plan, dw);
  string kept;
  for (int i = 0; i < N; i++)
    if (prefix.find((char)('a' + i)) == string::npos) kept.push_back((char)('a' + i));
  const char x = key[0];
  m[key] = contract_mode(m[prefix], kept, true, x, W[x - 'a'], dw);  // als_CP.cxx:407-408
}

void build_pp_operators(map<string, Tensor<>> &mttkrp_map, Tensor<> &V, Matrix<> *W, World &dw) {
  const int N = V.order;
  const string seq = all_modes(N);
  mttkrp_map.clear();
  const char sh = (char)('a' + dw.shard_mode);
  vector<string> pairs, singles;
  for (int ii = 0; ii < N; ii++)
    for (int jj = ii + 1; jj < N; jj++) pairs.push_back(without(seq, ii, jj));
  for (int ii = 0; ii < N; ii++) singles.push_back(without(seq, ii));
  if (getenv("PPX_PP_BUILD_REFERENCE_ORDER")) {  // the reference's contraction order (A/B timing, parity tests)
    for (const string &key : pairs) {
      Build_mttkrp_map(mttkrp_map, V, W, key.c_str(), dw);
      if (dw.np > 1 && key.find(sh) != string::npos) dw.allreduce(mttkrp_map[key].data, mttkrp_map[key].size);
    }
    for (const string &key : singles) Build_mttkrp_map(mttkrp_map, V, W, key.c_str(), dw);  // from reduced pair operators
  } else {
    const PPBuildPlan plan(pairs);
    // a result is a partial sum over the local slab iff this step contracted the sharded mode
    auto finish = [&](const string &key, bool from_V) {
      if (dw.np == 1 || key.find(sh) == string::npos) return;
      if (from_V || key.substr(1).find(sh) == string::npos) dw.allreduce(mttkrp_map[key].data, mttkrp_map[key].size);
    };
    // N = 4: a level-1 tensor has three kept modes, and every pair operator that descends from it is ONE Hadamard
    // contraction of it -- all of them in one pass over the 10.8 GB intermediate (ppx_mttv3) instead of one pass each
    if (N == 4) {
      for (int pm = N - 1; pm >= 0; pm--) {
        const string p(1, (char)('a' + pm));
        auto it = plan.uses.find(p);
        if (it == plan.uses.end() || it->second < 2) continue;
        string kept;
        for (int i = 0; i < N; i++)
          if (i != pm) kept.push_back((char)('a' + i));
        double *outs[3] = {nullptr, nullptr, nullptr};
        Tensor<> res[3];
        string keys[3];
        int wanted = 0;
        for (int q = 0; q < 3; q++) {
          if (kept[q] > p[0]) continue;  // that operator descends from the higher letter's level-1 tensor
          keys[q] = string(1, kept[q]) + p;
          if (std::find(pairs.begin(), pairs.end(), keys[q]) == pairs.end() || mttkrp_map.count(keys[q])) continue;
          int64_t out_lens[3];
          int n = 0;
          for (int z = 0; z < 3; z++)
            if (z != q) out_lens[n++] = V.lens[kept[z] - 'a'];
          out_lens[n++] = W[0].ncol;
          res[q] = Tensor<>(n, out_lens, dw, false);
          outs[q] = res[q].data;
          wanted++;
        }
        if (wanted < 2) continue;
        build_key_last_first(mttkrp_map, V, W, p, plan, dw);
        Tensor<> &T1 = mttkrp_map[p];
        Matrix<> &Wl = W[kept[0] - 'a'], &Wx = W[kept[1] - 'a'], &Wt = W[kept[2] - 'a'];
        PPXCK(dw, ppx_mttv3(dw.ctx, T1.data, T1.lens, (int)W[0].ncol, Wl.data, Wl.nrow, Wx.data, Wx.nrow, Wt.data, Wt.nrow,
                            outs[0], outs[1], outs[2]));
        for (int q = 0; q < 3; q++)
          if (outs[q]) {
            mttkrp_map[keys[q]] = std::move(res[q]);
            finish(keys[q], false);
          }
      }
    }
    for (const string &key : pairs) {
      const bool had = mttkrp_map.find(key) != mttkrp_map.end();
      build_key_last_first(mttkrp_map, V, W, key, plan, dw);
      if (!had) finish(key, plan.private_adjacent_chain(key));
    }
    const PPBuildPlan none(vector<string>{});
    for (const string &key : singles) {
      build_key_last_first(mttkrp_map, V, W, key, none, dw);
      finish(key, false);
    }
  }
  for (auto it = mttkrp_map.begin(); it != mttkrp_map.end();) {
    if ((int)it->first.size() < N - 2) it = mttkrp_map.erase(it);
    else ++it;
  }
}

// PP-corrected MTTKRP of mode i (als_CP.cxx:774-794), all operators in one launch.
// part = PP_ALL: the whole sum.  The sweep splits it so that only the term that depends on the mode updated just before
// sits on the critical path: PP_EARLY = M0 + every term except j = i-1 (enqueued on a lane while mode i-1 is being
// solved: dW_j for j < i-1 is final by then, dW_j for j > i is last sweep's), PP_LATE = M += the j = i-1 term.
enum PPPart { PP_ALL, PP_EARLY, PP_LATE };
void pp_corrected_mttkrp(map<string, Tensor<>> &mttkrp_map, Matrix<> *W, Matrix<> *dW, int i, int N, Matrix<> &M,
                         Matrix<> &zero_M, World &dw, PPPart part = PP_ALL) {
  const string seq = all_modes(N);
  const double *ops[16], *dws[16];
  int which[16];
  int64_t s_other[16];
  int n = 0;
  const bool reduce = dw.np > 1 && i != dw.shard_mode;
  for (int j = 0; j < N; j++) {
    if (j == i) continue;
    if (part == PP_EARLY && j == i - 1) continue;
    if (part == PP_LATE && j != i - 1) continue;
    // multi-GPU: P^(shard,i) is contracted over the sharded index -> partial sums; the replicated terms are added
    // on rank 0 only and the result is summed over ranks
    if (reduce && dw.rank != 0 && j != dw.shard_mode) continue;
    Tensor<> &P = mttkrp_map[without(seq, std::min(i, j), std::max(i, j))];
    ops[n] = P.data;
    which[n] = (j < i) ? 0 : 1;  // j<i: contract the operator's first index (:785); j>i: the second (:793)
    dws[n] = dW[j].data;
    s_other[n] = dW[j].nrow;
    n++;
  }
  const double *M0 = part == PP_LATE ? M.data  // in place: every element is read and written by the same thread
                                     : (reduce && dw.rank != 0) ? zero_M.data : mttkrp_map[without(seq, i)].data;  // :778
  if (part != PP_LATE || n > 0)
    PPXCK(dw, ppx_pp_correct(dw.ctx, M0, ops, which, dws, s_other, n, W[i].nrow, (int)W[i].ncol, M.data));
  if (reduce && part != PP_EARLY) dw.allreduce(M.data, M.size);
}

}  // namespace

double alsCP_DT_sub(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *dW, Matrix<> *F, double tol, double tol_init,
                    double timelimit, int maxiter, double &st_time, double lambda, ofstream &Plot_File,
                    double &projnorm, int &iter, int resprint, World &dw) {
  // als_CP.cxx:418-612
  (void)F;
  const int N = V.order;
  vector<Matrix<>> W_prev;  // zero-initialised: the first sweep can never switch (:428-431)
  for (int i = 0; i < N; i++) W_prev.emplace_back(W[i].nrow, W[i].ncol, dw);
  double diffnorm_V = 1000;
  Matrix<> S((int64_t)W[0].ncol, (int64_t)W[0].ncol, dw);
  map<string, string> parent, sibling;
  Construct_Dimension_Tree(parent, sibling, 0, N - 1);
  GramCache gc;
  gc.init(W, N, dw);
  DtSweepGraph sweep_graph(dw);
  for (; iter <= maxiter; iter++) {
    if (iter % resprint == 0 || iter == maxiter) {  // :457-498
      const double st_time1 = synced_time(dw);
      projnorm = gradnorm_global(grad_W, N, dw);
      diffnorm_V = residual_or_skip(V, W, dw);
      st_time += synced_time(dw) - st_time1;
      const double dtime = wall_time() - st_time;
      log_row(V, iter, projnorm, tol, 0, diffnorm_V, dtime, Plot_File, dw);
      if (projnorm < tol || wall_time() - st_time > timelimit) break;
    }
    sweep_graph.run([&]() { dt_sweep(V, W, grad_W, nullptr, lambda, false, parent, sibling, gc, S, dw); },
                    maxiter - iter);  // :499-592
    if (trace_sink()) trace_sink()->sweeps.push_back({0, iter});
    // :594-605  dW = W - W_prev; W_prev = W; switch when every ||dW_i||/||W_i|| < tol_init
    for (int i = 0; i < N; i++)
      PPXCK(dw, ppx_diff_update(dw.ctx, W[i].data, W_prev[i].data, dW[i].data, W[i].size, dw.scal_dev + 2 * i));
    if (dw.np > 1) dw.allreduce(dw.scal_dev + 2 * dw.shard_mode, 2);
    double h[32];
    dw.fetch(dw.scal_dev, h, 2 * N);
    int num_dw_break = 0;
    for (int i = 0; i < N; i++)
      if (std::fabs(std::sqrt(h[2 * i]) / std::sqrt(h[2 * i + 1])) < tol_init) num_dw_break++;
    if (num_dw_break == N) return diffnorm_V;  // returns BEFORE iter++ of this sweep (:604-605)
    if (iter % 10 == 0 && dw.rank == 0 && !trace_quiet()) printf(".");
  }
  return diffnorm_V;
}

namespace {

// One PP phase: the operators built at W_init and the approximate sweep that reads them (als_CP.cxx:672-694 and
// :754-825).  The sweep touches a fixed set of buffers, so it is captured once per phase as a CUDA graph.
struct PPPhase {
  Tensor<> &V;
  Matrix<> *W, *grad_W, *dW;
  double lambda, ratio_step;
  World &dw;
  int N, R;
  vector<Matrix<>> W_init, M;
  map<string, Tensor<>> ops;
  Matrix<> S, Sinv, zero_M;
  GramCache gc;
  void *graph = nullptr;
  // PPX_PP_TRACE=1: %globaltimer stamps between the kernels of the sweep, printed after the first two replays
  unsigned long long *trace_dev = nullptr;
  vector<string> trace_labels;
  int trace_prints = 0;
  void stamp(const char *label) {
    if (!trace_dev || trace_labels.size() >= 256) return;
    trace_labels.push_back(label);
    PPXCK(dw, ppx_stamp(dw.ctx, trace_dev + trace_labels.size() - 1));
  }
  void trace_print() {
    if (!trace_dev || trace_prints >= 2 || dw.rank != 0) return;
    trace_prints++;
    dw.sync();
    vector<unsigned long long> t(trace_labels.size());
    PPXCK(dw, ppx_memcpy_d2h(dw.ctx, t.data(), trace_dev, sizeof(unsigned long long) * t.size()));
    dw.sync();
    unsigned long long t0 = ~0ull;
    for (auto x : t) t0 = std::min(t0, x);
    fprintf(stderr, "ppx: PP sweep timeline (us since the first stamp)\n");
    for (size_t k = 0; k < t.size(); k++) fprintf(stderr, "  %8.2f  %s\n", (t[k] - t0) * 1e-3, trace_labels[k].c_str());
  }

  PPPhase(Tensor<> &V_, Matrix<> *W_, Matrix<> *grad_W_, Matrix<> *dW_, double lambda_, double ratio_step_, World &dw_)
      : V(V_), W(W_), grad_W(grad_W_), dW(dW_), lambda(lambda_), ratio_step(ratio_step_), dw(dw_), N(V_.order),
        R((int)W_[0].ncol), W_init(V_.order), S((int64_t)W_[0].ncol, (int64_t)W_[0].ncol, dw_),
        Sinv((int64_t)W_[0].ncol, (int64_t)W_[0].ncol, dw_) {
    for (int i = 0; i < N; i++) M.emplace_back(W[i].nrow, W[i].ncol, dw);
    if (dw.np > 1) {
      int64_t smax = 0;
      for (int i = 0; i < N; i++) smax = std::max(smax, W[i].nrow);
      zero_M = Matrix<>(smax, R, dw);
    }
    gc.init(W, N, dw);
    if (getenv("PPX_PP_TRACE")) PPXCK(dw, ppx_malloc(dw.ctx, 256 * sizeof(unsigned long long), (void **)&trace_dev));
  }
  ~PPPhase() {
    if (graph) ppx_graph_destroy(dw.ctx, graph);
    if (trace_dev) ppx_free(dw.ctx, trace_dev);
  }
  // W_init = W, dW = 0, build all pair operators and singles (:672-694)
  void build() {
    for (int j = 0; j < N; j++) {
      W_init[j] = W[j];
      dW[j].set_zero();
    }
    build_pp_operators(ops, V, W, dw);
    // The sweep touches a fixed set of buffers from here on: capture it now (capturing enqueues nothing), so that
    // every approximate sweep of the phase -- including the first, the one pp_bench times -- is one graph launch.
    if (dw.use_graph && !(dw.np > 1 && getenv("PPX_NO_MG_GRAPH")) && !graph) {
      PPXCK(dw, ppx_graph_begin(dw.ctx));
      try {
        enqueue_sweep();
      } catch (...) {
        // leave the context usable: back on the main stream, capture ended, nothing instantiated
        ppx_side_end(dw.ctx);
        void *g = nullptr;
        ppx_graph_end(dw.ctx, &g);
        if (g) ppx_graph_destroy(dw.ctx, g);
        throw;
      }
      PPXCK(dw, ppx_graph_end(dw.ctx, &graph));
    }
  }
  // per mode: correction -> Gram-Hadamard -> solve (+grad, dW) -> Gram; then Normalize and the 2N squared norms
  // the switching test needs, copied to pinned host memory (:754-825, :657-664)
  void enqueue_sweep() {
    // PPX_PP_SPLIT=1: the correction of mode i+1 minus its dW_i term on lane 1 under the solve of mode i, only the
    // dW_i term on the critical path.  Measured at cfg2 (stamps, PPX_PP_TRACE): the critical-path term drops from 24
    // to 11 us, but the 38-CTA apply kernel running beside the 500-CTA lane kernel takes 25-33 us instead of 10, and
    // the sweep is 0.197 ms against 0.173 in one launch per mode -- off by default (a timing experiment: the parity
    // suite runs the default path only).
    static const bool split = getenv("PPX_PP_SPLIT") != nullptr;
    trace_labels.clear();
    stamp("sweep begin");
    for (int i = 0; i < N; i++) {
      // S from the CURRENT W (:796-802) and its inverse depend only on the Grams: they run on the side stream while
      // the main stream forms the corrected MTTKRP (:774-794); then gradient + SVD_solve_mod (:811-812)
      const double *gp[16];
      for (int j = 0; j < N; j++) gp[j] = gc.G[j].data;
      PPXCK(dw, ppx_lane_begin(dw.ctx, 0));
      stamp("  lane0 inverse >");
      PPXCK(dw, ppx_spd_inverse_g(dw.ctx, gp, N, i, lambda, R, dw.solver, S.data, Sinv.data));
      stamp("  lane0 inverse <");
      PPXCK(dw, ppx_lane_end(dw.ctx));
      if (split && i > 0) PPXCK(dw, ppx_lane_join(dw.ctx, 1));
      stamp("main correction (late or all) >");
      pp_corrected_mttkrp(ops, W, dW, i, N, M[i], zero_M, dw, (split && i > 0) ? PP_LATE : PP_ALL);
      stamp("main correction <");
      // the correction of mode i+1 minus its dW_i term reads nothing this mode writes: lane 1, forked AFTER this
      // mode's own term so that the two do not share the HBM (started together, the short one finishes with the long
      // one); it then runs under the solve and the Gram of this mode, which move almost nothing
      if (split && i + 1 < N) {
        PPXCK(dw, ppx_lane_begin(dw.ctx, 1));
        stamp("    lane1 early >");
        pp_corrected_mttkrp(ops, W, dW, i + 1, N, M[i + 1], zero_M, dw, PP_EARLY);
        stamp("    lane1 early <");
        PPXCK(dw, ppx_lane_end(dw.ctx));
      }
      PPXCK(dw, ppx_lane_join(dw.ctx, 0));
      stamp("main apply >");
      PPXCK(dw, ppx_solve_apply(dw.ctx, M[i].data, S.data, Sinv.data, W[i].data, W[i].nrow, R, W_init[i].data,
                                ratio_step, grad_W[i].data, dW[i].data));
      stamp("main gram >");
      gc.refresh(W, i, dw);
      stamp("main gram <");
    }
    // Normalize after dW was taken -- W_init is never rescaled (:825) -- and the 2N squared norms of the switching test
    if (dw.np == 1) {
      double *wp[16], *gp[16];
      const double *dp[16];
      int64_t s[16];
      for (int i = 0; i < N; i++) {
        wp[i] = W[i].data;
        gp[i] = gc.G[i].data;
        dp[i] = dW[i].data;
        s[i] = W[i].nrow;
      }
      PPXCK(dw, ppx_normalize_norms(dw.ctx, wp, dp, s, N, R, gp, dw.scal_dev));
    } else {
      normalize_with_grams(W, N, gc, dw);
      const double *xs[32];
      int64_t ns[32];
      for (int i = 0; i < N; i++) {
        xs[2 * i] = dW[i].data;
        xs[2 * i + 1] = W[i].data;
        ns[2 * i] = ns[2 * i + 1] = W[i].size;
      }
      for (int b = 0; b < 2 * N; b += 16)
        PPXCK(dw, ppx_sqnorms(dw.ctx, xs + b, ns + b, std::min(16, 2 * N - b), dw.scal_dev + b));
      dw.allreduce(dw.scal_dev + 2 * dw.shard_mode, 2);
    }
    stamp("sweep end (normalised, norms taken)");
    PPXCK(dw, ppx_memcpy_d2h(dw.ctx, dw.scal_host, dw.scal_dev, sizeof(double) * 2 * N));
  }
  void sweep() {
    if (graph) PPXCK(dw, ppx_graph_launch(dw.ctx, graph));
    else enqueue_sweep();
    trace_print();
  }
};

}  // namespace

double alsCP_PP_sub(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *dW, Matrix<> *F, double tol, double tol_init,
                    double timelimit, int maxiter, double &st_time, double lambda, double ratio_step,
                    ofstream &Plot_File, double &projnorm, int &iter, int resprint, bool bench, World &dw) {
  // als_CP.cxx:621-833
  (void)F;
  const int N = V.order;
  double dtime_first = 0;
  const int init_iter = iter;
  double diffnorm_V = 1000;
  PPPhase phase(V, W, grad_W, dW, lambda, ratio_step, dw);
  bool norms_valid = false;  // scal_host holds ||dW_i||^2, ||W_i||^2 of the state after the last sweep
  for (; iter <= maxiter; iter++) {
    int num_dw_break = 0;
    if (!bench) {  // :657-665
      double h[32];
      if (norms_valid) {
        dw.sync();
        memcpy(h, dw.scal_host, sizeof(double) * 2 * N);
      } else {
        Matrix<> *ptrs[32];
        int modes[32];
        for (int i = 0; i < N; i++) {
          ptrs[2 * i] = &dW[i];
          ptrs[2 * i + 1] = &W[i];
          modes[2 * i] = modes[2 * i + 1] = i;
        }
        sqnorms_global(ptrs, 2 * N, h, dw, modes);
      }
      for (int i = 0; i < N; i++)
        if (std::fabs(std::sqrt(h[2 * i]) / std::sqrt(h[2 * i + 1])) > tol_init) num_dw_break++;
    }
    if ((iter - init_iter) % 15 == 0 || num_dw_break > 0) {  // :667-695
      if (num_dw_break > 0 || iter != init_iter) return diffnorm_V;
      phase.build();
      if (trace_sink()) trace_sink()->sweeps.push_back({2, iter});
    }
    if (iter % resprint == 0 || iter == maxiter || iter == init_iter) {  // :697-752
      const double st_time1 = synced_time(dw);
      projnorm = gradnorm_global(grad_W, N, dw);
      diffnorm_V = residual_or_skip(V, W, dw);
      norms_valid = false;  // scal_dev was reused
      st_time += synced_time(dw) - st_time1;
      const double dtime = wall_time() - st_time;
      if (!bench) {
        log_row(V, iter, projnorm, tol, 1, diffnorm_V, dtime, Plot_File, dw);
      } else if (iter != maxiter) {  // :736-738
        dtime_first = dtime;
        st_time = wall_time();
      } else {  // :739-747
        dtime_first = dtime_first + dtime;
        if (trace_sink()) {
          trace_sink()->bench_times.push_back(dtime_first);
          trace_sink()->bench_times.push_back(dtime);
        }
        if (dw.rank == 0) {
          if (!trace_quiet()) {
            cout << "  [PP first time]  " << dtime_first << "\n";
            cout << "  [PP second time]  " << dtime << "\n";
          }
          if (Plot_File.is_open()) {
            Plot_File << "  [PPfirst]  " << "," << dtime_first << "\n";
            Plot_File << "  [PPsecond]  " << "," << dtime << "\n";
          }
        }
      }
      if (projnorm < tol || wall_time() - st_time > timelimit) break;
    }
    phase.sweep();
    norms_valid = true;
    if (trace_sink()) trace_sink()->sweeps.push_back({1, iter});
    if (iter % 10 == 0 && dw.rank == 0 && !trace_quiet()) printf(".");
  }
  if (bench) iter++;
  return diffnorm_V;
}

// ---- timing helpers for bench.py (no logging, nothing but the sweeps on the stream) -----------------------------
void alsCP_DT_sweeps(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, int n_sweeps, double lambda, World &dw) {
  const int N = V.order;
  Matrix<> S((int64_t)W[0].ncol, (int64_t)W[0].ncol, dw);
  map<string, string> parent, sibling;
  Construct_Dimension_Tree(parent, sibling, 0, N - 1);
  GramCache gc;
  gc.init(W, N, dw);
  for (int it = 0; it < n_sweeps; it++) dt_sweep(V, W, grad_W, nullptr, lambda, true, parent, sibling, gc, S, dw);
}

void alsCP_PP_phase_timed(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, int n_sweeps, double lambda, double ratio_step,
                          World &dw, float *ms_build, float *ms_sweeps) {
  const int N = V.order;
  vector<Matrix<>> dW;
  for (int j = 0; j < N; j++) dW.emplace_back(W[j].nrow, W[j].ncol, dw);
  PPPhase phase(V, W, grad_W, dW.data(), lambda, ratio_step, dw);
  void *e0, *e1, *e2;
  PPXCK(dw, ppx_event_create(dw.ctx, &e0));
  PPXCK(dw, ppx_event_create(dw.ctx, &e1));
  PPXCK(dw, ppx_event_create(dw.ctx, &e2));
  PPXCK(dw, ppx_event_record(dw.ctx, e0));
  phase.build();
  PPXCK(dw, ppx_event_record(dw.ctx, e1));
  phase.sweep();  // first sweep captures the graph; not timed
  dw.sync();
  PPXCK(dw, ppx_event_record(dw.ctx, e2));
  PPXCK(dw, ppx_event_elapsed_ms(dw.ctx, e0, e1, ms_build));
  void *t0, *t1;
  PPXCK(dw, ppx_event_create(dw.ctx, &t0));
  PPXCK(dw, ppx_event_create(dw.ctx, &t1));
  PPXCK(dw, ppx_event_record(dw.ctx, t0));
  for (int it = 0; it < n_sweeps; it++) phase.sweep();
  PPXCK(dw, ppx_event_record(dw.ctx, t1));
  PPXCK(dw, ppx_event_elapsed_ms(dw.ctx, t0, t1, ms_sweeps));
  for (void *e : {e0, e1, e2, t0, t1}) ppx_event_destroy(dw.ctx, e);
}

void alsCP_DT_mttkrps(Tensor<> &V, Matrix<> *W, Matrix<> *M_out, World &dw) {
  const int N = V.order;
  map<string, string> parent, sibling;
  Construct_Dimension_Tree(parent, sibling, 0, N - 1);
  map<string, Tensor<>> mttkrp_map;
  for (int i = 0; i < N; i++) M_out[i] = leaf_mttkrp(mttkrp_map, parent, sibling, V, W, i, dw);
}

void alsCP_PP_operators(Tensor<> &V, Matrix<> *W, map<string, Tensor<>> &ops, World &dw) {
  build_pp_operators(ops, V, W, dw);
}

bool alsCP_PP(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *F, double tol, double tol_init, double timelimit,
              int maxiter, double lambda, double ratio_step, ofstream &Plot_File, int resprint, bool bench,
              World &dw) {
  // als_CP.cxx:1082-1137
  cout.precision(13);
  const int N = V.order;
  if (!bench && dw.rank == 0 && Plot_File.is_open()) Plot_File << kCsvHeader << "\n";
  double st_time = synced_time(dw);
  int iter = 0;
  double gradnorm = 10.;
  vector<Matrix<>> dW;
  for (int j = 0; j < N; j++) dW.emplace_back(W[j].nrow, W[j].ncol, dw);
  while (gradnorm > tol && iter <= maxiter) {
    if (!bench) {
      if (dw.rank == 0 && !trace_quiet()) printf("DT starts from %d\n", iter);
      if (trace_sink()) trace_sink()->events.push_back({0, iter});
      alsCP_DT_sub(V, W, grad_W, dW.data(), F, tol, tol_init, timelimit, maxiter, st_time, lambda, Plot_File, gradnorm,
                   iter, resprint, dw);
    }
    if (dw.rank == 0 && !trace_quiet()) printf("pairwise perturbation starts from %d\n", iter);
    if (trace_sink()) trace_sink()->events.push_back({1, iter});
    alsCP_PP_sub(V, W, grad_W, dW.data(), F, tol, tol_init, timelimit, maxiter, st_time, lambda, ratio_step, Plot_File,
                 gradnorm, iter, resprint, bench, dw);
  }
  if (dw.rank == 0 && !trace_quiet()) {
    printf("\nIter = %d Final grad norm %E \n", iter, gradnorm);
    printf("tf took %lf seconds\n", synced_time(dw) - st_time);
  }
  if (!bench && Plot_File.is_open()) Plot_File.close();
  return iter != maxiter + 1;
}

double alsCP_PP_partupdate_sub(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *dW, Matrix<> *F, double tol,
                               double tol_init, double timelimit, int maxiter, double update_percentage,
                               double &st_time, double lambda, double ratio_step, ofstream &Plot_File,
                               double &projnorm, int &iter, int resprint, bool bench, World &dw) {
  // als_CP.cxx:852-1073: like alsCP_PP_sub but the perturbation dM is PROPAGATED to the other modes after each
  // update (:1036-1053) and only the update_size most perturbed modes are solved per sweep (:991-1001)
  (void)F;
  const int N = V.order;
  const int R = (int)W[0].ncol;
  // multi-GPU: W_0, dW_0, M_0, dM_0 hold the local rows of the sharded mode; everything else is replicated.  The only
  // exchange inside the sweep is the propagation of dW_0 to the other modes (a contraction over the sharded index).
  Matrix<> dM_term;
  if (dw.np > 1) {
    int64_t smax = 0;
    for (int i = 0; i < N; i++) smax = std::max(smax, W[i].nrow);
    dM_term = Matrix<>(smax, R, dw);
  }
  double dtime_first = 0;
  const int init_iter = iter;
  double diffnorm_V = 1000;
  vector<Matrix<>> W_init(N), dM, M;
  for (int i = 0; i < N; i++) {
    dM.emplace_back(W[i].nrow, W[i].ncol, dw);
    M.emplace_back(W[i].nrow, W[i].ncol, dw);
  }
  map<string, Tensor<>> mttkrp_map;
  Matrix<> S((int64_t)R, (int64_t)R, dw);
  GramCache gc;
  gc.init(W, N, dw);
  vector<double> W_relative_perturbe(N, 0.);
  const int update_size = (int)(N * update_percentage);
  const string seq = all_modes(N);
  for (; iter <= maxiter; iter++) {
    int num_dw_break = 0;
    if (!bench) {
      Matrix<> *ptrs[32];
      int modes[32];
      double h[32];
      for (int i = 0; i < N; i++) {
        ptrs[2 * i] = &dW[i];
        ptrs[2 * i + 1] = &W[i];
        modes[2 * i] = modes[2 * i + 1] = i;
      }
      sqnorms_global(ptrs, 2 * N, h, dw, modes);
      for (int i = 0; i < N; i++)
        if (std::fabs(std::sqrt(h[2 * i]) / std::sqrt(h[2 * i + 1])) > tol_init) num_dw_break++;
    }
    if ((iter - init_iter) % 15 == 0 || num_dw_break > 0) {
      if (num_dw_break > 0 || iter != init_iter) return diffnorm_V;
      for (int j = 0; j < N; j++) {
        W_init[j] = W[j];
        dW[j].set_zero();
      }
      build_pp_operators(mttkrp_map, V, W, dw);
      if (trace_sink()) trace_sink()->sweeps.push_back({2, iter});
    }
    if (iter % resprint == 0 || iter == maxiter || iter == init_iter) {
      const double st_time1 = synced_time(dw);
      projnorm = gradnorm_global(grad_W, N, dw);
      diffnorm_V = residual_or_skip(V, W, dw);
      st_time += synced_time(dw) - st_time1;
      const double dtime = wall_time() - st_time;
      if (!bench) {
        log_row(V, iter, projnorm, tol, 1, diffnorm_V, dtime, Plot_File, dw);
      } else if (iter != maxiter) {
        dtime_first = dtime;
        st_time = wall_time();
      } else {
        dtime_first = dtime_first + dtime;
        if (trace_sink()) {
          trace_sink()->bench_times.push_back(dtime_first);
          trace_sink()->bench_times.push_back(dtime);
        }
        if (dw.rank == 0 && !trace_quiet()) {
          cout << "  [PP first time]  " << dtime_first << "\n";
          cout << "  [PP second time]  " << dtime << "\n";
        }
      }
      if (projnorm < tol || wall_time() - st_time > timelimit) break;
    }
    const vector<int> sorted_indices = sort_indexes(W_relative_perturbe);  // :992
    if (dw.rank == 0 && !trace_quiet()) cout << "new round" << endl;
    for (int t = 0; t < update_size; t++) {
      const int i = sorted_indices[t];
      if (dw.rank == 0 && !trace_quiet()) cout << i << endl;
      // M[i] = M_i(W_init) + dM[i]   (:1025)
      PPXCK(dw, ppx_memcpy_d2d(dw.ctx, M[i].data, mttkrp_map[without(seq, i)].data, sizeof(double) * M[i].size));
      PPXCK(dw, ppx_axpby(dw.ctx, 1.0, dM[i].data, 1.0, M[i].data, M[i].size));
      gc.hadamard(i, lambda, S, dw);
      solve_update_fused(M[i], S, W[i], &W_init[i], ratio_step, &grad_W[i], &dW[i], dw.solver, dw);  // :1034-1035
      gc.refresh(W, i, dw);
      dM[i].set_zero();  // :1037
      for (int ii = 0; ii < N; ii++) {  // :1038-1053  dM[ii] += P^(ii,i) x dW[i]
        if (ii == i) continue;
        Tensor<> &P = mttkrp_map[without(seq, std::min(i, ii), std::max(i, ii))];
        const double *ops[1] = {P.data}, *dws[1] = {dW[i].data};
        const int which[1] = {i < ii ? 0 : 1};  // contract the index that belongs to mode i
        const int64_t so[1] = {dW[i].nrow};
        if (dw.np > 1 && i == dw.shard_mode) {
          // partial sum over the local rows of mode i: 0 + P x dW_i into a scratch matrix, summed over ranks, then added
          const int64_t n = W[ii].nrow * (int64_t)R;
          PPXCK(dw, ppx_memset_zero(dw.ctx, dM_term.data, sizeof(double) * n));
          PPXCK(dw, ppx_pp_correct(dw.ctx, dM_term.data, ops, which, dws, so, 1, W[ii].nrow, R, dM_term.data));
          dw.allreduce(dM_term.data, n);
          PPXCK(dw, ppx_axpby(dw.ctx, 1.0, dM_term.data, 1.0, dM[ii].data, n));
        } else {
          PPXCK(dw, ppx_pp_correct(dw.ctx, dM[ii].data, ops, which, dws, so, 1, W[ii].nrow, R, dM[ii].data));
        }
      }
    }
    {  // :1060-1064
      Matrix<> *ptrs[32];
      int modes[32];
      double h[32];
      for (int i = 0; i < N; i++) {
        ptrs[2 * i] = &dM[i];
        ptrs[2 * i + 1] = &M[i];
        modes[2 * i] = modes[2 * i + 1] = i;
      }
      sqnorms_global(ptrs, 2 * N, h, dw, modes);
      for (int i = 0; i < N; i++) W_relative_perturbe[i] = std::sqrt(h[2 * i]) / std::sqrt(h[2 * i + 1]);
    }
    normalize_with_grams(W, N, gc, dw);
    if (trace_sink()) trace_sink()->sweeps.push_back({1, iter});
    if (iter % 10 == 0 && dw.rank == 0 && !trace_quiet()) printf(".");
  }
  if (bench) iter++;
  return diffnorm_V;
}

bool alsCP_PP_partupdate(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *F, double tol, double tol_init,
                         double timelimit, int maxiter, double lambda, double ratio_step, double update_percentage,
                         ofstream &Plot_File, int resprint, bool bench, World &dw) {
  // als_CP.cxx:1146-1207
  cout.precision(13);
  const int N = V.order;
  if (dw.rank == 0 && !trace_quiet()) cout << "alsCP_PP_partupdate starts. " << endl;
  if (!bench && dw.rank == 0 && Plot_File.is_open()) Plot_File << kCsvHeader << "\n";
  double st_time = synced_time(dw);
  int iter = 0;
  double gradnorm = 10.;
  vector<Matrix<>> dW;
  for (int j = 0; j < N; j++) dW.emplace_back(W[j].nrow, W[j].ncol, dw);
  while (gradnorm > tol && iter <= maxiter) {
    if (!bench) {
      if (dw.rank == 0 && !trace_quiet()) printf("DT starts from %d\n", iter);
      if (trace_sink()) trace_sink()->events.push_back({0, iter});
      alsCP_DT_sub(V, W, grad_W, dW.data(), F, tol, tol_init, timelimit, maxiter, st_time, lambda, Plot_File, gradnorm,
                   iter, resprint, dw);
    }
    if (dw.rank == 0 && !trace_quiet()) printf("pairwise perturbation starts from %d\n", iter);
    if (trace_sink()) trace_sink()->events.push_back({1, iter});
    alsCP_PP_partupdate_sub(V, W, grad_W, dW.data(), F, tol, tol_init, timelimit, maxiter, update_percentage, st_time,
                            lambda, ratio_step, Plot_File, gradnorm, iter, resprint, bench, dw);
  }
  if (dw.rank == 0 && !trace_quiet()) {
    printf("\nIter = %d Final grad norm %E \n", iter, gradnorm);
    printf("tf took %lf seconds\n", synced_time(dw) - st_time);
  }
  if (!bench && Plot_File.is_open()) Plot_File.close();
  return iter != maxiter + 1;
}
