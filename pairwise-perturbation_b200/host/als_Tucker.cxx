// als_Tucker.cxx -- Tucker drivers on the ppx C ABI; control flow and printed lines follow
// /root/reference/als_Tucker.cxx (cited per block).
#include "als_Tucker.h"
#include <cmath>

namespace {

const char *kCsvHeaderT = "[dim],[iter],[diffnorm],[tol],[pp_update],[diffV],[dtime]";

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
// ---- multi-GPU: V is sharded along mode 0, rows [row_begin, row_end) of it per rank; every factor is
// REPLICATED in full (the factor update is an eigenproblem every rank solves redundantly).  A tensor that still
// carries mode 0 holds the local rows and is complete; contracting mode 0 uses the local rows of W_0 and leaves a
// partial sum that stays partial down the chain (every later TTM is linear) until it is all-reduced.
bool sharded(World &dw) { return dw.np > 1; }
int64_t rows0(World &dw) { return dw.row_end - dw.row_begin; }

// all-gather along mode 0 (the fastest index): local rows pasted into a zeroed full-size tensor, then summed over ranks
Tensor<> gather_mode0(Tensor<> &Yloc, World &dw) {
  int64_t lens[16];
  for (int i = 0; i < Yloc.order; i++) lens[i] = Yloc.lens[i];
  lens[0] = dw.shard_global;
  Tensor<> full(Yloc.order, lens, dw);  // zero-initialised
  const int64_t cols = Yloc.size / Yloc.lens[0];
  PPXCK(dw, ppx_memcpy2d_d2d(dw.ctx, full.data + dw.row_begin, sizeof(double) * lens[0], Yloc.data,
                             sizeof(double) * Yloc.lens[0], sizeof(double) * Yloc.lens[0], (size_t)cols));
  dw.allreduce(full.data, full.size);
  return full;
}

// Y[.., q, ..] = sum_x T[.., x, ..] Wx[x, q]: the rank replaces mode x in place (als_Tucker.cxx:102,224,464-465).
// Multi-GPU: a contraction of mode 0 sums over the local rows only (partial result).
Tensor<> ttm_mode(Tensor<> &T, int x, Matrix<> &Wx, World &dw) {
  int64_t lens[16];
  for (int i = 0; i < T.order; i++) lens[i] = T.lens[i];
  lens[x] = Wx.ncol;
  Tensor<> out(T.order, lens, dw, false);
  if (sharded(dw) && x == 0) {
    if (T.lens[0] != rows0(dw)) throw std::runtime_error("ttm_mode: mode 0 of a sharded tensor must hold the local rows");
    PPXCK(dw, ppx_ttm(dw.ctx, T.data, T.lens, T.order, 0, Wx.data + dw.row_begin, Wx.nrow, (int)Wx.ncol, out.data));
  } else {
    PPXCK(dw, ppx_ttm(dw.ctx, T.data, T.lens, T.order, x, Wx.data, Wx.nrow, (int)Wx.ncol, out.data));
  }
  return out;
}

// W_i <- leading r eigenvectors of Gram(Y_(i))  == MTM.svd(U,S,VT,r); W[i]=U  (als_Tucker.cxx:399-406).
// Multi-GPU: Y arrives as computed from the local slab -- for i != 0 a partial sum (all-reduced here), for i == 0 the
// local rows (gathered here); the Gram and the eigenproblem are then replicated.  `complete`: Y is already global.
void factor_from_unfolding(Tensor<> &Y, int i, int r, Matrix<> &Wi, World &dw, bool complete = false) {
  if (sharded(dw) && !complete) {
    if (i != 0) dw.allreduce(Y.data, Y.size);
    else Y = gather_mode0(Y, dw);
  }
  Matrix<> MTM = unroll_tensor_contraction(Y, i);
  Matrix<> U(Y.lens[i], r, dw);
  // warm start from the eigenvectors this mode had one sweep earlier (same result, fewer Jacobi sweeps)
  World::EigBasis &eb = dw.eig_basis_for(i, Y.lens[i], r);
  PPXCK(dw, ppx_sym_eig_topk_warm(dw.ctx, MTM.data, Y.lens[i], r, U.data, nullptr, eb.data, eb.valid ? 1 : 0));
  eb.valid = true;
  Wi = std::move(U);
}

void sign_align(Matrix<> &Wi, Matrix<> &ref, World &dw) {  // als_Tucker.cxx:632-643
  PPXCK(dw, ppx_sign_align(dw.ctx, Wi.data, ref.data, Wi.nrow, (int)Wi.ncol));
}

// ||core x_j W_j^T - V||_F  (als_Tucker.cxx:295-310)
double tucker_residual(Tensor<> &V, Tensor<> &core, Matrix<> *W, World &dw) {
  if (trace_sink() && trace_sink()->skip_residual) return -1.0;
  const int N = V.order;
  vector<Matrix<>> W_T;
  for (int i = 0; i < N; i++) {
    W_T.emplace_back(W[i].ncol, W[i].nrow, dw);
    PPXCK(dw, ppx_transpose(dw.ctx, W[i].data, W[i].nrow, W[i].ncol, W_T[i].data));
  }
  // multi-GPU: only the local rows of mode 0 are reconstructed (W_0^T restricted to the columns row_begin..row_end,
  // a contiguous block of the R x s transposed factor) and the squared difference is summed over ranks
  Tensor<> cur;
  Tensor<> *src = &core;
  for (int i = 0; i < N; i++) {
    int64_t lens[16];
    for (int k = 0; k < N; k++) lens[k] = src->lens[k];
    const bool loc = sharded(dw) && i == 0;
    const int64_t ncol = loc ? rows0(dw) : W_T[i].ncol;
    const double *wt = loc ? W_T[i].data + W_T[i].nrow * dw.row_begin : W_T[i].data;
    lens[i] = ncol;
    Tensor<> nxt(N, lens, dw, false);
    PPXCK(dw, ppx_ttm(dw.ctx, src->data, src->lens, N, i, wt, W_T[i].nrow, (int)ncol, nxt.data));
    cur = std::move(nxt);
    src = &cur;
  }
  PPXCK(dw, ppx_diff_sqnorm(dw.ctx, cur.data, V.data, V.size, dw.scal_dev));
  dw.allreduce(dw.scal_dev, 1);
  double v;
  dw.fetch(dw.scal_dev, &v, 1);
  return std::sqrt(v);
}

// Y of leaf mode i from the dimension tree (als_Tucker.cxx:356-394 / 581-619); order-3 extension as in als_CP.cxx
Tensor<> leaf_ttmc(map<string, Tensor<>> &ttmc_map, map<string, string> &parent, map<string, string> &sibling,
                   Tensor<> &V, Matrix<> *W, int i, World &dw) {
  const string a(1, (char)('a' + i));
  const string par = parent[a];
  Tensor<> *src = &V;
  if ((int)par.size() != V.order) {
    if (ttmc_map.find(par) == ttmc_map.end()) ttmc_map_DT(ttmc_map, parent, sibling, V, W, par, dw);
    src = &ttmc_map[par];
  }
  Tensor<> cur;
  for (char c : par) {
    if (c == a[0]) continue;
    Tensor<> nxt = ttm_mode(*src, c - 'a', W[c - 'a'], dw);
    cur = std::move(nxt);
    src = &cur;
  }
  return cur;
}

void log_row_t(Tensor<> &V, int iter, double diffnorm, double tol, int pp_update, double diffV, double dtime,
               ofstream &Plot_File, World &dw) {
  if (trace_sink()) trace_sink()->rows.push_back({(double)iter, diffnorm, pp_update, diffV, dtime});
  if (dw.rank != 0) return;
  if (!trace_quiet())
    cout << "  [dim]=  " << (dw.np > 1 && dw.shard_mode == 0 ? dw.shard_global : V.lens[0]) << "  [iter]=  " << iter << "  [diffnorm]  " << diffnorm << "  [tol]  " << tol
         << "  [pp_update]  " << pp_update << "  [diffV]  " << diffV << "  [dtime]  " << dtime << "\n";
  if (Plot_File.is_open()) {
    Plot_File << (dw.np > 1 && dw.shard_mode == 0 ? dw.shard_global : V.lens[0]) << "," << iter << "," << diffnorm << "," << tol << "," << pp_update << "," << diffV << ","
              << dtime << "\n";
    if (iter % 100 == 0 && iter != 0) Plot_File << endl;
  }
}

}  // namespace

namespace {
// HOSVD factors of a tensor sharded along mode 0 (als_Tucker.cxx:12-40 on every rank's slab).  Modes i != 0: the Gram
// of the mode-i unfolding sums over mode 0, so the local Grams are all-reduced.  Mode 0: MTM_0[p,q] pairs rows that
// live on different ranks.  The sum over the other modes is split along the LAST mode instead: rank k takes the range
// [tb_k, te_k) of it and receives, from every rank, the rows that rank holds of that range (one personalised exchange,
// ppx_alltoallv: each rank sends and receives (np-1)/np of its slab -- contiguous blocks, since the last mode is the
// slowest index), pastes them into a full-height panel, forms the Gram of its panel and the partial Grams are
// all-reduced.  (Round 1 moved the WHOLE tensor through zero-padded all-reduces and formed the panels' Grams one rank
// at a time: hosvd took 116 ms on two GPUs against 61 ms on one.)
Matrix<> gram_mode0_sharded(Tensor<> &T, World &dw) {
  const int N = T.order;
  const int64_t s0 = dw.shard_global, r0 = rows0(dw);
  const int64_t last = T.lens[N - 1];
  const int64_t mid = T.size / r0 / last;  // product of the middle modes
  std::vector<int64_t> rb(dw.np), re(dw.np), tb(dw.np), te(dw.np);
  for (int k = 0; k < dw.np; k++) {
    PPXCK(dw, ppx_shard_range(s0, dw.np, k, &rb[k], &re[k]));
    PPXCK(dw, ppx_shard_range(last, dw.np, k, &tb[k], &te[k]));
  }
  const int me = dw.rank;
  const int64_t ct = te[me] - tb[me];  // my share of the last mode (may be empty when last < np)
  Matrix<> MTM(s0, s0, dw);
  // bounded staging: the exchange runs in rounds over sub-ranges of every rank's share (<= ~2^27 doubles received per round)
  int64_t ct_max = 0;
  for (int k = 0; k < dw.np; k++) ct_max = std::max(ct_max, te[k] - tb[k]);
  int64_t step = std::max<int64_t>(1, ((int64_t)1 << 27) / std::max<int64_t>(1, s0 * mid));
  for (int64_t off = 0; off < ct_max; off += step) {
    const int64_t my_n = std::max<int64_t>(0, std::min(step, ct - off));  // slices of the last mode I take this round
    std::vector<Tensor<>> stage(dw.np);
    std::vector<const double *> sb(dw.np, nullptr);
    std::vector<double *> rbuf(dw.np, nullptr);
    std::vector<int64_t> sc(dw.np, 0), rc(dw.np, 0);
    for (int k = 0; k < dw.np; k++) {
      const int64_t n_k = std::max<int64_t>(0, std::min(step, (te[k] - tb[k]) - off));  // what rank k takes this round
      if (n_k > 0) {
        sb[k] = T.data + r0 * mid * (tb[k] + off);
        sc[k] = r0 * mid * n_k;
      }
      if (my_n > 0) {
        const int64_t rows_k = re[k] - rb[k];
        int64_t l[1] = {rows_k * mid * my_n};
        stage[k] = Tensor<>(1, l, dw, false);
        rbuf[k] = stage[k].data;
        rc[k] = l[0];
      }
    }
    PPXCK(dw, ppx_alltoallv(dw.ctx, sb.data(), sc.data(), rbuf.data(), rc.data()));
    if (my_n == 0) continue;
    int64_t lens_p[16];
    for (int k = 0; k < N; k++) lens_p[k] = T.lens[k];
    lens_p[0] = s0;
    lens_p[N - 1] = my_n;
    Tensor<> full(N, lens_p, dw, false);
    for (int k = 0; k < dw.np; k++) {
      const int64_t rows_k = re[k] - rb[k];
      PPXCK(dw, ppx_memcpy2d_d2d(dw.ctx, full.data + rb[k], sizeof(double) * s0, stage[k].data, sizeof(double) * rows_k,
                                 sizeof(double) * rows_k, (size_t)(mid * my_n)));
    }
    Matrix<> part = unroll_tensor_contraction(full, 0);
    PPXCK(dw, ppx_axpby(dw.ctx, 1.0, part.data, 1.0, MTM.data, MTM.size));
  }
  dw.allreduce(MTM.data, MTM.size);
  return MTM;
}

void hosvd_sharded(Tensor<> &T, Matrix<> *factor_matrices, int *ranks, World &dw) {
  const int N = T.order;
  for (int i = 0; i < N; i++) {
    Matrix<> MTM;
    if (i != 0) {
      MTM = unroll_tensor_contraction(T, i);
      dw.allreduce(MTM.data, MTM.size);
    } else {
      MTM = gram_mode0_sharded(T, dw);
    }
    Matrix<> U(MTM.nrow, ranks[i], dw);
    // HOSVD is an initialisation: always a cold solve (whatever an earlier decomposition left in this World is
    // unrelated); the basis it leaves warm-starts the first HOOI sweep
    World::EigBasis &eb = dw.eig_basis_for(i, MTM.nrow, ranks[i]);
    PPXCK(dw, ppx_sym_eig_topk_warm(dw.ctx, MTM.data, MTM.nrow, ranks[i], U.data, nullptr, eb.data, 0));
    eb.valid = true;
    factor_matrices[i] = std::move(U);
  }
}
}  // namespace

void get_factor_matrices(Tensor<> &T, Matrix<> *factor_matrices, int ranks[], World &dw) {
  for (int i = 0; i < T.order; i++) factor_from_unfolding(T, i, ranks[i], factor_matrices[i], dw);
}

Tensor<> get_core_tensor(Tensor<> &T, Matrix<> *factor_matrices, int ranks[], World &dw) {
  (void)ranks;
  Tensor<> core;
  TTMc(core, T, factor_matrices, -1, dw);
  return core;
}

void hosvd(Tensor<> &T, Tensor<> &core, Matrix<> *factor_matrices, int *ranks, World &dw) {
  if (sharded(dw)) {
    hosvd_sharded(T, factor_matrices, ranks, dw);
  } else {
    get_factor_matrices(T, factor_matrices, ranks, dw);
  }
  core = get_core_tensor(T, factor_matrices, ranks, dw);
}

void TTMc(Tensor<> &Y, Tensor<> &V, Matrix<> *W, int i, World &dw) {
  Tensor<> *src = &V;
  Tensor<> cur;
  bool any = false;
  for (int index = 0; index < V.order; index++) {
    if (index == i) continue;
    Tensor<> nxt = ttm_mode(*src, index, W[index], dw);
    cur = std::move(nxt);
    src = &cur;
    any = true;
  }
  if (any) Y = std::move(cur);
  else Y = V;
  // multi-GPU: the core (i == -1) is summed over ranks here; for i >= 0 the caller gets the contribution of the
  // local slab (partial sum for i != 0, local rows for i == 0), which factor_from_unfolding completes
  if (sharded(dw) && i < 0 && V.lens[0] == rows0(dw)) dw.allreduce(Y.data, Y.size);
}

bool alsTucker(Tensor<> &V, Tensor<> &core, Matrix<> *W, double tol, double timelimit, int maxiter, World &dw) {
  // als_Tucker.cxx:120-176: no dimension tree, a full TTMc per mode
  double st_time = synced_time(dw);
  int iter;
  Tensor<> core_prev(core);
  double diffnorm = 0;
  for (iter = 0; iter <= maxiter; iter++) {
    if ((iter % 100 == 0 && iter != 0) || iter == maxiter) {
      TTMc(core, V, W, -1, dw);
      diffnorm = std::fabs(core.norm2() - core_prev.norm2());
      if (trace_sink()) trace_sink()->rows.push_back({(double)iter, diffnorm, 0, -1.0, 0.0});
      if (dw.rank == 0 && !trace_quiet())
        cout << "  [dim]=  " << (dw.np > 1 && dw.shard_mode == 0 ? dw.shard_global : V.lens[0]) << "  [iter]=  " << iter << "  [diffnorm]  " << diffnorm << "  [tol]  "
             << tol << "\n";
      if (diffnorm < tol || synced_time(dw) - st_time > timelimit) break;
      core_prev = core;
    }
    for (int i = 0; i < V.order; i++) {
      Tensor<> Y;
      TTMc(Y, V, W, i, dw);
      factor_from_unfolding(Y, i, (int)core.lens[i], W[i], dw);
    }
    if (iter % 10 == 0 && dw.rank == 0 && !trace_quiet()) printf(".");
  }
  if (dw.rank == 0 && !trace_quiet()) {
    printf("\nIter = %d Final Diff norm %E \n", iter, diffnorm);
    printf("tf took %lf seconds\n", synced_time(dw) - st_time);
  }
  return iter != maxiter + 1;
}

void ttmc_map_DT(map<string, Tensor<>> &ttmc_map, map<string, string> &parent, map<string, string> &sibling,
                 Tensor<> &V, Matrix<> *W, string args, World &dw) {
  // als_Tucker.cxx:178-230
  if (ttmc_map.find(args) != ttmc_map.end()) return;
  const string par = parent[args];
  Tensor<> *src = &V;
  if ((int)par.size() != V.order) {
    if (ttmc_map.find(par) == ttmc_map.end()) ttmc_map_DT(ttmc_map, parent, sibling, V, W, par, dw);
    src = &ttmc_map[par];
  }
  Tensor<> cur;
  for (char c : sibling[args]) {
    Tensor<> nxt = ttm_mode(*src, c - 'a', W[c - 'a'], dw);
    cur = std::move(nxt);
    src = &cur;
  }
  ttmc_map[args] = std::move(cur);
}

bool alsTucker_DT(Tensor<> &V, Tensor<> &core, Matrix<> *W, double tol, double timelimit, int maxiter,
                  ofstream &Plot_File, int resprint, bool bench, World &dw) {
  // als_Tucker.cxx:240-424
  cout.precision(13);
  const int N = V.order;
  if (!bench && Plot_File.is_open()) Plot_File << kCsvHeaderT << "\n";
  double st_time = synced_time(dw);
  int iter;
  Tensor<> core_prev(core);
  double diffnorm = 1000, diffnorm_V = 1000;
  map<string, Tensor<>> ttmc_map;
  map<string, string> parent, sibling;
  Construct_Dimension_Tree(parent, sibling, 0, N - 1);
  Tensor<> Y_end;
  for (iter = 0; iter <= maxiter; iter++) {
    if ((iter % resprint == 0 && iter != 0) || iter == 1 || iter == maxiter) {  // :288-338
      const double st_time1 = synced_time(dw);
      TTMc(core, V, W, -1, dw);
      diffnorm = std::fabs(core.norm2() - core_prev.norm2());
      diffnorm_V = tucker_residual(V, core, W, dw);
      st_time += synced_time(dw) - st_time1;
      const double dtime = wall_time() - st_time;
      if (!bench) {
        log_row_t(V, iter, diffnorm, tol, 0, diffnorm_V, dtime, Plot_File, dw);
      } else {
        if (trace_sink()) trace_sink()->bench_times.push_back(dtime);
        if (dw.rank == 0) {
          if (!trace_quiet()) cout << "  [dimension tree step time]  " << dtime << "\n";
          if (Plot_File.is_open()) Plot_File << "[DTtime]" << "," << dtime << "\n";
        }
      }
      if (diffnorm < tol || wall_time() - st_time > timelimit) break;
      core_prev = core;
    }
    ttmc_map.clear();
    for (int i = 0; i < N; i++) {
      Tensor<> Y = leaf_ttmc(ttmc_map, parent, sibling, V, W, i, dw);
      factor_from_unfolding(Y, i, (int)core.lens[i], W[i], dw);  // :399-406
      if (i == N - 1) Y_end = std::move(Y);
    }
    core = ttm_mode(Y_end, N - 1, W[N - 1], dw);  // :408
    if (trace_sink()) trace_sink()->sweeps.push_back({0, iter});
    if (iter % 10 == 0 && dw.rank == 0 && !trace_quiet()) printf(".");
  }
  if (dw.rank == 0 && !trace_quiet()) {
    printf("\nIter = %d Final Diff norm %E \n", iter, diffnorm);
    printf("tf took %lf seconds\n", synced_time(dw) - st_time);
  }
  if (!bench && Plot_File.is_open()) Plot_File.close();
  return iter != maxiter + 1;
}

void Build_ttmc_map(map<string, Tensor<>> &ttmc_map, Tensor<> &V, Matrix<> *W, const char *args_c, World &dw) {
  // als_Tucker.cxx:426-466: key = contracted modes; the operator keeps order N with R at the contracted positions
  const string args(args_c);
  Tensor<> *src = &V;
  if (args.size() > 1) {
    const string prefix = args.substr(0, args.size() - 1);
    if (ttmc_map.find(prefix) == ttmc_map.end()) Build_ttmc_map(ttmc_map, V, W, prefix.c_str(), dw);
    src = &ttmc_map[prefix];
  }
  const int x = args.back() - 'a';
  ttmc_map[args] = ttm_mode(*src, x, W[x], dw);
}

namespace {

void build_tucker_pp_operators(map<string, Tensor<>> &ttmc_map, Tensor<> &V, Matrix<> *W, World &dw) {
  // als_Tucker.cxx:742-760; intermediates that are neither pair operators nor singles are dropped afterwards
  const int N = V.order;
  const string seq = all_modes(N);
  ttmc_map.clear();
  for (int ii = 0; ii < N; ii++)
    for (int jj = ii + 1; jj < N; jj++) Build_ttmc_map(ttmc_map, V, W, without(seq, ii, jj).c_str(), dw);
  for (int ii = 0; ii < N; ii++) Build_ttmc_map(ttmc_map, V, W, without(seq, ii).c_str(), dw);
  for (auto it = ttmc_map.begin(); it != ttmc_map.end();) {
    if ((int)it->first.size() < N - 2) it = ttmc_map.erase(it);
    else ++it;
  }
  // multi-GPU: an operator that contracted mode 0 is a partial sum over the local slab
  if (sharded(dw))
    for (auto &kv : ttmc_map)
      if (kv.first.find('a') != string::npos) dw.allreduce(kv.second.data, kv.second.size);
}

// Y_i = Y_i(W_init) + sum_{j != i} T^(i,j) x_j dW_j   (als_Tucker.cxx:828-860)
Tensor<> tucker_pp_corrected(map<string, Tensor<>> &ttmc_map, Matrix<> *dW, int i, int N, World &dw) {
  const string seq = all_modes(N);
  // multi-GPU: for i != 0 everything is replicated except the j == 0 term, whose operator keeps the local rows of
  // mode 0 -- that term is a partial sum and is all-reduced on its own; for i == 0 all terms act on the local rows
  Tensor<> Y = ttmc_map[without(seq, i)];
  for (int j = 0; j < N; j++) {
    if (j == i) continue;
    Tensor<> &T = ttmc_map[without(seq, std::min(i, j), std::max(i, j))];
    if (sharded(dw) && j == 0) {
      Tensor<> term = ttm_mode(T, 0, dW[0], dw);
      dw.allreduce(term.data, term.size);
      PPXCK(dw, ppx_axpby(dw.ctx, 1.0, term.data, 1.0, Y.data, Y.size));
    } else {
      PPXCK(dw, ppx_ttm_acc(dw.ctx, T.data, T.lens, N, j, dW[j].data, dW[j].nrow, (int)dW[j].ncol, Y.data));
    }
  }
  return Y;
}

// the switching scalars ||dW_i|| / ||W_i||  (als_Tucker.cxx:648-656, 722-729)
void dw_ratios(Matrix<> *W, Matrix<> *dW, int N, double *ratio, World &dw) {
  const double *xs[32];
  int64_t ns[32];
  for (int i = 0; i < N; i++) {
    xs[2 * i] = dW[i].data;
    xs[2 * i + 1] = W[i].data;
    ns[2 * i] = dW[i].size;
    ns[2 * i + 1] = W[i].size;
  }
  for (int b = 0; b < 2 * N; b += 16)
    PPXCK(dw, ppx_sqnorms(dw.ctx, xs + b, ns + b, std::min(16, 2 * N - b), dw.scal_dev + b));
  double h[32];
  dw.fetch(dw.scal_dev, h, 2 * N);
  for (int i = 0; i < N; i++) ratio[i] = std::fabs(std::sqrt(h[2 * i]) / std::sqrt(h[2 * i + 1]));
}

}  // namespace

void alsTucker_DT_sub(Tensor<> &V, Tensor<> &core, Tensor<> &core_prev, Matrix<> *W, Matrix<> *dW, double tol,
                      double tol_init, double timelimit, int maxiter, double &st_time, ofstream &Plot_File,
                      double &diffnorm, int &iter, int resprint, World &dw) {
  // als_Tucker.cxx:476-669
  const int N = V.order;
  vector<Matrix<>> W_prev;  // zeros: the first sweep flips every column (sign(0) = -1, :636-641) and cannot switch
  for (int i = 0; i < N; i++) W_prev.emplace_back(W[i].nrow, W[i].ncol, dw);
  double diffnorm_V = 1000;
  map<string, Tensor<>> ttmc_map;
  map<string, string> parent, sibling;
  Construct_Dimension_Tree(parent, sibling, 0, N - 1);
  Tensor<> Y_end;
  for (; iter <= maxiter; iter++) {
    if ((iter % resprint == 0 && iter != 0) || iter == 1 || iter == maxiter) {  // :521-564
      const double st_time1 = synced_time(dw);
      TTMc(core, V, W, -1, dw);
      diffnorm = std::fabs(core.norm2() - core_prev.norm2());
      diffnorm_V = tucker_residual(V, core, W, dw);
      st_time += synced_time(dw) - st_time1;
      const double dtime = wall_time() - st_time;
      log_row_t(V, iter, diffnorm, tol, 0, diffnorm_V, dtime, Plot_File, dw);
      if (diffnorm < tol || wall_time() - st_time > timelimit) break;
      core_prev = core;
    }
    ttmc_map.clear();
    for (int i = 0; i < N; i++) {
      Tensor<> Y = leaf_ttmc(ttmc_map, parent, sibling, V, W, i, dw);
      factor_from_unfolding(Y, i, (int)core.lens[i], W[i], dw);
      sign_align(W[i], W_prev[i], dw);  // :632-643
      if (i == N - 1) Y_end = std::move(Y);
    }
    core = ttm_mode(Y_end, N - 1, W[N - 1], dw);  // :645
    if (trace_sink()) trace_sink()->sweeps.push_back({0, iter});
    for (int i = 0; i < N; i++)  // :649-651
      PPXCK(dw, ppx_diff_update(dw.ctx, W[i].data, W_prev[i].data, dW[i].data, W[i].size, dw.scal_dev + 2 * i));
    double h[32];
    dw.fetch(dw.scal_dev, h, 2 * N);
    int num_dw_break = 0;
    for (int i = 0; i < N; i++)
      if (std::fabs(std::sqrt(h[2 * i]) / std::sqrt(h[2 * i + 1])) < tol_init) num_dw_break++;
    if (num_dw_break == N) return;
    if (iter % 10 == 0 && dw.rank == 0 && !trace_quiet()) printf(".");
  }
}

void alsTucker_PP_sub(Tensor<> &V, Tensor<> &core, Tensor<> &core_prev, Matrix<> *W, Matrix<> *dW, double tol,
                      double tol_init, double timelimit, int maxiter, double &st_time, ofstream &Plot_File,
                      double &diffnorm, int &iter, int resprint, bool bench, World &dw) {
  // als_Tucker.cxx:679-896
  const int N = V.order;
  double dtime_first = 0;
  const int init_iter = iter;
  double diffnorm_V = 1000;
  vector<Matrix<>> W_init(N);
  map<string, Tensor<>> ttmc_map;
  Tensor<> Y_end;
  for (; iter <= maxiter; iter++) {
    int num_dw_break = 0;
    if (!bench) {  // :722-730
      double ratio[16];
      dw_ratios(W, dW, N, ratio, dw);
      for (int i = 0; i < N; i++)
        if (ratio[i] > tol_init) num_dw_break++;
    }
    if (iter == init_iter || num_dw_break > 0) {  // :733-761 (no 15-sweep cap here)
      if (num_dw_break > 0) return;
      for (int j = 0; j < N; j++) {
        W_init[j] = W[j];
        dW[j] = Matrix<>(W[j].nrow, W[j].ncol, dw);
      }
      build_tucker_pp_operators(ttmc_map, V, W, dw);
      if (trace_sink()) trace_sink()->sweeps.push_back({2, iter});
    }
    if ((iter % resprint == 0 && iter != 0) || iter == 1 || iter == maxiter || iter == init_iter) {  // :763-822
      const double st_time1 = synced_time(dw);
      TTMc(core, V, W, -1, dw);
      diffnorm = std::fabs(core.norm2() - core_prev.norm2());
      diffnorm_V = tucker_residual(V, core, W, dw);
      st_time += synced_time(dw) - st_time1;
      const double dtime = wall_time() - st_time;
      if (!bench) {
        log_row_t(V, iter, diffnorm, tol, 1, diffnorm_V, dtime, Plot_File, dw);
      } else if (iter != maxiter) {
        dtime_first = dtime;
        st_time = wall_time();
      } else {
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
      if (diffnorm < tol || wall_time() - st_time > timelimit || iter == maxiter) break;
      core_prev = core;
    }
    for (int i = 0; i < N; i++) {  // :824-890
      Tensor<> Y = tucker_pp_corrected(ttmc_map, dW, i, N, dw);
      factor_from_unfolding(Y, i, (int)core.lens[i], W[i], dw, /*complete=*/i != 0);
      sign_align(W[i], W_init[i], dw);  // :874-885
      PPXCK(dw, ppx_memcpy_d2d(dw.ctx, dW[i].data, W[i].data, sizeof(double) * W[i].size));
      PPXCK(dw, ppx_axpby(dw.ctx, -1.0, W_init[i].data, 1.0, dW[i].data, dW[i].size));  // dW = W - W_init (:887)
      if (i == N - 1) Y_end = std::move(Y);
    }
    core = ttm_mode(Y_end, N - 1, W[N - 1], dw);  // :891
    if (trace_sink()) trace_sink()->sweeps.push_back({1, iter});
  }
  if (bench) iter++;
}

bool alsTucker_PP(Tensor<> &V, Tensor<> &core, Matrix<> *W, double tol, double tol_init, double timelimit,
                  int maxiter, ofstream &Plot_File, int resprint, bool bench, World &dw) {
  // als_Tucker.cxx:906-962
  cout.precision(13);
  const int N = V.order;
  if (!bench && dw.rank == 0 && Plot_File.is_open()) Plot_File << kCsvHeaderT << "\n";
  double st_time = synced_time(dw);
  int iter = 0;
  Tensor<> core_prev(core);
  double diffnorm = 10.;
  vector<Matrix<>> dW;
  for (int j = 0; j < N; j++) dW.emplace_back(W[j].nrow, W[j].ncol, dw);
  while (diffnorm > tol && iter <= maxiter) {
    if (!bench) {
      if (dw.rank == 0 && !trace_quiet()) printf("DT starts from %d\n", iter);
      if (trace_sink()) trace_sink()->events.push_back({0, iter});
      alsTucker_DT_sub(V, core, core_prev, W, dW.data(), tol, tol_init, timelimit, maxiter, st_time, Plot_File,
                       diffnorm, iter, resprint, dw);
    }
    if (dw.rank == 0 && !trace_quiet()) printf("pairwise perturbation starts from %d\n", iter);
    if (trace_sink()) trace_sink()->events.push_back({1, iter});
    alsTucker_PP_sub(V, core, core_prev, W, dW.data(), tol, tol_init, timelimit, maxiter, st_time, Plot_File,
                     diffnorm, iter, resprint, bench, dw);
    if (tol_init > 5e-3) tol_init *= 0.9;  // :947-948
  }
  if (dw.rank == 0 && !trace_quiet()) {
    printf("\nIter = %d Final Diff norm %E \n", iter, diffnorm);
    printf("tf took %lf seconds\n", synced_time(dw) - st_time);
  }
  if (!bench && Plot_File.is_open()) Plot_File.close();
  return iter != maxiter + 1;
}
