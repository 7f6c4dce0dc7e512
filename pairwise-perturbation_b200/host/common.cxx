// common.cxx -- see common.h.  Every function cites the reference lines whose behaviour it reproduces.
#include "common.h"
#include <chrono>
#include <cmath>

TraceSink *&trace_sink() {
  static TraceSink *s = nullptr;
  return s;
}
bool trace_quiet() { return trace_sink() && trace_sink()->quiet; }

double wall_time() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

void vec2str(vector<int> vec, string &seq_out) {
  // common.cxx:10-18: local indices -> letters, '*' (rank) appended
  string s;
  for (int v : vec) s.push_back((char)('a' + v));
  s.push_back('*');
  seq_out = s;
}

void Construct_Dimension_Tree(map<string, string> &parent, map<string, string> &sibling, int start, int end) {
  // common.cxx:225-270: balanced binary split at (start+end)/2; a node is named by the modes it keeps
  if (end <= start) return;
  auto name = [](int lo, int hi) {
    string s;
    for (int i = lo; i <= hi; i++) s.push_back((char)('a' + i));
    return s;
  };
  const string whole = name(start, end);
  const int middle = (end == start + 1) ? start : (start + end) / 2;
  const string left = name(start, middle), right = name(middle + 1, end);
  parent[left] = whole;
  parent[right] = whole;
  sibling[left] = right;
  sibling[right] = left;
  Construct_Dimension_Tree(parent, sibling, start, middle);
  Construct_Dimension_Tree(parent, sibling, middle + 1, end);
}

Tensor<> contract_mode(Tensor<> &in, const string &modes, bool has_rank, char x, Matrix<> &Wx, World &dw) {
  const int k = (int)modes.size();
  const int pos = (int)modes.find(x);
  assert(pos >= 0 && pos < k);
  const int R = (int)Wx.ncol;
  int64_t out_lens[17];
  int n = 0;
  for (int i = 0; i < k; i++)
    if (i != pos) out_lens[n++] = in.lens[i];
  out_lens[n++] = R;
  Tensor<> out(n, out_lens, dw, false);
  if (!has_rank) {
    PPXCK(dw, ppx_ttm_first(dw.ctx, in.data, in.lens, k, pos, Wx.data, Wx.nrow, R, out.data));
  } else {
    PPXCK(dw, ppx_mttv(dw.ctx, in.data, in.lens, k, pos, Wx.data, Wx.nrow, R, out.data));
  }
  return out;
}

void mttkrp_map_DT(map<string, Tensor<>> &mttkrp_map, map<string, string> &parent, map<string, string> &sibling,
                   Tensor<> &V, Matrix<> *W, string args, World &dw) {
  // common.cxx:20-133
  if (mttkrp_map.find(args) != mttkrp_map.end()) return;
  const string par = parent[args];
  const string sib = sibling[args];
  if ((int)par.size() == V.order) {
    // first-level node (common.cxx:29-88): start from V, contract the sibling modes left to right; the first is
    // the GEMM (common.cxx:56), the rest are Hadamard-batched (common.cxx:83).
    // The sibling modes of a tree node are adjacent, so the whole chain is ONE GEMM against the Khatri-Rao rows of
    // their factors (ppx_ttm_multi): the level-1 tensors the reference materialises (10.8 GB each at N=4, s=300,
    // R=50) have a single consumer here and are never written.
    const int R = (int)W[0].ncol;
    const int x_first = sib[0] - 'a';
    const int n = (int)sib.size();
    int64_t out_lens[17];
    int k = 0;
    for (char c : args) out_lens[k++] = V.lens[c - 'a'];
    out_lens[k++] = R;
    Tensor<> out(k, out_lens, dw, false);
    const double *wp[8];
    int64_t ld[8];
    bool adjacent = n <= 8;
    for (int j = 0; j < n && adjacent; j++) {
      adjacent = (sib[j] - 'a') == x_first + j;
      wp[j] = W[sib[j] - 'a'].data;
      ld[j] = W[sib[j] - 'a'].nrow;
    }
    if (adjacent) {
      PPXCK(dw, ppx_ttm_multi(dw.ctx, V.data, V.lens, V.order, x_first, n, wp, ld, R, out.data));
      mttkrp_map[args] = std::move(out);
      return;
    }
    string modes = par;
    Tensor<> cur = contract_mode(V, modes, false, sib[0], W[sib[0] - 'a'], dw);
    modes.erase(modes.find(sib[0]), 1);
    for (size_t j = 1; j < sib.size(); j++) {
      Tensor<> nxt = contract_mode(cur, modes, true, sib[j], W[sib[j] - 'a'], dw);
      modes.erase(modes.find(sib[j]), 1);
      cur = std::move(nxt);
    }
    mttkrp_map[args] = std::move(cur);
    return;
  }
  if (mttkrp_map.find(par) == mttkrp_map.end()) mttkrp_map_DT(mttkrp_map, parent, sibling, V, W, par, dw);  // :89-91
  // deeper node (common.cxx:94-131): start from the cached parent
  string modes = par;
  Tensor<> *src = &mttkrp_map[par];
  Tensor<> cur;
  for (size_t j = 0; j < sib.size(); j++) {
    Tensor<> nxt = contract_mode(*src, modes, true, sib[j], W[sib[j] - 'a'], dw);
    modes.erase(modes.find(sib[j]), 1);
    cur = std::move(nxt);
    src = &cur;
  }
  mttkrp_map[args] = std::move(cur);
}

void build_V(Tensor<> &V, Matrix<> *W, int order, World &dw) {
  // common.cxx:135-197: V = [[W_0 .. W_{N-1}]]; one fused kernel instead of N-1 growing temporaries
  int64_t lens[16];
  const double *ptrs[16];
  for (int i = 0; i < order; i++) {
    lens[i] = W[i].nrow;
    ptrs[i] = W[i].data;
  }
  Tensor<> out(order, lens, dw, false);
  PPXCK(dw, ppx_cp_reconstruct(dw.ctx, lens, order, ptrs, (int)W[0].ncol, out.data));
  V = std::move(out);
}

// ---- low-rank update (common.cxx:760-786) -----------------------------------------------------------------------
void matrixDot(Matrix<> &result, Matrix<> &matrix1, Matrix<> &matrix2) {
  World &dw = *matrix1.wrld;
  result = Matrix<>(matrix1.nrow, matrix2.ncol, dw, false);
  PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 0, (int)matrix1.nrow, (int)matrix2.ncol, (int)matrix1.ncol, 1.0, matrix1.data,
                           matrix1.nrow, matrix2.data, matrix2.nrow, 0.0, result.data, result.nrow));
}

void get_rankR_update_cholesky(int R, Matrix<> &xU, Vector<> &xS, Matrix<> &xVT, Matrix<> &M, Matrix<> &A,
                               Matrix<> &gamma, bool random, uint64_t draw_id) {
  World &dw = *M.wrld;
  const int s = (int)M.nrow, n = (int)M.ncol, r = R;
  // Z = L^-1, gamma = L L^T  (:772-773)
  Matrix<> Z(n, n, dw, false);
  PPXCK(dw, ppx_spd_factor_inverse(dw.ctx, gamma.data, n, Z.data));
  // rhs = M - A gamma  (:774-776)
  Matrix<> rhs(M);
  PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 0, s, n, n, -1.0, A.data, A.nrow, gamma.data, n, 1.0, rhs.data, s));
  // X = rhs L^-T  (solve_tri from the right with L transposed, :777)
  Matrix<> X(s, n, dw, false);
  PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 1, s, n, n, 1.0, rhs.data, s, Z.data, n, 0.0, X.data, s));
  // Q (n x r): leading right singular vectors of X = leading eigenvectors of X^T X (:782), or the randomized range
  Matrix<> G(n, n, dw, false), Q(n, r, dw, false);
  PPXCK(dw, ppx_gram(dw.ctx, X.data, s, s, n, G.data));
  if (!random) {
    PPXCK(dw, ppx_sym_eig_topk(dw.ctx, G.data, n, r, Q.data, nullptr));
  } else {
    // Q = orth(X^T X Omega) by Cholesky QR: Y = G Omega, Y^T Y = C C^T, Q = Y C^-T  (:691-708 reduced, see the oracle)
    Matrix<> Omega(n, r, dw, false), Y(n, r, dw, false), YtY(r, r, dw, false), Ci(r, r, dw, false);
    Omega.fill_random(0, 1, dw.seed, draw_id);
    PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 0, n, r, n, 1.0, G.data, n, Omega.data, n, 0.0, Y.data, n));
    for (int pass = 0; pass < 2; pass++) {  // twice: Cholesky QR loses orthogonality with the square of the condition
      PPXCK(dw, ppx_gram(dw.ctx, Y.data, n, n, r, YtY.data));
      PPXCK(dw, ppx_spd_factor_inverse(dw.ctx, YtY.data, r, Ci.data));
      PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 1, n, r, r, 1.0, Y.data, n, Ci.data, r, 0.0, Q.data, n));
      if (pass == 0) PPXCK(dw, ppx_memcpy_d2d(dw.ctx, Y.data, Q.data, sizeof(double) * Q.size));
    }
  }
  // xU = X Q (= U diag(s)), xS = 1, xVT = Q^T L^-1  (:785)
  xU = Matrix<>(s, r, dw, false);
  PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 0, s, r, n, 1.0, X.data, s, Q.data, n, 0.0, xU.data, s));
  xVT = Matrix<>(r, n, dw, false);
  PPXCK(dw, ppx_gemm_small(dw.ctx, 1, 0, r, n, n, 1.0, Q.data, n, Z.data, n, 0.0, xVT.data, r));
  xS = Vector<>(r, dw);
  std::vector<double> ones((size_t)r, 1.0);
  xS.write_all(ones.data());
}

// ---- input generators ------------------------------------------------------------------------------------------
void laplacian_tensor(Tensor<> &V, int N, int s, bool sparse_V, World &dw) {
  // common.cxx:575-642: sum over the d = N/2 index pairs of D on one pair and identities on the others.  The
  // reference assembles it from identity tensors with d+1 contractions; the closed form is filled in place.
  if (sparse_V) throw std::runtime_error("sparse tensors are not supported (never exercised by the reference)");
  if (N < 2 || N % 2) throw std::runtime_error("laplacian_tensor: the order must be even");
  int64_t lens[16];
  for (int i = 0; i < N; i++) lens[i] = s;
  V = Tensor<>(N, lens, dw, false);
  PPXCK(dw, ppx_fill_laplacian(dw.ctx, V.data, N / 2, s));
}

void fold_unfold(Tensor<> &X, Tensor<> &Y) {
  // common.cxx:870-882: same global (first-index-fastest) order, different mode grouping -- a plain copy
  if (X.size != Y.size) throw std::runtime_error("fold_unfold: sizes differ");
  PPXCK(*X.wrld, ppx_memcpy_d2d(X.wrld->ctx, Y.data, X.data, sizeof(double) * X.size));
}

double host_u01(uint64_t seed, uint64_t tensor_id, uint64_t index) {
  uint64_t z = index + seed * 0x9E3779B97F4A7C15ULL + tensor_id * 0xD1B54A32D192ED03ULL;
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

double collinearity(const std::vector<double> &v1, const std::vector<double> &v2) {
  // common.cxx:297-302
  double ip = 0, n1 = 0, n2 = 0;
  for (size_t i = 0; i < v1.size(); i++) {
    ip += v1[i] * v2[i];
    n1 += v1[i] * v1[i];
    n2 += v2[i] * v2[i];
  }
  return ip / (std::sqrt(n1) * std::sqrt(n2));
}

Tensor<> Gen_collinearity(int *lens, int dim, int R, double col_min, double col_max, World &dw) {
  // common.cxx:361-423: R rank-one terms whose mode-j vectors have pairwise collinearity in [col_min, col_max]
  // (redrawn until they do), weights lambda_i = 0.2 + 0.6 (i+1)/R.  The vectors are short, so the rejection loop runs
  // on the host with the counter-based generator (draw k of the run uses tensor id 1000 + k); the sum of the R outer
  // products is one fused reconstruction on the device.
  vector<vector<vector<double>>> vec(R, vector<vector<double>>(dim));
  uint64_t draw = 1000;
  auto fill = [&](vector<double> &v, int n) {
    v.resize(n);
    for (int t = 0; t < n; t++) v[t] = host_u01(dw.seed, draw, (uint64_t)t);
    draw++;
  };
  for (int i = 0; i < R; i++)
    for (int j = 0; j < dim; j++) fill(vec[i][j], lens[j]);
  for (int j = 0; j < dim; j++)
    for (int i = 1; i < R; i++) {
      bool ok = false;
      int tries = 0;
      while (!ok) {
        int k = 0;
        for (; k < i; k++) {
          const double col = collinearity(vec[i][j], vec[k][j]);
          if (col < col_min || col > col_max) break;
        }
        if (k == i) {
          ok = true;
        } else {
          if (++tries > 100000) throw std::runtime_error("Gen_collinearity: no vector within the collinearity range");
          fill(vec[i][j], lens[j]);
        }
      }
    }
  vector<Matrix<>> W;
  vector<double> host;
  for (int j = 0; j < dim; j++) {
    // several GPUs: this rank's rows of the sharded mode (the vectors themselves are drawn identically everywhere)
    const bool cut = dw.np > 1 && j == dw.shard_mode;
    const int t0 = cut ? (int)dw.row_begin : 0, t1 = cut ? (int)dw.row_end : lens[j];
    const int rows = t1 - t0;
    W.emplace_back((int64_t)rows, (int64_t)R, dw, false);
    host.assign((size_t)rows * R, 0.0);
    for (int i = 0; i < R; i++) {
      const double lambda_ = (j == 0) ? 0.2 + 0.6 / R * (i + 1) : 1.0;  // :409
      for (int t = t0; t < t1; t++) host[(size_t)(t - t0) + (size_t)rows * i] = lambda_ * vec[i][j][t];
    }
    W[j].write_all(host.data());
  }
  Tensor<> X;
  build_V(X, W.data(), dim, dw);
  return X;
}

double cp_residual_norm(Tensor<> &V, Matrix<> *W, int order, World &dw) {
  const double *ptrs[16];
  for (int i = 0; i < order; i++) ptrs[i] = W[i].data;
  PPXCK(dw, ppx_cp_residual(dw.ctx, V.data, V.lens, order, ptrs, (int)W[0].ncol, dw.scal_dev));
  dw.allreduce(dw.scal_dev, 1);
  double v;
  dw.fetch(dw.scal_dev, &v, 1);
  return std::sqrt(v);
}

double gradient_norm(Matrix<> *grad_W, int order, World &dw) {
  const double *xs[16];
  int64_t ns[16];
  for (int i = 0; i < order; i++) {
    xs[i] = grad_W[i].data;
    ns[i] = grad_W[i].size;
  }
  PPXCK(dw, ppx_sqnorms(dw.ctx, xs, ns, order, dw.scal_dev));
  double h[16];
  dw.fetch(dw.scal_dev, h, order);
  double acc = 0;
  for (int i = 0; i < order; i++) acc += h[i];  // sum of norm2()^2, als_CP.cxx:175-180
  return std::sqrt(acc);
}

Matrix<> unroll_tensor_contraction(Tensor<> &T, int i) {
  // common.cxx:205-223
  World &dw = *T.wrld;
  Matrix<> MTM(T.lens[i], T.lens[i], dw);
  PPXCK(dw, ppx_unfold_gram(dw.ctx, T.data, T.lens, T.order, i, MTM.data));
  return MTM;
}

void Normalize(Matrix<> *W, int N, World &dw) {
  // common.cxx:680-688
  double *ptrs[16];
  int64_t s[16];
  for (int i = 0; i < N; i++) {
    ptrs[i] = W[i].data;
    s[i] = W[i].nrow;
  }
  PPXCK(dw, ppx_normalize(dw.ctx, ptrs, s, N, (int)W[0].ncol, nullptr));
}

void solve_update_fused(Matrix<> &M, Matrix<> &S, Matrix<> &W, Matrix<> *W_init, double ratio_step, Matrix<> *grad,
                        Matrix<> *dW, int mode, World &dw) {
  PPXCK(dw, ppx_solve_update(dw.ctx, M.data, S.data, W.data, W.nrow, (int)W.ncol, W_init ? W_init->data : nullptr,
                             ratio_step, mode, grad ? grad->data : nullptr, dW ? dW->data : nullptr, nullptr));
}

void SVD_solve(Matrix<> &M, Matrix<> &W, Matrix<> &S) {
  // common.cxx:710-725.  The world's `solver` picks the R x R factorisation (SVD pseudo-inverse semantics or the
  // Cholesky the north star asks for); both give W = M S^-1.
  World &dw = *M.wrld;
  solve_update_fused(M, S, W, nullptr, 1.0, nullptr, nullptr, dw.solver, dw);
}

void cholesky_solve(Matrix<> &M, Matrix<> &W, Matrix<> &S) {
  // common.cxx:727-737
  World &dw = *M.wrld;
  solve_update_fused(M, S, W, nullptr, 1.0, nullptr, nullptr, PPX_SOLVE_CHOL, dw);
}

void SVD_solve_mod(Matrix<> &M, Matrix<> &W, Matrix<> &W_init, Matrix<> &dW, Matrix<> &S, double ratio_step) {
  // common.cxx:739-758
  World &dw = *M.wrld;
  solve_update_fused(M, S, W, &W_init, ratio_step, nullptr, &dW, dw.solver, dw);
}

void gradsubprob(Matrix<> &M, Matrix<> &S, Matrix<> &W, Matrix<> &grad_W) {
  // common.cxx:1002-1004: grad = -M + W S (W unchanged): run the fused kernel on a scratch copy of W
  World &dw = *M.wrld;
  Matrix<> scratch(W);
  solve_update_fused(M, S, scratch, nullptr, 1.0, &grad_W, nullptr, PPX_SOLVE_CHOL, dw);
}

void KhatriRao_contract(Matrix<> &M, Tensor<> &V, Matrix<> *W, int *index, int *lens_H, World &dw) {
  // common.cxx:931-997: contract V with W[index[0]] (GEMM, :963) then W[index[1..N-2]] Hadamard-batched (:992).
  // The reference permutes the remaining modes into index[] order; the result M (s_{index[N-1]} x R) is the same.
  (void)lens_H;
  const int N = V.order;
  string modes;
  for (int i = 0; i < N; i++) modes.push_back((char)('a' + i));
  char x = (char)('a' + index[0]);
  Tensor<> cur = contract_mode(V, modes, false, x, W[index[0]], dw);
  modes.erase(modes.find(x), 1);
  for (int j = 1; j < N - 1; j++) {
    x = (char)('a' + index[j]);
    Tensor<> nxt = contract_mode(cur, modes, true, x, W[index[j]], dw);
    modes.erase(modes.find(x), 1);
    cur = std::move(nxt);
  }
  PPXCK(dw, ppx_memcpy_d2d(dw.ctx, M.data, cur.data, sizeof(double) * M.size));
}

void GramCache::init(Matrix<> *W, int N_, World &dw) {
  N = N_;
  R = (int)W[0].ncol;
  G.clear();
  for (int i = 0; i < N; i++) G.emplace_back(R, R, dw);
  for (int i = 0; i < N; i++) refresh(W, i, dw);
}

void GramCache::refresh(Matrix<> *W, int i, World &dw) {
  PPXCK(dw, ppx_gram(dw.ctx, W[i].data, W[i].nrow, W[i].nrow, R, G[i].data));
  if (dw.np > 1 && i == dw.shard_mode) dw.allreduce(G[i].data, (int64_t)R * R);  // rows of W_i are sharded
}

void GramCache::hadamard(int skip, double lambda, Matrix<> &S, World &dw) {
  const double *ptrs[16];
  for (int i = 0; i < N; i++) ptrs[i] = G[i].data;
  PPXCK(dw, ppx_hadamard_grams(dw.ctx, ptrs, N, skip, R, lambda, S.data));
}

void GramCache::solve(int skip, double lambda, Matrix<> &M, Matrix<> &W, Matrix<> *W_init, double ratio_step,
                      Matrix<> *grad, Matrix<> *dW, int mode, World &dw) {
  const double *ptrs[16];
  for (int i = 0; i < N; i++) ptrs[i] = G[i].data;
  PPXCK(dw, ppx_solve_update_g(dw.ctx, M.data, ptrs, N, skip, lambda, W.data, W.nrow, R,
                               W_init ? W_init->data : nullptr, ratio_step, mode, grad ? grad->data : nullptr,
                               dW ? dW->data : nullptr, nullptr));
}

void gradient_CP(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, World &dw) {
  // common.cxx:1009-1052: grad_W[i] = -MTTKRP_i + W[i] S_i for every mode, from scratch
  const int N = V.order;
  GramCache gc;
  gc.init(W, N, dw);
  Matrix<> S((int64_t)W[0].ncol, (int64_t)W[0].ncol, dw);
  for (int i = 0; i < N; i++) {
    int index[16], lens_H[16];
    int n = 0;
    for (int j = 0; j < N; j++)
      if (j != i) index[n++] = j;
    index[N - 1] = i;
    Matrix<> M(W[i].nrow, W[i].ncol, dw);
    KhatriRao_contract(M, V, W, index, lens_H, dw);
    // a replicated mode's MTTKRP is a partial sum over the local slab of the sharded mode (as in alsCP's own sweep)
    if (dw.np > 1 && i != dw.shard_mode) dw.allreduce(M.data, M.size);
    gc.hadamard(i, 0.0, S, dw);
    gradsubprob(M, S, W[i], grad_W[i]);
  }
}
