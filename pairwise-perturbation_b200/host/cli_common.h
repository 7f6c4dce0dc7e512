// cli_common.h -- flag parsing and tensor construction shared by test_ALS / pp_bench / run.
// Flag names and defaults are the reference's (test_ALS.cxx:64-196); new, GPU-side flags are listed at the end.
#ifndef PPX_HOST_CLI_COMMON_H__
#define PPX_HOST_CLI_COMMON_H__

#include <algorithm>
#include <cstring>
#include <string>
#include "als_CP.h"
#include "als_Tucker.h"

inline char *getCmdOption(char **begin, char **end, const std::string &option) {
  char **itr = std::find(begin, end, option);
  if (itr != end && ++itr != end) return *itr;
  return 0;
}

struct CliOptions {
  std::string model = "CP";     // CP | Tucker
  std::string tensor = "p";     // p / p2 / c / r / r2 / o1 / o2
  int pp = 0;                   // 0 dimension tree, 1 pairwise perturbation, 2 PP with partial update
  double update_percentage_pp = 1.0;
  int dim = 8, s = 10, R = 5;
  int issparse = 0;
  double tol = 1e-10, pp_res_tol = 1e-2, lambda_ = 0., magni = 1.;
  std::string filename = "out.csv";
  std::string tensorfile = "test";
  double col_min = 0.5, col_max = 0.9, ratio_noise = 0.01;
  double timelimit = 5e3;
  int maxiter = 5000;
  int resprint = 10;
  // additions of this build
  int device = -1;              // -device   (default: LOCAL_RANK or 0)
  std::string solver = "chol";  // -solver chol|svd : R x R solve behind SVD_solve
  uint64_t seed = 1;            // -seed     : counter-based generator seed for the tensor; factors use seed+1
  int graph = 1;                // -graph 0|1: replay the PP approximate sweep as a CUDA graph
  int fastres = 0;              // -fastres 0|1: alsCP_DT reports the residual from the MTTKRP identity (no pass over V)
  std::string lens;             // -lens a,b,c,.. : non-cubic synthetic tensor (overrides -dim/-size)
  int updaterank = 1, randomsvd = 0;  // run.cxx -pp 2 / 3: rank of the low-rank update, randomized range finder
};

inline CliOptions parse_cli(int argc, char **argv, int pp_max) {
  CliOptions o;
  char **b = argv, **e = argv + argc;
  auto get = [&](const char *name) { return getCmdOption(b, e, name); };
  if (char *v = get("-model")) o.model = (v[0] == 'C' || v[0] == 'T') ? v : "CP";
  if (char *v = get("-tensor")) o.tensor = v;
  if (char *v = get("-pp")) {
    o.pp = atoi(v);
    if (o.pp < 0 || o.pp > pp_max) o.pp = 0;
  }
  if (char *v = get("-update_percentage_pp")) {
    o.update_percentage_pp = atof(v);
    if (o.update_percentage_pp < 0 || o.update_percentage_pp > 1) o.update_percentage_pp = 1.0;
  }
  if (char *v = get("-dim")) {
    o.dim = atoi(v);
    if (o.dim < 0) o.dim = 8;
  }
  if (char *v = get("-maxiter")) {
    o.maxiter = atoi(v);
    if (o.maxiter < 0) o.maxiter = 5000;
  }
  if (char *v = get("-timelimit")) {
    o.timelimit = atof(v);
    if (o.timelimit < 0) o.timelimit = 5e3;
  }
  if (char *v = get("-size")) {
    o.s = atoi(v);
    if (o.s < 0) o.s = 10;
  }
  o.R = o.s / 2;
  if (char *v = get("-rank")) {
    o.R = atoi(v);
    if (o.R < 0 || o.R > o.s) o.R = o.s / 2;
  }
  if (char *v = get("-issparse")) {
    o.issparse = atoi(v);
    if (o.issparse < 0 || o.issparse > 1) o.issparse = 0;
  }
  if (char *v = get("-resprint")) {
    o.resprint = atoi(v);
    if (o.resprint < 0) o.resprint = 10;
  }
  if (char *v = get("-tol")) {
    o.tol = atof(v);
    if (o.tol < 0 || o.tol > 1) o.tol = 1e-10;
  }
  if (char *v = get("-pp_res_tol")) {
    o.pp_res_tol = atof(v);
    if (o.pp_res_tol < 0 || o.pp_res_tol > 1) o.pp_res_tol = 1e-2;
  }
  if (char *v = get("-lambda")) {
    o.lambda_ = atof(v);
    if (o.lambda_ < 0) o.lambda_ = 0.;
  }
  if (char *v = get("-magni")) {
    o.magni = atof(v);
    if (o.magni < 0) o.magni = 1.;
  }
  if (char *v = get("-filename")) o.filename = v;
  if (char *v = get("-tensorfile")) o.tensorfile = v;
  if (char *v = get("-colmin")) o.col_min = atof(v);
  if (char *v = get("-colmax")) o.col_max = atof(v);
  if (char *v = get("-rationoise")) {
    o.ratio_noise = atof(v);
    if (o.ratio_noise < 0) o.ratio_noise = 0.01;
  }
  if (char *v = get("-updaterank")) o.updaterank = atoi(v);
  if (char *v = get("-randomsvd")) o.randomsvd = atoi(v);
  if (char *v = get("-device")) o.device = atoi(v);
  if (char *v = get("-solver")) o.solver = v;
  if (char *v = get("-seed")) o.seed = strtoull(v, nullptr, 10);
  if (char *v = get("-graph")) o.graph = atoi(v);
  if (char *v = get("-fastres")) o.fastres = atoi(v);
  if (char *v = get("-lens")) o.lens = v;
  return o;
}

inline void print_options(const CliOptions &o, World &dw) {
  if (dw.rank != 0) return;
  cout << "  model=  " << o.model << "  tensor=  " << o.tensor << "  pp=  " << o.pp << endl;
  cout << "  dim=  " << o.dim << "  size=  " << o.s << "  rank=  " << o.R << endl;
  cout << "  issparse=  " << o.issparse << "  tolerance=  " << o.tol << "  restarttol=  " << o.pp_res_tol << endl;
  cout << "  lambda=  " << o.lambda_ << "  magnitude=  " << o.magni << "  filename=  " << o.filename << endl;
  cout << "  col_min=  " << o.col_min << "  col_max=  " << o.col_max << "  rationoise  " << o.ratio_noise << endl;
  cout << "  timelimit=  " << o.timelimit << "  maxiter=  " << o.maxiter << "  resprint=  " << o.resprint << endl;
  cout << "  tensorfile=  " << o.tensorfile << "  update_percentage_pp=  " << o.update_percentage_pp << endl;
  cout << "  solver=  " << o.solver << "  seed=  " << o.seed << "  graph=  " << o.graph << endl;
}

// One process per GPU, like the reference's mains under mpirun (test_ALS.cxx:58-60,200): rank, size and device come from
// the launcher's environment and World(argc, argv) joins the NCCL communicator by itself.
inline World *make_world(const CliOptions &o, int argc, char **argv) {
  World *dw = new World(argc, argv, (size_t)2 << 30, o.device);
  dw->solver = (o.solver == "svd") ? PPX_SOLVE_SVD_PINV : PPX_SOLVE_CHOL;
  dw->use_graph = o.graph != 0;
  dw->fast_residual = o.fastres != 0;
  dw->seed = o.seed;
  return dw;
}

// rows [row_begin, row_end) of a full tensor whose leading mode is being sharded (generators that have no slab form)
inline Tensor<> shard_leading(Tensor<> &full, World &dw) {
  if (dw.np == 1) return std::move(full);
  std::vector<int64_t> l(full.lens, full.lens + full.order);
  const int64_t Lg = l[0], rows = dw.row_end - dw.row_begin;
  l[0] = rows;
  Tensor<> loc(full.order, l.data(), dw, false);
  PPXCK(dw, ppx_memcpy2d_d2d(dw.ctx, loc.data, sizeof(double) * rows, full.data + dw.row_begin, sizeof(double) * Lg,
                             sizeof(double) * rows, (size_t)(full.size / Lg)));
  return loc;
}

// factor / gradient matrix of mode i filled like W[i].fill_random(0,1) of the reference (test_ALS.cxx:337-338); for the
// sharded mode the local rows of the same global matrix
inline Matrix<> seeded_factor(int64_t rows_local, int R, int mode, uint64_t seed, World &dw) {
  Matrix<> m(rows_local, R, dw, false);
  if (dw.np > 1 && mode == dw.shard_mode) m.fill_random_rows(0, 1, seed, (uint64_t)mode, dw.shard_global, dw.row_begin);
  else m.fill_random(0, 1, seed, (uint64_t)mode);
  return m;
}

// Builds the input tensor the way test_ALS.cxx:222-326 does: p / p2 (Poisson operator), c (constrained collinearity +
// noise), r, r2, o1 / o2 (raw files).
inline bool build_input_tensor(const CliOptions &o, Tensor<> &V, World &dw, bool bench_ranges) {
  std::vector<int64_t> lens;
  if (!o.lens.empty()) {
    size_t p = 0;
    while (p < o.lens.size()) {
      size_t q = o.lens.find(',', p);
      if (q == std::string::npos) q = o.lens.size();
      lens.push_back(atoll(o.lens.substr(p, q - p).c_str()));
      p = q + 1;
    }
  } else {
    lens.assign(o.dim, o.s);
  }
  const char t0 = o.tensor[0];
  const bool second = o.tensor.size() > 1 && o.tensor[1] == '2';
  const bool first = o.tensor.size() > 1 && o.tensor[1] == '1';
  if (t0 == 'o' && o.lens.empty()) {
    // o1: coil-100 (3 x 128 x 128 x 7200), o2: time-lapse (33 x 1344 x 1024 x 9)  (test_ALS.cxx:294-297, :313-316)
    const int64_t l1[4] = {3, 128, 128, 7200}, l2[4] = {33, 1344, 1024, 9};
    lens.assign(second ? l2 : l1, (second ? l2 : l1) + 4);
  }
  const int dim = (int)lens.size();
  // several GPUs: the tensor is sharded along its leading mode; every generator below produces this rank's
  // rows of the same global tensor a single process would build
  auto shard = [&](int64_t global_leading) {
    if (dw.np == 1) return true;
    if (global_leading < dw.np) {
      if (dw.rank == 0) fprintf(stderr, "the leading mode (%lld) is shorter than the number of GPUs (%d)\n",
                                (long long)global_leading, dw.np);
      return false;
    }
    dw.set_shard(0, global_leading);
    return true;
  };
  if (t0 == 'p') {
    // p2: Poisson operator as an order-dim tensor; p: the same entries folded to dim/2 modes of size s*s
    // (test_ALS.cxx:222-244)
    if (!o.lens.empty() || o.dim % 2) {
      if (dw.rank == 0) fprintf(stderr, "tensor 'p'/'p2' needs an even -dim and a cubic -size\n");
      return false;
    }
    Tensor<> V0;
    laplacian_tensor(V0, o.dim, o.s, o.issparse != 0, dw);
    if (second) {
      if (!shard(o.s)) return false;
      V = shard_leading(V0, dw);
    } else {
      std::vector<int64_t> l2(o.dim / 2, (int64_t)o.s * o.s);
      Tensor<> Vf(o.dim / 2, l2.data(), dw, false);
      fold_unfold(V0, Vf);
      if (!shard((int64_t)o.s * o.s)) return false;
      V = shard_leading(Vf, dw);
    }
    return true;
  }
  if (!shard(lens[0])) return false;
  std::vector<int64_t> lloc(lens);
  if (dw.np > 1) lloc[0] = dw.row_end - dw.row_begin;
  if (t0 == 'c') {
    // c: rank-R tensor with constrained collinearity plus uniform noise scaled to ratio_noise * ||V|| (test_ALS.cxx:245-261)
    std::vector<int> li(lens.begin(), lens.end());
    V = Gen_collinearity(li.data(), dim, o.R, o.col_min, o.col_max, dw);
    Tensor<> V_noise(dim, lloc.data(), dw, false);
    V_noise.fill_random_rows(-1, 1, o.seed, 101, lens[0], dw.np > 1 ? dw.row_begin : 0);
    const double noise_norm = V_noise.norm2_sharded(), V_norm = V.norm2_sharded();
    PPXCK(dw, ppx_axpby(dw.ctx, o.ratio_noise * V_norm / noise_norm, V_noise.data, 1.0, V.data, V.size));
    return true;
  }
  if (t0 == 'r' && second) {
    // r2: random tensor, uniform in [0.5,1) (test_ALS.cxx:266-273); pp_bench uses [-1,1) (pp_bench.cxx:249)
    V = Tensor<>(dim, lloc.data(), dw, false);
    V.fill_random_rows(bench_ranges ? -1 : 0.5, 1, o.seed, 100, lens[0], dw.np > 1 ? dw.row_begin : 0);
    return true;
  }
  if (t0 == 'r') {
    // r: tensor made by random matrices (test_ALS.cxx:274-286)
    std::vector<Matrix<>> Wt;
    for (int i = 0; i < dim; i++) Wt.push_back(seeded_factor(lloc[i], o.R, i, o.seed, dw));
    build_V(V, Wt.data(), dim, dw);
    return true;
  }
  if (t0 == 'o') {
    // raw little-endian doubles in global order (test_ALS.cxx:287-326); -lens overrides the extents for a file of
    // another shape in the same format
    if (!first && !second) return false;
    V = Tensor<>(dim, lloc.data(), dw, false);
    std::string path = (o.tensorfile != "test") ? o.tensorfile : (first ? "coil-100.bin" : "time-lapse.bin");
    if (dw.rank == 0) cout << "Read the tensor from file " << path << " ...... " << endl;
    V.read_dense_rows_from_file(path.c_str(), lens[0], dw.np > 1 ? dw.row_begin : 0);
    if (dw.rank == 0) cout << "Read dataset finished " << endl;
    return true;
  }
  if (dw.rank == 0)
    fprintf(stderr, "tensor '%s' is not known (p, p2, c, r, r2, o1, o2)\n",
            o.tensor.c_str());
  return false;
}

#endif
