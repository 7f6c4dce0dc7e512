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

inline World *make_world(const CliOptions &o) {
  int device = o.device;
  if (device < 0) {
    const char *lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) : 0;
  }
  World *dw = new World(device, (size_t)2 << 30);
  dw->solver = (o.solver == "svd") ? PPX_SOLVE_SVD_PINV : PPX_SOLVE_CHOL;
  dw->use_graph = o.graph != 0;
  dw->fast_residual = o.fastres != 0;
  dw->seed = o.seed;
  return dw;
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
  const int dim = (int)lens.size();
  const char t0 = o.tensor[0];
  const bool second = o.tensor.size() > 1 && o.tensor[1] == '2';
  const bool first = o.tensor.size() > 1 && o.tensor[1] == '1';
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
      V = std::move(V0);
    } else {
      std::vector<int64_t> l2(o.dim / 2, (int64_t)o.s * o.s);
      V = Tensor<>(o.dim / 2, l2.data(), dw, false);
      fold_unfold(V0, V);
    }
    return true;
  }
  if (t0 == 'c') {
    // c: rank-R tensor with constrained collinearity plus uniform noise scaled to ratio_noise * ||V|| (test_ALS.cxx:245-261)
    std::vector<int> li(lens.begin(), lens.end());
    V = Gen_collinearity(li.data(), dim, o.R, o.col_min, o.col_max, dw);
    Tensor<> V_noise(dim, lens.data(), dw, false);
    V_noise.fill_random(-1, 1, o.seed, 101);
    const double noise_norm = V_noise.norm2(), V_norm = V.norm2();
    PPXCK(dw, ppx_axpby(dw.ctx, o.ratio_noise * V_norm / noise_norm, V_noise.data, 1.0, V.data, V.size));
    return true;
  }
  if (t0 == 'r' && second) {
    // r2: random tensor, uniform in [0.5,1) (test_ALS.cxx:266-273); pp_bench uses [-1,1) (pp_bench.cxx:249)
    V = Tensor<>(dim, lens.data(), dw);
    if (bench_ranges) V.fill_random(-1, 1, o.seed, 100);
    else V.fill_random(0.5, 1, o.seed, 100);
    return true;
  }
  if (t0 == 'r') {
    // r: tensor made by random matrices (test_ALS.cxx:274-286)
    std::vector<Matrix<>> Wt;
    for (int i = 0; i < dim; i++) {
      Wt.emplace_back(lens[i], o.R, dw);
      Wt[i].fill_random(0, 1, o.seed, (uint64_t)i);
    }
    build_V(V, Wt.data(), dim, dw);
    return true;
  }
  if (t0 == 'o') {
    // o1: coil-100 (3 x 128 x 128 x 7200), o2: time-lapse (33 x 1344 x 1024 x 9); raw doubles, global order
    int64_t l1[4] = {3, 128, 128, 7200}, l2[4] = {33, 1344, 1024, 9};
    if (!first && !second) return false;
    V = Tensor<>(4, first ? l1 : l2, dw);
    std::string path = (o.tensorfile != "test") ? o.tensorfile : (first ? "coil-100.bin" : "time-lapse.bin");
    if (dw.rank == 0) cout << "Read the tensor from file " << path << " ...... " << endl;
    V.read_dense_from_file(path.c_str());
    if (dw.rank == 0) cout << "Read dataset finished " << endl;
    return true;
  }
  if (dw.rank == 0)
    fprintf(stderr, "tensor '%s' is not known (p, p2, c, r, r2, o1, o2)\n",
            o.tensor.c_str());
  return false;
}

#endif
