/* ref_driver.cxx -- TEST INFRASTRUCTURE (our code, not the reference's).
 *
 * A harness around the reference's own functions, linked against the objects that oracle/Makefile compiles from
 * /root/reference/{common,als_CP,als_Tucker}.cxx (unmodified) and oracle/ctf_standin/ctf.hpp.  It exists because the
 * reference's mains draw their inputs from fill_random; here the tensor and the initial factors come from raw-double
 * files (global first-index-fastest order, the format read_dense_from_file consumes, test_ALS.cxx:302), so the same
 * bytes can be handed to oracle/pp_oracle.py and to the CUDA path.
 *
 *   ref_driver -op <name> -lens a,b,c,.. -rank R [-ranks r0,r1,..] -V v.bin -W w.bin [-grad g.bin]
 *              [-tol t] [-tol_init t] [-maxiter n] [-lambda l] [-ratio_step x] [-update_pct p] [-resprint n]
 *              [-csv file] -out prefix
 *   op: alsCP | alsCP_DT | alsCP_PP | alsCP_PP_partupdate                     (als_CP.h:17-122)
 *       hosvd | alsTucker | alsTucker_DT | alsTucker_PP                       (als_Tucker.h:8-91; Tucker ops run hosvd first
 *                                                                             unless -W is given)
 *       kernels   : one call each of mttkrp_map_DT for every leaf, Build_mttkrp_map for every pair and single,
 *                   unroll_tensor_contraction, TTMc, build_V, Normalize, SVD_solve, cholesky_solve -> <out>.<name>.bin
 *   W / grad files hold all factors back to back (mode 0 first), each s_i x R column-major.
 *   Output: <out>.W.bin, <out>.grad.bin (CP), <out>.core.bin (Tucker), the CSV the reference writes (17 digits) and the
 *   reference's own stdout (the "DT starts from" / "pairwise perturbation starts from" markers).
 */
#include "als_CP.h"
#include "als_Tucker.h"
#include "common.h"

static vector<double> read_file(const string &path, size_t n) {
  vector<double> buf(n);
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "ref_driver: cannot open %s\n", path.c_str());
    exit(2);
  }
  size_t got = fread(buf.data(), sizeof(double), n, f);
  fclose(f);
  if (got != n) {
    fprintf(stderr, "ref_driver: %s holds %zu doubles, %zu wanted\n", path.c_str(), got, n);
    exit(2);
  }
  return buf;
}
static void write_file(const string &path, const vector<double> &buf) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) {
    fprintf(stderr, "ref_driver: cannot write %s\n", path.c_str());
    exit(2);
  }
  fwrite(buf.data(), sizeof(double), buf.size(), f);
  fclose(f);
}
static void dump(const string &path, TensorBase &t) { write_file(path, t.data); }
static void dump_all(const string &path, Matrix<> *W, int n) {
  vector<double> all;
  for (int i = 0; i < n; i++) all.insert(all.end(), W[i].data.begin(), W[i].data.end());
  write_file(path, all);
}
static vector<int> parse_ints(const char *s) {
  vector<int> v;
  string str(s), item;
  stringstream ss(str);
  while (getline(ss, item, ',')) v.push_back(atoi(item.c_str()));
  return v;
}

int main(int argc, char **argv) {
  map<string, string> a;
  for (int i = 1; i + 1 < argc; i += 2) a[argv[i]] = argv[i + 1];
  auto get = [&](const char *k, const char *d) { return a.count(k) ? a[k] : string(d); };
  const string op = get("-op", "alsCP_DT"), out = get("-out", "ref_out");
  vector<int> lens = parse_ints(get("-lens", "8,8,8,8").c_str());
  const int N = (int)lens.size(), R = atoi(get("-rank", "3").c_str());
  vector<int> ranks = a.count("-ranks") ? parse_ints(a["-ranks"].c_str()) : vector<int>(N, R);
  const double tol = atof(get("-tol", "1e-10").c_str()), tol_init = atof(get("-tol_init", "0.01").c_str());
  const double lambda = atof(get("-lambda", "0").c_str()), ratio_step = atof(get("-ratio_step", "1").c_str());
  const double update_pct = atof(get("-update_pct", "1").c_str());
  const int maxiter = atoi(get("-maxiter", "20").c_str()), resprint = atoi(get("-resprint", "10").c_str());
  const bool bench = atoi(get("-bench", "0").c_str()) != 0;
  const double timelimit = 1e9;

  World dw(argc, argv);
  Tensor<> V(N, lens.data(), dw);
  size_t P = 1;
  for (int i = 0; i < N; i++) P *= (size_t)lens[i];
  V.data = read_file(get("-V", "V.bin"), P);

  const bool tucker = op == "hosvd" || op.rfind("alsTucker", 0) == 0;
  Matrix<> *W = new Matrix<>[N], *grad_W = new Matrix<>[N], *F = new Matrix<>[N];
  if (a.count("-W")) {
    size_t tot = 0;
    for (int i = 0; i < N; i++) tot += (size_t)lens[i] * (tucker ? ranks[i] : R);
    vector<double> all = read_file(a["-W"], tot);
    size_t off = 0;
    for (int i = 0; i < N; i++) {
      const int r = tucker ? ranks[i] : R;
      W[i] = Matrix<>(lens[i], r, dw);
      copy(all.begin() + off, all.begin() + off + (size_t)lens[i] * r, W[i].data.begin());
      off += (size_t)lens[i] * r;
    }
  } else if (!tucker) {
    fprintf(stderr, "ref_driver: -W is required for CP ops\n");
    return 2;
  }
  if (!tucker) {
    size_t tot = 0;
    for (int i = 0; i < N; i++) tot += (size_t)lens[i] * R;
    vector<double> g = a.count("-grad") ? read_file(a["-grad"], tot) : vector<double>(tot, 1.0);
    size_t off = 0;
    for (int i = 0; i < N; i++) {
      grad_W[i] = Matrix<>(lens[i], R, dw);
      F[i] = Matrix<>(lens[i], R, dw);
      copy(g.begin() + off, g.begin() + off + (size_t)lens[i] * R, grad_W[i].data.begin());
      off += (size_t)lens[i] * R;
    }
  }

  ofstream csv(get("-csv", (out + ".csv").c_str()));
  csv.precision(17);
  cout.precision(17);

  if (op == "kernels") {
    /* the tree MTTKRP for every leaf, as alsCP_DT forms it (als_CP.cxx:236-284 does the last edge itself, so only
     * the first-level and inner nodes come from mttkrp_map_DT) */
    if (N >= 4) {
      map<string, string> parent, sibling;
      Construct_Dimension_Tree(parent, sibling, 0, N - 1);
      map<string, Tensor<>> mttkrp_map;
      for (auto &kv : parent) {
        if ((int)kv.first.size() == N || kv.first.size() < 2) continue;
        mttkrp_map_DT(mttkrp_map, parent, sibling, V, W, kv.first, dw);
      }
      for (auto &kv : mttkrp_map) dump(out + ".tree_" + kv.first + ".bin", kv.second);
      /* PP operators: every pair, then every single (als_CP.cxx:678-694) */
      map<string, Tensor<>> pp_map;
      char seq[32];
      for (int i = 0; i < N; i++)
        for (int j = i + 1; j < N; j++) {
          int k = 0;
          for (int m = 0; m < N; m++)
            if (m != i && m != j) seq[k++] = 'a' + m;
          seq[k] = '\0';
          Build_mttkrp_map(pp_map, V, W, seq, dw);
        }
      for (int i = 0; i < N; i++) {
        int k = 0;
        for (int m = 0; m < N; m++)
          if (m != i) seq[k++] = 'a' + m;
        seq[k] = '\0';
        Build_mttkrp_map(pp_map, V, W, seq, dw);
      }
      for (auto &kv : pp_map) dump(out + ".pp_" + kv.first + ".bin", kv.second);
    }
    for (int i = 0; i < N; i++) {
      Matrix<> G = unroll_tensor_contraction(V, i);
      dump(out + ".gram_" + to_string(i) + ".bin", G);
      Tensor<> Y;
      TTMc(Y, V, W, i, dw);
      dump(out + ".ttmc_" + to_string(i) + ".bin", Y);
      /* KhatriRao_contract as its callers set it up (als_CP.cxx:66-85, cp_simple_optimizer.cxx:30-45): mode i swapped
       * with the last one, lens_H[j] = extent of index[j] -- which only fits the intermediate it sizes (modes
       * index[1..]) when all extents are equal, so ragged shapes are skipped */
      bool uniform = true;
      for (int m = 1; m < N; m++) uniform = uniform && lens[m] == lens[0];
      if (uniform) {
        int index[32], lens_H[32];
        for (int m = 0; m < N; m++) index[m] = m;
        swap(index[i], index[N - 1]);
        for (int m = 0; m < N - 1; m++) lens_H[m] = lens[index[m]];
        lens_H[N - 1] = R;
        Matrix<> M(lens[i], R, dw);
        KhatriRao_contract(M, V, W, index, lens_H, dw);
        dump(out + ".krc_" + to_string(i) + ".bin", M);
      }
    }
    Tensor<> Vb;
    build_V(Vb, W, N, dw);
    dump(out + ".build_V.bin", Vb);
    /* normal equations of mode 0: S, then both solves */
    Matrix<> S(R, R, dw);
    S["ij"] = W[1]["ki"] * W[1]["kj"];
    for (int m = 2; m < N; m++) S["ij"] = S["ij"] * (W[m]["ki"] * W[m]["kj"]);
    Matrix<> X1(lens[0], R, dw), X2(lens[0], R, dw);
    SVD_solve(W[0], X1, S);
    cholesky_solve(W[0], X2, S);
    dump(out + ".S.bin", S);
    dump(out + ".svd_solve.bin", X1);
    dump(out + ".cholesky_solve.bin", X2);
    Normalize(W, N, dw);
    dump_all(out + ".normalized.bin", W, N);
    return 0;
  }

  bool ret = false;
  if (!tucker) {
    if (op == "alsCP")
      ret = alsCP(V, W, grad_W, F, tol, timelimit, maxiter, dw);
    else if (op == "alsCP_DT")
      ret = alsCP_DT(V, W, grad_W, F, tol, timelimit, maxiter, lambda, csv, resprint, bench, dw);
    else if (op == "alsCP_PP")
      ret = alsCP_PP(V, W, grad_W, F, tol, tol_init, timelimit, maxiter, lambda, ratio_step, csv, resprint, bench, dw);
    else if (op == "alsCP_PP_partupdate")
      ret = alsCP_PP_partupdate(V, W, grad_W, F, tol, tol_init, timelimit, maxiter, lambda, ratio_step, update_pct, csv,
                                resprint, bench, dw);
    else {
      fprintf(stderr, "ref_driver: unknown op %s\n", op.c_str());
      return 2;
    }
    dump_all(out + ".W.bin", W, N);
    dump_all(out + ".grad.bin", grad_W, N);
  } else {
    Tensor<> core;
    if (!a.count("-W")) {
      hosvd(V, core, W, ranks.data(), dw);
    } else {
      TTMc(core, V, W, -1, dw);
    }
    if (op == "hosvd") {
    } else if (op == "alsTucker")
      ret = alsTucker(V, core, W, tol, timelimit, maxiter, dw);
    else if (op == "alsTucker_DT")
      ret = alsTucker_DT(V, core, W, tol, timelimit, maxiter, csv, resprint, bench, dw);
    else if (op == "alsTucker_PP")
      ret = alsTucker_PP(V, core, W, tol, tol_init, timelimit, maxiter, csv, resprint, bench, dw);
    else {
      fprintf(stderr, "ref_driver: unknown op %s\n", op.c_str());
      return 2;
    }
    dump_all(out + ".W.bin", W, N);
    dump(out + ".core.bin", core);
  }
  csv.close();
  printf("ref_driver return %d\n", (int)ret);
  return 0;
}
