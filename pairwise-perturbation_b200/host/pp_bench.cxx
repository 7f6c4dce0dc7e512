// pp_bench -- per-step timing harness with the reference's protocol (pp_bench.cxx:295-348): `maxiter` repetitions,
// from the same starting factors, of ONE dimension-tree sweep ([dimension tree step time]) and of ONE PP operator
// build + approximate sweep ([PP first time]) followed by one approximate sweep alone ([PP second time]).
#include "cli_common.h"

int main(int argc, char **argv) {
  CliOptions o = parse_cli(argc, argv, 2);
  if (!getCmdOption(argv, argv + argc, "-resprint")) o.resprint = 1;
  const double start_time = wall_time();
  World *dwp;
  try {
    dwp = make_world(o, argc, argv);
  } catch (const std::exception &e) {
    fprintf(stderr, "pp_bench: %s\n", e.what());
    return 2;
  }
  World &dw = *dwp;
  int rc = 0;
  try {
    print_options(o, dw);
    Tensor<> V;
    if (!build_input_tensor(o, V, dw, true)) {
      delete dwp;
      return 3;
    }
    const double Vnorm = V.norm2_sharded();
    ofstream Plot_File;
    if (dw.rank == 0) Plot_File.open(o.filename);
    const int N = V.order;
    Matrix<> *W = new Matrix<>[N], *W_DT = new Matrix<>[N], *W_PP = new Matrix<>[N], *grad_W = new Matrix<>[N];
    Matrix<> *F = new Matrix<>[N];
    for (int i = 0; i < N; i++) {  // pp_bench.cxx:270-287
      if (o.model[0] == 'C') {
        W[i] = seeded_factor(V.lens[i], o.R, i, o.seed + 1, dw);
        grad_W[i] = seeded_factor(V.lens[i], o.R, i, o.seed + 2, dw);
      } else {  // Tucker factors are replicated in full on every rank (multi-GPU)
        const int64_t rows = (dw.np > 1 && i == dw.shard_mode) ? dw.shard_global : V.lens[i];
        W[i] = Matrix<>(rows, o.R, dw, false);
        W[i].fill_random(0, 1, o.seed + 1, (uint64_t)i);
        grad_W[i] = Matrix<>(rows, o.R, dw, false);
        grad_W[i].fill_random(0, 1, o.seed + 2, (uint64_t)i);
      }
      W_DT[i] = W[i];
      W_PP[i] = W[i];
      F[i] = Matrix<>(W[i].nrow, o.R, dw);
    }
    if (dw.rank == 0) Plot_File << "[timetype],[dtime]" << "\n";
    if (o.model[0] == 'C') {
      for (int i = 0; i < o.maxiter; i++) {  // :299-305
        alsCP_DT(V, W_DT, grad_W, F, o.tol * Vnorm, o.timelimit, 1, o.lambda_, Plot_File, o.resprint, true, dw);
        for (int j = 0; j < N; j++) W_DT[j] = W[j];
      }
      if (dw.rank == 0) Plot_File << endl;
      for (int i = 0; i < o.maxiter; i++) {  // :308-314
        alsCP_PP(V, W_PP, grad_W, F, o.tol * Vnorm, o.pp_res_tol, o.timelimit, 1, o.lambda_, o.magni, Plot_File,
                 o.resprint, true, dw);
        for (int j = 0; j < N; j++) W_PP[j] = W[j];
      }
      if (dw.rank == 0) Plot_File << endl;
    } else {
      int ranks[16];
      for (int i = 0; i < N; i++) ranks[i] = o.R;
      // the reference benchmarks from the RANDOM factors with a zero core (hosvd is commented out, pp_bench.cxx:326)
      Tensor<> hosvd_core(N, ranks, dw);
      for (int i = 0; i < o.maxiter; i++) {  // :328-335
        for (int j = 0; j < N; j++) W_DT[j] = W[j];
        alsTucker_DT(V, hosvd_core, W_DT, o.tol * Vnorm, o.timelimit, 1, Plot_File, o.resprint, true, dw);
      }
      if (dw.rank == 0) Plot_File << endl;
      for (int i = 0; i < o.maxiter; i++) {  // :338-345
        for (int j = 0; j < N; j++) W_PP[j] = W[j];
        alsTucker_PP(V, hosvd_core, W_PP, o.tol * Vnorm, o.pp_res_tol, o.timelimit, 1, Plot_File, o.resprint, true,
                     dw);
      }
      if (dw.rank == 0) Plot_File << endl;
    }
    if (dw.rank == 0) printf("experiment took %lf seconds\n", wall_time() - start_time);
    delete[] F;
    delete[] W;
    delete[] W_DT;
    delete[] W_PP;
    delete[] grad_W;
  } catch (const std::exception &e) {
    fprintf(stderr, "pp_bench: %s\n", e.what());
    rc = 1;
  }
  delete dwp;
  return rc;
}
