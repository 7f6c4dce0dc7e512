// test_ALS -- main command line: parse flags, build / load the tensor, random initial factors, dispatch to the CP
// or Tucker driver.  Same flags and defaults as the reference's test_ALS.cxx:22-415.
//   test_ALS -model CP -tensor r -dim 4 -size 300 -rank 50 -pp 1 -maxiter 50
#include "cli_common.h"

int main(int argc, char **argv) {
  CliOptions o = parse_cli(argc, argv, 2);
  const double start_time = wall_time();
  World *dwp;
  try {
    dwp = make_world(o, argc, argv);
  } catch (const std::exception &e) {
    fprintf(stderr, "test_ALS: %s\n", e.what());
    return 2;
  }
  World &dw = *dwp;
  int rc = 0;
  try {
    print_options(o, dw);
    Tensor<> V;
    if (!build_input_tensor(o, V, dw, false)) {
      delete dwp;
      return 3;
    }
    const double Vnorm = V.norm2_sharded();
    if (dw.rank == 0) cout << "Vnorm= " << Vnorm << endl;
    ofstream Plot_File;
    if (dw.rank == 0) Plot_File.open(o.filename);  // only rank 0 writes (als_CP.cxx:132,192)
    const int N = V.order;
    Matrix<> *W = new Matrix<>[N];
    Matrix<> *grad_W = new Matrix<>[N];
    Matrix<> *F = new Matrix<>[N];
    for (int i = 0; i < N; i++) {  // test_ALS.cxx:332-345
      W[i] = seeded_factor(V.lens[i], o.R, i, o.seed + 1, dw);
      grad_W[i] = seeded_factor(V.lens[i], o.R, i, o.seed + 2, dw);
      F[i] = Matrix<>(V.lens[i], o.R, dw);
    }
    if (o.model[0] == 'C') {
      if (o.pp == 0)
        alsCP_DT(V, W, grad_W, F, o.tol * Vnorm, o.timelimit, o.maxiter, o.lambda_, Plot_File, o.resprint, false, dw);
      else if (o.pp == 1)
        alsCP_PP(V, W, grad_W, F, o.tol * Vnorm, o.pp_res_tol, o.timelimit, o.maxiter, o.lambda_, o.magni, Plot_File,
                 o.resprint, false, dw);
      else
        alsCP_PP_partupdate(V, W, grad_W, F, o.tol * Vnorm, o.pp_res_tol, o.timelimit, o.maxiter, o.lambda_, o.magni,
                            o.update_percentage_pp, Plot_File, o.resprint, false, dw);
    } else {
      int ranks[16];
      for (int i = 0; i < N; i++) ranks[i] = o.R;
      if (o.tensor == "o1") {  // test_ALS.cxx:369-373
        ranks[0] = 3, ranks[1] = 10, ranks[2] = 10, ranks[3] = 70;
      } else if (o.tensor == "o2") {  // :375-379
        ranks[0] = 10, ranks[1] = 100, ranks[2] = 100, ranks[3] = 5;
      }
      Tensor<> hosvd_core;
      hosvd(V, hosvd_core, W, ranks, dw);
      if (o.pp == 0)
        alsTucker_DT(V, hosvd_core, W, o.tol * Vnorm, o.timelimit, o.maxiter, Plot_File, o.resprint, false, dw);
      else
        alsTucker_PP(V, hosvd_core, W, o.tol * Vnorm, o.pp_res_tol, o.timelimit, o.maxiter, Plot_File, o.resprint,
                     false, dw);
    }
    if (dw.rank == 0) printf("experiment took %lf seconds\n", wall_time() - start_time);
    delete[] W;
    delete[] grad_W;
    delete[] F;
  } catch (const std::exception &e) {
    fprintf(stderr, "test_ALS: %s\n", e.what());
    rc = 1;
  }
  delete dwp;
  return rc;
}
