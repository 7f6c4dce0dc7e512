// run -- driver for the class-based path: CPD<double, Optimizer>::als with -pp selecting the optimizer
// (reference: run.cxx:387-414).  -pp 0: CPDTOptimizer, 1: CPMSDTOptimizer, 2: CPDTLROptimizer, 3: CPMSDTLROptimizer
// (-updaterank, -randomsvd), 4: CPSimpleOptimizer.
#include "cli_common.h"
#include "src/CP.h"
#include "src/optimizer/cp_dt_optimizer.h"
#include "src/optimizer/cp_msdt_optimizer.h"
#include "src/optimizer/cp_dt_lr_optimizer.h"
#include "src/optimizer/cp_msdt_lr_optimizer.h"
#include "src/optimizer/cp_simple_optimizer.h"

template <class Opt>
static void run_with(const CliOptions &o, Tensor<> *V, Matrix<> *W, double Vnorm, ofstream &Plot_File, World &dw) {
  CPD<double, Opt> decom(V->order, (int)V->lens[0], o.R, dw);
  decom.Init(V, W, o.lambda_);
  decom.als(o.tol * Vnorm, o.timelimit, o.maxiter, o.resprint, Plot_File);
}

template <class Opt>
static void run_with_lr(const CliOptions &o, Tensor<> *V, Matrix<> *W, double Vnorm, ofstream &Plot_File, World &dw) {
  CPD<double, Opt> decom(V->order, (int)V->lens[0], o.R, o.updaterank, o.randomsvd, dw);  // run.cxx:400-409
  decom.Init(V, W, o.lambda_);
  decom.als(o.tol * Vnorm, o.timelimit, o.maxiter, o.resprint, Plot_File);
}

int main(int argc, char **argv) {
  CliOptions o = parse_cli(argc, argv, 4);
  const double start_time = wall_time();
  World *dwp;
  try {
    dwp = make_world(o, argc, argv);
  } catch (const std::exception &e) {
    fprintf(stderr, "run: %s\n", e.what());
    return 2;
  }
  World &dw = *dwp;
  int rc = 0;
  try {
    if (dw.np > 1) throw std::runtime_error("run (the class-based optimizers of src/optimizer) is single-GPU; use test_ALS / pp_bench under a multi-process launcher");
    print_options(o, dw);
    Tensor<> *V = new Tensor<>();  // adopted by the decomposition and never freed, as in the reference
    if (!build_input_tensor(o, *V, dw, false)) {
      delete dwp;
      return 3;
    }
    const double Vnorm = V->norm2();
    if (dw.rank == 0) cout << "Vnorm= " << Vnorm << endl;
    ofstream Plot_File(o.filename);
    const int N = V->order;
    Matrix<> *W = new Matrix<>[N];  // adopted (and deleted) by the decomposition
    for (int i = 0; i < N; i++) {
      W[i] = Matrix<>(V->lens[i], o.R, dw);
      W[i].fill_random(0, 1, o.seed + 1, (uint64_t)i);
    }
    if (o.pp == 0) run_with<CPDTOptimizer<double>>(o, V, W, Vnorm, Plot_File, dw);
    else if (o.pp == 1) run_with<CPMSDTOptimizer<double>>(o, V, W, Vnorm, Plot_File, dw);
    else if (o.pp == 2) run_with_lr<CPDTLROptimizer<double>>(o, V, W, Vnorm, Plot_File, dw);
    else if (o.pp == 3) run_with_lr<CPMSDTLROptimizer<double>>(o, V, W, Vnorm, Plot_File, dw);
    else run_with<CPSimpleOptimizer<double>>(o, V, W, Vnorm, Plot_File, dw);
    if (dw.rank == 0) printf("experiment took %lf seconds\n", wall_time() - start_time);
    delete V;
  } catch (const std::exception &e) {
    fprintf(stderr, "run: %s\n", e.what());
    rc = 1;
  }
  delete dwp;
  return rc;
}
