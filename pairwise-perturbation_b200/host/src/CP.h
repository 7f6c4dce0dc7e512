// CP.h -- CPD<dtype, Optimizer>: owns the optimizer and the gradient matrices, runs the outer ALS loop with the
// reference's logging (reference: src/CP.h:10-45, src/CP.cxx).
#ifndef PPX_HOST_CP_H__
#define PPX_HOST_CP_H__

#include <cmath>
#include "decomposition.h"

template <typename dtype, class Optimizer>
class CPD : public Decomposition<dtype> {
public:
  Matrix<> *grad_W = NULL;  // gradient in each dimension
  double gradnorm = 0.;
  Optimizer *optimizer = NULL;
  char seq_V[100];

  CPD(int order, int size, int r, World &dw) : Decomposition<dtype>(order, size, r, dw) {
    optimizer = new Optimizer(order, r, dw);
    make_seq();
  }
  // constructors for the low-rank-update optimizers (src/CP.cxx:21-46); those optimizers are out of scope here
  // (SURVEY.md 8f-4), any Optimizer with the matching constructor works
  CPD(int order, int size, int r, int update_rank, World &dw) : Decomposition<dtype>(order, size, r, dw) {
    optimizer = new Optimizer(order, r, update_rank, dw);
    make_seq();
  }
  CPD(int order, int size, int r, int update_rank, int randomsvd, World &dw)
      : Decomposition<dtype>(order, size, r, dw) {
    optimizer = new Optimizer(order, r, update_rank, randomsvd, dw);
    make_seq();
  }
  CPD(int order, int *size, int *r, World &dw) : Decomposition<dtype>(order, size, r, dw) {
    for (int i = 1; i < order; i++) assert(this->rank[i] == r[0]);
    optimizer = new Optimizer(order, r[0], dw);
    make_seq();
  }

  // src/CP.cxx:67-84.  grad_W is filled with uniform [0,1) numbers as the reference does (it only affects the
  // first printed gradient norm -- and keeps the iteration-0 test `gradnorm < tol` from firing).
  void Init(Tensor<dtype> *input, Matrix<dtype> *mat, double lambda = 0.) {
    Decomposition<dtype>::Init(input, mat);
    World *dw = this->world;
    if (grad_W != NULL) delete[] grad_W;
    grad_W = new Matrix<>[this->order];
    for (int i = 0; i < this->order; i++) {
      grad_W[i] = Matrix<dtype>(mat[i].nrow, this->rank[i], *dw);
      grad_W[i].fill_random(0, 1);
    }
    this->optimizer->configure(input, mat, grad_W, lambda);
  }

  ~CPD() {
    if (grad_W != NULL) delete[] grad_W;
    if (optimizer != NULL) delete optimizer;
  }

  void print_grad(int i) const {
    assert(grad_W != NULL);
    grad_W[i].print();
  }

  void update_gradnorm() { gradnorm = gradient_norm_sharded(); }  // src/CP.cxx:101-108

  // src/CP.cxx:110-187.  Returns true when it stopped before maxsweep+1.
  bool als(double tol, double timelimit, int maxsweep, int resprint, ofstream &Plot_File, bool bench = false) {
    cout.precision(13);
    World *dw = this->world;
    dw->sync();
    double st_time = wall_time();
    int iters = 0;
    double sweeps = 0;
    double diffnorm_V = 1000.;
    if (!bench && dw->rank == 0 && Plot_File.is_open())
      Plot_File << "[dim],[iter],[gradnorm],[tol],[pp_update],[diffV],[dtime]" << "\n";
    while (int(sweeps) <= maxsweep) {
      if (iters % resprint == 0 || sweeps >= maxsweep || sweeps == 0) {
        dw->sync();
        const double st_time1 = wall_time();
        update_gradnorm();
        diffnorm_V = (trace_sink() && trace_sink()->skip_residual) ? -1.0
                                                                   : cp_residual_norm(*this->V, this->W, this->order, *dw);
        dw->sync();
        st_time += wall_time() - st_time1;
        const double dtime = wall_time() - st_time;
        if (trace_sink()) trace_sink()->rows.push_back({sweeps, gradnorm, 0, diffnorm_V, dtime});
        if (!bench) {
          if (dw->rank == 0) {
            if (!trace_quiet())
              cout << "  [dim]=  " << this->size[0] << "  [sweeps]=  " << sweeps << "  [gradnorm]  " << gradnorm
                   << "  [tol]  " << tol << "  [pp_update]  " << 0 << "  [residual]  " << diffnorm_V << "  [dtime]  "
                   << dtime << "\n";
            if (Plot_File.is_open()) {
              Plot_File << this->size[0] << "," << sweeps << "," << gradnorm << "," << tol << "," << 0 << ","
                        << diffnorm_V << "," << dtime << "\n";
              if (iters % 100 == 0 && iters != 0) Plot_File << endl;
            }
          }
        } else if (iters != 0) {
          if (trace_sink()) trace_sink()->bench_times.push_back(dtime);
          if (dw->rank == 0) {
            if (!trace_quiet()) cout << "  [dimension tree step time]  " << dtime << "\n";
            if (Plot_File.is_open()) Plot_File << "[DTtime]" << "," << dtime << "\n";
          }
        }
        if (gradnorm < tol || wall_time() - st_time > timelimit) break;
      }
      sweeps += this->optimizer->step();
      iters += 1;
      if (iters % 10 == 0 && dw->rank == 0 && !trace_quiet()) printf(".");
    }
    if (dw->rank == 0 && !trace_quiet()) {
      dw->sync();
      printf("\nIters = %d Final proj-grad norm %E \n", iters, gradnorm);
      printf("tf took %lf seconds\n", wall_time() - st_time);
    }
    if (!bench && Plot_File.is_open()) Plot_File.close();
    return sweeps != maxsweep + 1;
  }

private:
  void make_seq() {
    seq_V[this->order] = '\0';
    for (int j = 0; j < this->order; j++) seq_V[j] = 'a' + j;
  }
  double gradient_norm_sharded() {
    World &dw = *this->world;
    const double *xs[16];
    int64_t ns[16];
    for (int i = 0; i < this->order; i++) {
      xs[i] = grad_W[i].data;
      ns[i] = grad_W[i].size;
    }
    PPXCK(dw, ppx_sqnorms(dw.ctx, xs, ns, this->order, dw.scal_dev));
    if (dw.np > 1) dw.allreduce(dw.scal_dev + dw.shard_mode, 1);
    double h[16], acc = 0;
    dw.fetch(dw.scal_dev, h, this->order);
    for (int i = 0; i < this->order; i++) acc += h[i];
    return std::sqrt(acc);
  }
};

#endif
