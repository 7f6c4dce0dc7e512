// cp_als_optimizer.h -- optimizer base: holds V / W / grad_W, the Gram-Hadamard matrix S and the regulariser
// (reference: src/optimizer/cp_als_optimizer.{h,cxx}).
#ifndef PPX_HOST_CP_ALS_OPTIMIZER_H__
#define PPX_HOST_CP_ALS_OPTIMIZER_H__

#include <cassert>
#include "../../common.h"

template <typename dtype>
class CPOptimizer {
public:
  CPOptimizer(int order_, int r, World &dw) : order(order_), rank(r), world(&dw) { S = Matrix<>(r, r, dw); }
  virtual ~CPOptimizer() {}

  // cp_als_optimizer.cxx:40-65.  The pointers are borrowed: the owning CPD/Decomposition frees them (the reference
  // stores them in both places and would delete them twice on a second configure).
  void configure(Tensor<dtype> *input, Matrix<dtype> *mat, Matrix<dtype> *grad, double lambda_) {
    assert(input->order == order);
    for (int i = 0; i < order; i++) assert(mat[i].ncol == rank);
    V = input;
    W = mat;
    grad_W = grad;
    lambda = lambda_;
    grams.init(W, order, *world);
  }

  // S = Hadamard_{j != update_index} W_j^T W_j + lambda I  (cp_als_optimizer.cxx:20-38), from the Gram cache
  void update_S(int update_index) { grams.hadamard(update_index, lambda, S, *world); }

  int order;
  int rank;
  Tensor<dtype> *V = NULL;
  Matrix<dtype> *W = NULL;
  Matrix<dtype> *grad_W = NULL;
  World *world;
  Matrix<dtype> S;
  double lambda = 0.;  // the reference keeps lambda*I as a matrix `regul`

protected:
  GramCache grams;
  // gradient + Cholesky solve of one mode, then refresh its Gram
  // (cp_simple_optimizer.cxx:47-52, cp_dt_optimizer.cxx:226-232)
  void solve_mode(int mode, Matrix<dtype> &M) {
    World &dw = *world;
    if (dw.np > 1 && mode != dw.shard_mode) dw.allreduce(M.data, M.size);
    update_S(mode);
    solve_update_fused(M, S, W[mode], nullptr, 1.0, &grad_W[mode], nullptr, PPX_SOLVE_CHOL, dw);
    grams.refresh(W, mode, dw);
  }
};

#endif
