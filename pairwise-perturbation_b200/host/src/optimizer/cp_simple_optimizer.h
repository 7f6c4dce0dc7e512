// cp_simple_optimizer.h -- plain ALS: one full MTTKRP from V per mode (reference: cp_simple_optimizer.cxx:23-56).
#ifndef PPX_HOST_CP_SIMPLE_OPTIMIZER_H__
#define PPX_HOST_CP_SIMPLE_OPTIMIZER_H__

#include <utility>
#include "cp_als_optimizer.h"

template <typename dtype>
class CPSimpleOptimizer : public CPOptimizer<dtype> {
public:
  CPSimpleOptimizer(int order, int r, World &dw) : CPOptimizer<dtype>(order, r, dw) {}
  ~CPSimpleOptimizer() {}

  double step() {
    World &dw = *this->world;
    const int order = this->order;
    for (int i = 0; i < order; i++) {
      int index[16], lens_H[16];
      for (int j = 0; j < order; j++) index[j] = j;
      std::swap(index[i], index[order - 1]);  // swap_char(seq_V, i, order-1)
      Matrix<dtype> M(this->W[i].nrow, this->W[i].ncol, dw);
      KhatriRao_contract(M, *(this->V), this->W, index, lens_H, dw);
      this->solve_mode(i, M);
    }
    return 1.;
  }
};

#endif
