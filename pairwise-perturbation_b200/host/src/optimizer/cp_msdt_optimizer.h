// cp_msdt_optimizer.h -- multi-sweep dimension tree: ONE first contraction updates N-1 factors, the root mode
// rotates N-1, N-2, ..., 0; step() returns (N-1)/N  (reference: src/optimizer/cp_msdt_optimizer.{h,cxx}).
#ifndef PPX_HOST_CP_MSDT_OPTIMIZER_H__
#define PPX_HOST_CP_MSDT_OPTIMIZER_H__

#include "cp_dt_optimizer.h"

template <typename dtype>
class CPMSDTOptimizer : public CPDTOptimizer<dtype> {
public:
  CPMSDTOptimizer(int order, int r, World &dw) : CPDTOptimizer<dtype>(order, r, dw) {
    this->left_index = order;  // cp_msdt_optimizer.cxx:28
  }
  ~CPMSDTOptimizer() {}

  // cp_msdt_optimizer.cxx:36-49
  void update_indexes() {
    this->left_index = (this->left_index + this->order - 1) % this->order;
    CPDTOptimizer<dtype>::update_indexes(this->indexes, this->left_index);
  }

  // cp_msdt_optimizer.cxx:173-208
  double step() {
    this->mttkrp_map.clear();
    update_indexes();
    this->mttkrp_map_init(this->left_index);
    for (int i = 0; i < (int)this->indexes.size(); i++) this->update_leaf(i);
    return 1. * (this->order - 1) / this->order;
  }
};

#endif
