// cp_msdt_lr_optimizer.h -- multi-sweep dimension tree with one cached root tensor per mode; when a mode comes round
// as the root again its cached tensor is patched by the rank-`update_rank` update that mode received instead of being
// recomputed (reference: src/optimizer/cp_msdt_lr_optimizer.{h,cxx}; run.cxx -pp 3).
#ifndef PPX_HOST_CP_MSDT_LR_OPTIMIZER_H__
#define PPX_HOST_CP_MSDT_LR_OPTIMIZER_H__

#include "cp_dt_lr_optimizer.h"
#include "cp_msdt_optimizer.h"

template <typename dtype>
class CPMSDTLROptimizer : public CPMSDTOptimizer<dtype> {
public:
  CPMSDTLROptimizer(int order, int r, int update_rank, int randomsvd_, World &dw)
      : CPMSDTOptimizer<dtype>(order, r, dw), randomsvd(randomsvd_ > 0) {
    rank = update_rank;
    low_rank_decomp = false;
    is_cached.assign(order, false);
    cached_tensors.resize(order);
    old_W.resize(order);
  }
  CPMSDTLROptimizer(int order, int r, int update_rank, World &dw) : CPMSDTLROptimizer(order, r, update_rank, 0, dw) {}
  ~CPMSDTLROptimizer() {}

  void mttkrp_map_init(int left) {  // :35-80
    if (low_rank_decomp && is_cached[left]) {
      update_cached_tensor(left);
      vector<int> axes;
      for (int i = 0; i < this->order; i++)
        if (i != left) axes.push_back(i);
      const string top = top_key();
      this->mttkrp_map[top] = cached_tensors[left];
      this->axes_map[top] = axes;
    } else {
      CPDTOptimizer<dtype>::mttkrp_map_init(left);
      cached_tensors[left] = this->mttkrp_map[top_key()];
      old_W[left] = this->W[left];
      is_cached[left] = true;
    }
  }

  void update_cached_tensor(int left) {  // :108-151
    lr_patch_root(cached_tensors[left], *this->V, left, U, VT, *this->world);
    old_W[left] = this->W[left];
    is_cached[left] = true;
  }

  double step() {  // :153-205
    World &dw = *this->world;
    this->mttkrp_map.clear();
    this->update_indexes();
    mttkrp_map_init(this->left_index);
    const int n = (int)this->indexes.size();
    for (int i = 0; i < n; i++) {
      const int m = this->indexes[i];
      if (!is_cached[m] || i != n - 1) {
        this->update_leaf(i);
      } else {
        Matrix<dtype> M = this->leaf(i);
        if (dw.np > 1 && m != dw.shard_mode) dw.allreduce(M.data, M.size);
        this->update_S(m);
        PPXCK(dw, ppx_memcpy_d2d(dw.ctx, this->grad_W[m].data, M.data, sizeof(double) * M.size));
        PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 0, (int)M.nrow, (int)M.ncol, (int)M.ncol, 1.0, this->W[m].data,
                                 this->W[m].nrow, this->S.data, this->S.nrow, -1.0, this->grad_W[m].data, M.nrow));
        get_rankR_update_cholesky(rank, U, s, VT, M, old_W[m], this->S, randomsvd, 5000 + draws++);
        this->W[m] = old_W[m];  // W = old_W + U s VT (:193-196)
        PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 0, (int)M.nrow, (int)M.ncol, rank, 1.0, U.data, U.nrow, VT.data, VT.nrow,
                                 1.0, this->W[m].data, this->W[m].nrow));
        this->grams.refresh(this->W, m, dw);
        low_rank_decomp = true;
      }
    }
    return 1. * (this->order - 1) / this->order;
  }

  int rank;
  bool low_rank_decomp;
  vector<bool> is_cached;
  vector<Tensor<dtype>> cached_tensors;
  vector<Matrix<dtype>> old_W;
  Matrix<dtype> U;
  Vector<dtype> s;
  Matrix<dtype> VT;
  bool randomsvd;

protected:
  uint64_t draws = 0;
  string top_key() {
    vector<int> ids(this->order - 1);
    for (int i = 0; i < this->order - 1; i++) ids[i] = i;
    string top;
    vec2str(ids, top);
    return top;
  }
};

#endif
