// cp_dt_lr_optimizer.h -- dimension-tree ALS whose two root contractions are kept up to date by rank-`update_rank`
// patches  cached += V x (U s) x VT  instead of full first contractions for num_subiteration - 2 of every
// num_subiteration sweeps (reference: src/optimizer/cp_dt_lr_optimizer.{h,cxx}; run.cxx -pp 2).
// The patch is a first contraction with "rank" update_rank (HBM-bound: one pass over V) followed by a rank expansion.
#ifndef PPX_HOST_CP_DT_LR_OPTIMIZER_H__
#define PPX_HOST_CP_DT_LR_OPTIMIZER_H__

#include "cp_dt_optimizer.h"

// cached (modes in increasing order, rank last) += V x_left (U s) x VT
// (cp_dt_lr_optimizer.cxx:127-159, cp_msdt_lr_optimizer.cxx:108-151)
template <typename dtype>
void lr_patch_root(Tensor<dtype> &cached, Tensor<dtype> &V, int left, Matrix<dtype> &Us, Matrix<dtype> &VT, World &dw) {
  const int r = (int)Us.ncol, R = (int)VT.ncol;
  const int64_t Mtot = V.size / V.lens[left];
  Tensor<dtype> temp;
  {
    int64_t lens[2] = {Mtot, r};
    temp = Tensor<dtype>(2, lens, dw, false);
  }
  PPXCK(dw, ppx_ttm_first(dw.ctx, V.data, V.lens, V.order, left, Us.data, Us.nrow, r, temp.data));
  PPXCK(dw, ppx_rank_expand_acc(dw.ctx, temp.data, Mtot, r, VT.data, VT.nrow, R, cached.data));
}

template <typename dtype>
class CPDTLROptimizer : public CPDTOptimizer<dtype> {
public:
  CPDTLROptimizer(int order, int r, int update_rank, int randomsvd_, World &dw)
      : CPDTOptimizer<dtype>(order, r, dw), randomsvd(randomsvd_ > 0) {
    num_subiteration = 5;
    rank = update_rank;
    initialize_low_rank_param();
  }
  CPDTLROptimizer(int order, int r, int update_rank, World &dw) : CPDTLROptimizer(order, r, update_rank, 0, dw) {}
  ~CPDTLROptimizer() {}

  void initialize_low_rank_param() {  // :24-28
    count_subiteration = 0;
    low_rank_decomp = false;
  }

  // :39-99: the root is the cached tensor patched by the last low-rank update, or a fresh first contraction
  void mttkrp_map_init(int left) {
    World &dw = *this->world;
    Tensor<dtype> &cached = this->first_subtree ? cached_tensor1 : cached_tensor2;
    if (low_rank_decomp && count_subiteration > 1) {
      update_cached_tensor(left);
      install_root(cached, left);
    } else {
      CPDTOptimizer<dtype>::mttkrp_map_init(left);
      cached = this->mttkrp_map[top_key()];
    }
    (void)dw;
  }

  void update_cached_tensor(int left) {  // :127-159
    Tensor<dtype> &cached = this->first_subtree ? cached_tensor1 : cached_tensor2;
    lr_patch_root(cached, *this->V, left, U, VT, *this->world);
  }

  double step() {  // :161-236
    World &dw = *this->world;
    const int order = this->order;
    if (this->first_subtree) {
      this->indexes = this->indexes1;
      this->left_index = this->left_index1;
    } else {
      this->indexes = this->indexes2;
      this->left_index = this->left_index2;
    }
    this->mttkrp_map.clear();
    mttkrp_map_init(this->left_index);
    const int n = (int)this->indexes.size();
    for (int i = 0; i < n; i++) {
      if (this->first_subtree && i < this->special_index) continue;
      if (!this->first_subtree && i > this->special_index) break;
      const bool lr_slot = (this->first_subtree && i == n - 1) || (!this->first_subtree && i == 0);
      if (lr_slot && count_subiteration >= 1) {
        const int m = this->indexes[i];
        Matrix<dtype> M = this->leaf(i);
        if (dw.np > 1 && m != dw.shard_mode) dw.allreduce(M.data, M.size);
        this->update_S(m);
        // gradient with the factor before the update (:209-211)
        PPXCK(dw, ppx_memcpy_d2d(dw.ctx, this->grad_W[m].data, M.data, sizeof(double) * M.size));
        PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 0, (int)M.nrow, (int)M.ncol, (int)M.ncol, 1.0, this->W[m].data,
                                 this->W[m].nrow, this->S.data, this->S.nrow, -1.0, this->grad_W[m].data, M.nrow));
        get_rankR_update_cholesky(rank, U, s, VT, M, this->W[m], this->S, randomsvd, 5000 + draws++);
        PPXCK(dw, ppx_gemm_small(dw.ctx, 0, 0, (int)M.nrow, (int)M.ncol, rank, 1.0, U.data, U.nrow, VT.data, VT.nrow,
                                 1.0, this->W[m].data, this->W[m].nrow));  // W += U s VT (:219-220)
        this->grams.refresh(this->W, m, dw);
        low_rank_decomp = true;
      } else {
        this->update_leaf(i);
      }
    }
    if (!this->first_subtree) count_subiteration++;
    if (count_subiteration == num_subiteration && !this->first_subtree) {
      this->special_index = (this->special_index + 1) % (order - 1);
      initialize_low_rank_param();
      if (this->special_index != 0) {
        this->left_index1 = (this->left_index1 + order - 1) % order;
        this->left_index2 = (this->left_index2 + order - 1) % order;
      } else {
        this->left_index = order - 1;
        this->left_index1 = this->left_index;
        this->left_index2 = (this->left_index + order - 1) % order;
      }
      this->update_indexes(this->indexes1, this->left_index1);
      this->update_indexes(this->indexes2, this->left_index2);
    }
    this->first_subtree = !this->first_subtree;
    return 0.5;
  }

  int num_subiteration;
  int count_subiteration;
  int rank;  // rank of the update
  bool low_rank_decomp;
  Tensor<dtype> cached_tensor1, cached_tensor2;
  Matrix<dtype> U;  // U diag(s), see get_rankR_update_cholesky
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
  // make `cached` (a copy of it) the root of the tree for root mode `left`
  void install_root(Tensor<dtype> &cached, int left) {
    vector<int> axes;
    for (int i = 0; i < this->order; i++)
      if (i != left) axes.push_back(i);
    const string top = top_key();
    this->mttkrp_map[top] = cached;
    this->axes_map[top] = axes;
  }
};

#endif
