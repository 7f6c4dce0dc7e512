// decomposition.h -- base class holding the input tensor, the factor matrices and their shapes
// (reference: src/decomposition.h:8-36, src/decomposition.cxx).
#ifndef PPX_HOST_DECOMPOSITION_H__
#define PPX_HOST_DECOMPOSITION_H__

#include <cassert>
#include "../common.h"

template <typename dtype>
class Decomposition {
public:
  Decomposition(int order_, int size_, int r, World &dw) : order(order_), world(&dw) {
    size = new int[order];
    rank = new int[order];
    for (int i = 0; i < order; i++) {
      size[i] = size_;
      rank[i] = r;
    }
  }
  // per-mode sizes and ranks.  (The reference dereferences an uninitialised World pointer here,
  // src/decomposition.cxx:23; this keeps the pointer instead.)
  Decomposition(int order_, int *size_, int *r, World &dw) : order(order_), world(&dw) {
    size = new int[order];
    rank = new int[order];
    for (int i = 0; i < order; i++) {
      size[i] = size_[i];
      rank[i] = r[i];
    }
  }
  Decomposition(const Decomposition &) = delete;
  Decomposition &operator=(const Decomposition &) = delete;

  virtual ~Decomposition() {
    // V is adopted but never freed, as in the reference (src/decomposition.cxx:39-41); W is owned.
    if (W != NULL) delete[] W;
    delete[] size;
    delete[] rank;
  }

  // adopts `input` and the `mat` array (src/decomposition.cxx:54-69)
  void Init(Tensor<dtype> *input, Matrix<dtype> *mat) {
    assert(input->order == order);
    for (int i = 0; i < order; i++) {
      // with a sharded leading mode the local extent is smaller than the global size
      assert(input->lens[i] == size[i] || (world->np > 1 && i == world->shard_mode));
      assert(mat[i].ncol == rank[i]);
    }
    if (W != NULL && W != mat) delete[] W;
    V = input;
    W = mat;
  }

  void print_V() const {
    assert(V != NULL);
    V->print();
  }
  void print_W(int i) const {
    assert(W != NULL);
    W[i].print();
  }

  Tensor<dtype> *V = NULL;  // input tensor
  int order;
  int *size;
  int *rank;
  Matrix<dtype> *W = NULL;  // output factors
  World *world;
};

#endif
