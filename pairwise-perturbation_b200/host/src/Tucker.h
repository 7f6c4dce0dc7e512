// Tucker.h -- Tucker<dtype>: the Tucker counterpart of CPD<dtype, Optimizer> (src/CP.h): holds the input tensor, the
// factor matrices and the core, initialises them by HOSVD and runs HOOI with the dimension tree (als) or with pairwise
// perturbation (als_pp) through the free functions of als_Tucker.h.
// The reference's own src/Tucker.h:1-146 is a non-compiling draft (a stale copy of old als_CP declarations under the
// DECOMPOSITION_H__ guard; src/Tucker.cxx is an old als_CP.cxx); what its Tucker path actually does is written in
// test_ALS.cxx:360-397: ranks per mode, hosvd(V, core, W, ranks), then alsTucker_DT or alsTucker_PP.  This class packages
// exactly that sequence behind the Decomposition interface (same constructor arguments and Init/als shape as CPD).
#ifndef PPX_HOST_TUCKER_H__
#define PPX_HOST_TUCKER_H__

#include "../als_Tucker.h"
#include "decomposition.h"

template <typename dtype>
class Tucker : public Decomposition<dtype> {
public:
  Tensor<dtype> core;  // R_0 x ... x R_{N-1}
  bool initialised = false;

  Tucker(int order, int size, int r, World &dw) : Decomposition<dtype>(order, size, r, dw) {}
  // per-mode sizes and ranks (test_ALS.cxx:366-379: coil-100 uses ranks 3,10,10,70)
  Tucker(int order, int *size, int *r, World &dw) : Decomposition<dtype>(order, size, r, dw) {}

  // Adopts `input` and the `mat` array like Decomposition::Init (mat[i] needs ncol == rank[i]; the contents are
  // replaced) and runs the HOSVD initialisation (test_ALS.cxx:381-382, als_Tucker.cxx:66-70).  With several GPUs the
  // factors are replicated in full: mat[i] has size[i] rows for every mode, the tensor holds the local slab.
  void Init(Tensor<dtype> *input, Matrix<dtype> *mat) {
    assert(input->order == this->order);
    for (int i = 0; i < this->order; i++) {
      assert(input->lens[i] == this->size[i] || (this->world->np > 1 && i == this->world->shard_mode));
      assert(mat[i].ncol == this->rank[i]);
    }
    if (this->W != NULL && this->W != mat) delete[] this->W;
    this->V = input;
    this->W = mat;
    hosvd_init();
  }
  void hosvd_init() {
    assert(this->V != NULL && this->W != NULL);
    hosvd(*this->V, core, this->W, this->rank, *this->world);
    initialised = true;
  }

  // HOOI with the dimension tree (als_Tucker.cxx:240-424).  Returns true when it stopped before maxiter + 1.
  bool als(double tol, double timelimit, int maxiter, int resprint, ofstream &Plot_File, bool bench = false) {
    assert(initialised);
    return alsTucker_DT(*this->V, core, this->W, tol, timelimit, maxiter, Plot_File, resprint, bench, *this->world);
  }
  // HOOI sweeps alternating with pairwise-perturbation sweeps (als_Tucker.cxx:906-962); tol_init is the switching
  // tolerance (-pp_res_tol)
  bool als_pp(double tol, double tol_init, double timelimit, int maxiter, int resprint, ofstream &Plot_File,
              bool bench = false) {
    assert(initialised);
    return alsTucker_PP(*this->V, core, this->W, tol, tol_init, timelimit, maxiter, Plot_File, resprint, bench,
                        *this->world);
  }
  // plain HOOI, one TTMc chain per mode (als_Tucker.cxx:112-170)
  bool als_plain(double tol, double timelimit, int maxiter) {
    assert(initialised);
    return alsTucker(*this->V, core, this->W, tol, timelimit, maxiter, *this->world);
  }
  double core_norm() const { return core.norm2(); }
};

#endif
