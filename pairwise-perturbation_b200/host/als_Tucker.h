// als_Tucker.h -- Tucker decomposition: HOSVD initialisation, HOOI with the dimension tree, pairwise perturbation.
// Same free-function surface as the reference's als_Tucker.h:8-91.
#ifndef PPX_HOST_ALS_TUCKER_H__
#define PPX_HOST_ALS_TUCKER_H__

#include "common.h"

void get_factor_matrices(Tensor<> &T, Matrix<> *factor_matrices, int ranks[], World &dw);   // als_Tucker.cxx:12-23
Tensor<> get_core_tensor(Tensor<> &T, Matrix<> *factor_matrices, int ranks[], World &dw);  // als_Tucker.cxx:25-64
void hosvd(Tensor<> &T, Tensor<> &core, Matrix<> *factor_matrices, int *ranks, World &dw); // als_Tucker.cxx:66-70

// Y = V x_{j != i} W_j in increasing mode order; i = -1 contracts every mode (als_Tucker.cxx:76-110)
void TTMc(Tensor<> &Y, Tensor<> &V, Matrix<> *W, int i, World &dw);

bool alsTucker(Tensor<> &V, Tensor<> &core, Matrix<> *W, double tol, double timelimit, int maxiter, World &dw);

void ttmc_map_DT(map<string, Tensor<>> &ttmc_map, map<string, string> &parent, map<string, string> &sibling,
                 Tensor<> &V, Matrix<> *W, string args, World &dw);                         // als_Tucker.cxx:178-230

bool alsTucker_DT(Tensor<> &V, Tensor<> &core, Matrix<> *W, double tol, double timelimit, int maxiter,
                  ofstream &Plot_File, int resprint, bool bench, World &dw);               // als_Tucker.cxx:240-424

void Build_ttmc_map(map<string, Tensor<>> &ttmc_map, Tensor<> &V, Matrix<> *W, const char *args, World &dw);

void alsTucker_DT_sub(Tensor<> &V, Tensor<> &core, Tensor<> &core_prev, Matrix<> *W, Matrix<> *dW, double tol,
                      double tol_init, double timelimit, int maxiter, double &st_time, ofstream &Plot_File,
                      double &diffnorm, int &iter, int resprint, World &dw);               // als_Tucker.cxx:476-669

void alsTucker_PP_sub(Tensor<> &V, Tensor<> &core, Tensor<> &core_prev, Matrix<> *W, Matrix<> *dW, double tol,
                      double tol_init, double timelimit, int maxiter, double &st_time, ofstream &Plot_File,
                      double &diffnorm, int &iter, int resprint, bool bench, World &dw);   // als_Tucker.cxx:679-896

bool alsTucker_PP(Tensor<> &V, Tensor<> &core, Matrix<> *W, double tol, double tol_init, double timelimit,
                  int maxiter, ofstream &Plot_File, int resprint, bool bench, World &dw);  // als_Tucker.cxx:906-962

#endif
