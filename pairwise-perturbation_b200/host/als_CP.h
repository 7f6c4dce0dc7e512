// als_CP.h -- CP-ALS with dimension tree and pairwise perturbation: the reference's free-function surface
// (als_CP.h:17-122 of /root/reference), same names, argument order and return meaning.
#ifndef PPX_HOST_ALS_CP_H__
#define PPX_HOST_ALS_CP_H__

#include "common.h"

// plain ALS, one full MTTKRP per mode (als_CP.cxx:20-115)
bool alsCP(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *F, double tol, double timelimit, int maxiter,
           World &dw);

// ALS with the balanced dimension tree (als_CP.cxx:127-320).  Returns true when it stopped before maxiter+1.
bool alsCP_DT(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *F, double tol, double timelimit, int maxiter,
              double lambda, ofstream &Plot_File, int resprint, bool bench, World &dw);

// "cd" -> "ab*" (als_CP.cxx:323-350)
void stringbuilder_mttkrp(const char *seq, char *seq_return, int N, World &dw);

// PP operator build; key = contracted modes in increasing order (als_CP.cxx:352-409)
void Build_mttkrp_map(map<string, Tensor<>> &mttkrp_map, Tensor<> &V, Matrix<> *W, const char *seq, World &dw);

double alsCP_DT_sub(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *dW, Matrix<> *F, double tol, double tol_init,
                    double timelimit, int maxiter, double &st_time, double lambda, ofstream &Plot_File,
                    double &projnorm, int &iter, int resprint, World &dw);

double alsCP_PP_sub(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *dW, Matrix<> *F, double tol, double tol_init,
                    double timelimit, int maxiter, double &st_time, double lambda, double ratio_step,
                    ofstream &Plot_File, double &projnorm, int &iter, int resprint, bool bench, World &dw);

double alsCP_PP_partupdate_sub(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *dW, Matrix<> *F, double tol,
                               double tol_init, double timelimit, int maxiter, double update_percentage,
                               double &st_time, double lambda, double ratio_step, ofstream &Plot_File,
                               double &projnorm, int &iter, int resprint, bool bench, World &dw);

// DT <-> PP switching driver (als_CP.cxx:1082-1137)
bool alsCP_PP(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *F, double tol, double tol_init, double timelimit,
              int maxiter, double lambda, double ratio_step, ofstream &Plot_File, int resprint, bool bench, World &dw);

// partial-update PP (als_CP.cxx:1146-1207)
bool alsCP_PP_partupdate(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, Matrix<> *F, double tol, double tol_init,
                         double timelimit, int maxiter, double lambda, double ratio_step, double update_percentage,
                         ofstream &Plot_File, int resprint, bool bench, World &dw);

vector<int> sort_indexes(const vector<double> &v);  // als_CP.cxx:835-843

// ---- timing helpers (bench.py): exactly n exact sweeps / one PP phase, no logging, no host synchronisation ----
void alsCP_DT_sweeps(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, int n_sweeps, double lambda, World &dw);
// builds the PP operators at W (timed: ms_build), captures the approximate sweep, then runs n_sweeps of it
// (timed: ms_sweeps); both with CUDA events on the world's stream
void alsCP_PP_phase_timed(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, int n_sweeps, double lambda, double ratio_step,
                          World &dw, float *ms_build, float *ms_sweeps);

// ---- probes for the full-size parity checks (tests/, bench.py parity_probe): what one tree pass / one operator build
// produce at FIXED factors, through exactly the calls the sweeps make -------------------------------------------
// M_out[i] = MTTKRP of mode i from the dimension tree (leaf_mttkrp of every mode, no update in between)
void alsCP_DT_mttkrps(Tensor<> &V, Matrix<> *W, Matrix<> *M_out, World &dw);
// all pair operators and singles of a PP phase built at W (build_pp_operators; key = contracted modes)
void alsCP_PP_operators(Tensor<> &V, Matrix<> *W, map<string, Tensor<>> &ops, World &dw);

#endif
