// common.h -- numerical helpers shared by the CP and Tucker drivers; same names and argument meaning as the
// reference's common.h (file:line citations are to /root/reference), implemented on the ppx C ABI.
#ifndef PPX_HOST_COMMON_H__
#define PPX_HOST_COMMON_H__

#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#include "ctf_shim.hpp"
using namespace CTF;
using std::map;
using std::ofstream;
using std::string;
using std::vector;
using std::cout;
using std::endl;

// ---- reference surface (common.h:7-115) -----------------------------------------------------------------------
void vec2str(vector<int> vec, string &seq_out);                                    // common.cxx:10-18
void build_V(Tensor<> &V, Matrix<> *W, int order, World &dw);                      // common.cxx:135-197
void mttkrp_map_DT(map<string, Tensor<>> &mttkrp_map, map<string, string> &parent, map<string, string> &sibling,
                   Tensor<> &V, Matrix<> *W, string args, World &dw);              // common.cxx:20-133
Matrix<> unroll_tensor_contraction(Tensor<> &T, int i);                            // common.cxx:205-223
void Construct_Dimension_Tree(map<string, string> &parent, map<string, string> &sibling, int start, int end);
void Normalize(Matrix<> *W, int N, World &dw);                                     // common.cxx:644-689
void SVD_solve(Matrix<> &M, Matrix<> &W, Matrix<> &S);                             // common.cxx:710-725
void cholesky_solve(Matrix<> &M, Matrix<> &W, Matrix<> &S);                        // common.cxx:727-737
void SVD_solve_mod(Matrix<> &M, Matrix<> &W, Matrix<> &W_init, Matrix<> &dW, Matrix<> &S, double ratio_step);
void KhatriRao_contract(Matrix<> &M, Tensor<> &V, Matrix<> *W, int *index, int *lens_H, World &dw);
void gradsubprob(Matrix<> &M, Matrix<> &S, Matrix<> &W, Matrix<> &grad_W);         // common.cxx:1002-1004
void gradient_CP(Tensor<> &V, Matrix<> *W, Matrix<> *grad_W, World &dw);           // common.cxx:1009-1052

// rank-R update of A towards M gamma^-1 (common.cxx:768-786): A_new ~ A + xU diag(xS) xVT.  Here xU already carries the
// singular values (xU = U diag(s), xS = ones): every caller only uses the products U diag(s) and VT.
// `random`: range finder of randomized_svd (common.cxx:691-708, one power iteration) instead of the exact
// truncated SVD; its test matrix is u(dw.seed, draw_id, .).
void get_rankR_update_cholesky(int R, Matrix<> &xU, Vector<> &xS, Matrix<> &xVT, Matrix<> &M, Matrix<> &A,
                               Matrix<> &gamma, bool random, uint64_t draw_id = 5000);
void matrixDot(Matrix<> &result, Matrix<> &matrix1, Matrix<> &matrix2);             // common.cxx:760-763

// input generators of test_ALS / pp_bench / run (test_ALS.cxx:222-286)
void laplacian_tensor(Tensor<> &V, int N, int s, bool sparse_V, World &dw);        // common.cxx:575-642
void fold_unfold(Tensor<> &X, Tensor<> &Y);                                        // common.cxx:870-882
double collinearity(const std::vector<double> &v1, const std::vector<double> &v2); // common.cxx:297-302
Tensor<> Gen_collinearity(int *lens, int dim, int R, double col_min, double col_max, World &dw);  // common.cxx:361-423
// u(seed, tensor_id, index) in [0,1) on the host: bit-identical to ppx_fill_uniform
double host_u01(uint64_t seed, uint64_t tensor_id, uint64_t index);

// ---- what the drivers use on top of it ------------------------------------------------------------------------
double wall_time();  // MPI_Wtime stand-in

// out = in contracted with Wx over mode-letter x.  `modes` lists the mode letters `in` still has (rank, if any, is
// last and implied by has_rank).  has_rank == false -> first contraction with V (GEMM, K1); true -> Hadamard-batched
// (K2).  Returns the tensor with the remaining modes + rank.
Tensor<> contract_mode(Tensor<> &in, const string &modes, bool has_rank, char x, Matrix<> &Wx, World &dw);

// Gram cache: G[j] = W_j^T W_j kept on the device, refreshed only for the factor that changed
// (the reference recomputes N-1 Grams for every mode update, als_CP.cxx:288-291).
struct GramCache {
  int N = 0, R = 0;
  vector<Matrix<>> G;
  void init(Matrix<> *W, int N, World &dw);
  void refresh(Matrix<> *W, int i, World &dw);
  // S = Hadamard_{j != skip} G[j] (+ lambda I)
  void hadamard(int skip, double lambda, Matrix<> &S, World &dw);
  // W = M S^-1 with S = Hadamard_{j != skip} G[j] + lambda I formed inside the inverse kernel; also the gradient
  // -M + W_old S and (W_init != NULL) dW = ratio (W - W_init).  Does NOT refresh G[skip].
  void solve(int skip, double lambda, Matrix<> &M, Matrix<> &W, Matrix<> *W_init, double ratio_step, Matrix<> *grad,
             Matrix<> *dW, int mode, World &dw);
};

// ||V - [[W]]||_F without materialising the reconstruction (replaces build_V + subtraction + norm2 at
// als_CP.cxx:183-187)
double cp_residual_norm(Tensor<> &V, Matrix<> *W, int order, World &dw);
// sqrt(sum_i ||grad_W[i]||^2)   (als_CP.cxx:174-181)
double gradient_norm(Matrix<> *grad_W, int order, World &dw);
// W = M S^-1 fused with grad = -M + W_old S and (W_init != NULL) dW = ratio (W - W_init)
void solve_update_fused(Matrix<> &M, Matrix<> &S, Matrix<> &W, Matrix<> *W_init, double ratio_step, Matrix<> *grad,
                        Matrix<> *dW, int mode, World &dw);

// ---- trace sink: what the reference prints, captured for tests / bench ---------------------------------------
struct TraceRow {
  double iter;  // iteration (or sweep count in CPD::als)
  double gradnorm;
  int pp_update;
  double diffV;
  double dtime;
};
struct TraceSink {
  vector<TraceRow> rows;
  vector<std::pair<int, int>> events;  // (0 = "DT starts from", 1 = "pairwise perturbation starts from", iter)
  vector<std::pair<int, int>> sweeps;  // (0 = DT sweep, 1 = PP sweep, 2 = PP operator build, iter)
  vector<double> bench_times;          // [DTtime] / [PPfirst] / [PPsecond] values in the order printed
  bool quiet = false;                  // suppress stdout
  bool skip_residual = false;          // do not evaluate ||V - [[W]]|| at print points (reported as -1)
};
TraceSink *&trace_sink();
bool trace_quiet();

#endif
