/* ppx.h -- C ABI of the B200-native engine for the ALS / pairwise-perturbation (PP) sweeps of CP and Tucker
 * decomposition.  This is the drop-in boundary: the reference (LinjianMa/pairwise-perturbation) performs every
 * operation below as a Cyclops-CTF Einstein-string expression or a ScaLAPACK-backed Matrix method; each entry point
 * names the reference call site (file:line under /root/reference) it replaces.
 *
 * Conventions
 *   - every tensor / matrix argument is a DEVICE pointer to dense FP64 data in the reference's (CTF) global order:
 *     first index fastest; a factor matrix W is s x R, column r contiguous (leading dimension ldw >= s);
 *   - an intermediate "T" of the dimension tree keeps its remaining tensor modes in increasing order and the rank
 *     index LAST (reference: common.cxx:44,53; als_CP.cxx:336);
 *   - every call is asynchronous on the context's stream and returns 0 (PPX_OK) or a negative PPX_E* code;
 *     ppx_last_error(ctx) gives the message.  No exceptions cross this boundary, no torch / C++ types appear in it;
 *   - buffers are caller-owned; one host thread per context.
 *   - there is NO CPU fallback: without a CUDA device ppx_ctx_create fails with PPX_ECUDA.
 */
#ifndef PPX_H_
#define PPX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ppx_ctx ppx_ctx;

enum {
  PPX_OK = 0,
  PPX_EINVAL = -1,       /* bad argument */
  PPX_ECUDA = -2,        /* CUDA runtime error (message has cudaGetErrorString) */
  PPX_ENOMEM = -3,       /* workspace / device allocation failed */
  PPX_ENCCL = -4,        /* NCCL missing or failed */
  PPX_EUNSUPPORTED = -5, /* shape outside what the kernels handle */
  PPX_ENUMERIC = -6      /* factorisation broke down (non-SPD matrix in Cholesky, ...) */
};

enum { PPX_SOLVE_CHOL = 0, PPX_SOLVE_SVD_PINV = 1 };

/* ---- context (replaces CTF::World; test_ALS.cxx:200) ------------------------------------------------------- */
const char *ppx_version(void);
/* stream: a cudaStream_t to run on, or NULL to create a private non-blocking stream. */
int ppx_ctx_create(int device, void *stream, size_t workspace_bytes, ppx_ctx **out);
int ppx_ctx_destroy(ppx_ctx *ctx);
int ppx_sync(ppx_ctx *ctx);
const char *ppx_last_error(ppx_ctx *ctx);
void *ppx_stream(ppx_ctx *ctx);
int ppx_device(ppx_ctx *ctx);
int ppx_sm_count(ppx_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t ppx_launch_count(ppx_ctx *ctx);

/* ---- memory (so the C++ host layer needs no CUDA headers) -------------------------------------------------- */
int ppx_malloc(ppx_ctx *ctx, size_t bytes, void **dptr);
int ppx_free(ppx_ctx *ctx, void *dptr);
int ppx_host_alloc(ppx_ctx *ctx, size_t bytes, void **hptr); /* pinned */
int ppx_host_free(ppx_ctx *ctx, void *hptr);
int ppx_memcpy_h2d(ppx_ctx *ctx, void *dst, const void *src, size_t bytes); /* async on the stream */
int ppx_memcpy_d2h(ppx_ctx *ctx, void *dst, const void *src, size_t bytes); /* async; ppx_sync before reading */
int ppx_memcpy_d2d(ppx_ctx *ctx, void *dst, const void *src, size_t bytes);
/* `height` columns of `width_bytes` each, column j at dst + j*dst_pitch / src + j*src_pitch (bytes): the row range of a
 * column-major matrix -- how the rows of a sharded factor are cut out of / pasted into the replicated one. */
int ppx_memcpy2d_d2d(ppx_ctx *ctx, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes,
                     size_t height);
int ppx_memset_zero(ppx_ctx *ctx, void *dst, size_t bytes);
int ppx_mem_info(ppx_ctx *ctx, size_t *free_bytes, size_t *total_bytes);

/* ---- events / graphs (timing as als_CP.cxx:167,189 does with MPI_Wtime; CUDA graphs for the PP sweep) ------- */
int ppx_event_create(ppx_ctx *ctx, void **ev);
int ppx_event_destroy(ppx_ctx *ctx, void *ev);
int ppx_event_record(ppx_ctx *ctx, void *ev);
int ppx_event_elapsed_ms(ppx_ctx *ctx, void *ev_start, void *ev_stop, float *ms); /* synchronises ev_stop */
int ppx_graph_begin(ppx_ctx *ctx);               /* start stream capture */
int ppx_graph_end(ppx_ctx *ctx, void **graph);   /* end capture, instantiate; *graph is an executable graph */
int ppx_graph_launch(ppx_ctx *ctx, void *graph);
int ppx_graph_destroy(ppx_ctx *ctx, void *graph);

/* ---- deterministic synthetic data (replaces fill_random; test_ALS.cxx:272,282,337-338) ---------------------- */
/* out[i] = lo + (hi-lo) * u(seed, tensor_id, start+i), u = SplitMix64 finaliser of the counter (53 bits). */
int ppx_fill_uniform(ppx_ctx *ctx, double *out, int64_t n, uint64_t seed, uint64_t tensor_id, int64_t start,
                     double lo, double hi);
/* The local slab of a tensor whose leading mode (global extent L_global) is sharded: rows [row_begin, row_begin+L_local)
 * of each of the n_cols columns, out[i + L_local*c] = lo + (hi-lo) * u(seed, tensor_id, row_begin + i + L_global*c) --
 * the same values ppx_fill_uniform gives the whole tensor, so any rank generates its slice without communication. */
int ppx_fill_uniform_rows(ppx_ctx *ctx, double *out, int64_t L_local, int64_t L_global, int64_t row_begin,
                          int64_t n_cols, uint64_t seed, uint64_t tensor_id, double lo, double hi);
/* out (s^(2d) doubles) = the Laplacian tensor of order 2d, V[a1,b1,..,ad,bd] = sum_k D[a_k,b_k] prod_{m!=k} delta(a_m,b_m),
 * D = tridiag(-1,2,-1): what laplacian_tensor builds from identity tensors (common.cxx:575-642; generators 'p','p2'). */
int ppx_fill_laplacian(ppx_ctx *ctx, double *out, int d, int64_t s);

/* ---- K1: first tensor-times-matrix contraction of the dimension tree ---------------------------------------
 * out[rest, r] = sum_x V[.., x, ..] * Wx[x, r]      (remaining modes in order, rank last)
 * replaces common.cxx:56, als_CP.cxx:378-379, cp_dt_optimizer.cxx:158-159, cp_msdt_optimizer.cxx:142-143,
 * common.cxx:963.  FP64 tensor-core (DMMA) GEMM. */
int ppx_ttm_first(ppx_ctx *ctx, const double *V, const int64_t *lens, int N, int x, const double *Wx, int64_t ldw,
                  int R, double *out);

/* Several ADJACENT modes x_first .. x_first+n_modes-1 contracted at once:
 *   out[rest, r] = sum_{x_1..x_n} V[.., x_1, .., x_n, ..] * prod_j W[j][x_j, r]
 * i.e. the first contraction followed by the Hadamard-batched contractions of the reference's tree
 * (common.cxx:56 then :83) as ONE GEMM against the Khatri-Rao rows of the factors; the level-1 intermediates are
 * never written.  W / ldw: HOST arrays of n_modes entries.  Needs 8*R*prod(lens[x_j]) bytes of workspace. */
int ppx_ttm_multi(ppx_ctx *ctx, const double *V, const int64_t *lens, int N, int x_first, int n_modes,
                  const double *const *W, const int64_t *ldw, int R, double *out);

/* ---- K2: Hadamard-batched contraction (rank index in all three operands) -----------------------------------
 * out[rest', r] = sum_x T[.., x, .., r] * Wx[x, r];  lens = the k non-rank modes of T.
 * replaces common.cxx:83,128; als_CP.cxx:258-259,407-408; cp_dt_optimizer.cxx:184-185. */
int ppx_mttv(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int x, const double *Wx, int64_t ldw, int R,
             double *out);
/* The three Hadamard contractions of ONE order-3 (+ rank) tensor in one pass over it -- what Build_mttkrp_map does to a
 * level-1 tensor with several consumers (als_CP.cxx:394-408: "ijkr,ir->jkr", "ijkr,jr->ikr", "ijkr,kr->ijr" read the
 * same intermediate once each):
 *   T[l, x, t, r], lens3 = {s_l, s_x, s_t};  out_l[x,t,r] = sum_l T Wl[l,r];  out_x[l,t,r] = sum_x T Wx[x,r];
 *   out_t[l,x,r] = sum_t T Wt[t,r].  A NULL output is skipped (its factor may be NULL too).
 * Streams T through shared memory with bulk copies; out_x is summed over x-tiles in a fixed order (workspace scratch).
 * Shapes the one-pass kernel does not take (odd s_l, a single requested output, small tensors, no scratch) run as one
 * ppx_mttv per output. */
int ppx_mttv3(ppx_ctx *ctx, const double *T, const int64_t *lens3, int R, const double *Wl, int64_t ldl,
              const double *Wx, int64_t ldx, const double *Wt, int64_t ldt, double *out_l, double *out_x, double *out_t);
/* two factors at once (leaf under a 3-mode node): out[rest'', r] = sum_{x1,x2} T * W1[x1,r] * W2[x2,r], x1 < x2.
 * replaces als_CP.cxx:281-283. */
int ppx_mttv2(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int x1, const double *W1, int64_t ldw1,
              int x2, const double *W2, int64_t ldw2, int R, double *out);
/* fused K1 -> K2 when the level-1 intermediate has a single consumer: contracts x1 by GEMM and x2 (x2 != x1)
 * Hadamard-batched without writing the level-1 tensor (N=4 ALS tree: abcd x W_c x W_d -> ab*). */
int ppx_ttm_first_mttv(ppx_ctx *ctx, const double *V, const int64_t *lens, int N, int x1, const double *W1,
                       int64_t ldw1, int x2, const double *W2, int64_t ldw2, int R, double *out);

/* ---- K3: PP first-order correction, all operators of one mode in one launch --------------------------------
 * M_out[i', r] = M0[i', r] + sum_j  sum_q op_j[..] * dW_j[q, r]
 * op_j is (which[j]==0) s_other[j] x s_i x R, contracted over its FIRST index (als_CP.cxx:785), or
 *         (which[j]==1) s_i x s_other[j] x R, contracted over its SECOND index (als_CP.cxx:793).
 * ops / dW / which / s_other are HOST arrays of n_ops entries (device pointers inside).
 * replaces als_CP.cxx:778-794. */
int ppx_pp_correct(ppx_ctx *ctx, const double *M0, const double *const *ops, const int *which,
                   const double *const *dW, const int64_t *s_other, int n_ops, int64_t s_i, int R, double *M_out);

/* ---- K4: Gram and Hadamard of Grams ----------------------------------------------------------------------- */
/* G = W^T W (R x R).  replaces the W["ki"]*W["kj"] factors of als_CP.cxx:288-291. */
int ppx_gram(ppx_ctx *ctx, const double *W, int64_t s, int64_t ldw, int R, double *G);
/* S = Hadamard_{j != skip} G[j] + lambda*I.  G: HOST array of nG device pointers, multiplied in array order.
 * replaces als_CP.cxx:288-292,573-579,796-802; cp_als_optimizer.cxx:33-37. */
int ppx_hadamard_grams(ppx_ctx *ctx, const double *const *G, int nG, int skip, int R, double lambda, double *S);

/* ---- K5 + K6: R x R solve fused with gradient, dW and norms ------------------------------------------------
 * grad_out = -M + W_old * S                       (als_CP.cxx:296,582,811; cp_dt_optimizer.cxx:229-230)
 * W_new    = M * S^-1                             (common.cxx:710-725 SVD_PINV | 727-737 CHOL)
 * if W_init: dW_out = ratio_step*(W_new - W_init); if ratio_step != 1: W_new = W_init + dW_out (common.cxx:753-756)
 * sq_norms_out (device, 3 doubles, may be NULL) = { ||W_new||_F^2, ||dW_out||_F^2, ||grad_out||_F^2 }.
 * W is read (old) and overwritten (new).  grad_out / dW_out / W_init may be NULL. */
int ppx_solve_update(ppx_ctx *ctx, const double *M, const double *S, double *W, int64_t s, int R,
                     const double *W_init, double ratio_step, int mode, double *grad_out, double *dW_out,
                     double *sq_norms_out);
/* Same with S = Hadamard_{j != skip} G[j] + lambda*I formed inside the inverse kernel from the cached Grams
 * (G: HOST array of nG device pointers) -- one launch less per mode update of the PP sweep. */
int ppx_solve_update_g(ppx_ctx *ctx, const double *M, const double *const *G, int nG, int skip, double lambda,
                       double *W, int64_t s, int R, const double *W_init, double ratio_step, int mode,
                       double *grad_out, double *dW_out, double *sq_norms_out);
/* The two halves of ppx_solve_update_g as separate calls, so that the R x R inverse (which depends only on the
 * Grams) can overlap the PP correction of the same mode on the side stream (ppx_side_begin/end/join):
 *   S_out (may be NULL) = Hadamard_{j != skip} G[j] + lambda*I ;  Sinv_out = S^-1   (R x R device buffers)         */
int ppx_spd_inverse_g(ppx_ctx *ctx, const double *const *G, int nG, int skip, double lambda, int R, int mode,
                      double *S_out, double *Sinv_out);
/*   grad_out = -M + W_old*S ; W = M*Sinv ; dW_out = ratio_step*(W - W_init)  (S may be NULL iff grad_out is NULL)  */
int ppx_solve_apply(ppx_ctx *ctx, const double *M, const double *S, const double *Sinv, double *W, int64_t s, int R,
                    const double *W_init, double ratio_step, double *grad_out, double *dW_out);
/* Linv_out (R x R, lower triangular, column-major) = L^-1 with S = L L^T (S symmetric positive definite): what the
 * solve_tri calls of get_rankR_update_cholesky apply from the right (common.cxx:774-785). */
int ppx_spd_factor_inverse(ppx_ctx *ctx, const double *S, int R, double *Linv_out);
/* C (m x n) = alpha op(A) op(B) + beta C, column-major, op = transpose when the flag is non-zero; for the small dense
 * products of the low-rank-update optimizers (common.cxx:760-786). */
int ppx_gemm_small(ppx_ctx *ctx, int transa, int transb, int m, int n, int k, double alpha, const double *A,
                   int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc);
/* out[m, c] += sum_q T[m, q] VT[q, c] for m < Mtot, c < R, q < r <= 16: the rank-r patch
 * cached += V x (U s) x VT of a dimension-tree root tensor, after T = V x (U s) was formed by ppx_ttm_first with
 * "rank" r (cp_dt_lr_optimizer.cxx:152-158, cp_msdt_lr_optimizer.cxx:142-146). */
int ppx_rank_expand_acc(ppx_ctx *ctx, const double *T, int64_t Mtot, int r, const double *VT, int64_t ldvt, int R,
                        double *out);
/* Fork / join for work that may overlap the main stream; valid eagerly and inside ppx_graph_begin/end:
 *   ppx_side_begin: the side stream waits for everything enqueued so far; subsequent calls go to the side stream
 *   ppx_side_end  : subsequent calls go to the main stream again (the side work keeps running concurrently)
 *   ppx_side_join : the main stream waits for the side work enqueued between begin and end                         */
int ppx_side_begin(ppx_ctx *ctx);
int ppx_side_end(ppx_ctx *ctx);
int ppx_side_join(ppx_ctx *ctx);
/* The same with four independent lanes (ppx_side_* is lane 0): the PP sweep runs the R x R inverse of mode i on lane 0
 * and the part of mode i+1's correction that does not depend on mode i's update on lane 1 (als_CP.cxx:774-812 has no
 * such overlap: CTF executes one contraction at a time).  Entry points that take scratch from the context workspace
 * must not run on two lanes at once; ppx_pp_correct does not split its grid (no scratch) while a lane is open.   */
int ppx_lane_begin(ppx_ctx *ctx, int lane);
int ppx_lane_end(ppx_ctx *ctx);
int ppx_lane_join(ppx_ctx *ctx, int lane);
/* Diagnostics: writes the GPU's nanosecond clock (%globaltimer) to *dev_slot, in stream order on the current stream or
 * lane -- a timeline of a captured sweep without a profiler (PPX_PP_TRACE=1, host/als_CP.cxx). */
int ppx_stamp(ppx_ctx *ctx, unsigned long long *dev_slot);
/* Normalize (common.cxx:680-688): every W_i scaled to the geometric mean of the Frobenius norms.  W, s: HOST
 * arrays.  If G != NULL, G[i] (cached Gram of W_i) is rescaled consistently. */
int ppx_normalize(ppx_ctx *ctx, double *const *W, const int64_t *s, int N, int R, double *const *G);
/* Same, but ||W_i||_F^2 is taken as trace(G[i]) instead of being recomputed from W_i: with the leading mode sharded
 * over GPUs the local rows of W_i do not give the global norm, the (all-reduced) Gram does. */
int ppx_normalize_g(ppx_ctx *ctx, double *const *W, const int64_t *s, int N, int R, double *const *G);
/* ppx_normalize_g fused with the 2N squared norms the PP switching test reads after every approximate sweep
 * (als_CP.cxx:657-664, :825), one launch:  sq_out_dev[2i] = ||dW[i]||^2 (0 if dW == NULL),
 * sq_out_dev[2i+1] = ||W[i]||^2 after the rescale. */
int ppx_normalize_norms(ppx_ctx *ctx, double *const *W, const double *const *dW, const int64_t *s, int N, int R,
                        double *const *G, double *sq_out_dev);
/* out_dev[j] = sum of squares of X[j][0..n[j]) for j < count (norm2()^2; als_CP.cxx:176-178,598-600). */
int ppx_sqnorms(ppx_ctx *ctx, const double *const *X, const int64_t *n, int count, double *out_dev);
/* out_dev[i] = <X[i], Y[i]> for count <= 16 pairs of n[i] doubles (deterministic).  With M = the MTTKRP of the last
 * mode, W its updated factor, S = Hadamard of the other Grams and G = W^T W, the residual follows without touching V:
 * ||V - [[W]]||^2 = ||V||^2 - 2 <M, W> + <S, G>   (a monitor: it cancels near convergence). */
int ppx_dots(ppx_ctx *ctx, const double *const *X, const double *const *Y, const int64_t *n, int count,
             double *out_dev);
/* dW = W - W_prev; W_prev = W; sq_out_dev = { ||dW||^2, ||W||^2 }  (als_CP.cxx:596-600). */
int ppx_diff_update(ppx_ctx *ctx, const double *W, double *W_prev, double *dW, int64_t n, double *sq_out_dev);
/* y = alpha*x + beta*y  (elementwise; M += F, als_CP.cxx:294). */
int ppx_axpby(ppx_ctx *ctx, double alpha, const double *x, double beta, double *y, int64_t n);

/* ---- K7: residual ||V - [[W_0..W_{N-1}]]||_F without materialising the reconstruction -----------------------
 * sq_out_dev[0] = sum of squares of the difference.  replaces common.cxx:135-197 + als_CP.cxx:183-187. */
int ppx_cp_residual(ppx_ctx *ctx, const double *V, const int64_t *lens, int N, const double *const *W, int R,
                    double *sq_out_dev);
/* V_out = [[W_0..W_{N-1}]]  (build_V, common.cxx:135-197; used to make the synthetic tensor 'r'). */
int ppx_cp_reconstruct(ppx_ctx *ctx, const int64_t *lens, int N, const double *const *W, int R, double *V_out);

/* ---- Tucker (K8-K11) --------------------------------------------------------------------------------------- */
/* out[.., q, ..] = sum_x T[.., x, ..] * Wx[x, q]  (rank replaces mode x in place).
 * replaces als_Tucker.cxx:102,224,372,389-391,464-465,845,858. */
int ppx_ttm(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int x, const double *Wx, int64_t ldw, int Q,
            double *out);
/* out += the same contraction (Tucker PP correction Y += T^(i,j) x_j dW_j, als_Tucker.cxx:845,858). */
int ppx_ttm_acc(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int x, const double *Wx, int64_t ldw,
                int Q, double *out);
/* MTM[p,q] = sum_rest T[..p..] T[..q..]  (common.cxx:205-223). */
int ppx_unfold_gram(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int i, double *MTM);
/* U (s x r) = eigenvectors of the r largest eigenvalues of the symmetric PSD matrix MTM (s x s), in decreasing
 * order -- what MTM.svd(U,S,VT,r) returns for such a matrix (als_Tucker.cxx:20,402,627,868).  MTM is destroyed.
 * evals_out (device, r doubles) may be NULL. */
int ppx_sym_eig_topk(ppx_ctx *ctx, double *MTM, int64_t s, int r, double *U, double *evals_out);
/* Same, warm started: `basis` (device, s x s, column-major) is an orthogonal matrix the iteration starts from when
 * basis_valid != 0 -- typically the eigenvectors this call left there for the same mode one HOOI sweep earlier -- and
 * receives ALL s eigenvectors on return.  Any orthogonal start gives the same result up to rounding; a close one
 * leaves a few Jacobi sweeps instead of a dozen.  basis may be NULL (cold start, nothing stored). */
int ppx_sym_eig_topk_warm(ppx_ctx *ctx, double *MTM, int64_t s, int r, double *U, double *evals_out, double *basis,
                          int basis_valid);
/* U <- U * diag(sign(diag(U^T Uref))), sign(b) = +1 if b > 0 else -1  (als_Tucker.cxx:632-643,874-885). */
int ppx_sign_align(ppx_ctx *ctx, double *U, const double *Uref, int64_t s, int r);
/* sq_out_dev[0] = sum of squares of (a - b), nothing else is written (Tucker residual, als_Tucker.cxx:309-310). */
int ppx_diff_sqnorm(ppx_ctx *ctx, const double *a, const double *b, int64_t n, double *sq_out_dev);
/* B (n x m) = A^T, A is m x n column-major (W_T, als_Tucker.cxx:296-300). */
int ppx_transpose(ppx_ctx *ctx, const double *A, int64_t m, int64_t n, double *B);

/* ---- K12: collectives (replace CTF-internal MPI) ------------------------------------------------------------ */
/* rows [*begin, *end) of a mode of size s owned by `rank` of `nranks` (first s%nranks ranks get one more). Pure host. */
int ppx_shard_range(int64_t s, int nranks, int rank, int64_t *begin, int64_t *end);
int ppx_comm_unique_id(void *id128);                                  /* 128 bytes; rank 0 calls, others receive */
int ppx_comm_init(ppx_ctx *ctx, const void *id128, int nranks, int rank);
/* The same without a launcher-side broadcast, for the SPMD command lines (the reference's mains call MPI_Init and
 * build World(argc, argv), test_ALS.cxx:58-60,200): rank 0 creates the id and hands it to the other nranks-1 processes
 * over TCP on addr:port (it listens, they connect, retrying for up to timeout_s seconds), then every rank calls
 * ppx_comm_init.  One node or several; nothing else ever goes over that socket. */
int ppx_comm_bootstrap(ppx_ctx *ctx, int nranks, int rank, const char *addr, int port, int timeout_s);
/* 1 when small all-reduces (ppx_allreduce_packed, up to 96 Ki doubles per call and 8 buffers) run as ONE kernel over
 * NVLink peer memory instead of NCCL: every rank stages its values in a buffer all peers have mapped (cudaIpc), raises
 * an epoch flag in every peer's memory, waits for the peers' flags and adds the P staged copies in rank order (bit-
 * identical on every rank).  Set up collectively inside ppx_comm_init; 0 when any rank could not map a peer (other
 * node, no peer access, IPC not permitted) or PPX_NO_P2P is set -- then every call is NCCL. */
int ppx_comm_p2p(ppx_ctx *ctx);
int ppx_comm_size(ppx_ctx *ctx);
int ppx_comm_rank(ppx_ctx *ctx);
/* in-place sum over ranks of n buffers as ONE NCCL group (bufs/sizes: HOST arrays; no-op when nranks == 1). */
int ppx_allreduce_packed(ppx_ctx *ctx, double *const *bufs, const int64_t *sizes, int n);

/* Personalised exchange (one NCCL group of send/recv pairs): sendcounts[k] doubles from sendbufs[k] go to rank k,
 * recvcounts[k] doubles from rank k land in recvbufs[k]; the entry of the calling rank is a device copy.  HOST arrays
 * of comm_size entries.  Used for the mode-0 unfolding Gram of a tensor sharded along mode 0 (HOSVD on several GPUs):
 * every rank gets ALL rows of its share of the last mode instead of moving the whole tensor through all-reduces. */
int ppx_alltoallv(ppx_ctx *ctx, const double *const *sendbufs, const int64_t *sendcounts, double *const *recvbufs,
                  const int64_t *recvcounts);

#ifdef __cplusplus
}
#endif
#endif /* PPX_H_ */
