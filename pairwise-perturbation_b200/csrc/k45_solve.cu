// K4 -- Gram matrices and their Hadamard product  S = o_{j != i} (W_j^T W_j) + lambda I
//       (als_CP.cxx:288-292,573-579,796-802; cp_als_optimizer.cxx:33-37)
// K5 -- R x R solve  W = M S^-1  fused with the gradient  G = -M + W_old S  and the PP update
//       dW = ratio*(W - W_init)   (common.cxx:710-758; als_CP.cxx:296,582,811-812)
// K6 -- Normalize (common.cxx:680-688)
//
// The R x R inverse is formed once by ONE CTA (Cholesky: S = L L^T, S^-1 = L^-T L^-1; or, for the reference's
// SVD_solve semantics, a cyclic Jacobi eigen-decomposition S = Q diag(e) Q^T, S^-1 = Q diag(1/e) Q^T with no
// truncation, common.cxx:720-722), then applied to the s x R right-hand side by a grid of row tiles.
#include "ppx_internal.h"

namespace {

// ---- Gram: one warp per (a,b), a <= b, mirrored --------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_kernel(const double *__restrict__ W, int64_t s, int64_t ldw, int R,
                                                   double *__restrict__ G) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int npairs = R * (R + 1) / 2;
  if (warp_global >= npairs) return;
  // unrank (a,b), a <= b, from the row-major upper triangle
  int a = 0, rem = warp_global;
  while (rem >= R - a) {
    rem -= R - a;
    a++;
  }
  const int b = a + rem;
  const double *wa = W + (int64_t)a * ldw, *wb = W + (int64_t)b * ldw;
  double acc0 = 0.0, acc1 = 0.0;
  int64_t i = lane;
  for (; i + 32 < s; i += 64) {
    acc0 += wa[i] * wb[i];
    acc1 += wa[i + 32] * wb[i + 32];
  }
  for (; i < s; i += 32) acc0 += wa[i] * wb[i];
  double v = ppx_warp_sum(acc0 + acc1);
  if (lane == 0) {
    G[a + R * b] = v;
    G[b + R * a] = v;
  }
}

struct HadArgs {
  const double *g[16];
  int n;
};
__global__ void hadamard_kernel(HadArgs h, int R, double lambda, double *__restrict__ S) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * R) return;
  double v = h.g[0][idx];
  for (int j = 1; j < h.n; j++) v *= h.g[j][idx];
  if (lambda != 0.0 && (idx / R) == (idx % R)) v += lambda;
  S[idx] = v;
}

// ---- R x R inverse, one CTA ------------------------------------------------------------------------------------
// shared: A[R][ld] (ld = R+1) and B[R][ld]
__global__ void __launch_bounds__(256) spd_inverse_chol_kernel(const double *__restrict__ S, int R,
                                                               double *__restrict__ Sinv) {
  extern __shared__ double sm[];
  const int ld = R + 1;
  double *A = sm;           // becomes L (lower)
  double *B = sm + R * ld;  // becomes L^-1 (lower)
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int idx = tid; idx < R * R; idx += nt) {
    int i = idx % R, j = idx / R;
    A[i * ld + j] = S[idx];
    B[i * ld + j] = 0.0;
  }
  __syncthreads();
  // right-looking Cholesky, lower triangle
  for (int k = 0; k < R; k++) {
    const double d = sqrt(A[k * ld + k]);  // every thread reads the same (pre-update) value
    __syncthreads();
    if (tid == 0) A[k * ld + k] = d;
    for (int i = k + 1 + tid; i < R; i += nt) A[i * ld + k] /= d;
    __syncthreads();
    const int m = R - k - 1;  // trailing update on the lower triangle of the m x m block
    for (int idx = tid; idx < m * m; idx += nt) {
      int i = k + 1 + idx / m, j = k + 1 + idx % m;
      if (j <= i) A[i * ld + j] -= A[i * ld + k] * A[j * ld + k];
    }
    __syncthreads();
  }
  // L^-1 by forward substitution, one thread per column
  for (int j = tid; j < R; j += nt) {
    for (int i = j; i < R; i++) {
      double v = (i == j) ? 1.0 : 0.0;
      for (int k = j; k < i; k++) v -= A[i * ld + k] * B[k * ld + j];
      B[i * ld + j] = v / A[i * ld + i];
    }
  }
  __syncthreads();
  // S^-1 = L^-T L^-1 :  Sinv[i][j] = sum_{k >= max(i,j)} Linv[k][i] Linv[k][j]
  for (int idx = tid; idx < R * R; idx += nt) {
    int i = idx % R, j = idx / R;
    if (i >= j) {
      double v = 0.0;
      for (int k = i; k < R; k++) v += B[k * ld + i] * B[k * ld + j];
      Sinv[i + R * j] = v;
      Sinv[j + R * i] = v;
    }
  }
}

// Cyclic Jacobi (parallel round-robin ordering) on a symmetric matrix in shared memory.
// A[n][ld], Q[n][ld]; n even (padded with an identity row/column when R is odd).
__device__ void jacobi_eig_shared(double *A, double *Q, double *cs, int n, int ld, int max_sweeps) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int half = n / 2;
  __shared__ double off_norm, diag_norm;
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    // convergence test: off-diagonal Frobenius norm
    __syncthreads();
    if (tid == 0) {
      off_norm = 0.0;
      diag_norm = 0.0;
    }
    __syncthreads();
    double o = 0.0, d = 0.0;
    for (int idx = tid; idx < n * n; idx += nt) {
      int i = idx / n, j = idx % n;
      double v = A[i * ld + j];
      if (i == j) d += v * v;
      else o += v * v;
    }
    o = ppx_warp_sum(o);
    d = ppx_warp_sum(d);
    if ((tid & 31) == 0) {
      atomicAdd(&off_norm, o);  // only used for the stopping test; the rotations themselves are deterministic
      atomicAdd(&diag_norm, d);
    }
    __syncthreads();
    if (off_norm <= 1e-30 * diag_norm) break;
    for (int round = 0; round < n - 1; round++) {
      // round-robin pairing: player n-1 fixed, others rotate
      if (tid < half) {
        int p, q;
        if (tid == 0) {
          p = n - 1;
          q = round % (n - 1);
        } else {
          p = (round + tid) % (n - 1);
          q = (round + n - 1 - tid) % (n - 1);
        }
        if (p > q) {
          int t = p;
          p = q;
          q = t;
        }
        const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
        double c = 1.0, s = 0.0;
        if (fabs(apq) > 1e-300) {
          const double tau = (aqq - app) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t);
          s = t * c;
        }
        cs[4 * tid + 0] = c;
        cs[4 * tid + 1] = s;
        cs[4 * tid + 2] = (double)p;
        cs[4 * tid + 3] = (double)q;
      }
      __syncthreads();
      // columns:  A <- A J,  Q <- Q J
      for (int idx = tid; idx < half * n; idx += nt) {
        const int pr = idx / n, i = idx % n;
        const double c = cs[4 * pr], s = cs[4 * pr + 1];
        const int p = (int)cs[4 * pr + 2], q = (int)cs[4 * pr + 3];
        const double aip = A[i * ld + p], aiq = A[i * ld + q];
        A[i * ld + p] = c * aip - s * aiq;
        A[i * ld + q] = s * aip + c * aiq;
        const double qip = Q[i * ld + p], qiq = Q[i * ld + q];
        Q[i * ld + p] = c * qip - s * qiq;
        Q[i * ld + q] = s * qip + c * qiq;
      }
      __syncthreads();
      // rows:  A <- J^T A
      for (int idx = tid; idx < half * n; idx += nt) {
        const int pr = idx / n, j = idx % n;
        const double c = cs[4 * pr], s = cs[4 * pr + 1];
        const int p = (int)cs[4 * pr + 2], q = (int)cs[4 * pr + 3];
        const double apj = A[p * ld + j], aqj = A[q * ld + j];
        A[p * ld + j] = c * apj - s * aqj;
        A[q * ld + j] = s * apj + c * aqj;
      }
      __syncthreads();
      if (tid < half) {  // the rotated pair is annihilated exactly
        const int p = (int)cs[4 * tid + 2], q = (int)cs[4 * tid + 3];
        A[p * ld + q] = 0.0;
        A[q * ld + p] = 0.0;
      }
      __syncthreads();
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) sym_inverse_jacobi_kernel(const double *__restrict__ S, int R,
                                                                 double *__restrict__ Sinv) {
  extern __shared__ double sm[];
  const int n = (R + 1) & ~1;
  const int ld = n + 1;
  double *A = sm;
  double *Q = sm + n * ld;
  double *cs = Q + n * ld;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int idx = tid; idx < n * n; idx += nt) {
    int i = idx / n, j = idx % n;
    double v = (i < R && j < R) ? 0.5 * (S[i + R * j] + S[j + R * i]) : (i == j ? 1.0 : 0.0);
    A[i * ld + j] = v;
    Q[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  jacobi_eig_shared(A, Q, cs, n, ld, 30);
  // S^-1 = Q diag(1/e) Q^T  (all R eigen-pairs, no truncation: common.cxx:720-722)
  for (int idx = tid; idx < R * R; idx += nt) {
    int i = idx % R, j = idx / R;
    if (i >= j) {
      double v = 0.0;
      for (int k = 0; k < n; k++) {
        if (n != R && k == n - 1 && fabs(Q[(n - 1) * ld + k]) > 0.5) continue;  // the padding eigenvector
        v += Q[i * ld + k] * Q[j * ld + k] / A[k * ld + k];
      }
      Sinv[i + R * j] = v;
      Sinv[j + R * i] = v;
    }
  }
}

// ---- apply: row tiles of 32 rows ----------------------------------------------------------------------------
constexpr int AP_WY = 8;
__global__ void __launch_bounds__(32 * AP_WY) solve_apply_kernel(const double *__restrict__ M,
                                                                 const double *__restrict__ S,
                                                                 const double *__restrict__ Sinv,
                                                                 double *__restrict__ W, int64_t s, int R,
                                                                 const double *__restrict__ W_init,
                                                                 double ratio_step, double *__restrict__ grad_out,
                                                                 double *__restrict__ dW_out) {
  extern __shared__ double sm[];
  double *Ss = sm;              // R*R  (S, column-major as given)
  double *Si = Ss + R * R;      // R*R
  double *Mt = Si + R * R;      // R*32: Mt[r*32 + lane]
  double *Wt = Mt + R * 32;     // R*32: old W
  const int lane = threadIdx.x, wy = threadIdx.y;
  const int tid = wy * 32 + lane, nt = 32 * AP_WY;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  for (int idx = tid; idx < R * R; idx += nt) {
    Ss[idx] = grad_out ? S[idx] : 0.0;
    Si[idx] = Sinv[idx];
  }
  for (int r = wy; r < R; r += AP_WY) {
    Mt[r * 32 + lane] = (i < s) ? M[i + s * r] : 0.0;
    Wt[r * 32 + lane] = (i < s) ? W[i + s * r] : 0.0;
  }
  __syncthreads();
  for (int c = wy; c < R; c += AP_WY) {
    double w = 0.0, g = 0.0;
    for (int r = 0; r < R; r++) {
      w += Mt[r * 32 + lane] * Si[r + R * c];
      g += Wt[r * 32 + lane] * Ss[r + R * c];
    }
    if (i < s) {
      const int64_t o = i + s * c;
      if (grad_out) grad_out[o] = -Mt[c * 32 + lane] + g;
      if (W_init) {
        const double wi = W_init[o];
        const double d = ratio_step * (w - wi);
        if (dW_out) dW_out[o] = d;
        if (ratio_step != 1.0) w = wi + d;
      }
      W[o] = w;
    }
  }
}

// ---- normalize ----------------------------------------------------------------------------------------------
struct NormArgs {
  double *w[16];
  double *g[16];
  int64_t n[16];
  int N;
  int R;
};
__global__ void __launch_bounds__(1024) norm_sq_kernel(NormArgs a, double *sq) {
  __shared__ double red[32];
  const double *x = a.w[blockIdx.x];
  const int64_t n = a.n[blockIdx.x];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i] * x[i];
  s = ppx_block_sum(s, red);
  if (threadIdx.x == 0) sq[blockIdx.x] = s;
}
__global__ void norm_sq_from_gram_kernel(NormArgs a, double *sq) {
  // ||W_i||_F^2 = trace(W_i^T W_i); one warp per factor
  const int i = blockIdx.x;
  double s = 0.0;
  for (int k = threadIdx.x; k < a.R; k += 32) s += a.g[i][k + (int64_t)a.R * k];
  s = ppx_warp_sum(s);
  if (threadIdx.x == 0) sq[i] = s;
}
__global__ void __launch_bounds__(256) norm_scale_kernel(NormArgs a, const double *__restrict__ sq) {
  const int m = blockIdx.y;
  double prod = 1.0;
  for (int j = 0; j < a.N; j++) prod *= sqrt(sq[j]);   // common.cxx:681-683
  const double gm = pow(prod, 1.0 / a.N);               // :684
  const double f = gm / sqrt(sq[m]);                    // :687
  double *x = a.w[m];
  const int64_t n = a.n[m];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = f * x[i];
  if (a.g[m]) {
    double *g = a.g[m];
    const int nn = a.R * a.R;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) g[i] = (f * f) * g[i];
  }
}

}  // namespace

int ppx_sum_partials(ppx_ctx *ctx, const double *partial, int n, double *out);

int ppx_k45_init(ppx_ctx *ctx) {
  const int big = 220 * 1024;
  PPX_CUDA(ctx, cudaFuncSetAttribute(spd_inverse_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(sym_inverse_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(solve_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  return PPX_OK;
}

extern "C" {

int ppx_gram(ppx_ctx *ctx, const double *W, int64_t s, int64_t ldw, int R, double *G) {
  PPX_REQUIRE(ctx, W && G && s >= 0 && R >= 1 && ldw >= s, "W, G non-null; ldw >= s; R >= 1");
  const int npairs = R * (R + 1) / 2;
  const int blocks = (npairs * 32 + 255) / 256;
  gram_kernel<<<blocks, 256, 0, ctx->stream>>>(W, s, ldw, R, G);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_hadamard_grams(ppx_ctx *ctx, const double *const *G, int nG, int skip, int R, double lambda, double *S) {
  PPX_REQUIRE(ctx, G && S && nG >= 1 && nG <= 16, "1 <= nG <= 16");
  HadArgs h;
  h.n = 0;
  for (int j = 0; j < nG; j++)
    if (j != skip) h.g[h.n++] = G[j];
  PPX_REQUIRE(ctx, h.n >= 1, "at least one Gram after skipping");
  hadamard_kernel<<<(R * R + 255) / 256, 256, 0, ctx->stream>>>(h, R, lambda, S);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_spd_inverse(ppx_ctx *ctx, const double *S, int R, int mode, double *Sinv) {
  if (mode == PPX_SOLVE_CHOL) {
    const size_t smem = sizeof(double) * 2 * R * (R + 1);
    if (smem > 220 * 1024) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "solve: R=%d too large for the one-CTA inverse", R);
    spd_inverse_chol_kernel<<<1, 256, smem, ctx->stream>>>(S, R, Sinv);
  } else {
    const int n = (R + 1) & ~1;
    const size_t smem = sizeof(double) * (2 * n * (n + 1) + 4 * (n / 2 + 1));
    if (smem > 220 * 1024) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "solve: R=%d too large for the one-CTA inverse", R);
    sym_inverse_jacobi_kernel<<<1, 256, smem, ctx->stream>>>(S, R, Sinv);
  }
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_solve_update(ppx_ctx *ctx, const double *M, const double *S, double *W, int64_t s, int R,
                     const double *W_init, double ratio_step, int mode, double *grad_out, double *dW_out,
                     double *sq_norms_out) {
  PPX_REQUIRE(ctx, M && S && W && s >= 1 && R >= 1, "M, S, W non-null; s, R >= 1");
  PPX_REQUIRE(ctx, mode == PPX_SOLVE_CHOL || mode == PPX_SOLVE_SVD_PINV, "mode is CHOL or SVD_PINV");
  ppx_ws_reset(ctx);
  double *Sinv = (double *)ppx_ws_alloc(ctx, sizeof(double) * R * R);
  if (!Sinv) return ppx_set_err(ctx, PPX_ENOMEM, "workspace too small");
  int rc = ppx_spd_inverse(ctx, S, R, mode, Sinv);
  if (rc) return rc;
  const size_t smem = sizeof(double) * (2 * (size_t)R * R + 64 * (size_t)R);
  if (smem > 220 * 1024) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "solve: R=%d too large", R);
  solve_apply_kernel<<<ppx_cdiv(s, 32), dim3(32, AP_WY), smem, ctx->stream>>>(M, S, Sinv, W, s, R, W_init, ratio_step,
                                                                             grad_out, dW_out);
  PPX_CHECK_LAUNCH(ctx);
  if (sq_norms_out) {
    const double *xs[3] = {W, dW_out ? dW_out : W, grad_out ? grad_out : W};
    int64_t ns[3] = {s * R, dW_out ? s * R : 0, grad_out ? s * R : 0};
    return ppx_sqnorms(ctx, xs, ns, 3, sq_norms_out);
  }
  return PPX_OK;
}

static int normalize_impl(ppx_ctx *ctx, double *const *W, const int64_t *s, int N, int R, double *const *G,
                          bool from_grams);

int ppx_normalize(ppx_ctx *ctx, double *const *W, const int64_t *s, int N, int R, double *const *G) {
  return normalize_impl(ctx, W, s, N, R, G, false);
}

int ppx_normalize_g(ppx_ctx *ctx, double *const *W, const int64_t *s, int N, int R, double *const *G) {
  PPX_REQUIRE(ctx, G != nullptr, "G != NULL");
  return normalize_impl(ctx, W, s, N, R, G, true);
}

}  // extern "C"

static int normalize_impl(ppx_ctx *ctx, double *const *W, const int64_t *s, int N, int R, double *const *G,
                          bool from_grams) {
  PPX_REQUIRE(ctx, W && s && N >= 1 && N <= 16, "1 <= N <= 16");
  NormArgs a;
  a.N = N;
  a.R = R;
  int64_t nmax = 0;
  for (int i = 0; i < N; i++) {
    a.w[i] = W[i];
    a.g[i] = G ? G[i] : nullptr;
    a.n[i] = s[i] * R;
    if (a.n[i] > nmax) nmax = a.n[i];
  }
  ppx_ws_reset(ctx);
  double *sq = (double *)ppx_ws_alloc(ctx, sizeof(double) * 16);
  if (!sq) return ppx_set_err(ctx, PPX_ENOMEM, "workspace too small");
  if (from_grams)
    norm_sq_from_gram_kernel<<<N, 32, 0, ctx->stream>>>(a, sq);
  else
    norm_sq_kernel<<<N, 1024, 0, ctx->stream>>>(a, sq);
  PPX_CHECK_LAUNCH(ctx);
  int bx = ppx_cdiv(nmax, 256 * 4);
  if (bx < 1) bx = 1;
  if (bx > ctx->sm_count * 4) bx = ctx->sm_count * 4;
  norm_scale_kernel<<<dim3(bx, N), 256, 0, ctx->stream>>>(a, sq);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}


