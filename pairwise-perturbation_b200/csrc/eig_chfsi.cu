// Leading eigenvectors of a large symmetric positive semi-definite matrix by Chebyshev-filtered subspace iteration
// with Rayleigh-Ritz and Hotelling deflation -- the GEMM-shaped alternative to the cooperative Jacobi of tucker_ops.cu
// for the Tucker factor update (MTM.svd(U,S,VT,r), als_Tucker.cxx:20,402,627,868) when r << s.
//
//   X (n x p, p = r + guard) orthonormal; repeat:
//     Rayleigh-Ritz:  H = X^T A X,  H = Z Theta Z^T (one-CTA one-sided Jacobi in shared memory),  X <- X Z
//     residuals  ||A x_j - theta_j x_j||: a leading run of converged Ritz pairs is LOCKED and removed from the matrix,
//       A <- A - theta_j x_j x_j^T  (their eigenvalue becomes 0, i.e. part of the damped interval: no projections, and
//       a dominant outlier -- the mean component of a positive tensor is 10^6 x the rest -- stops polluting the
//       products of the remaining block in floating point);
//     filter:  X_active <- T_d((A - c)/e) X_active,  [0, b] = [0, smallest Ritz value of the block] the damped
//       interval, degree d chosen so that the block's condition number stays below 5e6;
//     orthonormalise X_active by Cholesky QR (twice).
// Every step is a product with the n x n matrix (a small DFMA GEMM with the three-term recurrence fused into its
// epilogue) or p x p work: ~60 products of 80 MFLOP instead of the ~10^4 grid-synchronised rounds of the Jacobi sweep.
// Not converged within the iteration budget (or n too small to pay off) -> the caller falls back to Jacobi.
#include <algorithm>
#include <vector>
#include "ppx_internal.h"

namespace {

// ---- Y = alpha A X + beta X + gamma Z  (A symmetric n x n; X, Z, Y n x pa with leading dimension ld) -----------------
// Two launches: partial products over CG_S slices of K (16 rows x 64 columns per CTA, 128 threads, two rows x four
// columns each; the next 32-deep chunk is fetched into registers while the current one is multiplied), then the sum of
// the slices fused with the three-term recurrence.  (A first version without the K split and the register prefetch
// ran 50 CTAs at one exposed global-load latency per chunk: 129 us per product instead of ~12.)
constexpr int CG_ROWS = 16, CG_COLS = 64, CG_K = 32, CG_S = 4;
__global__ void __launch_bounds__(128) chfsi_gemm_part_kernel(const double *__restrict__ A, int n,
                                                              const double *__restrict__ X, int64_t ld, int pa,
                                                              int kslice, double *__restrict__ P) {
  __shared__ double As[CG_K][CG_ROWS + 1];
  __shared__ double Xs[CG_K][CG_COLS + 1];
  // thread (tx, ty): rows ty, ty + 8; columns tx, tx + 16, tx + 32, tx + 48 (consecutive lanes -> consecutive banks)
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int i0 = blockIdx.x * CG_ROWS, j0 = blockIdx.y * CG_COLS;
  const int kbeg = blockIdx.z * kslice;
  const int kend = min(n, kbeg + kslice);
  double acc[2][4];
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
  double ra[4], rx[16];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 4; u++) {  // A is symmetric: rows i0.. are read as columns (contiguous along k)
      const int idx = tid + 128 * u, kk = idx % CG_K, ii = idx / CG_K;
      const int k = k0 + kk, i = i0 + ii;
      ra[u] = (k < kend && i < n) ? A[k + (int64_t)n * i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int idx = tid + 128 * u, kk = idx % CG_K, jj = idx / CG_K;
      const int k = k0 + kk, j = j0 + jj;
      rx[u] = (k < kend && j < pa) ? X[k + ld * j] : 0.0;
    }
  };
  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += CG_K) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int idx = tid + 128 * u;
      As[idx % CG_K][idx / CG_K] = ra[u];
    }
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int idx = tid + 128 * u;
      Xs[idx % CG_K][idx / CG_K] = rx[u];
    }
    __syncthreads();
    if (k0 + CG_K < kend) fetch(k0 + CG_K);
#pragma unroll 8
    for (int kk = 0; kk < CG_K; kk++) {
      const double a0 = As[kk][ty], a1 = As[kk][ty + 8];
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const double x = Xs[kk][tx + 16 * b];
        acc[0][b] = fma(a0, x, acc[0][b]);
        acc[1][b] = fma(a1, x, acc[1][b]);
      }
    }
    __syncthreads();
  }
  double *o = P + (size_t)blockIdx.z * n * pa;
#pragma unroll
  for (int a = 0; a < 2; a++) {
    const int i = i0 + ty + 8 * a;
    if (i >= n) continue;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int j = j0 + tx + 16 * b;
      if (j < pa) o[i + (size_t)n * j] = acc[a][b];
    }
  }
}
__global__ void __launch_bounds__(256) chfsi_gemm_fin_kernel(const double *__restrict__ P, int n, int pa, int S,
                                                             const double *__restrict__ X, int64_t ld, double alpha,
                                                             double beta, const double *__restrict__ Z, double gamma,
                                                             double *__restrict__ Y) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * pa) return;
  const int i = (int)(idx % n), j = (int)(idx / n);
  double s = 0.0;
  for (int q = 0; q < S; q++) s += P[(size_t)q * n * pa + idx];  // fixed order: deterministic
  double v = alpha * s;
  if (beta != 0.0) v = fma(beta, X[i + ld * j], v);
  if (gamma != 0.0) v = fma(gamma, Z[i + ld * j], v);
  Y[i + ld * j] = v;
}

// H[a, b] = <X[:, a], Y[:, b]>  (pa x pa), one warp per element
__global__ void __launch_bounds__(256) chfsi_cross_gram_kernel(const double *__restrict__ X,
                                                               const double *__restrict__ Y, int n, int64_t ld, int pa,
                                                               double *__restrict__ H) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= pa * pa) return;
  const int a = w % pa, b = w / pa;
  const double *x = X + ld * a, *y = Y + ld * b;
  double s0 = 0.0, s1 = 0.0;
  int i = lane;
  for (; i + 32 < n; i += 64) {
    s0 = fma(x[i], y[i], s0);
    s1 = fma(x[i + 32], y[i + 32], s1);
  }
  if (i < n) s0 = fma(x[i], y[i], s0);
  const double v = ppx_warp_sum(s0 + s1);
  if (lane == 0) H[a + (int64_t)pa * b] = v;
}

// C (n x pa) = X (n x pa) op(Z), Z pa x pa (op = transpose when tz): 32 rows per CTA, Z and the row tile in shared memory
__global__ void __launch_bounds__(256) chfsi_right_mult_kernel(const double *__restrict__ X, int n, int64_t ld, int pa,
                                                               const double *__restrict__ Z, int tz,
                                                               double *__restrict__ C) {
  extern __shared__ double sm[];
  double *Zs = sm;                       // Zs[q * pa + j] = op(Z)[q, j]
  double *Xs = sm + (size_t)pa * pa;     // Xs[q * 33 + r] = X[i0 + r, q]
  const int tid = threadIdx.x, i0 = blockIdx.x * 32;
  for (int idx = tid; idx < pa * pa; idx += 256) {
    const int q = idx % pa, j = idx / pa;  // Z[q + pa*j]
    if (tz) Zs[j * pa + q] = Z[idx];       // op(Z)[j, q] = Z[q, j]
    else Zs[q * pa + j] = Z[idx];
  }
  for (int idx = tid; idx < 32 * pa; idx += 256) {
    const int r = idx % 32, q = idx / 32;
    Xs[q * 33 + r] = (i0 + r < n) ? X[i0 + r + ld * q] : 0.0;
  }
  __syncthreads();
  const int r = tid % 32;
  if (i0 + r >= n) return;
  for (int j = tid / 32; j < pa; j += 8) {
    double s0 = 0.0, s1 = 0.0;
    int q = 0;
    for (; q + 1 < pa; q += 2) {
      s0 = fma(Xs[q * 33 + r], Zs[q * pa + j], s0);
      s1 = fma(Xs[(q + 1) * 33 + r], Zs[(q + 1) * pa + j], s1);
    }
    if (q < pa) s0 = fma(Xs[q * 33 + r], Zs[q * pa + j], s0);
    C[i0 + r + ld * j] = s0 + s1;
  }
}

// A -= sum_j theta[j] x_j x_j^T over k newly locked Ritz pairs (Hotelling deflation)
__global__ void __launch_bounds__(256) chfsi_deflate_kernel(double *__restrict__ A, int n,
                                                            const double *__restrict__ X, int64_t ld,
                                                            const double *__restrict__ theta, int k) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * n) return;
  const int i = (int)(idx % n), j = (int)(idx / n);
  double s = 0.0;
  for (int q = 0; q < k; q++) s = fma(theta[q] * X[i + ld * q], X[j + ld * q], s);
  A[idx] -= s;
}

// res[j] = || AX_j - theta_j X_j ||, one CTA per column
__global__ void __launch_bounds__(256) chfsi_resid_kernel(const double *__restrict__ AX, const double *__restrict__ X,
                                                          int n, int64_t ld, const double *__restrict__ theta,
                                                          double *__restrict__ res) {
  __shared__ double red[32];
  const int j = blockIdx.x;
  const double th = theta[j];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = AX[i + ld * j] - th * X[i + ld * j];
    s = fma(d, d, s);
  }
  s = ppx_block_sum(s, red);
  if (threadIdx.x == 0) res[j] = sqrt(s);
}

// reciprocal square root / reciprocal from a float seed and three Newton steps (full double precision for arguments
// inside the float range, IEEE fallback outside): the rotation parameters are the dependent latency of a Jacobi round
__device__ __forceinline__ double fast_rsqrt(double w) {
  const float f = (float)w;
  if (!(f > 1e-30f && f < 1e30f)) return rsqrt(w);
  double y = (double)rsqrtf(f);
  const double hw = 0.5 * w;
  y = y * fma(-hw * y, y, 1.5);
  y = y * fma(-hw * y, y, 1.5);
  y = y * fma(-hw * y, y, 1.5);
  return y;
}
__device__ __forceinline__ double fast_rcp(double d) {
  const float f = (float)d;
  if (!(fabsf(f) > 1e-30f && fabsf(f) < 1e30f)) return 1.0 / d;
  double x = (double)__fdividef(1.0f, f);
  x = fma(x, fma(-d, x, 1.0), x);
  x = fma(x, fma(-d, x, 1.0), x);
  x = fma(x, fma(-d, x, 1.0), x);
  return x;
}

// ---- one-sided Jacobi of a p x p symmetric PSD matrix in shared memory, one CTA of 32 warps ----------------------------
// B = 0.5 (H + H^T); plane rotations of column pairs until a sweep applies none above sqrt(p) eps (or the largest was
// below 1e-8); column j converges to lambda_j z_j.  Output: Z (p x p, unit columns, DECREASING eigenvalue) and theta.
__global__ void __launch_bounds__(1024) chfsi_jacobi_smem_kernel(const double *__restrict__ H, int p,
                                                                 double *__restrict__ Zout,
                                                                 double *__restrict__ theta_out) {
  extern __shared__ double sm[];
  const int ld = p + 1;
  double *B = sm;                    // [p][ld], column c at B + c*ld
  double *nrm = sm + (size_t)p * ld; // [p]
  __shared__ int nrot_w[32];       // per warp, written once per sweep (shared atomics from 32 warps per round cost
  __shared__ double maxrot_w[32];  // more than the rotations themselves)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int idx = tid; idx < p * p; idx += blockDim.x) {
    const int i = idx % p, j = idx / p;
    B[j * ld + i] = 0.5 * (H[i + (int64_t)p * j] + H[j + (int64_t)p * i]);
  }
  const int m = (p + 1) & ~1, half = m / 2;
  const double tol = 2.3e-16 * sqrt((double)p);
  __syncthreads();
  for (int sweep = 0; sweep < 40; sweep++) {
    int my_rot = 0;
    double my_max = 0.0;
    // squared column norms, exact at the start of every sweep; inside the sweep a rotation updates the two it touches
    // (|a'|^2 = |a|^2 - t a.b, |b'|^2 = |b|^2 + t a.b), so a pair costs ONE inner product and one warp reduction
    // instead of three -- the single SM this kernel runs on is bound by its FP64 rate
    for (int c = warp; c < p; c += nw) {
      double q = 0.0;
      for (int i = lane; i < p; i += 32) q = fma(B[c * ld + i], B[c * ld + i], q);
      q = ppx_warp_sum(q);
      if (lane == 0) nrm[c] = q;
    }
    __syncthreads();
    for (int round = 0; round < m - 1; round++) {
      for (int pr = warp; pr < half; pr += nw) {
        int a, b;
        if (pr == 0) {
          a = m - 1;
          b = round % (m - 1);
        } else {
          a = (round + pr) % (m - 1);
          b = (round + m - 1 - pr) % (m - 1);
        }
        if (a > b) {
          const int t = a;
          a = b;
          b = t;
        }
        if (b >= p) continue;
        double *ca = B + a * ld, *cb = B + b * ld;
        double xa[4], xb[4];  // p <= 128
        double sab = 0.0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = lane + 32 * u;
          xa[u] = i < p ? ca[i] : 0.0;
          xb[u] = i < p ? cb[i] : 0.0;
          sab = fma(xa[u], xb[u], sab);
        }
        const double saa = nrm[a], sbb = nrm[b];
        sab = ppx_warp_sum(sab);
        if (!(fabs(sab) > tol * sqrt(saa * sbb) && fabs(sab) > 1e-300)) continue;
        // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b - a) / (2 g), written with one sqrt, one division and
        // one rsqrt (these are the dependent latencies of the round): t = 2 g / (h + sign(h) sqrt(h^2 + 4 g^2))
        const double hh = sbb - saa, g2 = 2.0 * sab;
        const double ww = fma(hh, hh, g2 * g2);
        const double rad = ww * fast_rsqrt(ww);
        const double t = g2 * fast_rcp(hh >= 0.0 ? hh + rad : hh - rad);
        const double c = fast_rsqrt(fma(t, t, 1.0)), s = c * t;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = lane + 32 * u;
          if (i < p) {
            ca[i] = c * xa[u] - s * xb[u];
            cb[i] = s * xa[u] + c * xb[u];
          }
        }
        if (lane == 0) {
          const double na = saa - t * sab, nb = sbb + t * sab;
          nrm[a] = na > 0.0 ? na : 0.0;
          nrm[b] = nb > 0.0 ? nb : 0.0;
        }
        my_rot++;
        my_max = fmax(my_max, fabs(s));
      }
      __syncthreads();
    }
    if (lane == 0) {
      nrot_w[warp] = my_rot;
      maxrot_w[warp] = my_max;
    }
    __syncthreads();
    int nrot = 0;
    double maxrot = 0.0;
    for (int w = 0; w < nw; w++) {
      nrot += nrot_w[w];
      maxrot = fmax(maxrot, maxrot_w[w]);
    }
    __syncthreads();
    if (nrot == 0 || maxrot < 1e-8) break;
  }
  // column norms = eigenvalues
  for (int c = warp; c < p; c += nw) {
    double s = 0.0;
    for (int i = lane; i < p; i += 32) s = fma(B[c * ld + i], B[c * ld + i], s);
    s = ppx_warp_sum(s);
    if (lane == 0) nrm[c] = sqrt(s);
  }
  __syncthreads();
  // rank by counting (ties broken by index), write normalised columns in decreasing order
  for (int c = warp; c < p; c += nw) {
    const double lc = nrm[c];
    int cnt = 0;
    for (int j = lane; j < p; j += 32) {
      const double lj = nrm[j];
      if (j != c && (lj > lc || (lj == lc && j < c))) cnt++;
    }
    cnt = (int)ppx_warp_sum((double)cnt);  // exact: small integers
    const double inv = lc > 0.0 ? 1.0 / lc : 0.0;
    for (int i = lane; i < p; i += 32) Zout[i + (int64_t)p * cnt] = lc > 0.0 ? B[c * ld + i] * inv : (i == c ? 1.0 : 0.0);
    if (lane == 0) theta_out[cnt] = lc;
  }
}

// U[:, k] = sign * X[:, src[k]], largest-magnitude component positive (the convention of eig_select_kernel)
__global__ void __launch_bounds__(256) chfsi_extract_kernel(const double *__restrict__ X, int n, int64_t ld,
                                                            const int *__restrict__ src, double *__restrict__ U) {
  __shared__ double red_v[256];
  __shared__ int red_i[256];
  const int k = blockIdx.x;
  const double *x = X + ld * src[k];
  double best = -1.0;
  int bi = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double v = fabs(x[j]);
    if (v > best) {
      best = v;
      bi = j;
    }
  }
  red_v[threadIdx.x] = best;
  red_i[threadIdx.x] = bi;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      if (red_v[threadIdx.x + st] > red_v[threadIdx.x] ||
          (red_v[threadIdx.x + st] == red_v[threadIdx.x] && red_i[threadIdx.x + st] < red_i[threadIdx.x])) {
        red_v[threadIdx.x] = red_v[threadIdx.x + st];
        red_i[threadIdx.x] = red_i[threadIdx.x + st];
      }
    }
    __syncthreads();
  }
  const double sgn = x[red_i[0]] < 0.0 ? -1.0 : 1.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) U[j + (int64_t)n * k] = sgn * x[j];
}

// P: scratch of CG_S * n * pa doubles
int gemm_cheb(ppx_ctx *ctx, const double *A, int n, const double *X, int64_t ld, int pa, double alpha, double beta,
              const double *Z, double gamma, double *Y, double *P) {
  const int kslice = ppx_cdiv(ppx_cdiv(n, CG_S), CG_K) * CG_K;
  const int S = ppx_cdiv(n, kslice);
  dim3 grid(ppx_cdiv(n, CG_ROWS), ppx_cdiv(pa, CG_COLS), S);
  chfsi_gemm_part_kernel<<<grid, 128, 0, ctx->stream>>>(A, n, X, ld, pa, kslice, P);
  PPX_CHECK_LAUNCH(ctx);
  chfsi_gemm_fin_kernel<<<ppx_cdiv((int64_t)n * pa, 256), 256, 0, ctx->stream>>>(P, n, pa, S, X, ld, alpha, beta, Z, gamma, Y);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int right_mult(ppx_ctx *ctx, const double *X, int n, int64_t ld, int pa, const double *Z, int tz, double *C) {
  static bool attr_set = false;
  if (!attr_set) {
    PPX_CUDA(ctx, cudaFuncSetAttribute(chfsi_right_mult_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  const size_t smem = sizeof(double) * ((size_t)pa * pa + 33 * (size_t)pa);
  chfsi_right_mult_kernel<<<ppx_cdiv(n, 32), 256, smem, ctx->stream>>>(X, n, ld, pa, Z, tz, C);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

#define CHK(x)            \
  do {                    \
    const int rc_ = (x);  \
    if (rc_) return rc_;  \
  } while (0)

// X (n x pa, leading dimension ld) <- orthonormal basis of its columns by Cholesky QR, twice; T is scratch of the same
// shape, G / Zi are pa x pa.  The result ends up in X.
int cholqr2(ppx_ctx *ctx, double *X, double *T, int n, int64_t ld, int pa, double *G, double *Zi) {
  for (int pass = 0; pass < 2; pass++) {
    CHK(ppx_gram(ctx, X, n, ld, pa, G));
    CHK(ppx_spd_factor_inverse(ctx, G, pa, Zi));
    CHK(right_mult(ctx, X, n, ld, pa, Zi, 1, T));
    PPX_CUDA(ctx, cudaMemcpyAsync(X, T, sizeof(double) * (size_t)ld * pa, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return PPX_OK;
}

}  // namespace

bool ppx_eig_chfsi_applicable(int64_t n, int r) {
  if (getenv("PPX_EIG_JACOBI")) return false;
  // p = r + guard columns (guard 24, down to 8 for large r): the Cholesky-QR factor kernel takes p <= 112
  return n >= 384 && r >= 1 && r + 8 <= 112 && 4 * (int64_t)r <= n;
}

// Returns PPX_OK with U / evals_out filled, 1 if the iteration did not converge (nothing usable was written: the caller
// falls back to the Jacobi solver), a negative code on errors.  `A` (n x n, symmetric) is destroyed.  `state`
// (>= n*(r+24) + 1 doubles, optional) holds the final block and a tag for the next call on a nearby matrix (same n and r).
int ppx_eig_chfsi(ppx_ctx *ctx, double *A, int n, int r, double *U, double *evals_out, double *state, int state_valid) {
  const int p = r + (112 - r < 24 ? 112 - r : 24);
  const int64_t ld = n;
  const size_t np = (size_t)n * p;
  double *X = (double *)ppx_ws_alloc(ctx, sizeof(double) * np);
  double *Y0 = (double *)ppx_ws_alloc(ctx, sizeof(double) * np);
  double *Y1 = (double *)ppx_ws_alloc(ctx, sizeof(double) * np);
  double *AX = (double *)ppx_ws_alloc(ctx, sizeof(double) * np);
  double *T = (double *)ppx_ws_alloc(ctx, sizeof(double) * np);
  double *H = (double *)ppx_ws_alloc(ctx, sizeof(double) * p * p);
  double *Zr = (double *)ppx_ws_alloc(ctx, sizeof(double) * p * p);
  double *G = (double *)ppx_ws_alloc(ctx, sizeof(double) * p * p);
  double *Zi = (double *)ppx_ws_alloc(ctx, sizeof(double) * p * p);
  double *theta = (double *)ppx_ws_alloc(ctx, sizeof(double) * 2 * p);
  int *src = (int *)ppx_ws_alloc(ctx, sizeof(int) * p);
  double *P = (double *)ppx_ws_alloc(ctx, sizeof(double) * CG_S * np);
  if (!X || !Y0 || !Y1 || !AX || !T || !H || !Zr || !G || !Zi || !theta || !src || !P) return 1;
  double *res = theta + p;
  const bool verbose = getenv("PPX_EIG_VERBOSE") != nullptr;

  // start block: the previous call's block, or random.  `state` is shared with the Jacobi solver (which stores an
  // n x n basis there when this iteration gave up): a block left by THIS solver carries a tag behind its last column
  // (np < n*n because 4 r <= n); anything else is not a block and a cold start is taken.  (Reading a Jacobi basis as
  // a block started one mode of a 2-GPU HOOI run orthogonal to its dominant eigenvector: the filter then diverged,
  // the fallback ran again, and the mode stayed on the 15 ms solver for the rest of the run.)
  const double kStateTag = 0x1.c4f51b5ea7ap+77 * (double)p + (double)n;
  if (state && state_valid) {
    double tag = 0.0;
    PPX_CUDA(ctx, cudaMemcpyAsync(&tag, state + np, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (tag != kStateTag) state_valid = 0;
  }
  if (state && state_valid) {
    PPX_CUDA(ctx, cudaMemcpyAsync(X, state, sizeof(double) * np, cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    CHK(ppx_fill_uniform(ctx, X, (int64_t)np, 0x5eedULL, 77, 0, -0.5, 0.5));
  }
  CHK(cholqr2(ctx, X, T, n, ld, p, G, Zi));

  std::vector<double> h(2 * (size_t)p), lam_locked;
  std::vector<int> order;
  int nlock = 0;  // columns [0, nlock) of X are locked (final), in the order they were locked
  double lam_max = -1.0;  // largest Ritz value of the first Rayleigh-Ritz step: the scale "numerically zero" refers to
  const size_t jac_smem = sizeof(double) * ((size_t)p * (p + 1) + p);
  static bool attr_set = false;
  if (!attr_set) {
    PPX_CUDA(ctx, cudaFuncSetAttribute(chfsi_jacobi_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  int iters = 0, products = 0;
  // Stagnation: the attainable residual sits near 1e-12 theta_j (deflation with a locked vector that is itself only
  // converged to that level, orthonormality of the block after Cholesky QR); a pair can stall a few per cent above the
  // threshold for ever (seen at 1.06e-12 on the all-reduced Gram of a 2-GPU run).  When the first open pair has not
  // improved by a factor 2 since the previous step with nothing newly locked, 1e-10 theta_j is accepted: still two
  // orders below the eps ||A|| / gap accuracy of a dense eigensolver on these matrices.
  double prev_rel = -1.0;
  int prev_nlock = -1;
  for (int outer = 0; outer < 24; outer++, iters++) {
    const int pa = p - nlock;
    double *Xa = X + ld * nlock;
    // Rayleigh-Ritz on the active block
    CHK(gemm_cheb(ctx, A, n, Xa, ld, pa, 1.0, 0.0, nullptr, 0.0, AX, P));
    products++;
    chfsi_cross_gram_kernel<<<ppx_cdiv((int64_t)pa * pa * 32, 256), 256, 0, ctx->stream>>>(Xa, AX, n, ld, pa, H);
    PPX_CHECK_LAUNCH(ctx);
    // (a two-sided Jacobi with H itself in shared memory -- rotation from three entries, no reductions -- was measured
    // SLOWER, 0.71 vs 0.60 ms per step at p = 64: on these graded Ritz matrices, one outlier 1e6 x the bulk, it needs
    // many more sweeps to push the small off-diagonal entries below the relative threshold)
    chfsi_jacobi_smem_kernel<<<1, 1024, jac_smem, ctx->stream>>>(H, pa, Zr, theta);
    PPX_CHECK_LAUNCH(ctx);
    CHK(right_mult(ctx, Xa, n, ld, pa, Zr, 0, T));
    PPX_CUDA(ctx, cudaMemcpyAsync(Xa, T, sizeof(double) * (size_t)ld * pa, cudaMemcpyDeviceToDevice, ctx->stream));
    CHK(right_mult(ctx, AX, n, ld, pa, Zr, 0, T));
    chfsi_resid_kernel<<<pa, 256, 0, ctx->stream>>>(T, Xa, n, ld, theta, res);
    PPX_CHECK_LAUNCH(ctx);
    PPX_CUDA(ctx, cudaMemcpyAsync(h.data(), theta, sizeof(double) * 2 * p, cudaMemcpyDeviceToHost, ctx->stream));
    PPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double *th = h.data(), *rs = h.data() + p;
    for (int j = 0; j < pa; j++)
      if (!(th[j] == th[j]) || !(rs[j] == rs[j])) return 1;  // NaN: the block lost rank somewhere
    if (lam_max < 0.0) lam_max = th[0];
    // lock the leading run of converged Ritz pairs (Ritz values are in decreasing order); below 1e-13 of the largest
    // eigenvalue the matrix is numerically rank deficient and any orthonormal vectors of the block serve
    const double rel0 = th[0] > 0.0 ? rs[0] / th[0] : 0.0;
    const bool stagnated = outer >= 2 && prev_nlock == nlock && prev_rel > 0.0 && rel0 > 0.5 * prev_rel && rel0 <= 1e-10;
    const double tolrel = stagnated ? 1e-10 : 1e-12;
    int nconv = 0;
    while (nconv < pa && nlock + nconv < r &&
           (rs[nconv] <= tolrel * th[nconv] + 1e-300 || th[nconv] <= 1e-13 * lam_max))
      nconv++;
    prev_rel = rel0;
    prev_nlock = nlock;
    if (verbose)
      fprintf(stderr, "chfsi n=%d r=%d it %d: locked %d (+%d), theta[0]=%.6e theta[last]=%.6e res[first unconv]=%.3e\n", n,
              r, outer, nlock, nconv, th[0], th[pa - 1], nconv < pa ? rs[nconv] : 0.0);
    if (nconv > 0) {
      for (int j = 0; j < nconv; j++) lam_locked.push_back(th[j]);
      chfsi_deflate_kernel<<<ppx_cdiv((int64_t)n * n, 256), 256, 0, ctx->stream>>>(A, n, Xa, ld, theta, nconv);
      PPX_CHECK_LAUNCH(ctx);
      nlock += nconv;
    }
    if (nlock >= r) break;
    if (outer == 23) return 1;
    // Chebyshev filter of the remaining block on the deflated matrix: damp [0, b]
    const int pa2 = p - nlock;
    double *Xb = X + ld * nlock;
    const double top = th[nconv];               // largest Ritz value still wanted
    double b = th[pa - 1];                       // smallest Ritz value of the block: upper end of the unwanted part
    if (!(b > 1e-8 * top)) b = 1e-8 * top;       // (rank-deficient matrices: everything below is zero)
    if (!(top > 0.0)) return 1;
    const double c = 0.5 * b, e = 0.5 * b;
    const double xmax = (top - c) / e;
    int d = 24;
    if (xmax > 1.0) {
      const double per = log(xmax + sqrt(xmax * xmax - 1.0));  // acosh
      const int cap = (int)(log(5e6) / per);  // Cholesky QR twice copes with a block condition number up to ~1e7
      if (cap < d) d = cap;
    }
    if (d < 2) d = 2;
    // T_0 = X, T_1 = (A - c) X / e, T_{k+1} = 2 (A - c) T_k / e - T_{k-1}
    // The Rayleigh-Ritz step (a p x p eigenproblem) costs as much as ~30 products: when the slowest wanted pair is
    // still far from converged, run several filter + orthonormalise passes before the next one.  A pass improves it by
    // about T_d(x_w), x_w the position of the smallest wanted Ritz value.
    int passes = 1;
    {
      const int rw = r - nlock;                        // wanted pairs still open
      const int jw = nconv + (rw < pa2 ? rw : pa2) - 1;  // slowest of them, in the pre-lock numbering
      const double xw = (th[jw] - c) / e;
      double worst = 0.0;
      for (int j = nconv; j <= jw; j++) worst = fmax(worst, rs[j] / th[j]);
      if (xw > 1.0 && worst > 1e-12) {
        const double gain = d * log(xw + sqrt(xw * xw - 1.0));  // log of T_d(x_w)
        if (gain > 0.1) passes = (int)ceil(log(worst / 1e-12) / gain);
      }
      if (passes < 1) passes = 1;
      if (passes > 3) passes = 3;
    }
    for (int pass = 0; pass < passes; pass++) {
      // three buffers rotate: the one holding T_{k-1} is free once T_{k+1} has been formed
      double *Tprev = Xb, *Tcur = Y0, *Tnext = Y1;
      CHK(gemm_cheb(ctx, A, n, Tprev, ld, pa2, 1.0 / e, -c / e, nullptr, 0.0, Tcur, P));
      for (int k = 2; k <= d; k++) {
        CHK(gemm_cheb(ctx, A, n, Tcur, ld, pa2, 2.0 / e, -2.0 * c / e, Tprev, -1.0, Tnext, P));
        double *freed = Tprev;
        Tprev = Tcur;
        Tcur = Tnext;
        Tnext = freed;
      }
      products += d;
      if (Tcur != Xb)
        PPX_CUDA(ctx, cudaMemcpyAsync(Xb, Tcur, sizeof(double) * (size_t)ld * pa2, cudaMemcpyDeviceToDevice, ctx->stream));
      CHK(cholqr2(ctx, Xb, T, n, ld, pa2, G, Zi));
    }
  }
  if (verbose) fprintf(stderr, "chfsi n=%d r=%d: converged after %d iterations, %d products with A\n", n, r, iters + 1, products);
  // order the locked vectors by decreasing eigenvalue and write U
  order.resize(r);
  for (int k = 0; k < r; k++) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lam_locked[a] > lam_locked[b]; });
  PPX_CUDA(ctx, cudaMemcpyAsync(src, order.data(), sizeof(int) * r, cudaMemcpyHostToDevice, ctx->stream));
  chfsi_extract_kernel<<<r, 256, 0, ctx->stream>>>(X, n, ld, src, U);
  PPX_CHECK_LAUNCH(ctx);
  if (evals_out) {
    std::vector<double> ev(r);
    for (int k = 0; k < r; k++) ev[k] = lam_locked[order[k]];
    PPX_CUDA(ctx, cudaMemcpyAsync(evals_out, ev.data(), sizeof(double) * r, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (state) {
    PPX_CUDA(ctx, cudaMemcpyAsync(state, X, sizeof(double) * np, cudaMemcpyDeviceToDevice, ctx->stream));
    PPX_CUDA(ctx, cudaMemcpyAsync(state + np, &kStateTag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  PPX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `order`, `ev` are host temporaries of the copies above
  return PPX_OK;
}
