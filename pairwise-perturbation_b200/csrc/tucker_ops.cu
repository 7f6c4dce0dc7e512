// Tucker-side operators (K9): Gram of an unfolding, top-r symmetric eigenvectors, sign alignment.
//   ppx_unfold_gram   MTM = T_(i) T_(i)^T                       (common.cxx:205-223)
//   ppx_sym_eig_topk  leading r eigenvectors of a symmetric PSD (what MTM.svd(U,S,VT,r) yields; als_Tucker.cxx:20,402)
//   ppx_sign_align    U <- U diag(sign(diag(U^T Uref)))         (als_Tucker.cxx:632-643)
// The TTM chain itself (K8/K10) is the DMMA GEMM of k1_ttm_first.cu with the rank written in place of mode x.
#include <cooperative_groups.h>
#include "ppx_internal.h"

namespace cg = cooperative_groups;

namespace {

// ---- SYRK on an unfolding: 64x64 output tile per CTA, 16-deep chunks, 4x4 per thread, split over z --------------
constexpr int UG_T = 64, UG_K = 16;
__global__ void __launch_bounds__(256) unfold_gram_kernel(const double *__restrict__ T, int64_t L, int64_t X,
                                                          int64_t Rt, int64_t C, int64_t c_per_z,
                                                          double *__restrict__ out_z) {
  __shared__ double Ap[UG_K][UG_T + 4];
  __shared__ double Aq[UG_K][UG_T + 4];
  const int tid = threadIdx.x;
  const int64_t p0 = (int64_t)blockIdx.x * UG_T, q0 = (int64_t)blockIdx.y * UG_T;
  if (q0 > p0) return;  // lower triangle of tiles only; mirrored by the reduction kernel
  const int64_t cb = (int64_t)blockIdx.z * c_per_z;
  int64_t ce = cb + c_per_z;
  if (ce > C) ce = C;
  const int tx = tid % 16, ty = tid / 16;  // thread computes rows p0+4*tx.., cols q0+4*ty..
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
  for (int64_t c0 = cb; c0 < ce; c0 += UG_K) {
    // load UG_K x UG_T elements of each panel; element (c, p) lives at l + L*(p + X*t), c = l + L*t
    for (int idx = tid; idx < UG_K * UG_T; idx += 256) {
      int cc, pp;
      if (L > 1) {
        cc = idx % UG_K;
        pp = idx / UG_K;
      } else {
        pp = idx % UG_T;
        cc = idx / UG_T;
      }
      const int64_t c = c0 + cc;
      double vp = 0.0, vq = 0.0;
      if (c < ce) {
        const int64_t t = c / L, l = c - t * L;
        const int64_t base = l + L * X * t;
        if (p0 + pp < X) vp = T[base + L * (p0 + pp)];
        if (q0 + pp < X) vq = T[base + L * (q0 + pp)];
      }
      Ap[cc][pp] = vp;
      Aq[cc][pp] = vq;
    }
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < UG_K; cc++) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = Ap[cc][4 * tx + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Aq[cc][4 * ty + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  double *o = out_z + (int64_t)blockIdx.z * X * X;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t p = p0 + 4 * tx + i, q = q0 + 4 * ty + j;
      if (p < X && q < X) o[p + X * q] = acc[i][j];
    }
}

__global__ void unfold_gram_reduce_kernel(const double *__restrict__ parts, int64_t X, int nz,
                                          double *__restrict__ MTM) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= X * X) return;
  int64_t p = idx % X, q = idx / X;
  // tiles with q0 > p0 were skipped: read the mirrored element
  int64_t pp = p, qq = q;
  if ((q / UG_T) > (p / UG_T)) {
    pp = q;
    qq = p;
  }
  double s = 0.0;
  for (int z = 0; z < nz; z++) s += parts[(int64_t)z * X * X + pp + X * qq];
  MTM[idx] = s;
}

// ---- cooperative two-sided Jacobi in global memory (matrix is L2 resident) ----------------------------------------
// A (n x n, symmetric, destroyed), Q (n x n) column-major, cs: 4*(n/2) doubles, flag: 2 doubles (off, diag norms)
__global__ void __launch_bounds__(256) jacobi_global_kernel(double *__restrict__ A, double *__restrict__ Q,
                                                            double *__restrict__ cs, double *__restrict__ norms,
                                                            int n, int max_sweeps) {
  cg::grid_group grid = cg::this_grid();
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
  const int half = n / 2;
  for (int64_t idx = gtid; idx < (int64_t)n * n; idx += gsz) Q[idx] = (idx % n == idx / n) ? 1.0 : 0.0;
  grid.sync();
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    if (gtid == 0) {
      norms[0] = 0.0;
      norms[1] = 0.0;
    }
    grid.sync();
    {
      double o = 0.0, d = 0.0;
      for (int64_t idx = gtid; idx < (int64_t)n * n; idx += gsz) {
        const double v = A[idx];
        if (idx % n == idx / n) d += v * v;
        else o += v * v;
      }
      o = ppx_warp_sum(o);
      d = ppx_warp_sum(d);
      if ((threadIdx.x & 31) == 0) {  // stopping test only
        atomicAdd(&norms[0], o);
        atomicAdd(&norms[1], d);
      }
    }
    grid.sync();
    if (norms[0] <= 1e-30 * norms[1]) break;
    for (int round = 0; round < n - 1; round++) {
      for (int64_t pr = gtid; pr < half; pr += gsz) {
        int p, q;
        if (pr == 0) {
          p = n - 1;
          q = round % (n - 1);
        } else {
          p = (int)((round + pr) % (n - 1));
          q = (int)((round + n - 1 - pr) % (n - 1));
        }
        if (p > q) {
          int t = p;
          p = q;
          q = t;
        }
        const double app = A[p + (int64_t)n * p], aqq = A[q + (int64_t)n * q], apq = A[p + (int64_t)n * q];
        double c = 1.0, s = 0.0;
        if (fabs(apq) > 1e-300) {
          const double tau = (aqq - app) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t);
          s = t * c;
        }
        cs[4 * pr + 0] = c;
        cs[4 * pr + 1] = s;
        cs[4 * pr + 2] = (double)p;
        cs[4 * pr + 3] = (double)q;
      }
      grid.sync();
      // columns p,q of A and Q (contiguous along i)
      for (int64_t idx = gtid; idx < (int64_t)half * n; idx += gsz) {
        const int pr = (int)(idx / n), i = (int)(idx % n);
        const double c = cs[4 * pr], s = cs[4 * pr + 1];
        const int64_t p = (int64_t)cs[4 * pr + 2], q = (int64_t)cs[4 * pr + 3];
        const double aip = A[i + n * p], aiq = A[i + n * q];
        A[i + n * p] = c * aip - s * aiq;
        A[i + n * q] = s * aip + c * aiq;
        const double qip = Q[i + n * p], qiq = Q[i + n * q];
        Q[i + n * p] = c * qip - s * qiq;
        Q[i + n * q] = s * qip + c * qiq;
      }
      grid.sync();
      // rows p,q of A
      for (int64_t idx = gtid; idx < (int64_t)half * n; idx += gsz) {
        const int pr = (int)(idx % half), j = (int)(idx / half);
        const double c = cs[4 * pr], s = cs[4 * pr + 1];
        const int64_t p = (int64_t)cs[4 * pr + 2], q = (int64_t)cs[4 * pr + 3];
        const double apj = A[p + (int64_t)n * j], aqj = A[q + (int64_t)n * j];
        A[p + (int64_t)n * j] = c * apj - s * aqj;
        A[q + (int64_t)n * j] = s * apj + c * aqj;
      }
      grid.sync();
      for (int64_t pr = gtid; pr < half; pr += gsz) {
        const int64_t p = (int64_t)cs[4 * pr + 2], q = (int64_t)cs[4 * pr + 3];
        A[p + n * q] = 0.0;
        A[q + n * p] = 0.0;
      }
      // no sync needed here: the next round's parameter phase reads A[p][p], A[q][q], A[p][q] of OTHER pairs only
      // after this loop?  No -- pairs change between rounds, so order it:
      grid.sync();
    }
  }
}

// pick the r largest eigenvalues (diag of A), write eigenvectors in decreasing order; pad index (if any) excluded
__global__ void __launch_bounds__(256) eig_select_kernel(const double *__restrict__ A, const double *__restrict__ Q,
                                                         int n, int s, int r, double *__restrict__ U,
                                                         double *__restrict__ evals) {
  // one block per candidate eigen-index i < n
  const int i = blockIdx.x;
  __shared__ int rank_s;
  __shared__ int is_pad;
  if (threadIdx.x == 0) {
    rank_s = 0;
    is_pad = (n != s) && (fabs(Q[(int64_t)(n - 1) + (int64_t)n * i]) > 0.5);
  }
  __syncthreads();
  const double li = A[i + (int64_t)n * i];
  int cnt = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    if (j == i) continue;
    const bool jpad = (n != s) && (fabs(Q[(int64_t)(n - 1) + (int64_t)n * j]) > 0.5);
    if (jpad) continue;
    const double lj = A[j + (int64_t)n * j];
    if (lj > li || (lj == li && j < i)) cnt++;
  }
  atomicAdd(&rank_s, cnt);  // integer: order independent
  __syncthreads();
  if (is_pad || rank_s >= r) return;
  const int k = rank_s;
  // deterministic sign: make the largest-magnitude component positive
  __shared__ double red_v[256];
  __shared__ int red_i[256];
  double best = -1.0;
  int bi = 0;
  for (int j = threadIdx.x; j < s; j += blockDim.x) {
    const double v = fabs(Q[j + (int64_t)n * i]);
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
  const double sgn = Q[red_i[0] + (int64_t)n * i] < 0.0 ? -1.0 : 1.0;
  for (int j = threadIdx.x; j < s; j += blockDim.x) U[j + (int64_t)s * k] = sgn * Q[j + (int64_t)n * i];
  if (threadIdx.x == 0 && evals) evals[k] = li;
}

__global__ void pad_copy_kernel(const double *__restrict__ M, int s, int n, double *__restrict__ A) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * n) return;
  const int i = (int)(idx % n), j = (int)(idx / n);
  A[idx] = (i < s && j < s) ? 0.5 * (M[i + (int64_t)s * j] + M[j + (int64_t)s * i]) : (i == j ? 1.0 : 0.0);
}

__global__ void __launch_bounds__(256) sign_align_kernel(double *__restrict__ U, const double *__restrict__ Uref,
                                                         int64_t s) {
  __shared__ double red[32];
  double *u = U + s * blockIdx.x;
  const double *v = Uref + s * blockIdx.x;
  double d = 0.0;
  for (int64_t j = threadIdx.x; j < s; j += blockDim.x) d += u[j] * v[j];
  d = ppx_block_sum(d, red);
  __shared__ double sg;
  if (threadIdx.x == 0) sg = (d > 0.0) ? 1.0 : -1.0;
  __syncthreads();
  if (sg < 0.0)
    for (int64_t j = threadIdx.x; j < s; j += blockDim.x) u[j] = -u[j];
}

}  // namespace

extern "C" {

int ppx_unfold_gram(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int i, double *MTM) {
  PPX_REQUIRE(ctx, T && lens && MTM && k >= 1 && k <= 16 && i >= 0 && i < k, "bad arguments");
  int64_t L, X, Rt;
  ppx_split3(lens, k, i, &L, &X, &Rt);
  const int64_t C = L * Rt;
  const int tiles = ppx_cdiv(X, UG_T);
  const int ntile_ctas = tiles * (tiles + 1) / 2;
  int nz = ppx_cdiv(4 * ctx->sm_count, ntile_ctas);
  if (nz > 32) nz = 32;
  int64_t max_z_by_c = (C + UG_K - 1) / UG_K;
  if (nz > max_z_by_c) nz = (int)max_z_by_c;
  if (nz < 1) nz = 1;
  ppx_ws_reset(ctx);
  double *parts = nullptr;
  while (nz >= 1 && !(parts = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)nz * X * X))) nz /= 2;
  if (!parts) return ppx_set_err(ctx, PPX_ENOMEM, "unfold_gram needs %lld bytes of workspace", (long long)(8 * X * X));
  int64_t c_per_z = (C + nz - 1) / nz;
  c_per_z = ((c_per_z + UG_K - 1) / UG_K) * UG_K;
  unfold_gram_kernel<<<dim3(tiles, tiles, nz), 256, 0, ctx->stream>>>(T, L, X, Rt, C, c_per_z, parts);
  PPX_CHECK_LAUNCH(ctx);
  unfold_gram_reduce_kernel<<<ppx_cdiv(X * X, 256), 256, 0, ctx->stream>>>(parts, X, nz, MTM);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_sym_eig_topk(ppx_ctx *ctx, double *MTM, int64_t s, int r, double *U, double *evals_out) {
  PPX_REQUIRE(ctx, MTM && U && s >= 1 && r >= 1 && r <= s, "1 <= r <= s");
  PPX_REQUIRE(ctx, s <= 16384, "s <= 16384");
  const int n = (int)((s + 1) & ~(int64_t)1);
  ppx_ws_reset(ctx);
  double *A = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)n * n);
  double *Q = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)n * n);
  double *cs = (double *)ppx_ws_alloc(ctx, sizeof(double) * (4 * (size_t)(n / 2) + 8));
  if (!A || !Q || !cs)
    return ppx_set_err(ctx, PPX_ENOMEM, "sym_eig_topk needs %lld bytes of workspace", (long long)(16LL * n * n + 4096));
  double *norms = cs + 4 * (n / 2);
  pad_copy_kernel<<<ppx_cdiv((int64_t)n * n, 256), 256, 0, ctx->stream>>>(MTM, (int)s, n, A);
  PPX_CHECK_LAUNCH(ctx);
  int max_sweeps = 40;
  int nn = n;
  int blocks_per_sm = 0;
  PPX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, jacobi_global_kernel, 256, 0));
  if (blocks_per_sm < 1) return ppx_set_err(ctx, PPX_ECUDA, "jacobi kernel cannot be resident");
  int grid = ctx->sm_count * (blocks_per_sm > 2 ? 2 : blocks_per_sm);
  // small problems: fewer CTAs make the grid barrier cheaper
  int64_t work = (int64_t)(n / 2) * n;
  int need = (int)((work + 255) / 256);
  if (need < 1) need = 1;
  if (grid > need) grid = need;
  void *args[] = {&A, &Q, &cs, &norms, &nn, &max_sweeps};
  PPX_CUDA(ctx, cudaLaunchCooperativeKernel((void *)jacobi_global_kernel, dim3(grid), dim3(256), args, 0, ctx->stream));
  ctx->launches++;
  eig_select_kernel<<<n, 256, 0, ctx->stream>>>(A, Q, n, (int)s, r, U, evals_out);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_sign_align(ppx_ctx *ctx, double *U, const double *Uref, int64_t s, int r) {
  PPX_REQUIRE(ctx, U && Uref && s >= 1 && r >= 1, "bad arguments");
  sign_align_kernel<<<r, 256, 0, ctx->stream>>>(U, Uref, s);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

}  // extern "C"
