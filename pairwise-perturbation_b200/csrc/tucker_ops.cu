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

// ---- the same for large unfoldings (HOSVD: X x prod(other modes)): 128 x 128 tile, 8 x 8 per thread (4 FMA per shared
// load instead of 2), the next chunk fetched into registers while the current one is multiplied, no division per element
constexpr int UH_T = 128, UH_K = 16;
__global__ void __launch_bounds__(256) unfold_gram128_kernel(const double *__restrict__ T, int64_t L, int64_t X,
                                                             int64_t Rt, int64_t C, int64_t c_per_z,
                                                             double *__restrict__ out_z) {
  __shared__ double Ap[UH_K][UH_T + 1];
  __shared__ double Aq[UH_K][UH_T + 1];
  const int tid = threadIdx.x;
  const int64_t p0 = (int64_t)blockIdx.x * UH_T, q0 = (int64_t)blockIdx.y * UH_T;
  if (q0 > p0) return;  // lower triangle of tiles only; mirrored by the reduction kernel
  const bool diag = (p0 == q0);
  const int64_t cb = (int64_t)blockIdx.z * c_per_z;
  int64_t ce = cb + c_per_z;
  if (ce > C) ce = C;
  const int tx = tid % 16, ty = tid / 16;  // rows p0 + tx + 16 i, columns q0 + ty + 16 j
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.0;
  // element (c, p) lives at l + L*(p + X*t), c = l + L*t.  L == 1: p is the fastest index -> threads run along p;
  // L > 1: c is -> threads run along c (16 consecutive c per p)
  const bool pfast = (L == 1);
  const int my_cc = pfast ? tid / 128 : tid % 16;  // + 2u (pfast) per load u
  const int my_pp = pfast ? tid % 128 : tid / 16;  // + 16u (!pfast)
  int64_t lt_l = 0, lt_t = 0;                       // (l, t) of c = chunk start + my_cc, kept incrementally
  if (!pfast) {
    const int64_t c = cb + my_cc;
    lt_t = c / L;
    lt_l = c - lt_t * L;
  }
  double rp[8], rq[8];
  auto fetch = [&](int64_t c0) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int cc = pfast ? my_cc + 2 * u : my_cc;
      const int pp = pfast ? my_pp : my_pp + 16 * u;
      const int64_t c = c0 + cc;
      double vp = 0.0, vq = 0.0;
      if (c < ce) {
        const int64_t base = pfast ? X * c : lt_l + L * X * lt_t;
        const int64_t st = pfast ? 1 : L;
        if (p0 + pp < X) vp = T[base + st * (p0 + pp)];
        if (!diag && q0 + pp < X) vq = T[base + st * (q0 + pp)];
      }
      rp[u] = vp;
      rq[u] = vq;
    }
  };
  auto advance = [&]() {  // (l, t) += UH_K along c
    if (!pfast) {
      lt_l += UH_K;
      while (lt_l >= L) {
        lt_l -= L;
        lt_t++;
      }
    }
  };
  fetch(cb);
  for (int64_t c0 = cb; c0 < ce; c0 += UH_K) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int cc = pfast ? my_cc + 2 * u : my_cc;
      const int pp = pfast ? my_pp : my_pp + 16 * u;
      Ap[cc][pp] = rp[u];
      if (!diag) Aq[cc][pp] = rq[u];
    }
    __syncthreads();
    if (c0 + UH_K < ce) {
      advance();
      fetch(c0 + UH_K);
    }
    const double(*Bq)[UH_T + 1] = diag ? Ap : Aq;
#pragma unroll 4
    for (int cc = 0; cc < UH_K; cc++) {
      double a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = Ap[cc][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; j++) b[j] = Bq[cc][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  double *o = out_z + (int64_t)blockIdx.z * X * X;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int64_t p = p0 + tx + 16 * i, q = q0 + ty + 16 * j;
      if (p < X && q < X) o[p + X * q] = acc[i][j];
    }
}

__global__ void unfold_gram_reduce_kernel(const double *__restrict__ parts, int64_t X, int nz, int tile,
                                          double *__restrict__ MTM) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= X * X) return;
  int64_t p = idx % X, q = idx / X;
  // tiles with q0 > p0 were skipped: read the mirrored element
  int64_t pp = p, qq = q;
  if ((q / tile) > (p / tile)) {
    pp = q;
    qq = p;
  }
  double s = 0.0;
  for (int z = 0; z < nz; z++) s += parts[(int64_t)z * X * X + pp + X * qq];
  MTM[idx] = s;
}

// ---- cooperative one-sided (Hestenes) Jacobi in global memory (matrix is L2 resident) -----------------------------
// B = A (n x n symmetric positive semi-definite, column-major, destroyed).  Plane rotations of column pairs make the
// columns mutually orthogonal: at convergence B = A V = V Lambda, i.e. column j is lambda_j v_j -- the eigenvalue is
// its norm and the eigenvector its direction, so the rotations never have to be accumulated separately.
// Every round of the round-robin ordering holds n/2 disjoint pairs: one CTA per pair, which reads its two columns
// once, forms the three inner products, rotates and writes back; pairs of a round touch disjoint columns, so a round
// needs ONE grid barrier (the two-sided variant this replaces needed five and also updated rows).  The iteration ends
// when a whole sweep applied no rotation above the threshold.  (Writing the pair back larger-column-first, de Rijk's
// sorting, was tried: with the round-robin ordering it DOUBLES the number of sweeps -- 20 instead of 11 at n = 128.)
// Warm start: any orthogonal V0 is a valid start, B = A V0; with V0 = the eigenvectors of a nearby matrix (the same
// mode in the previous HOOI sweep) the columns start almost orthogonal and few sweeps remain.  `Vout` (optional)
// receives all n normalised columns for that purpose.
// `count`: one int per sweep (zeroed by the launcher), rotations applied in that sweep (integer atomics: exact).
constexpr int JT = 128;   // threads per column pair
constexpr int JG = 1;     // pairs per CTA (4 pairs per 512-thread CTA was measured: 3.65 us per round instead of 3.2)
constexpr int JE = 8;     // column elements per thread held in registers per pass (n <= JT*JE in one pass)
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(JT) : "memory"); }
// `count`: [0, max_sweeps) rotations per sweep, [max_sweeps] sweeps done; `maxrot`: largest |sin| applied per sweep
// (positive doubles compared as integers).  Jacobi converges quadratically: after a sweep whose largest rotation is
// below 1e-8 what is left is of order 1e-16, so that sweep is the last -- no extra sweep just to see zero rotations.
__global__ void __launch_bounds__(JT *JG) jacobi_onesided_kernel(double *__restrict__ B, int n, int ncols,
                                                                 int max_sweeps, int *__restrict__ count,
                                                                 unsigned long long *__restrict__ maxrot,
                                                                 double *__restrict__ evals,
                                                                 double *__restrict__ Vout) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[2][JG][3][JT / 32];
  const int grp = threadIdx.x / JT, tid = threadIdx.x % JT, lane = tid & 31, warp = tid >> 5;
  const int m = ncols;            // even number of seats; seat index >= number of real columns: idle
  const double tol = 2.3e-16 * sqrt((double)n);
  const int half = m / 2;
  const bool onepass = n <= JT * JE;
  int sweep = 0, buf = 0;
  for (; sweep < max_sweeps; sweep++) {
    for (int round = 0; round < m - 1; round++) {
      for (int pr = blockIdx.x * JG + grp; pr < half; pr += gridDim.x * JG) {
        int p, q;
        if (pr == 0) {
          p = m - 1;
          q = round % (m - 1);
        } else {
          p = (round + pr) % (m - 1);
          q = (round + m - 1 - pr) % (m - 1);
        }
        if (p > q) {
          const int t = p;
          p = q;
          q = t;
        }
        if (q >= n) continue;  // padded seat (odd n)
        double *bp = B + (int64_t)n * p, *bq = B + (int64_t)n * q;
        double a = 0.0, b = 0.0, g = 0.0;
        double xp[JE], xq[JE];
        if (onepass) {
#pragma unroll
          for (int u = 0; u < JE; u++) {
            const int i = tid + JT * u;
            xp[u] = i < n ? bp[i] : 0.0;
            xq[u] = i < n ? bq[i] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < JE; u++) {
            a = fma(xp[u], xp[u], a);
            b = fma(xq[u], xq[u], b);
            g = fma(xp[u], xq[u], g);
          }
        } else {
          for (int i = tid; i < n; i += JT) {
            const double vp = bp[i], vq = bq[i];
            a = fma(vp, vp, a);
            b = fma(vq, vq, b);
            g = fma(vp, vq, g);
          }
        }
        a = ppx_warp_sum(a);
        b = ppx_warp_sum(b);
        g = ppx_warp_sum(g);
        // one barrier per pair: partial sums go to alternating buffers, every thread adds them up in the same order
        buf ^= 1;
        if (lane == 0) {
          red[buf][grp][0][warp] = a;
          red[buf][grp][1][warp] = b;
          red[buf][grp][2][warp] = g;
        }
        group_bar(grp);
        a = b = g = 0.0;
#pragma unroll
        for (int w = 0; w < JT / 32; w++) {
          a += red[buf][grp][0][w];
          b += red[buf][grp][1][w];
          g += red[buf][grp][2][w];
        }
        // rotate when the columns are not orthogonal to working precision
        // (threshold sqrt(n) eps, as LAPACK's dgesvj: a recomputed inner product of orthogonal columns is that large)
        if (!(fabs(g) > tol * sqrt(a * b) && fabs(g) > 1e-300)) continue;
        const double zeta = (b - a) / (2.0 * g);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t);
        const double sn = c * t;
        if (tid == 0) {
          atomicAdd(&count[sweep], 1);
          atomicMax(&maxrot[sweep], (unsigned long long)__double_as_longlong(fabs(sn)));
        }
        if (onepass) {
#pragma unroll
          for (int u = 0; u < JE; u++) {
            const int i = tid + JT * u;
            if (i < n) {
              bp[i] = c * xp[u] - sn * xq[u];
              bq[i] = sn * xp[u] + c * xq[u];
            }
          }
        } else {
          for (int i = tid; i < n; i += JT) {
            const double vp = bp[i], vq = bq[i];
            bp[i] = c * vp - sn * vq;
            bq[i] = sn * vp + c * vq;
          }
        }
      }
      grid.sync();
    }
    if (count[sweep] == 0 || __longlong_as_double((long long)maxrot[sweep]) < 1e-8) {
      sweep++;
      break;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) count[max_sweeps] = sweep;
  // eigenvalue estimates: column norms (one column per group)
  for (int j = blockIdx.x * JG + grp; j < n; j += gridDim.x * JG) {
    const double *bj = B + (int64_t)n * j;
    double a = 0.0;
    for (int i = tid; i < n; i += JT) a = fma(bj[i], bj[i], a);
    a = ppx_warp_sum(a);
    buf ^= 1;
    if (lane == 0) red[buf][grp][0][warp] = a;
    group_bar(grp);
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < JT / 32; w++) v += red[buf][grp][0][w];
    const double nrm = sqrt(v);
    if (tid == 0) evals[j] = nrm;
    if (Vout) {
      double *vj = Vout + (int64_t)n * j;
      for (int i = tid; i < n; i += JT) vj[i] = nrm > 0.0 ? bj[i] / nrm : (i == j ? 1.0 : 0.0);
    }
  }
}

// pick the r largest eigenvalues, write the normalised columns in decreasing order, largest component positive
__global__ void __launch_bounds__(256) eig_select_kernel(const double *__restrict__ B, const double *__restrict__ ev,
                                                         int n, int r, double *__restrict__ U,
                                                         double *__restrict__ evals_out) {
  const int i = blockIdx.x;  // one block per candidate column
  __shared__ int rank_s;
  if (threadIdx.x == 0) rank_s = 0;
  __syncthreads();
  const double li = ev[i];
  int cnt = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    if (j == i) continue;
    const double lj = ev[j];
    if (lj > li || (lj == li && j < i)) cnt++;
  }
  atomicAdd(&rank_s, cnt);  // integer: order independent
  __syncthreads();
  if (rank_s >= r) return;
  const int k = rank_s;
  const double *bi_ = B + (int64_t)n * i;
  __shared__ double red_v[256];
  __shared__ int red_i[256];
  double best = -1.0;
  int bi = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double v = fabs(bi_[j]);
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
  // a zero column (rank-deficient input) has no direction: emit the unit vector of its own index
  const double nrm = li;
  const double sc = nrm > 0.0 ? (bi_[red_i[0]] < 0.0 ? -1.0 : 1.0) / nrm : 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x)
    U[j + (int64_t)n * k] = nrm > 0.0 ? sc * bi_[j] : (j == i ? 1.0 : 0.0);
  if (threadIdx.x == 0 && evals_out) evals_out[k] = li;
}

// symmetrised copy of the input (the Gram kernels mirror exactly, callers' matrices may not)
__global__ void sym_copy_kernel(const double *__restrict__ M, int n, double *__restrict__ A) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * n) return;
  const int i = (int)(idx % n), j = (int)(idx / n);
  A[idx] = 0.5 * (M[i + (int64_t)n * j] + M[j + (int64_t)n * i]);
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
  // large unfoldings (HOSVD): DMMA SYRK (gram_dmma.cu); PPX_NO_GRAM_DMMA=1 selects the DFMA tile kernel (experiments)
  const bool big = X >= 192 && C >= 4096;
  static const bool no_dmma = getenv("PPX_NO_GRAM_DMMA") != nullptr;
  if (big && !no_dmma) {
    ppx_ws_reset(ctx);
    int max_splits = 64;
    double *parts = nullptr;
    while (max_splits >= 1 && !(parts = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)max_splits * X * X)))
      max_splits /= 2;
    if (!parts) return ppx_set_err(ctx, PPX_ENOMEM, "unfold_gram needs %lld bytes of workspace", (long long)(8 * X * X));
    const int nz = ppx_gram_dmma(ctx, T, L, X, Rt, parts, max_splits);
    if (nz < 0) return nz;
    unfold_gram_reduce_kernel<<<ppx_cdiv(X * X, 256), 256, 0, ctx->stream>>>(parts, X, nz, 128, MTM);
    PPX_CHECK_LAUNCH(ctx);
    return PPX_OK;
  }
  const int tile = big ? UH_T : UG_T;
  const int tiles = ppx_cdiv(X, tile);
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
  if (big)
    unfold_gram128_kernel<<<dim3(tiles, tiles, nz), 256, 0, ctx->stream>>>(T, L, X, Rt, C, c_per_z, parts);
  else
    unfold_gram_kernel<<<dim3(tiles, tiles, nz), 256, 0, ctx->stream>>>(T, L, X, Rt, C, c_per_z, parts);
  PPX_CHECK_LAUNCH(ctx);
  unfold_gram_reduce_kernel<<<ppx_cdiv(X * X, 256), 256, 0, ctx->stream>>>(parts, X, nz, tile, MTM);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_sym_eig_topk_warm(ppx_ctx *ctx, double *MTM, int64_t s, int r, double *U, double *evals_out, double *basis,
                          int basis_valid) {
  PPX_REQUIRE(ctx, MTM && U && s >= 1 && r >= 1 && r <= s, "1 <= r <= s");
  PPX_REQUIRE(ctx, s <= 16384, "s <= 16384");
  const int n = (int)s;
  const int seats = (n + 1) & ~1;
  const int max_sweeps = 40;
  // r << s: Chebyshev-filtered subspace iteration (GEMM-shaped, eig_chfsi.cu); `basis` then carries its block of
  // r + 24 vectors.  Not converged -> the Jacobi solver below, started cold.
  if (ppx_eig_chfsi_applicable(s, r)) {
    ppx_ws_reset(ctx);
    double *Ac = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)n * n);
    if (Ac) {
      sym_copy_kernel<<<ppx_cdiv((int64_t)n * n, 256), 256, 0, ctx->stream>>>(MTM, n, Ac);
      PPX_CHECK_LAUNCH(ctx);
      const int rc = ppx_eig_chfsi(ctx, Ac, n, r, U, evals_out, basis, basis_valid);
      if (rc <= 0) return rc;
      if (getenv("PPX_EIG_VERBOSE")) fprintf(stderr, "sym_eig_topk n=%d r=%d: subspace iteration gave up, Jacobi\n", n, r);
    }
    basis_valid = 0;
  }
  const bool warm = basis && basis_valid && n > 1;
  ppx_ws_reset(ctx);
  double *B = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)n * n);
  double *A = warm ? (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)n * n) : B;
  double *ev = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)n);
  int *count = (int *)ppx_ws_alloc(ctx, sizeof(int) * (max_sweeps + 2) + sizeof(unsigned long long) * max_sweeps);
  if (!B || !A || !ev || !count)
    return ppx_set_err(ctx, PPX_ENOMEM, "sym_eig_topk needs %lld bytes of workspace",
                       (long long)(16LL * n * n + 8LL * n + 4096));
  unsigned long long *maxrot = (unsigned long long *)(count + max_sweeps + 2);  // 8-byte aligned: ws blocks are 256-byte
  PPX_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * (max_sweeps + 2) + sizeof(unsigned long long) * max_sweeps,
                                ctx->stream));
  sym_copy_kernel<<<ppx_cdiv((int64_t)n * n, 256), 256, 0, ctx->stream>>>(MTM, n, A);
  PPX_CHECK_LAUNCH(ctx);
  if (warm) {  // B = A V0 (DMMA GEMM of the first-contraction kernel: rows l = i, contracted mode x, "rank" = n columns)
    const int rc = ppx_ttm_impl(ctx, A, n, n, 1, basis, n, n, B, 0, 0, true, false);
    if (rc) return rc;
  }
  if (n > 1) {
    int blocks_per_sm = 0;
    PPX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, jacobi_onesided_kernel, JT * JG, 0));
    if (blocks_per_sm < 1) return ppx_set_err(ctx, PPX_ECUDA, "jacobi kernel cannot be resident");
    int grid = ctx->sm_count * blocks_per_sm;
    const int need = (seats / 2 + JG - 1) / JG;  // JG pairs of a round per CTA
    if (grid > need) grid = need;
    int nn = n, ss = seats, ms = max_sweeps;
    void *args[] = {&B, &nn, &ss, &ms, &count, &maxrot, &ev, &basis};
    PPX_CUDA(ctx, cudaLaunchCooperativeKernel((void *)jacobi_onesided_kernel, dim3(grid), dim3(JT * JG), args, 0,
                                              ctx->stream));
    ctx->launches++;
  } else {
    PPX_CUDA(ctx, cudaMemcpyAsync(ev, B, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  eig_select_kernel<<<n, 256, 0, ctx->stream>>>(B, ev, n, r, U, evals_out);
  PPX_CHECK_LAUNCH(ctx);
  if (getenv("PPX_EIG_VERBOSE")) {  // development aid: rotations applied per sweep
    int h[64];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(h, count, sizeof(int) * (max_sweeps + 1), cudaMemcpyDeviceToHost);
    fprintf(stderr, "sym_eig_topk n=%d%s: %d sweeps, rotations:", n, warm ? " (warm)" : "", h[max_sweeps]);
    for (int i = 0; i < h[max_sweeps] && i < max_sweeps; i++) fprintf(stderr, " %d", h[i]);
    fprintf(stderr, "\n");
  }
  return PPX_OK;
}

int ppx_sym_eig_topk(ppx_ctx *ctx, double *MTM, int64_t s, int r, double *U, double *evals_out) {
  return ppx_sym_eig_topk_warm(ctx, MTM, s, r, U, evals_out, nullptr, 0);
}

int ppx_sign_align(ppx_ctx *ctx, double *U, const double *Uref, int64_t s, int r) {
  PPX_REQUIRE(ctx, U && Uref && s >= 1 && r >= 1, "bad arguments");
  sign_align_kernel<<<r, 256, 0, ctx->stream>>>(U, Uref, s);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

}  // extern "C"
