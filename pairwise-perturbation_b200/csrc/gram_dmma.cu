// Gram of a large mode-i unfolding on the FP64 tensor pipe:  MTM[p, q] = sum_c A[c, p] A[c, q],
// A[c, p] = T[l + L*(p + X*t)], c = (l, t)  -- unroll_tensor_contraction (common.cxx:205-223) as HOSVD calls it
// (als_Tucker.cxx:12-23): 800 x 640 000 -> 800 x 800 at BASELINE configs[2], 0.41 TFLOP per mode for the lower half.
//
// A SYRK with both operands streamed from the tensor.  CTA = 128 threads, tile 128 (p) x 64 (q), each warp 32 x 64 in
// registers fed by DMMA.8x8x4 (the geometry of the first-contraction kernel, k1_ttm_first.cu); only tile pairs in or
// below the diagonal of 128-blocks are computed (the reduction kernel mirrors the rest), the contracted range is
// split over gridDim.y so that the grid fills 2 x 148 CTA slots, and the partial tiles are summed in a fixed order.
// Operand tiles (16 deep) go through a 4-stage cp.async pipeline; two shared-memory layouts, both padded so that every
// DMMA fragment load is bank-conflict free:
//   CMAJOR (L > 1): c is contiguous in memory for a fixed p -> tiles [p][c], pitch 20
//   PMAJOR (L == 1): p is contiguous for a fixed c          -> tiles [c][p], pitch 132 / 68
#include "ppx_internal.h"

namespace {

constexpr int GB_M = 128, GB_N = 64, GB_K = 16, GB_THREADS = 128;
constexpr int GB_LDK = 20, GB_LDM = 132, GB_LDN = 68;

struct GramParams {
  const double *T;
  double *parts;  // [ksplit][X*X]
  int64_t L, X, Rt;
  int64_t lchunks;  // CMAJOR: 16-deep chunks per t (ceil(L/16)); PMAJOR: unused
  int64_t nk;       // chunks in total
  int64_t cps;      // chunks per split
  int ptiles;       // 128-row tiles
  int vec;          // 16-byte copies allowed
};

template <bool CMAJOR>
__host__ __device__ constexpr int gram_stage_doubles() {
  return CMAJOR ? (GB_M + GB_N) * GB_LDK : GB_K * (GB_LDM + GB_LDN);
}
// pipeline depth: what two CTAs per SM allow (30.7 KB / 25.6 KB per stage)
template <bool CMAJOR>
__host__ __device__ constexpr int gram_stages() {
  return CMAJOR ? 3 : 4;
}

template <bool CMAJOR>
__global__ void __launch_bounds__(GB_THREADS, 2) gram_dmma_kernel(GramParams p) {
  extern __shared__ __align__(16) double gsm[];
  constexpr int STAGE_D = gram_stage_doubles<CMAJOR>();
  constexpr int GB_STAGES = gram_stages<CMAJOR>();
  constexpr int A_D = CMAJOR ? GB_M * GB_LDK : GB_K * GB_LDM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  // tile pair: blockIdx.x enumerates (pt, qh) with qh (64-wide) <= 2*pt + 1
  int pt = 0, rem = blockIdx.x;
  while (rem >= 2 * pt + 2) {
    rem -= 2 * pt + 2;
    pt++;
  }
  const int64_t p0 = (int64_t)pt * GB_M, q0 = (int64_t)rem * GB_N;
  if (q0 >= p.X) return;
  const int64_t kb = (int64_t)blockIdx.y * p.cps;
  int64_t ke = kb + p.cps;
  if (ke > p.nk) ke = p.nk;
  const int total = ke > kb ? (int)(ke - kb) : 0;

  auto load_chunk = [&](int stage, int64_t kc) {
    double *As = gsm + stage * STAGE_D;
    double *Bs = As + A_D;
    if (CMAJOR) {
      const int64_t t = kc / p.lchunks;
      const int64_t l0 = (kc - t * p.lchunks) * GB_K;
      const double *base = p.T + p.L * p.X * t;
      const int kp = tid & 7, mb = tid >> 3;
      const int64_t lg = l0 + 2 * kp;
      const int kv = lg + 1 < p.L ? 2 : (lg < p.L ? 1 : 0);
#pragma unroll
      for (int i = 0; i < (GB_M + GB_N) / 16; i++) {  // rows mb + 16 i of the stacked [A ; B] tile
        const int row = mb + 16 * i;
        const bool isA = row < GB_M;
        const int64_t pg = isA ? p0 + row : q0 + (row - GB_M);
        const bool rv = pg < p.X;
        double *dst = (isA ? As + row * GB_LDK : Bs + (row - GB_M) * GB_LDK) + 2 * kp;
        const double *src = base + lg + p.L * pg;
        if (p.vec) {
          ppx_cp_async16(dst, (rv && kv) ? src : p.T, !rv ? 0 : (kv == 2 ? 16 : (kv == 1 ? 8 : 0)));
        } else {
          ppx_cp_async8(dst, (rv && kv >= 1) ? src : p.T, (rv && kv >= 1) ? 8 : 0);
          ppx_cp_async8(dst + 1, (rv && kv == 2) ? src + 1 : p.T, (rv && kv == 2) ? 8 : 0);
        }
      }
    } else {
      const int64_t c0 = kc * GB_K;
      const int64_t C = p.Rt;  // L == 1: c = t
      {  // A: [16 c][128 p]
        const int ml = 2 * (tid & 63), kb2 = tid >> 6;
        const int64_t pg = p0 + ml;
        const int pv = pg + 1 < p.X ? 2 : (pg < p.X ? 1 : 0);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int k = kb2 + 2 * i;
          const bool cv = c0 + k < C;
          double *dst = As + k * GB_LDM + ml;
          const double *src = p.T + pg + p.X * (c0 + k);
          if (p.vec) {
            ppx_cp_async16(dst, (cv && pv) ? src : p.T, !cv ? 0 : (pv == 2 ? 16 : (pv == 1 ? 8 : 0)));
          } else {
            ppx_cp_async8(dst, (cv && pv >= 1) ? src : p.T, (cv && pv >= 1) ? 8 : 0);
            ppx_cp_async8(dst + 1, (cv && pv == 2) ? src + 1 : p.T, (cv && pv == 2) ? 8 : 0);
          }
        }
      }
      {  // B: [16 c][64 q]
        const int nl = 2 * (tid & 31), kb2 = tid >> 5;
        const int64_t qg = q0 + nl;
        const int qv = qg + 1 < p.X ? 2 : (qg < p.X ? 1 : 0);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int k = kb2 + 4 * i;
          const bool cv = c0 + k < C;
          double *dst = Bs + k * GB_LDN + nl;
          const double *src = p.T + qg + p.X * (c0 + k);
          if (p.vec) {
            ppx_cp_async16(dst, (cv && qv) ? src : p.T, !cv ? 0 : (qv == 2 ? 16 : (qv == 1 ? 8 : 0)));
          } else {
            ppx_cp_async8(dst, (cv && qv >= 1) ? src : p.T, (cv && qv >= 1) ? 8 : 0);
            ppx_cp_async8(dst + 1, (cv && qv == 2) ? src + 1 : p.T, (cv && qv == 2) ? 8 : 0);
          }
        }
      }
    }
  };

  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int c = 0; c < GB_STAGES - 1; c++) {
    if (c < total) load_chunk(c, kb + c);
    ppx_cp_async_commit();
  }
  for (int c = 0; c < total; c++) {
    ppx_cp_async_wait<GB_STAGES - 2>();
    __syncthreads();
    {
      const int cn = c + GB_STAGES - 1;
      if (cn < total) load_chunk(cn % GB_STAGES, kb + cn);
      ppx_cp_async_commit();
    }
    const double *As = gsm + (c % GB_STAGES) * STAGE_D;
    const double *Bs = As + A_D;
#pragma unroll
    for (int kk = 0; kk < GB_K / 4; kk++) {
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; i++)
        a[i] = CMAJOR ? As[(32 * warp + 8 * i + g) * GB_LDK + 4 * kk + t4]
                      : As[(4 * kk + t4) * GB_LDM + 32 * warp + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 8; j++)
        b[j] = CMAJOR ? Bs[(8 * j + g) * GB_LDK + 4 * kk + t4] : Bs[(4 * kk + t4) * GB_LDN + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) ppx_dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  ppx_cp_async_wait<0>();
  double *o = p.parts + (int64_t)blockIdx.y * p.X * p.X;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int64_t pg = p0 + 32 * warp + 8 * i + g;
    if (pg >= p.X) continue;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int64_t qg = q0 + 8 * j + 2 * t4;
      if (qg < p.X) o[pg + p.X * qg] = acc[i][j][0];
      if (qg + 1 < p.X) o[pg + p.X * (qg + 1)] = acc[i][j][1];
    }
  }
}

}  // namespace

// Opt-in shared memory of the two kernels; per device, called from ppx_ctx_create.
int ppx_gram_init(ppx_ctx *ctx) {
  cudaError_t e = cudaFuncSetAttribute(gram_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sizeof(double) * gram_stages<true>() * gram_stage_doubles<true>()));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gram_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(double) * gram_stages<false>() * gram_stage_doubles<false>()));
  if (e != cudaSuccess) return ppx_set_err(ctx, PPX_ECUDA, "gram_dmma init: %s", cudaGetErrorString(e));
  return PPX_OK;
}

// Partial Grams into `parts` ([nz][X*X], only 128-blocks in or below the diagonal are written; the caller's reduction
// mirrors the others with tile = 128).  Returns the number of K splits used, 0 when the shape is not taken, < 0 on error.
int ppx_gram_dmma(ppx_ctx *ctx, const double *T, int64_t L, int64_t X, int64_t Rt, double *parts, int max_splits) {
  GramParams p;
  p.T = T;
  p.parts = parts;
  p.L = L;
  p.X = X;
  p.Rt = Rt;
  const bool cmajor = L > 1;
  p.lchunks = cmajor ? (L + GB_K - 1) / GB_K : 1;
  p.nk = cmajor ? p.lchunks * Rt : (Rt + GB_K - 1) / GB_K;
  p.ptiles = (int)((X + GB_M - 1) / GB_M);
  const bool a16 = (((uintptr_t)T) & 15) == 0;
  p.vec = cmajor ? (a16 && L % 2 == 0) : (a16 && X % 2 == 0);
  // tile pairs: for p-tile pt the 64-wide q tiles 0 .. 2 pt + 1
  int pairs = 0;
  for (int pt = 0; pt < p.ptiles; pt++) pairs += 2 * pt + 2;
  int nz = (2 * ctx->sm_count) / pairs;  // one wave of the 2 x 148 CTA slots (rounding up would leave a ragged second one)
  if (nz > max_splits) nz = max_splits;
  if ((int64_t)nz > p.nk / 8) nz = (int)(p.nk / 8);
  if (nz < 1) nz = 1;
  p.cps = (p.nk + nz - 1) / nz;
  nz = (int)((p.nk + p.cps - 1) / p.cps);
  const size_t smem = sizeof(double) * (cmajor ? gram_stages<true>() * gram_stage_doubles<true>()
                                              : gram_stages<false>() * gram_stage_doubles<false>());
  if (cmajor)
    gram_dmma_kernel<true><<<dim3(pairs, nz), GB_THREADS, smem, ctx->stream>>>(p);
  else
    gram_dmma_kernel<false><<<dim3(pairs, nz), GB_THREADS, smem, ctx->stream>>>(p);
  PPX_CHECK_LAUNCH(ctx);
  return nz;
}
