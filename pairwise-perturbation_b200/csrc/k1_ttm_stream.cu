// K1, streaming variant -- the first contraction when both the contracted extent and the rank are small
//     out[l, t, r] = sum_x V[l, x, t] * W[x, r],   X <= 64, R <= 16
// (reference: common.cxx:56, als_CP.cxx:378-379 -- the PP operator build of a deep tree, BASELINE configs[3]:
// order 6, s = 40, R = 10; the size-3 colour mode of the coil-shaped tensor, configs[4]).
//
// At these sizes the contraction is a bandwidth problem with a large output: 2 R flop per 8 (1 + R/X) bytes is
// 1.6-2.4 flop/B at R = 10, a third of what the FP64 pipe sustains at the HBM rate, and the output is R/X of the
// input (25 % at s = 40, 333 % at X = 3).  The DMMA tile kernels (k1_ttm_tma.cu, k1_ttm_first.cu) stage 128 x 16
// tiles and scatter 8-row fragments; measured on the PP build at configs[3] they reach 0.61-0.78 of the HBM rate.
// Here every thread owns one (or two adjacent) output rows, streams its X inputs with eight independent loads in
// flight, keeps the R accumulators in registers (W, zero padded to RB columns, is broadcast from shared memory) and
// writes each output column as one fully coalesced line per warp: plain DFMA, no tiles, no fragment layout.
//   L > 1 : lanes run along l (rows contiguous for a fixed x);
//   L == 1: the contracted mode is the fastest index; a CTA stages 128 rows x X (one contiguous slab, coalesced loads)
//           in shared memory with an odd pitch and every thread then reads its own row conflict free.
#include "ppx_internal.h"

namespace {

constexpr int ST_THREADS_M = 256;
constexpr int ST_ROWS_K = 128;  // rows (= threads) per CTA of the L == 1 kernel

template <int RB, int VEC>
__global__ void __launch_bounds__(ST_THREADS_M) ttm_stream_m_kernel(const double *__restrict__ V,
                                                                    const double *__restrict__ W, int64_t ldw,
                                                                    double *__restrict__ out, int64_t L, int64_t X,
                                                                    int64_t Rt, int R, int inplace, int accumulate) {
  extern __shared__ __align__(16) double st_smem[];
  double *Wsm = st_smem;  // [X][RB]
  for (int idx = threadIdx.x; idx < (int)X * RB; idx += ST_THREADS_M) {
    const int x = idx / RB, r = idx - x * RB;
    Wsm[idx] = r < R ? W[x + ldw * r] : 0.0;
  }
  __syncthreads();
  const int64_t Mtot = L * Rt;
  const int64_t nvec = Mtot / VEC;  // VEC == 2 requires L even: a pair never straddles two t
  const int64_t LX = L * X;
  for (int64_t mv = (int64_t)blockIdx.x * ST_THREADS_M + threadIdx.x; mv < nvec;
       mv += (int64_t)gridDim.x * ST_THREADS_M) {
    const int64_t m = mv * VEC;
    const int64_t t = m / L, l = m - t * L;
    const double *vp = V + l + LX * t;
    double acc[VEC][RB];
#pragma unroll
    for (int e = 0; e < VEC; e++)
#pragma unroll
      for (int r = 0; r < RB; r++) acc[e][r] = 0.0;
    for (int64_t x0 = 0; x0 < X; x0 += 8) {
      double v[8][VEC];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (x0 + u < X) {
          if (VEC == 2) {
            const double2 q = *reinterpret_cast<const double2 *>(vp + L * (x0 + u));
            v[u][0] = q.x;
            v[u][VEC - 1] = q.y;
          } else {
            v[u][0] = vp[L * (x0 + u)];
          }
        } else {
#pragma unroll
          for (int e = 0; e < VEC; e++) v[u][e] = 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (x0 + u < X) {
          const double2 *w2 = reinterpret_cast<const double2 *>(Wsm + (x0 + u) * RB);
#pragma unroll
          for (int r2 = 0; r2 < RB / 2; r2++) {
            const double2 w = w2[r2];
#pragma unroll
            for (int e = 0; e < VEC; e++) {
              acc[e][2 * r2] = fma(v[u][e], w.x, acc[e][2 * r2]);
              acc[e][2 * r2 + 1] = fma(v[u][e], w.y, acc[e][2 * r2 + 1]);
            }
          }
        }
      }
    }
    const int64_t base = inplace ? l + L * (int64_t)R * t : m;
    const int64_t cstride = inplace ? L : Mtot;
#pragma unroll
    for (int r = 0; r < RB; r++) {
      if (r < R) {
        double *o = out + base + cstride * r;
        if (VEC == 2) {
          double2 q = make_double2(acc[0][r], acc[VEC - 1][r]);
          if (accumulate) {
            const double2 old = *reinterpret_cast<double2 *>(o);
            q.x += old.x;
            q.y += old.y;
          }
          *reinterpret_cast<double2 *>(o) = q;
        } else {
          *o = accumulate ? (*o + acc[0][r]) : acc[0][r];
        }
      }
    }
  }
}

template <int RB>
__global__ void __launch_bounds__(ST_ROWS_K) ttm_stream_k_kernel(const double *__restrict__ V,
                                                                 const double *__restrict__ W, int64_t ldw,
                                                                 double *__restrict__ out, int64_t X, int64_t Mtot,
                                                                 int R, int pitch, int inplace, int accumulate) {
  extern __shared__ __align__(16) double st_smem[];
  double *Wsm = st_smem;                                // [X][RB]
  double *slab = st_smem + (((int)X * RB + 1) & ~1);  // [ST_ROWS_K][pitch]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < (int)X * RB; idx += ST_ROWS_K) {
    const int x = idx / RB, r = idx - x * RB;
    Wsm[idx] = r < R ? W[x + ldw * r] : 0.0;
  }
  const int Xi = (int)X;
  const int step_row = ST_ROWS_K / Xi, step_x = ST_ROWS_K - step_row * Xi;  // advancing the flat index by ST_ROWS_K
  for (int64_t t0 = (int64_t)blockIdx.x * ST_ROWS_K; t0 < Mtot; t0 += (int64_t)gridDim.x * ST_ROWS_K) {
    const int nrows = (int)(Mtot - t0 < ST_ROWS_K ? Mtot - t0 : ST_ROWS_K);
    const int n = nrows * Xi;
    const double *src = V + t0 * X;
    __syncthreads();  // the previous slab has been consumed (and Wsm is complete on the first pass)
    int row = tid / Xi, x = tid - row * Xi;
    for (int i0 = tid; i0 < n; i0 += 4 * ST_ROWS_K) {
      double v[4];
      int rr[4], xx[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        rr[u] = row;
        xx[u] = x;
        v[u] = (i0 + u * ST_ROWS_K < n) ? src[i0 + u * ST_ROWS_K] : 0.0;
        row += step_row;
        x += step_x;
        if (x >= Xi) {
          x -= Xi;
          row++;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (i0 + u * ST_ROWS_K < n) slab[rr[u] * pitch + xx[u]] = v[u];
    }
    __syncthreads();
    if (tid < nrows) {
      double acc[RB];
#pragma unroll
      for (int r = 0; r < RB; r++) acc[r] = 0.0;
      const double *my = slab + tid * pitch;
      for (int xq = 0; xq < Xi; xq++) {
        const double v = my[xq];
        const double2 *w2 = reinterpret_cast<const double2 *>(Wsm + xq * RB);
#pragma unroll
        for (int r2 = 0; r2 < RB / 2; r2++) {
          const double2 w = w2[r2];
          acc[2 * r2] = fma(v, w.x, acc[2 * r2]);
          acc[2 * r2 + 1] = fma(v, w.y, acc[2 * r2 + 1]);
        }
      }
      const int64_t m = t0 + tid;
      const int64_t base = inplace ? (int64_t)R * m : m;
      const int64_t cstride = inplace ? 1 : Mtot;
#pragma unroll
      for (int r = 0; r < RB; r++) {
        if (r < R) {
          double *o = out + base + cstride * r;
          *o = accumulate ? (*o + acc[r]) : acc[r];
        }
      }
    }
  }
}

template <int RB>
int launch_stream(ppx_ctx *ctx, const double *V, int64_t L, int64_t X, int64_t Rt, const double *W, int64_t ldw, int R,
                  double *out, int inplace, int accumulate) {
  const int64_t Mtot = L * Rt;
  if (L == 1) {
    const int pitch = (int)X | 1;
    const size_t smem = sizeof(double) * ((((size_t)X * RB + 1) & ~(size_t)1) + (size_t)ST_ROWS_K * pitch);
    auto kern = ttm_stream_k_kernel<RB>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return ppx_set_err(ctx, PPX_ECUDA, "ttm_stream: %s", cudaGetErrorString(e));
    }
    int64_t blocks = (Mtot + ST_ROWS_K - 1) / ST_ROWS_K;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    kern<<<(int)blocks, ST_ROWS_K, smem, ctx->stream>>>(V, W, ldw, out, X, Mtot, R, pitch, inplace, accumulate);
    PPX_CHECK_LAUNCH(ctx);
    return PPX_OK;
  }
  const bool vec = (L % 2 == 0) && ((((uintptr_t)V) | ((uintptr_t)out)) & 15) == 0;
  const int64_t nvec = vec ? Mtot / 2 : Mtot;
  int64_t blocks = (nvec + ST_THREADS_M - 1) / ST_THREADS_M;
  const int64_t cap = (int64_t)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  const size_t smem = sizeof(double) * (size_t)X * RB;
  if (vec)
    ttm_stream_m_kernel<RB, 2><<<(int)blocks, ST_THREADS_M, smem, ctx->stream>>>(V, W, ldw, out, L, X, Rt, R, inplace,
                                                                                   accumulate);
  else
    ttm_stream_m_kernel<RB, 1><<<(int)blocks, ST_THREADS_M, smem, ctx->stream>>>(V, W, ldw, out, L, X, Rt, R, inplace,
                                                                                   accumulate);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

}  // namespace

// Returns 1 when the shape is not eligible (the caller goes on to the DMMA tile kernels).
int ppx_ttm_stream_try(ppx_ctx *ctx, const double *V, int64_t L, int64_t X, int64_t Rt, const double *W, int64_t ldw,
                       int R, double *out, int inplace, int accumulate) {
  static const bool off = getenv("PPX_NO_STREAM") != nullptr;  // experiments only
  if (off || R > 16 || X < 1) return 1;
  // 1 < L < 16 (a short mode in front of the contracted one, e.g. the colour mode of the coil-shaped tensor): the
  // 128-row tiles of the DMMA kernels gather 8 L-byte pieces (0.41 of the HBM rate at L = 3, X = 128); here the three
  // threads of a slab sweep it front to back and the L1 absorbs the partial sectors, so longer X is taken as well
  if (X > (L > 1 && L < 16 ? 256 : 64)) return 1;
  // L == 1: the slab kernel wins for very short rows (X = 3: 0.92 of the HBM rate against 0.31); from X ~ 16 on the
  // per-row shared-memory pass costs more than the TMA tile kernel's 0.78 (measured at X = 40: 14.1 ms against 8.0 ms;
  // one thread per row with 16-byte loads straight from global memory, relying on L1 for the half-used sectors: 11.8 ms)
  if (L == 1 && X > 8) return 1;
  if (L * Rt < 512) return 1;      // tiny problems: nothing to stream
  if (R <= 4) return launch_stream<4>(ctx, V, L, X, Rt, W, ldw, R, out, inplace, accumulate);
  if (R <= 8) return launch_stream<8>(ctx, V, L, X, Rt, W, ldw, R, out, inplace, accumulate);
  if (R <= 12) return launch_stream<12>(ctx, V, L, X, Rt, W, ldw, R, out, inplace, accumulate);
  return launch_stream<16>(ctx, V, L, X, Rt, W, ldw, R, out, inplace, accumulate);
}
