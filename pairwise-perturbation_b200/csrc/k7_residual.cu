// K7 -- CP reconstruction fused with the residual:  sum (V - [[W_0..W_{N-1}]])^2  without materialising the
//       reconstruction (the reference builds two full-size temporaries: common.cxx:135-197 + als_CP.cxx:183-187),
//       and build_V itself (V_out = [[W]]), used to make the synthetic tensor 'r' (test_ALS.cxx:275-286).
//
// A CTA owns 128 consecutive "rows" m of the first N-1 modes (flattened, first index fastest).  It forms the
// Khatri-Rao rows K[m, r] = prod_{j<N-1} W_j[i_j(m), r] once in shared memory, then sweeps the last mode d in
// chunks: Vhat[m, d] = sum_r K[m, r] * W_last[d, r].  Off the timed path (als_CP.cxx:167,189 excludes it).
#include "ppx_internal.h"

namespace {

constexpr int RS_TM = 128;  // rows per CTA
constexpr int RS_TD = 32;   // last-mode chunk staged in shared memory
constexpr int RS_THREADS = 256;

struct ResArgs {
  const double *w[16];
  int64_t lens[16];
  int N;
  int R;
  int64_t P1;  // product of the first N-1 mode sizes
};

template <bool WRITE>
__global__ void __launch_bounds__(RS_THREADS) cp_reconstruct_kernel(const double *__restrict__ V, ResArgs a,
                                                                    double *__restrict__ Vout,
                                                                    double *__restrict__ partial) {
  extern __shared__ double sm[];
  const int R = a.R;
  double *Ks = sm;                 // [R][RS_TM]
  double *Ws = sm + R * RS_TM;     // [R][RS_TD]
  __shared__ double red[32];
  const int tid = threadIdx.x;
  const int ml = tid % RS_TM;      // row within the tile
  const int dg = tid / RS_TM;      // 0..1 : which half of the d-chunk
  const int64_t m = (int64_t)blockIdx.x * RS_TM + ml;
  const int64_t slast = a.lens[a.N - 1];
  const double *wl = a.w[a.N - 1];

  // Khatri-Rao rows
  {
    int64_t idx[16];
    int64_t q = m < a.P1 ? m : 0;
    for (int j = 0; j < a.N - 1; j++) {
      idx[j] = q % a.lens[j];
      q /= a.lens[j];
    }
    for (int r = dg; r < R; r += RS_THREADS / RS_TM) {
      double v = 1.0;
      for (int j = 0; j < a.N - 1; j++) v *= a.w[j][idx[j] + a.lens[j] * r];
      Ks[r * RS_TM + ml] = (m < a.P1) ? v : 0.0;
    }
  }
  double ss = 0.0;
  for (int64_t d0 = 0; d0 < slast; d0 += RS_TD) {
    __syncthreads();
    for (int idx = tid; idx < R * RS_TD; idx += RS_THREADS) {
      const int r = idx / RS_TD, dd = idx % RS_TD;
      Ws[r * RS_TD + dd] = (d0 + dd < slast) ? wl[d0 + dd + slast * r] : 0.0;
    }
    __syncthreads();
    // this thread: row ml, d = d0 + dg*16 + 0..15 in groups of 4
#pragma unroll
    for (int grp = 0; grp < 4; grp++) {
      const int db = dg * 16 + grp * 4;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      for (int r = 0; r < R; r++) {
        const double k = Ks[r * RS_TM + ml];
        const double2 w01 = *reinterpret_cast<const double2 *>(&Ws[r * RS_TD + db]);
        const double2 w23 = *reinterpret_cast<const double2 *>(&Ws[r * RS_TD + db + 2]);
        a0 += k * w01.x;
        a1 += k * w01.y;
        a2 += k * w23.x;
        a3 += k * w23.y;
      }
      if (m < a.P1) {
        const double est[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int64_t d = d0 + db + e;
          if (d < slast) {
            const int64_t off = m + a.P1 * d;
            if (WRITE) {
              Vout[off] = est[e];
            } else {
              const double df = V[off] - est[e];
              ss += df * df;
            }
          }
        }
      }
    }
  }
  if (!WRITE) {
    ss = ppx_block_sum(ss, red);
    if (tid == 0) partial[blockIdx.x] = ss;
  }
}

}  // namespace

int ppx_sum_partials(ppx_ctx *ctx, const double *partial, int n, double *out);

int ppx_k7_init(ppx_ctx *ctx) {
  const int big = 200 * 1024;
  PPX_CUDA(ctx, cudaFuncSetAttribute(cp_reconstruct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(cp_reconstruct_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  return PPX_OK;
}

static int fill_args(ppx_ctx *ctx, ResArgs &a, const int64_t *lens, int N, const double *const *W, int R) {
  PPX_REQUIRE(ctx, lens && W && N >= 2 && N <= 16 && R >= 1, "2 <= N <= 16, R >= 1");
  a.N = N;
  a.R = R;
  a.P1 = 1;
  for (int j = 0; j < N; j++) {
    a.w[j] = W[j];
    a.lens[j] = lens[j];
    if (j < N - 1) a.P1 *= lens[j];
  }
  const size_t smem = sizeof(double) * (size_t)R * (RS_TM + RS_TD);
  if (smem > 200 * 1024) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "cp_residual: R=%d too large", R);
  return PPX_OK;
}

extern "C" {

int ppx_cp_residual(ppx_ctx *ctx, const double *V, const int64_t *lens, int N, const double *const *W, int R,
                    double *sq_out_dev) {
  PPX_REQUIRE(ctx, V && sq_out_dev, "V, sq_out_dev non-null");
  ResArgs a;
  int rc = fill_args(ctx, a, lens, N, W, R);
  if (rc) return rc;
  const int64_t blocks = (a.P1 + RS_TM - 1) / RS_TM;
  if (blocks > 0x7fffffffLL) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "cp_residual: grid too large");
  ppx_ws_reset(ctx);
  double *partial = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)blocks);
  if (!partial) return ppx_set_err(ctx, PPX_ENOMEM, "cp_residual needs %lld bytes of workspace", (long long)blocks * 8);
  const size_t smem = sizeof(double) * (size_t)R * (RS_TM + RS_TD);
  cp_reconstruct_kernel<false><<<(unsigned)blocks, RS_THREADS, smem, ctx->stream>>>(V, a, nullptr, partial);
  PPX_CHECK_LAUNCH(ctx);
  return ppx_sum_partials(ctx, partial, (int)blocks, sq_out_dev);
}

int ppx_cp_reconstruct(ppx_ctx *ctx, const int64_t *lens, int N, const double *const *W, int R, double *V_out) {
  PPX_REQUIRE(ctx, V_out, "V_out non-null");
  ResArgs a;
  int rc = fill_args(ctx, a, lens, N, W, R);
  if (rc) return rc;
  const int64_t blocks = (a.P1 + RS_TM - 1) / RS_TM;
  if (blocks > 0x7fffffffLL) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "cp_reconstruct: grid too large");
  const size_t smem = sizeof(double) * (size_t)R * (RS_TM + RS_TD);
  cp_reconstruct_kernel<true><<<(unsigned)blocks, RS_THREADS, smem, ctx->stream>>>(nullptr, a, V_out, nullptr);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

}  // extern "C"
