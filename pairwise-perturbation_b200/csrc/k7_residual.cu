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

// ---- the same on the FP64 tensor pipe -----------------------------------------------------------------------------
// Vhat[m, d] = sum_r K[m, r] W_last[d, r] is a GEMM with a short inner dimension (R); the kernel above does it with
// DFMA out of shared memory (measured 10 TFLOP/s at BASELINE configs[1]: 78 ms per residual).  Here a CTA of 8 warps
// owns 128 rows m: it forms the Khatri-Rao rows K once in shared memory (DMMA A operand, leading dimension = 4 mod 16:
// conflict-free fragment loads), then sweeps the last mode in chunks of 64 columns: W_last's rows of the chunk are
// double buffered with cp.async (DMMA B operand, same layout), each warp computes 32 rows x up to 32 columns with
// mma.sync m8n8k4 f64 -- accumulators in the very fragment layout whose V elements it needs: the 32 values of V a lane
// compares against are requested BEFORE the MMA loop of the chunk, so their latency hides behind the 13 k-steps -- and
// squares the difference in registers.  V is read exactly once, in 64-byte pieces (8 consecutive rows per column).
constexpr int RD_TM = 128;
constexpr int RD_TN = 64;
constexpr int RD_THREADS = 256;

__device__ __forceinline__ void cp_async8(double *smem, const double *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem));
}

template <bool WRITE>
__global__ void __launch_bounds__(RD_THREADS, 1) cp_reconstruct_dmma_kernel(const double *__restrict__ V, ResArgs a,
                                                                            int KP, int ld, double *__restrict__ Vout,
                                                                            double *__restrict__ partial) {
  extern __shared__ double sm[];
  double *As = sm;               // [RD_TM][ld]
  double *Bs = sm + RD_TM * ld;  // [2][RD_TN][ld]
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int mg = warp & 3, nh = warp >> 2;
  const int R = a.R;
  const int64_t m0 = (int64_t)blockIdx.x * RD_TM;
  const int64_t slast = a.lens[a.N - 1];
  const double *wl = a.w[a.N - 1];
  const int nchunks = (int)((slast + RD_TN - 1) / RD_TN);

  auto load_B = [&](int c, int buf) {
    const int64_t n0 = (int64_t)c * RD_TN;
    double *dst = Bs + (size_t)buf * RD_TN * ld;
    for (int idx = tid; idx < KP * RD_TN; idx += RD_THREADS) {
      const int k = idx / RD_TN, n = idx - k * RD_TN;
      if (k < R && n0 + n < slast) cp_async8(dst + n * ld + k, wl + (n0 + n) + slast * k);
      else dst[n * ld + k] = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_B(0, 0);
  {
    // Khatri-Rao rows of the tile, K[m, k] = W_0[i_0(m), k] * Krest[q(m), k], q = m / lens[0] (the other N-2 modes):
    // a 128-row tile touches nq <= 128 / lens[0] + 2 consecutive q, so Krest is formed once per tile in shared memory
    // (the B buffer that is not being filled), then every thread forms 26 entries of K from independent, coalesced
    // loads of W_0 -- one thread per row walking k with a dependent product chain per entry took ~9 us per tile, more
    // than the tile's DMMA work.
    const int64_t s0 = a.lens[0];
    const int64_t q_first = m0 / s0;
    const int64_t m_last = (m0 + RD_TM <= a.P1 ? m0 + RD_TM : a.P1) - 1;
    const int nq = (int)(m_last / s0 - q_first) + 1;
    double *Kr = Bs + (size_t)RD_TN * ld;  // buffer 1: chunk 1 is loaded only after the first __syncthreads below
    for (int e = tid; e < nq * KP; e += RD_THREADS) {
      const int qi = e / KP, k = e - qi * KP;
      double v = 0.0;
      if (k < R) {
        v = 1.0;
        int64_t q = q_first + qi;
        for (int j = 1; j < a.N - 1; j++) {
          const int64_t ij = q % a.lens[j];
          q /= a.lens[j];
          v *= __ldg(a.w[j] + ij + a.lens[j] * k);
        }
      }
      Kr[qi * KP + k] = v;
    }
    __syncthreads();
    const int ml = tid & (RD_TM - 1), half = tid >> 7;
    const int64_t m = m0 + ml;
    const bool live = m < a.P1;
    const int64_t q = (live ? m : m0) / s0;
    const int64_t i0 = (live ? m : m0) - q * s0;
    const double *kr = Kr + (int)(q - q_first) * KP;
    const double *w0 = a.w[0] + i0;
    for (int kb = half; kb < KP; kb += 16) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = kb + 2 * u;
        v[u] = (live && k < R) ? __ldg(w0 + s0 * k) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = kb + 2 * u;
        if (k < KP) As[ml * ld + k] = k < R ? v[u] * kr[k] : 0.0;
      }
    }
    __syncthreads();  // Kr (in buffer 1) has been consumed before the main loop prefetches chunk 1 into it
  }
  double ss = 0.0;
  const int ksteps = KP >> 2;
  const double *Arow = As + (32 * mg + g) * ld + t4;
  for (int c = 0; c < nchunks; c++) {
    const int buf = c & 1;
    const int64_t n0 = (int64_t)c * RD_TN;
    if (c + 1 < nchunks) load_B(c + 1, buf ^ 1);
    // column blocks of this chunk, split between the two warp columns
    const int64_t left = slast - n0;
    const int nblk = left >= RD_TN ? 8 : (int)((left + 7) >> 3);
    const int h = (nblk + 1) >> 1;
    const int nbs = nh ? h : 0, cnt = nh ? nblk - h : h;
    double vreg[4][4][2];
    if (!WRITE) {
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int64_t m = m0 + 32 * mg + 8 * i + g;
            const int64_t n = n0 + 8 * (nbs + j) + 2 * t4 + e;
            vreg[i][j][e] = (j < cnt && m < a.P1 && n < slast) ? __ldg(V + m + a.P1 * n) : 0.0;
          }
    }
    if (c + 1 < nchunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // chunk c of W_last (and, for c == 0, the Khatri-Rao rows) are in shared memory
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double *Brow = Bs + (size_t)buf * RD_TN * ld + (8 * nbs + g) * ld + t4;
#pragma unroll 2
    for (int ks = 0; ks < ksteps; ks++) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; i++) af[i] = Arow[(8 * i) * ld + 4 * ks];
#pragma unroll
      for (int j = 0; j < 4; j++) bf[j] = j < cnt ? Brow[(8 * j) * ld + 4 * ks] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (j < cnt) {
#pragma unroll
          for (int i = 0; i < 4; i++) ppx_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          if (WRITE) {
            const int64_t m = m0 + 32 * mg + 8 * i + g;
            const int64_t n = n0 + 8 * (nbs + j) + 2 * t4 + e;
            if (j < cnt && m < a.P1 && n < slast) Vout[m + a.P1 * n] = acc[i][j][e];
          } else {
            // positions outside the tensor: vreg = 0 and acc = 0 (zero Khatri-Rao row / zero column of W_last)
            const double df = vreg[i][j][e] - (j < cnt ? acc[i][j][e] : 0.0);
            ss = fma(df, df, ss);
          }
        }
    __syncthreads();  // every warp is done with buffer `buf` before the prefetch of chunk c + 2 overwrites it
  }
  if (!WRITE) {
    ss = ppx_block_sum(ss, red);
    if (tid == 0) partial[blockIdx.x] = ss;
  }
}

// leading dimension of the shared operands: >= KP and = 4 (mod 16), so that the 32 lanes of a fragment load (8 rows x
// 4 consecutive k) fall into 32 different 8-byte banks per half-warp
inline int rd_ld(int KP) { return KP + ((4 - KP % 16 + 16) % 16); }
inline size_t rd_smem(int ld) { return sizeof(double) * (size_t)(RD_TM + 2 * RD_TN) * ld; }

}  // namespace

int ppx_sum_partials(ppx_ctx *ctx, const double *partial, int n, double *out);

int ppx_k7_init(ppx_ctx *ctx) {
  const int big = 200 * 1024;
  PPX_CUDA(ctx, cudaFuncSetAttribute(cp_reconstruct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(cp_reconstruct_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  const int dmma_max = 220 * 1024;
  PPX_CUDA(ctx, cudaFuncSetAttribute(cp_reconstruct_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dmma_max));
  PPX_CUDA(ctx, cudaFuncSetAttribute(cp_reconstruct_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dmma_max));
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
  const int KP = (R + 3) & ~3, ld = rd_ld(KP);
  // tensor-pipe kernel whenever its operands fit in shared memory and the tensor is big enough to care
  if (rd_smem(ld) <= 220 * 1024 && a.P1 * a.lens[N - 1] >= (1 << 16) && a.lens[0] >= 4 && !getenv("PPX_K7_DFMA")) {
    cp_reconstruct_dmma_kernel<false><<<(unsigned)blocks, RD_THREADS, rd_smem(ld), ctx->stream>>>(V, a, KP, ld, nullptr,
                                                                                              partial);
  } else {
    const size_t smem = sizeof(double) * (size_t)R * (RS_TM + RS_TD);
    cp_reconstruct_kernel<false><<<(unsigned)blocks, RS_THREADS, smem, ctx->stream>>>(V, a, nullptr, partial);
  }
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
  const int KP = (R + 3) & ~3, ld = rd_ld(KP);
  if (rd_smem(ld) <= 220 * 1024 && a.P1 * a.lens[N - 1] >= (1 << 16) && a.lens[0] >= 4 && !getenv("PPX_K7_DFMA")) {
    cp_reconstruct_dmma_kernel<true><<<(unsigned)blocks, RD_THREADS, rd_smem(ld), ctx->stream>>>(nullptr, a, KP, ld, V_out,
                                                                                             nullptr);
  } else {
    const size_t smem = sizeof(double) * (size_t)R * (RS_TM + RS_TD);
    cp_reconstruct_kernel<true><<<(unsigned)blocks, RS_THREADS, smem, ctx->stream>>>(nullptr, a, V_out, nullptr);
  }
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

}  // extern "C"
