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
// owns 128 rows m.  Per tile it puts into shared memory, in DMMA operand layout (leading dimension = 4 mod 16:
// conflict-free fragment loads), the Khatri-Rao rows K of the tile and as many rows of W_last as fit (all 300 at
// R = 50: 178 KB together) -- ONE barrier per tile (per super-chunk of W_last when the last mode is long).  After it
// the warps run free: warp (mg, nh) takes rows 32 mg .. 32 mg + 31 and every other group of 32 columns, computes
// 32 x 32 with mma.sync m8n8k4 f64 -- accumulators in the very fragment layout whose V elements it needs: the 32 values
// of V a lane compares against are requested BEFORE the 13 k-steps of the group, so their latency hides behind them --
// and squares the differences in registers while other warps are in their MMA loops (a first version with W_last in
// double-buffered 64-column chunks put two barriers around every chunk: all warps computed, then all warps compared;
// ncu showed the DMMA pipe 52 % busy).  V is read exactly once, in 64-byte pieces (8 consecutive rows per column).
constexpr int RD_TM = 128;
constexpr int RD_THREADS = 256;

__device__ __forceinline__ uint32_t rd_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rd_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rd_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void rd_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rd_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rd_mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(rd_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// one bulk (TMA) copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void rd_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   rd_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(rd_smem_u32(bar))
               : "memory");
}
// volatile: stays where it is written, BEFORE the (volatile) DMMA stream whose latency it is meant to hide behind
__device__ __forceinline__ double rd_ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// W_last in DMMA B-operand layout, once per call: Bp[n * ld + k] = W_last[n, k] (zero for k >= R and for the rows that
// pad the last mode to a multiple of 8), so that a tile fetches its rows with ONE bulk copy
__global__ void __launch_bounds__(256) rd_pack_last_kernel(const double *__restrict__ wl, int64_t slast, int R, int ld,
                                                           int64_t nrows_pad, double *__restrict__ Bp) {
  const int64_t total = nrows_pad * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / ld;
    const int k = (int)(i - n * ld);
    Bp[i] = (k < R && n < slast) ? wl[n + slast * k] : 0.0;
  }
}
constexpr int RD_NQ_MAX = 34;  // distinct "other modes" rows of a tile: 128 / lens[0] + 2 with lens[0] >= 4

template <bool WRITE>
__global__ void __launch_bounds__(RD_THREADS, 1) cp_reconstruct_dmma_kernel(const double *__restrict__ V, ResArgs a,
                                                                            int KP, int ld, int ncap,
                                                                            const double *__restrict__ Bp,
                                                                            double *__restrict__ Vout,
                                                                            double *__restrict__ partial) {
  extern __shared__ double sm[];
  double *As = sm;                          // [RD_TM][ld]
  double *Bs = sm + RD_TM * ld;             // [ncap][ld]: rows n0s .. n0s + ncap - 1 of W_last
  double *Kr = Bs + (size_t)ncap * ld;      // [RD_NQ_MAX][KP]
  __shared__ double red[32];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int mg = warp & 3, nh = warp >> 2;
  const int R = a.R;
  const int64_t m0 = (int64_t)blockIdx.x * RD_TM;
  const int64_t slast = a.lens[a.N - 1];
  const int64_t slast_pad = (slast + 7) & ~(int64_t)7;
  if (tid == 0) rd_mbar_init(&bar, 1);
  {
    // Khatri-Rao rows of the tile, K[m, k] = W_0[i_0(m), k] * Krest[q(m), k], q = m / lens[0] (the other N-2 modes):
    // a 128-row tile touches nq <= 128 / lens[0] + 2 consecutive q, so Krest is formed once per tile, then every
    // thread forms 26 entries of K from independent, coalesced loads of W_0 (one thread per row walking k with a
    // dependent product chain per entry took ~9 us per tile, more than the tile's DMMA work).
    const int64_t s0 = a.lens[0];
    const int64_t q_first = m0 / s0;
    const int64_t m_last = (m0 + RD_TM <= a.P1 ? m0 + RD_TM : a.P1) - 1;
    const int nq = (int)(m_last / s0 - q_first) + 1;
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
    __syncthreads();  // (also publishes the mbarrier)
    if (tid == 0) {   // the first super-chunk of W_last arrives while K is being formed
      const int64_t nr = slast_pad < ncap ? slast_pad : ncap;
      const uint32_t bytes = (uint32_t)(nr * ld * sizeof(double));
      rd_mbar_expect_tx(&bar, bytes);
      rd_bulk_load(Bs, Bp, bytes, &bar);
    }
    const int ml = tid & (RD_TM - 1), half = tid >> 7;
    const int64_t m = m0 + ml;
    const bool live = m < a.P1;
    const int64_t q = (live ? m : m0) / s0;
    const int64_t i0 = (live ? m : m0) - q * s0;
    const double *kr = Kr + (int)(q - q_first) * KP;
    const double *w0 = a.w[0] + i0;
    for (int kb = half; kb < KP; kb += 32) {  // 16 independent loads in flight per thread and batch
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; u++) {
        const int k = kb + 2 * u;
        v[u] = (live && k < R) ? __ldg(w0 + s0 * k) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 16; u++) {
        const int k = kb + 2 * u;
        if (k < KP) As[ml * ld + k] = k < R ? v[u] * kr[k] : 0.0;
      }
    }
  }
  double ss = 0.0;
  const int ksteps = KP >> 2;
  const double *Arow = As + (32 * mg + g) * ld + t4;
  int sc = 0;
  for (int64_t n0s = 0; n0s < slast; n0s += ncap, sc++) {  // super-chunks of W_last (one at BASELINE sizes)
    const int nrows_pad = (int)(slast_pad - n0s < ncap ? slast_pad - n0s : ncap);
    if (n0s > 0) {
      __syncthreads();  // everyone is done with the previous rows of W_last
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)((size_t)nrows_pad * ld * sizeof(double));
        rd_mbar_expect_tx(&bar, bytes);
        rd_bulk_load(Bs, Bp + n0s * ld, bytes, &bar);
      }
    }
    rd_mbar_wait(&bar, sc & 1);
    if (n0s == 0) __syncthreads();  // K is complete
    const int nblk_all = nrows_pad >> 3;
    for (int b0 = 4 * nh; b0 < nblk_all; b0 += 8) {  // groups of 4 column blocks, alternating between the warp columns
      const int cnt = nblk_all - b0 < 4 ? nblk_all - b0 : 4;
      const int64_t n0 = n0s + 8 * b0;
      double vreg[4][4][2];
      if (!WRITE) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int64_t m = m0 + 32 * mg + 8 * i + g;
              const int64_t n = n0 + 8 * j + 2 * t4 + e;
              vreg[i][j][e] = (j < cnt && m < a.P1 && n < slast) ? rd_ld_stream(V + m + a.P1 * n) : 0.0;
            }
      }
      double acc[4][4][2];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
      const double *Brow = Bs + (8 * b0 + g) * ld + t4;
      if (cnt == 4) {
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ks++) {
          double af[4], bf[4];
#pragma unroll
          for (int i = 0; i < 4; i++) af[i] = Arow[(8 * i) * ld + 4 * ks];
#pragma unroll
          for (int j = 0; j < 4; j++) bf[j] = Brow[(8 * j) * ld + 4 * ks];
#pragma unroll
          for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) ppx_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      } else {
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
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            if (WRITE) {
              const int64_t m = m0 + 32 * mg + 8 * i + g;
              const int64_t n = n0 + 8 * j + 2 * t4 + e;
              if (j < cnt && m < a.P1 && n < slast) Vout[m + a.P1 * n] = acc[i][j][e];
            } else {
              // positions outside the tensor: vreg = 0 and acc = 0 (zero Khatri-Rao row / zero row of W_last)
              const double df = vreg[i][j][e] - (j < cnt ? acc[i][j][e] : 0.0);
              ss = fma(df, df, ss);
            }
          }
    }
  }
  if (!WRITE) {
    ss = ppx_block_sum(ss, red);
    if (tid == 0) partial[blockIdx.x] = ss;
  }
}

// leading dimension of the shared operands: >= KP and = 4 (mod 16), so that the 32 lanes of a fragment load (8 rows x
// 4 consecutive k) fall into 32 different 8-byte banks per half-warp
inline int rd_ld(int KP) { return KP + ((4 - KP % 16 + 16) % 16); }
// rows of W_last kept in shared memory at a time: all of them when they fit beside K and Krest in 220 KB
inline int rd_ncap(int KP, int ld, int64_t slast) {
  const int64_t budget = 220 * 1024 / 8 - (int64_t)RD_TM * ld - (int64_t)RD_NQ_MAX * KP;
  int64_t cap = budget / ld;
  cap &= ~(int64_t)63;  // whole 64-column pairs of warp groups
  const int64_t need = (slast + 7) & ~(int64_t)7;
  if (cap > need) cap = need;
  return (int)cap;
}
inline size_t rd_smem(int KP, int ld, int ncap) {
  return sizeof(double) * ((size_t)(RD_TM + ncap) * ld + (size_t)RD_NQ_MAX * KP);
}

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

// packs W_last into the workspace (after whatever the caller allocated there); NULL when the workspace is too small
static double *rd_pack_last(ppx_ctx *ctx, const ResArgs &a, int ld) {
  const int64_t slast = a.lens[a.N - 1], nrows_pad = (slast + 7) & ~(int64_t)7;
  double *Bp = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)nrows_pad * ld);
  if (!Bp) return nullptr;
  int blocks = (int)((nrows_pad * ld + 255) / 256);
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  rd_pack_last_kernel<<<blocks, 256, 0, ctx->stream>>>(a.w[a.N - 1], slast, a.R, ld, nrows_pad, Bp);
  ctx->launches++;
  return Bp;
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
  const int ncap = rd_ncap(KP, ld, a.lens[N - 1]);
  // tensor-pipe kernel whenever its operands fit in shared memory and the tensor is big enough to care
  double *Bp = nullptr;
  if (ncap >= 8 && a.P1 * a.lens[N - 1] >= (1 << 16) && a.lens[0] >= 4 && !getenv("PPX_K7_DFMA"))
    Bp = rd_pack_last(ctx, a, ld);
  if (Bp) {
    cp_reconstruct_dmma_kernel<false><<<(unsigned)blocks, RD_THREADS, rd_smem(KP, ld, ncap), ctx->stream>>>(
        V, a, KP, ld, ncap, Bp, nullptr, partial);
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
  const int ncap = rd_ncap(KP, ld, a.lens[N - 1]);
  double *Bp = nullptr;
  if (ncap >= 8 && a.P1 * a.lens[N - 1] >= (1 << 16) && a.lens[0] >= 4 && !getenv("PPX_K7_DFMA")) {
    ppx_ws_reset(ctx);
    Bp = rd_pack_last(ctx, a, ld);
  }
  if (Bp) {
    cp_reconstruct_dmma_kernel<true><<<(unsigned)blocks, RD_THREADS, rd_smem(KP, ld, ncap), ctx->stream>>>(
        nullptr, a, KP, ld, ncap, Bp, V_out, nullptr);
  } else {
    const size_t smem = sizeof(double) * (size_t)R * (RS_TM + RS_TD);
    cp_reconstruct_kernel<true><<<(unsigned)blocks, RS_THREADS, smem, ctx->stream>>>(nullptr, a, V_out, nullptr);
  }
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

}  // extern "C"
