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
// DFMA out of shared memory (measured 10 TFLOP/s at BASELINE configs[1]: 78 ms per residual).  Here the grid is
// persistent, one CTA of 12 warps per SM.  W_last lives in shared memory for the whole kernel in DMMA operand layout
// (leading dimension = 4 mod 16: conflict-free fragment loads; all 300 rows at R = 50, super-chunks when the last mode
// is long), fetched once per CTA by a bulk copy.  After that one barrier the warps run on their own: a warp takes a
// contiguous range of STRIPS of 16 rows m, keeps the DMMA A fragments of a strip (its Khatri-Rao rows, 2 x 13 doubles
// per lane at R = 50, formed from independent loads of the factors; the part that is constant along the strip range
// cached in per-lane shared slots) in REGISTERS, and walks the columns RD_NB blocks (8 RD_NB columns) at a time: 2 RD_NB
// independent accumulators, B fragments from the resident W_last, then the 64-byte pieces of V that match its
// accumulator fragments (8 consecutive rows of a column), squared differences in registers.  V is read exactly once.
// History at BASELINE configs[1] (78 ms for the DFMA kernel), each step measured on a B200:
//   60 ms    128-row tile per CTA, K in shared memory built by one thread per row walking k (a dependent product chain)
//   45.7 ms  W_last in double-buffered 64-column chunks, two barriers around every chunk (ncu: DMMA pipe 52 % busy)
//   43.2 ms  W_last resident, one bulk copy per tile, V prefetched into the accumulator layout before the MMA loop
//   42.5 / 44.6 / 42.8 ms  16 warps as 4 x 4, then 8 x 2 balanced + persistent, then without the prefetch whose register
//            pressure made ptxas schedule the two k-steps of one accumulator back to back
//   ablation (PPX_K7_DBG): without the V loads 41.3 ms, without the MMA loop 18.6, without both 11.8 -- the per-tile work
//            outside the MMA (two barriers, 64-bit index arithmetic, the shared-memory image of K) was 8 us of a 30 us tile
//            and nothing overlapped it, because a tile-wide barrier keeps all warps in the same phase
//   40.4 ms  strips with K in registers, no barrier (strided strips, Krest recomputed per strip)
//   41.2 ms  contiguous strip ranges per warp with Krest cached (this version): ablation 7.7 ms without V loads and MMA,
//            14.6 without the MMA, 32.3 without the V loads -- the three parts add up to the total: with two warps per
//            scheduler partition (255 registers each) a warp's address/load phase is not covered by the other warp's
//            MMA phase often enough.  The MMA loop itself now runs at 93 % of the pipe (24.6 ms for 22.9 ms of DMMA).
//   29.6 ms  the same kernel with 12 warps per CTA and 5 column blocks per pass (this version): three warps per
//            scheduler at <= 168 registers -- 2 x 5 accumulators per warp instead of 2 x 13, so a third warp fits and
//            one warp's address/V-load phase is covered by the other two.  Grid of the A/B (threads_blocks, ms;
//            profiles/r02w_k7_warps_ab.json): 256_13 41.4, 320_4 35.7, 384_3 30.5, 384_4 29.8, 384_5 29.6, 384_6 32.2
//            (spills), 448_3 33.2, 512_3 30.2, 512_4 32.3.  27.3 TFLOP/s = 0.77 of the DGEMM peak.
#ifndef PPX_RD_THREADS
#define PPX_RD_THREADS 384
#endif
constexpr int RD_THREADS = PPX_RD_THREADS;  // 12 independent warps per CTA (-DPPX_RD_THREADS: A/B builds)

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

#ifndef PPX_RD_NB
#define PPX_RD_NB 5
#endif
constexpr int RD_NB = PPX_RD_NB;  // column blocks per pass: 2 x RD_NB independent DMMA accumulators per warp

template <bool WRITE, int NKS>
__global__ void __launch_bounds__(RD_THREADS, 1) cp_reconstruct_dmma_kernel(const double *__restrict__ V, ResArgs a,
                                                                            int KP, int ld, int ncap,
                                                                            const double *__restrict__ Bp,
                                                                            int64_t nstrips, int dbg,
                                                                            double *__restrict__ Vout,
                                                                            double *__restrict__ partial) {
  extern __shared__ double sm[];
  double *Bs = sm;                       // [ncap][ld]: rows n0s .. n0s + ncap - 1 of W_last
  double *Ks = sm + (size_t)ncap * ld;   // [NKS][RD_THREADS]: per-lane Krest slots
  __shared__ double red[32];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int R = a.R;
  const int ksteps = KP >> 2;
  const int64_t slast = a.lens[a.N - 1];
  const int64_t slast_pad = (slast + 7) & ~(int64_t)7;
  const int64_t s0 = a.lens[0];
  const int64_t gw = (int64_t)blockIdx.x * (RD_THREADS / 32) + warp, nw = (int64_t)gridDim.x * (RD_THREADS / 32);
  if (tid == 0) rd_mbar_init(&bar, 1);
  __syncthreads();
  double ss4[4] = {0.0, 0.0, 0.0, 0.0};  // four chains: a single running sum would be 156 dependent DFMAs per strip
  int nload = 0;
  for (int64_t n0s = 0; n0s < slast; n0s += ncap) {  // super-chunks of W_last: ONE, loaded once per CTA, at BASELINE sizes
    const int nrows_pad = (int)(slast_pad - n0s < ncap ? slast_pad - n0s : ncap);
    if (n0s > 0) __syncthreads();  // everyone is done with the previous rows of W_last
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)((size_t)nrows_pad * ld * sizeof(double));
      rd_mbar_expect_tx(&bar, bytes);
      rd_bulk_load(Bs, Bp + n0s * ld, bytes, &bar);
    }
    rd_mbar_wait(&bar, nload & 1);
    nload++;
    const int nblk_all = nrows_pad >> 3;
    // from here to the end of the super-chunk every warp runs on its own: a CONTIGUOUS range of 16-row strips, no CTA
    // barrier.  Contiguous, so that consecutive strips share the indices of modes 1 .. N-2: their product Krest[k] is
    // recomputed (two divisions per mode and 13 x (N-2) loads) only when a strip starts a new value of m / lens[0];
    // in between a strip costs 26 independent loads of W_0 and 26 multiplications.
    const int64_t per_warp = (nstrips + nw - 1) / nw;
    const int64_t strip_b = gw * per_warp, strip_e = strip_b + per_warp < nstrips ? strip_b + per_warp : nstrips;
    // Krest of the current value of m / lens[0] lives in shared memory, one private slot per lane and k-step (13 more
    // doubles per lane do not fit beside 26 accumulators, 26 A fragments and the B fragments)
    double *krs = Ks + (size_t)tid;  // slot of k-step ks at krs[ks * RD_THREADS]
    int64_t q_cached_a = -1;
    auto krest = [&](int64_t q, double *kr) {
#pragma unroll
      for (int ks = 0; ks < NKS; ks++) kr[ks] = 1.0;
      for (int j = 1; j < a.N - 1; j++) {
        const int64_t ij = q % a.lens[j];
        q /= a.lens[j];
        const double *wj = a.w[j] + ij;
#pragma unroll
        for (int ks = 0; ks < NKS; ks++) {
          const int k = 4 * ks + t4;
          if (ks < ksteps && k < R) kr[ks] *= __ldg(wj + a.lens[j] * k);
        }
      }
    };
    for (int64_t strip = strip_b; strip < strip_e; strip++) {
      const int64_t m0 = strip * 16;
      // ---- this lane's A fragments, in registers for the whole strip:  areg[i][ks] = K[m0 + 8 i + g, 4 ks + t4],
      // K[m, k] = W_0[i_0(m), k] * Krest[m / lens[0], k]
      double areg[2][NKS];
      {
        const int64_t ma = m0 + g, mb = m0 + 8 + g;
        const bool la = ma < a.P1, lb = mb < a.P1;
        // m / lens[0] without a division in the common case: the previous strip's quotient, advanced when the row passes
        // the end of its mode-0 fibre
        int64_t qa, qb;
        if (q_cached_a >= 0 && ma - q_cached_a * s0 < s0) qa = q_cached_a;
        else qa = (la ? ma : 0) / s0;
        if (mb - qa * s0 < s0) qb = qa;
        else qb = (lb ? mb : 0) / s0;
        const int64_t ia = (la ? ma : 0) - qa * s0, ib = (lb ? mb : 0) - qb * s0;
        const double *wa = a.w[0] + ia, *wb = a.w[0] + ib;
        double wva[NKS], wvb[NKS];
#pragma unroll
        for (int ks = 0; ks < NKS; ks++) {
          const int k = 4 * ks + t4;
          wva[ks] = (ks < ksteps && k < R && la) ? __ldg(wa + s0 * k) : 0.0;
          wvb[ks] = (ks < ksteps && k < R && lb) ? __ldg(wb + s0 * k) : 0.0;
        }
        // (the choice is per lane: the 8 rows of a block may straddle a boundary)
        if (qa != q_cached_a) {
          double kr[NKS];
          krest(qa, kr);
#pragma unroll
          for (int ks = 0; ks < NKS; ks++) krs[ks * RD_THREADS] = kr[ks];
          q_cached_a = qa;
        }
#pragma unroll
        for (int ks = 0; ks < NKS; ks++) areg[0][ks] = wva[ks] * krs[ks * RD_THREADS];
        if (qb == qa) {
#pragma unroll
          for (int ks = 0; ks < NKS; ks++) areg[1][ks] = wvb[ks] * krs[ks * RD_THREADS];
        } else {  // rare: the second row block has passed the end of the fibre
          double kr[NKS];
          krest(qb, kr);
#pragma unroll
          for (int ks = 0; ks < NKS; ks++) areg[1][ks] = wvb[ks] * kr[ks];
        }
      }
      for (int b0 = 0; b0 < nblk_all; b0 += RD_NB) {
        const int cnt = nblk_all - b0 < RD_NB ? nblk_all - b0 : RD_NB;
        const int64_t n0 = n0s + 8 * b0;
        double acc[2][RD_NB][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int j = 0; j < RD_NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        const double *Brow = Bs + (8 * b0 + g) * ld + t4;
        if (!(dbg & 2)) {
          if (cnt == RD_NB) {
#pragma unroll
            for (int ks = 0; ks < NKS; ks++) {
              if (ks < ksteps) {
                double bf[RD_NB];
#pragma unroll
                for (int j = 0; j < RD_NB; j++) bf[j] = Brow[(8 * j) * ld + 4 * ks];
#pragma unroll
                for (int j = 0; j < RD_NB; j++) {
                  ppx_dmma(acc[0][j][0], acc[0][j][1], areg[0][ks], bf[j]);
                  ppx_dmma(acc[1][j][0], acc[1][j][1], areg[1][ks], bf[j]);
                }
              }
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < NKS; ks++) {
              if (ks < ksteps) {
#pragma unroll
                for (int j = 0; j < RD_NB; j++)
                  if (j < cnt) {
                    const double bfj = Brow[(8 * j) * ld + 4 * ks];
                    ppx_dmma(acc[0][j][0], acc[0][j][1], areg[0][ks], bfj);
                    ppx_dmma(acc[1][j][0], acc[1][j][1], areg[1][ks], bfj);
                  }
              }
            }
          }
        }
        // epilogue: V in the accumulators' own layout, 16 values per batch (64-byte pieces: 8 consecutive rows of a column)
#pragma unroll
        for (int j0 = 0; j0 < RD_NB; j0 += 4) {
          double v[2][4][2];
          if (!WRITE) {
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
              for (int j = 0; j < 4; j++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                  const int64_t m = m0 + 8 * i + g;
                  const int64_t n = n0 + 8 * (j0 + j) + 2 * t4 + e;
                  v[i][j][e] = (j0 + j < cnt && m < a.P1 && n < slast && !(dbg & 1)) ? rd_ld_stream(V + m + a.P1 * n) : 0.0;
                }
          }
#pragma unroll
          for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
              for (int e = 0; e < 2; e++) {
                if (j0 + j < RD_NB) {
                  if (WRITE) {
                    const int64_t m = m0 + 8 * i + g;
                    const int64_t n = n0 + 8 * (j0 + j) + 2 * t4 + e;
                    if (j0 + j < cnt && m < a.P1 && n < slast) Vout[m + a.P1 * n] = acc[i][j0 + j][e];
                  } else {
                    // positions outside the tensor: v = 0 and acc = 0 (zero Khatri-Rao row / zero row of W_last)
                    const double df = v[i][j][e] - (j0 + j < cnt ? acc[i][j0 + j][e] : 0.0);
                    ss4[2 * i + e] = fma(df, df, ss4[2 * i + e]);
                  }
                }
              }
        }
      }
    }
  }
  if (!WRITE) {
    double ss = (ss4[0] + ss4[1]) + (ss4[2] + ss4[3]);
    ss = ppx_block_sum(ss, red);
    if (tid == 0) partial[blockIdx.x] = ss;
  }
}

// leading dimension of the shared operands: >= KP and = 4 (mod 16), so that the 32 lanes of a fragment load (8 rows x
// 4 consecutive k) fall into 32 different 8-byte banks per half-warp
inline int rd_ld(int KP) { return KP + ((4 - KP % 16 + 16) % 16); }
// rows of W_last kept in shared memory at a time: all of them when they fit in 172 KB
inline int rd_ncap(int ld, int64_t slast) {
  int64_t cap = (172 * 1024 / 8) / ld;  // + 13 x 256 doubles of Krest slots = 26 KB
  cap &= ~(int64_t)7;
  const int64_t need = (slast + 7) & ~(int64_t)7;
  if (cap > need) cap = need;
  return (int)cap;
}
inline size_t rd_smem(int ld, int ncap) { return sizeof(double) * ((size_t)ncap * ld + 13 * (size_t)RD_THREADS); }

}  // namespace

int ppx_sum_partials(ppx_ctx *ctx, const double *partial, int n, double *out);

int ppx_k7_init(ppx_ctx *ctx) {
  const int big = 200 * 1024;
  PPX_CUDA(ctx, cudaFuncSetAttribute(cp_reconstruct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(cp_reconstruct_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  const int dmma_max = 204 * 1024;
#define PPX_K7_ATTR(W, NKS) \
  PPX_CUDA(ctx, cudaFuncSetAttribute(cp_reconstruct_dmma_kernel<W, NKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dmma_max))
  PPX_K7_ATTR(true, 3);
  PPX_K7_ATTR(true, 7);
  PPX_K7_ATTR(true, 13);
  PPX_K7_ATTR(false, 3);
  PPX_K7_ATTR(false, 7);
  PPX_K7_ATTR(false, 13);
#undef PPX_K7_ATTR
  return PPX_OK;
}

// The strip kernel is used where the FP64 pipe is the limit: 24 <= R <= 52 (k-steps held in registers: 7 or 13 of
// them), a tensor big enough to care and a leading extent that keeps a 16-row strip within two values of the other
// modes' indices.  At R = 10 the residual is bandwidth bound and the DFMA kernel above is the faster one (order 6,
// s = 40: 18.6 ms = 1.76 TB/s against 22.6 ms; measured, profiles/).  PPX_K7_FORCE_DMMA / PPX_K7_DFMA override.
static bool rd_eligible(const ResArgs &a) {
  if (getenv("PPX_K7_DFMA")) return false;
  const bool shape_ok = a.R <= 52 && a.P1 * a.lens[a.N - 1] >= (1 << 16) && a.lens[0] >= 16;
  return shape_ok && (a.R >= 24 || getenv("PPX_K7_FORCE_DMMA"));
}
template <bool WRITE>
static void rd_launch(ppx_ctx *ctx, int grid, size_t smem, const double *V, const ResArgs &a, int KP, int ld, int ncap,
                      const double *Bp, int64_t nstrips, int dbg, double *Vout, double *partial) {
  const int ks = KP >> 2;
  if (ks <= 3)
    cp_reconstruct_dmma_kernel<WRITE, 3><<<grid, RD_THREADS, smem, ctx->stream>>>(V, a, KP, ld, ncap, Bp, nstrips, dbg, Vout, partial);
  else if (ks <= 7)
    cp_reconstruct_dmma_kernel<WRITE, 7><<<grid, RD_THREADS, smem, ctx->stream>>>(V, a, KP, ld, ncap, Bp, nstrips, dbg, Vout, partial);
  else
    cp_reconstruct_dmma_kernel<WRITE, 13><<<grid, RD_THREADS, smem, ctx->stream>>>(V, a, KP, ld, ncap, Bp, nstrips, dbg, Vout, partial);
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
  const int ncap = rd_ncap(ld, a.lens[N - 1]);
  double *Bp = nullptr;
  if (ncap >= 8 && rd_eligible(a)) Bp = rd_pack_last(ctx, a, ld);
  if (Bp) {
    const int64_t nstrips = (a.P1 + 15) / 16;
    const int64_t want = (nstrips + RD_THREADS / 32 - 1) / (RD_THREADS / 32);
    const int grid = (int)(want < ctx->sm_count ? want : ctx->sm_count);  // persistent; partial[] has >= grid entries
    const int dbg = getenv("PPX_K7_DBG") ? atoi(getenv("PPX_K7_DBG")) : 0;  // timing experiments: 1 no V loads, 2 no MMA
    rd_launch<false>(ctx, grid, rd_smem(ld, ncap), V, a, KP, ld, ncap, Bp, nstrips, dbg, nullptr, partial);
    PPX_CHECK_LAUNCH(ctx);
    return ppx_sum_partials(ctx, partial, grid, sq_out_dev);
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
  const int ncap = rd_ncap(ld, a.lens[N - 1]);
  double *Bp = nullptr;
  if (ncap >= 8 && rd_eligible(a)) {
    ppx_ws_reset(ctx);
    Bp = rd_pack_last(ctx, a, ld);
  }
  if (Bp) {
    const int64_t nstrips = (a.P1 + 15) / 16;
    const int64_t want = (nstrips + RD_THREADS / 32 - 1) / (RD_THREADS / 32);
    const int grid = (int)(want < ctx->sm_count ? want : ctx->sm_count);
    rd_launch<true>(ctx, grid, rd_smem(ld, ncap), nullptr, a, KP, ld, ncap, Bp, nstrips, 0, V_out, nullptr);
  } else {
    const size_t smem = sizeof(double) * (size_t)R * (RS_TM + RS_TD);
    cp_reconstruct_kernel<true><<<(unsigned)blocks, RS_THREADS, smem, ctx->stream>>>(nullptr, a, V_out, nullptr);
  }
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

}  // extern "C"
