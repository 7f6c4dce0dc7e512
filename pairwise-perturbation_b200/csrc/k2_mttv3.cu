// K2x3 -- the three Hadamard contractions of one level-1 tensor in ONE pass (PP operator build, als_CP.cxx:352-409)
//
//   T[l, x, t, r]  (s_l x s_x x s_t x R, first index fastest)   ->   out_l[x, t, r] = sum_l T[l,x,t,r] Wl[l,r]
//                                                                    out_x[l, t, r] = sum_x T[l,x,t,r] Wx[x,r]
//                                                                    out_t[l, x, r] = sum_t T[l,x,t,r] Wt[t,r]
// (any subset: a null output is skipped).  At BASELINE configs[1] the level-1 tensor `d` (10.8 GB) has three consumers
// and `c` two; one ppx_mttv per consumer reads them five times at 1.6-1.75 ms each.  This kernel reads each once.
//
// The tensor is streamed through shared memory by bulk (TMA) copies: for a fixed (t, r) the rows of an x-tile are ONE
// contiguous run of s_l * XT doubles, so a stage is a single cp.async.bulk of up to 56 KB issued by one elected thread
// of a producer warp; three stages are in flight per CTA.  Fourteen consumer warps read the staged tile:
//   out_t, out_x : a thread is a row l (and, for short rows, one of NG groups of the tile's x): ONE shared-memory read
//           of T[l, x, t] feeds both -- the accumulator of (l, x) over t in a register, and the sum over the thread's x,
//           which is complete in the thread: a PARTIAL sum over x-tiles (and groups), written to scratch and added in
//           a fixed order by a second kernel (deterministic);
//   out_l : warp <-> x, lanes stride l, one shuffle reduction per (x, t).
//           (ncu, profiles/r02n_ncu_summary.json: DRAM read = 1.000 x the tensor, issue slots 60 % active, shared-memory
//           load wavefronts 43 % of peak -- the consumers' instruction stream bounds the kernel, not DRAM.)
// (First version: every thread owned 28 arbitrary (l, x) pairs for out_t and out_x was a second loop over 300 of the 256
// threads -- 3.7 ms per pass at configs[1], bound by the slowest warp's instruction stream, against 1.7 per output for
// the separate kernels.)
// Work items (x-tile, r) are handed out through an atomic counter (persistent CTAs, one per SM): 14 x 50 = 700 items of
// 15 MB at configs[1].  Which CTA takes an item does not change any sum.
// Eligibility (else the separate kernels): s_l even (16-byte runs) and at most 320, enough workspace for the out_x partials.
#include "ppx_internal.h"

namespace {

constexpr int M3_CONSUMERS = 320;              // 10 row warps (thread = row l)
#ifndef PPX_M3_COLW
#define PPX_M3_COLW 4
#endif
constexpr int M3_COLW = PPX_M3_COLW;                     // column warps (warp = column x): the sum over l, beside the row warps
constexpr int M3_THREADS = M3_CONSUMERS + 32 * M3_COLW + 32;  // + the producer warp (the last one)
#ifndef PPX_M3_NST
#define PPX_M3_NST 3
#endif
#ifndef PPX_M3_STAGE
#define PPX_M3_STAGE 7168
#endif
constexpr int M3_NST = PPX_M3_NST;             // stages
constexpr int M3_STAGE_DOUBLES = PPX_M3_STAGE; // 56 KB per stage
constexpr int M3_CK = 28;                      // x of a tile per consumer thread (out_t accumulators)
constexpr int M3_WX = M3_CK * (M3_CONSUMERS / 32);  // padded length of the tile's Wx column in shared memory

__device__ __forceinline__ uint32_t m3_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m3_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m3_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void m3_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m3_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m3_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(m3_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void m3_mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(m3_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void m3_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   m3_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(m3_smem_u32(bar))
               : "memory");
}

struct M3Args {
  const double *T;
  const double *Wl, *Wx, *Wt;
  long long ldl, ldx, ldt;
  long long sl, sx, st;
  int R, XT, ntiles;
  int dbg;                         // PPX_M3_DBG: 1 = no row loop, 2 = no column loop (timing aid)
  int LT, NG;                      // consumer thread = (l = tid % LT, x group = tid / LT); LT = s_l rounded up to a warp
  double *out_l, *part_x, *out_t;  // part_x: [ntiles * NG][sl * st * R]
  int *counter;
};

__global__ void __launch_bounds__(M3_THREADS, 1) mttv3_kernel(M3Args a) {
  extern __shared__ __align__(128) double sm[];
  double *stage = sm;                                     // [NST][STAGE_DOUBLES]
  double *wx = stage + (size_t)M3_NST * M3_STAGE_DOUBLES;  // [M3_WX]: the tile's x, zero beyond it (16-byte aligned)
  double *wt = wx + M3_WX;                                // [st]
  __shared__ uint64_t full[M3_NST], empty[M3_NST];
  __shared__ int item_slot[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nitems = a.ntiles * a.R;
  if (tid == 0) {
    for (int s = 0; s < M3_NST; s++) {
      m3_mbar_init(&full[s], 1);
      m3_mbar_init(&empty[s], M3_CONSUMERS / 32 + M3_COLW);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long plane = a.sl * a.sx;  // one t
  const long long cube = plane * a.st;  // one r

  if (warp == M3_CONSUMERS / 32 + M3_COLW) {
    // ---- producer: one elected lane ----
    if (lane == 0) {
      long long g = 0;  // stage counter over the whole kernel
      int nit = 0;
      for (;;) {
        const int w = atomicAdd(a.counter, 1);
        // the item id travels with the first stage of the item: written before that stage's arrive (release)
        if (w >= nitems) {
          const int s = (int)(g % M3_NST);
          if (g >= M3_NST) m3_mbar_wait(&empty[s], (uint32_t)(((g / M3_NST) - 1) & 1));
          item_slot[nit & 1] = -1;
          m3_mbar_arrive(&full[s]);
          break;
        }
        const int tile = w % a.ntiles, r = w / a.ntiles;
        const long long x0 = (long long)tile * a.XT;
        const int xc = (int)(a.sx - x0 < a.XT ? a.sx - x0 : a.XT);
        const uint32_t bytes = (uint32_t)(a.sl * xc * sizeof(double));
        const double *src = a.T + cube * r + a.sl * x0;
        for (long long t = 0; t < a.st; t++, g++) {
          const int s = (int)(g % M3_NST);
          if (g >= M3_NST) m3_mbar_wait(&empty[s], (uint32_t)(((g / M3_NST) - 1) & 1));
          if (t == 0) item_slot[nit & 1] = w;
          m3_mbar_expect_tx(&full[s], bytes);
          m3_bulk_load(stage + (size_t)s * M3_STAGE_DOUBLES, src + plane * t, bytes, &full[s]);
        }
        nit++;
      }
    }
    return;
  }

  if (warp >= M3_CONSUMERS / 32) {
    // ---- column warps: out_l[x, t] = sum_l T[l, x, t] Wl[l], warp = column, lanes stride the rows (ten per lane,
    // s_l <= 320, Wl in registers).  Three columns go together: thirty loads in flight, three FMA chains, one
    // interleaved shuffle reduction.  They run BESIDE the row warps on the same staged tile (in the first versions
    // every warp did its rows, then its columns: 2.48 ms for three outputs against 1.80 for the rows alone).
    const int cw = warp - M3_CONSUMERS / 32;
    const int sl = (int)a.sl;
    long long g = 0;
    for (int nit = 0;; nit++) {
      {
        const int s = (int)(g % M3_NST);
        m3_mbar_wait(&full[s], (uint32_t)((g / M3_NST) & 1));
      }
      const int w = item_slot[nit & 1];
      if (w < 0) break;
      const int tile = w % a.ntiles, r = w / a.ntiles;
      const long long x0 = (long long)tile * a.XT;
      const int xc = (int)(a.sx - x0 < a.XT ? a.sx - x0 : a.XT);
      double wlr[10];
#pragma unroll
      for (int q = 0; q < 10; q++) wlr[q] = (a.out_l && lane + 32 * q < sl) ? a.Wl[lane + 32 * q + a.ldl * r] : 0.0;
      double *pl = a.out_l ? a.out_l + (size_t)r * (size_t)(a.sx * a.st) + x0 : nullptr;
      for (long long t = 0; t < a.st; t++, g++) {
        const int s = (int)(g % M3_NST);
        if (t > 0) m3_mbar_wait(&full[s], (uint32_t)((g / M3_NST) & 1));
        const double *S = stage + (size_t)s * M3_STAGE_DOUBLES;
        if (pl && !(a.dbg & 2)) {
          for (int xb = 0; xb < xc; xb += 3 * M3_COLW) {
            double d[3] = {0.0, 0.0, 0.0};
            double cv[3][10];
#pragma unroll
            for (int j = 0; j < 3; j++) {
              const int xx = xb + cw + M3_COLW * j;
              const double *col = S + sl * (xx < xc ? xx : 0);
#pragma unroll
              for (int q = 0; q < 10; q++) {
                const int i = lane + 32 * q;
                cv[j][q] = col[i < sl ? i : sl - 1];
              }
            }
#pragma unroll
            for (int q = 0; q < 10; q++)
#pragma unroll
              for (int j = 0; j < 3; j++) d[j] = fma(cv[j][q], wlr[q], d[j]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
              for (int j = 0; j < 3; j++) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
#pragma unroll
            for (int j = 0; j < 3; j++) {
              const int xx = xb + cw + M3_COLW * j;
              if (lane == 0 && xx < xc) pl[xx + a.sx * t] = d[j];
            }
          }
        }
        __syncwarp();
        if (lane == 0) m3_mbar_arrive(&empty[s]);
      }
    }
    return;
  }

  // ---- row warps ----
  long long g = 0;
  for (int nit = 0;; nit++) {
    {  // the first stage of the item carries its id
      const int s = (int)(g % M3_NST);
      m3_mbar_wait(&full[s], (uint32_t)((g / M3_NST) & 1));
    }
    const int w = item_slot[nit & 1];
    if (w < 0) break;
    const int tile = w % a.ntiles, r = w / a.ntiles;
    const long long x0 = (long long)tile * a.XT;
    const int xc = (int)(a.sx - x0 < a.XT ? a.sx - x0 : a.XT);
    const int sl = (int)a.sl;
    const int l = tid % a.LT, xg = tid / a.LT;
    const bool lrow = l < sl && xg < a.NG;
    // factor columns of this item (warps may be in different items: a barrier before the columns are overwritten)
    asm volatile("bar.sync 1, %0;" ::"n"(M3_CONSUMERS) : "memory");
    if (a.out_t)
      for (int i = tid; i < a.st; i += M3_CONSUMERS) wt[i] = a.Wt[i + a.ldt * r];
    for (int i = tid; i < M3_WX; i += M3_CONSUMERS) wx[i] = (a.part_x && i < xc) ? a.Wx[x0 + i + a.ldx * r] : 0.0;
    asm volatile("bar.sync 1, %0;" ::"n"(M3_CONSUMERS) : "memory");
    double acc[M3_CK];
#pragma unroll
    for (int k = 0; k < M3_CK; k++) acc[k] = 0.0;
    double *px = a.part_x ? a.part_x + (((size_t)tile * a.NG + xg) * a.R + r) * (size_t)(a.sl * a.st) : nullptr;
    const bool want_t = a.out_t != nullptr;
    for (long long t = 0; t < a.st; t++, g++) {
      const int s = (int)(g % M3_NST);
      if (t > 0) m3_mbar_wait(&full[s], (uint32_t)((g / M3_NST) & 1));
      const double *S = stage + (size_t)s * M3_STAGE_DOUBLES;
      // out_t and out_x from ONE read of the element: thread = row l, its x of the tile in registers.  Batches of four
      // with the loads first and no branch inside (a guard per element made ptxas emit a branch per element and the
      // loop ran at one shared-memory latency per FMA: 2.95 ms per pass instead of 1.6); x beyond the tile read the
      // last valid element with zero weights.  kcnt is uniform in a warp (the threads of a warp share xg).
      if (lrow && (want_t || px) && !(a.dbg & 1)) {
        const double c = want_t ? wt[t] : 0.0;
        const int kcnt = xg < xc ? (xc - xg + a.NG - 1) / a.NG : 0;
        const double *Sl = S + l + sl * xg;
        const int step = sl * a.NG;
        double b0 = 0.0, b1 = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < M3_CK; k0 += 4) {
          if (k0 < kcnt) {
            double v[4], wv[4], cv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int k = k0 + u, kk = k < kcnt ? k : kcnt - 1;
              v[u] = Sl[step * kk];
              cv[u] = k < kcnt ? c : 0.0;
            }
            if (a.NG == 1) {  // adjacent x: two 16-byte loads (wx is zero beyond the tile)
              const double2 w01 = *reinterpret_cast<const double2 *>(wx + k0);
              const double2 w23 = *reinterpret_cast<const double2 *>(wx + k0 + 2);
              wv[0] = w01.x, wv[1] = w01.y, wv[2] = w23.x, wv[3] = w23.y;
            } else {
#pragma unroll
              for (int u = 0; u < 4; u++) wv[u] = wx[xg + a.NG * (k0 + u)];
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              acc[k0 + u] = fma(v[u], cv[u], acc[k0 + u]);
              if (u & 1) b1 = fma(v[u], wv[u], b1);
              else b0 = fma(v[u], wv[u], b0);
            }
          }
        }
        if (px) px[l + a.sl * t] = b0 + b1;
      }
      __syncwarp();
      if (lane == 0) m3_mbar_arrive(&empty[s]);
    }
    if (want_t && lrow) {
      double *po = a.out_t + (size_t)r * (size_t)plane + a.sl * x0 + l;
#pragma unroll
      for (int k = 0; k < M3_CK; k++) {
        const int xx = xg + a.NG * k;
        if (xx < xc) po[a.sl * xx] = acc[k];
      }
    }
  }
}

__global__ void __launch_bounds__(256) mttv3_reduce_kernel(const double *__restrict__ part, long long n, int nparts,
                                                           double *__restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double s = part[i];
  for (int k = 1; k < nparts; k++) s += part[(long long)k * n + i];
  out[i] = s;
}

}  // namespace

int ppx_k2x3_init(ppx_ctx *ctx) {
  PPX_CUDA(ctx, cudaFuncSetAttribute(mttv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  return PPX_OK;
}

extern "C" {

int ppx_mttv3(ppx_ctx *ctx, const double *T, const int64_t *lens3, int R, const double *Wl, int64_t ldl,
              const double *Wx, int64_t ldx, const double *Wt, int64_t ldt, double *out_l, double *out_x, double *out_t) {
  PPX_REQUIRE(ctx, T && lens3 && R >= 1, "T, lens3 non-null; R >= 1");
  PPX_REQUIRE(ctx, (!out_l || Wl) && (!out_x || Wx) && (!out_t || Wt), "a factor for every requested output");
  const int64_t sl = lens3[0], sx = lens3[1], st = lens3[2];
  PPX_REQUIRE(ctx, sl >= 1 && sx >= 1 && st >= 1, "extents >= 1");
  PPX_REQUIRE(ctx, (!out_l || ldl >= sl) && (!out_x || ldx >= sx) && (!out_t || ldt >= st), "leading dimensions");
  if (!out_l && !out_x && !out_t) return PPX_OK;
  static const bool off = getenv("PPX_NO_MTTV3") != nullptr;  // A/B: one ppx_mttv per output
  // one pass: rows of 16 bytes that fit the consumer threads, a stage that holds at least two x, scratch for out_x
  const int LT = (int)((sl + 31) / 32 * 32);
  const int NG = LT <= M3_CONSUMERS ? M3_CONSUMERS / LT : 0;
  int XT = (int)(M3_STAGE_DOUBLES / sl);
  if (XT > M3_CK * NG) XT = M3_CK * NG;
  if (XT > sx) XT = (int)sx;
  const size_t smem = sizeof(double) * ((size_t)M3_NST * M3_STAGE_DOUBLES + M3_WX + st);
  // rows of at least 192: below that the column pass wastes most of its ten loads per lane and the tile's x are spread
  // over row groups -- measured on a 38-row shard (8 GPUs) the one-pass kernel made the operator build 13.8 ms instead
  // of 11.9 with one kernel per output
  bool fused = !off && sl % 2 == 0 && sl >= 192 && NG >= 1 && XT >= 2 && smem <= 220 * 1024 && ((uintptr_t)T & 15) == 0 &&
               (out_l != nullptr) + (out_x != nullptr) + (out_t != nullptr) >= 2 && sl * sx * st * R >= (1 << 22) && st >= 2;
  int ntiles = 0, nparts = 0;
  double *part = nullptr;
  int *counter = nullptr;
  if (fused) {
    ntiles = (int)((sx + XT - 1) / XT);
    XT = (int)((sx + ntiles - 1) / ntiles);  // even tiles
    nparts = ntiles * NG;
    ppx_ws_reset(ctx);
    counter = (int *)ppx_ws_alloc(ctx, 256);
    if (out_x && nparts > 1) part = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)nparts * sl * st * R);
    if (!counter || (out_x && nparts > 1 && !part)) fused = false;
  }
  if (!fused) {
    // the separate kernels (same sums; out_x in a different order of additions)
    int rc = PPX_OK;
    const int64_t lens[3] = {sl, sx, st};
    if (out_l) rc = ppx_mttv(ctx, T, lens, 3, 0, Wl, ldl, R, out_l);
    if (!rc && out_x) rc = ppx_mttv(ctx, T, lens, 3, 1, Wx, ldx, R, out_x);
    if (!rc && out_t) rc = ppx_mttv(ctx, T, lens, 3, 2, Wt, ldt, R, out_t);
    return rc;
  }
  PPX_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
  M3Args a;
  a.T = T;
  a.Wl = Wl, a.Wx = Wx, a.Wt = Wt;
  a.ldl = ldl, a.ldx = ldx, a.ldt = ldt;
  a.sl = sl, a.sx = sx, a.st = st;
  a.R = R, a.XT = XT, a.ntiles = ntiles;
  a.LT = LT, a.NG = NG;
  a.dbg = getenv("PPX_M3_DBG") ? atoi(getenv("PPX_M3_DBG")) : 0;
  a.out_l = out_l;
  a.part_x = out_x ? (nparts > 1 ? part : out_x) : nullptr;
  a.out_t = out_t;
  a.counter = counter;
  const int nitems = ntiles * R;
  const int grid = nitems < ctx->sm_count ? nitems : ctx->sm_count;
  mttv3_kernel<<<grid, M3_THREADS, smem, ctx->stream>>>(a);
  PPX_CHECK_LAUNCH(ctx);
  if (out_x && nparts > 1) {
    const long long n = sl * st * (long long)R;
    mttv3_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(part, n, nparts, out_x);
    PPX_CHECK_LAUNCH(ctx);
  }
  return PPX_OK;
}

}  // extern "C"
