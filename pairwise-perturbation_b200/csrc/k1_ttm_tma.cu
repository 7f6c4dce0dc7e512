// K1, TMA path -- the first tensor-times-matrix contraction with TMA-staged tensor tiles (sm_100a).
//
//     out[l, t, r] = sum_k V[l, k, t] * B[k, r]           V viewed as L x K x Rt (first index fastest)
//
// B is one factor (ppx_ttm_first, Tucker ppx_ttm) or the Khatri-Rao rows of several adjacent factors
// (ppx_ttm_multi).  Same math and same DMMA (mma.sync m8n8k4 f64) consumer layout as k1_ttm_first.cu; what changes is
// how the operands reach shared memory:
//   * one elected thread issues cp.async.bulk.tensor (TMA) loads of V straight from a CUtensorMap and one
//     cp.async.bulk copy of the pre-packed B slab per stage (the duty rotates over the four warps, chunk by chunk);
//     completion is signalled on a per-stage "full" mbarrier (complete_tx), the warps hand a stage back through an
//     "empty" mbarrier -- no __syncthreads in the main loop and no per-thread load/address instructions.  (A separate
//     producer warp was tried first: a fifth warp puts three warps on one SM sub-partition and caps the kernel at 168
//     registers, which spills the accumulators.)
//   * the V tile is stored dense with the hardware 128-byte swizzle; the k index each lane takes in an MMA step is
//     permuted so that every DMMA fragment load is bank-conflict free on the swizzled tile;
//   * B is packed once per call (krp_pack_kernel) into per-chunk slabs [chunk][8*NT + TAIL columns][20] with the same k
//     permutation, zero padded in k and in the columns, so a stage's slab is ONE contiguous bulk copy.
// Layouts:  KMAJOR (L == 1): tensor map 2-D {K, M}, one 16 x 128 box per stage;
//           M-major (L > 1): 128 l of one t per row tile -- one 4-D box of a {16, K, L/16, Rt} view for every row tile
//                            made of whole groups of 16 l; the last row tile of a t when L % 16 != 0 (a mode-0 shard
//                            of 37 or 38 rows gives L = 11100 / 11400) takes eight 16(l) x 16(k) boxes of the 3-D map
//                            {L, K, Rt}, whose out-of-bounds rows read as zero;
//           M-major, short L (128-row tiles would idle > 6 %): 16 l x 8 consecutive t per row tile, one 3-D box.
// ppx_ttm_tma_try returns 1 (caller falls back to the cp.async kernel) when the shape is not TMA-friendly: odd
// extents (TMA needs 16-byte global strides), unaligned base, R > 64, or an L that would waste > 6 % of a row tile.
#include <cuda.h>
#include "ppx_internal.h"

// which layouts spell the register double buffering of the operand fragments out (see the consumer loop)
#ifndef PPX_K1_EXPLICIT_PREFETCH
#define PPX_K1_EXPLICIT_PREFETCH(kmajor) (true)  // k-major: 24.0 ms with, 24.9 ms without; M-major: 24.1 / 24.9
#endif

namespace {

constexpr int TBM = 128;          // rows per tile
constexpr int TBK = 16;           // depth per stage
constexpr int TLDW = 20;          // padded leading dimension of the packed B slab (20 mod 16 == 4: conflict free)
constexpr int TSTAGES = 4;
constexpr int TCONSUMERS = 128;   // 4 warps, all of them consume; they take turns issuing the TMA loads
constexpr int TTHREADS = 128;
constexpr int A_STAGE_BYTES = TBM * TBK * 8;  // 16 KB

struct TmaParams {
  const double *Wpp;  // packed slabs
  double *out;
  int64_t L, K, Rt, Mtot;
  int64_t split_stride;
  int R;
  int tiles_per_t;  // M-major: row tiles per t
  int num_tiles;
  int nk, ksplit, cps;
  int inplace, accumulate;
  int onebox;  // M-major: the eight 16 x 16 boxes of a stage are one 4-D box (see ppx_ttm_tma_try)
  int64_t L16; // onebox: rows [0, L16) are whole groups of 16 l; a row tile reaching past L16 uses the 3-D map instead
  int tmulti;  // M-major with a short L: a row tile is 16 l x 8 consecutive t (one 3-D box); tiles_per_t = l groups
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {  // non-blocking
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Position p = 4*q + t4 of a packed slab row  <->  k offset inside the 16-deep chunk taken by lane-in-group t4 in MMA
// step q.  The permutation makes the A-fragment loads conflict free on the 128-byte-swizzled tile:
//   M-major tile (rows = k, 16 l per 128-byte row): a half-warp reads 4 l x 4 k; with k = 8*(q>>1) + 2*t4 + (q&1)
//     the swizzle term (k & 7) = 2*t4 + (q&1) spreads the four lanes of a group over four different 16-byte chunks;
//   k-major tile (rows = m, 16 k per 128-byte row): with k = (t4&1) + 8*(t4>>1) + 2*q the four lanes of a group read
//     chunks q and q+4 (x two 8-byte halves), which the row swizzle (m & 7) keeps distinct across the four rows.
__host__ __device__ __forceinline__ int k_of_pos(int p, bool kmajor) {
  const int q = p >> 2, t4 = p & 3;
  return kmajor ? ((t4 & 1) + 8 * (t4 >> 1) + 2 * q) : (8 * (q >> 1) + 2 * t4 + (q & 1));
}

// Packs B (one factor, or the Khatri-Rao rows of n adjacent factors) into slabs: out[(c*ncols + n)*TLDW + p]
struct PackArgs {
  const double *w[8];
  int64_t x[8];
  int64_t ld[8];
  int n;
};
__global__ void __launch_bounds__(256) krp_pack_kernel(PackArgs a, int64_t K, int R, int ncols, int nk, int kmajor,
                                                       double *__restrict__ out) {
  const int64_t total = (int64_t)nk * ncols * 16;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int p = (int)(i & 15);
    const int64_t cn = i >> 4;
    const int n = (int)(cn % ncols);
    const int64_t c = cn / ncols;
    int64_t k = c * 16 + k_of_pos(p, kmajor != 0);
    double v = 0.0;
    if (k < K && n < R) {
      v = 1.0;
      for (int j = 0; j < a.n; j++) {
        const int64_t xj = k % a.x[j];
        k /= a.x[j];
        v *= a.w[j][xj + a.ld[j] * n];
      }
    }
    out[cn * TLDW + p] = v;
  }
}

// out[i] = sum_s part[s][i] in a fixed order.  Few splits: one thread per output.  Many splits (a fused contraction with
// few output rows, e.g. 200 x 10 outputs from 148 K slices at BASELINE configs[0]): the serial sum of one thread is a
// chain of dependent loads (measured 101 us for 2000 outputs) -- eight groups of lanes take the splits round-robin with
// four independent partial sums each and are combined through shared memory in a fixed order.
__global__ void __launch_bounds__(256) tma_split_reduce_kernel(const double *__restrict__ part, int64_t n, int nsplit,
                                            double *__restrict__ out) {
  if (nsplit < 16) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
      double s = part[i];
      for (int k = 1; k < nsplit; k++) s += part[(int64_t)k * n + i];
      out[i] = s;
    }
    return;
  }
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  for (int64_t i0 = (int64_t)blockIdx.x * 32; i0 < n; i0 += (int64_t)gridDim.x * 32) {
    const int64_t i = i0 + lane;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (i < n) {
      int k = g;
      for (; k + 24 < nsplit; k += 32) {
        a0 += part[(int64_t)k * n + i];
        a1 += part[(int64_t)(k + 8) * n + i];
        a2 += part[(int64_t)(k + 16) * n + i];
        a3 += part[(int64_t)(k + 24) * n + i];
      }
      for (; k < nsplit; k += 8) a0 += part[(int64_t)k * n + i];
    }
    red[g][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (g == 0 && i < n) {
      double s = red[0][lane];
#pragma unroll
      for (int q = 1; q < 8; q++) s += red[q][lane];
      out[i] = s;
    }
    __syncthreads();
  }
}

// NT full 8-column DMMA tiles + TAIL (0..4) extra columns.  R = 50 as 7 DMMA tiles wastes 6 of 56 columns of the FP64
// pipe; DMMA and DFMA share that pipe (tools/dmma_bench.cu: interleaving them adds their times), so the R mod 8 <= 4
// leftover columns are cheaper as plain DFMA on the A fragments the thread already holds: per 16-deep chunk 16*TAIL
// DFMA per thread (about 2.9*TAIL DMMA-equivalents) instead of the 16 DMMA of a padded tile.  Each lane accumulates
// the partial sum over its own k positions; the four lanes of a group are combined by shuffles in the epilogue.
template <int NT, int TAIL, bool KMAJOR>
__global__ void __launch_bounds__(TTHREADS, 2) ttm_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                              const __grid_constant__ CUtensorMap tmap3, TmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NCOLS = 8 * NT + TAIL;
  constexpr int W_STAGE_BYTES = NCOLS * TLDW * 8;
  // 1024-byte aligned base (the 128-byte swizzle pattern is a function of the shared address)
  // (offset arithmetic on the __shared__ array keeps the address space: LDS, not generic LD)
  uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *A_base = base;
  uint8_t *W_base = base + TSTAGES * A_STAGE_BYTES;
  uint64_t *full = (uint64_t *)(W_base + TSTAGES * W_STAGE_BYTES);
  uint64_t *empty = full + TSTAGES;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < TSTAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], TCONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int num_units = p.num_tiles * p.ksplit;
  struct Unit {
    int u, tile, split, count;
  };
  auto set_unit = [&](Unit &q, int u) {
    q.u = u;
    q.tile = 0;
    q.split = 0;
    q.count = 0;
    if (u >= num_units) return;
    if (p.ksplit == 1) {
      q.tile = u;
      q.count = p.nk;
    } else {
      q.tile = u / p.ksplit;
      q.split = u - q.tile * p.ksplit;
      const int kb = q.split * p.cps;
      const int ke = min(p.nk, kb + p.cps);
      q.count = ke > kb ? ke - kb : 0;
    }
  };
  int total = 0;
  if (p.ksplit == 1) {
    if ((int)blockIdx.x < num_units) total = ((num_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) * p.nk;
  } else {
    Unit q;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      set_unit(q, u);
      total += q.count;
    }
  }

  // ---- producer duty, rotated over the warps: chunk n of this CTA's flat chunk stream is issued by lane 0 of warp
  // n % 4 (one producer warp would make its scheduler partition -- shared with the producer warp of the SM's other
  // CTA -- the slowest of the four, and every other warp waits for the data it issues).  Each warp tracks the
  // position (unit, chunk inside the unit) of the chunk it issues next.
  Unit ld;
  set_unit(ld, blockIdx.x);
  int ld_kc = 0, ld_n = 0;  // ld/ld_kc describe chunk ld_n
  auto issue_chunk = [&](int n) {
    if (KMAJOR) {
      // single producer thread (tid 0), chunks in order: ld/ld_kc already describe chunk n
      if (tid != 0) return;
      while (ld.count == 0) set_unit(ld, ld.u + gridDim.x);
    } else {
      int d = n - ld_n;
      ld_n = n;
      while (true) {
        while (ld.count == 0) set_unit(ld, ld.u + gridDim.x);
        const int rem = ld.count - ld_kc;
        if (d < rem) {
          ld_kc += d;
          break;
        }
        d -= rem;
        ld_kc = 0;
        set_unit(ld, ld.u + gridDim.x);
      }
      if (lane != 0) return;
    }
    const int stage = n % TSTAGES;
    const int round = n / TSTAGES;
    if (round > 0) mbar_wait(&empty[stage], (round - 1) & 1);  // every warp has released the stage
    mbar_expect_tx(&full[stage], A_STAGE_BYTES + W_STAGE_BYTES);
    const int chunk = ld.split * p.cps + ld_kc;
    uint8_t *As = A_base + stage * A_STAGE_BYTES;
    if (KMAJOR) {
      tma_load_2d(As, &tmap, &full[stage], chunk * TBK, ld.tile * TBM);
    } else {
      const int t = ld.tile / p.tiles_per_t;
      const int l0 = (ld.tile - t * p.tiles_per_t) * TBM;
      if (p.tmulti) {
        tma_load_3d(As, &tmap, &full[stage], (ld.tile - t * p.tiles_per_t) * 16, chunk * TBK, t * 8);
      } else if (p.onebox && l0 + TBM <= p.L16) {
        tma_load_4d(As, &tmap, &full[stage], 0, chunk * TBK, l0 >> 4, t);
      } else {
#pragma unroll
        for (int b = 0; b < 8; b++) tma_load_3d(As + b * 2048, &tmap3, &full[stage], l0 + 16 * b, chunk * TBK, t);
      }
    }
    bulk_load(W_base + stage * W_STAGE_BYTES, p.Wpp + (int64_t)chunk * (NCOLS * TLDW), W_STAGE_BYTES, &full[stage]);
    if (KMAJOR) {  // advance to the next chunk of the stream
      if (++ld_kc == ld.count) {
        ld_kc = 0;
        set_unit(ld, ld.u + gridDim.x);
      }
    }
  };
  // (measured at order-4 s=300 R=50: rotation gains 3 % for the M-major layout, whose issue path is longer, and
  // loses 4 % for the k-major one -- 25.06 vs 24.02 ms -- so that one keeps thread 0 as its only producer, walking
  // the chunk stream in order)
  constexpr int ROT = KMAJOR ? 0 : 3;
  for (int n = 0; n < TSTAGES - 1 && n < total; n++)
    if (KMAJOR ? tid == 0 : warp == (n & ROT)) issue_chunk(n);

  // ===================================== consumer warps =====================================
  const int g = lane >> 2, t4 = lane & 3;
  // byte offsets of this lane's A fragments inside a stage (see k_of_pos)
  int aoff[KMAJOR ? 4 : 8];
  if (KMAJOR) {
    // row m = 32*warp + 8*i + g (i adds 1024 bytes), k = (t4&1) + 8*(t4>>1) + 2*q
#pragma unroll
    for (int q = 0; q < 4; q++)
      aoff[q] = (32 * warp + g) * 128 + ((((q + 4 * (t4 >> 1)) ^ g) << 4) | ((t4 & 1) << 3));
  } else {
    // l = 32*warp + 8*i + g -> box 2*warp + (i>>1), l' = 8*(i&1) + g; k = 8*(q>>1) + 2*t4 + (q&1)
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int lp = 8 * (i & 1) + g;
        const int kk = 2 * t4 + e;
        aoff[2 * i + e] = (2 * warp + (i >> 1)) * 2048 + kk * 128 + ((((lp >> 1) ^ kk) << 4) | ((lp & 1) << 3));
      }
  }

  double acc[4][NT][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  double tacc[4][TAIL > 0 ? TAIL : 1];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int c = 0; c < (TAIL > 0 ? TAIL : 1); c++) tacc[i][c] = 0.0;

  Unit cu;
  set_unit(cu, blockIdx.x);
  int kc = 0;
  // Operand fragments are double buffered in registers: the shared loads of MMA step q+1 are issued BEFORE the DMMAs of
  // step q, also across the chunk boundary (the next stage's full barrier is polled without blocking; its data is
  // normally there, three chunks were requested ahead).  ptxas found this overlap on its own only in some variants
  // (254 registers: 23.6 ms, 206 registers: 25.0 ms on the same k-major problem), so it is spelled out.
  struct Frag {
    double a[4], b[NT], bt[TAIL > 0 ? TAIL : 1];
  };
  auto load_frag = [&](Frag &f, int stage, int q) {
    const uint8_t *As = A_base + stage * A_STAGE_BYTES;
    const double *Ws = (const double *)(W_base + stage * W_STAGE_BYTES);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (KMAJOR)
        f.a[i] = *(const double *)(As + aoff[q] + i * 1024);
      else
        f.a[i] = *(const double *)(As + aoff[2 * i + (q & 1)] + (q >> 1) * 1024);
    }
#pragma unroll
    for (int j = 0; j < NT; j++) f.b[j] = Ws[(8 * j + g) * TLDW + 4 * q + t4];
#pragma unroll
    for (int cc = 0; cc < TAIL; cc++) f.bt[cc] = Ws[(8 * NT + cc) * TLDW + 4 * q + t4];
  };
  auto compute = [&](const Frag &f) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < NT; j++) ppx_dmma(acc[i][j][0], acc[i][j][1], f.a[i], f.b[j]);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int cc = 0; cc < TAIL; cc++) tacc[i][cc] = fma(f.a[i], f.bt[cc], tacc[i][cc]);
  };
  Frag f0, f1;
  bool have_f0 = false;  // f0 already holds step 0 of the chunk about to be processed
  for (int c = 0; c < total; c++) {
    while (cu.count == 0) set_unit(cu, cu.u + gridDim.x);
    const int stage = c % TSTAGES;
    if (PPX_K1_EXPLICIT_PREFETCH(KMAJOR)) {
      if (!have_f0) {
        mbar_wait(&full[stage], (c / TSTAGES) & 1);
        load_frag(f0, stage, 0);
      }
      load_frag(f1, stage, 1);
      compute(f0);
      load_frag(f0, stage, 2);
      compute(f1);
      load_frag(f1, stage, 3);
      compute(f0);
      have_f0 = false;
      if (c + 1 < total) {
        const int ns = (c + 1) % TSTAGES;
        have_f0 = __all_sync(0xffffffffu, mbar_test(&full[ns], ((c + 1) / TSTAGES) & 1));  // warp-uniform
        if (have_f0) load_frag(f0, ns, 0);
      }
      compute(f1);
    } else {
      mbar_wait(&full[stage], (c / TSTAGES) & 1);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        load_frag(f0, stage, q);
        compute(f0);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);  // this warp is done reading the stage
    if ((KMAJOR ? tid == 0 : warp == ((c + TSTAGES - 1) & ROT)) && c + TSTAGES - 1 < total)
      issue_chunk(c + TSTAGES - 1);  // refills the stage of chunk c-1
    __syncwarp();
    if (++kc == cu.count) {
      const int tile = cu.tile, split = cu.split;
      double *outp = p.out + (int64_t)split * p.split_stride;
      int64_t row0, tt = 0;
      int64_t rows_valid;
      if (KMAJOR) {
        row0 = (int64_t)tile * TBM;
        rows_valid = p.Mtot - row0;
      } else {
        tt = tile / p.tiles_per_t;
        row0 = (int64_t)(tile - tt * p.tiles_per_t) * TBM;  // l0
        rows_valid = p.L - row0;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int rl = 32 * warp + 8 * i + g;
        bool rv = rl < rows_valid;
        int64_t basei, cstride;
        if (!KMAJOR && p.tmulti) {
          // row rl of the tile = (l, t) = (16 * group + rl % 16, 8 * tt + rl / 16)
          const int64_t l = (int64_t)(tile - tt * p.tiles_per_t) * 16 + (rl & 15);
          const int64_t t = tt * 8 + (rl >> 4);
          rv = l < p.L && t < p.Rt;
          basei = p.inplace ? l + p.L * (int64_t)p.R * t : l + p.L * t;
          cstride = p.inplace ? p.L : p.Mtot;
        } else if (p.inplace) {
          // out[l + L*(col + R*t)]
          basei = KMAJOR ? (row0 + rl) * (int64_t)p.R : (row0 + rl) + p.L * (int64_t)p.R * tt;
          cstride = KMAJOR ? 1 : p.L;
        } else {
          // out[m + Mtot*col], m = l + L*t
          basei = KMAJOR ? (row0 + rl) : (row0 + rl) + p.L * tt;
          cstride = p.Mtot;
        }
#pragma unroll
        for (int j = 0; j < NT; j++) {
          const int col = 8 * j + 2 * t4;
          if (rv) {
            if (col < p.R) {
              double *o = outp + basei + cstride * col;
              *o = p.accumulate ? (*o + acc[i][j][0]) : acc[i][j][0];
            }
            if (col + 1 < p.R) {
              double *o = outp + basei + cstride * (col + 1);
              *o = p.accumulate ? (*o + acc[i][j][1]) : acc[i][j][1];
            }
          }
          acc[i][j][0] = acc[i][j][1] = 0.0;
        }
        if (TAIL > 0) {
          // combine the four lanes of the group (fixed order), lane t4 == c stores column 8*NT + c
          double mine = 0.0;
#pragma unroll
          for (int c = 0; c < TAIL; c++) {
            double v = tacc[i][c];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (t4 == c) mine = v;
            tacc[i][c] = 0.0;
          }
          if (rv && t4 < TAIL) {
            double *o = outp + basei + cstride * (8 * NT + t4);
            *o = p.accumulate ? (*o + mine) : mine;
          }
        }
      }
      kc = 0;
      set_unit(cu, cu.u + gridDim.x);
    }
  }
}

template <int NT, int TAIL>
constexpr size_t tma_smem() {
  return 1024 + (size_t)TSTAGES * (A_STAGE_BYTES + (8 * NT + TAIL) * TLDW * 8) + 2 * TSTAGES * sizeof(uint64_t);
}

constexpr int TAIL_MAX = 4;  // leftover columns done as DFMA; 5..7 leftover columns get a padded DMMA tile
constexpr int NT_MAX = 8;

// (NT, TAIL) table of the instantiated kernels: NT = 1..8 with TAIL = 0, NT = 1..7 with TAIL = 1..4
typedef void (*TmaKernel)(const CUtensorMap, const CUtensorMap, TmaParams);
struct TmaEntry {
  TmaKernel kern[2];  // [kmajor]
  size_t smem;
};
template <int NT, int TAIL>
constexpr TmaEntry tma_entry() {
  return TmaEntry{{ttm_tma_kernel<NT, TAIL, false>, ttm_tma_kernel<NT, TAIL, true>}, tma_smem<NT, TAIL>()};
}
#define PPX_TMA_ROW(NT) {tma_entry<NT, 0>(), tma_entry<NT, 1>(), tma_entry<NT, 2>(), tma_entry<NT, 3>(), tma_entry<NT, 4>()}
const TmaEntry g_tma_table[NT_MAX][TAIL_MAX + 1] = {PPX_TMA_ROW(1), PPX_TMA_ROW(2), PPX_TMA_ROW(3), PPX_TMA_ROW(4),
                                                    PPX_TMA_ROW(5), PPX_TMA_ROW(6), PPX_TMA_ROW(7),
                                                    {tma_entry<8, 0>(), {}, {}, {}, {}}};

// R <= 64 -> (NT full tiles, TAIL DFMA columns)
inline void tma_split_rank(int R, int *nt, int *tail) {
  int n = R / 8, t = R % 8;
  if (n == 0 || t > TAIL_MAX) {
    n += 1;
    t = 0;
  }
  *nt = n;
  *tail = t;
}

int launch_tma(ppx_ctx *ctx, const CUtensorMap &map, const CUtensorMap &map3, const TmaParams &p, int nt, int tail,
               bool kmajor) {
  const TmaEntry &e = g_tma_table[nt - 1][tail];
  const int units = p.num_tiles * p.ksplit;
  const int gx = units < 2 * ctx->sm_count ? units : 2 * ctx->sm_count;
  e.kern[kmajor ? 1 : 0]<<<gx, TTHREADS, e.smem, ctx->stream>>>(map, map3, p);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_encode_tried = false;

}  // namespace

int ppx_k1_tma_init(ppx_ctx *ctx) {
  for (int n = 0; n < NT_MAX; n++)
    for (int t = 0; t <= TAIL_MAX; t++)
      for (int k = 0; k < 2; k++) {
        const TmaEntry &e = g_tma_table[n][t];
        if (!e.kern[k]) continue;
        cudaError_t err = cudaFuncSetAttribute((const void *)e.kern[k], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)e.smem);
        if (err != cudaSuccess) return ppx_set_err(ctx, PPX_ECUDA, "k1 tma init: %s", cudaGetErrorString(err));
      }
  if (!g_encode_tried) {
    g_encode_tried = true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
    cudaGetLastError();
  }
  return PPX_OK;
}

// Returns PPX_OK if the contraction was enqueued on the TMA path, 1 if the shape is not eligible (the caller falls
// back to the cp.async kernel), a negative PPX_E* code on error.  `fac`/`ld`/`xs`: the n_fac adjacent factors whose
// Khatri-Rao rows form B (n_fac == 1: a single factor), K = prod xs.
int ppx_ttm_tma_try(ppx_ctx *ctx, const double *V, int64_t L, int64_t K, int64_t Rt, const double *const *fac,
                    const int64_t *ld, const int64_t *xs, int n_fac, int R, double *out, int inplace, int accumulate,
                    bool ws_keep) {
  if (!g_encode || getenv("PPX_NO_TMA")) return 1;
  const bool kmajor = (L == 1);
  const int64_t Mtot = L * Rt;
  if (R > 64 || Mtot == 0 || K < 1) return 1;
  // M-major tiles arrive as eight 16 x 16 boxes (the 128-byte swizzle limits the inner box extent to 16 doubles):
  // fine when the FP64 pipe is the limit (R >= 25), but for small R the kernel is HBM bound and the 2 KB boxes reach
  // only ~75 % of the copy bandwidth where the 16-byte cp.async path reaches 99 % (measured, profiles/) -> use that.
  if (!kmajor && R <= 24 && !getenv("PPX_FORCE_TMA")) return 1;
  if ((((uintptr_t)V) & 15) != 0) return 1;
  if (kmajor ? (K % 2 != 0) : (L % 2 != 0)) return 1;                   // 16-byte global strides
  if (K >= ((int64_t)1 << 31) || Rt >= ((int64_t)1 << 31) || L >= ((int64_t)1 << 31)) return 1;
  int64_t tiles_per_t = 1, num_tiles;
  bool tmulti = false;
  if (kmajor) {
    num_tiles = (Mtot + TBM - 1) / TBM;
  } else {
    tiles_per_t = (L + TBM - 1) / TBM;
    num_tiles = tiles_per_t * Rt;
    if ((double)(tiles_per_t * TBM - L) > 0.06 * (double)L) {
      // short L: too many idle rows in a 128 l x 1 t tile; try 16 l x 8 t tiles (e.g. L = 300: 19 groups, 1.3 % idle)
      const int64_t lg = (L + 15) / 16, tb = (Rt + 7) / 8;
      if ((double)(lg * 16 * tb * 8) > 1.06 * (double)Mtot) return 1;
      tmulti = true;
      tiles_per_t = lg;
      num_tiles = lg * tb;
    }
  }
  if (num_tiles > 0x7fffffff / 64) return 1;
  const int64_t nk64 = (K + TBK - 1) / TBK;
  if (nk64 > 0x7fffffff / 2) return 1;

  TmaParams p;
  p.out = out;
  p.L = L;
  p.K = K;
  p.Rt = Rt;
  p.Mtot = Mtot;
  p.R = R;
  p.tiles_per_t = (int)tiles_per_t;
  p.num_tiles = (int)num_tiles;
  p.nk = (int)nk64;
  p.ksplit = 1;
  p.cps = p.nk;
  p.split_stride = 0;
  p.inplace = inplace;
  p.accumulate = accumulate;
  p.tmulti = tmulti ? 1 : 0;
  // one 4-D box per stage for every row tile made of whole groups of 16 l (needs at least one such tile); with
  // L % 16 != 0 the last row tile of each t goes through the 3-D map (zero fill past L), everything else is unchanged
  p.onebox = (!kmajor && !tmulti && L >= TBM && !getenv("PPX_NO_ONEBOX")) ? 1 : 0;
  p.L16 = (L % 16 == 0) ? ((int64_t)1 << 62) : (L / 16) * 16;  // L % 16 == 0: every tile, the 4-D view zero-fills past L
  int nt, tail;
  tma_split_rank(R, &nt, &tail);
  const int ncols = 8 * nt + tail;

  if (!ws_keep) ppx_ws_reset(ctx);
  double *Wpp = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)p.nk * ncols * TLDW);
  if (!Wpp) return 1;  // workspace too small for the packed slabs: use the other path

  // K split when the row tiles cannot balance the persistent grid (same rule as the cp.async kernel)
  double *partial = nullptr;
  const int G = 2 * ctx->sm_count;
  if (!inplace && !accumulate && p.nk >= 64 && p.num_tiles < 8 * G) {
    const int best = ppx_pick_ksplit(p.num_tiles, p.nk, G, Mtot * (int64_t)R, ctx->ws_bytes - ctx->ws_used);
    if (best > 1) {
      partial = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)best * Mtot * R);
      if (partial) {
        p.ksplit = best;
        p.cps = (p.nk + best - 1) / best;
        p.split_stride = Mtot * (int64_t)R;
        p.out = partial;
      }
    }
  }

  // tensor map(s) of V
  CUtensorMap map, map3;
  CUresult cr;
  if (kmajor) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Mtot};
    cuuint64_t strides[1] = {(cuuint64_t)K * 8};
    cuuint32_t box[2] = {TBK, TBM};
    cuuint32_t es[2] = {1, 1};
    cr = g_encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)V, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (p.onebox) {
    // L = 16 * Lo: view V as {li = 16, k = K, lo = Lo, t = Rt}; a {16, 16, 8, 1} box lands in shared memory as
    // [lo][k][li] -- byte for byte the image of eight consecutive {16 l, 16 k} boxes, in ONE instruction per stage
    // (rows past the last full group of a partial row tile are out of bounds in `lo` and read as zero).
    cuuint64_t dims[4] = {16, (cuuint64_t)K, (cuuint64_t)(L / 16), (cuuint64_t)Rt};
    cuuint64_t strides[3] = {(cuuint64_t)L * 8, 128, (cuuint64_t)L * (cuuint64_t)K * 8};
    cuuint32_t box[4] = {16, TBK, 8, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (strides[2] >= ((cuuint64_t)1 << 40)) return 1;
    cr = g_encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)V, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr == CUDA_SUCCESS && L % 16 != 0) {  // the last row tile of each t
      cuuint64_t dims3[3] = {(cuuint64_t)L, (cuuint64_t)K, (cuuint64_t)Rt};
      cuuint64_t strides3[2] = {(cuuint64_t)L * 8, (cuuint64_t)L * (cuuint64_t)K * 8};
      cuuint32_t box3[3] = {16, TBK, 1};
      cuuint32_t es3[3] = {1, 1, 1};
      cr = g_encode(&map3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)V, dims3, strides3, box3, es3,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      map3 = map;
    }
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)L, (cuuint64_t)K, (cuuint64_t)Rt};
    cuuint64_t strides[2] = {(cuuint64_t)L * 8, (cuuint64_t)L * (cuuint64_t)K * 8};
    cuuint32_t box[3] = {16, TBK, (cuuint32_t)(tmulti ? 8 : 1)};
    cuuint32_t es[3] = {1, 1, 1};
    if (strides[1] >= ((cuuint64_t)1 << 40)) return 1;
    cr = g_encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)V, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (cr != CUDA_SUCCESS) return 1;
  if (!p.onebox) map3 = map;  // the 3-D (or 2-D) map is the only one

  // pack B
  PackArgs a;
  a.n = n_fac;
  for (int j = 0; j < n_fac; j++) {
    a.w[j] = fac[j];
    a.x[j] = xs[j];
    a.ld[j] = ld[j];
  }
  {
    const int64_t total = (int64_t)p.nk * ncols * 16;
    int blocks = ppx_cdiv(total, 256 * 2);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    if (blocks < 1) blocks = 1;
    krp_pack_kernel<<<blocks, 256, 0, ctx->stream>>>(a, K, R, ncols, p.nk, kmajor ? 1 : 0, Wpp);
    PPX_CHECK_LAUNCH(ctx);
  }
  p.Wpp = Wpp;
  int rc = launch_tma(ctx, map, map3, p, nt, tail, kmajor);
  if (rc) return rc;
  if (p.ksplit > 1) {
    const int64_t n = Mtot * (int64_t)R;
    int blocks = p.ksplit < 16 ? ppx_cdiv(n, 256 * 4) : ppx_cdiv(n, 32);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    tma_split_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(partial, n, p.ksplit, out);
    PPX_CHECK_LAUNCH(ctx);
  }
  return PPX_OK;
}
