// K1 -- first tensor-times-matrix contraction of the dimension tree
//     out[l, t, r] = sum_x V[l, x, t] * W[x, r]         (V viewed as L x X x Rt, first index fastest)
// replaces the CTF expressions at common.cxx:56, als_CP.cxx:378-379, cp_dt_optimizer.cxx:158-159,
// cp_msdt_optimizer.cxx:142-143 and common.cxx:963 of the reference; with `inplace` it is the Tucker TTM
// (als_Tucker.cxx:102,224,464-465).
//
// Design (sm_100a): FP64 has no tcgen05 kind, so the tensor pipe is reached through warp-level DMMA
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4, measured 37.1 TFLOP/s on B200 = the FP64 peak).  Persistent CTAs
// stream 128-row x 16-deep tiles of V and the matching 16 x R slab of W through a multi-stage cp.async pipeline
// that runs across work-unit boundaries (no drain between tiles); each of the 4 warps owns a 32 x (8*NT)
// accumulator block in registers.  Shared-memory tiles are padded (ld = 4 mod 16 doubles) so every DMMA fragment
// load is bank-conflict free.  Two CTAs per SM.
//   KMAJOR = true : L == 1, the contracted mode is the fastest one (rows of V are contiguous along x).
//   KMAJOR = false: L  > 1, rows m = (l, t) are contiguous along l for a fixed x.
// A work unit is (row tile, K split): when there are too few row tiles to balance 2 x 148 persistent CTAs (the
// fused contraction of several adjacent modes has K = prod of their sizes and few rows) the K range is split and the
// partial tiles are summed by a second kernel in a fixed order (deterministic).
//
// ppx_ttm_multi: several ADJACENT modes contracted at once.  sum_{x1..xn} V[l, x1..xn, t] prod_j W_j[x_j, r] is one
// GEMM with the Khatri-Rao product of the factors as right-hand side: K[(x1..xn), r] = prod_j W_j[x_j, r] is formed
// by a small kernel (K*R doubles, L2 resident) and the level-1 intermediates of the reference's tree
// (common.cxx:56 followed by :83) are never written to HBM.
#include "ppx_internal.h"

namespace {

constexpr int BM = 128;   // rows of the output tile per CTA
constexpr int BK = 16;    // depth of one pipeline stage
constexpr int LDK = 20;   // padded leading dimension of k-contiguous tiles   (20 mod 16 == 4)
constexpr int LDM = 132;  // padded leading dimension of m-contiguous tiles   (132 mod 16 == 4)
constexpr int THREADS = 128;

struct TtmParams {
  const double *V;
  const double *W;
  double *out;
  int64_t L, X, Rt, Mtot, ldw;
  int64_t split_stride;  // elements between the partial outputs of consecutive K splits
  int R;
  int num_tiles;
  int nk;          // 16-deep chunks along K
  int ksplit;      // number of K splits (1 = none)
  int cps;         // chunks per split
  int vecA, vecW;
  int inplace;     // 0: out[m + Mtot*col] (rank last, CP) ; 1: out[l + L*(col + R*t)] (rank replaces mode x, Tucker)
  int accumulate;  // out += result
};

template <bool KMAJOR>
__host__ __device__ constexpr int a_doubles() {
  return KMAJOR ? BM * LDK : BK * LDM;
}

template <int NT, bool KMAJOR, int STAGES>
__global__ void __launch_bounds__(THREADS, 2) ttm_first_kernel(TtmParams p) {
  extern __shared__ __align__(16) double smem[];
  constexpr int A_D = a_doubles<KMAJOR>();
  constexpr int W_D = 8 * NT * LDK;
  constexpr int STAGE_D = A_D + W_D;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int ncol0 = blockIdx.y * 64;

  const int num_units = p.num_tiles * p.ksplit;
  // unit u = (tile, split); chunks [split*cps, min(nk, (split+1)*cps))
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

  // ---- loader state -----------------------------------------------------------------------------------
  Unit ld;  // work unit of the next chunk to load
  set_unit(ld, blockIdx.x);
  int ld_kc = 0;
  int cached_tile = -1;
  int64_t off0 = 0, off1 = 0;  // M-major: global offsets (without the k term) of this thread's two rows
  int nb0 = 0, nb1 = 0;        // bytes valid for row 0 / row 1 (0 or 8)

  auto load_chunk = [&](int stage) {
    while (ld.count == 0) set_unit(ld, ld.u + gridDim.x);  // skip empty units (only possible for the last split)
    const int tile = ld.tile, split = ld.split;
    double *As = smem + stage * STAGE_D;
    double *Ws = As + A_D;
    const int64_t k0 = (int64_t)(split * p.cps + ld_kc) * BK;
    // Fast path (all but the K-tail chunk and the last row tile of aligned problems): full 16-byte copies from a
    // per-thread base pointer plus a constant stride -- a handful of integer instructions per cp.async instead of
    // the bounds / zero-fill logic of the general path below.
    if (p.vecA && p.vecW && k0 + BK <= p.X && (int64_t)(tile + 1) * BM <= p.Mtot) {
      if (KMAJOR) {
        const int kp = tid & 7, mb = tid >> 3;
        const double *src = p.V + ((int64_t)tile * BM + mb) * p.X + k0 + 2 * kp;
        const int64_t stride = 16 * p.X;
        double *dst = As + mb * LDK + 2 * kp;
#pragma unroll
        for (int i = 0; i < 8; i++) ppx_cp_async16(dst + i * 16 * LDK, src + i * stride, 16);
      } else {
        const int ml = 2 * (tid & 63), kb = tid >> 6;
        if (tile != cached_tile) {
          cached_tile = tile;
          const int64_t mg = (int64_t)tile * BM + ml;
          nb0 = nb1 = 8;
          const int64_t t0 = mg / p.L, l0 = mg - t0 * p.L;
          off0 = l0 + p.L * p.X * t0;
          off1 = off0 + 1;
        }
        const double *src = p.V + off0 + (k0 + kb) * p.L;
        const int64_t stride = 2 * p.L;
        double *dst = As + kb * LDM + ml;
#pragma unroll
        for (int i = 0; i < 8; i++) ppx_cp_async16(dst + i * 2 * LDM, src + i * stride, 16);
      }
      {
        const int n0 = tid >> 3, kp = tid & 7;  // slab rows n0, n0+16, ...: same kp, constant row stride
        const double *src = p.W + (int64_t)(ncol0 + n0) * p.ldw + k0 + 2 * kp;
        const int64_t stride = 16 * p.ldw;
        double *dst = Ws + n0 * LDK + 2 * kp;
#pragma unroll
        for (int i = 0; i < (8 * NT + 15) / 16; i++) {
          if (n0 + 16 * i < 8 * NT) {
            const bool cv = ncol0 + n0 + 16 * i < p.R;
            ppx_cp_async16(dst + i * 16 * LDK, cv ? src + i * stride : p.W, cv ? 16 : 0);
          }
        }
      }
    } else {
    if (KMAJOR) {
      const int kp = tid & 7, mb = tid >> 3;
      const int64_t kg = k0 + 2 * kp;
      const int kv = (kg + 1 < p.X) ? 2 : ((kg < p.X) ? 1 : 0);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int m = mb + 16 * i;
        const int64_t mg = (int64_t)tile * BM + m;
        const bool mv = mg < p.Mtot;
        const double *src = p.V + (mv ? (mg * p.X + kg) : 0);
        double *dst = As + m * LDK + 2 * kp;
        if (p.vecA) {
          ppx_cp_async16(dst, (mv && kv) ? src : p.V, !mv ? 0 : (kv == 2 ? 16 : (kv == 1 ? 8 : 0)));
        } else {
          ppx_cp_async8(dst, (mv && kv >= 1) ? src : p.V, (mv && kv >= 1) ? 8 : 0);
          ppx_cp_async8(dst + 1, (mv && kv == 2) ? src + 1 : p.V, (mv && kv == 2) ? 8 : 0);
        }
      }
    } else {
      const int ml = 2 * (tid & 63), kb = tid >> 6;
      if (tile != cached_tile) {
        cached_tile = tile;
        const int64_t mg = (int64_t)tile * BM + ml;
        nb0 = (mg < p.Mtot) ? 8 : 0;
        nb1 = (mg + 1 < p.Mtot) ? 8 : 0;
        const int64_t t0 = mg / p.L, l0 = mg - t0 * p.L;
        off0 = nb0 ? (l0 + p.L * p.X * t0) : 0;
        if (p.vecA) {
          off1 = off0 + 1;
        } else {
          const int64_t t1 = (mg + 1) / p.L, l1 = (mg + 1) - t1 * p.L;
          off1 = nb1 ? (l1 + p.L * p.X * t1) : 0;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int k = kb + 2 * i;
        const int64_t kg = k0 + k;
        const bool kvd = kg < p.X;
        double *dst = As + k * LDM + ml;
        if (p.vecA) {
          const int nb = kvd ? (nb0 + nb1) : 0;
          ppx_cp_async16(dst, nb ? (p.V + off0 + kg * p.L) : p.V, nb);
        } else {
          ppx_cp_async8(dst, (kvd && nb0) ? (p.V + off0 + kg * p.L) : p.V, kvd ? nb0 : 0);
          ppx_cp_async8(dst + 1, (kvd && nb1) ? (p.V + off1 + kg * p.L) : p.V, kvd ? nb1 : 0);
        }
      }
    }
    // W slab: Ws[n][k], n < 8*NT, k < BK
#pragma unroll
    for (int i = 0; i < (8 * NT * 8 + THREADS - 1) / THREADS; i++) {
      const int idx = tid + THREADS * i;
      if (idx < 8 * NT * 8) {
        const int n = idx >> 3, kp = idx & 7;
        const int col = ncol0 + n;
        const int64_t kg = k0 + 2 * kp;
        const int kv = (col < p.R) ? ((kg + 1 < p.X) ? 2 : ((kg < p.X) ? 1 : 0)) : 0;
        const double *src = p.W + (kv ? ((int64_t)col * p.ldw + kg) : 0);
        double *dst = Ws + n * LDK + 2 * kp;
        if (p.vecW) {
          // kv == 1: the last row of an odd-length factor slice whose leading dimension is even (rows b..e of a
          // replicated factor in a sharded run): 8 bytes are copied, the other 8 zero-filled.  (Until round 2 this
          // copied nothing, dropping the slice's last row: wrong Tucker cores on shards with an odd number of rows.)
          ppx_cp_async16(dst, src, kv == 2 ? 16 : (kv == 1 ? 8 : 0));
        } else {
          ppx_cp_async8(dst, src, kv >= 1 ? 8 : 0);
          ppx_cp_async8(dst + 1, kv == 2 ? src + 1 : p.W, kv == 2 ? 8 : 0);
        }
      }
    }
    }  // general path
    if (++ld_kc == ld.count) {
      ld_kc = 0;
      set_unit(ld, ld.u + gridDim.x);
    }
  };

  // ---- accumulators ---------------------------------------------------------------------------------------
  double acc[4][NT][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // ---- pipeline -------------------------------------------------------------------------------------------
#pragma unroll
  for (int c = 0; c < STAGES - 1; c++) {
    if (c < total) load_chunk(c);
    ppx_cp_async_commit();
  }

  Unit cu;  // work unit being computed
  set_unit(cu, blockIdx.x);
  int kc = 0;
  for (int c = 0; c < total; c++) {
    while (cu.count == 0) set_unit(cu, cu.u + gridDim.x);
    ppx_cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int cn = c + STAGES - 1;
      if (cn < total) load_chunk(cn % STAGES);
      ppx_cp_async_commit();
    }
    const double *As = smem + (c % STAGES) * STAGE_D;
    const double *Ws = As + A_D;
#pragma unroll
    for (int kk = 0; kk < BK / 4; kk++) {
      double a[4], b[NT];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (KMAJOR)
          a[i] = As[(32 * warp + 8 * i + g) * LDK + 4 * kk + t4];
        else
          a[i] = As[(4 * kk + t4) * LDM + 32 * warp + 8 * i + g];
      }
#pragma unroll
      for (int j = 0; j < NT; j++) b[j] = Ws[(8 * j + g) * LDK + 4 * kk + t4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) ppx_dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (++kc == cu.count) {
      // epilogue of this work unit
      const int tile = cu.tile, split = cu.split;
      double *outp = p.out + (int64_t)split * p.split_stride;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int64_t mg = (int64_t)tile * BM + 32 * warp + 8 * i + g;
        int64_t base = mg, cstride = p.Mtot;
        if (p.inplace) {
          const int64_t tt = mg / p.L, ll = mg - tt * p.L;
          base = ll + p.L * (int64_t)p.R * tt;
          cstride = p.L;
        }
#pragma unroll
        for (int j = 0; j < NT; j++) {
          const int col = ncol0 + 8 * j + 2 * t4;
          if (mg < p.Mtot) {
            if (col < p.R) {
              double *o = outp + base + cstride * col;
              *o = p.accumulate ? (*o + acc[i][j][0]) : acc[i][j][0];
            }
            if (col + 1 < p.R) {
              double *o = outp + base + cstride * (col + 1);
              *o = p.accumulate ? (*o + acc[i][j][1]) : acc[i][j][1];
            }
          }
          acc[i][j][0] = acc[i][j][1] = 0.0;
        }
      }
      kc = 0;
      set_unit(cu, cu.u + gridDim.x);
    }
  }
  ppx_cp_async_wait<0>();
}

// out[i] = sum_s part[s][i] in a fixed order.  Few splits: one thread per output.  Many splits (a fused contraction with
// few output rows, e.g. 200 x 10 outputs from 148 K slices at BASELINE configs[0]): the serial sum of one thread is a
// chain of dependent loads (measured 101 us for 2000 outputs) -- eight groups of lanes take the splits round-robin with
// four independent partial sums each and are combined through shared memory in a fixed order.
__global__ void __launch_bounds__(256) split_reduce_kernel(const double *__restrict__ part, int64_t n, int nsplit,
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

// Khatri-Rao rows of n adjacent modes: K[k, r] = prod_j W_j[idx_j(k), r], k = x_1 + X_1*(x_2 + X_2*(...))
struct KrpArgs {
  const double *w[8];
  int64_t x[8];
  int64_t ld[8];
  int n;
};
__global__ void __launch_bounds__(256) krp_kernel(KrpArgs a, int64_t K, int R, double *__restrict__ out) {
  const int64_t total = K * (int64_t)R;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int64_t r = i / K;
    int64_t k = i - r * K;
    double v = 1.0;
    for (int j = 0; j < a.n; j++) {
      const int64_t xj = k % a.x[j];
      k /= a.x[j];
      v *= a.w[j][xj + a.ld[j] * r];
    }
    out[i] = v;
  }
}

// pipeline depth: as deep as two CTAs per SM allow (about 113 KB of shared memory each)
template <int NT, bool KMAJOR>
constexpr int ttm_stages() {
  return KMAJOR ? (NT <= 4 ? 4 : 3) : (NT <= 3 ? 5 : 4);
}
template <int NT, bool KMAJOR>
constexpr size_t ttm_smem() {
  return sizeof(double) * ttm_stages<NT, KMAJOR>() * (a_doubles<KMAJOR>() + 8 * NT * LDK);
}

template <int NT, bool KMAJOR>
int launch_ttm(ppx_ctx *ctx, const TtmParams &p, int col_blocks) {
  constexpr int STAGES = ttm_stages<NT, KMAJOR>();
  auto kern = ttm_first_kernel<NT, KMAJOR, STAGES>;
  const int units = p.num_tiles * p.ksplit;
  int gx = units < 2 * ctx->sm_count ? units : 2 * ctx->sm_count;
  dim3 grid(gx, col_blocks);
  kern<<<grid, THREADS, ttm_smem<NT, KMAJOR>(), ctx->stream>>>(p);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

template <int NT, bool KMAJOR>
cudaError_t init_one() {
  return cudaFuncSetAttribute(ttm_first_kernel<NT, KMAJOR, ttm_stages<NT, KMAJOR>()>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ttm_smem<NT, KMAJOR>());
}

template <bool KMAJOR>
int dispatch_nt(ppx_ctx *ctx, const TtmParams &p, int nt, int col_blocks) {
  switch (nt) {
    case 1: return launch_ttm<1, KMAJOR>(ctx, p, col_blocks);
    case 2: return launch_ttm<2, KMAJOR>(ctx, p, col_blocks);
    case 3: return launch_ttm<3, KMAJOR>(ctx, p, col_blocks);
    case 4: return launch_ttm<4, KMAJOR>(ctx, p, col_blocks);
    case 5: return launch_ttm<5, KMAJOR>(ctx, p, col_blocks);
    case 6: return launch_ttm<6, KMAJOR>(ctx, p, col_blocks);
    case 7: return launch_ttm<7, KMAJOR>(ctx, p, col_blocks);
    default: return launch_ttm<8, KMAJOR>(ctx, p, col_blocks);
  }
}

}  // namespace

int ppx_k1_init(ppx_ctx *ctx) {
  cudaError_t e = cudaSuccess;
#define PPX_K1_INIT(NT)                                  \
  if (e == cudaSuccess) e = init_one<NT, true>();        \
  if (e == cudaSuccess) e = init_one<NT, false>();
  PPX_K1_INIT(1) PPX_K1_INIT(2) PPX_K1_INIT(3) PPX_K1_INIT(4) PPX_K1_INIT(5) PPX_K1_INIT(6) PPX_K1_INIT(7) PPX_K1_INIT(8)
#undef PPX_K1_INIT
  if (e != cudaSuccess) return ppx_set_err(ctx, PPX_ECUDA, "k1 init: %s", cudaGetErrorString(e));
  return PPX_OK;
}

// Shared by ppx_ttm_first (CP, rank last), ppx_ttm / ppx_ttm_acc (Tucker, rank in place of mode x) and ppx_ttm_multi.
// Uses the context workspace for split-K partials (callers that hold workspace memory pass ws_keep = true).
int ppx_ttm_impl(ppx_ctx *ctx, const double *V, int64_t L, int64_t X, int64_t Rt, const double *Wx, int64_t ldw,
                 int R, double *out, int inplace, int accumulate, bool ws_keep, bool try_tma) {
  TtmParams p;
  p.V = V;
  p.W = Wx;
  p.out = out;
  p.L = L;
  p.X = X;
  p.Rt = Rt;
  p.Mtot = L * Rt;
  p.ldw = ldw;
  p.R = R;
  p.inplace = inplace;
  p.accumulate = accumulate;
  p.ksplit = 1;
  p.split_stride = 0;
  if (p.Mtot == 0 || R == 0) return PPX_OK;
  {  // small X and R: bandwidth problem with a large output -> streaming DFMA kernel
    const int rc = ppx_ttm_stream_try(ctx, V, L, X, Rt, Wx, ldw, R, out, inplace, accumulate);
    if (rc != 1) return rc;
  }
  if (try_tma) {  // TMA-staged tiles when the shape allows it
    const double *fac[1] = {Wx};
    const int64_t ld1[1] = {ldw}, xs1[1] = {X};
    const int rc = ppx_ttm_tma_try(ctx, V, L, X, Rt, fac, ld1, xs1, 1, R, out, inplace, accumulate, ws_keep);
    if (rc != 1) return rc;
  }
  int64_t tiles = (p.Mtot + BM - 1) / BM;
  if (tiles > 0x7fffffff / 64) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "ttm_first: too many row tiles");
  p.num_tiles = (int)tiles;
  const int64_t nk64 = (X + BK - 1) / BK;
  if (nk64 > 0x7fffffff) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "ttm_first: contracted extent too large");
  p.nk = (int)nk64;
  p.cps = p.nk;
  const bool kmajor = (L == 1);
  const bool v16 = (((uintptr_t)V) & 15) == 0;
  p.vecA = kmajor ? (v16 && (X % 2 == 0)) : (v16 && (L % 2 == 0));
  p.vecW = ((((uintptr_t)Wx) & 15) == 0) && (ldw % 2 == 0);
  const int col_blocks = (R + 63) / 64;
  const int ncols = R < 64 ? R : 64;
  const int nt = (ncols + 7) / 8;

  // K split: only when the row tiles cannot balance the persistent grid and K is deep enough to amortise it
  double *partial = nullptr;
  const int G = 2 * ctx->sm_count;
  if (!inplace && !accumulate && col_blocks == 1 && p.nk >= 64 && p.num_tiles < 8 * G) {
    const int best = ppx_pick_ksplit(p.num_tiles, p.nk, G, p.Mtot * (int64_t)R, ctx->ws_bytes - ctx->ws_used);
    if (best > 1) {
      if (!ws_keep) ppx_ws_reset(ctx);
      partial = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)best * p.Mtot * R);
      if (partial) {
        p.ksplit = best;
        p.cps = (p.nk + best - 1) / best;
        p.split_stride = p.Mtot * (int64_t)R;
        p.out = partial;
      }  // else: not enough workspace -> run unsplit (correct, just less balanced)
    }
  }
  int rc = kmajor ? dispatch_nt<true>(ctx, p, nt, col_blocks) : dispatch_nt<false>(ctx, p, nt, col_blocks);
  if (rc) return rc;
  if (p.ksplit > 1) {
    const int64_t n = p.Mtot * (int64_t)R;
    int blocks = p.ksplit < 16 ? ppx_cdiv(n, 256 * 4) : ppx_cdiv(n, 32);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    split_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(partial, n, p.ksplit, out);
    PPX_CHECK_LAUNCH(ctx);
  }
  return PPX_OK;
}

extern "C" {

int ppx_ttm_first(ppx_ctx *ctx, const double *V, const int64_t *lens, int N, int x, const double *Wx, int64_t ldw,
                  int R, double *out) {
  PPX_REQUIRE(ctx, V && lens && Wx && out, "non-null pointers");
  PPX_REQUIRE(ctx, N >= 1 && N <= 16 && x >= 0 && x < N && R >= 1, "1 <= N <= 16, 0 <= x < N, R >= 1");
  PPX_REQUIRE(ctx, ldw >= lens[x], "ldw >= lens[x]");
  int64_t L, X, Rt;
  ppx_split3(lens, N, x, &L, &X, &Rt);
  return ppx_ttm_impl(ctx, V, L, X, Rt, Wx, ldw, R, out, 0, 0, false, true);
}

int ppx_ttm_multi(ppx_ctx *ctx, const double *V, const int64_t *lens, int N, int x_first, int n_modes,
                  const double *const *W, const int64_t *ldw, int R, double *out) {
  PPX_REQUIRE(ctx, V && lens && W && ldw && out, "non-null pointers");
  PPX_REQUIRE(ctx, N >= 1 && N <= 16 && n_modes >= 1 && n_modes <= 8 && x_first >= 0 && x_first + n_modes <= N && R >= 1,
              "1 <= n_modes <= 8 adjacent modes inside [0, N)");
  if (n_modes == 1) return ppx_ttm_first(ctx, V, lens, N, x_first, W[0], ldw[0], R, out);
  int64_t L = 1, K = 1, Rt = 1;
  for (int i = 0; i < x_first; i++) L *= lens[i];
  KrpArgs a;
  a.n = n_modes;
  for (int j = 0; j < n_modes; j++) {
    PPX_REQUIRE(ctx, W[j] && ldw[j] >= lens[x_first + j], "ldw[j] >= lens of the contracted mode");
    a.w[j] = W[j];
    a.x[j] = lens[x_first + j];
    a.ld[j] = ldw[j];
    K *= lens[x_first + j];
  }
  for (int i = x_first + n_modes; i < N; i++) Rt *= lens[i];
  ppx_ws_reset(ctx);
  {  // TMA path: the Khatri-Rao rows are packed straight into the per-chunk slabs the kernel streams
    int64_t xs[8];
    for (int j = 0; j < n_modes; j++) xs[j] = lens[x_first + j];
    const int rc = ppx_ttm_tma_try(ctx, V, L, K, Rt, W, ldw, xs, n_modes, R, out, 0, 0, true);
    if (rc != 1) return rc;
    ppx_ws_reset(ctx);
  }
  double *krp = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)K * R);  // K x R, column-major, ld = K
  if (!krp)
    return ppx_set_err(ctx, PPX_ENOMEM, "ttm_multi needs %lld bytes of workspace for the Khatri-Rao rows",
                       (long long)(8 * K * R));
  const int64_t ld = K;
  const int64_t total = K * (int64_t)R;
  int blocks = ppx_cdiv(total, 256 * 2);
  if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  krp_kernel<<<blocks, 256, 0, ctx->stream>>>(a, K, R, krp);
  PPX_CHECK_LAUNCH(ctx);
  return ppx_ttm_impl(ctx, V, L, K, Rt, krp, ld, R, out, 0, 0, true, false);
}

int ppx_ttm(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int x, const double *Wx, int64_t ldw, int Q,
            double *out) {
  PPX_REQUIRE(ctx, T && lens && Wx && out, "non-null pointers");
  PPX_REQUIRE(ctx, k >= 1 && k <= 16 && x >= 0 && x < k && Q >= 1, "1 <= k <= 16, 0 <= x < k, Q >= 1");
  PPX_REQUIRE(ctx, ldw >= lens[x], "ldw >= lens[x]");
  int64_t L, X, Rt;
  ppx_split3(lens, k, x, &L, &X, &Rt);
  return ppx_ttm_impl(ctx, T, L, X, Rt, Wx, ldw, Q, out, 1, 0, false, true);
}

int ppx_ttm_acc(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int x, const double *Wx, int64_t ldw,
                int Q, double *out) {
  PPX_REQUIRE(ctx, T && lens && Wx && out, "non-null pointers");
  PPX_REQUIRE(ctx, k >= 1 && k <= 16 && x >= 0 && x < k && Q >= 1, "1 <= k <= 16, 0 <= x < k, Q >= 1");
  PPX_REQUIRE(ctx, ldw >= lens[x], "ldw >= lens[x]");
  int64_t L, X, Rt;
  ppx_split3(lens, k, x, &L, &X, &Rt);
  return ppx_ttm_impl(ctx, T, L, X, Rt, Wx, ldw, Q, out, 1, 1, false, true);
}

}  // extern "C"
