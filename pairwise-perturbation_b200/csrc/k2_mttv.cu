// K2 -- Hadamard-batched contraction (the rank index appears in all three operands, so it is a batch of
//       matrix-vector products, one per rank column: no operand reuse, HBM-bound)
//     out[l, t, r] = sum_x T[l, x, t, r] * W[x, r]            (T viewed as L x X x Rt x R)
// replaces common.cxx:83,128; als_CP.cxx:258-259,281-283,407-408; cp_dt_optimizer.cxx:184-185.
// K3 -- PP first-order correction  M = M0 + sum_j op_j (x) dW_j   (als_CP.cxx:778-794), all operators of one
//       mode in ONE launch, summed in registers / shared memory in a fixed order (deterministic).
#include "ppx_internal.h"

namespace {

// ---- L == 1: contiguous dot products.  A group of G lanes owns FR consecutive outputs per pass. ------------------
// All FR x FU loads of a pass are issued before the first multiply (16 independent 8-byte loads per lane): the dot
// products are short, so memory-level parallelism, not instruction count, decides (one output per pass measured
// 0.48-0.77 of the HBM rate at X = 40 ... 300).  When the FR rows share one rank column (all but the rows straddling
// a column boundary) the factor values are loaded once per pass.
constexpr int FR = 4;  // rows per group and pass
constexpr int FU = 4;  // loads per row, lane and x step
template <int G>
__global__ void __launch_bounds__(256) mttv_first_kernel(const double *__restrict__ T, const double *__restrict__ W,
                                                         double *__restrict__ out, int64_t X, int64_t Rt, int R,
                                                         int64_t ldw) {
  const int64_t n_out = Rt * (int64_t)R;
  const int lane_g = threadIdx.x % G;
  int64_t o0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G) * FR;
  const int64_t ostride = (((int64_t)gridDim.x * blockDim.x) / G) * FR;
  const int64_t niter = (n_out + ostride - 1) / ostride;  // same trip count for every lane (shuffles below)
  for (int64_t it = 0; it < niter; ++it, o0 += ostride) {
    double acc[FR];
#pragma unroll
    for (int j = 0; j < FR; j++) acc[j] = 0.0;
    if (o0 < n_out) {
      const int64_t r0 = o0 / Rt;
      const bool full = o0 + FR <= n_out;
      const bool same_r = full && (o0 + FR - 1) / Rt == r0;
      const double *tp = T + o0 * X;
      if (same_r) {
        const double *wp = W + r0 * ldw;
        for (int64_t x0 = lane_g; x0 < X; x0 += (int64_t)FU * G) {
          double v[FR][FU], w[FU];
#pragma unroll
          for (int u = 0; u < FU; u++) {
            const int64_t x = x0 + (int64_t)u * G;
            const bool ok = x < X;
            w[u] = ok ? wp[x] : 0.0;
#pragma unroll
            for (int j = 0; j < FR; j++) v[j][u] = ok ? tp[j * X + x] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < FU; u++)
#pragma unroll
            for (int j = 0; j < FR; j++) acc[j] = fma(v[j][u], w[u], acc[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < FR; j++) {
          const int64_t o = o0 + j;
          if (o < n_out) {
            const double *wp = W + (o / Rt) * ldw;
            for (int64_t x = lane_g; x < X; x += G) acc[j] = fma(tp[j * X + x], wp[x], acc[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < FR; j++) {
#pragma unroll
      for (int sft = G / 2; sft > 0; sft >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], sft);
    }
    if (lane_g == 0) {
#pragma unroll
      for (int j = 0; j < FR; j++)
        if (o0 + j < n_out) out[o0 + j] = acc[j];
    }
  }
}

// ---- L >= 32: lanes along l (coalesced), warps split x, fixed-order reduction across warps ------------------
// VEC = 2: every lane owns two consecutive l (16-byte loads, 64 l per block); VEC = 1: one l per lane.
// MT_B loads are in flight per thread and the next batch is requested before the current one is consumed: with
// tens of thousands of small blocks the kernel is otherwise bound by the number of dependent round trips per block.
constexpr int MT_WY = 8;
constexpr int MT_B = 8;
template <int VEC>
__global__ void __launch_bounds__(32 * MT_WY) mttv_mid_kernel(const double *__restrict__ T,
                                                              const double *__restrict__ W,
                                                              double *__restrict__ out, int64_t L, int64_t X,
                                                              int64_t Rt, int R, int64_t ldw, int64_t ltiles) {
  __shared__ double part[MT_WY][32 * VEC];
  const int lane = threadIdx.x, wy = threadIdx.y;
  int64_t b = blockIdx.x;  // over (ltile, t, r)
  const int64_t lt = b % ltiles;
  b /= ltiles;
  const int64_t t = b % Rt;
  const int64_t r = b / Rt;
  const int64_t l = (lt * 32 + lane) * VEC;
  double acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; v++) acc[v] = 0.0;
  if (l < L) {
    const double *tp = T + l + L * X * (t + Rt * r);
    const double *wp = W + r * ldw;
    double cur[MT_B][VEC], nxt[MT_B][VEC];
    auto load = [&](double (&dst)[MT_B][VEC], int64_t x0) {
#pragma unroll
      for (int u = 0; u < MT_B; u++) {
        const int64_t x = x0 + (int64_t)u * MT_WY;
        if (x < X) {
          if (VEC == 2) {
            const double2 v = *reinterpret_cast<const double2 *>(tp + L * x);
            dst[u][0] = v.x;
            dst[u][VEC - 1] = v.y;
          } else {
            dst[u][0] = tp[L * x];
          }
        } else {
#pragma unroll
          for (int v = 0; v < VEC; v++) dst[u][v] = 0.0;
        }
      }
    };
    load(cur, wy);
    for (int64_t x0 = wy; x0 < X; x0 += MT_B * MT_WY) {
      const int64_t x1 = x0 + MT_B * MT_WY;
      if (x1 < X) load(nxt, x1);
#pragma unroll
      for (int u = 0; u < MT_B; u++) {
        const int64_t x = x0 + (int64_t)u * MT_WY;
        const double w = x < X ? wp[x] : 0.0;
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[v] += cur[u][v] * w;
      }
#pragma unroll
      for (int u = 0; u < MT_B; u++)
#pragma unroll
        for (int v = 0; v < VEC; v++) cur[u][v] = nxt[u][v];
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; v++) part[wy][lane * VEC + v] = acc[v];
  __syncthreads();
  if (wy == 0 && l < L) {
#pragma unroll
    for (int v = 0; v < VEC; v++) {
      double s = 0.0;
#pragma unroll
      for (int y = 0; y < MT_WY; y++) s += part[y][lane * VEC + v];
      out[l + v + L * (t + Rt * r)] = s;
    }
  }
}

// ---- 1 < L < 32: one thread per output (l,t), loop over all x -----------------------------------------------
__global__ void __launch_bounds__(256) mttv_small_kernel(const double *__restrict__ T, const double *__restrict__ W,
                                                         double *__restrict__ out, int64_t L, int64_t X, int64_t Rt,
                                                         int R, int64_t ldw) {
  const int64_t n_out = L * Rt * (int64_t)R;
  int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; o < n_out; o += stride) {
    const int64_t l = o % L;
    const int64_t tr = o / L;  // t + Rt*r
    const int64_t r = tr / Rt;
    const double *tp = T + l + L * X * tr;
    const double *wp = W + r * ldw;
    double acc = 0.0;
    for (int64_t x = 0; x < X; x++) acc += tp[L * x] * wp[x];
    out[o] = acc;
  }
}


// ---- small X (<= 64), many outputs: one thread per output (two when VEC == 2), no cross-warp reduction ---------
// The block-per-(l tile, t, r) kernel below splits x over eight warps; with X = 40 that is five loads per warp and a
// shared-memory reduction for 20 KB of traffic per CTA (order-6 PP build, BASELINE configs[3]: 400 000 such CTAs,
// 0.34-0.54 of the HBM rate).  Here lanes run along the flat output index (l fastest, so loads for a fixed x are
// coalesced), every thread loops over all x with eight independent loads in flight and writes its own output.
template <int VEC>
__global__ void __launch_bounds__(256) mttv_flat_m_kernel(const double *__restrict__ T, const double *__restrict__ W,
                                                          double *__restrict__ out, int64_t L, int64_t X, int64_t Rt,
                                                          int R, int64_t ldw) {
  const int64_t n_out = L * Rt * (int64_t)R, nvec = n_out / VEC;
  const int64_t LX = L * X;
  for (int64_t ov = (int64_t)blockIdx.x * 256 + threadIdx.x; ov < nvec; ov += (int64_t)gridDim.x * 256) {
    const int64_t o = ov * VEC;
    const int64_t tr = o / L, l = o - tr * L;
    const int64_t r = tr / Rt;
    const double *tp = T + l + LX * tr;
    const double *wp = W + r * ldw;
    double acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; e++) acc[e] = 0.0;
    for (int64_t x0 = 0; x0 < X; x0 += 8) {
      double v[8][VEC], w[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (x0 + u < X) {
          if (VEC == 2) {
            const double2 q = *reinterpret_cast<const double2 *>(tp + L * (x0 + u));
            v[u][0] = q.x;
            v[u][VEC - 1] = q.y;
          } else {
            v[u][0] = tp[L * (x0 + u)];
          }
          w[u] = wp[x0 + u];
        } else {
#pragma unroll
          for (int e = 0; e < VEC; e++) v[u][e] = 0.0;
          w[u] = 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++)
#pragma unroll
        for (int e = 0; e < VEC; e++) acc[e] = fma(v[u][e], w[u], acc[e]);
    }
    if (VEC == 2)
      *reinterpret_cast<double2 *>(out + o) = make_double2(acc[0], acc[VEC - 1]);
    else
      out[o] = acc[0];
  }
}

// ---- few outputs, long x (e.g. the leaf contraction over the 7200-image mode of the coil-shaped tensor): the x range
// is split over gridDim.y CTAs and eight warps each; partial sums go to the workspace and are added in a fixed order.
constexpr int SX_WY = 8;
__global__ void __launch_bounds__(32 * SX_WY) mttv_splitx_kernel(const double *__restrict__ T,
                                                                 const double *__restrict__ W,
                                                                 double *__restrict__ part, int64_t L, int64_t X,
                                                                 int64_t Rt, int R, int64_t ldw, int64_t xchunk) {
  __shared__ double red[SX_WY][32];
  const int lane = threadIdx.x, wy = threadIdx.y;
  const int64_t n_out = L * Rt * (int64_t)R;
  const int64_t o = (int64_t)blockIdx.x * 32 + lane;
  double acc = 0.0;
  if (o < n_out) {
    const int64_t tr = o / L, l = o - tr * L;
    const double *tp = T + l + L * X * tr;
    const double *wp = W + (tr / Rt) * ldw;
    const int64_t xb = (int64_t)blockIdx.y * xchunk;
    const int64_t xe = xb + xchunk < X ? xb + xchunk : X;
    for (int64_t x0 = xb + wy; x0 < xe; x0 += 4 * SX_WY) {
      double v[4], w[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int64_t x = x0 + u * SX_WY;
        v[u] = x < xe ? tp[L * x] : 0.0;
        w[u] = x < xe ? wp[x] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) acc = fma(v[u], w[u], acc);
    }
  }
  red[wy][lane] = acc;
  __syncthreads();
  if (wy == 0 && o < n_out) {
    double s = 0.0;
#pragma unroll
    for (int y = 0; y < SX_WY; y++) s += red[y][lane];
    part[(int64_t)blockIdx.y * n_out + o] = s;
  }
}
__global__ void __launch_bounds__(256) mttv_splitx_reduce_kernel(const double *__restrict__ part, int64_t n, int nsplit,
                                                                 double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double s = part[i];
  for (int k = 1; k < nsplit; k++) s += part[(int64_t)k * n + i];
  out[i] = s;
}

// ---- K3 ------------------------------------------------------------------------------------------------------
constexpr int PP_MAX_OPS = 15;
struct PpArgs {
  const double *op[PP_MAX_OPS];
  const double *dw[PP_MAX_OPS];
  int64_t sj[PP_MAX_OPS];
  int which[PP_MAX_OPS];
  int n_ops;
};

// One CTA per (32 rows of M, rank column r); 8 warps; up to 4 CTAs per SM, so the whole grid (500 CTAs at s = 300,
// R = 50) is resident at once and every thread keeps PP_B 16-byte loads in flight: at PP-operator sizes (3 x 36 MB per
// mode) the kernel lives or dies by memory-level parallelism, not by instruction count.
//   which == 1, op[i', q, r] (rows contiguous along i'): lanes along i', warps (and half-warps when VEC == 2) split q;
//   which == 0, op[q, i', r] (contiguous along q): one warp per row i', lanes along q.
// VEC == 2 needs even s_i / s_j and 16-byte aligned operators (checked by the launcher per call).
constexpr int PP_WY = 8;  // warps per CTA
constexpr int PP_B = 8;   // loads in flight per thread and batch
template <int VEC>
__global__ void __launch_bounds__(32 * PP_WY, 4) pp_correct_kernel(const double *__restrict__ M0, PpArgs a,
                                                                   int64_t s_i, int R, double *__restrict__ Mout,
                                                                   double *__restrict__ part) {
  // gridDim.z > 1: the contracted index q of every operator is split into gridDim.z (even-sized) ranges; each z writes
  // its partial sums to part[z][s_i x R] and pp_correct_reduce_kernel adds them to M0 in a fixed order.  Used when
  // s_i x R alone gives too few CTAs (the size-3 and size-128 modes of the coil-shaped tensor against s_j = 7200).
  __shared__ double psum[PP_WY][32];  // which==1 partial sums (per warp, per row)
  __shared__ double dots[32];         // which==0 sums (per row; row rr is owned by warp rr % PP_WY)
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int r = blockIdx.y;
  const int lrow = VEC == 2 ? 2 * (lane & 15) : lane;  // which==1: first row owned by this lane
  const int qlane = VEC == 2 ? (lane >> 4) : 0;        // which==1, VEC==2: the two half-warps take alternate q
  double acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; v++) acc[v] = 0.0;
  double dsum[32 / PP_WY];
#pragma unroll
  for (int k = 0; k < 32 / PP_WY; k++) dsum[k] = 0.0;

  for (int j = 0; j < a.n_ops; j++) {
    const int64_t sj = a.sj[j];
    const double *dw = a.dw[j] + sj * r;
    int64_t qb = 0, qe = sj;
    if (gridDim.z > 1) {
      const int64_t chunk = (((sj + gridDim.z - 1) / gridDim.z) + 1) & ~(int64_t)1;
      qb = (int64_t)blockIdx.z * chunk;
      qe = qb + chunk < sj ? qb + chunk : sj;
    }
    if (a.which[j]) {
      if (i0 + lrow < s_i) {
        const double *pp = a.op[j] + i0 + lrow + s_i * sj * (int64_t)r;
        constexpr int QSTEP = PP_WY * VEC;
        for (int64_t q0 = qb + wy * VEC + qlane; q0 < qe; q0 += PP_B * QSTEP) {
          double v[PP_B][VEC];
#pragma unroll
          for (int u = 0; u < PP_B; u++) {
            const int64_t q = q0 + u * QSTEP;
            if (q < qe) {
              if (VEC == 2) {
                const double2 t = *reinterpret_cast<const double2 *>(pp + s_i * q);
                v[u][0] = t.x;
                v[u][VEC - 1] = t.y;
              } else {
                v[u][0] = pp[s_i * q];
              }
            } else {
#pragma unroll
              for (int e = 0; e < VEC; e++) v[u][e] = 0.0;
            }
          }
#pragma unroll
          for (int u = 0; u < PP_B; u++) {
            const int64_t q = q0 + u * QSTEP;
            const double w = q < qe ? dw[q] : 0.0;
#pragma unroll
            for (int e = 0; e < VEC; e++) acc[e] = fma(v[u][e], w, acc[e]);
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 32 / PP_WY; k++) {
        const int64_t ii = i0 + wy + PP_WY * k;
        double d = 0.0;
        if (ii < s_i) {
          const double *pp = a.op[j] + sj * (ii + s_i * (int64_t)r);
          for (int64_t q0 = qb + lane * VEC; q0 < qe; q0 += 32 * VEC * PP_B) {
            double v[PP_B][VEC];
#pragma unroll
            for (int u = 0; u < PP_B; u++) {
              const int64_t q = q0 + 32 * VEC * u;
              if (q < qe) {
                if (VEC == 2) {
                  const double2 t = *reinterpret_cast<const double2 *>(pp + q);
                  v[u][0] = t.x;
                  v[u][VEC - 1] = t.y;
                } else {
                  v[u][0] = pp[q];
                }
              } else {
#pragma unroll
                for (int e = 0; e < VEC; e++) v[u][e] = 0.0;
              }
            }
#pragma unroll
            for (int u = 0; u < PP_B; u++) {
              const int64_t q = q0 + 32 * VEC * u;
              if (q < qe) {
#pragma unroll
                for (int e = 0; e < VEC; e++) d = fma(v[u][e], dw[q + e], d);
              }
            }
          }
        }
        dsum[k] += ppx_warp_sum(d);
      }
    }
  }
  if (VEC == 2) {
#pragma unroll
    for (int e = 0; e < VEC; e++) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    if (lane < 16) {
#pragma unroll
      for (int e = 0; e < VEC; e++) psum[wy][lrow + e] = acc[e];
    }
  } else {
    psum[wy][lane] = acc[0];
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 32 / PP_WY; k++) dots[wy + PP_WY * k] = dsum[k];
  }
  __syncthreads();
  if (wy == 0 && i0 + lane < s_i) {
    double s = gridDim.z > 1 ? 0.0 : M0[i0 + lane + s_i * r];
#pragma unroll
    for (int y = 0; y < PP_WY; y++) s += psum[y][lane];
    s += dots[lane];
    if (gridDim.z > 1)
      part[(int64_t)blockIdx.z * s_i * R + i0 + lane + s_i * r] = s;
    else
      Mout[i0 + lane + s_i * r] = s;
  }
}

__global__ void __launch_bounds__(256) pp_correct_reduce_kernel(const double *__restrict__ M0,
                                                                const double *__restrict__ part, int64_t n, int nz,
                                                                double *__restrict__ Mout) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double s = M0[i];
  for (int z = 0; z < nz; z++) s += part[(int64_t)z * n + i];
  Mout[i] = s;
}

}  // namespace

int ppx_mttv_impl(ppx_ctx *ctx, const double *T, int64_t L, int64_t X, int64_t Rt, const double *Wx, int64_t ldw,
                  int R, double *out) {
  const int64_t n_out = L * Rt * (int64_t)R;
  if (n_out == 0) return PPX_OK;
  static const bool no_flat = getenv("PPX_NO_FLAT") != nullptr;  // experiments only
  // few outputs and a long contracted mode: split x over CTAs (partials in the workspace, fixed-order sum)
  if (!no_flat && X >= 512 && (n_out + 31) / 32 < ctx->sm_count) {
    const int64_t ob = (n_out + 31) / 32;
    int64_t nsplit = (4 * (int64_t)ctx->sm_count + ob - 1) / ob;
    if (nsplit > (X + 63) / 64) nsplit = (X + 63) / 64;
    const size_t mark = ctx->ws_used;
    double *part = nsplit > 1 ? (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)nsplit * n_out) : nullptr;
    if (part) {
      const int64_t xchunk = (X + nsplit - 1) / nsplit;
      mttv_splitx_kernel<<<dim3((unsigned)ob, (unsigned)nsplit), dim3(32, SX_WY), 0, ctx->stream>>>(T, Wx, part, L, X, Rt,
                                                                                                   R, ldw, xchunk);
      PPX_CHECK_LAUNCH(ctx);
      mttv_splitx_reduce_kernel<<<ppx_cdiv(n_out, 256), 256, 0, ctx->stream>>>(part, n_out, (int)nsplit, out);
      PPX_CHECK_LAUNCH(ctx);
      ctx->ws_used = mark;  // stream ordered: the next user of this scratch runs after the two kernels above
      return PPX_OK;
    }
    ctx->ws_used = mark;
  }
  if (!no_flat && X <= 64 && n_out >= 512 && L >= 16) {
    const bool vec = (L % 2 == 0) && ((((uintptr_t)T) | ((uintptr_t)out)) & 15) == 0;
    const int64_t nvec = vec ? n_out / 2 : n_out;
    int64_t blocks = (nvec + 255) / 256;
    if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
    if (vec)
      mttv_flat_m_kernel<2><<<(int)blocks, 256, 0, ctx->stream>>>(T, Wx, out, L, X, Rt, R, ldw);
    else
      mttv_flat_m_kernel<1><<<(int)blocks, 256, 0, ctx->stream>>>(T, Wx, out, L, X, Rt, R, ldw);
    PPX_CHECK_LAUNCH(ctx);
    return PPX_OK;
  }
  if (L == 1) {
    // lanes per output: every lane should have about FU elements of a row (one x step), at least 4 lanes
    int G = 32;
    while (G > 4 && X <= (int64_t)(G / 2) * FU + G / 4) G >>= 1;
    const int64_t groups = (n_out + FR - 1) / FR;
    int64_t blocks = (groups * G + 255) / 256;
    if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
    switch (G) {
      case 32: mttv_first_kernel<32><<<(int)blocks, 256, 0, ctx->stream>>>(T, Wx, out, X, Rt, R, ldw); break;
      case 16: mttv_first_kernel<16><<<(int)blocks, 256, 0, ctx->stream>>>(T, Wx, out, X, Rt, R, ldw); break;
      case 8: mttv_first_kernel<8><<<(int)blocks, 256, 0, ctx->stream>>>(T, Wx, out, X, Rt, R, ldw); break;
      default: mttv_first_kernel<4><<<(int)blocks, 256, 0, ctx->stream>>>(T, Wx, out, X, Rt, R, ldw); break;
    }
    PPX_CHECK_LAUNCH(ctx);
  } else if (L >= 32) {
    const bool vec = (L % 2 == 0) && ((((uintptr_t)T) & 15) == 0) && L >= 64;
    const int64_t ltiles = vec ? (L + 63) / 64 : (L + 31) / 32;
    const int64_t blocks = ltiles * Rt * R;
    if (blocks > 0x7fffffffLL) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "mttv: grid too large");
    if (vec)
      mttv_mid_kernel<2><<<(unsigned)blocks, dim3(32, MT_WY), 0, ctx->stream>>>(T, Wx, out, L, X, Rt, R, ldw, ltiles);
    else
      mttv_mid_kernel<1><<<(unsigned)blocks, dim3(32, MT_WY), 0, ctx->stream>>>(T, Wx, out, L, X, Rt, R, ldw, ltiles);
    PPX_CHECK_LAUNCH(ctx);
  } else {
    int64_t blocks = (n_out + 255) / 256;
    if (blocks > (int64_t)ctx->sm_count * 32) blocks = (int64_t)ctx->sm_count * 32;
    mttv_small_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(T, Wx, out, L, X, Rt, R, ldw);
    PPX_CHECK_LAUNCH(ctx);
  }
  return PPX_OK;
}

extern "C" {

int ppx_mttv(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int x, const double *Wx, int64_t ldw, int R,
             double *out) {
  PPX_REQUIRE(ctx, T && lens && Wx && out, "non-null pointers");
  PPX_REQUIRE(ctx, k >= 1 && k <= 16 && x >= 0 && x < k && R >= 1, "1 <= k <= 16, 0 <= x < k, R >= 1");
  PPX_REQUIRE(ctx, ldw >= lens[x], "ldw >= lens[x]");
  int64_t L, X, Rt;
  ppx_split3(lens, k, x, &L, &X, &Rt);
  return ppx_mttv_impl(ctx, T, L, X, Rt, Wx, ldw, R, out);
}

int ppx_mttv2(ppx_ctx *ctx, const double *T, const int64_t *lens, int k, int x1, const double *W1, int64_t ldw1,
              int x2, const double *W2, int64_t ldw2, int R, double *out) {
  PPX_REQUIRE(ctx, T && lens && W1 && W2 && out, "non-null pointers");
  PPX_REQUIRE(ctx, k >= 2 && k <= 16 && x1 >= 0 && x1 < x2 && x2 < k && R >= 1, "0 <= x1 < x2 < k <= 16");
  // contract the later mode first (keeps x1's position), through the context workspace
  int64_t L, X, Rt;
  ppx_split3(lens, k, x2, &L, &X, &Rt);
  ppx_ws_reset(ctx);
  double *tmp = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)(L * Rt * R));
  if (!tmp)
    return ppx_set_err(ctx, PPX_ENOMEM, "mttv2 needs %lld bytes of workspace (have %zu)",
                       (long long)(8 * L * Rt * R), ctx->ws_bytes);
  int rc = ppx_mttv_impl(ctx, T, L, X, Rt, W2, ldw2, R, tmp);
  if (rc) return rc;
  int64_t lens2[16];
  int kk = 0;
  for (int i = 0; i < k; i++)
    if (i != x2) lens2[kk++] = lens[i];
  ppx_split3(lens2, kk, x1, &L, &X, &Rt);
  return ppx_mttv_impl(ctx, tmp, L, X, Rt, W1, ldw1, R, out);
}

int ppx_pp_correct(ppx_ctx *ctx, const double *M0, const double *const *ops, const int *which,
                   const double *const *dW, const int64_t *s_other, int n_ops, int64_t s_i, int R, double *M_out) {
  PPX_REQUIRE(ctx, M0 && M_out && n_ops >= 0 && n_ops <= PP_MAX_OPS, "0 <= n_ops <= 15");
  PPX_REQUIRE(ctx, s_i >= 1 && R >= 1 && R <= 65535, "s_i >= 1, 1 <= R <= 65535");
  PpArgs a;
  a.n_ops = n_ops;
  for (int j = 0; j < n_ops; j++) {
    a.op[j] = ops[j];
    a.dw[j] = dW[j];
    a.sj[j] = s_other[j];
    a.which[j] = which[j];
  }
  // 16-byte loads when every operator slab and dW column stays 16-byte aligned
  bool vec = (s_i % 2 == 0);
  for (int j = 0; j < n_ops && vec; j++)
    vec = (s_other[j] % 2 == 0) && ((((uintptr_t)ops[j]) | ((uintptr_t)dW[j])) & 15) == 0;
  dim3 grid(ppx_cdiv(s_i, 32), R), block(32 * PP_WY);
  // too few CTAs for the machine and long operators: split the contracted index over gridDim.z
  int64_t max_sj = 0;
  for (int j = 0; j < n_ops; j++) max_sj = s_other[j] > max_sj ? s_other[j] : max_sj;
  const int64_t ctas = (int64_t)grid.x * grid.y;
  double *part = nullptr;
  int nz = 1;
  static const bool no_split = getenv("PPX_NO_FLAT") != nullptr;  // experiments only
  // (not on a lane: the partial sums live in the context workspace, which the main stream may be using)
  if (!no_split && ctx->open_lane < 0 && ctas < 2 * ctx->sm_count && max_sj >= 256) {
    nz = (int)((4 * (int64_t)ctx->sm_count + ctas - 1) / ctas);
    if (nz > (max_sj + 63) / 64) nz = (int)((max_sj + 63) / 64);
    if (nz > 64) nz = 64;
    if (nz > 1) {
      ppx_ws_reset(ctx);
      part = (double *)ppx_ws_alloc(ctx, sizeof(double) * (size_t)nz * s_i * R);
    }
    if (!part) nz = 1;
  }
  grid.z = nz;
  if (vec)
    pp_correct_kernel<2><<<grid, block, 0, ctx->stream>>>(M0, a, s_i, R, M_out, part);
  else
    pp_correct_kernel<1><<<grid, block, 0, ctx->stream>>>(M0, a, s_i, R, M_out, part);
  PPX_CHECK_LAUNCH(ctx);
  if (nz > 1) {
    pp_correct_reduce_kernel<<<ppx_cdiv(s_i * R, 256), 256, 0, ctx->stream>>>(M0, part, s_i * (int64_t)R, nz, M_out);
    PPX_CHECK_LAUNCH(ctx);
  }
  return PPX_OK;
}

}  // extern "C"
