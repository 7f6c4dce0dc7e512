// K4 -- Gram matrices and their Hadamard product  S = o_{j != i} (W_j^T W_j) + lambda I
//       (als_CP.cxx:288-292,573-579,796-802; cp_als_optimizer.cxx:33-37)
// K5 -- R x R solve  W = M S^-1  fused with the gradient  G = -M + W_old S  and the PP update
//       dW = ratio*(W - W_init)   (common.cxx:710-758; als_CP.cxx:296,582,811-812)
// K6 -- Normalize (common.cxx:680-688)
//
// The R x R inverse is formed once by ONE CTA (square-root-free Cholesky: S = L D L^T, S^-1 = L^-T D^-1 L^-1; or, for the reference's
// SVD_solve semantics, a cyclic Jacobi eigen-decomposition S = Q diag(e) Q^T, S^-1 = Q diag(1/e) Q^T with no
// truncation, common.cxx:720-722), then applied to the s x R right-hand side by a grid of 8-row tiles.  The
// Hadamard product of the cached Grams is fused into the inverse kernel (ppx_solve_update_g), so one mode update of
// the PP sweep is: correction kernel, inverse kernel, apply kernel, Gram kernel.
#include "ppx_internal.h"

namespace {

// ---- Gram: one warp per (a,b), a <= b, mirrored --------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_kernel(const double *__restrict__ W, int64_t s, int64_t ldw, int R,
                                                   double *__restrict__ G) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int npairs = R * (R + 1) / 2;
  if (warp_global >= npairs) return;
  // unrank (a,b), a <= b, from the row-major upper triangle
  int a = 0, rem = warp_global;
  while (rem >= R - a) {
    rem -= R - a;
    a++;
  }
  const int b = a + rem;
  const double *wa = W + (int64_t)a * ldw, *wb = W + (int64_t)b * ldw;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
  int64_t i = lane;
  for (; i + 96 < s; i += 128) {
    acc0 += wa[i] * wb[i];
    acc1 += wa[i + 32] * wb[i + 32];
    acc2 += wa[i + 64] * wb[i + 64];
    acc3 += wa[i + 96] * wb[i + 96];
  }
  for (; i < s; i += 32) acc0 += wa[i] * wb[i];
  double v = ppx_warp_sum((acc0 + acc1) + (acc2 + acc3));
  if (lane == 0) {
    G[a + R * b] = v;
    G[b + R * a] = v;
  }
}

struct HadArgs {
  const double *g[16];
  int n;
};
__device__ __forceinline__ double hadamard_at(const HadArgs &h, int idx, int R, double lambda) {
  double v = h.g[0][idx];
  for (int j = 1; j < h.n; j++) v *= h.g[j][idx];
  if (lambda != 0.0 && (idx / R) == (idx % R)) v += lambda;
  return v;
}
__global__ void hadamard_kernel(HadArgs h, int R, double lambda, double *__restrict__ S) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * R) return;
  S[idx] = hadamard_at(h, idx, R, lambda);
}

// ---- R x R SPD inverse, one CTA of 16 x 16 threads, matrix in registers -------------------------------------------
// S is the Hadamard product of h.n matrices (+ lambda I); it is written to S_out (the apply kernel needs it for the
// gradient) and inverted through the square-root-free Cholesky factorisation S = L D L^T (L unit lower triangular):
//      S^-1 = Y^T D^-1 Y = sum_k (1/d_k) y_k^T y_k,   Y = L^-1,  y_k = row k of Y.
// Thread (tx, ty) owns the elements (i, j), i = ty (mod 16), j = tx (mod 16), of a T x T register tile Z that holds
// the partially factorised S in its columns j > k and the rows of Y built so far in the columns j <= k (every position
// is touched by exactly one of the two updates, so they share storage), and of a second tile that accumulates S^-1:
//      step k:  p_j = z_jk (j > k, column k of S),  p_j = z_kj (j < k, row k of Y),  p_k = 1,  t_i = p_i / d_k,
//               rows i >  k:  z_ij   -= t_i p_j            (column k was zeroed by its owners: z_ik = -t_i)
//               rows i <= k:  inv_ij += t_i p_j  (j <= k)  (row k of Y is final: its term of S^-1)
// The owners publish p through a double-buffered shared vector: ONE barrier per column.
// The kernel is ONE CTA with two warps per scheduler: it runs at the latency of its dependent instruction stream
// (measured with clock64 probes, tools/inv_bench.cu: about 5 cycles per instruction, 365 for an IEEE division), so
//   * tile indices are compile-time (unrolled over the tile column kk, a real loop over the 16 columns inside it)
//     and the update has no selects;
//   * the reciprocal of the NEXT pivot is computed one step ahead, off the critical path: the owner of z_{k+1,k+1}
//     publishes it (before the update) with column k, and every thread forms d_{k+1} = z_{k+1,k+1} - p_{k+1}^2 / d_k
//     and its Newton reciprocal while the updates of step k are in flight;
//   * S^-1 is accumulated inside the loop (the latency-bound loop has issue slots to spare) instead of a separate
//     triangular product through shared memory.
constexpr int INV_B = 16;

// 1/d for the pivots of an SPD matrix: float seed + three Newton steps (full double precision); IEEE division outside
// the float range.
__device__ __forceinline__ double inv_pivot(double d) {
  const float f = (float)d;
  if (!(fabsf(f) > 1e-30f && fabsf(f) < 1e30f)) return 1.0 / d;
  double x = (double)__fdividef(1.0f, f);
  x = fma(x, fma(-d, x, 1.0), x);
  x = fma(x, fma(-d, x, 1.0), x);
  x = fma(x, fma(-d, x, 1.0), x);
  return x;
}

template <int T>
__global__ void __launch_bounds__(INV_B *INV_B) spd_inverse_ldl_kernel(HadArgs h, int R, double lambda,
                                                                       double *__restrict__ S_out,
                                                                       double *__restrict__ Sinv,
                                                                       double *__restrict__ Linv_out) {
  constexpr int B = INV_B;
  constexpr int NTH = B * B;
  constexpr int RP = B * T;   // padded order
  constexpr int PV = RP + 2;  // p_0..p_{RP-1}, slot RP: z_{k+1,k+1} before the update of step k
  extern __shared__ double sm[];
  const int ld = R + 1;
  double *pv = sm;            // [2][PV]
  double *Ss = sm + 2 * PV;   // [R][ld] staging of S
  const int tx = threadIdx.x % B, ty = threadIdx.x / B;
#ifdef PPX_INV_PROFILE
#define PPX_PROF(slot) if (threadIdx.x == 0) g_prof[slot] = clock64()
#else
#define PPX_PROF(slot)
#endif
  PPX_PROF(0);
  // S = Hadamard product (+ lambda I) -> S_out and, through shared memory, the registers; four independent elements
  // per thread and pass keep the global loads of a pass in flight together
  for (int e0 = threadIdx.x; e0 < R * R; e0 += 4 * NTH) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u * NTH;
      v[u] = e < R * R ? h.g[0][e] : 0.0;
    }
#pragma unroll 1
    for (int m = 1; m < h.n; m++) {
      const double *g = h.g[m];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * NTH;
        if (e < R * R) v[u] *= g[e];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u * NTH;
      if (e < R * R) {
        const int j = e / R, i = e - j * R;
        if (i == j) v[u] += lambda;
        if (S_out) S_out[e] = v[u];
        Ss[i * ld + j] = v[u];
      }
    }
  }
  __syncthreads();
  double z[T][T], acc[T][T];
#pragma unroll
  for (int ii = 0; ii < T; ii++)
#pragma unroll
    for (int jj = 0; jj < T; jj++) {
      const int i = ty + B * ii, j = tx + B * jj;
      // identity padding keeps the factorisation of the padded matrix trivial
      z[ii][jj] = (i < R && j < R) ? Ss[i * ld + j] : ((i == j) ? 1.0 : 0.0);
      acc[ii][jj] = 0.0;
    }
  double inv = inv_pivot(Ss[0]);  // 1 / d_0
  PPX_PROF(1);
#pragma unroll
  for (int kk = 0; kk < T; kk++) {
#pragma unroll 1
    for (int kt = 0; kt < B; kt++) {
      const int k = B * kk + kt;
      if (k >= R) break;
      double *p = pv + (k & 1) * PV;
      if (tx == kt) {  // owners of column k: p_i = z_ik (i > k); the column is cleared for Y; p_k = 1
#pragma unroll
        for (int ii = 0; ii < T; ii++) {
          const int i = ty + B * ii;
          if (i > k) {
            p[i] = z[ii][kk];
            z[ii][kk] = 0.0;
          } else if (i == k) {
            p[i] = 1.0;
          }
        }
      }
      if (ty == kt) {  // owners of row k: p_j = z_kj for j < k
#pragma unroll
        for (int jj = 0; jj < T; jj++) {
          const int j = tx + B * jj;
          if (j < k) p[j] = z[kk][jj];
        }
      }
      if (tx == ty) {  // owner of the next diagonal element
        if (kt + 1 < B) {
          if (tx == kt + 1) p[RP] = z[kk][kk];
        } else if (kk + 1 < T) {
          if (tx == 0) p[RP] = z[kk + 1 < T ? kk + 1 : kk][kk + 1 < T ? kk + 1 : kk];
        }
      }
      __syncthreads();
      if (Linv_out) {
        // row k of L^-1 (S = L L^T): sqrt(1/d_k) * (row k of Y, unit diagonal); columns past k are zero
        const double sc = sqrt(inv);
        for (int j = threadIdx.x; j < R; j += NTH) Linv_out[k + R * j] = j <= k ? p[j] * sc : 0.0;
      }
      double t[T], pj[T], pjy[T];
#pragma unroll
      for (int ii = 0; ii < T; ii++) t[ii] = p[ty + B * ii] * inv;
#pragma unroll
      for (int jj = 0; jj < T; jj++) {
        pj[jj] = p[tx + B * jj];
        pjy[jj] = (tx + B * jj <= k) ? pj[jj] : 0.0;
      }
      // next pivot and its reciprocal (used in the next step)
      double inv_next = 0.0;
      if (k + 1 < R) {
        const double pn = p[k + 1];
        inv_next = inv_pivot(fma(-(pn * inv), pn, p[RP]));
      }
#pragma unroll
      for (int ii = 0; ii < T; ii++) {
        if (ty + B * ii > k) {
#pragma unroll
          for (int jj = 0; jj < T; jj++) z[ii][jj] = fma(-t[ii], pj[jj], z[ii][jj]);
        } else {
#pragma unroll
          for (int jj = 0; jj < T; jj++) acc[ii][jj] = fma(t[ii], pjy[jj], acc[ii][jj]);
        }
      }
      inv = inv_next;
    }
  }
  PPX_PROF(2);
#pragma unroll
  for (int ii = 0; ii < T; ii++)
#pragma unroll
    for (int jj = 0; jj < T; jj++) {
      const int i = ty + B * ii, j = tx + B * jj;
      if (Sinv && i < R && j < R) Sinv[i + R * j] = acc[ii][jj];
    }
  PPX_PROF(4);
}

// (A BLOCKED variant -- 8 columns per step: in-register Gauss-Jordan of the 8 x 8 pivot block by one warp, panel and
// trailing update with one thread per element, Y = L^-1 by block forward substitution, S^-1 = Y^T Dinv Y; NumPy model in
// tools/inv_blocked_proto.py -- was written and measured in round 2: 36.8 us at R = 50 and 123 us at R = 100 against
// 32.8 / 86.0 us for the kernel above.  One CTA runs at the latency of its dependent shared-memory loads and FMAs
// whichever way the work is cut; the seven block steps each pay an 8-round pivot inversion plus three barriers, and
// the two triangular products that the column kernel folds into its loop come on top.  Not adopted.)

// ---- R x R SPD inverse for R <= 64: symmetric sweep with the matrix in registers, two columns per step ---------------
// The column kernel above spends its 33 us at R = 50 on latency: a step is a barrier of 8 warps, a shared-memory
// broadcast, a reciprocal chain and two dependent-issue FMA groups, R times in sequence.  This kernel halves the
// number of sequential steps and shortens each:
//   * four threads own a COLUMN (thread (j, h): rows [h NR/4, (h+1) NR/4) of column j, NR = R rounded up to 8; the
//     threads of a warp share h, so every lane reads the SAME 16 bytes of a published column: one wavefront per load);
//   * the loop over the pivot blocks is unrolled, every register index is compile-time;
//   * a step eliminates the 2 x 2 pivot block K = {k, k+1} (the symmetric sweep operator in block form):
//         P = A_KK^-1,   (u1, u2)_j = P (a_kj, a_k+1,j),   a_ij -= a_ik u1_j + a_i,k+1 u2_j  (i, j not in K),
//         a_Kj = (u1, u2)_j,   a_iK = a_iK P,   A_KK = -P;
//     after R/2 steps the registers hold -S^-1.  Only the PUBLISHED columns k, k+1 are read (a_jk stands in for a_kj,
//     equal up to rounding, so the rows never travel); the columns' own threads start from zero with (u1, u2) =
//     -(column of P), which gives a_iK P and -P by the same FMAs;
//   * the owners of columns k+2, k+3 update those first, one of them (a shuffle brings the third element) forms
//     the next P = adj / det with a Newton reciprocal while its remaining FMAs issue, and publishes columns and P
//     through double-buffered shared vectors: ONE barrier per two columns, no divisions.
// An odd R is padded with an identity row and column.  The result is symmetrised through shared memory on the way out
// (the sweep keeps symmetry only up to rounding).  NumPy model of the scalar form: tools/inv_sweep_proto.py.
// Measured (tools/time_inverse.py, 200 back-to-back launches): see profiles/r02x_inverse_*.json.
__device__ __forceinline__ double rcp_newton(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));  // MUFU.RCP64H: about 20 bits
  // x (1 + e + e^2 + e^3)(1 + e^4), e = 1 - d x: 2^-80 in four dependent operations
  const double e = fma(-d, x, 1.0);
  const double e2 = e * e;
  const double y = fma(x, e, x);
  const double z = fma(y, e2, y);
  return fma(z, e2 * e2, z);
}

#ifdef PPX_INV_PROFILE  // tools/inv_bench.cu: clock probes of the middle step taken by the thread that publishes P
#define PPX_PROF_STEP(slot) if (K == (R / 4) * 2 && j == K + 2 && hh == (K + 2) / NH) g_prof[slot] = clock64()
#else
#define PPX_PROF_STEP(slot)
#endif

constexpr int SWEEP_H = 4;  // threads per column

template <int NR>
struct SweepSmem {
  double col[2][2][NR];  // [buffer][column k or k+1][row]
  double P[2][4];        // [buffer]: P00, P01, P11
  double st[NR][NR + 1];
};

template <int NR, int K>
struct SweepSteps {
  static constexpr int NH = NR / SWEEP_H;
  __device__ __forceinline__ static void run(double (&c)[NH], SweepSmem<NR> &sm, int R, int j, int hh) {
    if (K < R) {
      PPX_PROF_STEP(16);
      constexpr int b = (K / 2) & 1;
      const double *ca = sm.col[b][0], *cb = sm.col[b][1];
      const double P00 = sm.P[b][0], P01 = sm.P[b][1], P11 = sm.P[b][2];
      const double cj1 = ca[j], cj2 = cb[j];
      const bool own1 = (j == K), own2 = (j == K + 1);
      double u1 = fma(P00, cj1, P01 * cj2), u2 = fma(P01, cj1, P11 * cj2);
      if (own1) u1 = -P00, u2 = -P01;
      if (own2) u1 = -P01, u2 = -P11;
      double va[NH], vb[NH];
      const double2 *a2 = reinterpret_cast<const double2 *>(ca + hh * NH);
      const double2 *b2 = reinterpret_cast<const double2 *>(cb + hh * NH);
#pragma unroll
      for (int i = 0; i < NH; i += 2) {
        const double2 t = a2[i / 2], w = b2[i / 2];
        va[i] = t.x, va[i + 1] = t.y, vb[i] = w.x, vb[i + 1] = w.y;
      }
      if (own1 || own2) {  // the two columns of the block start from zero (one warp per row block takes the branch)
#pragma unroll
        for (int i = 0; i < NH; i++) c[i] = 0.0;
      }
      PPX_PROF_STEP(17);
      constexpr int kl = K % NH, kh = K / NH, nl = (K + 2) % NH, nh = ((K + 2) / NH) % SWEEP_H;
      // rows k+2, k+3 first: in columns k+2, k+3 they are the next pivot block
      c[nl] = fma(-vb[nl], u2, fma(-va[nl], u1, c[nl]));
      c[nl + 1] = fma(-vb[nl + 1], u2, fma(-va[nl + 1], u1, c[nl + 1]));
      const double pa = c[nl], pb = c[nl + 1], pd = __shfl_down_sync(0xffffffffu, c[nl + 1], 1);
      const double rdet = rcp_newton(fma(pa, pd, -pb * pb));
      double n00 = pd * rdet, n01 = -pb * rdet, n11 = pa * rdet;
#pragma unroll
      for (int i = 0; i < NH; i++)
        if (i != nl && i != nl + 1) c[i] = fma(-vb[i], u2, fma(-va[i], u1, c[i]));
      if (hh == kh) c[kl] = u1, c[kl + 1] = u2;
      asm volatile("" : "+d"(n00), "+d"(n01), "+d"(n11));
      PPX_PROF_STEP(19);
      if (K + 2 < R) {
        if (j == K + 2 || j == K + 3) {
          double2 *dst = reinterpret_cast<double2 *>(sm.col[b ^ 1][j - (K + 2)] + hh * NH);
#pragma unroll
          for (int i = 0; i < NH; i += 2) dst[i / 2] = make_double2(c[i], c[i + 1]);
          if (j == K + 2 && hh == nh) sm.P[b ^ 1][0] = n00, sm.P[b ^ 1][1] = n01, sm.P[b ^ 1][2] = n11;
        }
      }
      PPX_PROF_STEP(20);
      __syncthreads();
      PPX_PROF_STEP(21);
    }
    SweepSteps<NR, K + 2>::run(c, sm, R, j, hh);
  }
};
template <int NR>
struct SweepSteps<NR, NR> {
  __device__ __forceinline__ static void run(double (&)[NR / SWEEP_H], SweepSmem<NR> &, int, int, int) {}
};

template <int NR>
__global__ void __launch_bounds__(SWEEP_H *((NR + 31) / 32) * 32)
    spd_inverse_sweep_kernel(HadArgs h, int R, double lambda, double *__restrict__ S_out, double *__restrict__ Sinv) {
  constexpr int NH = NR / SWEEP_H, NRP = (NR + 31) / 32 * 32;
  __shared__ __align__(16) SweepSmem<NR> sm;
  const int hh = threadIdx.x / NRP, j = threadIdx.x % NRP, r0 = hh * NH;
  double c[NH];
  PPX_PROF(8);
  // S = Hadamard product (+ lambda I), read through its symmetry (element (j, row)) so that a warp reads rows
#pragma unroll
  for (int i = 0; i < NH; i++) c[i] = (r0 + i < R && j < R) ? h.g[0][j + R * (r0 + i)] : 0.0;
#pragma unroll 1
  for (int m = 1; m < h.n; m++) {
    const double *g = h.g[m];
#pragma unroll
    for (int i = 0; i < NH; i++)
      if (r0 + i < R && j < R) c[i] *= g[j + R * (r0 + i)];
  }
#pragma unroll
  for (int i = 0; i < NH; i++) {
    if (r0 + i == j && lambda != 0.0) c[i] += lambda;
    if (S_out && r0 + i < R && j < R) S_out[j + R * (r0 + i)] = c[i];
    if (r0 + i == j && j >= R) c[i] = 1.0;  // identity padding: an odd R shares its last pivot block with index R
  }
  {  // block 0: columns 0, 1 and P
    const double pd = __shfl_down_sync(0xffffffffu, c[1], 1);
    const double rdet = rcp_newton(fma(c[0], pd, -c[1] * c[1]));
    if (j < 2) {
#pragma unroll
      for (int i = 0; i < NH; i++) sm.col[0][j][r0 + i] = c[i];
      if (j == 0 && hh == 0) sm.P[0][0] = pd * rdet, sm.P[0][1] = -c[1] * rdet, sm.P[0][2] = c[0] * rdet;
    }
  }
  __syncthreads();
  PPX_PROF(9);
  SweepSteps<NR, 0>::run(c, sm, R, j, hh);
  PPX_PROF(10);
  // -S^-1 is in the registers: symmetrise through shared memory and write
  if (j < NR) {
#pragma unroll
    for (int i = 0; i < NH; i++) sm.st[r0 + i][j] = c[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NH; i++)
    if (r0 + i < R && j < R) Sinv[j + R * (r0 + i)] = -0.5 * (c[i] + sm.st[j][r0 + i]);
  PPX_PROF(11);
}

// Cyclic Jacobi (parallel round-robin ordering) on a symmetric matrix in shared memory.
// A[n][ld], Q[n][ld]; n even (padded with an identity row/column when R is odd).
__device__ void jacobi_eig_shared(double *A, double *Q, double *cs, int n, int ld, int max_sweeps) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int half = n / 2;
  __shared__ double off_norm, diag_norm;
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    // convergence test: off-diagonal Frobenius norm
    __syncthreads();
    if (tid == 0) {
      off_norm = 0.0;
      diag_norm = 0.0;
    }
    __syncthreads();
    double o = 0.0, d = 0.0;
    for (int idx = tid; idx < n * n; idx += nt) {
      int i = idx / n, j = idx % n;
      double v = A[i * ld + j];
      if (i == j) d += v * v;
      else o += v * v;
    }
    o = ppx_warp_sum(o);
    d = ppx_warp_sum(d);
    if ((tid & 31) == 0) {
      atomicAdd(&off_norm, o);  // only used for the stopping test; the rotations themselves are deterministic
      atomicAdd(&diag_norm, d);
    }
    __syncthreads();
    if (off_norm <= 1e-30 * diag_norm) break;
    for (int round = 0; round < n - 1; round++) {
      // round-robin pairing: player n-1 fixed, others rotate
      if (tid < half) {
        int p, q;
        if (tid == 0) {
          p = n - 1;
          q = round % (n - 1);
        } else {
          p = (round + tid) % (n - 1);
          q = (round + n - 1 - tid) % (n - 1);
        }
        if (p > q) {
          int t = p;
          p = q;
          q = t;
        }
        const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
        double c = 1.0, s = 0.0;
        if (fabs(apq) > 1e-300) {
          const double tau = (aqq - app) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t);
          s = t * c;
        }
        cs[4 * tid + 0] = c;
        cs[4 * tid + 1] = s;
        cs[4 * tid + 2] = (double)p;
        cs[4 * tid + 3] = (double)q;
      }
      __syncthreads();
      // columns:  A <- A J,  Q <- Q J
      for (int idx = tid; idx < half * n; idx += nt) {
        const int pr = idx / n, i = idx % n;
        const double c = cs[4 * pr], s = cs[4 * pr + 1];
        const int p = (int)cs[4 * pr + 2], q = (int)cs[4 * pr + 3];
        const double aip = A[i * ld + p], aiq = A[i * ld + q];
        A[i * ld + p] = c * aip - s * aiq;
        A[i * ld + q] = s * aip + c * aiq;
        const double qip = Q[i * ld + p], qiq = Q[i * ld + q];
        Q[i * ld + p] = c * qip - s * qiq;
        Q[i * ld + q] = s * qip + c * qiq;
      }
      __syncthreads();
      // rows:  A <- J^T A
      for (int idx = tid; idx < half * n; idx += nt) {
        const int pr = idx / n, j = idx % n;
        const double c = cs[4 * pr], s = cs[4 * pr + 1];
        const int p = (int)cs[4 * pr + 2], q = (int)cs[4 * pr + 3];
        const double apj = A[p * ld + j], aqj = A[q * ld + j];
        A[p * ld + j] = c * apj - s * aqj;
        A[q * ld + j] = s * apj + c * aqj;
      }
      __syncthreads();
      if (tid < half) {  // the rotated pair is annihilated exactly
        const int p = (int)cs[4 * tid + 2], q = (int)cs[4 * tid + 3];
        A[p * ld + q] = 0.0;
        A[q * ld + p] = 0.0;
      }
      __syncthreads();
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) sym_inverse_jacobi_kernel(HadArgs h, int R, double lambda,
                                                                         double *__restrict__ S_out,
                                                                         double *__restrict__ Sinv) {
  extern __shared__ double sm[];
  const int n = (R + 1) & ~1;
  const int ld = n + 1;
  double *A = sm;
  double *Q = sm + n * ld;
  double *cs = Q + n * ld;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int idx = tid; idx < R * R; idx += nt) {
    const double v = hadamard_at(h, idx, R, lambda);
    if (S_out) S_out[idx] = v;
    A[(idx % R) * ld + idx / R] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += nt) {
    int i = idx / n, j = idx % n;
    if (i < R && j < R) {
      if (i < j) {  // symmetrise in place (each pair once)
        const double v = 0.5 * (A[i * ld + j] + A[j * ld + i]);
        A[i * ld + j] = v;
        A[j * ld + i] = v;
      }
    } else {
      A[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    Q[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  jacobi_eig_shared(A, Q, cs, n, ld, 30);
  // S^-1 = Q diag(1/e) Q^T  (all R eigen-pairs, no truncation: common.cxx:720-722)
  for (int idx = tid; idx < R * R; idx += nt) {
    int i = idx % R, j = idx / R;
    if (i >= j) {
      double v = 0.0;
      for (int k = 0; k < n; k++) {
        if (n != R && fabs(Q[(n - 1) * ld + k]) > 0.5) continue;  // the padding eigenvector
        v += Q[i * ld + k] * Q[j * ld + k] / A[k * ld + k];
      }
      Sinv[i + R * j] = v;
      Sinv[j + R * i] = v;
    }
  }
}

// ---- apply: 8-row tiles; warp = row, lanes = columns --------------------------------------------------------------
constexpr int AP_ROWS = 8;
__global__ void __launch_bounds__(32 * AP_ROWS) solve_apply_kernel(const double *__restrict__ M,
                                                                   const double *__restrict__ S,
                                                                   const double *__restrict__ Sinv,
                                                                   double *__restrict__ W, int64_t s, int R,
                                                                   const double *__restrict__ W_init,
                                                                   double ratio_step, double *__restrict__ grad_out,
                                                                   double *__restrict__ dW_out) {
  extern __shared__ double sm[];
  double *Ss = sm;                 // R*R
  double *Si = Ss + R * R;         // R*R
  double *Mt = Si + R * R;         // AP_ROWS * R : Mt[row*R + r]
  double *Wt = Mt + AP_ROWS * R;   // old W
  const int lane = threadIdx.x, row = threadIdx.y;
  const int tid = row * 32 + lane, nt = 32 * AP_ROWS;
  const int64_t i = (int64_t)blockIdx.x * AP_ROWS + row;
  // (compact compute loops: the kernel runs once per launch on cold instruction caches, code size is latency.  The
  // loads are the opposite: issued one dependent round trip at a time they were 8 of the kernel's 13 us, so every
  // thread now has its row elements and eight elements of each matrix in flight before the first one is stored.)
  double mv[2], wv[2];
#pragma unroll
  for (int u = 0; u < 2; u++) {
    const int r = lane + 32 * u;
    mv[u] = (i < s && r < R) ? M[i + s * r] : 0.0;
    wv[u] = (i < s && r < R) ? W[i + s * r] : 0.0;
  }
#pragma unroll 1
  for (int idx0 = tid; idx0 < R * R; idx0 += 8 * nt) {
    double a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int idx = idx0 + u * nt;
      a[u] = (grad_out && idx < R * R) ? S[idx] : 0.0;
      b[u] = idx < R * R ? Sinv[idx] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int idx = idx0 + u * nt;
      if (idx < R * R) {
        Ss[idx] = a[u];
        Si[idx] = b[u];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; u++) {
    const int r = lane + 32 * u;
    if (r < R) {
      Mt[row * R + r] = mv[u];
      Wt[row * R + r] = wv[u];
    }
  }
#pragma unroll 1
  for (int r = lane + 64; r < R; r += 32) {  // R > 64
    Mt[row * R + r] = (i < s) ? M[i + s * r] : 0.0;
    Wt[row * R + r] = (i < s) ? W[i + s * r] : 0.0;
  }
  __syncthreads();
  if (i >= s) return;
  const double *mrow = Mt + row * R, *wrow = Wt + row * R;
#pragma unroll 1
  for (int c = lane; c < R; c += 32) {
    // S and S^-1 are symmetric: column c is read as row c, contiguous across the lanes (no bank conflicts)
    double w0 = 0.0, w1 = 0.0, g0 = 0.0, g1 = 0.0;
    int r = 0;
#pragma unroll 1
    for (; r + 1 < R; r += 2) {
      w0 = fma(mrow[r], Si[c + R * r], w0);
      w1 = fma(mrow[r + 1], Si[c + R * (r + 1)], w1);
      g0 = fma(wrow[r], Ss[c + R * r], g0);
      g1 = fma(wrow[r + 1], Ss[c + R * (r + 1)], g1);
    }
    if (r < R) {
      w0 = fma(mrow[r], Si[c + R * r], w0);
      g0 = fma(wrow[r], Ss[c + R * r], g0);
    }
    double w = w0 + w1;
    const double g = g0 + g1;
    const int64_t o = i + s * c;
    if (grad_out) grad_out[o] = -mrow[c] + g;
    if (W_init) {
      const double wi = W_init[o];
      const double d = ratio_step * (w - wi);
      if (dW_out) dW_out[o] = d;
      if (ratio_step != 1.0) w = wi + d;
    }
    W[o] = w;
  }
}

// ---- normalize ----------------------------------------------------------------------------------------------
struct NormArgs {
  double *w[16];
  double *g[16];
  int64_t n[16];
  int N;
  int R;
};
__global__ void __launch_bounds__(1024) norm_sq_kernel(NormArgs a, double *sq) {
  __shared__ double red[32];
  const double *x = a.w[blockIdx.x];
  const int64_t n = a.n[blockIdx.x];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i] * x[i];
  s = ppx_block_sum(s, red);
  if (threadIdx.x == 0) sq[blockIdx.x] = s;
}
__global__ void norm_sq_from_gram_kernel(NormArgs a, double *sq) {
  // ||W_i||_F^2 = trace(W_i^T W_i); one warp per factor
  const int i = blockIdx.x;
  double s = 0.0;
  for (int k = threadIdx.x; k < a.R; k += 32) s += a.g[i][k + (int64_t)a.R * k];
  s = ppx_warp_sum(s);
  if (threadIdx.x == 0) sq[i] = s;
}
__global__ void __launch_bounds__(256) norm_scale_kernel(NormArgs a, const double *__restrict__ sq) {
  const int m = blockIdx.y;
  double prod = 1.0;
  for (int j = 0; j < a.N; j++) prod *= sqrt(sq[j]);   // common.cxx:681-683
  const double gm = pow(prod, 1.0 / a.N);               // :684
  const double f = gm / sqrt(sq[m]);                    // :687
  double *x = a.w[m];
  const int64_t n = a.n[m];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = f * x[i];
  if (a.g[m]) {
    double *g = a.g[m];
    const int nn = a.R * a.R;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) g[i] = (f * f) * g[i];
  }
}

// Normalize from the cached Grams + the squared norms the PP switching test needs, one block per mode:
// ||W_j||^2 = trace(G_j) for all j comes in through `tr` (filled by norm_sq_from_gram_kernel in a launch of its own:
// block m rescales G_m in place, so reading the other Grams here would race with the blocks that own them), W_m and
// G_m are rescaled, sq[2m] = ||dW_m||^2, sq[2m+1] = ||W_m||^2 after the rescale.
struct NormNormsArgs {
  double *w[16];
  double *g[16];
  const double *dw[16];
  int64_t n[16];
  int N;
  int R;
};
__global__ void __launch_bounds__(1024) normalize_norms_kernel(NormNormsArgs a, const double *__restrict__ tr,
                                                               double *__restrict__ sq_out) {
  __shared__ double red[32];
  const int m = blockIdx.x;
  double prod = 1.0;
  for (int j = 0; j < a.N; j++) prod *= sqrt(tr[j]);
  const double f = pow(prod, 1.0 / a.N) / sqrt(tr[m]);
  double *x = a.w[m];
  const double *d = a.dw[m];
  const int64_t n = a.n[m];
  double sw = 0.0, sd = 0.0;
  // eight elements of W and dW per thread in flight at a time (one block per mode: the loop is latency, not bandwidth)
  for (int64_t i0 = threadIdx.x; i0 < n; i0 += 8 * (int64_t)blockDim.x) {
    double xv[8], dv[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int64_t i = i0 + u * (int64_t)blockDim.x;
      xv[u] = i < n ? x[i] : 0.0;
      dv[u] = (d && i < n) ? d[i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int64_t i = i0 + u * (int64_t)blockDim.x;
      const double v = f * xv[u];
      if (i < n) x[i] = v;
      sw += v * v;
      sd += dv[u] * dv[u];
    }
  }
  double *g = a.g[m];
  for (int i = threadIdx.x; i < a.R * a.R; i += blockDim.x) g[i] = (f * f) * g[i];
  sw = ppx_block_sum(sw, red);
  sd = ppx_block_sum(sd, red);
  if (threadIdx.x == 0) {
    sq_out[2 * m] = sd;
    sq_out[2 * m + 1] = sw;
  }
}

int inverse_launch(ppx_ctx *ctx, const HadArgs &h, int R, double lambda, int mode, double *S_out, double *Sinv,
                   double *Linv = nullptr) {
  static const bool inv_column = getenv("PPX_INV_COLUMN") != nullptr;  // A/B: the column kernel for every R
  if (mode == PPX_SOLVE_CHOL && R <= 64 && !Linv && !inv_column) {
    // (asking for all the shared memory of an SM, so that CTAs of kernels on other lanes stay off this one's SM, was
    // tried: no measurable effect on the PP sweep)
#define PPX_SWEEP_LAUNCH(NR) \
  spd_inverse_sweep_kernel<NR><<<1, SWEEP_H *((NR + 31) / 32) * 32, 0, ctx->stream>>>(h, R, lambda, S_out, Sinv)
    switch ((R + 7) / 8) {
      case 1: PPX_SWEEP_LAUNCH(8); break;
      case 2: PPX_SWEEP_LAUNCH(16); break;
      case 3: PPX_SWEEP_LAUNCH(24); break;
      case 4: PPX_SWEEP_LAUNCH(32); break;
      case 5: PPX_SWEEP_LAUNCH(40); break;
      case 6: PPX_SWEEP_LAUNCH(48); break;
      case 7: PPX_SWEEP_LAUNCH(56); break;
      default: PPX_SWEEP_LAUNCH(64); break;
    }
#undef PPX_SWEEP_LAUNCH
  } else if (mode == PPX_SOLVE_CHOL) {
    if (R > 112) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "solve: R=%d too large for the one-CTA inverse (max 112)", R);
    const int T = (R + INV_B - 1) / INV_B;
    const size_t smem = sizeof(double) * (2 * ((size_t)INV_B * T + 2) + (size_t)R * (R + 1));
#define PPX_INV_LAUNCH(T) \
  spd_inverse_ldl_kernel<T><<<1, INV_B * INV_B, smem, ctx->stream>>>(h, R, lambda, S_out, Sinv, Linv)
    switch (T) {
      case 1: PPX_INV_LAUNCH(1); break;
      case 2: PPX_INV_LAUNCH(2); break;
      case 3: PPX_INV_LAUNCH(3); break;
      case 4: PPX_INV_LAUNCH(4); break;
      case 5: PPX_INV_LAUNCH(5); break;
      case 6: PPX_INV_LAUNCH(6); break;
      default: PPX_INV_LAUNCH(7); break;
    }
#undef PPX_INV_LAUNCH
  } else {
    const int n = (R + 1) & ~1;
    const size_t smem = sizeof(double) * (2 * (size_t)n * (n + 1) + 4 * (n / 2 + 1));
    if (smem > 220 * 1024) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "solve: R=%d too large for the one-CTA inverse", R);
    sym_inverse_jacobi_kernel<<<1, 1024, smem, ctx->stream>>>(h, R, lambda, S_out, Sinv);
  }
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int solve_impl(ppx_ctx *ctx, const double *M, const HadArgs &h, double lambda, const double *S_given, double *W,
               int64_t s, int R, const double *W_init, double ratio_step, int mode, double *grad_out,
               double *dW_out, double *sq_norms_out) {
  PPX_REQUIRE(ctx, M && W && s >= 1 && R >= 1, "M, W non-null; s, R >= 1");
  PPX_REQUIRE(ctx, mode == PPX_SOLVE_CHOL || mode == PPX_SOLVE_SVD_PINV, "mode is CHOL or SVD_PINV");
  ppx_ws_reset(ctx);
  double *Sinv = (double *)ppx_ws_alloc(ctx, sizeof(double) * R * R);
  double *Sws = (double *)ppx_ws_alloc(ctx, sizeof(double) * R * R);
  if (!Sinv || !Sws) return ppx_set_err(ctx, PPX_ENOMEM, "workspace too small");
  // S is needed by the apply kernel only for the gradient
  int rc = inverse_launch(ctx, h, R, lambda, mode, (grad_out && !S_given) ? Sws : nullptr, Sinv);
  if (rc) return rc;
  const double *S_apply = S_given ? S_given : Sws;
  const size_t smem = sizeof(double) * (2 * (size_t)R * R + 2 * AP_ROWS * (size_t)R);
  if (smem > 220 * 1024) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "solve: R=%d too large", R);
  solve_apply_kernel<<<ppx_cdiv(s, AP_ROWS), dim3(32, AP_ROWS), smem, ctx->stream>>>(M, S_apply, Sinv, W, s, R, W_init,
                                                                                    ratio_step, grad_out, dW_out);
  PPX_CHECK_LAUNCH(ctx);
  if (sq_norms_out) {
    const double *xs[3] = {W, dW_out ? dW_out : W, grad_out ? grad_out : W};
    int64_t ns[3] = {s * R, dW_out ? s * R : 0, grad_out ? s * R : 0};
    return ppx_sqnorms(ctx, xs, ns, 3, sq_norms_out);
  }
  return PPX_OK;
}

// ---- small dense products of the low-rank-update optimizers (operands of a few hundred x R; latency-size work) -----
// C = alpha op(A) op(B) + beta C, column-major, one thread per element of C
__global__ void __launch_bounds__(256) gemm_small_kernel(int ta, int tb, int m, int n, int k, double alpha,
                                                         const double *__restrict__ A, int64_t lda,
                                                         const double *__restrict__ B, int64_t ldb, double beta,
                                                         double *__restrict__ C, int64_t ldc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)m * n) return;
  const int i = (int)(idx % m), j = (int)(idx / m);
  double acc = 0.0;
  for (int q = 0; q < k; q++) {
    const double a = ta ? A[q + lda * i] : A[i + lda * q];
    const double b = tb ? B[j + ldb * q] : B[q + ldb * j];
    acc = fma(a, b, acc);
  }
  double *c = C + i + ldc * j;
  *c = beta == 0.0 ? alpha * acc : alpha * acc + beta * *c;
}

// out[m, c] += sum_q T[m, q] VT[q, c]:  the rank-r patch of a cached root tensor expanded to R columns
// (cp_dt_lr_optimizer.cxx:152-158).  HBM-bound read-modify-write of the Mtot x R tensor; VT lives in shared memory.
constexpr int RX_MAX_r = 16;
__global__ void __launch_bounds__(256) rank_expand_acc_kernel(const double *__restrict__ T, int64_t Mtot, int r,
                                                              const double *__restrict__ VT, int64_t ldvt, int R,
                                                              double *__restrict__ out) {
  extern __shared__ double vt[];  // [r][R]
  for (int i = threadIdx.x; i < r * R; i += blockDim.x) vt[i] = VT[(i / R) + ldvt * (i % R)];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < Mtot; m += stride) {
    double t[RX_MAX_r];
#pragma unroll
    for (int q = 0; q < RX_MAX_r; q++) t[q] = q < r ? T[m + Mtot * q] : 0.0;
    for (int c = 0; c < R; c++) {
      double acc = out[m + Mtot * c];
#pragma unroll
      for (int q = 0; q < RX_MAX_r; q++)
        if (q < r) acc = fma(t[q], vt[q * R + c], acc);
      out[m + Mtot * c] = acc;
    }
  }
}

}  // namespace

int ppx_k45_init(ppx_ctx *ctx) {
  const int big = 220 * 1024;
  PPX_CUDA(ctx, cudaFuncSetAttribute(spd_inverse_ldl_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(spd_inverse_ldl_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(spd_inverse_ldl_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(spd_inverse_ldl_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(sym_inverse_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  PPX_CUDA(ctx, cudaFuncSetAttribute(solve_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  return PPX_OK;
}

static int normalize_impl(ppx_ctx *ctx, double *const *W, const int64_t *s, int N, int R, double *const *G,
                          bool from_grams) {
  PPX_REQUIRE(ctx, W && s && N >= 1 && N <= 16, "1 <= N <= 16");
  NormArgs a;
  a.N = N;
  a.R = R;
  int64_t nmax = 0;
  for (int i = 0; i < N; i++) {
    a.w[i] = W[i];
    a.g[i] = G ? G[i] : nullptr;
    a.n[i] = s[i] * R;
    if (a.n[i] > nmax) nmax = a.n[i];
  }
  ppx_ws_reset(ctx);
  double *sq = (double *)ppx_ws_alloc(ctx, sizeof(double) * 16);
  if (!sq) return ppx_set_err(ctx, PPX_ENOMEM, "workspace too small");
  if (from_grams)
    norm_sq_from_gram_kernel<<<N, 32, 0, ctx->stream>>>(a, sq);
  else
    norm_sq_kernel<<<N, 1024, 0, ctx->stream>>>(a, sq);
  PPX_CHECK_LAUNCH(ctx);
  int bx = ppx_cdiv(nmax, 256 * 4);
  if (bx < 1) bx = 1;
  if (bx > ctx->sm_count * 4) bx = ctx->sm_count * 4;
  norm_scale_kernel<<<dim3(bx, N), 256, 0, ctx->stream>>>(a, sq);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

extern "C" {

int ppx_gram(ppx_ctx *ctx, const double *W, int64_t s, int64_t ldw, int R, double *G) {
  PPX_REQUIRE(ctx, W && G && s >= 0 && R >= 1 && ldw >= s, "W, G non-null; ldw >= s; R >= 1");
  const int npairs = R * (R + 1) / 2;
  const int blocks = (npairs * 32 + 255) / 256;
  gram_kernel<<<blocks, 256, 0, ctx->stream>>>(W, s, ldw, R, G);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_hadamard_grams(ppx_ctx *ctx, const double *const *G, int nG, int skip, int R, double lambda, double *S) {
  PPX_REQUIRE(ctx, G && S && nG >= 1 && nG <= 16, "1 <= nG <= 16");
  HadArgs h;
  h.n = 0;
  for (int j = 0; j < nG; j++)
    if (j != skip) h.g[h.n++] = G[j];
  PPX_REQUIRE(ctx, h.n >= 1, "at least one Gram after skipping");
  hadamard_kernel<<<(R * R + 255) / 256, 256, 0, ctx->stream>>>(h, R, lambda, S);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_solve_update(ppx_ctx *ctx, const double *M, const double *S, double *W, int64_t s, int R,
                     const double *W_init, double ratio_step, int mode, double *grad_out, double *dW_out,
                     double *sq_norms_out) {
  PPX_REQUIRE(ctx, S != nullptr, "S != NULL");
  HadArgs h;
  h.n = 1;
  h.g[0] = S;
  return solve_impl(ctx, M, h, 0.0, S, W, s, R, W_init, ratio_step, mode, grad_out, dW_out, sq_norms_out);
}

int ppx_solve_update_g(ppx_ctx *ctx, const double *M, const double *const *G, int nG, int skip, double lambda,
                       double *W, int64_t s, int R, const double *W_init, double ratio_step, int mode,
                       double *grad_out, double *dW_out, double *sq_norms_out) {
  PPX_REQUIRE(ctx, G && nG >= 1 && nG <= 16, "1 <= nG <= 16");
  HadArgs h;
  h.n = 0;
  for (int j = 0; j < nG; j++)
    if (j != skip) h.g[h.n++] = G[j];
  PPX_REQUIRE(ctx, h.n >= 1, "at least one Gram after skipping");
  return solve_impl(ctx, M, h, lambda, nullptr, W, s, R, W_init, ratio_step, mode, grad_out, dW_out, sq_norms_out);
}

int ppx_spd_inverse_g(ppx_ctx *ctx, const double *const *G, int nG, int skip, double lambda, int R, int mode,
                      double *S_out, double *Sinv_out) {
  PPX_REQUIRE(ctx, G && Sinv_out && nG >= 1 && nG <= 16 && R >= 1, "1 <= nG <= 16, R >= 1, Sinv_out != NULL");
  PPX_REQUIRE(ctx, mode == PPX_SOLVE_CHOL || mode == PPX_SOLVE_SVD_PINV, "mode is CHOL or SVD_PINV");
  HadArgs h;
  h.n = 0;
  for (int j = 0; j < nG; j++)
    if (j != skip) h.g[h.n++] = G[j];
  PPX_REQUIRE(ctx, h.n >= 1, "at least one Gram after skipping");
  return inverse_launch(ctx, h, R, lambda, mode, S_out, Sinv_out);
}

int ppx_solve_apply(ppx_ctx *ctx, const double *M, const double *S, const double *Sinv, double *W, int64_t s, int R,
                    const double *W_init, double ratio_step, double *grad_out, double *dW_out) {
  PPX_REQUIRE(ctx, M && Sinv && W && s >= 1 && R >= 1, "M, Sinv, W non-null; s, R >= 1");
  PPX_REQUIRE(ctx, S || !grad_out, "S is required when the gradient is requested");
  const size_t smem = sizeof(double) * (2 * (size_t)R * R + 2 * AP_ROWS * (size_t)R);
  if (smem > 220 * 1024) return ppx_set_err(ctx, PPX_EUNSUPPORTED, "solve: R=%d too large", R);
  solve_apply_kernel<<<ppx_cdiv(s, AP_ROWS), dim3(32, AP_ROWS), smem, ctx->stream>>>(M, S, Sinv, W, s, R, W_init,
                                                                                    ratio_step, grad_out, dW_out);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_normalize_norms(ppx_ctx *ctx, double *const *W, const double *const *dW, const int64_t *s, int N, int R,
                        double *const *G, double *sq_out_dev) {
  PPX_REQUIRE(ctx, W && s && G && sq_out_dev && N >= 1 && N <= 16, "W, s, G, sq_out_dev non-null; 1 <= N <= 16");
  NormNormsArgs a;
  a.N = N;
  a.R = R;
  for (int i = 0; i < N; i++) {
    PPX_REQUIRE(ctx, W[i] && G[i], "W[i], G[i] non-null");
    a.w[i] = W[i];
    a.g[i] = G[i];
    a.dw[i] = dW ? dW[i] : nullptr;
    a.n[i] = s[i] * R;
  }
  NormArgs t;
  t.N = N;
  t.R = R;
  for (int i = 0; i < N; i++) t.g[i] = G[i];
  ppx_ws_reset(ctx);
  double *tr = (double *)ppx_ws_alloc(ctx, sizeof(double) * 16);
  if (!tr) return ppx_set_err(ctx, PPX_ENOMEM, "workspace too small");
  norm_sq_from_gram_kernel<<<N, 32, 0, ctx->stream>>>(t, tr);
  PPX_CHECK_LAUNCH(ctx);
  normalize_norms_kernel<<<N, 1024, 0, ctx->stream>>>(a, tr, sq_out_dev);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_normalize(ppx_ctx *ctx, double *const *W, const int64_t *s, int N, int R, double *const *G) {
  return normalize_impl(ctx, W, s, N, R, G, false);
}

int ppx_normalize_g(ppx_ctx *ctx, double *const *W, const int64_t *s, int N, int R, double *const *G) {
  PPX_REQUIRE(ctx, G != nullptr, "G != NULL");
  return normalize_impl(ctx, W, s, N, R, G, true);
}

int ppx_spd_factor_inverse(ppx_ctx *ctx, const double *S, int R, double *Linv_out) {
  PPX_REQUIRE(ctx, S && Linv_out && R >= 1, "S, Linv_out non-null; R >= 1");
  HadArgs h;
  h.n = 1;
  h.g[0] = S;
  return inverse_launch(ctx, h, R, 0.0, PPX_SOLVE_CHOL, nullptr, nullptr, Linv_out);
}

int ppx_gemm_small(ppx_ctx *ctx, int transa, int transb, int m, int n, int k, double alpha, const double *A,
                   int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc) {
  PPX_REQUIRE(ctx, C && (k == 0 || (A && B)) && m >= 0 && n >= 0 && k >= 0, "non-null operands, non-negative sizes");
  if (m == 0 || n == 0) return PPX_OK;
  gemm_small_kernel<<<ppx_cdiv((int64_t)m * n, 256), 256, 0, ctx->stream>>>(transa, transb, m, n, k, alpha, A, lda, B,
                                                                         ldb, beta, C, ldc);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_rank_expand_acc(ppx_ctx *ctx, const double *T, int64_t Mtot, int r, const double *VT, int64_t ldvt, int R,
                        double *out) {
  PPX_REQUIRE(ctx, T && VT && out && Mtot >= 0 && R >= 1, "non-null operands");
  PPX_REQUIRE(ctx, r >= 1 && r <= RX_MAX_r, "1 <= r <= 16");
  if (Mtot == 0) return PPX_OK;
  int64_t blocks = (Mtot + 255) / 256;
  if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
  rank_expand_acc_kernel<<<(int)blocks, 256, sizeof(double) * (size_t)r * R, ctx->stream>>>(T, Mtot, r, VT, ldvt, R, out);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

}  // extern "C"
