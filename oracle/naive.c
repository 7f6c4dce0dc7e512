/* Independent loop-based CPU restatement of the contractions on the ALS / pairwise-perturbation hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/pp_oracle.py header: the oracle is pinned against the reference's own sources
 * built on a CTF stand-in, oracle/_ref/; CTF's own arithmetic stays unpinned).  This file shares no code with pp_oracle.py:
 * plain loops over the first-index-fastest buffers with long double accumulators; tests/test_oracle_cpu.py checks
 * the two against each other and against tests/golden/.  Links nothing but libm.
 *
 * All tensors are dense FP64, first index fastest (the reference's CTF global order);
 * factor matrices are s x R column-major.
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

/* u(seed, tensor_id, linear_index) in [0,1): SplitMix64 finaliser, top 53 bits (same as pp_oracle.u01 and the
 * CUDA generator ppx_fill_uniform). */
double naive_u01(uint64_t seed, uint64_t tensor_id, uint64_t idx) {
  uint64_t z = idx + seed * 0x9E3779B97F4A7C15ULL + tensor_id * 0xD1B54A32D192ED03ULL;
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

void naive_fill_uniform(double *out, int64_t n, uint64_t seed, uint64_t tensor_id, double lo, double hi) {
  for (int64_t i = 0; i < n; i++) out[i] = lo + (hi - lo) * naive_u01(seed, tensor_id, (uint64_t)i);
}

static void split3(const int64_t *lens, int k, int x, int64_t *L, int64_t *X, int64_t *Rt) {
  *L = 1; *Rt = 1;
  for (int i = 0; i < x; i++) *L *= lens[i];
  for (int i = x + 1; i < k; i++) *Rt *= lens[i];
  *X = lens[x];
}

/* First tensor-times-matrix contraction  out[rest, r] = sum_x V[..x..] W[x, r]
 * (reference: common.cxx:56, als_CP.cxx:378-379, cp_dt_optimizer.cxx:158-159). */
void naive_ttm_first(const double *V, const int64_t *lens, int N, int x, const double *W, int R, double *out) {
  int64_t L, X, Rt;
  split3(lens, N, x, &L, &X, &Rt);
  for (int r = 0; r < R; r++)
    for (int64_t t = 0; t < Rt; t++)
      for (int64_t l = 0; l < L; l++) {
        long double acc = 0;
        for (int64_t k = 0; k < X; k++) acc += (long double)V[l + L * (k + X * t)] * W[k + X * r];
        out[l + L * (t + Rt * (int64_t)r)] = (double)acc;
      }
}

/* Hadamard-batched contraction  out[rest', r] = sum_x T[..x.., r] W[x, r]; `lens` are the k non-rank modes of T
 * (reference: common.cxx:83,128; als_CP.cxx:258-259,407-408). */
void naive_mttv(const double *T, const int64_t *lens, int k, int x, const double *W, int R, double *out) {
  int64_t L, X, Rt;
  split3(lens, k, x, &L, &X, &Rt);
  for (int r = 0; r < R; r++)
    for (int64_t t = 0; t < Rt; t++)
      for (int64_t l = 0; l < L; l++) {
        long double acc = 0;
        for (int64_t j = 0; j < X; j++) acc += (long double)T[l + L * (j + X * (t + Rt * (int64_t)r))] * W[j + X * r];
        out[l + L * (t + Rt * (int64_t)r)] = (double)acc;
      }
}

/* Tucker tensor-times-matrix, rank replaces mode x in place: out[..q..] = sum_x T[..x..] W[x, q]
 * (reference: als_Tucker.cxx:102, 224, 464-465). */
void naive_ttm(const double *T, const int64_t *lens, int k, int x, const double *W, int Q, double *out) {
  int64_t L, X, Rt;
  split3(lens, k, x, &L, &X, &Rt);
  for (int64_t t = 0; t < Rt; t++)
    for (int q = 0; q < Q; q++)
      for (int64_t l = 0; l < L; l++) {
        long double acc = 0;
        for (int64_t j = 0; j < X; j++) acc += (long double)T[l + L * (j + X * t)] * W[j + X * q];
        out[l + L * (q + (int64_t)Q * t)] = (double)acc;
      }
}

/* Gram G = W^T W (reference: als_CP.cxx:288). */
void naive_gram(const double *W, int64_t s, int R, double *G) {
  for (int a = 0; a < R; a++)
    for (int b = 0; b < R; b++) {
      long double acc = 0;
      for (int64_t i = 0; i < s; i++) acc += (long double)W[i + s * a] * W[i + s * b];
      G[a + R * b] = (double)acc;
    }
}

/* PP first-order correction  M = M0 + sum_j op_j (x) dW_j  (reference: als_CP.cxx:778-794).
 * op_j is s_a x s_b x R; which[j]==0 contracts the FIRST operator index (operator's second index is the output
 * row, als_CP.cxx:785), which[j]==1 contracts the SECOND (first index is the output row, :793). */
void naive_pp_correct(const double *M0, const double *const *ops, const int *which, const double *const *dW,
                      const int64_t *s_other, int n_ops, int64_t s_i, int R, double *M) {
  for (int r = 0; r < R; r++)
    for (int64_t i = 0; i < s_i; i++) {
      long double acc = M0[i + s_i * r];
      for (int j = 0; j < n_ops; j++) {
        int64_t sj = s_other[j];
        const double *P = ops[j];
        const double *d = dW[j];
        if (which[j] == 0) {
          for (int64_t q = 0; q < sj; q++) acc += (long double)P[q + sj * (i + s_i * (int64_t)r)] * d[q + sj * r];
        } else {
          for (int64_t q = 0; q < sj; q++) acc += (long double)P[i + s_i * (q + sj * (int64_t)r)] * d[q + sj * r];
        }
      }
      M[i + s_i * r] = (double)acc;
    }
}

/* ||V - [[W_0..W_{N-1}]]||_F  (reference: common.cxx:135-197 + als_CP.cxx:183-187); W packed back to back. */
double naive_cp_residual(const double *V, const int64_t *lens, int N, const double *const *W, int R) {
  int64_t P = 1;
  for (int i = 0; i < N; i++) P *= lens[i];
  long double ss = 0;
  int64_t idx[16];
  for (int64_t p = 0; p < P; p++) {
    int64_t q = p;
    for (int i = 0; i < N; i++) { idx[i] = q % lens[i]; q /= lens[i]; }
    long double v = 0;
    for (int r = 0; r < R; r++) {
      long double t = 1;
      for (int i = 0; i < N; i++) t *= W[i][idx[i] + lens[i] * r];
      v += t;
    }
    long double d = (long double)V[p] - v;
    ss += d * d;
  }
  return (double)sqrtl(ss);
}

/* Gram of the mode-i unfolding  MTM[p,q] = sum_rest T[..p..] T[..q..]  (reference: common.cxx:205-223). */
void naive_unfold_gram(const double *T, const int64_t *lens, int k, int i, double *MTM) {
  int64_t L, X, Rt;
  split3(lens, k, i, &L, &X, &Rt);
  for (int64_t p = 0; p < X; p++)
    for (int64_t q = 0; q < X; q++) {
      long double acc = 0;
      for (int64_t t = 0; t < Rt; t++)
        for (int64_t l = 0; l < L; l++) acc += (long double)T[l + L * (p + X * t)] * T[l + L * (q + X * t)];
      MTM[p + X * q] = (double)acc;
    }
}
