// Internal definitions shared by the kernels behind include/ppx.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include "../../include/ppx.h"

struct ppx_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;       // the stream operators are enqueued on (main, or side between side_begin/end)
  cudaStream_t main_stream = nullptr;
  // lanes: streams for work that may overlap the main stream (ppx_lane_*; ppx_side_* is lane 0)
  static constexpr int N_LANES = 4;
  cudaStream_t lane_stream[N_LANES] = {};
  cudaEvent_t ev_fork[N_LANES] = {}, ev_join[N_LANES] = {};
  int open_lane = -1;  // the lane `stream` points at, -1: the main stream
  bool own_stream = false;
  char *ws = nullptr;  // workspace arena (bump allocated per call)
  size_t ws_bytes = 0;
  size_t ws_used = 0;
  int64_t launches = 0;
  std::string err;
  // pinned staging for pointer tables handed to kernels
  void *comm = nullptr;  // ncclComm_t
  int nranks = 1, rank = 0;
  void *p2p = nullptr;   // peer-memory state of the one-shot all-reduce (comm.cu), nullptr: NCCL only
};

int ppx_set_err(ppx_ctx *ctx, int code, const char *fmt, ...);
// per-file one-time kernel attribute setup (opt-in shared memory), called from ppx_ctx_create
int ppx_k1_init(ppx_ctx *ctx);
int ppx_k1_tma_init(ppx_ctx *ctx);
// TMA path of the first contraction; returns 1 when the shape is not eligible (caller falls back to cp.async)
int ppx_ttm_tma_try(ppx_ctx *ctx, const double *V, int64_t L, int64_t K, int64_t Rt, const double *const *fac,
                    const int64_t *ld, const int64_t *xs, int n_fac, int R, double *out, int inplace, int accumulate,
                    bool ws_keep);
// streaming variant of the first contraction for X <= 64, R <= 16 (k1_ttm_stream.cu); returns 1 when not eligible
int ppx_ttm_stream_try(ppx_ctx *ctx, const double *V, int64_t L, int64_t X, int64_t Rt, const double *W, int64_t ldw,
                       int R, double *out, int inplace, int accumulate);
// out[l,t,r] = sum_x V[l,x,t] W[x,r] (rank last, or in place of mode x); shared by the CP and Tucker entry points
int ppx_ttm_impl(ppx_ctx *ctx, const double *V, int64_t L, int64_t X, int64_t Rt, const double *Wx, int64_t ldw, int R,
                 double *out, int inplace, int accumulate, bool ws_keep, bool try_tma);
// Chebyshev-filtered subspace iteration for the leading eigenvectors (eig_chfsi.cu); see there for the contract
bool ppx_eig_chfsi_applicable(int64_t n, int r);
int ppx_eig_chfsi(ppx_ctx *ctx, double *A, int n, int r, double *U, double *evals_out, double *state, int state_valid);
// DMMA SYRK for large unfoldings (gram_dmma.cu): partial Grams, returns the number of K splits written (0: not taken)
int ppx_gram_dmma(ppx_ctx *ctx, const double *T, int64_t L, int64_t X, int64_t Rt, double *parts, int max_splits);
int ppx_gram_init(ppx_ctx *ctx);
int ppx_k45_init(ppx_ctx *ctx);
int ppx_k7_init(ppx_ctx *ctx);
int ppx_k2x3_init(ppx_ctx *ctx);
void ppx_comm_destroy_internal(ppx_ctx *ctx);
int ppx_sum_partials(ppx_ctx *ctx, const double *partial, int n, double *out);

#define PPX_CUDA(ctx, expr)                                                                            \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return ppx_set_err(ctx, PPX_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),      \
                         __FILE__, __LINE__);                                                          \
  } while (0)

#define PPX_CHECK_LAUNCH(ctx)                                                                          \
  do {                                                                                                 \
    (ctx)->launches++;                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                                              \
    if (e__ != cudaSuccess)                                                                            \
      return ppx_set_err(ctx, PPX_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),  \
                         __FILE__, __LINE__);                                                          \
  } while (0)

#define PPX_REQUIRE(ctx, cond, msg)                                                                    \
  do {                                                                                                 \
    if (!(cond)) return ppx_set_err(ctx, PPX_EINVAL, "%s: requirement failed: %s", __func__, msg);     \
  } while (0)

// workspace: reset at the start of an API call that needs scratch, then bump.
static inline void ppx_ws_reset(ppx_ctx *ctx) { ctx->ws_used = 0; }
void *ppx_ws_alloc(ppx_ctx *ctx, size_t bytes);  // nullptr if exhausted (256-byte aligned)

// split a k-mode tensor around mode x:  index = l + L*(j + X*t)
static inline void ppx_split3(const int64_t *lens, int k, int x, int64_t *L, int64_t *X, int64_t *Rt) {
  int64_t l = 1, r = 1;
  for (int i = 0; i < x; i++) l *= lens[i];
  for (int i = x + 1; i < k; i++) r *= lens[i];
  *L = l;
  *X = lens[x];
  *Rt = r;
}

static inline int ppx_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Number of K splits for the first-contraction kernels: `num_tiles` row tiles of `nk` 16-deep chunks on a persistent
// grid of G CTAs, output of `out_elems` doubles per split.  Cost model in units of one chunk step of a CTA (~1.75 us
// with two CTAs per SM): rounds of the grid x chunks per split, plus writing and re-reading the partial outputs at
// ~6 TB/s; a split must keep >= 16 chunks and its partial buffer must fit in the free workspace.  (A plain "best grid
// efficiency" rule picked 296 splits for 87 tiles -- efficiency exactly 1 -- whose 1.3 GB of partials did not fit.)
static inline int ppx_pick_ksplit(int num_tiles, int nk, int G, int64_t out_elems, size_t ws_free) {
  if (const char *f = getenv("PPX_KSPLIT"))  // experiments only
    if (atoi(f) > 0) return atoi(f);
  int best = 1;
  double best_cost = 1e300;
  for (int S = 1; S <= 512 && (S == 1 || nk / S >= 16); S++) {
    if (S > 1 && (double)S * (double)out_elems * 8.0 + 4096.0 > (double)ws_free) break;
    const int64_t units = (int64_t)num_tiles * S;
    const double rounds = (double)((units + G - 1) / G);
    const double cps = (double)((nk + S - 1) / S);
    const double traffic_us = S > 1 ? (double)S * (double)out_elems * 16.0 / 6.0e6 : 0.0;
    const double cost = rounds * cps + traffic_us / 1.75 + (S > 1 ? 4.0 : 0.0);  // + the reduce kernel's launch
    if (cost < best_cost * 0.995) {  // prefer the smaller split on near ties
      best_cost = cost;
      best = S;
    }
  }
  return best;
}

// ---- device helpers -------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double ppx_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block reduction; result valid in thread 0.  `red` is >= 32 doubles of shared memory.
__device__ __forceinline__ double ppx_block_sum(double v, double *red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = ppx_warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (w == 0) v = ppx_warp_sum(v);
  return v;
}

__device__ __forceinline__ void ppx_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void ppx_cp_async16(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void ppx_cp_async8(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void ppx_cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void ppx_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
#endif
