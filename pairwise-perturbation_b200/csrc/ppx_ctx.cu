// Context, memory, events/graphs, generator and small elementwise / reduction kernels behind include/ppx.h.
#include <stdarg.h>
#include <string.h>
#include "ppx_internal.h"

int ppx_set_err(ppx_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else fprintf(stderr, "ppx: %s\n", buf);
  return code;
}

void *ppx_ws_alloc(ppx_ctx *ctx, size_t bytes) {
  size_t off = (ctx->ws_used + 255) & ~(size_t)255;
  if (off + bytes > ctx->ws_bytes) return nullptr;
  ctx->ws_used = off + bytes;
  return ctx->ws + off;
}

extern "C" {

const char *ppx_version(void) { return "ppx 0.1 (sm_100a)"; }

int ppx_ctx_create(int device, void *stream, size_t workspace_bytes, ppx_ctx **out) {
  if (!out) return PPX_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return ppx_set_err(nullptr, PPX_ECUDA, "no CUDA device (%s); there is no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return ppx_set_err(nullptr, PPX_EINVAL, "device %d out of range", device);
  ppx_ctx *ctx = new ppx_ctx();
  ctx->device = device;
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete ctx; return ppx_set_err(nullptr, PPX_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e)); }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->sm_count = prop.multiProcessorCount;
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return ppx_set_err(nullptr, PPX_ECUDA, "stream: %s", cudaGetErrorString(e)); }
    ctx->own_stream = true;
  }
  ctx->main_stream = ctx->stream;
  e = cudaSuccess;
  for (int l = 0; l < ppx_ctx::N_LANES && e == cudaSuccess; l++) {
    e = cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork[l], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join[l], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) { delete ctx; return ppx_set_err(nullptr, PPX_ECUDA, "side stream: %s", cudaGetErrorString(e)); }
  if (workspace_bytes < ((size_t)8 << 20)) workspace_bytes = (size_t)8 << 20;
  e = cudaMalloc((void **)&ctx->ws, workspace_bytes);
  if (e != cudaSuccess) { delete ctx; return ppx_set_err(nullptr, PPX_ENOMEM, "workspace: %s", cudaGetErrorString(e)); }
  ctx->ws_bytes = workspace_bytes;
  {
    int rc = ppx_k1_init(ctx);
    if (!rc) rc = ppx_k1_tma_init(ctx);
    if (!rc) rc = ppx_k45_init(ctx);
    if (!rc) rc = ppx_k7_init(ctx);
    if (!rc) rc = ppx_k2x3_init(ctx);
    if (!rc) rc = ppx_gram_init(ctx);
    if (rc) {
      fprintf(stderr, "ppx: kernel init failed: %s\n", ctx->err.c_str());
      cudaFree(ctx->ws);
      delete ctx;
      return rc;
    }
  }
  *out = ctx;
  return PPX_OK;
}

int ppx_ctx_destroy(ppx_ctx *ctx) {
  if (!ctx) return PPX_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->main_stream);
  for (int l = 0; l < ppx_ctx::N_LANES; l++)
    if (ctx->lane_stream[l]) cudaStreamSynchronize(ctx->lane_stream[l]);
  ppx_comm_destroy_internal(ctx);
  if (ctx->ws) cudaFree(ctx->ws);
  for (int l = 0; l < ppx_ctx::N_LANES; l++) {
    if (ctx->lane_stream[l]) cudaStreamDestroy(ctx->lane_stream[l]);
    if (ctx->ev_fork[l]) cudaEventDestroy(ctx->ev_fork[l]);
    if (ctx->ev_join[l]) cudaEventDestroy(ctx->ev_join[l]);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->main_stream);
  delete ctx;
  return PPX_OK;
}

int ppx_sync(ppx_ctx *ctx) {
  PPX_CUDA(ctx, cudaStreamSynchronize(ctx->main_stream));
  return PPX_OK;
}
// Fork/join for work that may overlap the main stream (works eagerly and inside a graph capture):
//   ppx_lane_begin(l): lane l waits for everything enqueued on the main stream so far; later calls go to lane l
//   ppx_lane_end      : later calls go to the main stream again (the lane's work keeps running concurrently)
//   ppx_lane_join(l)  : the main stream waits for the work enqueued on lane l up to its last ppx_lane_end
//   ppx_side_*        : lane 0
int ppx_lane_begin(ppx_ctx *ctx, int lane) {
  PPX_REQUIRE(ctx, lane >= 0 && lane < ppx_ctx::N_LANES, "0 <= lane < 4");
  PPX_REQUIRE(ctx, ctx->stream == ctx->main_stream, "ppx_lane_begin inside an open lane");
  PPX_CUDA(ctx, cudaEventRecord(ctx->ev_fork[lane], ctx->main_stream));
  PPX_CUDA(ctx, cudaStreamWaitEvent(ctx->lane_stream[lane], ctx->ev_fork[lane], 0));
  ctx->stream = ctx->lane_stream[lane];
  ctx->open_lane = lane;
  return PPX_OK;
}
int ppx_lane_end(ppx_ctx *ctx) {
  if (ctx->open_lane >= 0) PPX_CUDA(ctx, cudaEventRecord(ctx->ev_join[ctx->open_lane], ctx->lane_stream[ctx->open_lane]));
  ctx->open_lane = -1;
  ctx->stream = ctx->main_stream;
  return PPX_OK;
}
int ppx_lane_join(ppx_ctx *ctx, int lane) {
  PPX_REQUIRE(ctx, lane >= 0 && lane < ppx_ctx::N_LANES, "0 <= lane < 4");
  PPX_CUDA(ctx, cudaStreamWaitEvent(ctx->main_stream, ctx->ev_join[lane], 0));
  return PPX_OK;
}
// diagnostics: the GPU's nanosecond clock into a device slot, on the current stream (a timeline of a captured sweep
// without a profiler: PPX_PP_TRACE=1 in the PP phase of host/als_CP.cxx)
__global__ void stamp_kernel(unsigned long long *slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}
int ppx_stamp(ppx_ctx *ctx, unsigned long long *dev_slot) {
  PPX_REQUIRE(ctx, dev_slot != nullptr, "dev_slot != NULL");
  stamp_kernel<<<1, 1, 0, ctx->stream>>>(dev_slot);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}
int ppx_side_begin(ppx_ctx *ctx) { return ppx_lane_begin(ctx, 0); }
int ppx_side_end(ppx_ctx *ctx) { return ppx_lane_end(ctx); }
int ppx_side_join(ppx_ctx *ctx) { return ppx_lane_join(ctx, 0); }
const char *ppx_last_error(ppx_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }
void *ppx_stream(ppx_ctx *ctx) { return (void *)ctx->main_stream; }
int ppx_device(ppx_ctx *ctx) { return ctx->device; }
int ppx_sm_count(ppx_ctx *ctx) { return ctx->sm_count; }
int64_t ppx_launch_count(ppx_ctx *ctx) { return ctx->launches; }

int ppx_malloc(ppx_ctx *ctx, size_t bytes, void **dptr) {
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 8);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return ppx_set_err(ctx, PPX_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  }
  return PPX_OK;
}
int ppx_free(ppx_ctx *ctx, void *dptr) {
  if (dptr) PPX_CUDA(ctx, cudaFree(dptr));
  return PPX_OK;
}
int ppx_host_alloc(ppx_ctx *ctx, size_t bytes, void **hptr) {
  PPX_CUDA(ctx, cudaMallocHost(hptr, bytes ? bytes : 8));
  return PPX_OK;
}
int ppx_host_free(ppx_ctx *ctx, void *hptr) {
  if (hptr) PPX_CUDA(ctx, cudaFreeHost(hptr));
  return PPX_OK;
}
int ppx_memcpy_h2d(ppx_ctx *ctx, void *dst, const void *src, size_t bytes) {
  PPX_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return PPX_OK;
}
int ppx_memcpy_d2h(ppx_ctx *ctx, void *dst, const void *src, size_t bytes) {
  PPX_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return PPX_OK;
}
int ppx_memcpy_d2d(ppx_ctx *ctx, void *dst, const void *src, size_t bytes) {
  PPX_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return PPX_OK;
}
int ppx_memcpy2d_d2d(ppx_ctx *ctx, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes,
                     size_t height) {
  if (width_bytes == 0 || height == 0) return PPX_OK;
  PPX_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, cudaMemcpyDeviceToDevice,
                                  ctx->stream));
  return PPX_OK;
}
int ppx_memset_zero(ppx_ctx *ctx, void *dst, size_t bytes) {
  PPX_CUDA(ctx, cudaMemsetAsync(dst, 0, bytes, ctx->stream));
  return PPX_OK;
}
int ppx_mem_info(ppx_ctx *ctx, size_t *free_bytes, size_t *total_bytes) {
  PPX_CUDA(ctx, cudaMemGetInfo(free_bytes, total_bytes));
  return PPX_OK;
}

int ppx_event_create(ppx_ctx *ctx, void **ev) {
  cudaEvent_t e;
  PPX_CUDA(ctx, cudaEventCreate(&e));
  *ev = (void *)e;
  return PPX_OK;
}
int ppx_event_destroy(ppx_ctx *ctx, void *ev) {
  PPX_CUDA(ctx, cudaEventDestroy((cudaEvent_t)ev));
  return PPX_OK;
}
int ppx_event_record(ppx_ctx *ctx, void *ev) {
  PPX_CUDA(ctx, cudaEventRecord((cudaEvent_t)ev, ctx->main_stream));
  return PPX_OK;
}
int ppx_event_elapsed_ms(ppx_ctx *ctx, void *a, void *b, float *ms) {
  PPX_CUDA(ctx, cudaEventSynchronize((cudaEvent_t)b));
  PPX_CUDA(ctx, cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return PPX_OK;
}
int ppx_graph_begin(ppx_ctx *ctx) {
  PPX_CUDA(ctx, cudaStreamBeginCapture(ctx->main_stream, cudaStreamCaptureModeThreadLocal));
  return PPX_OK;
}
int ppx_graph_end(ppx_ctx *ctx, void **graph) {
  cudaGraph_t g;
  PPX_CUDA(ctx, cudaStreamEndCapture(ctx->main_stream, &g));
  cudaGraphExec_t ge;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return ppx_set_err(ctx, PPX_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  *graph = (void *)ge;
  return PPX_OK;
}
int ppx_graph_launch(ppx_ctx *ctx, void *graph) {
  PPX_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)graph, ctx->main_stream));
  return PPX_OK;
}
int ppx_graph_destroy(ppx_ctx *ctx, void *graph) {
  if (graph) PPX_CUDA(ctx, cudaGraphExecDestroy((cudaGraphExec_t)graph));
  return PPX_OK;
}

}  // extern "C"

// ---- Laplacian (Poisson) tensor of the reference's generator 'p' / 'p2' (common.cxx:575-642) --------------------
// Order 2d, extents s: V[a1,b1,...,ad,bd] = sum_k D[a_k,b_k] prod_{m != k} delta(a_m,b_m), D = tridiag(-1, 2, -1).
// Closed form per element: no pair off the diagonal -> 2d; exactly one pair (a,b) off the diagonal -> D[a,b]
// (-1 if |a-b| == 1 else 0); two or more -> 0.
__global__ void fill_laplacian_kernel(double *out, int64_t n, int d, int64_t s) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t r = i;
    int off = 0;
    double v = 0.0;
    for (int m = 0; m < d; m++) {
      const int64_t a = r % s;
      r /= s;
      const int64_t b = r % s;
      r /= s;
      if (a != b) {
        off++;
        v = (a - b == 1 || b - a == 1) ? -1.0 : 0.0;
      }
    }
    out[i] = off == 0 ? 2.0 * d : (off == 1 ? v : 0.0);
  }
}

// ---- generator ----------------------------------------------------------------------------------------------
__global__ void fill_uniform_kernel(double *out, int64_t n, uint64_t base, int64_t start, double lo, double span) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint64_t z = (uint64_t)(start + i) + base;
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    out[i] = lo + span * ((double)(z >> 11) * (1.0 / 9007199254740992.0));
  }
}

// rows [row_begin, row_begin + L_local) of every column of a (L_global x n_cols) first-index-fastest array: the local
// slab of a tensor whose leading mode is sharded, drawn from the same global counter
__global__ void fill_uniform_rows_kernel(double *out, int64_t L_local, int64_t L_global, int64_t row_begin,
                                         int64_t n_cols, uint64_t base, double lo, double span) {
  const int64_t n = L_local * n_cols;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int64_t c = i / L_local, r = i - c * L_local;
    uint64_t z = (uint64_t)(row_begin + r + L_global * c) + base;
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    out[i] = lo + span * ((double)(z >> 11) * (1.0 / 9007199254740992.0));
  }
}

// ---- small elementwise / reduction kernels --------------------------------------------------------------------
struct SqnormArgs {
  const double *x[16];
  int64_t n[16];
};

// one block per array: deterministic sum of squares
__global__ void __launch_bounds__(1024) sqnorms_kernel(SqnormArgs a, double *out) {
  __shared__ double red[32];
  const double *x = a.x[blockIdx.x];
  int64_t n = a.n[blockIdx.x];
  double s = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    double v = x[i];
    s += v * v;
  }
  s = ppx_block_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// tensor-sized arrays: many blocks write partial sums (fixed grid, fixed order -> deterministic), sum_stage2 adds them
__global__ void __launch_bounds__(256) sq_stage1(const double *__restrict__ x, int64_t n, double *partial) {
  __shared__ double red[32];
  double s0 = 0, s1 = 0;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i + stride < n; i += 2 * stride) {
    const double a = x[i], b = x[i + stride];
    s0 = fma(a, a, s0);
    s1 = fma(b, b, s1);
  }
  if (i < n) s0 = fma(x[i], x[i], s0);
  const double s = ppx_block_sum(s0 + s1, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// one block per pair of arrays: deterministic inner product
struct DotArgs {
  const double *x[16];
  const double *y[16];
  int64_t n[16];
};
__global__ void __launch_bounds__(1024) dots_kernel(DotArgs a, double *out) {
  __shared__ double red[32];
  const double *x = a.x[blockIdx.x], *y = a.y[blockIdx.x];
  const int64_t n = a.n[blockIdx.x];
  double s = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = fma(x[i], y[i], s);
  s = ppx_block_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) diff_update_kernel(const double *W, double *Wp, double *dW, int64_t n,
                                                           double *out) {
  __shared__ double red[32];
  double s0 = 0, s1 = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    double w = W[i];
    double d = w - Wp[i];
    dW[i] = d;
    Wp[i] = w;
    s0 += d * d;
    s1 += w * w;
  }
  s0 = ppx_block_sum(s0, red);
  s1 = ppx_block_sum(s1, red);
  if (threadIdx.x == 0) {
    out[0] = s0;
    out[1] = s1;
  }
}

__global__ void axpby_kernel(double alpha, const double *x, double beta, double *y, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = alpha * x[i] + (beta == 0.0 ? 0.0 : beta * y[i]);
}

// two-stage deterministic  out = a - b, sum of squares
__global__ void __launch_bounds__(256) diff_sq_stage1(const double *a, const double *b, int64_t n, double *partial) {
  __shared__ double red[32];
  double s = 0;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double d = a[i] - b[i];
    s += d * d;
  }
  s = ppx_block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(1024) sum_stage2(const double *partial, int n, double *out) {
  __shared__ double red[32];
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = ppx_block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void transpose_kernel(const double *A, int64_t m, int64_t n, double *B) {
  __shared__ double tile[32][33];
  int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = bx + threadIdx.x, c = by + j;
    if (r < m && c < n) tile[j][threadIdx.x] = A[r + m * c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t c = by + threadIdx.x, r = bx + j;
    if (r < m && c < n) B[c + n * r] = tile[threadIdx.x][j];
  }
}

int ppx_sum_partials(ppx_ctx *ctx, const double *partial, int n, double *out) {
  sum_stage2<<<1, 1024, 0, ctx->stream>>>(partial, n, out);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

extern "C" {

int ppx_fill_uniform(ppx_ctx *ctx, double *out, int64_t n, uint64_t seed, uint64_t tensor_id, int64_t start,
                     double lo, double hi) {
  PPX_REQUIRE(ctx, out && n >= 0, "out != NULL, n >= 0");
  if (n == 0) return PPX_OK;
  uint64_t base = seed * 0x9E3779B97F4A7C15ULL + tensor_id * 0xD1B54A32D192ED03ULL;
  int blocks = (int)((n + 255) / 256);
  if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  fill_uniform_kernel<<<blocks, 256, 0, ctx->stream>>>(out, n, base, start, lo, hi - lo);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_fill_uniform_rows(ppx_ctx *ctx, double *out, int64_t L_local, int64_t L_global, int64_t row_begin,
                          int64_t n_cols, uint64_t seed, uint64_t tensor_id, double lo, double hi) {
  PPX_REQUIRE(ctx, out && L_local >= 0 && n_cols >= 0 && row_begin >= 0 && row_begin + L_local <= L_global,
              "out != NULL, 0 <= row_begin, row_begin + L_local <= L_global");
  const int64_t n = L_local * n_cols;
  if (n == 0) return PPX_OK;
  uint64_t base = seed * 0x9E3779B97F4A7C15ULL + tensor_id * 0xD1B54A32D192ED03ULL;
  int blocks = (int)((n + 255) / 256);
  if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  fill_uniform_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(out, L_local, L_global, row_begin, n_cols, base, lo, hi - lo);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_fill_laplacian(ppx_ctx *ctx, double *out, int d, int64_t s) {
  PPX_REQUIRE(ctx, out && d >= 1 && d <= 8 && s >= 1, "out != NULL, 1 <= d <= 8, s >= 1");
  int64_t n = 1;
  for (int m = 0; m < 2 * d; m++) {
    if (n > ((int64_t)1 << 62) / s) return ppx_set_err(ctx, PPX_EINVAL, "fill_laplacian: s^(2d) overflows");
    n *= s;
  }
  int blocks = (int)((n + 255) / 256 > (int64_t)ctx->sm_count * 16 ? (int64_t)ctx->sm_count * 16 : (n + 255) / 256);
  fill_laplacian_kernel<<<blocks, 256, 0, ctx->stream>>>(out, n, d, s);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_sqnorms(ppx_ctx *ctx, const double *const *X, const int64_t *n, int count, double *out_dev) {
  PPX_REQUIRE(ctx, count >= 1 && count <= 16, "1 <= count <= 16");
  SqnormArgs a;
  for (int i = 0; i < count; i++) {
    a.x[i] = X[i];
    a.n[i] = n[i];
  }
  // factor-sized arrays: one block each, one launch.  An array of tensor size would run through a single SM (64.8 GB at
  // BASELINE configs[1]: about a second), so those go through the grid-wide two-stage sum.
  bool big = false;
  for (int i = 0; i < count; i++) big = big || n[i] > ((int64_t)1 << 22);
  if (!big) {
    sqnorms_kernel<<<count, 1024, 0, ctx->stream>>>(a, out_dev);
    PPX_CHECK_LAUNCH(ctx);
    return PPX_OK;
  }
  ppx_ws_reset(ctx);
  const int blocks = ctx->sm_count * 8;
  double *partial = (double *)ppx_ws_alloc(ctx, sizeof(double) * blocks);
  if (!partial) return ppx_set_err(ctx, PPX_ENOMEM, "workspace too small");
  for (int i = 0; i < count; i++) {
    sq_stage1<<<blocks, 256, 0, ctx->stream>>>(X[i], n[i], partial);
    PPX_CHECK_LAUNCH(ctx);
    int rc = ppx_sum_partials(ctx, partial, blocks, out_dev + i);
    if (rc) return rc;
  }
  return PPX_OK;
}

int ppx_dots(ppx_ctx *ctx, const double *const *X, const double *const *Y, const int64_t *n, int count,
             double *out_dev) {
  PPX_REQUIRE(ctx, X && Y && n && out_dev && count >= 1 && count <= 16, "1 <= count <= 16");
  DotArgs a;
  for (int i = 0; i < count; i++) {
    a.x[i] = X[i];
    a.y[i] = Y[i];
    a.n[i] = n[i];
  }
  dots_kernel<<<count, 1024, 0, ctx->stream>>>(a, out_dev);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_diff_update(ppx_ctx *ctx, const double *W, double *W_prev, double *dW, int64_t n, double *sq_out_dev) {
  diff_update_kernel<<<1, 1024, 0, ctx->stream>>>(W, W_prev, dW, n, sq_out_dev);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_axpby(ppx_ctx *ctx, double alpha, const double *x, double beta, double *y, int64_t n) {
  if (n == 0) return PPX_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  axpby_kernel<<<blocks, 256, 0, ctx->stream>>>(alpha, x, beta, y, n);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_diff_sqnorm(ppx_ctx *ctx, const double *a, const double *b, int64_t n, double *sq_out_dev) {
  ppx_ws_reset(ctx);
  int blocks = ctx->sm_count * 8;
  double *partial = (double *)ppx_ws_alloc(ctx, sizeof(double) * blocks);
  if (!partial) return ppx_set_err(ctx, PPX_ENOMEM, "workspace too small");
  diff_sq_stage1<<<blocks, 256, 0, ctx->stream>>>(a, b, n, partial);
  PPX_CHECK_LAUNCH(ctx);
  return ppx_sum_partials(ctx, partial, blocks, sq_out_dev);
}

int ppx_transpose(ppx_ctx *ctx, const double *A, int64_t m, int64_t n, double *B) {
  dim3 grid(ppx_cdiv(m, 32), ppx_cdiv(n, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, ctx->stream>>>(A, m, n, B);
  PPX_CHECK_LAUNCH(ctx);
  return PPX_OK;
}

int ppx_shard_range(int64_t s, int nranks, int rank, int64_t *begin, int64_t *end) {
  if (nranks < 1 || rank < 0 || rank >= nranks || s < 0 || !begin || !end) return PPX_EINVAL;
  int64_t q = s / nranks, r = s % nranks;
  *begin = rank * q + (rank < r ? rank : r);
  *end = *begin + q + (rank < r ? 1 : 0);
  return PPX_OK;
}

}  // extern "C"
