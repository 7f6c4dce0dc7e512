// K12 -- collectives.  The reference's communication is implicit in every CTF expression (MPI all-to-all
// redistribution, SUMMA broadcasts, allreduce inside norm2).  With the tensor sharded along one mode the only
// exchange left is the sum over ranks of each s x R partial MTTKRP, the R x R partial Grams and a few scalars:
// ppx_allreduce_packed issues them as ONE NCCL group on the context stream.
//
// NCCL is loaded lazily with dlopen so that the single-GPU path has no link-time dependency on it (inside a
// Python process that already imported torch this resolves to torch's bundled libnccl.so.2).
#include <arpa/inet.h>
#include <dlfcn.h>
#include <errno.h>
#include <netinet/in.h>
#include <poll.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>
#include "ppx_internal.h"

namespace {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_p;
typedef int ncclResult_e;  // 0 == ncclSuccess
enum { NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct NcclApi {
  void *lib = nullptr;
  ncclResult_e (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
  ncclResult_e (*CommInitRank)(ncclComm_p *, int, ncclUniqueId_t, int) = nullptr;
  ncclResult_e (*CommDestroy)(ncclComm_p) = nullptr;
  ncclResult_e (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  ncclResult_e (*Send)(const void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  ncclResult_e (*Recv)(void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  ncclResult_e (*GroupStart)() = nullptr;
  ncclResult_e (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_e) = nullptr;
  bool tried = false;
};
NcclApi g_nccl;

bool load_nccl() {
  if (g_nccl.tried) return g_nccl.lib != nullptr;
  g_nccl.tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.lib; i++) g_nccl.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.lib) return false;
#define LOAD(sym) *(void **)(&g_nccl.sym) = dlsym(g_nccl.lib, "nccl" #sym)
  LOAD(GetUniqueId);
  LOAD(CommInitRank);
  LOAD(CommDestroy);
  LOAD(AllReduce);
  LOAD(Send);
  LOAD(Recv);
  LOAD(GroupStart);
  LOAD(GroupEnd);
  LOAD(GetErrorString);
#undef LOAD
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
    g_nccl.lib = nullptr;
    return false;
  }
  return true;
}

// ---- one-shot all-reduce over NVLink peer memory ---------------------------------------------------------------------
// What crosses GPUs in a sweep is small (an s x R partial MTTKRP: 120 KB at BASELINE configs[1]; an R x R Gram; a few
// scalars) and sits on the critical path of every mode update, so its cost is latency: NCCL needs 29 us for 120 KB on
// 8 GPUs.  Every rank owns a staging buffer that all peers have mapped (cudaIpc); the kernel
//   1. copies its slice of the input to its own staging buffer, fences, and writes the call's epoch into a flag word in
//      EVERY peer's memory (one remote store per peer and CTA);
//   2. spins on its own flag words until every peer's epoch has arrived (the peers' data are then visible);
//   3. reads the slice from every peer's staging buffer through NVLink, adds the P values in rank order -- every rank
//      forms bit-identical sums, replicated quantities stay replicated -- and writes the result in place.
// CTA c talks only to CTA c of the peers (flags per CTA), so there is no grid-wide barrier; staging and flags are double
// buffered by the parity of the epoch, which needs no closing barrier: a rank can only start call e+2 (same parity as e)
// after completing e+1, and that took every peer's flag e+1, which a peer writes after it has finished reading in
// call e.  The epoch lives in device memory and is advanced by the kernel itself, so the kernel can sit in a captured
// CUDA graph.  Several buffers go through one launch (ppx_allreduce_packed).  A spin that sees no flag for 300 s gives
// up, reports through a device-side error word and lets the stream drain instead of hanging the GPU.
constexpr int P2P_MAXR = 8;       // ranks (one NVSwitch domain)
constexpr int P2P_MAXB = 8;       // buffers per launch
constexpr int P2P_MAXCTA = 16;    // CTAs per launch
constexpr int P2P_THREADS = 512;
constexpr size_t P2P_CAP = 96 * 1024;  // doubles per call and parity (768 KB)

struct P2PState {
  int nranks = 0, rank = 0;
  char *base[P2P_MAXR] = {};  // mapped regions, [rank] is the local allocation
  bool opened[P2P_MAXR] = {};
};
// region layout (bytes)
constexpr size_t P2P_OFF_FLAGS = 0;                                               // [2][MAXCTA][MAXR] u64
constexpr size_t P2P_OFF_EPOCH = P2P_OFF_FLAGS + 2 * P2P_MAXCTA * P2P_MAXR * 8;   // [MAXCTA] u64
constexpr size_t P2P_OFF_ERR = P2P_OFF_EPOCH + P2P_MAXCTA * 8;                    // u64
constexpr size_t P2P_OFF_DONE = P2P_OFF_ERR + 64;                                 // [MAXR] u64: teardown handshake
constexpr size_t P2P_OFF_STAGE = 4096;                                            // [2][CAP] double
constexpr size_t P2P_BYTES = P2P_OFF_STAGE + 2 * P2P_CAP * 8;
static_assert(P2P_OFF_DONE + 8 * P2P_MAXR <= P2P_OFF_STAGE, "header fits");

struct P2PArgs {
  char *base[P2P_MAXR];
  double *buf[P2P_MAXB];
  long long off[P2P_MAXB + 1];  // packed offsets of the buffers, off[nbuf] = total
  int nranks, rank, nbuf;
};

__device__ __forceinline__ unsigned long long p2p_ld_acquire(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void p2p_st_release(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double p2p_ld_data(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(P2P_THREADS) allreduce_oneshot_kernel(P2PArgs a) {
  const int c = blockIdx.x, tid = threadIdx.x;
  char *mine = a.base[a.rank];
  unsigned long long *epoch = reinterpret_cast<unsigned long long *>(mine + P2P_OFF_EPOCH) + c;
  const unsigned long long e = *epoch + 1;  // only this CTA writes this word (at the end)
  const int par = (int)(e & 1);
  const long long total = a.off[a.nbuf];
  long long chunk = (total + gridDim.x - 1) / gridDim.x;
  const long long lo = c * chunk, hi = lo + chunk < total ? lo + chunk : total;
  double *stage = reinterpret_cast<double *>(mine + P2P_OFF_STAGE) + (size_t)par * P2P_CAP;
  // 1. own slice -> own staging (four elements per thread and round, loads first)
  constexpr int U = 4;
  for (long long i0 = lo + tid; i0 < hi; i0 += U * P2P_THREADS) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long i = i0 + u * P2P_THREADS;
      if (i < hi) {
        int b = 0;
        while (i >= a.off[b + 1]) b++;
        v[u] = a.buf[b][i - a.off[b]];
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long i = i0 + u * P2P_THREADS;
      if (i < hi) stage[i] = v[u];
    }
  }
  // the barrier orders every thread's staging stores before the release stores below (release is cumulative)
  __syncthreads();
  const size_t fslot = ((size_t)par * P2P_MAXCTA + c) * P2P_MAXR;
  if (tid < a.nranks && tid != a.rank) {
    p2p_st_release(reinterpret_cast<unsigned long long *>(a.base[tid] + P2P_OFF_FLAGS) + fslot + a.rank, e);
    // 2. wait for that peer's epoch
    const unsigned long long *f = reinterpret_cast<const unsigned long long *>(mine + P2P_OFF_FLAGS) + fslot + tid;
    unsigned long long t0 = 0;
    int spins = 0;
    while (p2p_ld_acquire(f) < e) {
      if (++spins < 4096) continue;  // (the clock is read only once the wait is long)
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t0 == 0) t0 = t1;
      if (t1 - t0 > 300000000000ull) {  // a peer never arrived: do not hang the GPU
        *reinterpret_cast<unsigned long long *>(mine + P2P_OFF_ERR) = e;
        printf("ppx: one-shot all-reduce: rank %d saw no flag from rank %d for 300 s (call %llu, CTA %d); results are wrong\n",
               a.rank, tid, e, c);
        break;
      }
    }
  }
  __syncthreads();
  // 3. sum in rank order; all P x U loads of a round are in flight before the first addition
  for (long long i0 = lo + tid; i0 < hi; i0 += U * P2P_THREADS) {
    double v[P2P_MAXR][U];
#pragma unroll
    for (int r = 0; r < P2P_MAXR; r++) {
      if (r < a.nranks) {
        const double *src = reinterpret_cast<const double *>(a.base[r] + P2P_OFF_STAGE) + (size_t)par * P2P_CAP;
#pragma unroll
        for (int u = 0; u < U; u++) {
          const long long i = i0 + u * P2P_THREADS;
          v[r][u] = i < hi ? (r == a.rank ? stage[i] : p2p_ld_data(src + i)) : 0.0;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long i = i0 + u * P2P_THREADS;
      if (i < hi) {
        double sum = v[0][u];
#pragma unroll
        for (int r = 1; r < P2P_MAXR; r++)
          if (r < a.nranks) sum += v[r][u];
        int b = 0;
        while (i >= a.off[b + 1]) b++;
        a.buf[b][i - a.off[b]] = sum;
      }
    }
  }
  __syncthreads();
  if (tid == 0) *epoch = e;
}

void p2p_teardown(ppx_ctx *ctx, bool live = false) {
  P2PState *st = (P2PState *)ctx->p2p;
  if (!st) return;
  if (live && st->base[st->rank]) {
    // A peer may still be reading this rank's staging buffer in its last call: nobody frees before everybody is done.
    // Host-side handshake through the mapped regions, bounded (3 s): a peer that died or never tears down must not
    // hang this rank, which an NCCL barrier here would.
    cudaStreamSynchronize(ctx->main_stream);
    const unsigned long long one = 1;
    for (int r = 0; r < st->nranks; r++)
      if (r != st->rank && st->opened[r])
        cudaMemcpy(st->base[r] + P2P_OFF_DONE + 8 * st->rank, &one, 8, cudaMemcpyDefault);
    for (int tries = 0; tries < 15000; tries++) {
      unsigned long long seen[P2P_MAXR] = {};
      if (cudaMemcpy(seen, st->base[st->rank] + P2P_OFF_DONE, sizeof(seen), cudaMemcpyDefault) != cudaSuccess) break;
      bool all = true;
      for (int r = 0; r < st->nranks; r++)
        if (r != st->rank && st->opened[r] && !seen[r]) all = false;
      if (all) break;
      usleep(200);
    }
    cudaGetLastError();
  }
  for (int r = 0; r < st->nranks; r++) {
    if (r == st->rank) {
      if (st->base[r]) cudaFree(st->base[r]);
    } else if (st->opened[r]) {
      cudaIpcCloseMemHandle(st->base[r]);
    }
  }
  delete st;
  ctx->p2p = nullptr;
}

// Collective over the communicator: map every peer's staging region.  Any failure on any rank (peers on another node,
// no peer access, IPC not permitted in this container, PPX_NO_P2P set) leaves every rank on NCCL.
void p2p_setup(ppx_ctx *ctx) {
  const int P = ctx->nranks;
  if (P < 2 || P > P2P_MAXR) return;
  P2PState *st = new P2PState();
  st->nranks = P;
  st->rank = ctx->rank;
  unsigned long long fail = getenv("PPX_NO_P2P") != nullptr;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "64-byte IPC handles");
  if (!fail) {
    if (cudaMalloc((void **)&st->base[ctx->rank], P2P_BYTES) != cudaSuccess ||
        cudaMemset(st->base[ctx->rank], 0, P2P_BYTES) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, st->base[ctx->rank]) != cudaSuccess) {
      cudaGetLastError();
      fail = 1;
    }
  }
  // all-gather of the handles (+ a failure count) as an NCCL sum of zero-padded u64 words
  const int words = P * 8 + 1;
  unsigned long long *dev = nullptr, host[P2P_MAXR * 8 + 1];
  memset(host, 0, sizeof(host));
  memcpy(host + 8 * ctx->rank, &mine, 64);
  host[P * 8] = fail;
  bool ok = cudaMalloc((void **)&dev, sizeof(unsigned long long) * words) == cudaSuccess &&
            cudaMemcpyAsync(dev, host, sizeof(unsigned long long) * words, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
            g_nccl.AllReduce(dev, dev, (size_t)words, NCCL_UINT64, NCCL_SUM, (ncclComm_p)ctx->comm, ctx->stream) == 0 &&
            cudaMemcpyAsync(host, dev, sizeof(unsigned long long) * words, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
            cudaStreamSynchronize(ctx->stream) == cudaSuccess;
  if (!ok) {  // cannot even agree: NCCL only (the other ranks fail the same collective or see our absence as an error)
    cudaGetLastError();
    if (dev) cudaFree(dev);
    ctx->p2p = st;
    p2p_teardown(ctx);
    return;
  }
  unsigned long long bad = host[P * 8];
  if (!bad) {
    for (int r = 0; r < P; r++) {
      if (r == ctx->rank) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, host + 8 * r, 64);
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        bad = 1;
        break;
      }
      st->base[r] = (char *)ptr;
      st->opened[r] = true;
    }
  }
  // second agreement: did every rank map every peer?
  host[0] = bad;
  ok = cudaMemcpyAsync(dev, host, sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
       g_nccl.AllReduce(dev, dev, 1, NCCL_UINT64, NCCL_SUM, (ncclComm_p)ctx->comm, ctx->stream) == 0 &&
       cudaMemcpyAsync(host, dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
       cudaStreamSynchronize(ctx->stream) == cudaSuccess;
  cudaFree(dev);
  ctx->p2p = st;
  if (!ok || host[0] != 0) {
    cudaGetLastError();
    p2p_teardown(ctx);
    if (getenv("PPX_COMM_VERBOSE") && ctx->rank == 0) fprintf(stderr, "ppx: peer-memory all-reduce unavailable, using NCCL\n");
    return;
  }
  if (getenv("PPX_COMM_VERBOSE") && ctx->rank == 0)
    fprintf(stderr, "ppx: one-shot all-reduce over peer memory enabled (%d ranks, up to %zu doubles per call)\n", P, P2P_CAP);
}

}  // namespace

void ppx_comm_destroy_internal(ppx_ctx *ctx) {
  p2p_teardown(ctx, true);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_p)ctx->comm);
  ctx->comm = nullptr;
}

extern "C" {

int ppx_comm_unique_id(void *id128) {
  if (!id128) return PPX_EINVAL;
  if (!load_nccl()) return ppx_set_err(nullptr, PPX_ENCCL, "libnccl.so.2 not found: %s", dlerror());
  ncclUniqueId_t id;
  ncclResult_e r = g_nccl.GetUniqueId(&id);
  if (r) return ppx_set_err(nullptr, PPX_ENCCL, "ncclGetUniqueId failed (%d)", r);
  memcpy(id128, &id, 128);
  return PPX_OK;
}

int ppx_comm_init(ppx_ctx *ctx, const void *id128, int nranks, int rank) {
  PPX_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks, "0 <= rank < nranks");
  ctx->nranks = nranks;
  ctx->rank = rank;
  if (nranks == 1) return PPX_OK;
  PPX_REQUIRE(ctx, id128 != nullptr, "id128 != NULL");
  if (!load_nccl()) return ppx_set_err(ctx, PPX_ENCCL, "libnccl.so.2 not found");
  ncclUniqueId_t id;
  memcpy(&id, id128, 128);
  PPX_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclComm_p comm = nullptr;
  ncclResult_e r = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (r)
    return ppx_set_err(ctx, PPX_ENCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  ctx->comm = comm;
  p2p_setup(ctx);
  return PPX_OK;
}

int ppx_comm_bootstrap(ppx_ctx *ctx, int nranks, int rank, const char *addr, int port, int timeout_s) {
  PPX_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks, "0 <= rank < nranks");
  if (nranks == 1) return ppx_comm_init(ctx, nullptr, 1, 0);
  PPX_REQUIRE(ctx, addr && port > 0 && port < 65536, "addr != NULL, 0 < port < 65536");
  if (timeout_s <= 0) timeout_s = 120;
  unsigned char id[128];
  struct sockaddr_in sa;
  memset(&sa, 0, sizeof(sa));
  sa.sin_family = AF_INET;
  sa.sin_port = htons((uint16_t)port);
  auto send_all = [](int fd, const void *buf, size_t n) {
    const char *p = (const char *)buf;
    while (n) {
      ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
      if (k <= 0) return false;
      p += k;
      n -= (size_t)k;
    }
    return true;
  };
  auto recv_all = [](int fd, void *buf, size_t n) {
    char *p = (char *)buf;
    while (n) {
      ssize_t k = recv(fd, p, n, 0);
      if (k <= 0) return false;
      p += k;
      n -= (size_t)k;
    }
    return true;
  };
  if (rank == 0) {
    int rc = ppx_comm_unique_id(id);
    if (rc) return ppx_set_err(ctx, PPX_ENCCL, "ncclGetUniqueId failed");
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    if (ls < 0) return ppx_set_err(ctx, PPX_ENCCL, "bootstrap: socket(): %s", strerror(errno));
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sa.sin_addr.s_addr = htonl(INADDR_ANY);
    if (bind(ls, (struct sockaddr *)&sa, sizeof(sa)) != 0 || listen(ls, nranks) != 0) {
      const int e = errno;
      close(ls);
      return ppx_set_err(ctx, PPX_ENCCL, "bootstrap: cannot listen on port %d: %s (set PPX_BOOT_PORT)", port, strerror(e));
    }
    for (int got = 0; got < nranks - 1; got++) {
      struct pollfd pf = {ls, POLLIN, 0};
      if (poll(&pf, 1, timeout_s * 1000) <= 0) {
        close(ls);
        return ppx_set_err(ctx, PPX_ENCCL, "bootstrap: only %d of %d peers connected within %d s", got, nranks - 1, timeout_s);
      }
      int fd = accept(ls, nullptr, nullptr);
      if (fd < 0) {
        got--;
        continue;
      }
      int peer = -1;
      const bool ok = recv_all(fd, &peer, sizeof(peer)) && peer > 0 && peer < nranks && send_all(fd, id, sizeof(id));
      close(fd);
      if (!ok) {
        close(ls);
        return ppx_set_err(ctx, PPX_ENCCL, "bootstrap: exchange with a peer failed");
      }
    }
    close(ls);
  } else {
    if (inet_pton(AF_INET, strcmp(addr, "localhost") == 0 ? "127.0.0.1" : addr, &sa.sin_addr) != 1)
      return ppx_set_err(ctx, PPX_ENCCL, "bootstrap: '%s' is not an IPv4 address", addr);
    bool ok = false;
    for (int attempt = 0; attempt < timeout_s * 20 && !ok; attempt++) {
      int fd = socket(AF_INET, SOCK_STREAM, 0);
      if (fd < 0) break;
      if (connect(fd, (struct sockaddr *)&sa, sizeof(sa)) == 0)
        ok = send_all(fd, &rank, sizeof(rank)) && recv_all(fd, id, sizeof(id));
      close(fd);
      if (!ok) usleep(50000);
    }
    if (!ok) return ppx_set_err(ctx, PPX_ENCCL, "bootstrap: rank %d could not reach rank 0 at %s:%d", rank, addr, port);
  }
  return ppx_comm_init(ctx, id, nranks, rank);
}

int ppx_comm_p2p(ppx_ctx *ctx) { return ctx->p2p != nullptr; }
int ppx_comm_size(ppx_ctx *ctx) { return ctx->nranks; }
int ppx_comm_rank(ppx_ctx *ctx) { return ctx->rank; }

int ppx_allreduce_packed(ppx_ctx *ctx, double *const *bufs, const int64_t *sizes, int n) {
  if (ctx->nranks == 1 || n == 0) return PPX_OK;
  // timing aid: PPX_SKIP_ALLREDUCE=1 turns the exchange into a no-op (results are then wrong on purpose) so that the
  // cost of the collectives, including the waiting they impose on skewed ranks, can be read off as a difference
  static const bool skip = getenv("PPX_SKIP_ALLREDUCE") != nullptr;
  if (skip) return PPX_OK;
  PPX_REQUIRE(ctx, ctx->comm != nullptr, "communicator initialised (ppx_comm_init)");
  PPX_REQUIRE(ctx, bufs && sizes && n > 0, "bufs, sizes non-null");
  if (ctx->p2p && n <= P2P_MAXB) {
    P2PState *st = (P2PState *)ctx->p2p;
    P2PArgs a;
    a.nranks = st->nranks;
    a.rank = st->rank;
    a.nbuf = 0;
    a.off[0] = 0;
    for (int i = 0; i < n; i++)
      if (sizes[i] > 0) {
        a.buf[a.nbuf] = bufs[i];
        a.off[a.nbuf + 1] = a.off[a.nbuf] + sizes[i];
        a.nbuf++;
      }
    const long long total = a.off[a.nbuf];
    if (total == 0) return PPX_OK;
    if ((size_t)total <= P2P_CAP) {
      for (int r = 0; r < P2P_MAXR; r++) a.base[r] = r < st->nranks ? st->base[r] : nullptr;
      int nblk = (int)((total + 1023) / 1024);
      nblk = nblk < 1 ? 1 : nblk > P2P_MAXCTA ? P2P_MAXCTA : nblk;
      // the number of CTAs is a function of the sizes alone: every rank launches the same grid
      allreduce_oneshot_kernel<<<nblk, P2P_THREADS, 0, ctx->stream>>>(a);
      PPX_CHECK_LAUNCH(ctx);
      return PPX_OK;
    }
  }
  ncclResult_e r = g_nccl.GroupStart();
  for (int i = 0; i < n && !r; i++)
    if (sizes[i] > 0)
      r = g_nccl.AllReduce(bufs[i], bufs[i], (size_t)sizes[i], NCCL_FLOAT64, NCCL_SUM, (ncclComm_p)ctx->comm, ctx->stream);
  ncclResult_e r2 = g_nccl.GroupEnd();
  if (r || r2)
    return ppx_set_err(ctx, PPX_ENCCL, "ncclAllReduce failed: %s",
                       g_nccl.GetErrorString ? g_nccl.GetErrorString(r ? r : r2) : "?");
  ctx->launches += n;
  return PPX_OK;
}

int ppx_alltoallv(ppx_ctx *ctx, const double *const *sendbufs, const int64_t *sendcounts, double *const *recvbufs,
                  const int64_t *recvcounts) {
  if (ctx->nranks == 1) return PPX_OK;
  PPX_REQUIRE(ctx, ctx->comm != nullptr, "communicator initialised (ppx_comm_init)");
  PPX_REQUIRE(ctx, sendbufs && sendcounts && recvbufs && recvcounts, "non-null arrays");
  if (!g_nccl.Send || !g_nccl.Recv) return ppx_set_err(ctx, PPX_ENCCL, "this NCCL has no ncclSend/ncclRecv");
  ncclResult_e r = g_nccl.GroupStart();
  for (int k = 0; k < ctx->nranks && !r; k++) {
    if (k == ctx->rank) continue;
    if (sendcounts[k] > 0)
      r = g_nccl.Send(sendbufs[k], (size_t)sendcounts[k], NCCL_FLOAT64, k, (ncclComm_p)ctx->comm, ctx->stream);
    if (!r && recvcounts[k] > 0)
      r = g_nccl.Recv(recvbufs[k], (size_t)recvcounts[k], NCCL_FLOAT64, k, (ncclComm_p)ctx->comm, ctx->stream);
  }
  ncclResult_e r2 = g_nccl.GroupEnd();
  if (r || r2)
    return ppx_set_err(ctx, PPX_ENCCL, "ncclSend/ncclRecv failed: %s",
                       g_nccl.GetErrorString ? g_nccl.GetErrorString(r ? r : r2) : "?");
  // the piece a rank keeps for itself is a plain copy
  const int me = ctx->rank;
  if (sendcounts[me] > 0 && recvbufs[me] && recvbufs[me] != sendbufs[me]) {
    PPX_REQUIRE(ctx, recvcounts[me] == sendcounts[me], "self counts agree");
    PPX_CUDA(ctx, cudaMemcpyAsync(recvbufs[me], sendbufs[me], sizeof(double) * (size_t)sendcounts[me],
                                  cudaMemcpyDeviceToDevice, ctx->stream));
  }
  ctx->launches += ctx->nranks;
  return PPX_OK;
}

int ppx_ttm_first_mttv(ppx_ctx *ctx, const double *V, const int64_t *lens, int N, int x1, const double *W1,
                       int64_t ldw1, int x2, const double *W2, int64_t ldw2, int R, double *out) {
  // adjacent modes: one GEMM against the Khatri-Rao rows (ppx_ttm_multi); otherwise the two contractions back to
  // back through the context workspace
  PPX_REQUIRE(ctx, V && lens && W1 && W2 && out, "non-null pointers");
  PPX_REQUIRE(ctx, N >= 2 && N <= 16 && x1 >= 0 && x1 < N && x2 >= 0 && x2 < N && x1 != x2, "x1 != x2 in [0,N)");
  if (x2 == x1 + 1 || x1 == x2 + 1) {
    const double *W[2] = {x1 < x2 ? W1 : W2, x1 < x2 ? W2 : W1};
    const int64_t ld[2] = {x1 < x2 ? ldw1 : ldw2, x1 < x2 ? ldw2 : ldw1};
    return ppx_ttm_multi(ctx, V, lens, N, x1 < x2 ? x1 : x2, 2, W, ld, R, out);
  }
  int64_t P = 1;
  for (int i = 0; i < N; i++) P *= lens[i];
  const int64_t n1 = P / lens[x1] * R;
  double *tmp = nullptr;
  cudaError_t e = cudaMalloc((void **)&tmp, sizeof(double) * (size_t)n1);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return ppx_set_err(ctx, PPX_ENOMEM, "ttm_first_mttv: cannot allocate the %lld-byte level-1 tensor", (long long)(8 * n1));
  }
  int rc = ppx_ttm_first(ctx, V, lens, N, x1, W1, ldw1, R, tmp);
  if (!rc) {
    int64_t lens2[16];
    int k = 0;
    for (int i = 0; i < N; i++)
      if (i != x1) lens2[k++] = lens[i];
    rc = ppx_mttv(ctx, tmp, lens2, k, x2 > x1 ? x2 - 1 : x2, W2, ldw2, R, out);
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  return rc;
}

}  // extern "C"
