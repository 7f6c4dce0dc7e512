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
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

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

}  // namespace

void ppx_comm_destroy_internal(ppx_ctx *ctx) {
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
