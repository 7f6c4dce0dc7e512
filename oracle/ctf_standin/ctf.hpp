/* ctf.hpp -- TEST INFRASTRUCTURE.  A single-process, dense, FP64 stand-in for the part of the Cyclops Tensor
 * Framework API (and of MPI) that the reference sources use, so that the reference's OWN files
 * (/root/reference/{common,als_CP,als_Tucker,test_ALS,pp_bench,run}.cxx and src/**) compile unmodified, from where
 * they lie, into oracle/_ref/ (recipe: oracle/Makefile, target `ref`).  Nothing under pairwise-perturbation_b200/
 * includes or links this file; only tests/ and bench.py's cpu legs execute what is built from it.
 *
 * What it is: Tensor<>/Matrix<>/Vector<> as dense column-major (first index fastest = CTF's global order) arrays with
 * deep-copy semantics and zero initialisation; the Einstein-string expression algebra (`C["ij"] = 2.*A["ik"]*B["kj"] -
 * D["ij"]`, `+=`, `-=`, repeated indices = diagonals, indices absent from the output are summed) evaluated with plain
 * loops; svd (one-sided Jacobi), qr, cholesky, solve_tri; norm2, write, read_local, read_dense_from_file, fill_random;
 * Transform<>; Timer/Timer_epoch; World; the handful of MPI symbols the mains call.
 * Optional: with CTF_STANDIN_BLAS=<path of an OpenBLAS shared library> (bench.py's CPU arm points it at the one NumPy
 * ships) a product of two tensors whose index pattern folds into a matrix product -- the tensor-times-matrix and
 * Hadamard-batched contractions of the dimension tree, the Grams, the solves -- is handed to DGEMM slice by slice, the
 * way CTF itself maps a contraction to local GEMMs; everything else, and everything when the variable is unset (all
 * parity tests), runs through the plain loops.  tests/test_reference_pin.py checks the two paths against each other.
 * What it is not: CTF.  Summation order, the SVD algorithm and fill_random's stream differ from the real library, so
 * agreement with a CTF build holds to rounding (and up to the sign of singular vectors), not bit for bit.
 * fill_random(lo, hi) draws u(seed, id, linear_index) from the repo's counter-based generator (oracle/pp_oracle.py
 * u01, oracle/naive.c naive_u01, ppx_fill_uniform); call k of the process uses the k-th "seed:id" pair of the
 * environment variable CTF_STANDIN_FILLS (comma separated), or (CTF_STANDIN_SEED or 1, k) when the list is shorter.
 */
#ifndef CTF_STANDIN_HPP
#define CTF_STANDIN_HPP

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

/* ---------------------------------------------------------------- MPI symbols the mains use (one process) */
typedef int MPI_Comm;
typedef int MPI_Info;
typedef FILE *MPI_File;
#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_INFO_NULL 0
#define MPI_MODE_RDWR 2
#define MPI_MODE_CREATE 1
#define MPI_MODE_RDONLY 4
static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline double MPI_Wtime() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline int MPI_File_open(MPI_Comm, const char *name, int, MPI_Info, MPI_File *fh) {
  *fh = fopen(name, "rb");  /* the reference only reads */
  return *fh ? 0 : 1;
}
static inline int MPI_File_close(MPI_File *fh) {
  if (fh && *fh) fclose(*fh);
  if (fh) *fh = nullptr;
  return 0;
}

namespace CTF {
using namespace std; /* the reference relies on ctf.hpp bringing std names into scope */

enum { NS = 0, SY = 1, AS = 2, SH = 3 };

class World {
public:
  int rank = 0, np = 1;
  MPI_Comm comm = MPI_COMM_WORLD;
  World() {}
  World(int, char **) {}
  World(MPI_Comm c) : comm(c) {}
  World(const char *) {}
};
inline World &get_universe() {
  static World w;
  return w;
}

namespace standin {
inline double u01(uint64_t seed, uint64_t tensor_id, uint64_t idx) {
  uint64_t z = idx + seed * 0x9E3779B97F4A7C15ULL + tensor_id * 0xD1B54A32D192ED03ULL;
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
struct FillSchedule {
  vector<pair<uint64_t, uint64_t>> list;
  uint64_t seed = 1, calls = 0;
  FillSchedule() {
    if (const char *s = getenv("CTF_STANDIN_SEED")) seed = strtoull(s, nullptr, 10);
    if (const char *f = getenv("CTF_STANDIN_FILLS")) {
      string str(f), item;
      stringstream ss(str);
      while (getline(ss, item, ',')) {
        size_t c = item.find(':');
        if (c == string::npos) continue;
        list.push_back({strtoull(item.substr(0, c).c_str(), nullptr, 10),
                        strtoull(item.substr(c + 1).c_str(), nullptr, 10)});
      }
    }
  }
  pair<uint64_t, uint64_t> next() {
    uint64_t k = calls++;
    if (k < list.size()) return list[k];
    return {seed, k};
  }
};
inline FillSchedule &fills() {
  static FillSchedule f;
  return f;
}
} // namespace standin

/* ---------------------------------------------------------------- dense storage */
class TensorBase {
public:
  int order = 0;
  int64_t lens[24] = {0};
  int *sym = nullptr;
  bool is_sparse = false;
  World *wrld = nullptr;
  vector<double> data;

  int64_t size() const {
    int64_t n = 1;
    for (int i = 0; i < order; i++) n *= lens[i];
    return n;
  }
  void init(int order_, const int64_t *l, World *w) {
    assert(order_ <= 24);
    order = order_;
    for (int i = 0; i < order; i++) lens[i] = l[i];
    wrld = w ? w : &get_universe();
    data.assign((size_t)size(), 0.0);
  }
  void init(int order_, const int *l, World *w) {
    int64_t l64[24];
    for (int i = 0; i < order_; i++) l64[i] = l[i];
    init(order_, l64, w);
  }
  double norm2() const {
    long double s = 0;
    for (double x : data) s += (long double)x * x;
    return (double)sqrtl(s);
  }
  double norm2(double &nrm) const { return nrm = norm2(); }
  void fill_random(double lo, double hi) {
    pair<uint64_t, uint64_t> si = standin::fills().next();
    for (size_t i = 0; i < data.size(); i++) data[i] = lo + (hi - lo) * standin::u01(si.first, si.second, i);
  }
  void write(int64_t n, const int64_t *inds, const double *vals) {
    for (int64_t i = 0; i < n; i++) {
      assert(inds[i] >= 0 && inds[i] < (int64_t)data.size());
      data[inds[i]] = vals[i];
    }
  }
  /* inds comes from malloc, vals from new[] (the reference frees them that way, common.cxx:878-879) */
  void read_local(int64_t *n, int64_t **inds, double **vals) const {
    *n = (int64_t)data.size();
    *inds = (int64_t *)malloc(sizeof(int64_t) * (data.size() + 1));
    *vals = new double[data.size() + 1];
    for (size_t i = 0; i < data.size(); i++) {
      (*inds)[i] = (int64_t)i;
      (*vals)[i] = data[i];
    }
  }
  void read_dense_from_file(MPI_File &fh) {
    assert(fh && "tensor file not found");
    size_t got = fread(data.data(), sizeof(double), data.size(), fh);
    assert(got == data.size() && "tensor file too short");
    (void)got;
  }
  void write_dense_to_file(MPI_File &) {}
  void print(FILE *fp = stdout) const {
    for (size_t i = 0; i < data.size(); i++) fprintf(fp, "[%zu] %.16e\n", i, data[i]);
  }
  void set_zero() { std::fill(data.begin(), data.end(), 0.0); }
  /* this += tensor living on a sub-world (run.cxx:299,317,378); one process: the tensor itself */
  void add_from_subworld(const TensorBase *t) {
    if (!t) return;
    assert(t->data.size() == data.size());
    for (size_t i = 0; i < data.size(); i++) data[i] += t->data[i];
  }
};

/* ---------------------------------------------------------------- expression algebra */
struct Leaf {
  const TensorBase *t;
  string idx;
};
struct Prod {
  double coef = 1.0;
  vector<Leaf> f;
};
class Term {
public:
  vector<Prod> p;
  Term() {}
  Term(double c) {
    Prod q;
    q.coef = c;
    p.push_back(q);
  }
  /* full contraction to a scalar: `double ip = v1["i"] * v2["i"];` (common.cxx:292-298) */
  inline operator double() const;
};
inline Term operator+(const Term &a, const Term &b) {
  Term r = a;
  r.p.insert(r.p.end(), b.p.begin(), b.p.end());
  return r;
}
inline Term operator-(const Term &a) {
  Term r = a;
  for (auto &q : r.p) q.coef = -q.coef;
  return r;
}
inline Term operator-(const Term &a, const Term &b) { return a + (-b); }
inline Term operator*(const Term &a, const Term &b) {
  Term r;
  for (const auto &x : a.p)
    for (const auto &y : b.p) {
      Prod q;
      q.coef = x.coef * y.coef;
      q.f = x.f;
      q.f.insert(q.f.end(), y.f.begin(), y.f.end());
      r.p.push_back(q);
    }
  return r;
}
inline Term operator*(double a, const Term &b) { return Term(a) * b; }
inline Term operator*(const Term &a, double b) { return a * Term(b); }
inline Term operator+(double a, const Term &b) { return Term(a) + b; }
inline Term operator+(const Term &a, double b) { return a + Term(b); }
inline Term operator-(double a, const Term &b) { return Term(a) - b; }
inline Term operator-(const Term &a, double b) { return a - Term(b); }

namespace standin {
inline void strides_of(const TensorBase *t, vector<int64_t> &st) {
  st.resize(t->order);
  int64_t s = 1;
  for (int i = 0; i < t->order; i++) {
    st[i] = s;
    s *= t->lens[i];
  }
}
/* acc[out positions] += coef * prod_k F_k[...], summed over every index that is not an output index */
inline void accumulate(vector<double> &acc, const TensorBase *out, const string &oidx, const Prod &q) {
  assert((int)oidx.size() == out->order);
  string chars;
  vector<int64_t> ext;
  auto add_char = [&](char c, int64_t e) {
    size_t pos = chars.find(c);
    if (pos == string::npos) {
      chars.push_back(c);
      ext.push_back(e);
    } else {
      assert(ext[pos] == e && "index extent mismatch");
    }
  };
  for (int i = 0; i < out->order; i++) add_char(oidx[i], out->lens[i]);
  const size_t n_out_chars = chars.size();
  for (const auto &l : q.f) {
    assert((int)l.idx.size() == l.t->order && "index string length != tensor order");
    for (int i = 0; i < l.t->order; i++) add_char(l.idx[i], l.t->lens[i]);
  }
  (void)n_out_chars;
  const int nd = (int)chars.size(), nf = (int)q.f.size();
  vector<int64_t> os(nd, 0), st;
  vector<vector<int64_t>> fs(nf, vector<int64_t>(nd, 0));
  strides_of(out, st);
  for (int i = 0; i < out->order; i++) os[chars.find(oidx[i])] += st[i];
  for (int k = 0; k < nf; k++) {
    strides_of(q.f[k].t, st);
    for (int i = 0; i < q.f[k].t->order; i++) fs[k][chars.find(q.f[k].idx[i])] += st[i];
  }
  vector<const double *> F(nf);
  for (int k = 0; k < nf; k++) F[k] = q.f[k].t->data.data();
  for (int d = 0; d < nd; d++)
    if (ext[d] == 0) return;
  if (nd == 0) {
    double v = q.coef;
    for (int k = 0; k < nf; k++) v *= F[k][0];
    acc[0] += v;
    return;
  }
  vector<int64_t> ctr(nd, 0), foff(nf, 0);
  int64_t ooff = 0;
  const int64_t e0 = ext[0], os0 = os[0];
  while (true) {
    if (nf == 2) {
      const double *a = F[0] + foff[0], *b = F[1] + foff[1];
      const int64_t sa = fs[0][0], sb = fs[1][0];
      for (int64_t i = 0; i < e0; i++) acc[ooff + i * os0] += q.coef * a[i * sa] * b[i * sb];
    } else {
      for (int64_t i = 0; i < e0; i++) {
        double v = q.coef;
        for (int k = 0; k < nf; k++) v *= F[k][foff[k] + i * fs[k][0]];
        acc[ooff + i * os0] += v;
      }
    }
    int d = 1;
    for (; d < nd; d++) {
      ctr[d]++;
      ooff += os[d];
      for (int k = 0; k < nf; k++) foff[k] += fs[k][d];
      if (ctr[d] < ext[d]) break;
      ooff -= os[d] * ext[d];
      for (int k = 0; k < nf; k++) foff[k] -= fs[k][d] * ext[d];
      ctr[d] = 0;
    }
    if (d == nd) break;
  }
}
/* ---- optional DGEMM path (CTF_STANDIN_BLAS) ---------------------------------------------------------------------- */
typedef void (*dgemm64_t)(int order, int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const double *a,
                          int64_t lda, const double *b, int64_t ldb, double beta, double *c, int64_t ldc);
typedef void (*dgemm32_t)(int order, int ta, int tb, int m, int n, int k, double alpha, const double *a, int lda,
                          const double *b, int ldb, double beta, double *c, int ldc);
struct Blas {
  dgemm64_t g64 = nullptr;
  dgemm32_t g32 = nullptr;
  void (*set_threads)(int) = nullptr;
  Blas() {
    const char *path = getenv("CTF_STANDIN_BLAS");
    if (!path || !*path) return;
    void *lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) {
      fprintf(stderr, "ctf stand-in: cannot load %s: %s\n", path, dlerror());
      return;
    }
    for (const char *n : {"scipy_cblas_dgemm64_", "cblas_dgemm64_"})
      if (!g64) g64 = (dgemm64_t)dlsym(lib, n);
    if (!g64) g32 = (dgemm32_t)dlsym(lib, "cblas_dgemm");
    for (const char *n : {"scipy_openblas_set_num_threads64_", "openblas_set_num_threads64_", "openblas_set_num_threads"})
      if (!set_threads) set_threads = (void (*)(int))dlsym(lib, n);
  }
  bool ok() const { return g64 || g32; }
  /* column-major C (m x n, ldc) += alpha op(A) op(B) */
  void gemm(bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const double *a, int64_t lda, const double *b,
            int64_t ldb, double *c, int64_t ldc) const {
    if (g64) g64(102, ta ? 112 : 111, tb ? 112 : 111, m, n, k, alpha, a, lda, b, ldb, 1.0, c, ldc);
    else g32(102, ta ? 112 : 111, tb ? 112 : 111, (int)m, (int)n, (int)k, alpha, a, (int)lda, b, (int)ldb, 1.0, c, (int)ldc);
  }
};
inline const Blas &blas() {
  static Blas b;
  return b;
}
struct IdxInfo {
  int64_t ext, sa, sb, so; /* stride in A, B, out; -1 when absent */
};
/* merges, in increasing order of `key`, the indices of `grp` that are contiguous in both tensors of the group:
 * returns the merged extent, fills the two strides, moves what is left to `rest` */
inline int64_t merge_group(vector<IdxInfo> grp, int which1, int which2, int64_t &s1, int64_t &s2, vector<IdxInfo> &rest) {
  auto st = [](const IdxInfo &x, int w) { return w == 0 ? x.sa : (w == 1 ? x.sb : x.so); };
  if (grp.empty()) {
    s1 = s2 = 1;
    return 1;
  }
  sort(grp.begin(), grp.end(), [&](const IdxInfo &x, const IdxInfo &y) { return st(x, which2) < st(y, which2); });
  int64_t ext = grp[0].ext;
  s1 = st(grp[0], which1);
  s2 = st(grp[0], which2);
  size_t used = 1;
  while (used < grp.size() && st(grp[used], which1) == s1 * ext && st(grp[used], which2) == s2 * ext) {
    ext *= grp[used].ext;
    used++;
  }
  for (size_t i = used; i < grp.size(); i++) rest.push_back(grp[i]);
  return ext;
}
/* acc += coef * A * B through DGEMM when the index pattern allows it; false = not taken (nothing written) */
inline bool accumulate_blas(vector<double> &acc, const TensorBase *out, const string &oidx, const Prod &q) {
  const Blas &bl = blas();
  if (!bl.ok() || q.f.size() != 2) return false;
  const Leaf &A = q.f[0], &B = q.f[1];
  auto has_repeat = [](const string &x) {
    for (size_t i = 0; i < x.size(); i++)
      if (x.find(x[i]) != i) return true;
    return false;
  };
  if (has_repeat(A.idx) || has_repeat(B.idx) || has_repeat(oidx)) return false;
  vector<int64_t> stA, stB, stO;
  strides_of(A.t, stA);
  strides_of(B.t, stB);
  strides_of(out, stO);
  string chars = oidx;
  for (char c : A.idx + B.idx)
    if (chars.find(c) == string::npos) chars.push_back(c);
  vector<IdxInfo> Mg, Ng, Kg, loops;
  int64_t work = 1;
  for (char c : chars) {
    IdxInfo x{0, -1, -1, -1};
    size_t pa = A.idx.find(c), pb = B.idx.find(c), po = oidx.find(c);
    if (pa != string::npos) x.sa = stA[pa], x.ext = A.t->lens[pa];
    if (pb != string::npos) {
      if (pa != string::npos && B.t->lens[pb] != x.ext) return false;
      x.sb = stB[pb], x.ext = B.t->lens[pb];
    }
    if (po != string::npos) {
      if (x.ext && out->lens[po] != x.ext) return false;
      x.so = stO[po], x.ext = out->lens[po];
    }
    if (x.ext == 0) return true; /* empty contraction: nothing to add */
    work *= x.ext;
    const bool a = x.sa >= 0, b = x.sb >= 0, o = x.so >= 0;
    if (a && b && o) loops.push_back(x); /* batch (Hadamard) index */
    else if (a && !b && o) Mg.push_back(x);
    else if (!a && b && o) Ng.push_back(x);
    else if (a && b && !o) Kg.push_back(x);
    else return false; /* summed over one operand only, or broadcast: plain loops */
  }
  if (work < 4096) return false; /* tiny: not worth a library call */
  int64_t am, cm, bn, cn, ak, bk;
  const int64_t M = merge_group(Mg, 0, 2, am, cm, loops);
  const int64_t N = merge_group(Ng, 1, 2, bn, cn, loops);
  const int64_t K = merge_group(Kg, 1, 0, bk, ak, loops);
  /* leftover indices become loops; absent strides count as 0 */
  for (auto &x : loops) {
    if (x.sa < 0) x.sa = 0;
    if (x.sb < 0) x.sb = 0;
    if (x.so < 0) x.so = 0;
  }
  /* orientation: C column-major needs a unit stride along m (or n, then C^T = B^T A^T) */
  bool swap_roles = false;
  if (!(cm == 1 || M == 1)) {
    if (cn == 1 || N == 1) swap_roles = true;
    else return false;
  }
  int64_t m = M, n = N, k = K, a_m = am, a_k = ak, b_k = bk, b_n = bn, c_m = cm, c_n = cn;
  const double *pa0 = A.t->data.data(), *pb0 = B.t->data.data();
  vector<IdxInfo> lp = loops;
  if (swap_roles) { /* C^T (n x m) = B^T (n x k) A^T (k x m) */
    m = N, n = M;
    a_m = bn, a_k = bk, b_k = ak, b_n = am, c_m = cn, c_n = cm;
    std::swap(pa0, pb0);
    for (auto &x : lp) std::swap(x.sa, x.sb);
  }
  bool ta, tb;
  int64_t lda, ldb, ldc;
  if (a_m == 1 || m == 1) ta = false, lda = (k == 1 ? std::max<int64_t>(m, 1) : a_k);
  else if (a_k == 1 || k == 1) ta = true, lda = a_m;
  else return false;
  if (b_k == 1 || k == 1) tb = false, ldb = (n == 1 ? std::max<int64_t>(k, 1) : b_n);
  else if (b_n == 1 || n == 1) tb = true, ldb = b_k;
  else return false;
  ldc = (n == 1 ? std::max<int64_t>(m, 1) : c_n);
  if (lda < (ta ? k : m) || ldb < (tb ? n : k) || ldc < m) return false;
  int64_t nloop = 1;
  for (auto &x : lp) nloop *= x.ext;
  /* indices that are summed (K leftovers) make different loop iterations write the same C: those stay serial */
  bool loops_disjoint = true;
  for (auto &x : lp)
    if (x.so == 0 && x.ext > 1) loops_disjoint = false;
  auto run = [&](int64_t it0, int64_t it1) {
    vector<int64_t> ctr(lp.size(), 0);
    int64_t oa = 0, ob = 0, oc = 0, rem = it0;
    for (size_t d = 0; d < lp.size(); d++) {
      ctr[d] = rem % lp[d].ext;
      rem /= lp[d].ext;
      oa += ctr[d] * lp[d].sa, ob += ctr[d] * lp[d].sb, oc += ctr[d] * lp[d].so;
    }
    for (int64_t it = it0; it < it1; it++) {
      bl.gemm(ta, tb, m, n, k, q.coef, pa0 + oa, lda, pb0 + ob, ldb, acc.data() + oc, ldc);
      for (size_t d = 0; d < lp.size(); d++) {
        ctr[d]++;
        oa += lp[d].sa, ob += lp[d].sb, oc += lp[d].so;
        if (ctr[d] < lp[d].ext) break;
        oa -= lp[d].sa * lp[d].ext, ob -= lp[d].sb * lp[d].ext, oc -= lp[d].so * lp[d].ext;
        ctr[d] = 0;
      }
    }
  };
  /* many small products: one BLAS thread each, the loop spread over the cores; few large ones: BLAS threads */
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const char *nt_env = getenv("OMP_NUM_THREADS");
  const unsigned nthr = nt_env ? (unsigned)std::max(1, atoi(nt_env)) : hw;
  const double flops_each = 2.0 * (double)m * (double)n * (double)k;
  if (loops_disjoint && nthr > 1 && nloop >= 4 * (int64_t)nthr && flops_each < 2e8 && bl.set_threads) {
    bl.set_threads(1);
    vector<std::thread> pool;
    for (unsigned t = 0; t < nthr; t++)
      pool.emplace_back(run, nloop * t / nthr, nloop * (t + 1) / nthr);
    for (auto &th : pool) th.join();
    bl.set_threads((int)nthr);
  } else {
    run(0, nloop);
  }
  return true;
}
/* out (+)= coef * A with identical index strings: a plain vector update (the residual V - V_build of als_CP.cxx:183-187) */
inline bool accumulate_axpy(vector<double> &acc, const TensorBase *out, const string &oidx, const Prod &q) {
  if (!blas().ok() || q.f.size() != 1 || q.f[0].idx != oidx || q.f[0].t->data.size() != acc.size()) return false;
  for (size_t i = 0; i < oidx.size(); i++)
    if (oidx.find(oidx[i]) != i) return false;
  const double *a = q.f[0].t->data.data();
  const double c = q.coef;
  const size_t n = acc.size();
  const unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < (1u << 22)) {
    for (size_t i = 0; i < n; i++) acc[i] += c * a[i];
  } else {
    vector<std::thread> pool;
    for (unsigned t = 0; t < nthr; t++)
      pool.emplace_back([&, t]() {
        for (size_t i = n * t / nthr; i < n * (t + 1) / nthr; i++) acc[i] += c * a[i];
      });
    for (auto &th : pool) th.join();
  }
  return true;
}

/* mode 0: out = rhs, +1: out += rhs, -1: out -= rhs.  The right-hand side is evaluated completely first (it may
 * mention the output tensor). */
inline void assign(TensorBase *out, const string &oidx, const Term &rhs, int mode) {
  vector<double> acc(out->data.size(), 0.0);
  static const bool verbose = getenv("CTF_STANDIN_VERBOSE") != nullptr;
  for (const auto &q : rhs.p) {
    const double t0 = verbose ? MPI_Wtime() : 0.0;
    int path = 0;
    if (accumulate_blas(acc, out, oidx, q)) path = 1;
    else if (accumulate_axpy(acc, out, oidx, q)) path = 2;
    else accumulate(acc, out, oidx, q);
    if (verbose && MPI_Wtime() - t0 > 0.05) {
      string desc;
      for (const auto &l : q.f) desc += l.idx + " ";
      fprintf(stderr, "ctf stand-in: %s-> %s  %s  %.3f s\n", desc.c_str(), oidx.c_str(),
              path == 1 ? "dgemm" : (path == 2 ? "axpy" : "loops"), MPI_Wtime() - t0);
    }
  }
  bool repeated = false;
  for (size_t i = 0; i < oidx.size(); i++)
    if (oidx.find(oidx[i]) != i) repeated = true;
  if (mode == 0 && !repeated) {
    out->data.swap(acc);
  } else if (mode == 0) { /* only the addressed (diagonal) positions are overwritten */
    vector<double> mask(out->data.size(), 0.0);
    Prod one;
    accumulate(mask, out, oidx, one);
    for (size_t i = 0; i < acc.size(); i++)
      if (mask[i] != 0.0) out->data[i] = acc[i];
  } else {
    for (size_t i = 0; i < acc.size(); i++) out->data[i] += mode * acc[i];
  }
}
} // namespace standin

inline Term::operator double() const {
  TensorBase sc;
  sc.order = 0;
  sc.data.assign(1, 0.0);
  standin::assign(&sc, "", *this, 0);
  return sc.data[0];
}

class Idx_Tensor : public Term {
public:
  TensorBase *parent;
  string idx;
  /* like CTF, exactly `order` characters of the index string are read (build_V passes one without its
   * terminator, common.cxx:184-193) */
  Idx_Tensor(TensorBase *t, const char *s) : parent(t), idx(s, (size_t)t->order) {
    Prod q;
    q.f.push_back(Leaf{t, idx});
    p.push_back(q);
  }
  Idx_Tensor(const Idx_Tensor &) = default;
  void operator=(const Term &B) { standin::assign(parent, idx, B, 0); }
  void operator=(const Idx_Tensor &B) { standin::assign(parent, idx, (const Term &)B, 0); }
  void operator=(double c) { standin::assign(parent, idx, Term(c), 0); }
  void operator+=(const Term &B) { standin::assign(parent, idx, B, +1); }
  void operator-=(const Term &B) { standin::assign(parent, idx, B, -1); }
  void operator+=(double c) { standin::assign(parent, idx, Term(c), +1); }
  void operator-=(double c) { standin::assign(parent, idx, Term(c), -1); }
  void operator*=(double c) {
    for (double &x : parent->data) x *= c;
  }
};

template <typename A = double, typename B = double> class Transform {
public:
  std::function<void(double &)> f;
  template <typename Fn> Transform(Fn fn) : f(fn) {}
  void operator()(Idx_Tensor t) const {
    for (double &x : t.parent->data) f(x);
  }
};

/* ---------------------------------------------------------------- Tensor / Matrix / Vector */
template <typename dtype = double> class Tensor : public TensorBase {
public:
  Tensor() {}
  Tensor(int order_, const int *l, World &w = get_universe()) { init(order_, l, &w); }
  Tensor(int order_, const int64_t *l, World &w = get_universe()) { init(order_, l, &w); }
  Tensor(int order_, const int *l, const int *, World &w = get_universe()) { init(order_, l, &w); }
  Tensor(int order_, const int64_t *l, const int *, World &w = get_universe()) { init(order_, l, &w); }
  Tensor(int order_, bool sparse, const int *l, World &w = get_universe()) {
    assert(!sparse && "the stand-in is dense only");
    init(order_, l, &w);
  }
  Tensor(int order_, bool sparse, const int64_t *l, World &w = get_universe()) {
    assert(!sparse && "the stand-in is dense only");
    init(order_, l, &w);
  }
  Tensor(const Tensor &o) = default;
  Tensor(const TensorBase &o) : TensorBase(o) {}
  /* "same data, different symmetry": symmetric storage is not modelled (only Gauss_Seidel, off the path, asks) */
  Tensor(const TensorBase &o, const int *) : TensorBase(o) {}
  Tensor &operator=(const Tensor &o) = default;
  Idx_Tensor operator[](const char *s) { return Idx_Tensor(this, s); }
  Idx_Tensor operator[](const string &s) { return Idx_Tensor(this, s.c_str()); }
  Idx_Tensor i(const char *s) { return Idx_Tensor(this, s); }
};

template <typename dtype> class Vector;

template <typename dtype = double> class Matrix : public Tensor<dtype> {
public:
  int64_t nrow = 0, ncol = 0;
  Matrix() {}
  void init2(int64_t r, int64_t c, World *w) {
    int64_t l[2] = {r, c};
    this->init(2, l, w);
    nrow = r;
    ncol = c;
  }
  Matrix(int64_t r, int64_t c) { init2(r, c, nullptr); }
  Matrix(int64_t r, int64_t c, World &w) { init2(r, c, &w); }
  Matrix(int64_t r, int64_t c, int, World &w = get_universe()) { init2(r, c, &w); }
  Matrix(const Matrix &o) = default;
  Matrix(const TensorBase &o) : Tensor<dtype>(o) {
    assert(o.order == 2);
    nrow = o.lens[0];
    ncol = o.lens[1];
  }
  Matrix &operator=(const Matrix &o) = default;
  double &at(int64_t i, int64_t j) { return this->data[i + nrow * j]; }
  double at(int64_t i, int64_t j) const { return this->data[i + nrow * j]; }

  /* this = U diag(S) VT, singular values in decreasing order, truncated to `rank` (0 = all).  One-sided Jacobi. */
  void svd(Matrix &U, Vector<dtype> &S, Matrix &VT, int rank = 0, double = 0.0) const;
  void qr(Matrix &Q, Matrix &R) const;
  /* this = L L^T, L lower triangular */
  void cholesky(Matrix &L, bool lower = true) const;
  /* solves op(L) X = this (from_left) or X op(L) = this; op(L) = L^T when transp_L; only the `lower` (or upper)
   * triangle of L is referenced */
  void solve_tri(Matrix &L, Matrix &X, bool lower = true, bool from_left = true, bool transp_L = false) const;
};

template <typename dtype = double> class Vector : public Tensor<dtype> {
public:
  int64_t len = 0;
  Vector() {}
  void init1(int64_t n, World *w) {
    int64_t l[1] = {n};
    this->init(1, l, w);
    len = n;
  }
  Vector(int64_t n) { init1(n, nullptr); }
  Vector(int64_t n, World &w) { init1(n, &w); }
  Vector(const Vector &o) = default;
  Vector(const TensorBase &o) : Tensor<dtype>(o) {
    assert(o.order == 1);
    len = o.lens[0];
  }
  Vector &operator=(const Vector &o) = default;
};

namespace standin {
/* thin SVD of the m x n column-major matrix a (m >= n): a = u diag(s) v^T by Hestenes rotations of the columns */
inline void jacobi_svd_tall(int64_t m, int64_t n, vector<double> g, vector<double> &u, vector<double> &s,
                            vector<double> &v) {
  v.assign((size_t)(n * n), 0.0);
  for (int64_t j = 0; j < n; j++) v[j + n * j] = 1.0;
  const double eps = 2.220446049250313e-16;
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
    for (int64_t p = 0; p < n - 1; p++)
      for (int64_t q = p + 1; q < n; q++) {
        long double app = 0, aqq = 0, apq = 0;
        const double *gp = &g[m * p], *gq = &g[m * q];
        for (int64_t i = 0; i < m; i++) {
          app += (long double)gp[i] * gp[i];
          aqq += (long double)gq[i] * gq[i];
          apq += (long double)gp[i] * gq[i];
        }
        if (apq == 0 || fabsl(apq) <= eps * sqrtl(app * aqq)) continue;
        rotated = true;
        long double zeta = (aqq - app) / (2 * apq);
        long double t = (zeta >= 0 ? 1.0L : -1.0L) / (fabsl(zeta) + sqrtl(1 + zeta * zeta));
        double c = (double)(1 / sqrtl(1 + t * t)), sn = (double)(t / sqrtl(1 + t * t));
        double *wp = &g[m * p], *wq = &g[m * q];
        for (int64_t i = 0; i < m; i++) {
          double x = wp[i], y = wq[i];
          wp[i] = c * x - sn * y;
          wq[i] = sn * x + c * y;
        }
        double *vp = &v[n * p], *vq = &v[n * q];
        for (int64_t i = 0; i < n; i++) {
          double x = vp[i], y = vq[i];
          vp[i] = c * x - sn * y;
          vq[i] = sn * x + c * y;
        }
      }
    if (!rotated) break;
  }
  vector<double> nrm(n);
  for (int64_t j = 0; j < n; j++) {
    long double a = 0;
    for (int64_t i = 0; i < m; i++) a += (long double)g[i + m * j] * g[i + m * j];
    nrm[j] = (double)sqrtl(a);
  }
  vector<int64_t> ord(n);
  iota(ord.begin(), ord.end(), 0);
  stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return nrm[a] > nrm[b]; });
  u.assign((size_t)(m * n), 0.0);
  s.assign((size_t)n, 0.0);
  vector<double> v2((size_t)(n * n));
  const double tiny = (n ? nrm[ord[0]] : 0.0) * 1e-300;
  for (int64_t j = 0; j < n; j++) {
    const int64_t o = ord[j];
    s[j] = nrm[o];
    for (int64_t i = 0; i < n; i++) v2[i + n * j] = v[i + n * o];
    if (nrm[o] > tiny) {
      for (int64_t i = 0; i < m; i++) u[i + m * j] = g[i + m * o] / nrm[o];
    } else { /* null direction: any unit vector orthogonal to the previous columns */
      for (int64_t e = 0; e < m; e++) {
        vector<double> c(m, 0.0);
        c[e] = 1.0;
        for (int pass = 0; pass < 2; pass++)
          for (int64_t jj = 0; jj < j; jj++) {
            double d = 0;
            for (int64_t i = 0; i < m; i++) d += u[i + m * jj] * c[i];
            for (int64_t i = 0; i < m; i++) c[i] -= d * u[i + m * jj];
          }
        double nn = 0;
        for (int64_t i = 0; i < m; i++) nn += c[i] * c[i];
        if (nn > 0.25) {
          nn = sqrt(nn);
          for (int64_t i = 0; i < m; i++) u[i + m * j] = c[i] / nn;
          break;
        }
      }
    }
  }
  v.swap(v2);
}
/* solve T Y = B in place, T k x k triangular (column-major), B k x nb */
inline void trsm_left(int64_t k, const vector<double> &T, bool lowerT, int64_t nb, vector<double> &B) {
  for (int64_t c = 0; c < nb; c++) {
    double *b = &B[k * c];
    if (lowerT) {
      for (int64_t i = 0; i < k; i++) {
        long double a = b[i];
        for (int64_t j = 0; j < i; j++) a -= (long double)T[i + k * j] * b[j];
        b[i] = (double)(a / T[i + k * i]);
      }
    } else {
      for (int64_t i = k - 1; i >= 0; i--) {
        long double a = b[i];
        for (int64_t j = i + 1; j < k; j++) a -= (long double)T[i + k * j] * b[j];
        b[i] = (double)(a / T[i + k * i]);
      }
    }
  }
}
} // namespace standin

template <typename dtype>
void Matrix<dtype>::svd(Matrix &U, Vector<dtype> &S, Matrix &VT, int rank, double) const {
  const int64_t m = nrow, n = ncol, k = std::min(m, n);
  int64_t r = (rank <= 0 || rank > k) ? k : rank;
  vector<double> u, s, v;
  if (m >= n) {
    standin::jacobi_svd_tall(m, n, this->data, u, s, v); /* u m x n, v n x n */
  } else {
    vector<double> at((size_t)(m * n));
    for (int64_t i = 0; i < m; i++)
      for (int64_t j = 0; j < n; j++) at[j + n * i] = this->data[i + m * j];
    vector<double> ut, vt;
    standin::jacobi_svd_tall(n, m, at, ut, s, vt); /* A^T = ut diag(s) vt^T  ->  A = vt diag(s) ut^T */
    u = vt;                                          /* m x m */
    v = ut;                                          /* n x m */
  }
  Matrix Uo(m, r), VTo(r, n);
  Vector<dtype> So(r);
  for (int64_t j = 0; j < r; j++) {
    So.data[j] = s[j];
    for (int64_t i = 0; i < m; i++) Uo.data[i + m * j] = u[i + m * j];
    for (int64_t i = 0; i < n; i++) VTo.data[j + r * i] = v[i + n * j];
  }
  U = Uo;
  S = So;
  VT = VTo;
}

template <typename dtype> void Matrix<dtype>::qr(Matrix &Q, Matrix &R) const {
  const int64_t m = nrow, n = ncol;
  assert(m >= n);
  Matrix Qo(m, n), Ro(n, n);
  vector<double> a = this->data;
  for (int64_t j = 0; j < n; j++) {
    double *aj = &a[m * j];
    for (int pass = 0; pass < 2; pass++)
      for (int64_t p = 0; p < j; p++) {
        long double d = 0;
        for (int64_t i = 0; i < m; i++) d += (long double)Qo.data[i + m * p] * aj[i];
        for (int64_t i = 0; i < m; i++) aj[i] -= (double)d * Qo.data[i + m * p];
        Ro.data[p + n * j] += (double)d;
      }
    long double nn = 0;
    for (int64_t i = 0; i < m; i++) nn += (long double)aj[i] * aj[i];
    double nr = (double)sqrtl(nn);
    Ro.data[j + n * j] = nr;
    for (int64_t i = 0; i < m; i++) Qo.data[i + m * j] = nr > 0 ? aj[i] / nr : 0.0;
  }
  Q = Qo;
  R = Ro;
}

template <typename dtype> void Matrix<dtype>::cholesky(Matrix &L, bool lower) const {
  const int64_t n = nrow;
  assert(nrow == ncol);
  Matrix Lo(n, n);
  for (int64_t j = 0; j < n; j++) {
    long double d = this->data[j + n * j];
    for (int64_t k = 0; k < j; k++) d -= (long double)Lo.data[j + n * k] * Lo.data[j + n * k];
    double djj = (double)sqrtl(d);
    Lo.data[j + n * j] = djj;
    for (int64_t i = j + 1; i < n; i++) {
      long double a = this->data[i + n * j];
      for (int64_t k = 0; k < j; k++) a -= (long double)Lo.data[i + n * k] * Lo.data[j + n * k];
      Lo.data[i + n * j] = (double)(a / djj);
    }
  }
  if (!lower) {
    Matrix Uo(n, n);
    for (int64_t i = 0; i < n; i++)
      for (int64_t j = 0; j < n; j++) Uo.data[j + n * i] = Lo.data[i + n * j];
    L = Uo;
  } else {
    L = Lo;
  }
}

template <typename dtype>
void Matrix<dtype>::solve_tri(Matrix &L, Matrix &X, bool lower, bool from_left, bool transp_L) const {
  const int64_t k = L.nrow;
  assert(L.nrow == L.ncol);
  /* T = op(L) restricted to the referenced triangle */
  vector<double> T((size_t)(k * k), 0.0);
  for (int64_t j = 0; j < k; j++)
    for (int64_t i = 0; i < k; i++) {
      const bool in_tri = lower ? (i >= j) : (i <= j);
      if (!in_tri) continue;
      if (transp_L)
        T[j + k * i] = L.data[i + k * j];
      else
        T[i + k * j] = L.data[i + k * j];
    }
  bool lowerT = (lower != transp_L);
  const int64_t m = nrow, n = ncol;
  Matrix Xo(m, n);
  if (from_left) {
    assert(m == k);
    vector<double> B = this->data;
    standin::trsm_left(k, T, lowerT, n, B);
    Xo.data = B;
  } else { /* X T = B  <=>  T^T X^T = B^T */
    assert(n == k);
    vector<double> Tt((size_t)(k * k)), Bt((size_t)(m * n));
    for (int64_t i = 0; i < k; i++)
      for (int64_t j = 0; j < k; j++) Tt[j + k * i] = T[i + k * j];
    for (int64_t i = 0; i < m; i++)
      for (int64_t j = 0; j < n; j++) Bt[j + n * i] = this->data[i + m * j];
    standin::trsm_left(k, Tt, !lowerT, m, Bt);
    for (int64_t i = 0; i < m; i++)
      for (int64_t j = 0; j < n; j++) Xo.data[i + m * j] = Bt[j + n * i];
  }
  X = Xo;
}

/* ---------------------------------------------------------------- timers (no-ops) */
class Timer {
public:
  Timer(const char *) {}
  Timer(const string &) {}
  void start() {}
  void stop() {}
  void exit() {}
};
class Timer_epoch {
public:
  Timer_epoch(const char *) {}
  void begin() {}
  void end() {}
};

} // namespace CTF

#endif
