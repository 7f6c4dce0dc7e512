/* standin_selftest.cxx -- TEST INFRASTRUCTURE: exercises oracle/ctf_standin/ctf.hpp on its own (no reference code), so
 * that the checker the reference sources run on is itself checked against NumPy (tests/test_reference_pin.py).
 * Inputs come from fill_random = the repo's counter-based generator with (seed 11, id = call number); every result
 * is written to <prefix>.<name>.bin as raw doubles in first-index-fastest order. */
#include <ctf.hpp>
using namespace CTF;

static std::string g_prefix;
static void dump(const char *name, TensorBase &t) {
  FILE *f = fopen((g_prefix + "." + name + ".bin").c_str(), "wb");
  fwrite(t.data.data(), sizeof(double), t.data.size(), f);
  fclose(f);
}

int main(int argc, char **argv) {
  g_prefix = argc > 1 ? argv[1] : "selftest";
  World dw(argc, argv);
  int l4[4] = {5, 4, 3, 6};
  Tensor<> V(4, l4, dw);
  V.fill_random(-1, 1);                       // call 0
  Matrix<> A(5, 3, dw), B(4, 3, dw), C(3, 3, dw), D(6, 3, dw);
  A.fill_random(0, 1);                        // 1
  B.fill_random(0, 1);                        // 2
  C.fill_random(0, 1);                        // 3
  D.fill_random(0, 1);                        // 4
  dump("V", V);
  // 1. tensor times matrix, rank index last
  int l1[4] = {5, 4, 6, 3};
  Tensor<> T1(4, l1, dw);
  T1["abd*"] = V["abcd"] * C["c*"];
  dump("ttm", T1);
  // 2. Hadamard-batched contraction (the rank index in all three operands)
  int l2[3] = {5, 4, 3};
  Tensor<> T2(3, l2, dw);
  T2["ab*"] = T1["abd*"] * D["d*"];
  dump("mttv", T2);
  // 3. three factors at once, += and -=, scalar factors
  Matrix<> M(5, 3, dw);
  M["a*"] = T2["ab*"] * B["b*"];
  M["a*"] += 2.0 * A["a*"];
  M["a*"] -= 0.5 * (A["a*"] - M["a*"]);
  dump("chain", M);
  // 4. Gram, Hadamard of Grams, the right-hand side mentioning the output, diagonal assignment
  Matrix<> S(3, 3, dw), reg(3, 3, dw);
  S["ij"] = A["ki"] * A["kj"];
  S["ij"] = S["ij"] * (B["ki"] * B["kj"]);
  reg["ii"] = 1. * 0.25;
  S["ij"] += reg["ij"];
  dump("S", S);
  // 5. diagonal extraction, Transform, diagonal write (als_Tucker.cxx:634-642)
  Vector<> dg(3, dw);
  dg["i"] = S["ii"];
  Transform<double, double>([](double &b) { b = b > 2.0 ? 1 : -1; })(dg["i"]);
  Matrix<> Dg(3, 3, dw);
  Dg["ii"] = dg["i"];
  dump("diag", Dg);
  // 6. full contraction to a scalar, norm2
  double ip = A["ij"] * A["ij"];
  Vector<> sc(2, dw);
  int64_t idx[2] = {0, 1};
  double vals[2] = {ip, V.norm2()};
  sc.write(2, idx, vals);
  dump("scalars", sc);
  // 7. Gram of an unfolding with the reference's index characters (common.cxx:205-223)
  Matrix<> G(4, 4, dw);
  G["^&"] = V["i^kl"] * V["i&kl"];
  dump("unfold_gram", G);
  // 8. dense linear algebra
  Matrix<> U, VT, Q, Rr, L, X;
  Vector<> sv;
  Matrix<> Wd(6, 4, dw);
  Wd.fill_random(-1, 1);                      // 5
  Wd.svd(U, sv, VT, 3);
  dump("svd_in", Wd);
  dump("svd_U", U);
  dump("svd_s", sv);
  dump("svd_VT", VT);
  Matrix<> Wt(3, 5, dw);                      // wide
  Wt.fill_random(-1, 1);                      // 6
  Wt.svd(U, sv, VT, 3);
  dump("svdw_in", Wt);
  dump("svdw_U", U);
  dump("svdw_s", sv);
  dump("svdw_VT", VT);
  Wd.qr(Q, Rr);
  dump("qr_Q", Q);
  dump("qr_R", Rr);
  S.cholesky(L);
  dump("chol_L", L);
  A.solve_tri(L, X, true, false, true);       // X L^T = A
  dump("tri_right_T", X);
  A.solve_tri(L, X, true, false, false);      // X L = A
  dump("tri_right_N", X);
  Matrix<> At(3, 5, dw);
  At["ij"] = A["ji"];
  At.solve_tri(L, X, true, true, false);      // L X = A^T
  dump("tri_left_N", X);
  At.solve_tri(L, X, true, true, true);       // L^T X = A^T
  dump("tri_left_T", X);
  printf("selftest done\n");
  return 0;
}
