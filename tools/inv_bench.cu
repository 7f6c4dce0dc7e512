// Cycle-level breakdown of the one-CTA R x R inverse kernel (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DPPX_INV_PROFILE -I include -o /tmp/inv_bench tools/inv_bench.cu && /tmp/inv_bench 50
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ long long g_prof[64];
#include "../pairwise-perturbation_b200/csrc/k45_solve.cu"

int ppx_set_err(ppx_ctx *, int code, const char *, ...) { return code; }
void *ppx_ws_alloc(ppx_ctx *, size_t) { return nullptr; }
extern "C" int ppx_sqnorms(ppx_ctx *, const double *const *, const int64_t *, int, double *) { return 0; }

int main(int argc, char **argv) {
  int R = argc > 1 ? atoi(argv[1]) : 50;
  int s = 300;
  std::vector<double> W(s * R), G(R * R);
  srand(1);
  for (auto &w : W) w = rand() / (double)RAND_MAX;
  for (int a = 0; a < R; a++)
    for (int b = 0; b < R; b++) {
      double t = 0;
      for (int i = 0; i < s; i++) t += W[i + s * a] * W[i + s * b];
      G[a + R * b] = t;
    }
  double *dG, *dS, *dSi;
  cudaMalloc(&dG, 8 * R * R);
  cudaMalloc(&dS, 8 * R * R);
  cudaMalloc(&dSi, 8 * R * R);
  cudaMemcpy(dG, G.data(), 8 * R * R, cudaMemcpyHostToDevice);
  HadArgs h;
  h.n = 3;
  h.g[0] = h.g[1] = h.g[2] = dG;
  const int T = (R + INV_B - 1) / INV_B;
  const size_t smem = sizeof(double) * (2 * ((size_t)INV_B * T + 2) + (size_t)R * (R + 1));
  cudaFuncSetAttribute(spd_inverse_ldl_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(spd_inverse_ldl_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    switch (T) {
      case 1: spd_inverse_ldl_kernel<1><<<1, 256, smem>>>(h, R, 0.0, dS, dSi, nullptr); break;
      case 2: spd_inverse_ldl_kernel<2><<<1, 256, smem>>>(h, R, 0.0, dS, dSi, nullptr); break;
      case 3: spd_inverse_ldl_kernel<3><<<1, 256, smem>>>(h, R, 0.0, dS, dSi, nullptr); break;
      case 4: spd_inverse_ldl_kernel<4><<<1, 256, smem>>>(h, R, 0.0, dS, dSi, nullptr); break;
      default: spd_inverse_ldl_kernel<7><<<1, 256, smem>>>(h, R, 0.0, dS, dSi, nullptr); break;
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long prof[64];
    cudaMemcpyFromSymbol(prof, g_prof, sizeof(prof));
    printf("rep %d: %.1f us; cycles: load %lld, loop %lld (%.0f/step), store %lld", rep, ms * 1e3,
           prof[1] - prof[0], prof[2] - prof[1], (double)(prof[2] - prof[1]) / R, prof[4] - prof[2]);
    printf("\n");
  }
  if (R <= 64) {  // the register-resident sweep kernel (R <= 64)
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      switch ((R + 7) / 8) {
        case 2: spd_inverse_sweep_kernel<16><<<1, SWEEP_H * 32>>>(h, R, 0.0, dS, dSi); break;
        case 5: spd_inverse_sweep_kernel<40><<<1, SWEEP_H * 64>>>(h, R, 0.0, dS, dSi); break;
        case 7: spd_inverse_sweep_kernel<56><<<1, SWEEP_H * 64>>>(h, R, 0.0, dS, dSi); break;
        default: spd_inverse_sweep_kernel<64><<<1, SWEEP_H * 64>>>(h, R, 0.0, dS, dSi); break;
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      long long prof[64];
      cudaMemcpyFromSymbol(prof, g_prof, sizeof(prof));
      printf("sweep rep %d: %.1f us; cycles: load %lld, loop %lld (%.0f/step), store %lld; step R/2 of the publisher: "
             "read+u %lld, fma %lld, publish %lld, barrier %lld\n", rep, ms * 1e3, prof[9] - prof[8],
             prof[10] - prof[9], (double)(prof[10] - prof[9]) / R, prof[11] - prof[10], prof[17] - prof[16],
             prof[19] - prof[17], prof[20] - prof[19], prof[21] - prof[20]);
    }
  }
  // check: S * Sinv = I
  std::vector<double> S(R * R), Si(R * R);
  cudaMemcpy(S.data(), dS, 8 * R * R, cudaMemcpyDeviceToHost);
  cudaMemcpy(Si.data(), dSi, 8 * R * R, cudaMemcpyDeviceToHost);
  double err = 0;
  for (int a = 0; a < R; a++)
    for (int b = 0; b < R; b++) {
      double t = 0;
      for (int k = 0; k < R; k++) t += S[a + R * k] * Si[k + R * b];
      err = fmax(err, fabs(t - (a == b)));
    }
  printf("max |S Sinv - I| = %.3e (%s)\n", err, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
