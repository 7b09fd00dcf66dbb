/* Plain-C consumer of libac_b200.so: no Python, no torch -- the boundary is include/ac_b200.h only.
 * gcc abi_smoke.c -I<repo>/include -I/usr/local/cuda/include -L<pkg> -lac_b200 -L/usr/local/cuda/lib64 -lcudart -lm
 * Computes X = sum_p alpha_p Z_p and the pairwise Euclidean matrix for a tiny problem and checks both against
 * host loops.  Exit code 0 = OK. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ac_b200.h"

int main(void) {
  const int N = 5, P = 7, D = 12;
  if (ac_device_ok(0) != AC_OK) {
    printf("no sm_100 device: %s\n", ac_strerror(ac_device_ok(0)));
    return 2;
  }
  float *hZ = malloc(sizeof(float) * N * P * D), *hA = malloc(sizeof(float) * N * P);
  float *hX = malloc(sizeof(float) * N * D), *hD = malloc(sizeof(float) * N * N);
  for (int i = 0; i < N * P * D; ++i) hZ[i] = (float)((i * 37 % 101) - 50) / 25.0f;
  for (int i = 0; i < N; ++i) {
    float s = 0;
    for (int p = 0; p < P; ++p) s += (hA[i * P + p] = 1.0f + (float)((i + 3 * p) % 5));
    for (int p = 0; p < P; ++p) hA[i * P + p] /= s;
  }
  float *dZ, *dA, *dX, *dD;
  cudaMalloc((void**)&dZ, sizeof(float) * N * P * D);
  cudaMalloc((void**)&dA, sizeof(float) * N * P);
  cudaMalloc((void**)&dX, sizeof(float) * N * D);
  cudaMalloc((void**)&dD, sizeof(float) * N * N);
  cudaMemcpy(dZ, hZ, sizeof(float) * N * P * D, cudaMemcpyHostToDevice);
  cudaMemcpy(dA, hA, sizeof(float) * N * P, cudaMemcpyHostToDevice);
  int rc = ac_weighted_embed(dA, dZ, N, P, D, dX, NULL);
  if (rc) { printf("ac_weighted_embed: %s\n", ac_strerror(rc)); return 1; }
  rc = ac_pairwise_l2(dX, N, D, dD, NULL);
  if (rc) { printf("ac_pairwise_l2: %s\n", ac_strerror(rc)); return 1; }
  if (ac_pairwise_l2(NULL, N, D, dD, NULL) != AC_ERR_INVALID) { printf("null pointer not rejected\n"); return 1; }
  cudaDeviceSynchronize();
  cudaMemcpy(hX, dX, sizeof(float) * N * D, cudaMemcpyDeviceToHost);
  cudaMemcpy(hD, dD, sizeof(float) * N * N, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int i = 0; i < N; ++i)
    for (int d = 0; d < D; ++d) {
      double x = 0;
      for (int p = 0; p < P; ++p) x += (double)hA[i * P + p] * hZ[(i * P + p) * D + d];
      worst = fmax(worst, fabs(x - hX[i * D + d]));
    }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int d = 0; d < D; ++d) { double t = (double)hX[i * D + d] - hX[j * D + d]; s += t * t; }
      worst = fmax(worst, fabs(sqrt(s) - hD[i * N + j]));
    }
  printf("ac_version %d, max abs error %.3e\n", ac_version(), worst);
  return worst < 1e-5 ? 0 : 1;
}
