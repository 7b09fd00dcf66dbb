// Exact fp32 patch-versus-bank nearest neighbour (AC_PREC_F32): no tensor cores, sum (x-y)^2.
//
// Replaces torch.cdist + torch.min(dim=1) of Weight_Distance_* (reference:
// Anomaly-Clustering/models/patchcore/utils.py:226, :233-234) for one (query block, bank image)
// pair per CTA.  This is the precision-reference mode on the GPU (the reference's own cdist uses
// the |x|^2+|y|^2-2xy expansion in fp32 and is *less* exact than this kernel) and the cross-check
// for the tcgen05 path; it is not the fast path.
#include "common.cuh"

namespace ac {

static constexpr int TM = 64, TN = 64, TK = 32;

__global__ void __launch_bounds__(256) mindist_simt_kernel(const float* __restrict__ Q, long long Mq, const float* __restrict__ Bk,
                                                           int P, int D, float* __restrict__ dmin) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int j = blockIdx.y;                       // bank image
  const long long m0 = (long long)blockIdx.x * TM;
  const float* Bj = Bk + (long long)j * P * D;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float best[4] = {INFINITY, INFINITY, INFINITY, INFINITY};

  for (int n0 = 0; n0 < P; n0 += TN) {
    float acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
    for (int k0 = 0; k0 < D; k0 += TK) {
#pragma unroll
      for (int it = 0; it < (TM * TK) / 256; ++it) {
        const int e = threadIdx.x + it * 256;
        const int r = e / TK, c = e - r * TK;
        const int d = k0 + c;
        const long long qr = m0 + r;
        As[c][r] = (qr < Mq && d < D) ? __ldg(Q + qr * D + d) : 0.f;
        const int br = n0 + r;
        Bs[c][r] = (br < P && d < D) ? __ldg(Bj + (long long)br * D + d) : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int c = 0; c < TK; ++c) {
        const float4 a = *reinterpret_cast<const float4*>(&As[c][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[c][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float t = av[u] - bv[v];
            acc[u][v] = fmaf(t, t, acc[u][v]);
          }
      }
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (n0 + tx * 4 + v < P) best[u] = fminf(best[u], acc[u][v]);
  }
  // combine the 16 threads (tx) that share a query row: they are 16 consecutive lanes
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    float b = best[u];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
    const long long qr = m0 + ty * 4 + u;
    if (tx == 0 && qr < Mq) dmin[(long long)j * Mq + qr] = sqrtf(b);
  }
}

int launch_mindist_simt(const float* Q, long long Mq, const float* Bk, int nb_img, int P, int D, float* dmin, cudaStream_t st) {
  dim3 grid((unsigned)((Mq + TM - 1) / TM), nb_img);
  mindist_simt_kernel<<<grid, 256, 0, st>>>(Q, Mq, Bk, P, D, dmin);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
