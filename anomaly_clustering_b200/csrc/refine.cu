// Exact re-evaluation of the nearest-neighbour distances the tensor-core pass SELECTED (precision mode "f16r").
//
// The tcgen05 pass (mindist_tc.cu) finds, per (query row r, bank image j), the nearest bank row -- torch.cdist +
// torch.min(dim=1) of Weight_Distance_* (reference: Anomaly-Clustering/models/patchcore/utils.py:226, :233-234) -- but
// its distance carries the tensor core's fp32 accumulation error over K = D products of magnitude |q||b| (systematic,
// ~3e-3 absolute at config 2: too much for softmax(w / tau) once tau < 1).  The arg-min is far more robust than the
// value, so that pass also records WHICH row won (ac_min_dist_arg / ac_min_dist_sym_arg) and this kernel recomputes
//      d(r, j) = sqrt( sum_k (q_r[k] - b_(j,c)[k])^2 ),   c = arg-min row of image j for query row r,
// in fp32 with no |x|^2+|y|^2-2xy cancellation: one gathered bank row (fp16/bf16 operand, + its lo part if present)
// per pair against the query row held in shared memory as fp32 (from fp32 Z when the caller has it, else from the
// operand).  What is left is the zero-mean rounding of the operands (~1e-4 per entry, averaged away by the mean over
// bank images).  Cost: one 2*D-byte row per pair from L2 -- 64 GB at config 2 against 18 ms of GEMM.
//
// Blocking: grid = (query chunks of kRQ rows, groups of Jb bank images); blocks are dispatched x-fastest, so all
// resident blocks gather from the same Jb bank images (they stay in L2) while the query chunks stream past once per group.
#include "common.cuh"
#include <algorithm>

namespace ac {

static constexpr int kRQ = 4;            // query rows per block (fp32 in shared memory)
static constexpr int kRefThreads = 256;  // 8 warps: two per query row, taking alternate bank images

__host__ __device__ inline bool pair_owned_r(int i, int j, int N) {   // same rule as mindist_tc.cu: pair_owned
  int d = j - i;
  if (d < 0) d += N;
  if (d == 0) return false;
  if (2 * d < N) return true;
  if (2 * d == N) return i < j;
  return false;
}

struct RefineParams {
  const float* Zq;                    // [Mq, D] fp32 or null
  const void* Qhi; const void* Qlo;   // [Mq, D] operand copies (used when Zq is null)
  const void* Bhi; const void* Blo;   // [nb_img*P, D]
  long long Mq;
  int nb_img, P, D, Pq;
  const int* rowarg;                  // [nb_img, Mq]
  const unsigned long long* colkey;   // [nb_img, Mq] (sym) or null
  int sym, q_img0;
  const int* q_self;                  // non-sym: bank index of each query image (its pair is skipped) or null
  const int* groups;                  // sym: (first image, count) of the category of every bank image, or null
  float* dex;                         // [nb_img, Mq]
  int Jb;
};

template <typename T> __device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]);
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 v = __half22float2(h[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
}
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 v = __bfloat1622float2(h[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
}

template <typename T>
__global__ void __launch_bounds__(kRefThreads) refine_kernel(const RefineParams p) {
  // query rows as fp32, split so that a lane's 8 values are two conflict-free 16-byte reads:
  // element 8g+t lives in lo4[g] (t < 4) or hi4[g] (t >= 4)
  extern __shared__ __align__(16) float4 s_q[];          // [kRQ][2][D/8]
  const int G8 = p.D >> 3;
  const long long m0 = (long long)blockIdx.x * kRQ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < kRQ * G8; e += kRefThreads) {
    const int rr = e / G8, g = e - rr * G8;
    const long long r = m0 + rr;
    float f[8];
    if (r >= p.Mq) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    } else if (p.Zq) {
      const float4* z = reinterpret_cast<const float4*>(p.Zq + r * p.D) + 2 * g;
      const float4 a = __ldg(z), b = __ldg(z + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
      unpack8<T>(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.Qhi) + r * p.D) + g), f);
      if (p.Qlo) {
        float l[8];
        unpack8<T>(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.Qlo) + r * p.D) + g), l);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += l[i];
      }
    }
    s_q[(rr * 2 + 0) * G8 + g] = make_float4(f[0], f[1], f[2], f[3]);
    s_q[(rr * 2 + 1) * G8 + g] = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();

  const int rr = warp >> 1, par = warp & 1;
  const long long r = m0 + rr;
  if (r >= p.Mq) return;
  const int qi = (int)(r / p.Pq);
  const int i = p.sym ? p.q_img0 + qi : (p.q_self ? __ldg(p.q_self + qi) : -1);
  const float4* qlo4 = s_q + (rr * 2 + 0) * G8;
  const float4* qhi4 = s_q + (rr * 2 + 1) * G8;
  const int j0 = blockIdx.y * p.Jb, j1 = min(p.nb_img, j0 + p.Jb);
  // categories: only the images of the query image's own category are its bank (the rest is never read by the reduction)
  int g0 = 0, gn = p.nb_img;
  if (p.sym && p.groups) { g0 = __ldg(p.groups + 2 * i); gn = __ldg(p.groups + 2 * i + 1); }
  for (int j = j0 + par; j < j1; j += 2) {
    if (j < g0 || j >= g0 + gn) continue;
    if (j == i) {
      if (lane == 0) p.dex[(long long)j * p.Mq + r] = 0.f;     // own image: excluded by the reduction
      continue;
    }
    int c = 0;
    if (lane == 0) {
      const long long e = (long long)j * p.Mq + r;
      if (p.sym && !pair_owned_r(i - g0, j - g0, gn)) c = (int)(unsigned int)(__ldg(p.colkey + e) & 0xffffffffull);
      else c = __ldg(p.rowarg + e);
      c = min(max(c, 0), p.P - 1);
    }
    c = __shfl_sync(0xffffffffu, c, 0);
    const uint4* bh = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.Bhi) + ((long long)j * p.P + c) * p.D);
    const uint4* bl = p.Blo ? reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.Blo) + ((long long)j * p.P + c) * p.D) : nullptr;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int g = lane; g < G8; g += 32) {
      float b[8];
      unpack8<T>(__ldg(bh + g), b);
      if (bl) {
        float l[8];
        unpack8<T>(__ldg(bl + g), l);
#pragma unroll
        for (int t = 0; t < 8; ++t) b[t] += l[t];
      }
      const float4 ql = qlo4[g], qh = qhi4[g];
      float d;
      d = ql.x - b[0]; a0 = fmaf(d, d, a0);
      d = ql.y - b[1]; a1 = fmaf(d, d, a1);
      d = ql.z - b[2]; a2 = fmaf(d, d, a2);
      d = ql.w - b[3]; a3 = fmaf(d, d, a3);
      d = qh.x - b[4]; a0 = fmaf(d, d, a0);
      d = qh.y - b[5]; a1 = fmaf(d, d, a1);
      d = qh.z - b[6]; a2 = fmaf(d, d, a2);
      d = qh.w - b[7]; a3 = fmaf(d, d, a3);
    }
    const float s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) p.dex[(long long)j * p.Mq + r] = sqrtf(s);
  }
}

}  // namespace ac

using namespace ac;

static int g_refine_l2_mb = 64;   // debug knob (ac_debug_set key 5): bytes of bank operand rows one group keeps L2-resident
extern "C" int ac_debug_set_refine(int mb) {
  if (mb < 1 || mb > 512) return AC_ERR_INVALID;
  g_refine_l2_mb = mb;
  return AC_OK;
}

extern "C" int ac_refine_min_dist(const float* Zq, const void* Qhi, const void* Qlo, int64_t Mq, const void* Bhi, const void* Blo,
                                  int op_dtype, int nb_img, int P, int D, const int32_t* rowarg, const uint64_t* colkey, int sym,
                                  int q_img0, const int32_t* q_self, int Pq, const int32_t* groups, float* dmin, ac_stream_t stream) {
  if ((!Zq && !Qhi) || !Bhi || !rowarg || !dmin || Mq < 0 || nb_img < 1 || P < 1 || D < 1 || Pq < 1 || q_img0 < 0) return AC_ERR_INVALID;
  if (op_dtype != AC_DT_F16 && op_dtype != AC_DT_BF16) return AC_ERR_INVALID;
  if (sym && (!colkey || Pq != P)) return AC_ERR_INVALID;
  if (D % 8 != 0) return AC_ERR_UNSUPPORTED;                       // 16-byte operand vectors
  const size_t smem = (size_t)kRQ * D * sizeof(float);
  if (smem > 200 * 1024) return AC_ERR_UNSUPPORTED;
  int rc = check_device();
  if (rc) return rc;
  if (Mq == 0) return AC_OK;
  RefineParams p;
  p.Zq = Zq; p.Qhi = Qhi; p.Qlo = Qlo; p.Bhi = Bhi; p.Blo = Blo;
  p.Mq = Mq; p.nb_img = nb_img; p.P = P; p.D = D; p.Pq = Pq;
  p.rowarg = rowarg; p.colkey = (const unsigned long long*)colkey; p.sym = sym; p.q_img0 = q_img0; p.q_self = q_self; p.dex = dmin;
  p.groups = sym ? groups : nullptr;
  // bank images per group: their operand rows (64 MB by default) stay L2-resident while every query chunk passes
  const double img_bytes = (double)P * D * 2.0 * (Blo ? 2 : 1);
  p.Jb = (int)std::max(2.0, std::min(64.0, g_refine_l2_mb * 1.0e6 / img_bytes));
  p.Jb &= ~1;                                                      // the two warps of a row take alternate images
  const long long chunks = (Mq + kRQ - 1) / kRQ;
  const int nbg = (nb_img + p.Jb - 1) / p.Jb;
  if (chunks > 0x7fffffffLL || nbg > 65535) return AC_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)chunks, (unsigned)nbg);
  if (op_dtype == AC_DT_F16) {
    AC_CUDA(cudaFuncSetAttribute(refine_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    refine_kernel<__half><<<grid, kRefThreads, smem, st>>>(p);
  } else {
    AC_CUDA(cudaFuncSetAttribute(refine_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    refine_kernel<__nv_bfloat16><<<grid, kRefThreads, smem, st>>>(p);
  }
  AC_LAUNCH_CHECK();
  return AC_OK;
}
