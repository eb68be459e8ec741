// K4b (SIMT, fp32, any shape): out[e,o,p] (+)= sum_{i,f} R[e,o,i,f] * T[e,i,f,p],
//   R[e,(o,i,f)] = W3[(o*Ci+i)*F+f, :] . g[e,:] + b3[...]          (RadialFunc.net.6, S:294, 299)
// i.e. PairwiseConv.forward (S:326-343) + the per-edge mat-vec of ConvSE3 (S:251-254) in factored form.
// R is produced 64 rows x 32 edges at a time in registers and consumed immediately; it never reaches memory.
// This kernel is the correctness anchor and the path for shapes the tensor-core kernel does not take.
#include "common.cuh"

namespace se3 {

constexpr int kSE = 32;     // edges per CTA
constexpr int kSO = 16;     // output channels per CTA
constexpr int kSR = kSO * SE3_TILE_IF;   // 64 weight rows per step
constexpr int kLd = SE3_RADIAL_MID + 4;  // padded smem row (conflict-free float4 rows)

template <int PH>
__global__ void __launch_bounds__(256)
pairwise_simt_kernel(const float* __restrict__ g, const float* __restrict__ W3, const float* __restrict__ b3,
                     const float* __restrict__ T, int64_t E, int Co, int CiF, int P, int accumulate,
                     float* __restrict__ out) {
  extern __shared__ __align__(16) float smem_simt[];
  float (*Ws)[kLd] = reinterpret_cast<float (*)[kLd]>(smem_simt);
  float (*gs)[kLd] = reinterpret_cast<float (*)[kLd]>(smem_simt + kSR * kLd);
  float* bs = smem_simt + (kSR + kSE) * kLd;
  const int tid = threadIdx.x;
  const int rr = tid & 63;          // weight row of the step: o_l*4 + if_l
  const int eg = tid >> 6;          // edge group (8 edges)
  const int64_t e0 = (int64_t)blockIdx.x * kSE;
  const int o0 = blockIdx.y * kSO;
  const int NIFB = (CiF + SE3_TILE_IF - 1) / SE3_TILE_IF;
  const int64_t mt = e0 / SE3_TILE_E;
  const int el0 = (int)(e0 % SE3_TILE_E) + eg * 8;
  const float4* Tt = reinterpret_cast<const float4*>(T) + (size_t)mt * NIFB * SE3_TILE_IF * PH * SE3_TILE_E;

  for (int t = tid; t < kSE * (SE3_RADIAL_MID / 4); t += 256) {
    const int e = t / (SE3_RADIAL_MID / 4), c4 = t % (SE3_RADIAL_MID / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e0 + e < E) v = *reinterpret_cast<const float4*>(g + (size_t)(e0 + e) * SE3_RADIAL_MID + c4 * 4);
    *reinterpret_cast<float4*>(&gs[e][c4 * 4]) = v;
  }

  float acc[8][PH * 4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int p = 0; p < PH * 4; ++p) acc[a][p] = 0.f;

  for (int ifb = 0; ifb < NIFB; ++ifb) {
    __syncthreads();
    for (int t = tid; t < kSR * (SE3_RADIAL_MID / 4); t += 256) {
      const int r = t / (SE3_RADIAL_MID / 4), c4 = t % (SE3_RADIAL_MID / 4);
      const int o = o0 + (r >> 2), ifx = ifb * SE3_TILE_IF + (r & 3);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o < Co && ifx < CiF) v = *reinterpret_cast<const float4*>(W3 + ((size_t)o * CiF + ifx) * SE3_RADIAL_MID + c4 * 4);
      *reinterpret_cast<float4*>(&Ws[r][c4 * 4]) = v;
    }
    if (tid < kSR) {
      const int o = o0 + (tid >> 2), ifx = ifb * SE3_TILE_IF + (tid & 3);
      bs[tid] = (o < Co && ifx < CiF) ? b3[(size_t)o * CiF + ifx] : 0.f;
    }
    __syncthreads();
    float R[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) R[a] = 0.f;
#pragma unroll 4
    for (int c = 0; c < SE3_RADIAL_MID; c += 4) {
      const float4 w = *reinterpret_cast<const float4*>(&Ws[rr][c]);
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const float4 gv = *reinterpret_cast<const float4*>(&gs[eg * 8 + a][c]);
        R[a] = fmaf(w.x, gv.x, R[a]);
        R[a] = fmaf(w.y, gv.y, R[a]);
        R[a] = fmaf(w.z, gv.z, R[a]);
        R[a] = fmaf(w.w, gv.w, R[a]);
      }
    }
    const float bias = bs[rr];
    const float4* Tb = Tt + ((size_t)ifb * SE3_TILE_IF + (rr & 3)) * PH * SE3_TILE_E + el0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const float r = R[a] + bias;
#pragma unroll
      for (int ph = 0; ph < PH; ++ph) {
        const float4 tv = Tb[(size_t)ph * SE3_TILE_E + a];
        acc[a][ph * 4 + 0] = fmaf(r, tv.x, acc[a][ph * 4 + 0]);
        acc[a][ph * 4 + 1] = fmaf(r, tv.y, acc[a][ph * 4 + 1]);
        acc[a][ph * 4 + 2] = fmaf(r, tv.z, acc[a][ph * 4 + 2]);
        acc[a][ph * 4 + 3] = fmaf(r, tv.w, acc[a][ph * 4 + 3]);
      }
    }
  }
  // reduce over the 4 (i,f) lanes of each output channel (adjacent lanes) with shuffles
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int p = 0; p < PH * 4; ++p) {
      float v = acc[a][p];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      acc[a][p] = v;
    }
  const int o = o0 + (rr >> 2);
  if ((rr & 3) == 0 && o < Co) {
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int64_t e = e0 + eg * 8 + a;
      if (e < E) {
        float* dst = out + ((size_t)e * Co + o) * P;
#pragma unroll
        for (int p = 0; p < PH * 4; ++p)
          if (p < P) dst[p] = accumulate ? dst[p] + acc[a][p] : acc[a][p];
      }
    }
  }
}

}  // namespace se3

extern "C" int se3_pairwise_simt_fwd(const float* g, const float* W3, const float* b3, const float* T, int64_t E, int Co,
                                     int Ci, int F, int P, int accumulate, float* out, void* stream) {
  using namespace se3;
  SE3_REQUIRE(E > 0 && Co > 0 && Ci > 0 && F > 0, "se3_pairwise_simt_fwd: bad sizes");
  SE3_REQUIRE(P >= 1 && P <= 11, "se3_pairwise_simt_fwd: P=%d out of range", P);
  const int CiF = Ci * F;
  dim3 grid((unsigned)ceil_div(E, kSE), (unsigned)ceil_div(Co, kSO));
  cudaStream_t s = as_stream(stream);
  const size_t smem = ((size_t)(kSR + kSE) * kLd + kSR) * sizeof(float);
  switch ((P + 3) / 4) {
    case 1:
      SE3_CUDA_OK(cudaFuncSetAttribute(pairwise_simt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pairwise_simt_kernel<1><<<grid, 256, smem, s>>>(g, W3, b3, T, E, Co, CiF, P, accumulate, out);
      break;
    case 2:
      SE3_CUDA_OK(cudaFuncSetAttribute(pairwise_simt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pairwise_simt_kernel<2><<<grid, 256, smem, s>>>(g, W3, b3, T, E, Co, CiF, P, accumulate, out);
      break;
    default:
      SE3_CUDA_OK(cudaFuncSetAttribute(pairwise_simt_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pairwise_simt_kernel<3><<<grid, 256, smem, s>>>(g, W3, b3, T, E, Co, CiF, P, accumulate, out);
      break;
  }
  SE3_LAUNCH_OK();
  return SE3_OK;
}
