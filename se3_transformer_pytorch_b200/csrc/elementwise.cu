// NormSE3 (se3_transformer_pytorch.py:97-152), the norm nonlinearity that sits either side of the attention block
// (prenorm of AttentionBlockSE3 / FeedForwardBlockSE3, the nonlinearity inside FeedForwardSE3, the optional output
// norm): one pass over the features instead of the reference's ~7 elementwise ATen kernels.
//   norm = max(||x[b,n,c,:]||_2, eps);  out = nonlin(norm * scale[c]) * (x / norm)
#include "common.cuh"

namespace se3 {

__device__ __forceinline__ float gelu_erf_e(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int M>
__global__ void __launch_bounds__(256)
norm_kernel(const float* __restrict__ x, const float* __restrict__ scale, int64_t rows, int C, float eps, int use_gelu,
            float* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* xp = x + r * M;
  float v[M];
  float ss = 0.f;
#pragma unroll
  for (int m = 0; m < M; ++m) { v[m] = xp[m]; ss = fmaf(v[m], v[m], ss); }
  const float nrm = fmaxf(sqrtf(ss), eps);
  float t = nrm * scale[r % C];
  if (use_gelu) t = gelu_erf_e(t);
  float* op = out + r * M;
#pragma unroll
  for (int m = 0; m < M; ++m) op[m] = t * (v[m] / nrm);
}

}  // namespace se3

extern "C" int se3_norm_fwd(const float* x, const float* scale, int64_t rows, int C, int M, float eps, int use_gelu, float* out,
                            void* stream) {
  using namespace se3;
  SE3_REQUIRE(rows > 0 && C > 0, "se3_norm_fwd: bad sizes");
  SE3_REQUIRE(M >= 1 && M <= 11 && (M & 1), "se3_norm_fwd: M=%d must be odd and <= 11", M);
  const unsigned blocks = (unsigned)ceil_div(rows, 256);
  cudaStream_t s = as_stream(stream);
  switch (M) {
    case 1: norm_kernel<1><<<blocks, 256, 0, s>>>(x, scale, rows, C, eps, use_gelu, out); break;
    case 3: norm_kernel<3><<<blocks, 256, 0, s>>>(x, scale, rows, C, eps, use_gelu, out); break;
    case 5: norm_kernel<5><<<blocks, 256, 0, s>>>(x, scale, rows, C, eps, use_gelu, out); break;
    case 7: norm_kernel<7><<<blocks, 256, 0, s>>>(x, scale, rows, C, eps, use_gelu, out); break;
    case 9: norm_kernel<9><<<blocks, 256, 0, s>>>(x, scale, rows, C, eps, use_gelu, out); break;
    default: norm_kernel<11><<<blocks, 256, 0, s>>>(x, scale, rows, C, eps, use_gelu, out); break;
  }
  SE3_LAUNCH_OK();
  return SE3_OK;
}
