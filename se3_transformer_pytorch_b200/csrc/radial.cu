// K3: radial trunk -- the first six stages of RadialFunc.net (se3_transformer_pytorch.py:287-293) for all
// (degree_in, degree_out) pairs of one ConvSE3 in a single launch.  The last Linear (net.6) is NOT applied here:
// its output is consumed on-chip by the pairwise kernels.
//
// Output: g fp32 [pairs, E, 128] and/or the bf16 hi/lo operand image for the tcgen05 kernel:
//   per (pair, edge tile of 128): 4 sub-tiles [hi|lo][k-half] of 128 rows x 64 bf16, 128-byte swizzled, K-major
//   (the canonical UMMA SWIZZLE_128B layout), so the pairwise kernel can bulk-copy 64 KiB straight into smem.
#include "common.cuh"
#include <cuda_bf16.h>

namespace se3 {

constexpr int kMid = SE3_RADIAL_MID;  // 128
constexpr int kTrunkEB = 32;          // edges per CTA

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// byte offset of element (row r, k) inside one 128x128 (hi or lo) operand image made of two 128x64 SW128 sub-tiles
__device__ __forceinline__ uint32_t sw128_offset(int r, int k) {
  const int kh = k >> 6, kk = k & 63;
  const int chunk = (kk >> 3) ^ (r & 7);
  return (uint32_t)(kh * 16384 + r * 128 + chunk * 16 + (kk & 7) * 2);
}

// LayerNorm (eps 1e-5, biased variance) + GELU over the 128 hidden units of each edge; one warp per edge.
__device__ __forceinline__ void ln_gelu_rows(float (*h)[kMid + 4], int ne, const float* __restrict__ w,
                                             const float* __restrict__ bsh) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = warp; e < ne; e += 4) {
    float v[4];
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) { v[t] = h[e][lane + 32 * t]; s += v[t]; }
    const float mu = warp_sum(s) * (1.f / kMid);
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) { const float d = v[t] - mu; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / kMid) + 1e-5f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = lane + 32 * t;
      h[e][c] = gelu_erf((v[t] - mu) * rstd * w[c] + bsh[c]);
    }
  }
}

__global__ void __launch_bounds__(128)
radial_trunk_kernel(const float* __restrict__ feat, int64_t E, int in_dim, const float* __restrict__ params,
                    int64_t param_stride, float* __restrict__ out_g, uint8_t* __restrict__ out_img, int64_t n_mtiles) {
  __shared__ __align__(16) float h[kTrunkEB][kMid + 4];
  __shared__ float fs[kTrunkEB][64];
  const int pair = blockIdx.y;
  const int64_t e0 = (int64_t)blockIdx.x * kTrunkEB;
  const int ne = (int)max((int64_t)0, min((int64_t)kTrunkEB, E - e0));
  const int t = threadIdx.x;
  const float* P = params + (size_t)pair * param_stride;
  const float* W1T = P;                       // [in_dim][128]
  const float* b1 = W1T + (size_t)in_dim * kMid;
  const float* ln1w = b1 + kMid;
  const float* ln1b = ln1w + kMid;
  const float* W2T = ln1b + kMid;             // [128][128]  (c, t)
  const float* b2 = W2T + kMid * kMid;
  const float* ln2w = b2 + kMid;
  const float* ln2b = ln2w + kMid;

  for (int idx = t; idx < ne * in_dim; idx += 128) fs[idx / in_dim][idx % in_dim] = feat[(e0 + idx / in_dim) * in_dim + idx % in_dim];
  __syncthreads();
  // layer 1: thread t = hidden unit t
  {
    float acc[kTrunkEB];
    const float bias = b1[t];
#pragma unroll
    for (int e = 0; e < kTrunkEB; ++e) acc[e] = bias;
    for (int d = 0; d < in_dim; ++d) {
      const float w = W1T[d * kMid + t];
#pragma unroll
      for (int e = 0; e < kTrunkEB; ++e) acc[e] = fmaf(fs[e][d], w, acc[e]);
    }
#pragma unroll
    for (int e = 0; e < kTrunkEB; ++e) h[e][t] = acc[e];
  }
  __syncthreads();
  ln_gelu_rows(h, ne, ln1w, ln1b);
  __syncthreads();
  // layer 2
  {
    float acc[kTrunkEB];
    const float bias = b2[t];
#pragma unroll
    for (int e = 0; e < kTrunkEB; ++e) acc[e] = bias;
    for (int c = 0; c < kMid; c += 4) {
      const float w0 = W2T[(c + 0) * kMid + t], w1 = W2T[(c + 1) * kMid + t];
      const float w2 = W2T[(c + 2) * kMid + t], w3 = W2T[(c + 3) * kMid + t];
#pragma unroll
      for (int e = 0; e < kTrunkEB; ++e) {
        const float4 a = *reinterpret_cast<const float4*>(&h[e][c]);
        acc[e] = fmaf(a.x, w0, acc[e]);
        acc[e] = fmaf(a.y, w1, acc[e]);
        acc[e] = fmaf(a.z, w2, acc[e]);
        acc[e] = fmaf(a.w, w3, acc[e]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kTrunkEB; ++e) h[e][t] = acc[e];
  }
  __syncthreads();
  ln_gelu_rows(h, ne, ln2w, ln2b);
  __syncthreads();
  if (out_g) {
    float* og = out_g + ((size_t)pair * E + e0) * kMid;
    for (int e = 0; e < ne; ++e) og[(size_t)e * kMid + t] = h[e][t];
  }
  if (out_img) {
    // tile image: [pair][mtile] x 64 KiB = [hi: 2 x 16 KiB][lo: 2 x 16 KiB]; rows beyond E are zero.
    for (int e = 0; e < kTrunkEB; ++e) {
      const int64_t eg = e0 + e;
      if (eg >= n_mtiles * SE3_TILE_E) break;
      const float x = (e < ne) ? h[e][t] : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
      const int64_t mt = eg / SE3_TILE_E;
      const int r = (int)(eg % SE3_TILE_E);
      uint8_t* tile = out_img + ((size_t)pair * n_mtiles + mt) * 65536;
      const uint32_t off = sw128_offset(r, t);
      *reinterpret_cast<__nv_bfloat16*>(tile + off) = hi;
      *reinterpret_cast<__nv_bfloat16*>(tile + 32768 + off) = lo;
    }
  }
}

}  // namespace se3

extern "C" int se3_radial_trunk_fwd(const float* feat, int64_t E, int in_dim, int num_pairs, const float* params,
                                    float* out_g, void* out_img, void* stream) {
  using namespace se3;
  SE3_REQUIRE(E > 0 && num_pairs > 0, "se3_radial_trunk_fwd: bad sizes");
  SE3_REQUIRE(in_dim >= 1 && in_dim <= 64, "se3_radial_trunk_fwd: in_dim %d unsupported (1..64)", in_dim);
  SE3_REQUIRE(out_g || out_img, "se3_radial_trunk_fwd: no output requested");
  const int64_t n_mtiles = ceil_div(E, SE3_TILE_E);
  const int64_t param_stride = (int64_t)in_dim * kMid + 3 * kMid + kMid * kMid + 3 * kMid;
  // the image path also has to zero-fill the padded rows of the last tile: cover them with the grid
  const int64_t e_cover = out_img ? n_mtiles * SE3_TILE_E : E;
  dim3 grid((unsigned)ceil_div(e_cover, kTrunkEB), (unsigned)num_pairs);
  radial_trunk_kernel<<<grid, 128, 0, as_stream(stream)>>>(feat, E, in_dim, params, param_stride, out_g,
                                                            reinterpret_cast<uint8_t*>(out_img), n_mtiles);
  SE3_LAUNCH_OK();
  return SE3_OK;
}
