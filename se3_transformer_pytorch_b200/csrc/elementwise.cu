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

// Basis fold of the input-side contraction.  For a degree pair with 2 l_in + 1 = Q < P = 2 l_out + 1 the
// reference's per-edge product  out[e,o,p] = sum_{i,f} R[e,o,i,f] * (sum_q B[e,p,q,f] x[j(e),i,q])  (S:336-343, 251) is
// evaluated as  S[f,e,o,q] = sum_i R[e,o,i,f] x[j(e),i,q]  by the fused pairwise kernel (Q instead of P FMAs per R
// element) followed by this fold:  out[e,o,p] (+)= sum_{f,q} B[e,p,q,f] S[f,e,o,q].
// One thread per (edge, channel); the P*Q*F basis values of the block's edges are staged in shared memory.
constexpr int kFoldThreads = 256;
constexpr int kFoldMaxB = 7 * 7 * 7;

template <bool CM>      // CM: S is [F,E,Q,Co] (component-major) instead of [F,E,Co,Q]
__global__ void __launch_bounds__(kFoldThreads)
fold_basis_kernel(const float* __restrict__ S, const float* __restrict__ basis, int64_t E, int Co, int P, int Q, int F,
                  int accumulate, float* __restrict__ out) {
  extern __shared__ float sB[];                       // [edges of this block][P*Q*F]
  const int64_t first = (int64_t)blockIdx.x * kFoldThreads;
  const int64_t e_first = first / Co;
  const int64_t last = min(first + kFoldThreads, E * Co) - 1;
  const int n_edges = (int)(last / Co - e_first) + 1;
  const int nb = P * Q * F;
  for (int t = threadIdx.x; t < n_edges * nb; t += kFoldThreads) sB[t] = basis[e_first * nb + t];
  __syncthreads();
  const int64_t idx = first + threadIdx.x;
  if (idx >= E * Co) return;
  const int64_t e = idx / Co;
  const float* b = sB + (e - e_first) * nb;           // [p][q][f]
  float acc[7];
#pragma unroll
  for (int p = 0; p < 7; ++p) acc[p] = 0.f;
  for (int f = 0; f < F; ++f) {
    const float* sp = CM ? S + ((size_t)f * E + e) * Q * Co + (idx - e * Co) : S + ((size_t)f * E * Co + idx) * Q;
    for (int q = 0; q < Q; ++q) {
      const float sv = CM ? sp[(size_t)q * Co] : sp[q];
#pragma unroll
      for (int p = 0; p < 7; ++p)
        if (p < P) acc[p] = fmaf(b[(p * Q + q) * F + f], sv, acc[p]);
    }
  }
  float* op = out + idx * P;
#pragma unroll
  for (int p = 0; p < 7; ++p)
    if (p < P) op[p] = accumulate ? op[p] + acc[p] : acc[p];
}

// Rotation of the edge-aligned outputs back to the global frame: out[e,o,:] = D_lo(e) out'[e,o,:], where
// out' arrives as one dense buffer per |m|: part0 [E,Co] (m = 0), part_m [E,Co,2] = components (+m, -m).
__global__ void __launch_bounds__(kFoldThreads)
rotate_back_kernel(const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                   const float* __restrict__ p3, const float* __restrict__ D, int64_t E, int Co, int lo, float* __restrict__ out) {
  extern __shared__ float sB[];                       // [edges of this block][P*P]
  const int P = 2 * lo + 1;
  const int64_t first = (int64_t)blockIdx.x * kFoldThreads;
  const int64_t e_first = first / Co;
  const int64_t last = min(first + kFoldThreads, E * Co) - 1;
  const int n_edges = (int)(last / Co - e_first) + 1;
  const int nb = P * P;
  for (int t = threadIdx.x; t < n_edges * nb; t += kFoldThreads) sB[t] = D[e_first * nb + t];
  __syncthreads();
  const int64_t idx = first + threadIdx.x;
  if (idx >= E * Co) return;
  const float* d = sB + (idx / Co - e_first) * nb;    // [p][n]
  float v[7];
#pragma unroll
  for (int n = 0; n < 7; ++n) v[n] = 0.f;
  v[lo] = p0 ? p0[idx] : 0.f;
  const float* parts[3] = {p1, p2, p3};
#pragma unroll
  for (int m = 1; m <= 3; ++m) {
    if (m <= lo && parts[m - 1] != nullptr) {
      const float2 pm = reinterpret_cast<const float2*>(parts[m - 1])[idx];
#pragma unroll
      for (int n = 0; n < 7; ++n) {
        if (n == lo + m) v[n] = pm.x;
        if (n == lo - m) v[n] = pm.y;
      }
    }
  }
  float* op = out + idx * P;
#pragma unroll
  for (int pp = 0; pp < 7; ++pp) {
    if (pp < P) {
      float acc = 0.f;
#pragma unroll
      for (int n = 0; n < 7; ++n)
        if (n < P) acc = fmaf(d[pp * P + n], v[n], acc);
      op[pp] = acc;
    }
  }
}

}  // namespace se3

extern "C" int se3_rotate_back_fwd(const float* part0, const float* part1, const float* part2, const float* part3, const float* D,
                                   int64_t E, int Co, int lo, float* out, void* stream) {
  using namespace se3;
  SE3_REQUIRE(E > 0 && Co > 0 && lo >= 1 && lo <= 3, "se3_rotate_back_fwd: bad sizes (degree_out 1..3)");
  const int P = 2 * lo + 1;
  const unsigned blocks = (unsigned)ceil_div(E * Co, (int64_t)kFoldThreads);
  const size_t smem = (size_t)(kFoldThreads / Co + 2) * P * P * sizeof(float);
  SE3_REQUIRE(smem <= 48 * 1024, "se3_rotate_back_fwd: Co=%d too small for the staging buffer", Co);
  rotate_back_kernel<<<blocks, kFoldThreads, smem, as_stream(stream)>>>(part0, part1, part2, part3, D, E, Co, lo, out);
  SE3_LAUNCH_OK();
  return SE3_OK;
}

static int fold_basis_impl(const float* S, const float* basis_pair, int64_t E, int Co, int P, int Q, int F, int accumulate,
                           float* out, bool cm, void* stream) {
  using namespace se3;
  SE3_REQUIRE(E > 0 && Co > 0, "se3_fold_basis_fwd: bad sizes");
  SE3_REQUIRE(P >= 1 && P <= 7 && Q >= 1 && Q <= 7 && F >= 1 && F <= 7, "se3_fold_basis_fwd: P=%d Q=%d F=%d unsupported (degrees <= 3)", P, Q, F);
  const int64_t total = E * Co;
  const unsigned blocks = (unsigned)ceil_div(total, (int64_t)kFoldThreads);
  const int max_edges = kFoldThreads / Co + 2;
  const size_t smem = (size_t)max_edges * P * Q * F * sizeof(float);
  SE3_REQUIRE(smem <= 48 * 1024, "se3_fold_basis_fwd: Co=%d too small for the staging buffer", Co);
  if (cm) fold_basis_kernel<true><<<blocks, kFoldThreads, smem, as_stream(stream)>>>(S, basis_pair, E, Co, P, Q, F, accumulate, out);
  else fold_basis_kernel<false><<<blocks, kFoldThreads, smem, as_stream(stream)>>>(S, basis_pair, E, Co, P, Q, F, accumulate, out);
  SE3_LAUNCH_OK();
  return SE3_OK;
}

extern "C" int se3_fold_basis_fwd(const float* S, const float* basis_pair, int64_t E, int Co, int P, int Q, int F, int accumulate,
                                  float* out, void* stream) {
  return fold_basis_impl(S, basis_pair, E, Co, P, Q, F, accumulate, out, false, stream);
}

extern "C" int se3_fold_basis_cm_fwd(const float* S, const float* basis_pair, int64_t E, int Co, int P, int Q, int F, int accumulate,
                                     float* out, void* stream) {
  return fold_basis_impl(S, basis_pair, E, Co, P, Q, F, accumulate, out, true, stream);
}

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
