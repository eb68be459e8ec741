// Geometry and data movement of the edge-aligned formulation, all per forward, all on the device:
//   frames_kernel      rotation R_e taking the polar axis a = (0,1,0) of the reference's harmonics (B:57-95) to the edge
//                      direction, and the real Wigner matrices D_l(R_e), l = 1..lmax, in float64
//   rotgather_kernel   x'[e,i,:] = D_li(R_e)^T x[j(e),i,:]: the neighbour gather of ConvSE3 (S:237-238, U:56-70) fused with the
//                      rotation into the edge frame, written in the layout the Z generators of zgemm.cu read
//   rowabsmax / edge_scale   power-of-two scale per edge that keeps the fp16 operands of zgemm in range
#include "common.cuh"
#include <cmath>

namespace se3 {

constexpr int kMaxL = 5;

struct FrameTables {
  const double* xs[kMaxL + 1];     // [S_l, 3] unit sample points of degree l
  const double* pin[kMaxL + 1];    // [2l+1, S_l] pseudo-inverse of Y_l(xs)
  int S[kMaxL + 1];
  float* D[kMaxL + 1];             // out: [E, 2l+1, 2l+1] fp32
};

// Real spherical harmonics of degree l of the reference (SH:34-123 with theta = pi - beta, phi = alpha, IR:103-104, axis
// permutation (x,y,z) = (c2,c0,c1) of B:57-95) for a direction d (any length > 0), float64, m = -l..l.
__device__ void real_sh_l(const double d[3], int l, double* Y) {
  const double cx = d[2], cy = d[0], cz = d[1];
  const double rxy = cx * cx + cy * cy;
  const double nrm = sqrt(rxy + cz * cz);
  const double ct = -cz / nrm;                       // cos(pi - beta), beta = atan2(sqrt(rxy), cz)
  const double st = sqrt(fmax(1.0 - ct * ct, 0.0));
  double cphi = 1.0, sphi = 0.0;                     // phi = atan2(cy, cx); atan2(0, 0) = 0
  if (rxy > 0.0) { const double r = sqrt(rxy); cphi = cx / r; sphi = cy / r; }
  double P[kMaxL + 1];                               // P_l^m for m = 0..l
  for (int m = 0; m <= l; ++m) {
    double pmm = 1.0;                                // (-1)^m (2m-1)!! st^m
    for (int k = 1; k <= m; ++k) pmm *= -(2.0 * k - 1.0) * st;
    double p_prev = 0.0, p_cur = pmm;                // P_{m-1}^m = 0, P_m^m
    for (int ll = m + 1; ll <= l; ++ll) {
      const double p_next = ((2.0 * ll - 1.0) * ct * p_cur - (ll + m - 1.0) * p_prev) / (ll - m);
      p_prev = p_cur;
      p_cur = p_next;
    }
    P[m] = p_cur;
  }
  const double n0 = sqrt((2.0 * l + 1.0) / (4.0 * 3.14159265358979323846));
  Y[l] = n0 * P[0];
  double cm = 1.0, sm = 0.0;
  for (int m = 1; m <= l; ++m) {
    const double c2 = cm * cphi - sm * sphi, s2 = sm * cphi + cm * sphi;
    cm = c2; sm = s2;
    double poch = 1.0;
    for (int n = l - m + 1; n <= l + m; ++n) poch *= n;
    const double nm = n0 * sqrt(2.0 / poch);
    Y[l + m] = nm * cm * P[m];
    Y[l - m] = nm * sm * P[m];
  }
}

__global__ void __launch_bounds__(128)
frames_kernel(const float* __restrict__ rel_pos, int64_t E, int lmax, FrameTables tb) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double d[3] = {(double)rel_pos[e * 3 + 0], (double)rel_pos[e * 3 + 1], (double)rel_pos[e * 3 + 2]};
  const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  // coincident points: the reference evaluates its harmonics at beta = atan2(0,0) = 0, i.e. on the axis: identity frame
  double rh[3] = {0.0, 1.0, 0.0};
  if (nrm > 0.0) { rh[0] = d[0] / nrm; rh[1] = d[1] / nrm; rh[2] = d[2] / nrm; }
  // Rodrigues from src = +-a (the hemisphere of rhat) to rhat; directions in the lower hemisphere start from -a after a
  // half turn about x
  const bool flip = rh[1] < 0.0;
  const double sy = flip ? -1.0 : 1.0;               // src = (0, sy, 0)
  const double v[3] = {sy * rh[2], 0.0, -sy * rh[0]};   // src x rhat
  const double c = sy * rh[1];
  const double vx[3][3] = {{0.0, -v[2], v[1]}, {v[2], 0.0, -v[0]}, {-v[1], v[0], 0.0}};
  double R[3][3];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double v2 = 0.0;
      for (int k = 0; k < 3; ++k) v2 += vx[a][k] * vx[k][b];
      R[a][b] = (a == b ? 1.0 : 0.0) + vx[a][b] + v2 / (1.0 + c);
    }
  if (flip)
    for (int a = 0; a < 3; ++a) { R[a][1] = -R[a][1]; R[a][2] = -R[a][2]; }
  for (int l = 1; l <= lmax; ++l) {
    const int M = 2 * l + 1, S = tb.S[l];
    double Dm[(2 * kMaxL + 1) * (2 * kMaxL + 1)];
    for (int i = 0; i < M * M; ++i) Dm[i] = 0.0;
    for (int s = 0; s < S; ++s) {
      const double* x = tb.xs[l] + 3 * s;
      double pt[3];
      for (int a = 0; a < 3; ++a) pt[a] = R[a][0] * x[0] + R[a][1] * x[1] + R[a][2] * x[2];
      double Y[2 * kMaxL + 1];
      real_sh_l(pt, l, Y);
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < M; ++n) Dm[m * M + n] += Y[m] * tb.pin[l][n * S + s];   // Y_l(R x_s) = D_l Y_l(x_s)
    }
    float* out = tb.D[l] + (size_t)e * M * M;
    for (int i = 0; i < M * M; ++i) out[i] = (float)Dm[i];
  }
}

// x'[e,i,n] = sum_q D[e][q][n] x[b(e), idx[e], i, q]   ->   X[tile][i][n][edge_local]  (128 edges per tile)
//
// CTA = one edge tile x a slab of 8 input channels, 256 threads.  Gather phase: thread = (edge, channel) reads its Q
// contiguous floats of the neighbour's feature row (a warp covers 4 edges x 8 channels: 8*Q*4 contiguous bytes per edge),
// rotates them with the edge's Wigner block (staged in shared memory, read as a broadcast) and parks the Q results in a
// padded shared-memory tile [channel][component][edge]; store phase: the tile goes out as whole 512-byte rows
// X[tile][i][n][0..127].  (Round-2 first version: thread = edge with a channel loop and 4-byte stores per (channel, component):
// 1.4 TB/s.)
constexpr int kRgCh = 8;            // channels per CTA
constexpr int kRgPad = 132;         // padded row of the transposing tile (conflict-free for the (edge, channel) -> [channel][n][edge] write)

template <int Q>
__global__ void __launch_bounds__(256)
rotgather_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ D, int64_t E,
                 int64_t mt_begin, int n, int k, int Ci, float* __restrict__ X) {
  extern __shared__ float smem_rg[];
  float* Ds = smem_rg;                                   // [128 edges][Q*Q] (only Q > 1)
  float* tile = smem_rg + (Q > 1 ? SE3_TILE_E * Q * Q : 0);   // [kRgCh][Q][kRgPad]
  const int64_t mt = blockIdx.x;
  const int i0 = blockIdx.y * kRgCh;
  const int64_t e0 = (mt_begin + mt) * SE3_TILE_E;
  const int nvalid = (int)max((int64_t)0, min((int64_t)SE3_TILE_E, E - e0));
  if (Q > 1) {
    for (int t = threadIdx.x; t < nvalid * Q * Q; t += 256) Ds[t] = D[(size_t)e0 * Q * Q + t];      // contiguous block of the tile's edges
    __syncthreads();
  }
  const int ch = threadIdx.x & (kRgCh - 1);              // channel inside the slab
  const int i = i0 + ch;
#pragma unroll 2
  for (int el = threadIdx.x >> 3; el < SE3_TILE_E; el += 32) {
    float o[Q];
#pragma unroll
    for (int nn = 0; nn < Q; ++nn) o[nn] = 0.f;
    if (el < nvalid && i < Ci) {
      const int64_t e = e0 + el;
      const int64_t bb = (e / k) / n;
      const float* xr = x + (((size_t)(bb * n + idx[e]) * Ci) + i) * Q;
      float xv[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) xv[q] = __ldg(xr + q);
      if (Q == 1) {
        o[0] = xv[0];
      } else {
        const float* d = Ds + el * Q * Q;
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
          for (int nn = 0; nn < Q; ++nn) o[nn] = fmaf(d[q * Q + nn], xv[q], o[nn]);
      }
    }
#pragma unroll
    for (int nn = 0; nn < Q; ++nn) tile[(ch * Q + nn) * kRgPad + el] = o[nn];
  }
  __syncthreads();
  // rows (channel, component) of 128 edges each: float4 stores, one warp per row at a time
  float* Xt = X + ((size_t)mt * Ci * Q) * SE3_TILE_E;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRgCh * Q; r += 8) {
    const int c = r / Q;
    if (i0 + c >= Ci) continue;
    const float4 v = *reinterpret_cast<const float4*>(tile + r * kRgPad + lane * 4);
    *reinterpret_cast<float4*>(Xt + ((size_t)(i0 + c) * Q + (r - c * Q)) * SE3_TILE_E + lane * 4) = v;
  }
}

// out[row] = max |x[row, :]|  (combine: max with the value already there); one warp per row
__global__ void __launch_bounds__(256)
rowabsmax_kernel(const float* __restrict__ x, int64_t rows, int W, int combine, float* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = x + row * W;
  float m = 0.f;
  for (int c = lane; c < W; c += 32) m = fmaxf(m, fabsf(p[c]));
  m = warp_max(m);
  if (lane == 0) out[row] = combine ? fmaxf(out[row], m) : m;
}

// sx[e] = 2^j with  nodemax[b(e), idx[e]] * sqrt(2 lmax + 1) * sx < 2^10  (1 if the node's features are all zero or not finite)
__global__ void __launch_bounds__(256)
edge_scale_kernel(const float* __restrict__ nodemax, const int64_t* __restrict__ idx, int64_t E, int n, int k, float comp_bound,
                  float* __restrict__ sx) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t bb = (e / k) / n;
  const float v = nodemax[bb * n + idx[e]] * comp_bound;
  float s = 1.f;
  if (v > 0.f && v < 3.0e38f) {
    int ex;
    frexpf(v, &ex);                                    // v = f * 2^ex, f in [0.5, 1)
    s = ldexpf(1.f, max(-100, min(100, 10 - ex)));
  }
  sx[e] = s;
}

// Pooled ConvSE3 epilogue (conv_in / conv_out; reference S:256-266, utils.py:72-80) fused with the rotation back to the global
// frame: out[node,o,:] = masked_mean_j ( D_lo(e_j) out'[e_j,:,o] ) (+ self[node,o,:]), e_j = node*K + j.  The edge-level
// [E, C, 2lo+1] tensor of the unfused path is never written.  Thread = (node, channel); D of 32 edges at a time in shared memory.
constexpr int kRpEdges = 32;

template <int P>
__global__ void __launch_bounds__(256)
rotate_pool_kernel(const float* __restrict__ Op, const float* __restrict__ D, const uint8_t* __restrict__ mask,
                   const float* __restrict__ self_add, int K, int Co, float* __restrict__ out) {
  __shared__ float sD[kRpEdges * P * P];
  __shared__ uint8_t sM[kRpEdges];
  const int64_t node = blockIdx.x;
  const int o = blockIdx.y * blockDim.x + threadIdx.x;
  float acc[P];
#pragma unroll
  for (int p = 0; p < P; ++p) acc[p] = 0.f;
  int cnt = 0;
  for (int j0 = 0; j0 < K; j0 += kRpEdges) {
    const int nj = min(kRpEdges, K - j0);
    const int64_t e0 = node * K + j0;
    __syncthreads();
    if (P > 1)
      for (int t = threadIdx.x; t < nj * P * P; t += blockDim.x) sD[t] = D[e0 * P * P + t];
    for (int t = threadIdx.x; t < nj; t += blockDim.x) sM[t] = mask ? mask[e0 + t] : (uint8_t)1;
    __syncthreads();
    for (int j = 0; j < nj; ++j) {
      if (!sM[j]) continue;                       // uniform over the block
      ++cnt;
      if (o < Co) {
        const float* src = Op + ((size_t)(e0 + j) * P) * Co + o;
        float v[P];
#pragma unroll
        for (int n = 0; n < P; ++n) v[n] = src[(size_t)n * Co];
        if (P == 1) {
          acc[0] += v[0];
        } else {
          const float* d = sD + j * P * P;
#pragma unroll
          for (int p = 0; p < P; ++p) {
            float a = acc[p];
#pragma unroll
            for (int n = 0; n < P; ++n) a = fmaf(d[p * P + n], v[n], a);
            acc[p] = a;
          }
        }
      }
    }
  }
  if (o >= Co) return;
  const float inv = 1.f / (float)max(cnt, 1);     // masked_mean: sum / clamp(count, 1), zero when nothing is left (utils.py:76-79)
  float* dst = out + ((size_t)node * Co + o) * P;
  const float* sa = self_add ? self_add + ((size_t)node * Co + o) * P : nullptr;
#pragma unroll
  for (int p = 0; p < P; ++p) dst[p] = (cnt ? acc[p] * inv : 0.f) + (sa ? sa[p] : 0.f);
}

// sx[node] = 2^j with rowmax[node] * sx < 2^target_exp (1 for all-zero / non-finite rows)
__global__ void __launch_bounds__(256)
pow2_scale_kernel(const float* __restrict__ rowmax, int64_t rows, int target_exp, float* __restrict__ sx) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float v = rowmax[r];
  float s = 1.f;
  if (v > 0.f && v < 3.0e38f) {
    int ex;
    frexpf(v, &ex);
    s = ldexpf(1.f, max(-100, min(100, target_exp - ex)));
  }
  sx[r] = s;
}

template <int Q>
static void launch_rg(dim3 grid, cudaStream_t s, const float* x, const int64_t* idx, const float* D, int64_t E, int64_t tb, int n, int k,
                      int Ci, float* X) {
  const size_t smem = sizeof(float) * ((Q > 1 ? (size_t)SE3_TILE_E * Q * Q : 0) + (size_t)kRgCh * Q * kRgPad);
  cudaFuncSetAttribute(rotgather_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rotgather_kernel<Q><<<grid, 256, smem, s>>>(x, idx, D, E, tb, n, k, Ci, X);
}

}  // namespace se3

extern "C" int se3_frames_fwd(const float* rel_pos, int64_t E, int lmax, const double* const* xs, const double* const* pin,
                              const int* n_samples, float* const* D_out, void* stream) {
  using namespace se3;
  SE3_REQUIRE(E > 0 && lmax >= 1 && lmax <= kMaxL, "se3_frames_fwd: lmax=%d out of range (1..%d)", lmax, kMaxL);
  FrameTables tb = {};
  for (int l = 1; l <= lmax; ++l) {
    SE3_REQUIRE(xs[l] != nullptr && pin[l] != nullptr && D_out[l] != nullptr && n_samples[l] >= 2 * l + 1, "se3_frames_fwd: bad tables for degree %d", l);
    tb.xs[l] = xs[l];
    tb.pin[l] = pin[l];
    tb.S[l] = n_samples[l];
    tb.D[l] = D_out[l];
  }
  frames_kernel<<<(unsigned)ceil_div(E, 128), 128, 0, as_stream(stream)>>>(rel_pos, E, lmax, tb);
  SE3_LAUNCH_OK();
  return SE3_OK;
}

extern "C" int se3_rotgather_fwd(const float* x, const int64_t* idx, const float* D, int b, int n, int k, int Ci, int Q,
                                 int64_t tile_begin, int64_t tile_count, float* X, void* stream) {
  using namespace se3;
  SE3_REQUIRE(b > 0 && n > 0 && k > 0 && Ci > 0, "se3_rotgather_fwd: bad sizes");
  SE3_REQUIRE(Q >= 1 && Q <= 11 && (Q & 1), "se3_rotgather_fwd: Q=%d must be odd and <= 11", Q);
  SE3_REQUIRE(Q == 1 || D != nullptr, "se3_rotgather_fwd: D is required for degree > 0");
  const int64_t E = (int64_t)b * n * k;
  const int64_t n_all = ceil_div(E, SE3_TILE_E);
  SE3_REQUIRE(tile_begin >= 0 && tile_count > 0 && tile_begin + tile_count <= n_all, "se3_rotgather_fwd: tile range out of bounds");
  dim3 grid((unsigned)tile_count, (unsigned)ceil_div(Ci, kRgCh));
  cudaStream_t st = as_stream(stream);
  switch (Q) {
    case 1: launch_rg<1>(grid, st, x, idx, D, E, tile_begin, n, k, Ci, X); break;
    case 3: launch_rg<3>(grid, st, x, idx, D, E, tile_begin, n, k, Ci, X); break;
    case 5: launch_rg<5>(grid, st, x, idx, D, E, tile_begin, n, k, Ci, X); break;
    case 7: launch_rg<7>(grid, st, x, idx, D, E, tile_begin, n, k, Ci, X); break;
    case 9: launch_rg<9>(grid, st, x, idx, D, E, tile_begin, n, k, Ci, X); break;
    default: launch_rg<11>(grid, st, x, idx, D, E, tile_begin, n, k, Ci, X); break;
  }
  SE3_LAUNCH_OK();
  return SE3_OK;
}

extern "C" int se3_rotate_pool_fwd(const float* Oprime, const float* D, const uint8_t* mask, const float* self_add, int64_t nodes, int K,
                                   int Co, int lo, float* out, void* stream) {
  using namespace se3;
  SE3_REQUIRE(nodes > 0 && K > 0 && Co > 0 && lo >= 0 && lo <= kMaxL, "se3_rotate_pool_fwd: bad sizes (degree_out 0..%d)", kMaxL);
  SE3_REQUIRE(lo == 0 || D != nullptr, "se3_rotate_pool_fwd: D is required for degree > 0");
  SE3_REQUIRE(nodes <= 2147483647ll && Co <= 65535 * 256, "se3_rotate_pool_fwd: too many nodes / channels");
  dim3 grid((unsigned)nodes, (unsigned)ceil_div(Co, 256));
  cudaStream_t s = as_stream(stream);
  switch (lo) {
    case 0: rotate_pool_kernel<1><<<grid, 256, 0, s>>>(Oprime, D, mask, self_add, K, Co, out); break;
    case 1: rotate_pool_kernel<3><<<grid, 256, 0, s>>>(Oprime, D, mask, self_add, K, Co, out); break;
    case 2: rotate_pool_kernel<5><<<grid, 256, 0, s>>>(Oprime, D, mask, self_add, K, Co, out); break;
    case 3: rotate_pool_kernel<7><<<grid, 256, 0, s>>>(Oprime, D, mask, self_add, K, Co, out); break;
    case 4: rotate_pool_kernel<9><<<grid, 256, 0, s>>>(Oprime, D, mask, self_add, K, Co, out); break;
    default: rotate_pool_kernel<11><<<grid, 256, 0, s>>>(Oprime, D, mask, self_add, K, Co, out); break;
  }
  SE3_LAUNCH_OK();
  return SE3_OK;
}

extern "C" int se3_rowabsmax_fwd(const float* x, int64_t rows, int W, int combine, float* out, void* stream) {
  using namespace se3;
  SE3_REQUIRE(rows > 0 && W > 0, "se3_rowabsmax_fwd: bad sizes");
  rowabsmax_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(x, rows, W, combine, out);
  SE3_LAUNCH_OK();
  return SE3_OK;
}

extern "C" int se3_pow2_scale_fwd(const float* rowmax, int64_t rows, int target_exp, float* sx, void* stream) {
  using namespace se3;
  SE3_REQUIRE(rows > 0 && target_exp >= -60 && target_exp <= 60, "se3_pow2_scale_fwd: bad sizes");
  pow2_scale_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, as_stream(stream)>>>(rowmax, rows, target_exp, sx);
  SE3_LAUNCH_OK();
  return SE3_OK;
}

extern "C" int se3_edge_scale_fwd(const float* nodemax, const int64_t* idx, int b, int n, int k, int max_degree, float* sx, void* stream) {
  using namespace se3;
  SE3_REQUIRE(b > 0 && n > 0 && k > 0 && max_degree >= 0, "se3_edge_scale_fwd: bad sizes");
  const int64_t E = (int64_t)b * n * k;
  edge_scale_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, as_stream(stream)>>>(nodemax, idx, E, n, k, sqrtf(2.f * max_degree + 1.f), sx);
  SE3_LAUNCH_OK();
  return SE3_OK;
}
