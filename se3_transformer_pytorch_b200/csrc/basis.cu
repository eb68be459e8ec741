// K2: real spherical harmonics + Clebsch-Gordan ("Q_J") basis, one pass over the edges.
//
// Reference path: get_spherical_from_cartesian (basis.py:57-95) -> precompute_sh (basis.py:140-151) ->
// lpmv / tesseral harmonics (spherical_harmonics.py:34-123, irr_repr.py:103-104) -> Y_J @ Q_J^T per pair
// (basis.py:184-198).  Here: Y_J straight from Cartesian components (no atan2/cos/pow), and all pairs'
// blocks written by one CTA per 32 edges through a CSR view of the (very sparse) Q_J tables.
#include "common.cuh"

namespace se3 {

constexpr int kMaxJ = 10;                       // 2 * (max supported degree 5)
constexpr int kMaxY = (kMaxJ + 1) * (kMaxJ + 1);  // 121
__constant__ float c_sh_norm[(kMaxJ + 1) * (kMaxJ + 1)];   // [l][m], m >= 0
__constant__ float c_pmm_coef[kMaxJ + 1];                  // (-1)^m (2m-1)!!

// Y[J*J + (m + J)] for J = 0..maxJ.  Angle conventions of the reference:
//   cartesian (x,y,z) := (c[2], c[0], c[1])  (basis.py:76);  beta = atan2(sqrt(x^2+y^2), z), alpha = atan2(y, x);
//   theta = pi - beta, phi = alpha (irr_repr.py:104)  =>  cos(theta) = -c1/r, sin(theta) = rxy/r,
//   cos(phi) = c2/rxy, sin(phi) = c0/rxy, with atan2(0,0) = 0 at the poles / origin.
__device__ void real_sh(float c0, float c1, float c2, int maxJ, float* __restrict__ Y, int ystride) {
  const float rxy2 = c2 * c2 + c0 * c0;
  const float rxy = sqrtf(rxy2);
  const float r = sqrtf(rxy2 + c1 * c1);
  float ct = -1.f, st = 0.f, cp = 1.f, sp = 0.f;
  if (r > 0.f) { ct = -c1 / r; st = rxy / r; }
  if (rxy > 0.f) { cp = c2 / rxy; sp = c0 / rxy; }
  float stm = 1.f;            // sin(theta)^m
  float cm = 1.f, sm = 0.f;   // cos(m phi), sin(m phi)
  for (int m = 0; m <= maxJ; ++m) {
    float p_lm2 = 0.f;
    float p_lm1 = c_pmm_coef[m] * stm;  // P_m^m
    for (int l = m; l <= maxJ; ++l) {
      float p;
      if (l == m) {
        p = p_lm1;
      } else {
        p = ((float)(2 * l - 1) / (float)(l - m)) * ct * p_lm1;
        if (l - m > 1) p -= ((float)(l + m - 1) / (float)(l - m)) * p_lm2;
        p_lm2 = p_lm1;
        p_lm1 = p;
      }
      const float nrm = c_sh_norm[l * (kMaxJ + 1) + m];
      float* Yl = Y + (size_t)(l * l + l) * ystride;
      if (m == 0) {
        Yl[0] = nrm * p;
      } else {
        Yl[(size_t)m * ystride] = cm * p * nrm;
        Yl[-(ptrdiff_t)m * ystride] = sm * p * nrm;
      }
    }
    stm *= st;
    const float cn = cm * cp - sm * sp;
    const float sn = sm * cp + cm * sp;
    cm = cn; sm = sn;
  }
}

constexpr int kEB = 32;   // edges per CTA

__global__ void __launch_bounds__(128)
basis_kernel(const float* __restrict__ rel_pos, int64_t E, int maxJ, int ny,
             const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const float* __restrict__ val,
             const int32_t* __restrict__ pair_row0, const int32_t* __restrict__ pair_base, int num_pairs,
             float* __restrict__ out) {
  __shared__ float Ys[kMaxY * (kEB + 1)];   // [y][edge], padded
  const int64_t e0 = (int64_t)blockIdx.x * kEB;
  const int ne = (int)min((int64_t)kEB, E - e0);
  if (threadIdx.x < kEB) {
    const int el = threadIdx.x;
    if (el < ne) {
      const float* c = rel_pos + (e0 + el) * 3;
      real_sh(c[0], c[1], c[2], maxJ, Ys + el, kEB + 1);
    }
  }
  __syncthreads();
  for (int p = 0; p < num_pairs; ++p) {
    const int r0 = pair_row0[p];
    const int S = pair_row0[p + 1] - r0;
    float* o = out + (size_t)pair_base[p] * E + (size_t)e0 * S;
    const int total = ne * S;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
      const int el = t / S;
      const int rl = t - el * S;
      const int r = r0 + rl;
      float acc = 0.f;
      for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) acc = fmaf(val[q], Ys[col[q] * (kEB + 1) + el], acc);
      o[t] = acc;
    }
  }
}

static bool g_tables_ready[16] = {};

static int upload_tables() {
  int dev = 0;
  SE3_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 16 && g_tables_ready[dev]) return SE3_OK;
  float norm[(kMaxJ + 1) * (kMaxJ + 1)] = {};
  float pmm[kMaxJ + 1];
  const double pi = 3.14159265358979323846;
  for (int l = 0; l <= kMaxJ; ++l) {
    for (int m = 0; m <= l; ++m) {
      double nn = sqrt((2.0 * l + 1.0) / (4.0 * pi));
      if (m > 0) {
        double poch = 1.0;   // (l-m+1)(l-m+2)...(l+m)   (spherical_harmonics.py:25-27, 104)
        for (int t = l - m + 1; t <= l + m; ++t) poch *= t;
        nn *= sqrt(2.0 / poch);
      }
      norm[l * (kMaxJ + 1) + m] = (float)nn;
    }
  }
  for (int m = 0; m <= kMaxJ; ++m) {
    double semi = 1.0;
    for (int t = 2 * m - 1; t > 1; t -= 2) semi *= t;
    pmm[m] = (float)((m & 1) ? -semi : semi);
  }
  SE3_CUDA_OK(cudaMemcpyToSymbol(c_sh_norm, norm, sizeof(norm)));
  SE3_CUDA_OK(cudaMemcpyToSymbol(c_pmm_coef, pmm, sizeof(pmm)));
  if (dev < 16) g_tables_ready[dev] = true;
  return SE3_OK;
}

}  // namespace se3

extern "C" int se3_basis_fwd(const float* rel_pos, int64_t E, int max_degree, const int32_t* csr_row_ptr,
                             const int32_t* csr_col, const float* csr_val, const int32_t* pair_row0,
                             const int32_t* pair_base, int num_pairs, float* out, void* stream) {
  using namespace se3;
  SE3_REQUIRE(E > 0, "se3_basis_fwd: E must be positive");
  SE3_REQUIRE(max_degree >= 0 && 2 * max_degree <= kMaxJ, "se3_basis_fwd: max_degree %d unsupported (max %d)", max_degree,
              kMaxJ / 2);
  SE3_REQUIRE(num_pairs == (max_degree + 1) * (max_degree + 1), "se3_basis_fwd: num_pairs must be (max_degree+1)^2");
  int rc = upload_tables();
  if (rc != SE3_OK) return rc;
  const int maxJ = 2 * max_degree;
  const int ny = (maxJ + 1) * (maxJ + 1);
  basis_kernel<<<(unsigned)ceil_div(E, kEB), 128, 0, as_stream(stream)>>>(rel_pos, E, maxJ, ny, csr_row_ptr, csr_col, csr_val,
                                                                          pair_row0, pair_base, num_pairs, out);
  SE3_LAUNCH_OK();
  return SE3_OK;
}
