// Micro-test of the 2-CTA tensor-core form used by zgemm's pair mode: tcgen05.mma.cta_group::2.kind::f16, A from tensor memory
// (.ts), B split over the shared memory of the two CTAs of a cluster, one commit multicast to both.  Checks D = A B^T exactly
// (small integers) and prints which B-half convention the hardware uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I se3_transformer_pytorch_b200/csrc -o tools/ubench/mma2sm tools/ubench/mma2sm.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace se3;

constexpr int N = 128;            // accumulator columns (full N of the pair instruction)
constexpr int NH = N / 2;         // B rows held by each CTA

__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(float* out, int swap_halves) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const uint32_t sB = base;                       // NH rows x 128 B (SW128), only K chunk 0 used
  const uint32_t bar = base + 16384, slot = bar + 16;
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // B half of this CTA: logical row n = half*NH + r, value B[n][k] = ((n % 7) - 3) for k == n % 16 else (k == 0 ? 1 : 0)
  const int half = swap_halves ? 1 - (int)rank : (int)rank;
  for (int t = threadIdx.x; t < NH * 64; t += blockDim.x) {
    const int r = t >> 6, kk = t & 63;
    const int n = half * NH + r;
    float v = 0.f;
    if (kk < 16) v = (kk == n % 16) ? (float)((n % 7) - 3) : (kk == 0 ? 1.f : 0.f);
    *reinterpret_cast<__half*>(bp + (sB - base) + r * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2) = __float2half(v);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(bp + (slot - base));
  // A rows: global row g = rank*128 + (warp*32 + lane): A[g][k] = (g % 5) + k  (fp16 exact), columns 256..263 of TMEM
  {
    const int g = rank * 128 + warp * 32 + lane;
    uint32_t r[16];
    for (int c = 0; c < 8; ++c) {
      const __half2 h = __floats2half2_rn((float)((g % 5) + 2 * c), (float)((g % 5) + 2 * c + 1));
      r[c] = *reinterpret_cast<const uint32_t*>(&h);
      r[8 + c] = 0u;
    }
    tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256u, r);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (rank == 0 && warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);     // fp16 x fp16 -> fp32, M = 256 (pair), N
      mma2_ts(tmem, tmem + 256u, umma_desc_sw128(sB), idesc, 0u);
      commit2(bar, (uint16_t)3);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  {
    const int g = rank * 128 + warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[(size_t)g * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  float* d;
  cudaMalloc(&d, 256 * N * 4);
  for (int swap = 0; swap < 2; ++swap) {
    cudaMemset(d, 0xff, 256 * N * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
    k<<<2, 128, 20000>>>(d, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("swap=%d CUDA error: %s\n", swap, cudaGetErrorString(e)); return 1; }
    std::vector<float> h(256 * N);
    cudaMemcpy(h.data(), d, 256 * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int g = 0; g < 256; ++g)
      for (int n = 0; n < N; ++n) {
        float ref = 0.f;
        for (int kk = 0; kk < 16; ++kk) {
          const float b = (kk == n % 16) ? (float)((n % 7) - 3) : (kk == 0 ? 1.f : 0.f);
          ref += (float)((g % 5) + kk) * b;
        }
        if (h[(size_t)g * N + n] != ref) { if (bad < 4) printf("  swap=%d mismatch row %d col %d: got %g want %g\n", swap, g, n, h[(size_t)g * N + n], ref); ++bad; }
      }
    printf("swap_halves=%d: %d mismatches of %d\n", swap, bad, 256 * N);
  }
  return 0;
}
