// Micro-benchmark: sustained issue rate of fma.rn.f32x2 (SASS FFMA2) in the operand pattern of the pairwise epilogue
// (accumulator pair += channel pair * broadcast scalar), 4 warps per SM sub-partition.   nvcc -arch=sm_100a -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  return (unsigned long long)__float_as_uint(x) | ((unsigned long long)__float_as_uint(y) << 32);
}
template <int P, bool SCALAR>
__global__ void __launch_bounds__(512, 1) k(const float* in, float* out, int iters, long long* cyc) {
  unsigned long long acc[4][P];
  float r[8], t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { r[i] = in[threadIdx.x + i * 32]; t[i] = in[threadIdx.x + 300 + i]; }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int p = 0; p < P; ++p) acc[a][p] = 0ull;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int slot = 0; slot < 4; ++slot) {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const unsigned long long t2 = SCALAR ? pack2(t[p], t[p]) : pack2(t[p], t[(p + 1) & 7]);
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a][p] = fma2(pack2(r[2 * a], r[2 * a + 1]), t2, acc[a][p]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] += 1.0f;   // new "R" values every slot (8 FADD per 4P FFMA2)
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int p = 0; p < P; ++p) s += __uint_as_float((unsigned)acc[a][p]) + __uint_as_float((unsigned)(acc[a][p] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int P, bool SCALAR>
void run(int threads, const float* in, float* out, long long* cyc) {
  const int iters = 2000;
  k<P, SCALAR><<<148, threads>>>(in, out, iters, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / ((double)iters * 16 * P * (threads / 128));   // FFMA2 per SMSP: warps/SMSP * 16P per iteration
  printf("P=%d %s warps/SMSP=%d : %.2f cycles per FFMA2 per sub-partition (%s)\n", P, SCALAR ? "scalar-b" : "pair-b", threads / 128, per,
         cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  run<7, true>(512, in, out, cyc); run<7, true>(256, in, out, cyc); run<7, true>(128, in, out, cyc);
  run<7, false>(512, in, out, cyc);
  run<5, true>(512, in, out, cyc); run<3, true>(512, in, out, cyc); run<1, true>(512, in, out, cyc);
  return 0;
}
