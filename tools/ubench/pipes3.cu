// pipes3.cu -- do SHFL and LDS/STS share the shared-memory data pipe?  (decides whether the FFT exchange should move
// from the padded shared-memory buffer to warp shuffles).  Modes: LDS.64 alone, SHFL alone, both interleaved, STS+LDS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes3 pipes3.cu && ./pipes3
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* clk, float seed) {
  __shared__ float4 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  int idx = threadIdx.x;
  const float2* s = reinterpret_cast<const float2*>(sm);
  float2* sw = reinterpret_cast<float2*>(sm);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0 || MODE == 2) {   // 4 LDS.64
      float2 v;
      v = s[idx & 4095]; a0 += v.x; a1 += v.y; v = s[(idx + 32) & 4095]; a2 += v.x; a3 += v.y;
      v = s[(idx + 64) & 4095]; a4 += v.x; a5 += v.y; v = s[(idx + 96) & 4095]; a6 += v.x; a7 += v.y;
      idx += 128;
    }
    if (MODE == 1 || MODE == 2) {   // 8 SHFL
      a0 = __shfl_xor_sync(0xffffffffu, a0, 1); a1 = __shfl_xor_sync(0xffffffffu, a1, 2);
      a2 = __shfl_xor_sync(0xffffffffu, a2, 4); a3 = __shfl_xor_sync(0xffffffffu, a3, 8);
      a4 = __shfl_xor_sync(0xffffffffu, a4, 16); a5 = __shfl_xor_sync(0xffffffffu, a5, 3);
      a6 = __shfl_xor_sync(0xffffffffu, a6, 5); a7 = __shfl_xor_sync(0xffffffffu, a7, 7);
    }
    if (MODE == 3) {   // 4 STS.64 + 4 LDS.64 (an exchange through shared memory: 8 values out, 8 values in)
      sw[idx & 4095] = make_float2(a0, a1); sw[(idx + 32) & 4095] = make_float2(a2, a3);
      sw[(idx + 64) & 4095] = make_float2(a4, a5); sw[(idx + 96) & 4095] = make_float2(a6, a7);
      __syncwarp();
      float2 v;
      v = s[(idx ^ 1) & 4095]; a0 += v.x; a1 += v.y; v = s[((idx ^ 1) + 32) & 4095]; a2 += v.x; a3 += v.y;
      v = s[((idx ^ 1) + 64) & 4095]; a4 += v.x; a5 += v.y; v = s[((idx ^ 1) + 96) & 4095]; a6 += v.x; a7 += v.y;
      __syncwarp();
      idx += 128;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + idx;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&clk, 148 * sizeof(long long));
  k<MODE><<<148, threads>>>(out, clk, 1.0f); cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(out, clk, 1.0f); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; i++) cyc += (double)h[i]; cyc /= 148;
  printf("%-34s threads=%4d  %.2f clk per iteration per warp-slot  (%.2f clk/iter/SM-warp)  err=%s\n", name, threads,
         cyc / ITERS, cyc / ITERS / (threads / 32), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int threads : {256, 1024}) {
    run<0>("4 LDS.64", threads);
    run<1>("8 SHFL", threads);
    run<2>("4 LDS.64 + 8 SHFL", threads);
    run<3>("4 STS.64 + 4 LDS.64 (+2 syncwarp)", threads);
  }
  return 0;
}
