// pipes.cu -- issue-rate microbenchmarks for the B200 SM: scalar vs packed FP32, shared-memory loads, shuffles.
// Used once to pick the instruction mix of the filter-bank kernels (DESIGN.md "instruction budget").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
  unsigned long long dd, aa, bb;
  aa = *reinterpret_cast<unsigned long long*>(&a);
  bb = *reinterpret_cast<unsigned long long*>(&b);
  dd = *reinterpret_cast<unsigned long long*>(&d);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void fadd2(float2& d, float2 a) {
  unsigned long long dd, aa;
  aa = *reinterpret_cast<unsigned long long*>(&a);
  dd = *reinterpret_cast<unsigned long long*>(&d);
  asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(dd) : "l"(aa));
  d = *reinterpret_cast<float2*>(&dd);
}

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* clk, float seed) {
  __shared__ float4 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  float2 p0 = {seed, seed}, p1 = {seed + 1, seed}, p2 = {seed + 2, seed}, p3 = {seed + 3, seed}, p4 = {seed + 4, seed},
         p5 = {seed + 5, seed}, p6 = {seed + 6, seed}, p7 = {seed + 7, seed};
  const float b = 1.0001f, c = 0.5f;
  const float2 b2 = {1.0001f, 0.9999f}, c2 = {0.5f, 0.25f};
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {   // scalar FFMA, 8 independent chains
      a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    } else if (MODE == 1) {  // packed FFMA2
      ffma2(p0, b2, c2); ffma2(p1, b2, c2); ffma2(p2, b2, c2); ffma2(p3, b2, c2);
      ffma2(p4, b2, c2); ffma2(p5, b2, c2); ffma2(p6, b2, c2); ffma2(p7, b2, c2);
    } else if (MODE == 2) {  // scalar FADD
      a0 += b; a1 += b; a2 += b; a3 += b; a4 += b; a5 += b; a6 += b; a7 += b;
    } else if (MODE == 3) {  // packed FADD2
      fadd2(p0, b2); fadd2(p1, b2); fadd2(p2, b2); fadd2(p3, b2); fadd2(p4, b2); fadd2(p5, b2); fadd2(p6, b2); fadd2(p7, b2);
    } else if (MODE == 4) {  // LDS.32 conflict-free
      const float* s = reinterpret_cast<const float*>(sm);
      a0 += s[idx & 8191]; a1 += s[(idx + 32) & 8191]; a2 += s[(idx + 64) & 8191]; a3 += s[(idx + 96) & 8191];
      a4 += s[(idx + 128) & 8191]; a5 += s[(idx + 160) & 8191]; a6 += s[(idx + 192) & 8191]; a7 += s[(idx + 224) & 8191];
      idx += 256;
    } else if (MODE == 5) {  // LDS.64
      const float2* s = reinterpret_cast<const float2*>(sm);
      float2 v;
      v = s[idx & 4095]; a0 += v.x + v.y; v = s[(idx + 32) & 4095]; a1 += v.x + v.y;
      v = s[(idx + 64) & 4095]; a2 += v.x + v.y; v = s[(idx + 96) & 4095]; a3 += v.x + v.y;
      idx += 128;
    } else if (MODE == 6) {  // LDS.128
      float4 v;
      v = sm[idx & 2047]; a0 += v.x + v.y + v.z + v.w; v = sm[(idx + 32) & 2047]; a1 += v.x + v.y + v.z + v.w;
      idx += 64;
    } else if (MODE == 7) {  // SHFL
      a0 = __shfl_xor_sync(0xffffffffu, a0, 1); a1 = __shfl_xor_sync(0xffffffffu, a1, 2);
      a2 = __shfl_xor_sync(0xffffffffu, a2, 4); a3 = __shfl_xor_sync(0xffffffffu, a3, 8);
      a4 = __shfl_xor_sync(0xffffffffu, a4, 16); a5 = __shfl_xor_sync(0xffffffffu, a5, 3);
      a6 = __shfl_xor_sync(0xffffffffu, a6, 5); a7 = __shfl_xor_sync(0xffffffffu, a7, 7);
    } else if (MODE == 8) {  // FFMA + LDS.64 mixed: 8 FFMA per LDS.64
      const float2* s = reinterpret_cast<const float2*>(sm);
      float2 v = s[idx & 4095]; idx += 32;
      a0 = fmaf(a0, b, v.x); a1 = fmaf(a1, b, v.y); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    } else if (MODE == 9) {  // FFMA2 + scalar FADD mix (4 + 4)
      ffma2(p0, b2, c2); a0 += b; ffma2(p1, b2, c2); a1 += b; ffma2(p2, b2, c2); a2 += b; ffma2(p3, b2, c2); a3 += b;
    } else if (MODE == 10) {  // STS.64
      float2* s = reinterpret_cast<float2*>(sm);
      s[idx & 4095] = make_float2(a0, a1); s[(idx + 32) & 4095] = make_float2(a1, a2);
      s[(idx + 64) & 4095] = make_float2(a2, a3); s[(idx + 96) & 4095] = make_float2(a3, a0);
      idx += 128; a0 += 1.f;
    } else if (MODE == 11) {  // FFMA with IADD interleaved (dual-pipe issue test): 8 FFMA + 8 IADD
      a0 = fmaf(a0, b, c); idx += 3; a1 = fmaf(a1, b, c); idx ^= 5; a2 = fmaf(a2, b, c); idx += 7; a3 = fmaf(a3, b, c); idx ^= 9;
      a4 = fmaf(a4, b, c); idx += 11; a5 = fmaf(a5, b, c); idx ^= 13; a6 = fmaf(a6, b, c); idx += 17; a7 = fmaf(a7, b, c); idx ^= 19;
    }
  }
  long long t1 = clock64();
  float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y + p4.x + p4.y +
            p5.x + p5.y + p6.x + p6.y + p7.x + p7.y + idx;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int per_iter, int threads) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&clk, 148 * sizeof(long long));
  k<MODE><<<148, threads>>>(out, clk, 1.0f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, clk, 1.0f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; i++) cyc += (double)h[i]; cyc /= 148;
  double winst = (double)ITERS * per_iter * (threads / 32);
  printf("%-28s threads=%4d  %8.0f clk  %.3f warp-inst/clk/SM  (%.1f lane-ops/clk/SM)  %.3f ms  err=%s\n", name, threads, cyc,
         winst / cyc, winst * 32 / cyc, ms, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int threads : {256, 512, 1024}) {
    run<0>("FFMA scalar", 8, threads);
    run<1>("FFMA2 packed", 8, threads);
    run<2>("FADD scalar", 8, threads);
    run<3>("FADD2 packed", 8, threads);
    run<4>("LDS.32", 8, threads);
    run<5>("LDS.64", 4, threads);
    run<6>("LDS.128", 2, threads);
    run<7>("SHFL", 8, threads);
    run<8>("8 FFMA + 1 LDS.64", 9, threads);
    run<9>("4 FFMA2 + 4 FADD", 8, threads);
    run<10>("STS.64", 4, threads);
    run<11>("8 FFMA + 8 IADD/LOP", 16, threads);
  }
  return 0;
}
