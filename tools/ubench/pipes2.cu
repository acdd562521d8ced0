// pipes2.cu -- FP32 issue rate vs operand pattern (register-bank / reuse-cache effects), scalar vs packed.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 8192
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { float2 t = {a, b}; return *reinterpret_cast<u64*>(&t); }
__device__ __forceinline__ float lo(u64 v) { return reinterpret_cast<float2*>(&v)->x; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* clk, const float* in) {
  float a[8], b[8], c[8];
  u64 A[8], B[8], Cc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; c[i] = in[threadIdx.x + 16 + i];
    A[i] = pk(a[i], b[i]); B[i] = pk(b[i], c[i]); Cc[i] = pk(c[i], a[i]);
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = fmaf(a[i], a[i], a[i]);            // 1 distinct source
      if (MODE == 1) a[i] = fmaf(a[i], b[0], a[i]);            // 2 distinct, one shared
      if (MODE == 2) a[i] = fmaf(a[i], b[i], c[i]);            // 3 distinct per chain (no reuse)
      if (MODE == 3) a[i] = fmaf(b[i], c[(i + 1) & 7], a[i]);  // accumulate pattern, 3 distinct
      if (MODE == 4) a[i] = a[i] + b[i];                       // FADD 2 distinct
      if (MODE == 5) a[i] = a[i] * b[i];                       // FMUL 2 distinct
      if (MODE == 6) A[i] = fma2(A[i], A[i], A[i]);
      if (MODE == 7) A[i] = fma2(A[i], B[0], A[i]);
      if (MODE == 8) A[i] = fma2(A[i], B[i], Cc[i]);
      if (MODE == 9) A[i] = fma2(B[i], Cc[(i + 1) & 7], A[i]);
      if (MODE == 10) A[i] = add2(A[i], B[i]);
      if (MODE == 11) A[i] = mul2(A[i], B[i]);
      if (MODE == 12) { a[i] = a[i] + b[i]; c[i] = c[i] - b[i]; }   // butterfly-like: 2 FADD sharing an operand
      if (MODE == 13) { A[i] = add2(A[i], B[i]); Cc[i] = add2(Cc[i], B[i]); }
    }
  }
  long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r += a[i] + b[i] + c[i] + lo(A[i]) + lo(B[i]) + lo(Cc[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  __shared__ long long s0, s1;
  if (threadIdx.x == 0) { s0 = t0; s1 = t1; }
  __syncthreads();
  atomicMin(&s0, t0); atomicMax(&s1, t1);
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = s1 - s0;
}

template <int MODE> void run(const char* name, int per_iter, int lanes_per_inst, int threads) {
  float *out, *in; long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&in, 4096 * sizeof(float)); cudaMemset(in, 0, 4096 * sizeof(float));
  cudaMalloc(&clk, 148 * sizeof(long long));
  k<MODE><<<148, threads>>>(out, clk, in);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, clk, in);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; i++) cyc += (double)h[i]; cyc /= 148;
  double winst = (double)ITERS * per_iter * (threads / 32);
  printf("%-36s thr=%4d %.3f warp-inst/clk/SM  %.1f flop-lanes/clk/SM  [%.3f ms -> %.2f GHz] %s\n", name, threads, winst / cyc,
         winst * 32 * lanes_per_inst / cyc, ms, cyc / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(clk); cudaFree(in);
}
int main() {
  for (int threads : {256, 512, 1024}) {
    run<0>("FFMA a=a*a+a", 8, 1, threads);
    run<1>("FFMA a=a*b0+a", 8, 1, threads);
    run<2>("FFMA a=a*b+c (3 distinct)", 8, 1, threads);
    run<3>("FFMA a=b*c'+a (accumulate)", 8, 1, threads);
    run<4>("FADD a=a+b", 8, 1, threads);
    run<5>("FMUL a=a*b", 8, 1, threads);
    run<6>("FFMA2 A=A*A+A", 8, 2, threads);
    run<7>("FFMA2 A=A*B0+A", 8, 2, threads);
    run<8>("FFMA2 A=A*B+C (3 distinct)", 8, 2, threads);
    run<9>("FFMA2 A=B*C'+A (accumulate)", 8, 2, threads);
    run<10>("FADD2 A=A+B", 8, 2, threads);
    run<11>("FMUL2 A=A*B", 8, 2, threads);
    run<12>("FADD butterfly a+b, c-b", 16, 1, threads);
    run<13>("FADD2 butterfly", 16, 2, threads);
  }
  return 0;
}
