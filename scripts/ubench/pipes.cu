// Micro-benchmarks of the FP32 pipe on sm_100a: scalar FFMA (3 distinct registers) against the packed FFMA2
// (fma.rn.f32x2), dependent chains (latency) and 8 independent chains (throughput), 1 / 2 / 4 warps per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmul(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

#define ITERS 2048
// mode 0: FFMA dependent chain; 1: FFMA 8 chains; 2: FFMA2 dependent; 3: FFMA2 8 chains; 4: FMUL 8 chains (2 regs);
// 5: FFMA 8 chains with 2 shared multiplicands (b, c loaded from memory: 3-reg form, 2 of them common); 6: FFMA2 4 chains;
// 7: FFMA 4 chains; 8: mixed FFMA2 + FFMA alternating 8 chains; 9: SHFL.BFLY dependent; 10: MUFU.RCP dependent; 11: FFMA 2 chains; 12: FFMA2 2 chains
__global__ void bench(int mode, const float* in, float* out, long long* cyc) {
  float x[8], y[8];
  for (int i = 0; i < 8; ++i) { x[i] = in[threadIdx.x + 32 * i]; y[i] = in[threadIdx.x + 32 * i + 7]; }
  float b = in[300 + (threadIdx.x & 1)], c = in[310 + (threadIdx.x & 1)];
  u64 X[8], B, C;
  for (int i = 0; i < 8; ++i) X[i] = ((u64)__float_as_uint(x[i]) << 32) | __float_as_uint(y[i]);
  B = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b); C = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) x[0] = ffma(x[0], b, c); } }
  else if (mode == 1 || mode == 5) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = ffma(x[r], b, c); } }
  else if (mode == 2) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) X[0] = ffma2(X[0], B, C); } }
  else if (mode == 3) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) X[r] = ffma2(X[r], B, C); } }
  else if (mode == 4) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = fmul(x[r], b); } }
  else if (mode == 6) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) X[r & 3] = ffma2(X[r & 3], B, C); } }
  else if (mode == 7) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r & 3] = ffma(x[r & 3], b, c); } }
  else if (mode == 8) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) { X[r] = ffma2(X[r], B, C); x[r + 4] = ffma(x[r + 4], b, c); } } }
  else if (mode == 9) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) x[0] = __shfl_xor_sync(0xffffffffu, x[0], 1 + (r & 1)); } }
  else if (mode == 10) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) { float d; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(x[0])); x[0] = d; } } }
  else if (mode == 11) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r & 1] = ffma(x[r & 1], b, c); } }
  else if (mode == 12) { for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) X[r & 1] = ffma2(X[r & 1], B, C); } }
  else if (mode == 13) {   // 8 chains, all three operands distinct per chain (x = x*y + x' pattern): register-bank pressure
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = ffma(x[r], y[r], y[(r + 1) & 7]); } }
  else if (mode == 14) {
    u64 Y[8]; for (int i = 0; i < 8; ++i) Y[i] = X[(i + 3) & 7] ^ 0x0000100000001000ull;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) X[r] = ffma2(X[r], Y[r], Y[(r + 1) & 7]); } }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float((unsigned)(X[i] >> 32)) + __uint_as_float((unsigned)X[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* in; float* out; long long* cyc;
  cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 0.5f + 1e-3f * (i % 97);
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  const char* names[] = {"FFMA dep", "FFMA x8", "FFMA2 dep", "FFMA2 x8", "FMUL x8", "FFMA x8 (again)", "FFMA2 x4", "FFMA x4", "FFMA2+FFMA x4+4",
                         "SHFL dep", "MUFU.RCP dep", "FFMA x2", "FFMA2 x2", "FFMA x8 3 distinct regs", "FFMA2 x8 3 distinct regs"};
  for (int wps = 1; wps <= 4; wps *= 2)
    for (int mode = 0; mode < 15; ++mode) {
      bench<<<1, 128 * wps>>>(mode, in, out, cyc); cudaDeviceSynchronize();
      bench<<<1, 128 * wps>>>(mode, in, out, cyc); cudaDeviceSynchronize();
      printf("warps/sched %d  %-28s %7.3f cycles per instruction (per warp)\n", wps, names[mode], (double)*cyc / (ITERS * 8.0));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
