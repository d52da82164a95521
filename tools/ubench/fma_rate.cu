// fma_rate.cu -- measures FFMA / FFMA2 issue rates on sm_100a (tools only).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fma_rate tools/ubench/fma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
struct P { float m[64]; };
template <int MODE, int CH>
__global__ void __launch_bounds__(256) k(float* out, const __grid_constant__ P p, int iters, long long* cyc) {
  float s = threadIdx.x * 1e-3f;
  long long t0 = clock64();
  if constexpr (MODE == 0) {          // scalar FFMA, uniform-register multiplier
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = s + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = fmaf(acc[c], p.m[u], p.m[u + 8]);
    }
    float r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else if constexpr (MODE == 1) {   // FFMA2, pair x broadcast uniform scalar
    uint64_t acc[CH], x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { acc[c] = pack2(s + c, s - c); x[c] = pack2(s * c, s + 2 * c); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = fma2(x[c], pack2(p.m[u], p.m[u]), acc[c]);
    }
    uint64_t r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r ^= acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float) r;
  } else {                            // FFMA2, three register pairs
    uint64_t acc[CH], x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { acc[c] = pack2(s + c, s - c); x[c] = pack2(s * c, s + 2 * c); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = fma2(x[c], x[(c + 1) % CH], acc[c]);
    }
    uint64_t r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r ^= acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float) r;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int CH>
void run(const char* name, int blocks_per_sm, int threads) {
  int iters = 4096;
  float* out; long long* cyc;
  int blocks = 148 * blocks_per_sm;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
  P p; for (int i = 0; i < 64; ++i) p.m[i] = 1.0f + i * 1e-6f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, CH><<<blocks, threads>>>(out, p, 16, cyc);
  cudaEventRecord(e0);
  k<MODE, CH><<<blocks, threads>>>(out, p, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
  double instr_per_warp = (double) iters * 8 * CH;
  int warps_per_smsp = blocks_per_sm * threads / 32 / 4;
  double cyc_per_instr_smsp = (double) c0 / (instr_per_warp * warps_per_smsp);
  double fma_per_lane = MODE == 0 ? 1 : 2;
  double tflops = 2.0 * fma_per_lane * instr_per_warp * (blocks * threads) / (ms * 1e-3) / 1e12;
  printf("%-28s CH=%d warps/SMSP=%d  cycles/instr/SMSP=%.3f  %.2f ms  %.1f TFLOP/s  (clk est %.0f MHz)\n", name, CH, warps_per_smsp,
         cyc_per_instr_smsp, ms, tflops, c0 / (ms * 1e3));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 8>("FFMA  R*UR+UR", 2, 256);
  run<0, 8>("FFMA  R*UR+UR", 4, 256);
  run<1, 8>("FFMA2 pair*URscalar", 1, 128);
  run<1, 8>("FFMA2 pair*URscalar", 2, 256);
  run<1, 8>("FFMA2 pair*URscalar", 4, 256);
  run<1, 4>("FFMA2 pair*URscalar", 4, 256);
  run<1, 2>("FFMA2 pair*URscalar", 4, 256);
  run<2, 8>("FFMA2 pair*pair", 2, 256);
  run<2, 8>("FFMA2 pair*pair", 4, 256);
  return 0;
}
