// Microbenchmark (round 2): which instructions mixed into an FFMA stream produce `dispatch` / `wait` stalls with two warps per
// sub-partition? Eight independent FFMA chains; every 8th slot is replaced by the probe instruction. Body 1536 instructions (24 KB:
// inside the first instruction-delivery tier). Run under ncu with the smsp__average_warps_issue_stalled_* metrics.
// usage: ./mix
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, const float* in, int iters, float a, float b, long long* cyc) {
  __shared__ float sm[256 * 4];
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  float y = in[threadIdx.x & 3];
  sm[threadIdx.x] = x0; sm[256 + threadIdx.x] = x1;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < 192; q++) {
      x0 = fmaf(x0, x1, b); x1 = fmaf(x1, x2, a); x2 = fmaf(x2, x3, b); x3 = fmaf(x3, x4, a);
      x4 = fmaf(x4, x5, b); x5 = fmaf(x5, x6, a); x6 = fmaf(x6, x7, b);
      if (MODE == 0) x7 = fmaf(x7, x0, a);                                        // pure FFMA
      if (MODE == 1) x7 = (x0 > y) ? x7 : x1;                                     // FSETP + FSEL (alu pipe)
      if (MODE == 2) { if (x3 > y) x7 += x5; }                                    // FSETP + predicated FADD
      if (MODE == 3) x7 = fmaxf(x7, x0);                                          // FMNMX (alu pipe)
      if (MODE == 4) x7 += __shfl_xor_sync(0xffffffffu, x5, 1);                   // SHFL + FADD
      if (MODE == 5) x7 += sm[(q & 1) * 256 + threadIdx.x];                       // LDS + FADD
      if (MODE == 6) x7 = x7 * x0;                                                // FMUL
      if (MODE == 7) x7 = x7 + x0;                                                // FADD
      if (MODE == 8) asm("fma.rn.f32 %0, %1, 0f3F800000, %2;" : "=f"(x7) : "f"(x7), "f"(x0));   // the same sum as an FFMA with multiplier 1.0
      if (MODE == 9) asm("fma.rn.f32 %0, %1, %2, 0f80000000;" : "=f"(x7) : "f"(x7), "f"(x0));   // the same product as an FFMA with addend -0.0
      if (MODE == 10) { float t = (x3 > y) ? x5 : 0.f; asm("fma.rn.f32 %0, %1, 0f3F800000, %2;" : "=f"(x7) : "f"(t), "f"(x7)); }  // select then FFMA-add
    }
  }
  long long t1 = clock64();
  const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name) {
  float *d, *in; long long* c; cudaMalloc(&d, 4); cudaMalloc(&in, 16); cudaMalloc(&c, 8);
  float h4[4] = {1e30f, 1e30f, 1e30f, 1e30f}; cudaMemcpy(in, h4, 16, cudaMemcpyHostToDevice);
  const int iters = 2000;
  k<MODE><<<148, 256>>>(d, in, iters, 1.0000001f, 1e-9f, c);
  k<MODE><<<148, 256>>>(d, in, iters, 1.0000001f, 1e-9f, c);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-28s : %.3f cycles per source slot per warp\n", name, (double)h / ((double)iters * 192 * 8));
  cudaFree(d); cudaFree(in); cudaFree(c);
}
int main() {
  run<0>("pure FFMA"); run<1>("1/8 FSETP+FSEL"); run<2>("1/8 FSETP+@P FADD"); run<3>("1/8 FMNMX"); run<4>("1/8 SHFL+FADD");
  run<5>("1/8 LDS+FADD"); run<6>("1/8 FMUL"); run<7>("1/8 FADD"); run<8>("1/8 FADD as FFMA(x,1,y)"); run<9>("1/8 FMUL as FFMA(x,y,-0)"); run<10>("1/8 FSETP+FSEL+FFMA");
  return 0;
}
