// Microbenchmark (round 2): does a kernel whose loop is TWO code phases A and B (each under the 128 KB instruction-delivery tier,
// together beyond it) run in the fast tier when each phase is repeated K times before switching (loop order "for phase: for set"
// instead of "for set: for phase")? `desync`: odd CTAs start with phase B, so neighbouring SMs stream different phases.
// usage: ./icache3
#include <cstdio>
#include <cuda_runtime.h>
#define BODY8(a, b)                                                                                  \
  x0 = fmaf(x0, x1, b); x1 = fmaf(x1, x2, a); x2 = fmaf(x2, x3, b); x3 = fmaf(x3, x4, a);            \
  x4 = fmaf(x4, x5, b); x5 = fmaf(x5, x6, a); x6 = fmaf(x6, x7, b); x7 = fmaf(x7, x0, a);
template <int BODY> __global__ void __launch_bounds__(256) k(float* out, int iters, int K, int desync, float a, float b, long long* cyc) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  const int first = (desync && (blockIdx.x & 1)) ? 1 : 0;
  long long t0 = clock64();
  for (int i = 0; i < 2 * iters; i++) {
    if (((i + first) & 1) == 0) {
#pragma unroll 1
      for (int r = 0; r < K; r++) {
#pragma unroll
        for (int q = 0; q < BODY / 8; q++) { BODY8(a, b) }
      }
    } else {
#pragma unroll 1
      for (int r = 0; r < K; r++) {
#pragma unroll
        for (int q = 0; q < BODY / 8; q++) { BODY8(b, a) }
      }
    }
  }
  long long t1 = clock64();
  const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int BODY> void run(int tpb, int grid, int K, int desync) {
  float* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
  int iters = (1 << 21) / (BODY * K); if (iters < 2) iters = 2;
  k<BODY><<<grid, tpb>>>(d, iters, K, desync, 1.0000001f, 1e-9f, c);
  k<BODY><<<grid, tpb>>>(d, iters, K, desync, 1.0000001f, 1e-9f, c);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("2 phases x %5d instr (%3d KB each)  K %2d  %s  tpb %3d grid %4d : %.3f cycles/instr per warp\n", BODY, BODY * 16 / 1024, K,
         desync ? "desync" : "sync  ", tpb, grid, (double)h / ((double)iters * 2 * K * BODY));
  cudaFree(d); cudaFree(c);
}
int main() {
  const int Ks[5] = {1, 2, 4, 8, 16};
  for (int desync = 0; desync < 2; desync++)
    for (int K : Ks) { run<3072>(256, 148, K, desync); run<5632>(256, 148, K, desync); run<6144>(256, 148, K, desync); run<7168>(256, 148, K, desync); }
  for (int K : Ks) { run<5632>(256, 64, K, 0); run<6144>(32, 512, K, 0); }
  return 0;
}
