// Microbenchmark (round 2): instruction delivery with 8 warps per CTA sharing one SM's instruction stream, body sizes around the
// 128 KB tier (the biped kernel's step is 171 KB, the Barkour kernel's 63 KB). usage: ./icache2
#include <cstdio>
#include <cuda_runtime.h>
template <int BODY> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, long long* cyc) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < BODY / 8; k++) {
      x0 = fmaf(x0, x1, b); x1 = fmaf(x1, x2, a); x2 = fmaf(x2, x3, b); x3 = fmaf(x3, x4, a);
      x4 = fmaf(x4, x5, b); x5 = fmaf(x5, x6, a); x6 = fmaf(x6, x7, b); x7 = fmaf(x7, x0, a);
    }
  }
  long long t1 = clock64();
  const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int BODY> void run(int tpb, int grid) {
  float* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
  int iters = (1 << 22) / BODY; if (iters < 4) iters = 4;
  k<BODY><<<grid, tpb>>>(d, iters, 1.0000001f, 1e-9f, c);
  k<BODY><<<grid, tpb>>>(d, iters, 1.0000001f, 1e-9f, c);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("body %6d instr (%4d KB)  tpb %3d grid %4d : %.3f cycles/instr per warp\n", BODY, BODY * 16 / 1024, tpb, grid, (double)h / ((double)iters * BODY));
  cudaFree(d); cudaFree(c);
}
int main() {
  const int cfgs[4][2] = {{32, 512}, {256, 64}, {256, 148}, {256, 296}};
  for (auto& c : cfgs) {
    run<4096>(c[0], c[1]); run<6144>(c[0], c[1]); run<7168>(c[0], c[1]); run<8192>(c[0], c[1]); run<9216>(c[0], c[1]); run<10240>(c[0], c[1]);
    run<11264>(c[0], c[1]); run<12288>(c[0], c[1]); run<14336>(c[0], c[1]); run<16384>(c[0], c[1]);
  }
  return 0;
}
