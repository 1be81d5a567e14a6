// Microbenchmark: FFMA vs FFMA2 (packed f32x2, sm_100) issue rate, for a lone warp per CTA and for a full SM,
// at loop bodies inside and outside the 32 KB instruction-cache tier. Prints cycles per warp-instruction.
#include <cstdio>
#include <cuda_runtime.h>
template <int N, bool PACKED> struct Rep {
  static __device__ __forceinline__ void go(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4, float2& x5, float2& x6, float2& x7, float2 A, float2 B) {
    Rep<N / 2, PACKED>::go(x0, x1, x2, x3, x4, x5, x6, x7, A, B);
    Rep<N / 2, PACKED>::go(x0, x1, x2, x3, x4, x5, x6, x7, A, B);
  }
};
template <bool PACKED> struct Rep<8, PACKED> {
  static __device__ __forceinline__ void go(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4, float2& x5, float2& x6, float2& x7, float2 A, float2 B) {
    if (PACKED) {
      x0 = __ffma2_rn(x0, x1, B); x1 = __ffma2_rn(x1, x2, A); x2 = __ffma2_rn(x2, x3, B); x3 = __ffma2_rn(x3, x4, A);
      x4 = __ffma2_rn(x4, x5, B); x5 = __ffma2_rn(x5, x6, A); x6 = __ffma2_rn(x6, x7, B); x7 = __ffma2_rn(x7, x0, A);
    } else {
      x0.x = fmaf(x0.x, x1.x, B.x); x1.x = fmaf(x1.x, x2.x, A.x); x2.x = fmaf(x2.x, x3.x, B.x); x3.x = fmaf(x3.x, x4.x, A.x);
      x4.x = fmaf(x4.x, x5.x, B.x); x5.x = fmaf(x5.x, x6.x, A.x); x6.x = fmaf(x6.x, x7.x, B.x); x7.x = fmaf(x7.x, x0.x, A.x);
    }
  }
};
template <int BODY, bool PACKED> __global__ void __launch_bounds__(1024) k(float* out, int iters, float a, float b, long long* cyc) {
  float2 x0 = {threadIdx.x + 0.f, 1.f}, x1 = {x0.x + 1.f, 2.f}, x2 = {x0.x + 2.f, 3.f}, x3 = {x0.x + 3.f, 4.f};
  float2 x4 = {x0.x + 4.f, 5.f}, x5 = {x0.x + 5.f, 6.f}, x6 = {x0.x + 6.f, 7.f}, x7 = {x0.x + 7.f, 8.f};
  const float2 A = {a, a}, B = {b, b};
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) Rep<BODY, PACKED>::go(x0, x1, x2, x3, x4, x5, x6, x7, A, B);
  long long t1 = clock64();
  const float s = x0.x + x1.x + x2.x + x3.x + x4.x + x5.x + x6.x + x7.x + x0.y + x1.y + x2.y + x3.y + x4.y + x5.y + x6.y + x7.y;
  if (s == 12345.678f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int BODY, bool PACKED> void run(int tpb, int grid) {
  float* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
  int iters = (1 << 22) / BODY; if (iters < 4) iters = 4;
  k<BODY, PACKED><<<grid, tpb>>>(d, iters, 1.0000001f, 1e-9f, c);
  k<BODY, PACKED><<<grid, tpb>>>(d, iters, 1.0000001f, 1e-9f, c);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%s body %6d instr (%4d KB)  tpb %4d grid %4d : %.3f cycles/instr per warp, %.3f per SM sub-partition\n", PACKED ? "FFMA2" : "FFMA ",
         BODY, BODY * 16 / 1024, tpb, grid, (double)h / ((double)iters * BODY), (double)h / ((double)iters * BODY) / ((tpb + 127) / 128));
  cudaFree(d); cudaFree(c);
}
int main() {
  const int tpbs[5] = {32, 128, 256, 512, 1024};  // 1 warp per CTA (one sub-partition busy) ... 8 warps per sub-partition
  for (int c = 0; c < 5; c++) {
    run<1024, false>(tpbs[c], 148); run<1024, true>(tpbs[c], 148);
    run<8192, false>(tpbs[c], 148); run<8192, true>(tpbs[c], 148);
  }
  return 0;
}
