// Probe (B200): throughput of MUFU.EX2 (ex2.approx.ftz.f32), of the bf16x2 pack (cvt.rn.bf16x2.f32) and of an FMA-pipe
// exp2 (Cody-Waite + cubic) per SM, with W warps resident -- what a softmax warp can expect per clock.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ float exp2_poly(float x) {
  const float t = x + 12582912.f;                    // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);              // f in [-0.5, 0.5]
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__global__ void probe(long long* out, float* sink, int mode, int reps) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t u = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) { a0 = ex2f(a0) - 1.f; a1 = ex2f(a1) - 1.f; a2 = ex2f(a2) - 1.f; a3 = ex2f(a3) - 1.f; }
    else if (mode == 1) { u ^= pack(a0, a1); u += pack(a2, a3); a0 += 1.f; a2 += 1.f; u ^= pack(a1, a0); u += pack(a3, a2); }
    else { a0 = exp2_poly(a0) - 1.f; a1 = exp2_poly(a1) - 1.f; a2 = exp2_poly(a2) - 1.f; a3 = exp2_poly(a3) - 1.f; }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (a0 + a1 + a2 + a3 + (float)u == 1234.5f) sink[0] = a0;
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  float* sink; cudaMalloc(&sink, 64);
  long long h;
  const int reps = 4000;
  const char* names[3] = {"ex2.approx.ftz.f32 (MUFU)   ", "cvt.rn.bf16x2.f32 (pack)    ", "exp2 on the FMA pipe (cubic)"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      probe<<<1, warps * 32>>>(d, sink, mode, reps);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%s  %2d warps: %6.2f results per clk per SM\n", names[mode], warps, (double)warps * 32 * 4 * reps / (double)h);
    }
  return 0;
}
