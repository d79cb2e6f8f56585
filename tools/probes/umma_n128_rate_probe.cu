// Probe (B200): execution rate of M=128, N=128, K=16 bf16 tcgen05.mma in the operand forms the attention kernel uses:
// SS K-major / K-major (S = Q K^T), SS with an MN-major B (P V with P in shared memory), TS (A from tensor memory) with
// K-major or MN-major B.  One converged warp issues (elected lane), descriptors are loop invariants.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include "../../speech_resynth_b200/csrc/srb_ptx.cuh"
using namespace srb;

__device__ __forceinline__ void umma_ts_pred(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_mn128(uint32_t smem_addr, uint32_t atom_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((atom_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe(long long* out, int mode, int n, int reps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (warp == 1) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 32768;
    const uint32_t idesc_k = umma_idesc_bf16(128, n), idesc_mn = umma_idesc_bf16(128, n) | (1u << 16);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
        if (mode == 0) umma_bf16_pred(1u, tbase, umma_smem_desc<128>(a + off), umma_smem_desc<128>(b + off), idesc_k, 1u);
        else if (mode == 1) umma_bf16_pred(1u, tbase, umma_smem_desc<128>(a + off), desc_mn128(b + kk * 2048, 16384), idesc_mn, 1u);
        else if (mode == 2) umma_ts_pred(tbase, tbase + 256 + kk * 8, umma_smem_desc<128>(b + off), idesc_k, 1u);
        else umma_ts_pred(tbase, tbase + 256 + kk * 8, desc_mn128(b + kk * 2048, 16384), idesc_mn, 1u);
      }
    }
    umma_commit_pred(1u, smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  const int reps = 500;
  const char* names[4] = {"SS  A K-major smem, B K-major ", "SS  A K-major smem, B MN-major", "TS  A tensor memory, B K-major ", "TS  A tensor memory, B MN-major"};
  for (int n : {64, 128, 256})
    for (int mode = 0; mode < 4; ++mode) {
      if (n == 256 && (mode == 1 || mode == 3)) continue;   // the MN-major layout built here has two 64-wide atoms
      probe<<<1, 128, 100000>>>(d, mode, n, reps);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("M128 N%3d K16  %s: %6.1f clk per MMA  (%.0f flop/clk/SM)\n", n, names[mode], (double)h / (reps * 8), 2.0 * 128 * n * 16 / ((double)h / (reps * 8)));
    }
  return 0;
}
