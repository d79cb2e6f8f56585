// Probe: can a SWIZZLE_128B K-major A operand be addressed at a start shifted by whole 128-byte rows (implicit-conv
// taps over one halo-resident tile)?  Tries base_offset = 0 and base_offset = (start >> 7) & 7.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include "../../speech_resynth_b200/csrc/srb_ptx.cuh"
using namespace srb;

__device__ uint64_t mkdesc128(uint32_t addr, int base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= 2ull << 61;
  return d;
}

// A: R rows x 64 bf16 (128 B rows), TMA-style 128B swizzle: 16-byte chunk c of row r stored at chunk (c ^ (r & 7)).
// value A[r][k] = r (+0.5 at k == 3).  B = [n][k] 16 x 64 with B[n][k] = (k == n): picks columns 0..15 of A.
__global__ void probe(float* out, int shift, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int R = 192;
  __nv_bfloat16* A = (__nv_bfloat16*)smem;
  __nv_bfloat16* B = (__nv_bfloat16*)(smem + R * 128);
  for (int i = threadIdx.x; i < R * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64;
    int c = k / 8, e = k % 8;
    A[r * 64 + ((c ^ (r & 7)) * 8) + e] = __float2bfloat16_rn((float)r + (k == 3 ? 0.5f : 0.f));
  }
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    int n = i / 64, k = i % 64;
    int c = k / 8, e = k % 8;
    B[n * 64 + ((c ^ (n & 7)) * 8) + e] = __float2bfloat16_rn(k == n ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tbase = tslot;
  if (threadIdx.x == 0) {
    uint32_t a_addr = smem_u32(A) + shift * 128;
    int bo = mode == 0 ? 0 : ((a_addr >> 7) & 7);
    for (int k = 0; k < 4; ++k)   // K = 64 = 4 x 16
      umma_bf16(tbase, mkdesc128(a_addr, bo) + 2 * k, mkdesc128(smem_u32(B), 0) + 2 * k, umma_idesc_bf16(128, 16), k != 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[16];
  tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 32);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  float h[128 * 16];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 1, 3, 8, 13, 27}) {
      cudaMemset(d, 0xff, sizeof(h));
      probe<<<1, 128, 192 * 128 + 16 * 128 + 2048>>>(d, shift, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: %s\n", mode, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 100; ++r) for (int n = 0; n < 16; ++n) {
        float want = (float)(r + shift) + (n == 3 ? 0.5f : 0.f);
        if (h[r * 16 + n] != want) ++bad;
      }
      printf("mode %d (base_offset %s) shift %2d: mismatches %4d / 1600   row0: %g %g %g %g  row1: %g %g %g %g  row9: %g %g\n", mode,
             mode ? "=(addr>>7)&7" : "=0", shift, bad, h[0], h[1], h[2], h[3], h[16], h[17], h[18], h[19], h[9 * 16], h[9 * 16 + 3]);
    }
  return 0;
}
