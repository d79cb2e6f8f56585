// Probe: tcgen05.mma with un-swizzled (INTERLEAVE) K-major smem descriptors; checks operand layout hypotheses and
// whether the start address may be shifted by whole rows (16 B).  nvcc -arch=sm_100a -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include "../../speech_resynth_b200/csrc/srb_ptx.cuh"
using namespace srb;

__device__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}

// A: rows R=256 (+ shift room), K=16: layout [chunk(2)][row][8]; value A[r][k] = r + k/100
// B: identity 16x16 in [chunk][n][8]
__global__ void probe(float* out, int shift, int lbo_a, int sbo_a, int lbo_b, int sbo_b) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int R = 320;
  __nv_bfloat16* A = (__nv_bfloat16*)smem;                 // 2 * R * 8 elems
  __nv_bfloat16* B = (__nv_bfloat16*)(smem + 2 * R * 16);  // 2 * 16 * 8
  for (int i = threadIdx.x; i < 2 * R * 8; i += blockDim.x) {
    int ch = i / (R * 8), r = (i / 8) % R, e = i % 8;
    int k = ch * 8 + e;
    A[i] = __float2bfloat16_rn((float)r + (k == 3 ? 0.5f : 0.f));   // only column 3 carries +0.5 so columns are distinguishable
  }
  for (int i = threadIdx.x; i < 2 * 16 * 8; i += blockDim.x) {
    int ch = i / (16 * 8), n = (i / 8) % 16, e = i % 8;
    int k = ch * 8 + e;
    B[i] = __float2bfloat16_rn(k == n ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tbase = tslot;
  if (threadIdx.x == 0) {
    uint64_t da = mkdesc(smem_u32(A) + shift * 16, lbo_a, sbo_a);
    uint64_t db = mkdesc(smem_u32(B), lbo_b, sbo_b);
    umma_bf16(tbase, da, db, umma_idesc_bf16(128, 16), 0);
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
  const int R = 320;
  struct { int shift, lbo_a, sbo_a, lbo_b, sbo_b; const char* name; } cases[] = {
    {0, R * 16, 128, 16 * 16, 128, "LBO=chunk stride, SBO=128, shift 0"},
    {1, R * 16, 128, 16 * 16, 128, "shift 1 row"},
    {5, R * 16, 128, 16 * 16, 128, "shift 5 rows"},
    {8, R * 16, 128, 16 * 16, 128, "shift 8 rows"},
    {27, R * 16, 128, 16 * 16, 128, "shift 27 rows"},
  };
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (auto& c : cases) {
    cudaMemset(d, 0xff, sizeof(h));
    probe<<<1, 128, 2 * R * 16 + 1024, 0>>>(d, c.shift, c.lbo_a, c.sbo_a, c.lbo_b, c.sbo_b);
    cudaError_t e = cudaDeviceSynchronize();
    printf("== %s : %s\n", c.name, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r) for (int n = 0; n < 16; ++n) {
      float want = (float)(r + c.shift) + (n == 3 ? 0.5f : 0.f);
      if (h[r * 16 + n] != want) ++bad;
    }
    printf("   mismatches vs expected D[r][n] = r+shift (+.5 at n=3): %d / 2048\n", bad);
    for (int r : {0, 1, 9, 127}) { printf("   row %3d:", r); for (int n = 0; n < 16; ++n) printf(" %g", h[r * 16 + n]); printf("\n"); }
  }
  return 0;
}
