// Probe (B200): two tcgen05.mma operand forms the attention kernel wants.
//  (1) A operand from TENSOR MEMORY (kind::f16 "TS" form): P = bf16 pairs packed in 32-bit TMEM columns, written with
//      tcgen05.st by the thread that owns the row.  Which half of a column is the lower K index?
//  (2) B operand MN-major (N contiguous) in 128-byte-swizzled shared memory: a V tile [keys][d] as TMA delivers it from a
//      (frames, d) row-major buffer.  Descriptor: LBO = distance between 64-element N atoms, SBO = distance between groups
//      of 8 K rows (1024 B); instruction descriptor bit 16 (b_major) = 1.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include "../../speech_resynth_b200/csrc/srb_ptx.cuh"
using namespace srb;

__device__ uint64_t mkdesc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;   // SWIZZLE_128B
  return d;
}
__device__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

// mode 0: TS form.  A[r][k] = (r % 4) * 64 + k in TMEM (k < 64: 32 packed columns at column 64), pack order `order`;
//         B[n][k] = (k == n + noff) K-major SW128 -> D[r][n] = A[r][n + noff].
// mode 1: MN-major B.  A[r][k] = (k == r % 64) K-major SW128 smem; V[k][n] = (k * 7 + n) % 256, n < ncols (64 or 128),
//         stored [k][64-element atom] rows of 128 B, SW128; D[r][n] = V[r % 64][n].  variant: 0 = (LBO atom, SBO 1024), 1 = swapped
__global__ void probe(float* out, int mode, int arg, int variant) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* A = (__nv_bfloat16*)smem;                 // 128 rows x 64 (16 KB)
  __nv_bfloat16* B = (__nv_bfloat16*)(smem + 16384);       // up to 32 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  const int ncols = mode == 0 ? 16 : arg;
  if (mode == 0) {
    const int noff = arg;
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
      int n = i / 64, k = i % 64, c = k / 8, e = k % 8;
      B[n * 64 + ((c ^ (n & 7)) * 8) + e] = __float2bfloat16_rn(k == n + noff ? 1.f : 0.f);
    }
    // thread <-> row: 32 packed columns
    const int r = warp * 32 + lane;
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) {
      float lo = (float)((r % 4) * 64 + 2 * c), hi = lo + 1.f;
      v[c] = variant == 0 ? pack_bf16(lo, hi) : pack_bf16(hi, lo);
    }
    tmem_st32(tbase + ((uint32_t)(warp * 32) << 16) + 64, v);
    tmem_st_wait();
  } else {
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
      int r = i / 64, k = i % 64, c = k / 8, e = k % 8;
      A[r * 64 + ((c ^ (r & 7)) * 8) + e] = __float2bfloat16_rn(k == r % 64 ? 1.f : 0.f);
    }
    for (int i = threadIdx.x; i < 64 * ncols; i += blockDim.x) {
      int k = i / ncols, n = i % ncols, atom = n / 64, nn = n % 64, c = nn / 8, e = nn % 8;
      B[atom * 64 * 64 + k * 64 + ((c ^ (k & 7)) * 8) + e] = __float2bfloat16_rn((float)((k * 7 + n) % 256));
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    if (mode == 0) {
      for (int k = 0; k < 4; ++k)
        umma_ts(tbase, tbase + 64 + 8 * k, mkdesc(smem_u32(B), 16, 1024) + 2 * k, umma_idesc_bf16(128, 16), k != 0);
    } else {
      const uint32_t idesc = umma_idesc_bf16(128, ncols) | (1u << 16);
      const uint32_t atom = 64 * 128, grp = 1024;
      for (int k = 0; k < 4; ++k) {
        const uint64_t bd = variant == 0 ? mkdesc(smem_u32(B) + k * 2048, atom, grp) : mkdesc(smem_u32(B) + k * 2048, grp, atom);
        umma_bf16(tbase, mkdesc(smem_u32(A), 16, 1024) + 2 * k, bd, idesc, k != 0);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 256);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 128 * 4);
  static float h[128 * 128];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int variant = 0; variant < 2; ++variant)
    for (int noff : {0, 16, 32, 48}) {
      cudaMemset(d, 0xff, sizeof(h));
      probe<<<1, 128, 60000>>>(d, 0, noff, variant);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("TS variant %d noff %d: %s\n", variant, noff, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 128; ++r) for (int n = 0; n < 16; ++n) bad += h[r * 128 + n] != (float)((r % 4) * 64 + n + noff);
      printf("TS  pack %s noff %2d: mismatches %4d / 2048   row1: %g %g %g %g\n", variant ? "(hi,lo)" : "(lo,hi)", noff, bad,
             h[128], h[129], h[130], h[131]);
    }
  for (int variant = 0; variant < 2; ++variant)
    for (int ncols : {64, 128}) {
      cudaMemset(d, 0xff, sizeof(h));
      probe<<<1, 128, 60000>>>(d, 1, ncols, variant);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("MN variant %d ncols %d: %s\n", variant, ncols, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 128; ++r) for (int n = 0; n < ncols; ++n) bad += h[r * 128 + n] != (float)(((r % 64) * 7 + n) % 256);
      printf("MN  %s N %3d: mismatches %5d / %d   row1: %g %g %g %g  col64: %g %g\n", variant ? "LBO=1024 SBO=atom" : "LBO=atom SBO=1024", ncols,
             bad, 128 * ncols, h[128], h[129], h[130], h[131], h[128 + 64], h[128 + 65]);
    }
  return 0;
}
