// Probe (B200): tensor-memory read throughput of tcgen05.ld (how many bytes per clock can the softmax warps pull, and
// does it depend on warps per SM sub-partition / loads in flight?), alone and while the tensor core runs SS-form or
// TS-form (A operand from tensor memory) MMAs.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include "../../speech_resynth_b200/csrc/srb_ptx.cuh"
using namespace srb;

__device__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

// warps 0 .. nld-1: tcgen05.ld loops (depth 1 or 2 in flight); warp 16: MMA issuer (mode 0 none, 1 SS, 2 TS)
__global__ void __launch_bounds__(544) probe(long long* out, int nld, int depth, int mma_mode, int reps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  long long t0 = clock64();
  if (warp < nld) {
    const uint32_t addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    if (depth == 1) {
      for (int r = 0; r < reps; ++r) {
        uint32_t v[32];
        tmem_ld32(addr + (r & 7) * 32, v);
        tmem_ld_wait_dep(v);
        acc ^= v[r & 31];
      }
    } else {
      for (int r = 0; r < reps; r += 2) {
        uint32_t v[32], w[32];
        tmem_ld32(addr + (r & 7) * 32, v);
        tmem_ld32(addr + ((r + 1) & 7) * 32, w);
        tmem_ld_wait_dep(v);
        acc ^= v[r & 31] ^ w[r & 31];
      }
    }
    long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    if (acc == 0x12345678u) out[63] = acc;
  } else if (warp == 16 && mma_mode != 0) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const int n_mma = reps * 2;
      for (int k = 0; k < n_mma; ++k) {
        if (mma_mode == 1)
          umma_bf16(tbase + 256, umma_smem_desc<128>(smem_u32(smem)), umma_smem_desc<128>(smem_u32(smem) + 16384), idesc, 1);
        else
          umma_ts(tbase + 256, tbase + 384 + (k & 7) * 8, umma_smem_desc<128>(smem_u32(smem) + 16384), idesc, 1);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      long long t1 = clock64();
      out[32] = t1 - t0;
      out[33] = n_mma;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  long long h[64];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const int reps = 2000;
  for (int mma = 0; mma < 3; ++mma)
    for (int nld : {0, 4, 8, 16})
      for (int depth : {1, 2}) {
        if (nld == 0 && (depth == 2 || mma == 0)) continue;
        cudaMemset(d, 0, sizeof(h));
        probe<<<1, 544, 70000>>>(d, nld, depth, mma, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma %d nld %d depth %d: %s\n", mma, nld, depth, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < nld; ++w) mx = h[w] > mx ? h[w] : mx;
        double bpc = nld ? (double)nld * reps * 4096.0 / (double)mx : 0.0;
        printf("mma %s  ld warps %2d depth %d: ld %7lld clk  -> %6.1f B/clk/SM (%5.1f clk per x32 load per warp)   mma: %lld clk for %lld MMAs = %.1f clk each\n",
               mma == 0 ? "none" : (mma == 1 ? "SS  " : "TS  "), nld, depth, mx, bpc, nld ? (double)mx / reps : 0.0, h[32], h[33],
               h[33] ? (double)h[32] / h[33] : 0.0);
      }
  return 0;
}
