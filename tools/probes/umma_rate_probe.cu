// How long does one tcgen05.mma (M = 128, K = 16, bf16) take as a function of N?  (A in un-swizzled interleaved smem
// layout as in srb_mrf_fused.cu, B likewise.)  Issues `iters` MMAs back to back from one thread and times them with
// clock64 around a final commit + mbarrier wait.   ./umma_rate_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../speech_resynth_b200/csrc/srb_ptx.cuh"
using namespace srb;

__device__ __forceinline__ uint64_t desc_interleaved(uint32_t addr, uint32_t chunk_stride) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((chunk_stride >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(128 >> 4) << 32;
  d |= 1ull << 46;
  return d;
}

template <int N>
__global__ void probe(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t s_a = sbase, s_b = sbase + 64 * 1024, bar = sbase + 96 * 1024, slot = bar + 16;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 96 * 1024 + 16);
  if (threadIdx.x < 32) {
    constexpr uint32_t IDESC = umma_idesc_bf16(128, N);
    const uint64_t a0 = desc_interleaved(s_a, 2048 * 16), b0 = desc_interleaved(s_b, N * 16);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // rotate over 8 accumulators and shifted A views like a conv does
      umma_bf16_pred(1u, tmem + (i & 7) * N, a0 + (uint64_t)(i & 31), b0, IDESC, 1u);
    }
    umma_commit_pred(1u, bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N>
void run(long long* d) {
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 20000;
  probe<N><<<1, 64, 100 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d: %s  %.1f clk per MMA (1 SM alone)\n", N, cudaGetErrorString(e), (double)h / iters);
  // all SMs at once (power / clocks as in a real kernel)
  probe<N><<<148, 64, 100 * 1024>>>(iters, d);
  e = cudaDeviceSynchronize();
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d: %s  %.1f clk per MMA (148 SMs)\n", N, cudaGetErrorString(e), (double)h / iters);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<16>(d);
  run<32>(d);
  run<64>(d);
  run<128>(d);
  run<256>(d);
  return 0;
}
