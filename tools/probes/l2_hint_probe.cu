// Which L2 cache-hint forms run on sm_100a?  (each kernel separately; an illegal instruction poisons the context,
// so the host runs every probe in a fresh process: ./l2_hint_probe <n>)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pol_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__global__ void k_policy(uint64_t* out) { out[0] = pol_last(); }
__global__ void k_st_hint(uint4* dst) {
  const uint64_t p = pol_last();
  uint4 v = make_uint4(1, 2, 3, 4);
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(dst + threadIdx.x), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(p) : "memory");
}
__global__ void k_cpasync_hint_srcsize(const uint4* src, uint4* dst) {
  __shared__ uint4 buf[32];
  const uint64_t p = pol_last();
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(buf + threadIdx.x);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(d), "l"(src + threadIdx.x), "r"(16u), "l"(p) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  dst[threadIdx.x] = buf[threadIdx.x];
}
__global__ void k_cpasync_hint(const uint4* src, uint4* dst) {
  __shared__ uint4 buf[32];
  const uint64_t p = pol_last();
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(buf + threadIdx.x);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(src + threadIdx.x), "l"(p) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  dst[threadIdx.x] = buf[threadIdx.x];
}
__global__ void k_ld_hint(const uint4* src, uint4* dst) {
  const uint64_t p = pol_last();
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.b32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + threadIdx.x), "l"(p));
  dst[threadIdx.x] = v;
}
int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  uint4 *a, *b;
  cudaMalloc(&a, 4096);
  cudaMalloc(&b, 4096);
  cudaMemset(a, 1, 4096);
  switch (which) {
    case 0: k_policy<<<1, 1>>>((uint64_t*)b); break;
    case 1: k_st_hint<<<1, 32>>>(b); break;
    case 2: k_cpasync_hint_srcsize<<<1, 32>>>(a, b); break;
    case 3: k_cpasync_hint<<<1, 32>>>(a, b); break;
    case 4: k_ld_hint<<<1, 32>>>(a, b); break;
  }
  cudaError_t e = cudaDeviceSynchronize();
  uint64_t h = 0;
  cudaMemcpy(&h, b, 8, cudaMemcpyDeviceToHost);
  printf("probe %d: %s (first word %llx)\n", which, cudaGetErrorString(e), (unsigned long long)h);
  return 0;
}
