// Key-padding-masked softmax attention for the velocity transformer (transformer.py:115-127):
// 2 heads x d_head 128, scale 1/sqrt(128), keys >= len_b masked, no dropout.  Flash-style single pass with an
// online softmax; the (B, H, N, N) boolean mask of the reference is replaced by the per-utterance length.
//
// Round-1 implementation: warp-level mma.sync (m16n8k16 bf16 -> fp32), cp.async double-buffered K/V tiles with an
// XOR swizzle, 64 query rows per CTA (16 per warp).  The tcgen05/TMEM version is the next step.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/srb.h"
#include "srb_common.h"
#include "srb_ptx.cuh"

namespace srb {

constexpr int kHeadDim = 128;
constexpr int kQRows = 64;
constexpr int kKVRows = 64;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tile of `rows` x 128 bf16 (256 B per row = 16 chunks of 16 B); chunk index XOR (row & 7)
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 256 + ((chunk ^ (row & 7)) << 4);
}

__device__ __forceinline__ void load_tile(uint32_t smem, const __nv_bfloat16* g, long long row_stride, int row0,
                                          int rows_total, int nrows, int tid, int nthreads) {
  for (int i = tid; i < nrows * 16; i += nthreads) {
    const int r = i >> 4, c = i & 15;
    const int gr = row0 + r;
    const bool ok = gr < rows_total;
    cp_async16(tile_addr(smem, r, c), g + (long long)(ok ? gr : 0) * row_stride + c * 8, ok);
  }
}

__global__ void __launch_bounds__(128) attention_kernel(const __nv_bfloat16* __restrict__ qkv, const int* __restrict__ lengths,
                                                        __nv_bfloat16* __restrict__ out, int frames) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQRows;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int len = min(lengths[b], frames);
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + kQRows * 256;            // 2 buffers
  const uint32_t sV = sK + 2 * kKVRows * 256;       // 2 buffers
  const __nv_bfloat16* base = qkv + (long long)b * frames * 768;
  const __nv_bfloat16* gq = base + h * kHeadDim;
  const __nv_bfloat16* gk = base + 256 + h * kHeadDim;
  const __nv_bfloat16* gv = base + 512 + h * kHeadDim;

  const int n_kv = (len + kKVRows - 1) / kKVRows;
  load_tile(sQ, gq, 768, q0, frames, kQRows, tid, 128);
  if (n_kv > 0) {
    load_tile(sK, gk, 768, 0, len, kKVRows, tid, 128);
    load_tile(sV, gv, 768, 0, len, kKVRows, tid, 128);
  }
  cp_async_commit();

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float sl2 = 0.08838834764831845f * 1.4426950408889634f;  // (1/sqrt(128)) * log2(e)

  uint32_t qf[8][4];
  for (int kv = 0; kv < n_kv; ++kv) {
    const int buf = kv & 1;
    if (kv + 1 < n_kv) {
      load_tile(sK + (buf ^ 1) * kKVRows * 256, gk, 768, (kv + 1) * kKVRows, len, kKVRows, tid, 128);
      load_tile(sV + (buf ^ 1) * kKVRows * 256, gv, 768, (kv + 1) * kKVRows, len, kKVRows, tid, 128);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kv == 0) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        ldmatrix_x4(tile_addr(sQ, warp * 16 + (lane & 15), kk * 2 + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
    }
    const uint32_t kb = sK + buf * kKVRows * 256, vb = sV + buf * kKVRows * 256;

    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(tile_addr(kb, j * 8 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
        mma_bf16(s[j], qf[kk], b0, b1);
        mma_bf16(s[j + 1], qf[kk], b2, b3);
      }
    }
    // mask keys >= len (only the last tile can be partial), scale into log2 domain
    const int key0 = kv * kKVRows + 2 * (lane & 3);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = key0 + j * 8 + (e & 1);
        const float v = key < len ? s[j][e] * sl2 : -INFINITY;
        s[j][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);   // finite: every processed tile has >= 1 valid key
      corr[r] = exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] - m_run[0]), p1 = exp2f(s[j][1] - m_run[0]);
      const float p2 = exp2f(s[j][2] - m_run[1]), p3 = exp2f(s[j][3] - m_run[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(tile_addr(vb, kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), j + (lane >> 4)), b0, b1, b2, b3);
        mma_bf16(o[j], pf[kk], b0, b1);
        mma_bf16(o[j + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();  // all warps done with this K/V buffer before it is refilled
  }
  if (n_kv == 0) cp_async_wait<0>();

  // finalise: divide by the row sums (quad reduction) and store bf16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  const int row0 = q0 + warp * 16 + (lane >> 2);
  __nv_bfloat16* ob = out + (long long)b * frames * 256 + h * kHeadDim + 2 * (lane & 3);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (row0 < frames)
      *reinterpret_cast<uint32_t*>(ob + (long long)row0 * 256 + i * 8) = pack_bf16(o[i][0] * inv0, o[i][1] * inv0);
    if (row0 + 8 < frames)
      *reinterpret_cast<uint32_t*>(ob + (long long)(row0 + 8) * 256 + i * 8) = pack_bf16(o[i][2] * inv1, o[i][3] * inv1);
  }
}

}  // namespace srb

using namespace srb;

extern "C" int srb_cfm_attention(const void* qkv_bf16, const int32_t* lengths, void* o_bf16, int32_t batch,
                                 int32_t frames, void* stream) {
  if (batch <= 0 || frames <= 0) return 0;
  const int smem = (kQRows + 4 * kKVRows) * 256;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    SRB_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = true;
  }
  dim3 grid((frames + kQRows - 1) / kQRows, 2, batch);
  attention_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16*>(qkv_bf16), lengths,
                                                              static_cast<__nv_bfloat16*>(o_bf16), frames);
  return after_launch("attention_kernel");
}
