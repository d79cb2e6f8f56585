// Implicit-GEMM conv1d / linear core for sm_100a: TMA -> swizzled smem -> tcgen05.mma (accumulators in TMEM)
// -> fused epilogue straight from TMEM.  One kernel template serves every dense op of the path:
//   Linear (1 tap), k=3 FFN convs, dilated HiFi-GAN convs (k = 3/7/11), conv_pre (k = 7), the fused MRF tail
//   (21 taps over 3 source tensors) and the polyphase transposed convs (one tap group per output phase).
//
// GEMM view:  D[128 rows (time) x BN (out channels)] += A[128 x KB] * W[BN x KB]^T  per (tap, K chunk).  The
// activation tile is loaded ONCE per K chunk with its halo (128 + tap span rows, 3-D tensor map (channels, rows,
// batch); TMA zero-fills rows outside [0, rows), which is exactly the conv's zero padding and also isolates
// utterances) and every tap reads it through a descriptor whose start address is moved by (tap shift) rows --
// the swizzle is a function of the absolute smem address, so row-granular starts are legal (verified on B200,
// tools/probes/umma_probe_sw128.cu).  Only the weight slab is re-fetched per tap.  This removes the k-fold re-read
// of activations from L2 that made the k = 7 / 11 convs L2-bandwidth bound.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue
// (thread <-> accumulator row / TMEM lane).  Two smem rings (activation boxes, weight slabs); two TMEM buffers so
// the epilogue of tile i overlaps the MMAs of tile i+1; persistent over tiles.
#pragma once
#include <cuda.h>
#include "srb_ptx.cuh"

namespace srb {

constexpr int kMaxTaps = 32;
constexpr int kMaxGroups = 5;
constexpr int kMaxSrc = 3;
constexpr int kMaxSegs = 8;
constexpr int kTileM = 128;

enum Epilogue : int { EPI_GENERIC = 0, EPI_GLU = 1, EPI_RESNORM = 2, EPI_QKV_ROPE = 3, EPI_EULER = 4 };

struct ConvGemmParams {
  CUtensorMap tmA[kMaxSrc];
  CUtensorMap tmW;
  // problem
  int batch, kchunks, n_groups, n_tiles;          // n_tiles = n_total / BN
  int m_tiles[kMaxGroups];                        // row tiles per batch, per group
  int tile_begin[kMaxGroups + 1];                 // prefix sum of batch*m_tiles*n_tiles
  int group_tap_begin[kMaxGroups + 1];
  int group_rows[kMaxGroups];                     // valid output rows q per batch
  int group_row_add[kMaxGroups];                  // output row = q*row_mul + row_add
  int row_mul;
  short tap_shift[kMaxTaps];
  signed char tap_src[kMaxTaps];
  // taps of a group are split into segments of equal source tensor: one halo-resident A box per (segment, K chunk)
  int group_seg_begin[kMaxGroups + 1];
  short seg_tap_begin[kMaxSegs + 1];
  short seg_min_shift[kMaxSegs];
  signed char seg_src[kMaxSegs];
  int a_box_rows;                                 // 128 + largest tap span of the launch (<= 256)
  int a_box_bytes;                                // a_box_rows * KB * 2, rounded up to 1024
  int a_stages, w_stages;
  // epilogue operands
  const float* bias;
  void* out0;                                     // bf16 "activated"/normalised output
  void* out1;                                     // raw output (bf16 for GENERIC, fp32 for RESNORM/EULER)
  long long out_row_stride, out_batch_stride;     // elements
  const void* res[3];
  long long res_row_stride, res_batch_stride;
  const int* lengths;
  const float* vec0;
  const float* vec1;
  float scale, slope;
  int norm_mode;                                  // RESNORM: 0 none, 1 adaptive (L2), 2 rms
  float f0, f1, f2, f3;                           // EULER: dt, std, mean, pad
  void* aux0;                                     // EULER: mel fp32 (or null)
  void* aux1;                                     // EULER: mel bf16
};

template <int BN>
struct TmemCols {
  static constexpr int buf = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int total = 2 * buf;
};

template <int BN, int KB>
struct StageLayout {
  static constexpr int a_bytes = kTileM * KB * 2;
  static constexpr int w_bytes_raw = BN * KB * 2;
  static constexpr int w_bytes = (w_bytes_raw + 1023) & ~1023;
  static constexpr int stage_bytes = a_bytes + w_bytes;
};

struct TileCoord {
  int group, b, m, n;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvGemmParams& p, int tile) {
  TileCoord c;
  int g = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroups; ++i)
    if (i < p.n_groups && tile >= p.tile_begin[i]) g = i;
  int local = tile - p.tile_begin[g];
  c.group = g;
  c.n = local % p.n_tiles;
  int rest = local / p.n_tiles;
  c.m = rest % p.m_tiles[g];
  c.b = rest / p.m_tiles[g];
  return c;
}

// ---------------------------------------------------------------------------------------------- epilogues
// Thread <-> one accumulator row (TMEM lane).  The epilogue is the throughput limiter of these kernels when it is
// under-populated (one warp per SM sub-partition cannot hide TMEM / global latency), so wide tiles use EIGHT
// epilogue warps: warps sharing a TMEM lane quarter split the columns in two halves (`half`, `nhalf`).
// `tacc` is the TMEM address of (lane base, column 0) of this tile's accumulator buffer.  All 32 lanes of a warp
// execute every tcgen05.ld/st (sync.aligned); global loads that do not depend on the accumulator are issued
// before the TMEM wait so their latency overlaps it.

template <int BN, int NHALF>
__device__ __forceinline__ void epi_generic(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q, int half) {
  constexpr int CW = BN < 32 ? BN : 32;
  constexpr int COLS = BN / NHALF;
  const bool valid = q < p.group_rows[tc.group];
  const long long orow = (long long)q * p.row_mul + p.group_row_add[tc.group];
  const long long obase = (long long)tc.b * p.out_batch_stride + orow * p.out_row_stride + (long long)tc.n * BN;
  const long long rbase = (long long)tc.b * p.res_batch_stride + orow * p.res_row_stride + (long long)tc.n * BN;
  const float sc = p.scale, sl = p.slope;
#pragma unroll 1
  for (int c0 = half * COLS; c0 < (half + 1) * COLS; c0 += CW) {
    uint4 rv[3][CW / 8];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (valid && p.res[r]) {
        const uint4* r4 = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.res[r]) + rbase + c0);
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) rv[r][j] = __ldg(r4 + j);
      }
    }
    float y[CW];
    tmem_ld_f<CW>(tacc + c0, y);
    if (valid) {
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + tc.n * BN + c0);
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          float4 bb = __ldg(b4 + j);
          y[4 * j + 0] += bb.x; y[4 * j + 1] += bb.y; y[4 * j + 2] += bb.z; y[4 * j + 3] += bb.w;
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (p.res[r]) {
#pragma unroll
          for (int j = 0; j < CW / 8; ++j) {
            const uint4 u = rv[r][j];
            y[8 * j + 0] += bf16_lo(u.x); y[8 * j + 1] += bf16_hi(u.x);
            y[8 * j + 2] += bf16_lo(u.y); y[8 * j + 3] += bf16_hi(u.y);
            y[8 * j + 4] += bf16_lo(u.z); y[8 * j + 5] += bf16_hi(u.z);
            y[8 * j + 6] += bf16_lo(u.w); y[8 * j + 7] += bf16_hi(u.w);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CW; ++j) y[j] *= sc;
      if (p.out1) {
        uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out1) + obase + c0);
#pragma unroll
        for (int j = 0; j < CW / 8; ++j)
          o4[j] = make_uint4(pack_bf16(y[8 * j], y[8 * j + 1]), pack_bf16(y[8 * j + 2], y[8 * j + 3]),
                             pack_bf16(y[8 * j + 4], y[8 * j + 5]), pack_bf16(y[8 * j + 6], y[8 * j + 7]));
      }
      if (p.out0) {
        uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out0) + obase + c0);
#pragma unroll
        for (int j = 0; j < CW / 8; ++j)
          o4[j] = make_uint4(pack_bf16(lrelu(y[8 * j], sl), lrelu(y[8 * j + 1], sl)),
                             pack_bf16(lrelu(y[8 * j + 2], sl), lrelu(y[8 * j + 3], sl)),
                             pack_bf16(lrelu(y[8 * j + 4], sl), lrelu(y[8 * j + 5], sl)),
                             pack_bf16(lrelu(y[8 * j + 6], sl), lrelu(y[8 * j + 7], sl)));
      }
    }
  }
}

// FFN conv1 tile = [128 value columns | 128 gate columns] -> 128 outputs  (fastspeech/modules.py:27-30, 62-69)
template <int NHALF>
__device__ __forceinline__ void epi_glu(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q, int half) {
  constexpr int COLS = 128 / NHALF;
  const bool valid = q < p.group_rows[0];
  const bool keep = valid && q < p.lengths[tc.b];
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out0) + (long long)tc.b * p.out_batch_stride +
                       (long long)q * p.out_row_stride + tc.n * 128;
  const float* bias = p.bias + tc.n * 256;
#pragma unroll 1
  for (int c0 = half * COLS; c0 < (half + 1) * COLS; c0 += 32) {
    uint32_t v[32], g[32];
    tmem_ld32(tacc + c0, v);
    tmem_ld32(tacc + 128 + c0, g);
    tmem_ld_wait();
    if (valid) {
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + c0) + j);
        const float4 bg = __ldg(reinterpret_cast<const float4*>(bias + 128 + c0) + j);
        const float h0 = silu_fast(__uint_as_float(g[4 * j + 0]) + bg.x) * (__uint_as_float(v[4 * j + 0]) + bv.x);
        const float h1 = silu_fast(__uint_as_float(g[4 * j + 1]) + bg.y) * (__uint_as_float(v[4 * j + 1]) + bv.y);
        const float h2 = silu_fast(__uint_as_float(g[4 * j + 2]) + bg.z) * (__uint_as_float(v[4 * j + 2]) + bv.z);
        const float h3 = silu_fast(__uint_as_float(g[4 * j + 3]) + bg.w) * (__uint_as_float(v[4 * j + 3]) + bv.w);
        o[2 * j] = keep ? pack_bf16(h0, h1) : 0u;       // pads are zeroed before conv2 (fastspeech/modules.py:69)
        o[2 * j + 1] = keep ? pack_bf16(h2, h3) : 0u;
      }
      uint4* o4 = reinterpret_cast<uint4*>(out + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) o4[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
  }
}

// y = acc + bias + residual (fp32 stream, in place allowed); x_out = y; xn = bf16(norm(y) * g) with pad rows zeroed.
// A row's 256 columns live in ONE TMEM lane; with two column halves per lane quarter the two warps exchange their
// partial sums of squares through `red` (smem, [2][128] floats) and a 64-thread named barrier.
// 64-thread named barrier shared by the two epilogue warps of one TMEM lane quarter (ids 1-4, immediates so that
// ptxas accounts for them; id 0 is __syncthreads)
__device__ __forceinline__ void pair_barrier(int quarter) {
  switch (quarter) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

// Global operands of an epilogue that do not depend on the accumulator (the fp32 residual stream, the rotary
// tables) are fetched in two batches: the first BEFORE the thread waits for the tile's MMAs (so its L2 latency
// overlaps the tensor work), the second right after the wait while the first is being consumed.  Profiling showed
// these epilogues stalled on exactly these loads (long-scoreboard on the first dependent FADD/FMUL).
template <int NHALF>
__device__ __forceinline__ void resnorm_prefetch(const ConvGemmParams& p, const TileCoord& tc, int q, int half, float4 (&pre)[16]) {
  constexpr int COLS = 256 / NHALF;
  if (q < p.group_rows[0]) {
    const float* res = static_cast<const float*>(p.res[0]) + (long long)tc.b * p.res_batch_stride +
                       (long long)q * p.res_row_stride + half * COLS;
#pragma unroll
    for (int j = 0; j < 16; ++j) pre[j] = *reinterpret_cast<const float4*>(res + 4 * j);
  }
}

template <int NHALF>
__device__ __forceinline__ void epi_resnorm(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q, int half,
                                            float* red, int row_in_tile, int quarter, const float4 (&pre)[16]) {
  constexpr int COLS = 256 / NHALF;
  constexpr int NCHUNK = COLS / 32;
  static_assert(NCHUNK == 4, "RESNORM runs with eight epilogue warps");
  const bool valid = q < p.group_rows[0];
  const long long off = (long long)tc.b * p.out_batch_stride + (long long)q * p.out_row_stride;
  const float* res = static_cast<const float*>(p.res[0]) + (long long)tc.b * p.res_batch_stride +
                     (long long)q * p.res_row_stride + half * COLS;
  float* xout = static_cast<float*>(p.out1) + off + half * COLS;
  const uint32_t tcol = tacc + half * COLS;
  float4 late[16];   // residual of chunks 2, 3
  if (valid) {
#pragma unroll
    for (int j = 0; j < 16; ++j) late[j] = *reinterpret_cast<const float4*>(res + 64 + 4 * j);
  }
  float sumsq = 0.f;
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    uint32_t v[32];
    tmem_ld32(tcol + c * 32, v);
    tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 r = c < 2 ? pre[(c & 1) * 8 + j] : late[(c & 1) * 8 + j];
        float4 y;
        y.x = __uint_as_float(v[4 * j + 0]) + r.x;
        y.y = __uint_as_float(v[4 * j + 1]) + r.y;
        y.z = __uint_as_float(v[4 * j + 2]) + r.z;
        y.w = __uint_as_float(v[4 * j + 3]) + r.w;
        if (p.bias) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + half * COLS + c * 32 + 4 * j));
          y.x += bb.x; y.y += bb.y; y.z += bb.z; y.w += bb.w;
        }
        sumsq += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
        *reinterpret_cast<float4*>(xout + c * 32 + 4 * j) = y;
        v[4 * j + 0] = __float_as_uint(y.x); v[4 * j + 1] = __float_as_uint(y.y);
        v[4 * j + 2] = __float_as_uint(y.z); v[4 * j + 3] = __float_as_uint(y.w);
      }
    }
    if (p.norm_mode != 0) tmem_st32(tcol + c * 32, v);  // stash y for the second pass
  }
  if (p.norm_mode == 0) return;
  tmem_st_wait();
  if constexpr (NHALF == 2) {
    red[half * 128 + row_in_tile] = sumsq;
    pair_barrier(quarter);
    sumsq += red[(half ^ 1) * 128 + row_in_tile];
  }
  float inv;
  if (p.norm_mode == 1) inv = 1.f / fmaxf(sqrtf(sumsq), 1e-12f);           // F.normalize (norm.py:41)
  else inv = rsqrtf(sumsq * (1.f / 256.f) + 1.1920928955078125e-07f);      // nn.RMSNorm eps = finfo(fp32).eps
  const bool keep = valid && (p.lengths == nullptr || q < p.lengths[tc.b]);
  __nv_bfloat16* xn = static_cast<__nv_bfloat16*>(p.out0) + off + half * COLS;
  const float* gv = p.vec0 + half * COLS;
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    uint32_t v[32];
    tmem_ld32(tcol + c * 32, v);
    tmem_ld_wait();
    if (valid) {
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gv + c * 32 + 4 * j));
        // select (not multiply) so a non-finite pad row can never leak into the conv taps of valid frames
        o[2 * j] = keep ? pack_bf16(__uint_as_float(v[4 * j]) * inv * g.x, __uint_as_float(v[4 * j + 1]) * inv * g.y) : 0u;
        o[2 * j + 1] = keep ? pack_bf16(__uint_as_float(v[4 * j + 2]) * inv * g.z, __uint_as_float(v[4 * j + 3]) * inv * g.w) : 0u;
      }
      uint4* o4 = reinterpret_cast<uint4*>(xn + c * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) o4[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
  }
  if constexpr (NHALF == 2) {
    // the pair must not overwrite `red` for the next tile before both have read it
    pair_barrier(quarter);
  }
}

// to_qkv tile n: 0 = q (2 heads x 128), 1 = k, 2 = v.  Rotary (transformer.py:66-73) on q,k:
// out[i] = t[i] cos - t[i+64] sin ; out[i+64] = t[i+64] cos + t[i] sin, angle = pos * inv_freq[i], i < 64.
// With eight epilogue warps, column half h of a lane quarter is head h.
__device__ __forceinline__ void rope_prefetch(const ConvGemmParams& p, const TileCoord& tc, int q, float4 (&pre)[16]) {
  if (tc.n != 2 && q < p.group_rows[0]) {
    const float* cs = p.vec0 + (long long)q * 64;
    const float* sn = p.vec1 + (long long)q * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      pre[j] = __ldg(reinterpret_cast<const float4*>(cs) + j);        // cos, frequencies 0..31
      pre[8 + j] = __ldg(reinterpret_cast<const float4*>(sn) + j);    // sin, frequencies 0..31
    }
  }
}

template <int NHALF>
__device__ __forceinline__ void epi_qkv_rope(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q, int half,
                                             const float4 (&pre)[16]) {
  static_assert(NHALF == 2, "QKV_ROPE runs with eight epilogue warps (one head per column half)");
  const bool valid = q < p.group_rows[0];
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out0) + (long long)tc.b * p.out_batch_stride +
                       (long long)q * p.out_row_stride + tc.n * 256 + half * 128;
  const uint32_t tcol = tacc + half * 128;
  if (tc.n == 2) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tcol + c * 32, v);
      tmem_ld_wait();
      if (valid) {
        uint4* o4 = reinterpret_cast<uint4*>(out + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o4[j] = make_uint4(pack_bf16(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                             pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                             pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                             pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
      }
    }
    return;
  }
  float4 late[16];   // cos / sin of frequencies 32..63
  if (valid) {
    const float* cs = p.vec0 + (long long)q * 64 + 32;
    const float* sn = p.vec1 + (long long)q * 64 + 32;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      late[j] = __ldg(reinterpret_cast<const float4*>(cs) + j);
      late[8 + j] = __ldg(reinterpret_cast<const float4*>(sn) + j);
    }
  }
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    uint32_t lo[32], hi[32];
    tmem_ld32(tcol + f * 32, lo);
    tmem_ld32(tcol + 64 + f * 32, hi);
    tmem_ld_wait();
    if (valid) {
      uint32_t olo[16], ohi[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 c = f == 0 ? pre[j] : late[j];
        const float4 s = f == 0 ? pre[8 + j] : late[8 + j];
        const float a0 = __uint_as_float(lo[4 * j]), a1 = __uint_as_float(lo[4 * j + 1]);
        const float a2 = __uint_as_float(lo[4 * j + 2]), a3 = __uint_as_float(lo[4 * j + 3]);
        const float b0 = __uint_as_float(hi[4 * j]), b1 = __uint_as_float(hi[4 * j + 1]);
        const float b2 = __uint_as_float(hi[4 * j + 2]), b3 = __uint_as_float(hi[4 * j + 3]);
        olo[2 * j] = pack_bf16(a0 * c.x - b0 * s.x, a1 * c.y - b1 * s.y);
        olo[2 * j + 1] = pack_bf16(a2 * c.z - b2 * s.z, a3 * c.w - b3 * s.w);
        ohi[2 * j] = pack_bf16(b0 * c.x + a0 * s.x, b1 * c.y + a1 * s.y);
        ohi[2 * j + 1] = pack_bf16(b2 * c.z + a2 * s.z, b3 * c.w + a3 * s.w);
      }
      uint4* l4 = reinterpret_cast<uint4*>(out + f * 32);
      uint4* h4 = reinterpret_cast<uint4*>(out + 64 + f * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        l4[j] = make_uint4(olo[4 * j], olo[4 * j + 1], olo[4 * j + 2], olo[4 * j + 3]);
        h4[j] = make_uint4(ohi[4 * j], ohi[4 * j + 1], ohi[4 * j + 2], ohi[4 * j + 3]);
      }
    }
  }
}

template <int N>
__device__ __forceinline__ void euler_chunk(const uint32_t (&v)[N], int c0, float* xt, __nv_bfloat16* xtb, float* mel,
                                            __nv_bfloat16* melb, long long off, bool is_pad, float dt, float sd,
                                            float mean, float padv) {
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    float4 x = *reinterpret_cast<const float4*>(xt + c0 + j);
    x.x += dt * __uint_as_float(v[j]);
    x.y += dt * __uint_as_float(v[j + 1]);
    x.z += dt * __uint_as_float(v[j + 2]);
    x.w += dt * __uint_as_float(v[j + 3]);
    *reinterpret_cast<float4*>(xt + c0 + j) = x;
    *reinterpret_cast<uint2*>(xtb + c0 + j) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
    if (mel) {
      float4 m;
      m.x = is_pad ? padv : x.x * sd + mean;
      m.y = is_pad ? padv : x.y * sd + mean;
      m.z = is_pad ? padv : x.z * sd + mean;
      m.w = is_pad ? padv : x.w * sd + mean;
      *reinterpret_cast<float4*>(mel + off + c0 + j) = m;
      *reinterpret_cast<uint2*>(melb + off + c0 + j) = make_uint2(pack_bf16(m.x, m.y), pack_bf16(m.z, m.w));
    }
  }
}

// to_pred (N = 80) + Euler step in fp32 (models.py:183-184); last step also de-normalises and fills pads (:186-187)
__device__ __forceinline__ void epi_euler(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q) {
  const bool valid = q < p.group_rows[0];
  const long long off = (long long)tc.b * p.out_batch_stride + (long long)q * p.out_row_stride;
  float* xt = static_cast<float*>(p.out1) + off;
  __nv_bfloat16* xtb = static_cast<__nv_bfloat16*>(p.out0) + off;
  float* mel = static_cast<float*>(p.aux0);
  __nv_bfloat16* melb = static_cast<__nv_bfloat16*>(p.aux1);
  const bool is_pad = valid && p.lengths != nullptr && q >= p.lengths[tc.b];
  const float dt = p.f0, sd = p.f1, mean = p.f2, padv = p.f3;
  {
    uint32_t v[32];
    tmem_ld32(tacc, v);
    tmem_ld_wait();
    if (valid) euler_chunk<32>(v, 0, xt, xtb, mel, melb, off, is_pad, dt, sd, mean, padv);
    tmem_ld32(tacc + 32, v);
    tmem_ld_wait();
    if (valid) euler_chunk<32>(v, 32, xt, xtb, mel, melb, off, is_pad, dt, sd, mean, padv);
  }
  {
    uint32_t v[16];
    tmem_ld16(tacc + 64, v);
    tmem_ld_wait();
    if (valid) euler_chunk<16>(v, 64, xt, xtb, mel, melb, off, is_pad, dt, sd, mean, padv);
  }
}

// ---------------------------------------------------------------------------------------------- kernel
template <int BN, int EPI>
struct EpiWarps {
  // eight epilogue warps (two column halves per TMEM lane quarter) for the wide tiles, four otherwise
  static constexpr int value = (BN >= 128 && EPI != EPI_EULER) ? 8 : 4;
};

template <int BN, int KB, int EPI>
__global__ void __launch_bounds__(64 + 32 * EpiWarps<BN, EPI>::value)
convgemm_kernel(const __grid_constant__ ConvGemmParams p, int total_tiles) {
  using L = StageLayout<BN, KB>;
  constexpr int SW = KB * 2;
  constexpr int TBUF = TmemCols<BN>::buf;
  constexpr int TCOLS = TmemCols<BN>::total;
  constexpr int EW = EpiWarps<BN, EPI>::value;
  constexpr int NHALF = EW / 4;
  constexpr uint32_t IDESC = umma_idesc_bf16(kTileM, BN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int a_stages = p.a_stages, w_stages = p.w_stages;
  const uint32_t a_ring = smem_base;
  const uint32_t w_ring = smem_base + a_stages * p.a_box_bytes;
  const uint32_t bar_base = w_ring + w_stages * L::w_bytes;       // 1024-aligned
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (a_stages + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * a_stages + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * a_stages + w_stages + s); };
  const int n_ring_bars = 2 * (a_stages + w_stages);
  auto tfull_bar = [&](int b) { return bar_base + 8u * (n_ring_bars + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (n_ring_bars + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (n_ring_bars + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* red = reinterpret_cast<float*>(smem_gen + (bar_base - smem_base) + 8 * (n_ring_bars + 4) + 16);  // [2][128]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform (uniform datapath)
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a_stages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < w_stages; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), EW);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmA[0]);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    // TMA producer: converged warp, one elected lane issues (coordinates stay in uniform registers)
    {
      int ai = 0, wi = 0;
      uint32_t aph = 0, wph = 0;
      const uint32_t a_tx = static_cast<uint32_t>(p.a_box_rows) * KB * 2;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int t0 = tc.m * kTileM;
        for (int sg = p.group_seg_begin[tc.group]; sg < p.group_seg_begin[tc.group + 1]; ++sg) {
          const int src = p.seg_src[sg];
          const int row = t0 + p.seg_min_shift[sg];
          const int tb = p.seg_tap_begin[sg], te = p.seg_tap_begin[sg + 1];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(a_empty(ai), aph ^ 1u);
            mbar_expect_tx_elect(a_full(ai), a_tx);
            tma_load_3d_elect(a_ring + ai * p.a_box_bytes, &p.tmA[src], a_full(ai), kc * KB, row, tc.b);
            if (++ai == a_stages) { ai = 0; aph ^= 1u; }
            for (int t = tb; t < te; ++t) {
              mbar_wait(w_empty(wi), wph ^ 1u);
              mbar_expect_tx_elect(w_full(wi), L::w_bytes_raw);
              tma_load_2d_elect(w_ring + wi * L::w_bytes, &p.tmW, w_full(wi), (t * p.kchunks + kc) * KB, tc.n * BN);
              if (++wi == w_stages) { wi = 0; wph ^= 1u; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // MMA issuer: the whole warp runs the (warp-uniform) loop converged so descriptors stay in uniform registers;
    // one elected lane issues.  See umma_bf16_pred.
    {
      int ai = 0, wi = 0;
      uint32_t aph = 0, wph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const TileCoord tc = decode_tile(p, tile);
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(buf), bphase ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * TBUF;
        uint32_t first = 1;
        for (int sg = p.group_seg_begin[tc.group]; sg < p.group_seg_begin[tc.group + 1]; ++sg) {
          const int min_shift = p.seg_min_shift[sg];
          const int tb = p.seg_tap_begin[sg], te = p.seg_tap_begin[sg + 1];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(a_full(ai), aph);
            const uint32_t a_addr = a_ring + ai * p.a_box_bytes;
            for (int t = tb; t < te; ++t) {
              mbar_wait(w_full(wi), wph);
              tc_fence_after();
              const uint64_t adesc = umma_smem_desc<SW>(a_addr + (p.tap_shift[t] - min_shift) * (KB * 2));
              const uint64_t wdesc = umma_smem_desc<SW>(w_ring + wi * L::w_bytes);
#pragma unroll
              for (int k = 0; k < KB / 16; ++k)
                umma_bf16_pred(1u, tacc, adesc + 2 * k, wdesc + 2 * k, IDESC, (first && k == 0) ? 0u : 1u);
              first = 0;
              umma_commit_pred(1u, w_empty(wi));
              if (++wi == w_stages) { wi = 0; wph ^= 1u; }
            }
            umma_commit_pred(1u, a_empty(ai));
            if (++ai == a_stages) { ai = 0; aph ^= 1u; }
          }
        }
        umma_commit_pred(1u, tfull_bar(buf));
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;           // column half (0 when EW == 4)
    const int lane_base = quarter * 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const TileCoord tc = decode_tile(p, tile);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const int q = tc.m * kTileM + lane_base + lane;
      float4 pre[16];
      if constexpr (EPI == EPI_RESNORM) resnorm_prefetch<NHALF>(p, tc, q, half, pre);
      if constexpr (EPI == EPI_QKV_ROPE) rope_prefetch(p, tc, q, pre);
      mbar_wait(tfull_bar(buf), bphase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(lane_base) << 16) + buf * TBUF;
      if constexpr (EPI == EPI_GENERIC) epi_generic<BN, NHALF>(p, tacc, tc, q, half);
      else if constexpr (EPI == EPI_GLU) epi_glu<NHALF>(p, tacc, tc, q, half);
      else if constexpr (EPI == EPI_RESNORM) epi_resnorm<NHALF>(p, tacc, tc, q, half, red, lane_base + lane, quarter, pre);
      else if constexpr (EPI == EPI_QKV_ROPE) epi_qkv_rope<NHALF>(p, tacc, tc, q, half, pre);
      else epi_euler(p, tacc, tc, q);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TCOLS);
}

}  // namespace srb
