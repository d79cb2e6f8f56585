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

#include <type_traits>

#include "srb_ptx.cuh"

namespace srb {

constexpr int kMaxTaps = 32;
constexpr int kMaxGroups = 5;
constexpr int kMaxSrc = 3;
constexpr int kMaxSegs = 8;
constexpr int kTileM = 128;

enum Epilogue : int { EPI_GENERIC = 0, EPI_GLU = 1, EPI_RESNORM = 2, EPI_QKV_ROPE = 3, EPI_EULER = 4, EPI_ARGMAX = 5 };

struct ConvGemmParams {
  CUtensorMap tmA[kMaxSrc];
  CUtensorMap tmW;
  CUtensorMap tmWh;                               // weight map with a half-height box (2-CTA multicast variant)
  // epilogue I/O through TMA (RESNORM, QKV_ROPE): 3-D maps (columns, rows, batch) with a 32-column x 32-row box --
  // one epilogue warp's block.  tmR: fp32 residual, tmO1: fp32 output stream, tmO0: bf16 output.
  CUtensorMap tmR, tmO1, tmO0;
  // QKV_ROPE with a transposed V: 3-D view (frames, batch, 256) of V^T [256][m_pad] (element (q, b, d) at d * m_pad +
  // b * frames + q), box (32 keys, 1, 32 d): rows beyond an utterance's frames are clipped by the map
  CUtensorMap tmVt;
  void* vt_out;                                   // non-null: tile n == 2 (V) is stored transposed through tmVt
  // GENERIC: bf16 residuals (same geometry as the output) and, per tap group, the raw / activated outputs: a group of
  // a polyphase transposed conv writes rows q * row_mul + phase, which is a plain 3-D view with a row_mul-fold pitch
  CUtensorMap tmRes[3];
  CUtensorMap tmOut[2][kMaxGroups];
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
  int res_bufs;                                   // RESNORM: residual-stream buffers per epilogue warp (1 or 2);
                                                  // GENERIC: staging blocks per epilogue warp (-1: LSU epilogue)
  int n_res;                                      // GENERIC: number of residual tensors (0..3)
  int gen_lsu;                                    // GENERIC: 1 = register / LSU epilogue (MMA-bound launches, fused tails)
  int gen_nbuf;                                   // GENERIC TMA epilogue: staging depth (1 or 2) of residual and output blocks
  int l2_keep;                                    // RESNORM: 1 = evict_last policy on the residual stream's TMA transfers
  int res_tail;                                   // RESNORM: 1 = the CTA's LAST tile takes its remaining residual chunks all at once
                                                  // into the (by then idle) weight ring instead of one latency after the other
  int w_early;                                    // 1 = request the first weight slabs before the dependency wait (PDL)
  int w_dynamic;                                  // 1 = the "weight" operand is produced by a predecessor kernel (swapped-operand v^T GEMM)
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
  float res_unact;                                // GENERIC: negative residual values are multiplied by this (1 = residuals are raw;
                                                  // 1 / slope = they carry a leaky_relu(slope) and the raw value is recovered here)
  int norm_mode;                                  // RESNORM: 0 none, 1 adaptive (L2), 2 rms
  float f0, f1, f2, f3;                           // EULER: dt, std, mean, pad
  void* aux0;                                     // EULER: mel fp32 (or null)
  void* aux1;                                     // EULER: mel bf16
  int aux_rows;                                   // EULER: rows per utterance of the (compact) mel outputs; rows beyond are not written
  int flat_frames;                                // GLU: > 0 = the launch sees ONE sequence of batch * flat_frames rows (row tiles span
                                                  // utterances, which a zero pad row separates); lengths are looked up per row
#ifdef SRB_TRACE
  unsigned long long* trace;                      // debug build only: %globaltimer stamps, see SRB_TRACE_AT
#endif
};

// Debug-only timeline (built with -DSRB_TRACE into libsrb_trace.so by tools/trace_kernels.py; never in libsrb.so):
// stamp slot (CTA, role, tile, k) with the global nanosecond timer.
#ifdef SRB_TRACE
__device__ __forceinline__ void srb_trace_at(const ConvGemmParams& p, int role, int tile_it, int k) {
  if (p.trace != nullptr && blockIdx.x < 8 && tile_it < 8) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[((blockIdx.x * 7 + role) * 8 + tile_it) * 4 + k] = t;
  }
}
#define SRB_TRACE_AT(role, it, k) do { if (lane == 0) srb_trace_at(p, role, it, k); } while (0)
// inside epi_resnorm (roles 3..6), first epilogue warp only
#define SRB_TRACE_EPI(role, it, k) do { if (w.lane == 0 && half == 0 && quarter == 2) srb_trace_at(p, role, it, k); } while (0)
#else
#define SRB_TRACE_AT(role, it, k) do { } while (0)
#define SRB_TRACE_EPI(role, it, k) do { } while (0)
#endif

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
// Thread <-> one accumulator row (TMEM lane).  Wide tiles use EIGHT epilogue warps (two column halves per TMEM lane
// quarter): one warp per SM sub-partition cannot hide TMEM / global latency.  `tacc` is the TMEM address of
// (lane base, column 0) of this tile's accumulator buffer.  All 32 lanes execute every tcgen05.ld/st (sync.aligned).
//
// Global traffic is COALESCED through a 4 KB per-warp staging buffer: with thread <-> row every 16-byte access of a
// warp would touch 32 different cache lines (measured: the LSU, not the tensor core, paced these kernels).  A warp
// owns a 32-row x (P x 16 B) block; in global memory lane l moves the pieces l, l+32, ... of the row-major block
// (32/P rows x P*16 contiguous bytes per instruction), in shared memory it reads/writes its own row; 16-byte pieces
// are XOR-swizzled by row so both sides are bank-conflict free.
struct EpiWarp {
  uint8_t* stage;   // 32 rows x (P * 16) bytes, P = pieces per row of the widest block the epilogue moves
  int lane;
  int row0;         // first tile-local output row q of this warp
};

// row pitch = P * 16 bytes; the XOR keeps every 8-lane phase of a 128-bit access on distinct banks
template <int P>
__device__ __forceinline__ uint4* stage_slot(uint8_t* stage, int row, int piece) {
  const int swz = P == 8 ? (row & 7) : (P == 4 ? ((row >> 1) & 3) : ((row >> 2) & 1));
  return reinterpret_cast<uint4*>(stage + row * (P * 16) + ((piece ^ swz) << 4));
}

// issue the coalesced loads of a 32 x (P*16 B) block (rows >= valid_rows read as zero)
template <int P>
__device__ __forceinline__ void gather_issue(const void* gbase, long long row_stride_bytes, int valid_rows, int lane, uint4* t) {
  const uint8_t* g = static_cast<const uint8_t*>(gbase);
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const int idx = lane + 32 * i, row = idx / P, piece = idx % P;
    t[i] = row < valid_rows ? __ldg(reinterpret_cast<const uint4*>(g + row * row_stride_bytes + piece * 16)) : make_uint4(0, 0, 0, 0);
  }
}
// route the loaded pieces through the staging buffer; afterwards v[j] is piece j of this lane's own row
template <int P>
__device__ __forceinline__ void gather_finish(const EpiWarp& w, const uint4* t, uint4* v) {
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const int idx = w.lane + 32 * i;
    *stage_slot<P>(w.stage, idx / P, idx % P) = t[i];
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < P; ++j) v[j] = *stage_slot<P>(w.stage, w.lane, j);
  __syncwarp();
}
// v[j] = piece j of this lane's row  ->  coalesced global stores of the 32 x (P*16 B) block
template <int P>
__device__ __forceinline__ void scatter_store(const EpiWarp& w, const uint4* v, void* gbase, long long row_stride_bytes, int valid_rows) {
#pragma unroll
  for (int j = 0; j < P; ++j) *stage_slot<P>(w.stage, w.lane, j) = v[j];
  __syncwarp();
  uint8_t* g = static_cast<uint8_t*>(gbase);
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const int idx = w.lane + 32 * i, row = idx / P, piece = idx % P;
    const uint4 x = *stage_slot<P>(w.stage, row, piece);
    if (row < valid_rows) *reinterpret_cast<uint4*>(g + row * row_stride_bytes + piece * 16) = x;
  }
  __syncwarp();
}

// asynchronous form of gather_issue: the block goes global -> (swizzled) warp-private shared buffer with cp.async, no
// registers involved, so several blocks can be in flight per warp while it works.  Rows >= valid_rows are zero-filled
// (src-size 0; `safe` is any valid address of the tensor, used so that no out-of-range address is ever formed).
template <int P>
__device__ __forceinline__ void async_gather(uint8_t* buf, const void* gbase, long long row_stride_bytes, int valid_rows,
                                             int lane, const void* safe) {
  const uint8_t* g = static_cast<const uint8_t*>(gbase);
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const int idx = lane + 32 * i, row = idx / P, piece = idx % P;
    const bool ok = row < valid_rows;
    const void* src = ok ? static_cast<const void*>(g + row * row_stride_bytes + piece * 16) : safe;
    cp_async_16(smem_u32(stage_slot<P>(buf, row, piece)), src, ok ? 16u : 0u);
  }
}

__device__ __forceinline__ int clamp_rows(int total, int row0) {
  const int v = total - row0;
  return v < 0 ? 0 : (v > 32 ? 32 : v);
}

// residual value as stored -> raw value (see ConvGemmParams::res_unact): the activated copy of a tensor carries the same
// information as the raw one (leaky_relu with a non-zero slope is invertible, and bf16(slope * y) / slope is y to within the
// same 2^-9 a bf16 copy of y itself has), so the resblock chains keep ONE copy and the consumer undoes the activation
__device__ __forceinline__ float unact(float v, float k) { return v < 0.f ? v * k : v; }

// Register / LSU form of the GENERIC epilogue (kept for launches with two or three residual tensors -- the fused MRF
// tails, whose long MMA phase hides it: their TMA staging would take the shared memory of two weight-ring stages).
template <int BN, int NHALF>
__device__ __forceinline__ void epi_generic_lsu(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q, int half,
                                            const EpiWarp& w) {
  constexpr int CW = BN < 32 ? BN : 32;
  constexpr int COLS = BN / NHALF;
  constexpr int P = CW / 8;   // 16-byte pieces of bf16 per chunk row
  const float sc = p.scale, sl = p.slope;
  const float ru = kSplit == 1 ? p.res_unact : 1.f;   // (the split build's residuals are always raw hi + lo pairs)
  if (p.row_mul == 1) {
    // contiguous output rows: coalesced path.  (Split build: bf16 tensors are kSplit times as wide -- [hi | lo | hi]
    // blocks, out_row_stride / res_row_stride columns apart; residuals are read as hi + lo.)
    const int vrows = clamp_rows(p.group_rows[tc.group], w.row0);
    const long long row0 = (long long)w.row0 + p.group_row_add[tc.group];
    const long long obase = ((long long)tc.b * p.out_batch_stride + row0 * p.out_row_stride) * kSplit + (long long)tc.n * BN;
    const long long rbase = ((long long)tc.b * p.res_batch_stride + row0 * p.res_row_stride) * kSplit + (long long)tc.n * BN;
    const long long opitch = p.out_row_stride * 2 * kSplit, rpitch = p.res_row_stride * 2 * kSplit;
    constexpr int NRP = kSplit == 3 ? 2 : 1;   // parts of a residual tensor
#pragma unroll 1
    for (int c0 = half * COLS; c0 < (half + 1) * COLS; c0 += CW) {
      uint4 rt[3][NRP][P];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        if (p.res[r]) {
#pragma unroll
          for (int h = 0; h < NRP; ++h)
            gather_issue<P>(static_cast<const __nv_bfloat16*>(p.res[r]) + rbase + h * p.res_row_stride + c0, rpitch, vrows, w.lane,
                            rt[r][h]);
        }
      float y[CW];
      tmem_ld_f<CW>(tacc + c0, y);
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + tc.n * BN + c0);
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          const float4 bb = __ldg(b4 + j);
          y[4 * j + 0] += bb.x; y[4 * j + 1] += bb.y; y[4 * j + 2] += bb.z; y[4 * j + 3] += bb.w;
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (p.res[r]) {
#pragma unroll
          for (int h = 0; h < NRP; ++h) {
            uint4 rv[P];
            gather_finish<P>(w, rt[r][h], rv);
#pragma unroll
            for (int j = 0; j < P; ++j) {
              const uint4 u = rv[j];
              y[8 * j + 0] += unact(bf16_lo(u.x), ru); y[8 * j + 1] += unact(bf16_hi(u.x), ru);
              y[8 * j + 2] += unact(bf16_lo(u.y), ru); y[8 * j + 3] += unact(bf16_hi(u.y), ru);
              y[8 * j + 4] += unact(bf16_lo(u.z), ru); y[8 * j + 5] += unact(bf16_hi(u.z), ru);
              y[8 * j + 6] += unact(bf16_lo(u.w), ru); y[8 * j + 7] += unact(bf16_hi(u.w), ru);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CW; ++j) y[j] *= sc;
      // store one bf16 output tensor: hi (and, split build, the rounding rest and hi again)
      auto put = [&](void* base, auto act_tag) {
        constexpr bool act = decltype(act_tag)::value;
        uint4 o[P], ol[P];
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
        uint32_t* lw = reinterpret_cast<uint32_t*>(ol);
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) {
          const float a = act ? lrelu(y[2 * j], sl) : y[2 * j], b2 = act ? lrelu(y[2 * j + 1], sl) : y[2 * j + 1];
          ow[j] = pack_bf16(a, b2);
          if constexpr (kSplit == 3) lw[j] = pack_bf16_rest(a, b2, ow[j]);
        }
        __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(base) + obase + c0;
        scatter_store<P>(w, o, ob, opitch, vrows);
        if constexpr (kSplit == 3) {
          scatter_store<P>(w, ol, ob + p.out_row_stride, opitch, vrows);
          scatter_store<P>(w, o, ob + 2 * p.out_row_stride, opitch, vrows);
        }
      };
      if (p.out1) put(p.out1, std::false_type{});
      if (p.out0) put(p.out0, std::true_type{});
    }
    return;
  }
  // strided output rows (polyphase transposed conv: row = q * stride + phase): per-row stores, no residuals
  const bool valid = q < p.group_rows[tc.group];
  const long long orow = (long long)q * p.row_mul + p.group_row_add[tc.group];
  const long long obase = ((long long)tc.b * p.out_batch_stride + orow * p.out_row_stride) * kSplit + (long long)tc.n * BN;
#pragma unroll 1
  for (int c0 = half * COLS; c0 < (half + 1) * COLS; c0 += CW) {
    float y[CW];
    tmem_ld_f<CW>(tacc + c0, y);
    if (valid) {
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + tc.n * BN + c0);
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          const float4 bb = __ldg(b4 + j);
          y[4 * j + 0] += bb.x; y[4 * j + 1] += bb.y; y[4 * j + 2] += bb.z; y[4 * j + 3] += bb.w;
        }
      }
#pragma unroll
      for (int j = 0; j < CW; ++j) y[j] *= sc;
      auto put = [&](void* base, auto act_tag) {
        constexpr bool act = decltype(act_tag)::value;
        uint4 o[P], ol[P];
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
        uint32_t* lw = reinterpret_cast<uint32_t*>(ol);
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) {
          const float a = act ? lrelu(y[2 * j], sl) : y[2 * j], b2 = act ? lrelu(y[2 * j + 1], sl) : y[2 * j + 1];
          ow[j] = pack_bf16(a, b2);
          if constexpr (kSplit == 3) lw[j] = pack_bf16_rest(a, b2, ow[j]);
        }
        __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(base) + obase + c0;
#pragma unroll
        for (int j = 0; j < P; ++j) {
          reinterpret_cast<uint4*>(ob)[j] = o[j];
          if constexpr (kSplit == 3) {
            reinterpret_cast<uint4*>(ob + p.out_row_stride)[j] = ol[j];
            reinterpret_cast<uint4*>(ob + 2 * p.out_row_stride)[j] = o[j];
          }
        }
      };
      if (p.out1) put(p.out1, std::false_type{});
      if (p.out0) put(p.out0, std::true_type{});
    }
  }
}

// GENERIC epilogue (vocoder convs, v^T): y = (acc + bias + residuals) * scale -> raw bf16 and / or leaky_relu bf16.
// All global traffic is TMA (see the note above epi_resnorm): the bf16 residual blocks (32 rows x CW columns) of the
// warp form a stream over chunks and tiles, double buffered, one mbarrier per buffer armed for all n_res blocks of a
// chunk; outputs are staged in two alternating blocks and stored with cp.async.bulk.tensor.  Per-warp staging:
// [2 buffers][n_res blocks] then [2 output blocks], every block CW * 64 bytes.  (n_res <= 1 here; see epi_generic_lsu.)
struct GenStream {
  uint32_t buf, bar;
  uint8_t* gen;
  int n_res;
  int nbuf;   // staging depth: 2 = double buffered, 1 = single
  int seq;    // chunks consumed so far
  int nout;   // output blocks stored so far
};

template <int BN, int NHALF, typename IssueLoad>
__device__ __forceinline__ void epi_generic(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int half,
                                            const EpiWarp& w, GenStream& gs, IssueLoad&& issue_load) {
  constexpr int CW = BN < 32 ? BN : 32;
  constexpr int COLS = BN / NHALF;
  constexpr int P = CW / 8;             // 16-byte pieces of bf16 per block row
  constexpr int BLK = CW * 64;          // bytes of a 32-row block
  const float sc = p.scale, sl = p.slope, ru = p.res_unact;
  const int n_res = gs.n_res, nbuf = gs.nbuf;
  const uint32_t out_stage = gs.buf + nbuf * n_res * BLK;
  uint8_t* out_gen = gs.gen + nbuf * n_res * BLK;
  const int g = tc.group;
  auto put_block = [&](const uint4 (&o)[P], const CUtensorMap* tm, int col) {
    // the block last stored from this staging slot has been read
    if (w.lane == 0) {
      if (nbuf == 2) bulk_wait_read<1>();
      else bulk_wait_read<0>();
    }
    __syncwarp();
    const int slot = nbuf == 2 ? (gs.nout & 1) : 0;
    uint8_t* hb = out_gen + slot * BLK;
#pragma unroll
    for (int j = 0; j < P; ++j) *stage_slot<P>(hb, w.lane, j) = o[j];
    fence_proxy_async_smem();
    __syncwarp();
    if (w.lane == 0) {
      tma_store_3d(tm, out_stage + slot * BLK, col, w.row0, tc.b);
      bulk_commit();
    }
    ++gs.nout;
  };
#pragma unroll 1
  for (int c0 = half * COLS; c0 < (half + 1) * COLS; c0 += CW) {
    float y[CW];
    tmem_ld_f<CW>(tacc + c0, y);
    if (p.bias) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + tc.n * BN + c0);
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) {
        const float4 bb = __ldg(b4 + j);
        y[4 * j + 0] += bb.x; y[4 * j + 1] += bb.y; y[4 * j + 2] += bb.z; y[4 * j + 3] += bb.w;
      }
    }
    if (n_res == 1) {
      const int s = gs.seq, bi = nbuf == 2 ? (s & 1) : 0;
      mbar_wait(gs.bar + 8u * bi, nbuf == 2 ? ((s >> 1) & 1) : (s & 1));
#pragma unroll 1
      for (int r = 0; r < n_res; ++r) {
        const uint8_t* rb = gs.gen + (bi * n_res + r) * BLK;
#pragma unroll
        for (int j = 0; j < P; ++j) {
          const uint4 u = *stage_slot<P>(const_cast<uint8_t*>(rb), w.lane, j);
          y[8 * j + 0] += unact(bf16_lo(u.x), ru); y[8 * j + 1] += unact(bf16_hi(u.x), ru);
          y[8 * j + 2] += unact(bf16_lo(u.y), ru); y[8 * j + 3] += unact(bf16_hi(u.y), ru);
          y[8 * j + 4] += unact(bf16_lo(u.z), ru); y[8 * j + 5] += unact(bf16_hi(u.z), ru);
          y[8 * j + 6] += unact(bf16_lo(u.w), ru); y[8 * j + 7] += unact(bf16_hi(u.w), ru);
        }
      }
      __syncwarp();   // every lane has its rows: the buffer takes the chunk after next (the next one when single)
      issue_load(s + nbuf, bi);
      ++gs.seq;
    }
#pragma unroll
    for (int j = 0; j < CW; ++j) y[j] *= sc;
    if (p.out1) {
      uint4 o[P];
#pragma unroll
      for (int j = 0; j < P; ++j)
        o[j] = make_uint4(pack_bf16(y[8 * j], y[8 * j + 1]), pack_bf16(y[8 * j + 2], y[8 * j + 3]),
                          pack_bf16(y[8 * j + 4], y[8 * j + 5]), pack_bf16(y[8 * j + 6], y[8 * j + 7]));
      put_block(o, &p.tmOut[0][g], tc.n * BN + c0);
    }
    if (p.out0) {
      uint4 o[P];
#pragma unroll
      for (int j = 0; j < P; ++j)
        o[j] = make_uint4(pack_bf16(lrelu(y[8 * j], sl), lrelu(y[8 * j + 1], sl)),
                          pack_bf16(lrelu(y[8 * j + 2], sl), lrelu(y[8 * j + 3], sl)),
                          pack_bf16(lrelu(y[8 * j + 4], sl), lrelu(y[8 * j + 5], sl)),
                          pack_bf16(lrelu(y[8 * j + 6], sl), lrelu(y[8 * j + 7], sl)));
      put_block(o, &p.tmOut[1][g], tc.n * BN + c0);
    }
  }
}

// FFN conv1 tile = [128 value columns | 128 gate columns] -> 128 outputs  (fastspeech/modules.py:27-30, 62-69)
template <int NHALF>
__device__ __forceinline__ void epi_glu(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q, int half,
                                        const EpiWarp& w) {
  constexpr int COLS = 128 / NHALF;
  const int vrows = clamp_rows(p.group_rows[0], w.row0);
  bool keep = false;
  if (q < p.group_rows[0]) {
    if (p.flat_frames > 0) {
      const int b = q / p.flat_frames;
      keep = q - b * p.flat_frames < p.lengths[b];
    } else {
      keep = q < p.lengths[tc.b];
    }
  }
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out0) +
                       ((long long)tc.b * p.out_batch_stride + (long long)w.row0 * p.out_row_stride) * kSplit + tc.n * 128;
  const float* bias = p.bias + tc.n * 256;
#pragma unroll 1
  for (int c0 = half * COLS; c0 < (half + 1) * COLS; c0 += 32) {
    uint32_t v[32], g[32];
    tmem_ld32(tacc + c0, v);
    tmem_ld32(tacc + 128 + c0, g);
    tmem_ld_wait();
    uint4 o[4], ol[4];
    uint32_t* ow = reinterpret_cast<uint32_t*>(o);
    uint32_t* lw = reinterpret_cast<uint32_t*>(ol);
    // (the tanh.approx form of SiLU is good to ~5e-4, below one bf16 rounding; the split build takes exp / divide)
    auto act = [](float x) { return kSplit == 3 ? silu(x) : silu_tanh(x); };
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + c0) + j);
      const float4 bg = __ldg(reinterpret_cast<const float4*>(bias + 128 + c0) + j);
      const float h0 = act(__uint_as_float(g[4 * j + 0]) + bg.x) * (__uint_as_float(v[4 * j + 0]) + bv.x);
      const float h1 = act(__uint_as_float(g[4 * j + 1]) + bg.y) * (__uint_as_float(v[4 * j + 1]) + bv.y);
      const float h2 = act(__uint_as_float(g[4 * j + 2]) + bg.z) * (__uint_as_float(v[4 * j + 2]) + bv.z);
      const float h3 = act(__uint_as_float(g[4 * j + 3]) + bg.w) * (__uint_as_float(v[4 * j + 3]) + bv.w);
      ow[2 * j] = keep ? pack_bf16(h0, h1) : 0u;       // pads are zeroed before conv2 (fastspeech/modules.py:69)
      ow[2 * j + 1] = keep ? pack_bf16(h2, h3) : 0u;
      if constexpr (kSplit == 3) {
        lw[2 * j] = keep ? pack_bf16_rest(h0, h1, ow[2 * j]) : 0u;
        lw[2 * j + 1] = keep ? pack_bf16_rest(h2, h3, ow[2 * j + 1]) : 0u;
      }
    }
    scatter_store<4>(w, o, out + c0, p.out_row_stride * 2 * kSplit, vrows);
    if constexpr (kSplit == 3) {
      scatter_store<4>(w, ol, out + p.out_row_stride + c0, p.out_row_stride * 2 * kSplit, vrows);
      scatter_store<4>(w, o, out + 2 * p.out_row_stride + c0, p.out_row_stride * 2 * kSplit, vrows);
    }
  }
}

// y = acc + bias + residual (fp32 stream, in place allowed); x_out = y; xn = bf16(norm(y) * g) with pad rows zeroed.
// A row's 256 columns live in ONE TMEM lane; with two column halves per lane quarter the two warps exchange their
// partial sums of squares through `red` (smem, [2][128] floats) and a 64-thread named barrier (ids 1-4, immediates so
// that ptxas accounts for them; id 0 is __syncthreads).
__device__ __forceinline__ void pair_barrier(int quarter) {
  switch (quarter) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

// Epilogue I/O goes through TMA.  In-kernel timelines showed the register / LSU path (ld.global or cp.async into a
// staging buffer, st.shared + ld.shared + st.global out of it) pacing every small-K launch: 7-10 us of epilogue per
// 128 x 256 tile against 1 us of MMAs, bound by LSU / shared-memory-pipe throughput (neither more warps nor L2
// residency moved it).  With TMA a block costs the warp one instruction on one lane; the staging buffers use the
// tensor maps' own swizzle (128-byte rows: stage_slot<8>, 64-byte rows: stage_slot<4>), so a thread reads / writes its
// own row bank-conflict free and rows outside the tensor are zero-filled on load and clipped on store.
//
// The fp32 residual blocks (32 rows x 32 columns = 4 KB) of a warp's column half form a STREAM -- chunk after chunk,
// tile after tile.  Chunk s lands in load buffer s % nb and signals that buffer's mbarrier (phase (s / nb) & 1);
// `issue_load(s, buf, bar)` asks for chunk s (nothing past the end of the stream).
//   nb == 2: separate 4 KB store stage; a load buffer is refilled as soon as the warp has read it, also across tiles.
//   nb == 1 (long K loops, shared memory goes to the weight ring): the output is staged in place and the buffer is
//            refilled once the store has read it (the latency hides under the MMAs of the next tile).
struct ResStream {
  uint32_t buf;        // shared address of the nb load buffers (1024-byte aligned), then the store stage when nb == 2
  uint32_t bar;        // nb mbarriers
  uint8_t* gen;        // generic pointer to `buf`
  int nb;
  int seq;             // chunks consumed so far
  // Last tile of the CTA (see ConvGemmParams::res_tail): once that tile's MMAs are complete nothing uses the weight ring
  // any more, so the chunks the stream has not requested yet (all but the first `nb`) are requested at once into this
  // warp's 3 x 4 KB slice of it, each with its own single-use mbarrier; they are staged for the store in place.
  uint32_t tail_buf;   // shared address of the slice (1024-byte aligned)
  uint8_t* tail_gen;   // generic pointer to it
  uint32_t tail_bar;   // 3 mbarriers
  int tail;            // 1 while the current tile is served that way
};

template <int NHALF, typename IssueLoad>
__device__ __forceinline__ void epi_resnorm(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q, int half,
                                            float* red, int row_in_tile, int quarter, const EpiWarp& w, ResStream& rs,
                                            IssueLoad&& issue_load) {
  constexpr int COLS = 256 / NHALF;
  constexpr int NCHUNK = COLS / 32;
  static_assert(NCHUNK == 4, "RESNORM runs with eight epilogue warps");
  const uint32_t tcol = tacc + half * COLS;
  const int nb = rs.nb;
  const uint32_t st_stage = rs.buf + (nb == 2 ? 2 * 4096 : 0);
  uint8_t* st_gen = rs.gen + (nb == 2 ? 2 * 4096 : 0);
  float sumsq = 0.f;
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    const int s = rs.seq;
    const int bi = nb == 2 ? (s & 1) : 0;
    const bool from_tail = rs.tail && c >= nb;          // chunk c of the CTA's last tile sits in the weight-ring slice
    if (from_tail) mbar_wait(rs.tail_bar + 8u * (c - nb), 0u);
    else mbar_wait(rs.bar + 8u * bi, nb == 2 ? ((s >> 1) & 1) : (s & 1));
    SRB_TRACE_EPI(3, s >> 2, c);   // residual chunk c is in shared memory
    uint8_t* src_gen = from_tail ? rs.tail_gen + (c - nb) * 4096 : rs.gen + bi * 4096;
    uint4 rr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) rr[j] = *stage_slot<8>(src_gen, w.lane, j);
    __syncwarp();   // every lane has its row
    if (nb == 2 && !rs.tail) issue_load(s + 2, bi);
    uint32_t v[32];
    tmem_ld32(tcol + c * 32, v);
    tmem_ld_wait();
    uint4 yo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 y;
      y.x = __uint_as_float(v[4 * j + 0]) + __uint_as_float(rr[j].x);
      y.y = __uint_as_float(v[4 * j + 1]) + __uint_as_float(rr[j].y);
      y.z = __uint_as_float(v[4 * j + 2]) + __uint_as_float(rr[j].z);
      y.w = __uint_as_float(v[4 * j + 3]) + __uint_as_float(rr[j].w);
      if (p.bias) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + half * COLS + c * 32 + 4 * j));
        y.x += bb.x; y.y += bb.y; y.z += bb.z; y.w += bb.w;
      }
      sumsq += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
      v[4 * j + 0] = __float_as_uint(y.x); v[4 * j + 1] = __float_as_uint(y.y);
      v[4 * j + 2] = __float_as_uint(y.z); v[4 * j + 3] = __float_as_uint(y.w);
      yo[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    // store stage of this chunk: the separate stage (nb == 2), the load buffer itself (nb == 1), or -- a chunk that came
    // through the weight-ring slice -- its own slot there (used once, so nothing to wait for)
    const bool own_slot = from_tail && nb != 2;
    const uint32_t o_stage = own_slot ? rs.tail_buf + (c - nb) * 4096 : st_stage;
    uint8_t* o_gen = own_slot ? src_gen : st_gen;
    if (nb == 2) {
      if (w.lane == 0) bulk_wait_read<0>();   // the previous block has left the store stage
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) *stage_slot<8>(o_gen, w.lane, j) = yo[j];
    fence_proxy_async_smem();
    __syncwarp();
    if (w.lane == 0) {
      if (p.l2_keep) tma_store_3d_hint(&p.tmO1, o_stage, half * COLS + c * 32, w.row0, tc.b, kL2EvictLast);
      else tma_store_3d(&p.tmO1, o_stage, half * COLS + c * 32, w.row0, tc.b);
      bulk_commit();
      // in place: the buffer takes the next chunk once the store has read it; the last chunk of a tile waits until
      // the bf16 pass below has used the buffer as its stage
      if (nb != 2 && !rs.tail && (c < NCHUNK - 1 || p.norm_mode == 0)) {
        bulk_wait_read<0>();
        issue_load(s + 1, 0);
      }
    }
    SRB_TRACE_EPI(4, s >> 2, c);   // fp32 block c stored
    ++rs.seq;
    if (p.norm_mode != 0) tmem_st32(tcol + c * 32, v);  // stash y for the second pass
  }
  if (p.norm_mode == 0) return;
  tmem_st_wait();
  if constexpr (NHALF == 2) {
    red[half * 128 + row_in_tile] = sumsq;
    pair_barrier(quarter);
    sumsq += red[(half ^ 1) * 128 + row_in_tile];
  }
  SRB_TRACE_EPI(6, (rs.seq >> 2) - 1, 1);   // row norms exchanged
  float inv;
  if (p.norm_mode == 1) inv = 1.f / fmaxf(sqrtf(sumsq), 1e-12f);           // F.normalize (norm.py:41)
  else inv = rsqrtf(sumsq * (1.f / 256.f) + 1.1920928955078125e-07f);      // nn.RMSNorm eps = finfo(fp32).eps
  const bool keep = q < p.group_rows[0] && (p.lengths == nullptr || q < p.lengths[tc.b]);
  const float* gv = p.vec0 + half * COLS;
  // bf16 blocks (32 rows x 64 B = 2 KB) alternate between the two halves of the 4 KB store stage
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    uint32_t v[32];
    tmem_ld32(tcol + c * 32, v);
    tmem_ld_wait();
    uint4 o[4], ol[4];
    uint32_t* ow = reinterpret_cast<uint32_t*>(o);
    uint32_t* lw = reinterpret_cast<uint32_t*>(ol);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gv + c * 32 + 4 * j));
      const float n0 = __uint_as_float(v[4 * j]) * inv * g.x, n1 = __uint_as_float(v[4 * j + 1]) * inv * g.y;
      const float n2 = __uint_as_float(v[4 * j + 2]) * inv * g.z, n3 = __uint_as_float(v[4 * j + 3]) * inv * g.w;
      // select (not multiply) so a non-finite pad row can never leak into the conv taps of valid frames
      ow[2 * j] = keep ? pack_bf16(n0, n1) : 0u;
      ow[2 * j + 1] = keep ? pack_bf16(n2, n3) : 0u;
      if constexpr (kSplit == 3) {
        lw[2 * j] = keep ? pack_bf16_rest(n0, n1, ow[2 * j]) : 0u;
        lw[2 * j + 1] = keep ? pack_bf16_rest(n2, n3, ow[2 * j + 1]) : 0u;
      }
    }
    if constexpr (kSplit == 3) {
      // split build: three blocks per chunk (hi | rest | hi, 256 columns apart), one after the other through the stage
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        if (w.lane == 0) bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) *stage_slot<4>(st_gen, w.lane, j) = part == 1 ? ol[j] : o[j];
        fence_proxy_async_smem();
        __syncwarp();
        if (w.lane == 0) {
          tma_store_3d(&p.tmO0, st_stage, part * 256 + half * COLS + c * 32, w.row0, tc.b);
          bulk_commit();
        }
      }
      continue;
    }
    if (w.lane == 0) {
      if (c < 2) bulk_wait_read<0>();   // fp32 blocks (and, in place, the whole buffer) have been read
      else bulk_wait_read<1>();         // the other half may still be draining
    }
    __syncwarp();
    uint8_t* hb = st_gen + (c & 1) * 2048;
#pragma unroll
    for (int j = 0; j < 4; ++j) *stage_slot<4>(hb, w.lane, j) = o[j];
    fence_proxy_async_smem();
    __syncwarp();
    if (w.lane == 0) {
      tma_store_3d(&p.tmO0, st_stage + (c & 1) * 2048, half * COLS + c * 32, w.row0, tc.b);
      bulk_commit();
    }
    SRB_TRACE_EPI(5, (rs.seq >> 2) - 1, c);   // bf16 block c stored
  }
  if (nb != 2 && !rs.tail && w.lane == 0) {
    bulk_wait_read<0>();
    issue_load(rs.seq, 0);
  }
  if constexpr (NHALF == 2) {
    // the pair must not overwrite `red` for the next tile before both have read it
    pair_barrier(quarter);
  }
}

// to_qkv tile n: 0 = q (2 heads x 128), 1 = k, 2 = v.  Rotary (transformer.py:66-73) on q,k:
// out[i] = t[i] cos - t[i+64] sin ; out[i+64] = t[i+64] cos + t[i] sin, angle = pos * inv_freq[i], i < 64.
// With eight epilogue warps, column half h of a lane quarter is head h.
// The (rows, 64) fp32 cos/sin tables are NOT streamed per row (that was 128 KB of L2 reads per tile, register-staged
// and spilled): a warp's 32 positions are pos = r0 + lane with r0 a multiple of 32, so
//   cos(pos f) = cos(r0 f) cos(lane f) - sin(r0 f) sin(lane f),  sin(pos f) = sin(r0 f) cos(lane f) + cos(r0 f) sin(lane f)
// with row r0 of the tables read warp-uniformly (broadcast, L1 resident) and rows 0..31 kept in shared memory for the
// whole kernel (`tab_b`: [cos | sin][32 rows][256 B], 16-byte pieces XOR-swizzled by row).  The sum formula differs from
// the tabulated fp32 value only by the rounding of the fp32 angle the reference itself carries (<= ulp(pos f) ~ 3e-5).
__device__ __forceinline__ void rope_table_to_smem(const ConvGemmParams& p, uint8_t* tab_b) {
  // 2 tables x 32 rows x 16 pieces of 16 bytes
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const int t = i >> 9, row = (i >> 4) & 31, piece = i & 15;
    const float4 v = __ldg(reinterpret_cast<const float4*>((t ? p.vec1 : p.vec0) + row * 64) + piece);
    *reinterpret_cast<float4*>(tab_b + t * 8192 + row * 256 + ((piece ^ (row & 7)) << 4)) = v;
  }
}

// row r0 of the cos | sin tables, one 16-byte piece per lane (issued before the warp waits for the tile's MMAs)
__device__ __forceinline__ float4 rope_row_fetch(const ConvGemmParams& p, const TileCoord& tc, const EpiWarp& w) {
  if (tc.n == 2) return make_float4(0.f, 0.f, 0.f, 0.f);
  // the row exists whenever the warp has a valid row; otherwise nothing is stored and row 0 stands in
  const long long r0 = clamp_rows(p.group_rows[0], w.row0) > 0 ? w.row0 : 0;
  return __ldg(reinterpret_cast<const float4*>((w.lane < 16 ? p.vec0 : p.vec1) + r0 * 64) + (w.lane & 15));
}

template <int NHALF>
__device__ __forceinline__ void epi_qkv_rope(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int half,
                                             const EpiWarp& w, const uint8_t* tab_b, const float4& row_piece) {
  static_assert(NHALF == 2, "QKV_ROPE runs with eight epilogue warps (one head per column half)");
  const int col0 = tc.n * 256 + half * 128;   // first output column of this warp
  const uint32_t tcol = tacc + half * 128;
  const uint32_t stage_s = smem_u32(w.stage);
  // output blocks (32 rows x 64 B = 2 KB, 64-byte swizzle) alternate between the two halves of the 4 KB stage
  int blk = 0;
  auto put_one = [&](const uint4 (&o)[4], int col) {
    if (w.lane == 0) bulk_wait_read<1>();   // the block stored from this half two blocks ago has been read
    __syncwarp();
    uint8_t* hb = w.stage + (blk & 1) * 2048;
#pragma unroll
    for (int j = 0; j < 4; ++j) *stage_slot<4>(hb, w.lane, j) = o[j];
    fence_proxy_async_smem();
    __syncwarp();
    if (w.lane == 0) {
      tma_store_3d(&p.tmO0, stage_s + (blk & 1) * 2048, col, w.row0, tc.b);
      bulk_commit();
    }
    ++blk;
  };
  // `f` holds the 32 fp32 values of this lane's row; split build: [hi | rest | hi] blocks out_row_stride columns apart
  auto put_block = [&](const float (&f)[32], int col) {
    uint4 o[4], ol[4];
    uint32_t* ow = reinterpret_cast<uint32_t*>(o);
    uint32_t* lw = reinterpret_cast<uint32_t*>(ol);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      ow[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
      if constexpr (kSplit == 3) lw[j] = pack_bf16_rest(f[2 * j], f[2 * j + 1], ow[j]);
    }
    put_one(o, col);
    if constexpr (kSplit == 3) {
      put_one(ol, col + (int)p.out_row_stride);
      put_one(o, col + 2 * (int)p.out_row_stride);
    }
  };
  if (tc.n == 2 && p.vt_out != nullptr) {
    // V stored TRANSPOSED ([d][utterance * frames + key], keys contiguous) for the attention kernel's K-major P V product:
    // a thread holds one key's 32 d-values per chunk and writes them as one bf16 each into a [32 d][32 keys] block
    // (64-byte rows, the tensor map's 64-byte swizzle; the 32 lanes of a warp fill one row per store: conflict free),
    // which leaves as one TMA store per chunk.  Replaces a separate swapped-operand GEMM launch.
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tcol + c * 32, v);
      tmem_ld_wait();
      if (w.lane == 0) bulk_wait_read<1>();   // the block stored from this half two blocks ago has been read
      __syncwarp();
      uint8_t* hb = w.stage + (blk & 1) * 2048;
      const int piece = w.lane >> 3, within = (w.lane & 7) * 2;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const __nv_bfloat16 h = __float2bfloat16_rn(__uint_as_float(v[j]));
        *reinterpret_cast<__nv_bfloat16*>(hb + j * 64 + ((piece ^ ((j >> 1) & 3)) << 4) + within) = h;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (w.lane == 0) {
        tma_store_3d(&p.tmVt, stage_s + (blk & 1) * 2048, w.row0, tc.b, half * 128 + c * 32);
        bulk_commit();
      }
      ++blk;
    }
    return;
  }
  if (tc.n == 2) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tcol + c * 32, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      put_block(f, col0 + c * 32);
    }
    return;
  }
  // table row r0 (cos pieces 0..15 | sin pieces 0..15) goes through a warp-private 512-byte area: broadcast reads
  float4* tab_a = reinterpret_cast<float4*>(w.stage + 4096);
  tab_a[w.lane] = row_piece;
  __syncwarp();
  const uint8_t* cb_row = tab_b + w.lane * 256;
  const uint8_t* sb_row = tab_b + 8192 + w.lane * 256;
  const int swz = w.lane & 7;
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    float norm2 = 0.f;   // squared norm of columns i, i + 64 (i in this frequency half) of the row: rotation invariant
    uint32_t lo[32], hi[32];
    tmem_ld32(tcol + f * 32, lo);
    tmem_ld32(tcol + 64 + f * 32, hi);
    tmem_ld_wait();
    float flo[32], fhi[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int piece = f * 8 + j;
      const float4 ca = tab_a[piece], sa = tab_a[16 + piece];
      const float4 cb = *reinterpret_cast<const float4*>(cb_row + ((piece ^ swz) << 4));
      const float4 sb = *reinterpret_cast<const float4*>(sb_row + ((piece ^ swz) << 4));
      const float cx = ca.x * cb.x - sa.x * sb.x, sx = sa.x * cb.x + ca.x * sb.x;
      const float cy = ca.y * cb.y - sa.y * sb.y, sy = sa.y * cb.y + ca.y * sb.y;
      const float cz = ca.z * cb.z - sa.z * sb.z, sz = sa.z * cb.z + ca.z * sb.z;
      const float cw = ca.w * cb.w - sa.w * sb.w, sw = sa.w * cb.w + ca.w * sb.w;
      const float a0 = __uint_as_float(lo[4 * j]), a1 = __uint_as_float(lo[4 * j + 1]);
      const float a2 = __uint_as_float(lo[4 * j + 2]), a3 = __uint_as_float(lo[4 * j + 3]);
      const float b0 = __uint_as_float(hi[4 * j]), b1 = __uint_as_float(hi[4 * j + 1]);
      const float b2 = __uint_as_float(hi[4 * j + 2]), b3 = __uint_as_float(hi[4 * j + 3]);
      norm2 += (a0 * a0 + b0 * b0) + (a1 * a1 + b1 * b1) + (a2 * a2 + b2 * b2) + (a3 * a3 + b3 * b3);
      flo[4 * j] = a0 * cx - b0 * sx; flo[4 * j + 1] = a1 * cy - b1 * sy;
      flo[4 * j + 2] = a2 * cz - b2 * sz; flo[4 * j + 3] = a3 * cw - b3 * sw;
      fhi[4 * j] = b0 * cx + a0 * sx; fhi[4 * j + 1] = b1 * cy + a1 * sy;
      fhi[4 * j + 2] = b2 * cz + a2 * sz; fhi[4 * j + 3] = b3 * cw + a3 * sw;
    }
    put_block(flo, col0 + f * 32);
    put_block(fhi, col0 + 64 + f * 32);
    if (p.aux0 != nullptr) {
      // per-(utterance, q|k, head, f) maximum of the partial squared row norms for the attention kernel's single-pass
      // test (max_f0 + max_f1 bounds the maximum of the sum); non-negative floats order like their bit patterns, rows
      // outside the tensor have zero accumulators
      const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(norm2));
      if (w.lane == 0) atomicMax(static_cast<unsigned*>(p.aux0) + ((tc.b * 2 + tc.n) * 2 + half) * 2 + f, mx);
    }
  }
}

// One N-column chunk (N = 32 or 16) of the Euler epilogue for a warp's 32 rows.  With thread <-> row every 16-byte access
// of a warp touched 32 different lines of the 320-byte xt rows (11 us of epilogue per 128 x 80 tile in the in-kernel
// timeline); the chunk now travels through the warp's staging buffer like the GLU output: coalesced loads of the
// 32 x (N x 4 B) block, own-row reads, coalesced stores of xt / its bf16 copy / the two mel outputs.
template <int N>
__device__ __forceinline__ void euler_chunk(const ConvGemmParams& p, uint32_t tacc, int c0, const EpiWarp& w, float* xt,
                                            __nv_bfloat16* xtb, int vrows, float* mel, __nv_bfloat16* melb, int mrows,
                                            bool is_pad) {
  constexpr int P = N / 4;   // 16-byte pieces of fp32 per chunk row
  const float dt = p.f0, sd = p.f1, mean = p.f2, padv = p.f3;
  uint4 t[P], xin[P];
  gather_issue<P>(xt + c0, 320, vrows, w.lane, t);
  uint32_t v[N];
  if constexpr (N == 32) tmem_ld32(tacc + c0, v);
  else tmem_ld16(tacc + c0, v);
  tmem_ld_wait();
  gather_finish<P>(w, t, xin);
  // bf16 copies (split build: [hi | rest | hi], 80 columns apart, rows of 3 x 80 values)
  auto put_bf16 = [&](const uint4 (&f32v)[P], __nv_bfloat16* dst, int rows) {
    uint4 hb[P / 2], lb[P / 2];
    uint32_t* hw = reinterpret_cast<uint32_t*>(hb);
    uint32_t* lw = reinterpret_cast<uint32_t*>(lb);
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const float a = __uint_as_float(f32v[j].x), b2 = __uint_as_float(f32v[j].y);
      const float c2 = __uint_as_float(f32v[j].z), d2 = __uint_as_float(f32v[j].w);
      hw[2 * j] = pack_bf16(a, b2);
      hw[2 * j + 1] = pack_bf16(c2, d2);
      if constexpr (kSplit == 3) {
        lw[2 * j] = pack_bf16_rest(a, b2, hw[2 * j]);
        lw[2 * j + 1] = pack_bf16_rest(c2, d2, hw[2 * j + 1]);
      }
    }
    scatter_store<P / 2>(w, hb, dst + c0, 160 * kSplit, rows);
    if constexpr (kSplit == 3) {
      scatter_store<P / 2>(w, lb, dst + 80 + c0, 160 * kSplit, rows);
      scatter_store<P / 2>(w, hb, dst + 160 + c0, 160 * kSplit, rows);
    }
  };
  uint4 xo[P];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    float4 x;
    x.x = __uint_as_float(xin[j].x) + dt * __uint_as_float(v[4 * j + 0]);
    x.y = __uint_as_float(xin[j].y) + dt * __uint_as_float(v[4 * j + 1]);
    x.z = __uint_as_float(xin[j].z) + dt * __uint_as_float(v[4 * j + 2]);
    x.w = __uint_as_float(xin[j].w) + dt * __uint_as_float(v[4 * j + 3]);
    xo[j] = make_uint4(__float_as_uint(x.x), __float_as_uint(x.y), __float_as_uint(x.z), __float_as_uint(x.w));
  }
  scatter_store<P>(w, xo, xt + c0, 320, vrows);
  put_bf16(xo, xtb, vrows);
  if (mel != nullptr) {
    uint4 mo[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      float4 m;
      m.x = is_pad ? padv : __uint_as_float(xo[j].x) * sd + mean;
      m.y = is_pad ? padv : __uint_as_float(xo[j].y) * sd + mean;
      m.z = is_pad ? padv : __uint_as_float(xo[j].z) * sd + mean;
      m.w = is_pad ? padv : __uint_as_float(xo[j].w) * sd + mean;
      mo[j] = make_uint4(__float_as_uint(m.x), __float_as_uint(m.y), __float_as_uint(m.z), __float_as_uint(m.w));
    }
    scatter_store<P>(w, mo, mel + c0, 320, mrows);
    put_bf16(mo, melb, mrows);
  }
}

// to_pred (N = 80) + Euler step in fp32 (models.py:183-184); last step also de-normalises and fills pads (:186-187)
__device__ __forceinline__ void epi_euler(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q, const EpiWarp& w) {
  const int rows = p.group_rows[0];
  const int vrows = clamp_rows(rows, w.row0);
  const long long xoff = (long long)tc.b * p.out_batch_stride + (long long)w.row0 * p.out_row_stride;
  float* xt = static_cast<float*>(p.out1) + xoff;
  __nv_bfloat16* xtb = static_cast<__nv_bfloat16*>(p.out0) + xoff * kSplit;
  // the mel outputs are compact: aux_rows (<= frames) rows per utterance, what the vocoder must see
  const bool has_mel = p.aux0 != nullptr;
  const int mrows = has_mel ? clamp_rows(p.aux_rows < rows ? p.aux_rows : rows, w.row0) : 0;
  const long long off = ((long long)tc.b * p.aux_rows + w.row0) * p.out_row_stride;
  float* mel = has_mel ? static_cast<float*>(p.aux0) + off : nullptr;
  __nv_bfloat16* melb = has_mel ? static_cast<__nv_bfloat16*>(p.aux1) + off * kSplit : nullptr;
  const bool is_pad = q < rows && p.lengths != nullptr && q >= p.lengths[tc.b];
  euler_chunk<32>(p, tacc, 0, w, xt, xtb, vrows, mel, melb, mrows, is_pad);
  euler_chunk<32>(p, tacc, 32, w, xt, xtb, vrows, mel, melb, mrows, is_pad);
  euler_chunk<16>(p, tacc, 64, w, xt, xtb, vrows, mel, melb, mrows, is_pad);
}

// Nearest-centroid assignment (the k-means unit quantiser on the input side of the path, utils/textless.py:9-21 ->
// sklearn KMeans.predict): score[row][j] = acc + bias[j] with bias[j] = -|c_j|^2 / 2 (-inf for padding columns), so
// argmax_j score = argmin_j |x - c_j|^2.  A thread scans its row's columns of this tile, then the best (score, column)
// pairs of the two column halves and of all N tiles are merged by ONE 64-bit atomicMax per (row, warp) on
// keys[row] = (order-preserving bits of the score << 32) | (0xFFFFFFFF - column): equal scores resolve to the SMALLEST
// column, numpy / sklearn argmin's first-occurrence rule.  `out1` = keys (u64 per row, zeroed by the caller).
template <int NHALF>
__device__ __forceinline__ void epi_argmax(const ConvGemmParams& p, uint32_t tacc, const TileCoord& tc, int q, int half) {
  constexpr int COLS = 256 / NHALF;
  float best = -INFINITY;
  int best_col = 0;
  const int col0 = tc.n * 256 + half * COLS;
  const float* bias = p.bias + col0;
#pragma unroll 1
  for (int c0 = 0; c0 < COLS; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tacc + half * COLS + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0) + j);
      const float s0 = __uint_as_float(v[4 * j]) + b4.x, s1 = __uint_as_float(v[4 * j + 1]) + b4.y;
      const float s2 = __uint_as_float(v[4 * j + 2]) + b4.z, s3 = __uint_as_float(v[4 * j + 3]) + b4.w;
      if (s0 > best) { best = s0; best_col = col0 + c0 + 4 * j; }
      if (s1 > best) { best = s1; best_col = col0 + c0 + 4 * j + 1; }
      if (s2 > best) { best = s2; best_col = col0 + c0 + 4 * j + 2; }
      if (s3 > best) { best = s3; best_col = col0 + c0 + 4 * j + 3; }
    }
  }
  if (q < p.group_rows[0] && best > -INFINITY) {
    uint32_t u = __float_as_uint(best);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // unsigned order == float order
    const unsigned long long key = (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(best_col));
    atomicMax(static_cast<unsigned long long*>(p.out1) + (long long)tc.b * p.out_batch_stride + q, key);
  }
}

// ---------------------------------------------------------------------------------------------- kernel
template <int BN, int EPI>
struct EpiWarps {
  // eight epilogue warps (two column halves per TMEM lane quarter) for the wide tiles, four otherwise
  // (sixteen for the FFN GLU tile, whose SiLU epilogue is transcendental-bound and needs the extra latency hiding)
  // (eight warps for the BN = 64 tiles too -- -DSRB_EPI_WARPS_MIN_BN=64 -- measured at config 2: k = 11 convs 254 -> 239 us,
  // but k = 7 180 -> 210, k = 3 152 -> 177 and the fused tail 509 -> 666: the second resident CTA is worth more)
#ifndef SRB_EPI_WARPS_MIN_BN
#define SRB_EPI_WARPS_MIN_BN 128
#endif
  static constexpr int value = EPI == EPI_GLU ? 16 : ((BN >= SRB_EPI_WARPS_MIN_BN && EPI != EPI_EULER) ? 8 : 4);
  // staging bytes per epilogue warp: bf16 output blocks are 2 KB (4 pieces per row); RESNORM adds `res_bufs` 4 KB
  // buffers for the asynchronous fp32 residual stream (they double as the fp32 output stage)
  // (multiples of 1024: the TMA-staged blocks need their swizzle alignment)
  // GENERIC: res_bufs = number of staging blocks (32 rows x min(BN, 32) bf16 columns) of the TMA epilogue, or -1 for
  // the LSU epilogue's 2 KB
  __host__ __device__ static constexpr int stage_bytes(int res_bufs) {
    return EPI == EPI_RESNORM ? (res_bufs == 2 ? 3 * 4096 : 4096)
           : EPI == EPI_QKV_ROPE ? 5120
           : EPI == EPI_GENERIC ? (res_bufs < 0 ? 2048 : ((res_bufs * (BN < 32 ? BN : 32) * 64 + 1023) & ~1023))
           : EPI == EPI_EULER ? 4096      // 32 rows x 128 B: one fp32 chunk of the staged Euler update
                                : 2048;
  }
  // CTA-wide extra: the rotary offset table (cos | sin of positions 0..31)
  static constexpr int extra_bytes = EPI == EPI_QKV_ROPE ? 16384 : 0;
};

// MC = 2: CTA-pair MMA (tcgen05 cta_group::2, M = 256 across two SMs): each CTA stages its own 128 activation rows
// and HALF of the weight slab; the leader CTA issues the MMAs for both; accumulators land in each CTA's own TMEM and
// each CTA runs its own epilogue.  All loads signal the leader's barriers, the leader's commits release both CTAs.
// MC = 1: launched as 2-CTA clusters.  The two CTAs of a cluster work on the same output-channel tile and on adjacent
// row tiles, so they consume the same weight slabs: each CTA fetches HALF of every slab and TMA-multicasts it into both
// shared memories, halving the L2 -> SM weight traffic that bounds the wide (BN = 256) GEMMs.  A slab slot is reused
// only after BOTH CTAs have consumed it (the MMA warps commit onto both CTAs' w_empty barriers).
__device__ __forceinline__ TileCoord decode_tile_mc(const ConvGemmParams& p, int pair_tile, int rank) {
  TileCoord c;
  c.group = 0;
  c.n = pair_tile % p.n_tiles;
  const int fm = 2 * (pair_tile / p.n_tiles) + rank;
  c.b = fm / p.m_tiles[0];
  c.m = fm % p.m_tiles[0];
  return c;
}

// WS = 1: weight-stationary.  When every (tap, K chunk) slab of a CTA's output-channel tile fits in shared memory
// (the K = 80 / 256 linears of the transformer, the C = 64 vocoder convs), a CTA keeps ONE channel tile for its whole
// life, loads its slabs once and then streams only activation boxes.  The L2 -> SM fabric (~43 B/clk per SM with all
// SMs pulling), not the tensor core, bounds the streamed form: a 128 x 256 x 16 MMA step needs 8 KB of weights per
// 128 clk.  `total_tiles` counts ROW tiles here; CTA c owns channel tile c % n_tiles and walks row tiles
// c / n_tiles, + gridDim / n_tiles, ...  (the host makes the grid a multiple of n_tiles).
template <int BN, int KB, int EPI, int MC = 0, int WS = 0>
__global__ void __launch_bounds__(64 + 32 * EpiWarps<BN, EPI>::value)
convgemm_kernel(const __grid_constant__ ConvGemmParams p, int total_tiles) {
  static_assert(!(WS && MC), "weight-stationary mode runs without clusters");
  using L = StageLayout<BN, KB>;
  constexpr int SW = KB * 2;
  constexpr int TBUF = TmemCols<BN>::buf;
  constexpr int TCOLS = TmemCols<BN>::total;
  constexpr int EW = EpiWarps<BN, EPI>::value;
  constexpr int NHALF = EW / 4;
  constexpr uint32_t IDESC = umma_idesc_bf16(MC == 2 ? 2 * kTileM : kTileM, BN);
  constexpr int W_STAGE_BYTES = MC == 2 ? L::w_bytes / 2 : L::w_bytes;   // CTA-pair mode keeps half a slab per CTA

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int a_stages = p.a_stages, w_stages = p.w_stages;
  const uint32_t a_ring = smem_base;
  const uint32_t w_ring = smem_base + a_stages * p.a_box_bytes;
  const uint32_t bar_base = w_ring + w_stages * W_STAGE_BYTES;       // 1024-aligned
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (a_stages + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * a_stages + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * a_stages + w_stages + s); };
  const int n_ring_bars = 2 * (a_stages + w_stages);
  auto tfull_bar = [&](int b) { return bar_base + 8u * (n_ring_bars + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (n_ring_bars + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (n_ring_bars + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t red_off = (bar_base - smem_base) + 8 * (n_ring_bars + 4) + 16;
  float* red = reinterpret_cast<float*>(smem_gen + red_off);  // [2][128]
  const uint32_t epi_bar = smem_base + ((red_off + 1024 + 15) & ~15u);               // EW x 2 residual-stream mbarriers, then
                                                                                     // (RESNORM) EW x 3 single-use ones of the last tile
  // the bookkeeping block (ring barriers, TMEM slot, `red`, epilogue barriers) is 2 KB; bar_base is 1024-aligned
  // (written as an explicitly aligned offset: with the plain sum the compiler lost the 1024-byte alignment of the
  // staging areas and spent ~50 % more integer instructions on every swizzled slot address)
  uint8_t* stage_base = smem_gen + ((red_off + 1024 + 384 + 1023) & ~1023u);         // EW warp-private staging areas
  const int stage_bytes = EpiWarps<BN, EPI>::stage_bytes(p.res_bufs);
  uint8_t* extra_base = stage_base + EW * stage_bytes;                               // CTA-wide epilogue tables

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform (uniform datapath)
  const int lane = threadIdx.x & 31;
  // tile walk: CTA (or, with MC, cluster) `walker` of `n_walkers` takes tiles walker, walker + n_walkers, ...
  const int mc_rank = MC ? static_cast<int>(cluster_ctarank()) : 0;
  const int ws_n = WS ? static_cast<int>(blockIdx.x) % p.n_tiles : 0;
  const int walker = MC ? static_cast<int>(blockIdx.x >> 1)
                        : (WS ? static_cast<int>(blockIdx.x) / p.n_tiles : static_cast<int>(blockIdx.x));
  const int n_walkers = MC ? static_cast<int>(gridDim.x >> 1)
                           : (WS ? static_cast<int>(gridDim.x) / p.n_tiles : static_cast<int>(gridDim.x));
  auto decode = [&](int tile) {
    if constexpr (WS) {
      TileCoord c;
      c.group = 0;
      c.n = ws_n;
      c.b = tile / p.m_tiles[0];
      c.m = tile - c.b * p.m_tiles[0];
      return c;
    } else {
      return MC ? decode_tile_mc(p, tile, mc_rank) : decode_tile(p, tile);
    }
  };

  if (warp == 0) SRB_TRACE_AT(0, 0, 0);   // kernel entry
  if (threadIdx.x == 0) {
    for (int s = 0; s < a_stages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < w_stages; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), MC == 1 ? 2 : 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), MC == 2 ? 2 * EW : EW);
    }
    if constexpr (EPI == EPI_RESNORM || EPI == EPI_GENERIC) {
      for (int i = 0; i < (EPI == EPI_RESNORM ? 5 : 2) * EW; ++i) mbar_init(epi_bar + 8u * i, 1);
    }
    if constexpr (EPI == EPI_RESNORM) {
      tma_prefetch_desc(&p.tmR);
      tma_prefetch_desc(&p.tmO1);
    }
    if constexpr (EPI == EPI_RESNORM || EPI == EPI_QKV_ROPE) tma_prefetch_desc(&p.tmO0);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmA[0]);
  }
  if constexpr (EPI == EPI_QKV_ROPE) rope_table_to_smem(p, extra_base);
  if (warp == 1) {
    if constexpr (MC == 2) {
      tmem_alloc_2sm(tmem_slot, TCOLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, TCOLS);
      tmem_relinquish();
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();   // peer barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  // everything above touched only shared / tensor memory and constant tables; activations come after this point
  if (warp == 0) SRB_TRACE_AT(0, 0, 1);   // prologue done
  // Weights never depend on a predecessor: the producer requests its first ring of slabs (all of them in the
  // weight-stationary form) BEFORE waiting for the previous kernel, so they travel while that kernel drains.
  int w_pre = 0;   // slabs already requested (producer warp only; same order as the main loop below)
  auto request_first_slabs = [&]() {
    if (warp != 0 || walker >= total_tiles) return;
    if constexpr (WS) {
      // slab s = tap * kchunks + chunk lives in slot s for the whole kernel
      for (int s = 0; s < w_stages; ++s) {
        mbar_expect_tx_elect(w_full(s), L::w_bytes_raw);
        tma_load_2d_elect(w_ring + s * L::w_bytes, &p.tmW, w_full(s), s * KB, ws_n * BN);
      }
    } else if constexpr (MC == 0) {
      const TileCoord tc = decode(walker);
      for (int sg = p.group_seg_begin[tc.group]; sg < p.group_seg_begin[tc.group + 1] && w_pre < w_stages; ++sg)
        for (int kc = 0; kc < p.kchunks && w_pre < w_stages; ++kc)
          for (int t = p.seg_tap_begin[sg]; t < p.seg_tap_begin[sg + 1] && w_pre < w_stages; ++t, ++w_pre) {
            mbar_expect_tx_elect(w_full(w_pre), L::w_bytes_raw);
            tma_load_2d_elect(w_ring + w_pre * L::w_bytes, &p.tmW, w_full(w_pre), (t * p.kchunks + kc) * KB, tc.n * BN);
          }
    }
  };
  if (p.w_early) request_first_slabs();
  pdl_wait();
  if (warp == 0) SRB_TRACE_AT(0, 0, 2);   // dependencies satisfied
  if (!p.w_early) request_first_slabs();
  if constexpr (EPI == EPI_QKV_ROPE) {
    // clear the norm-bound buffer of the NEXT q|k projection (nobody reads or fills it while this launch runs)
    if (blockIdx.x == 0 && p.aux1 != nullptr)
      for (int i = threadIdx.x; i < p.batch * 8; i += blockDim.x) static_cast<float*>(p.aux1)[i] = 0.f;
  }

  if (warp == 0) {
    // TMA producer: converged warp, one elected lane issues (coordinates stay in uniform registers)
    {
      int ai = 0, wi = 0;
      uint32_t aph = 0, wph = 0;
      const uint32_t a_tx = static_cast<uint32_t>(p.a_box_rows) * KB * 2;
      for (int tile = walker; tile < total_tiles; tile += n_walkers) {
        const TileCoord tc = decode(tile);
        const int t0 = tc.m * kTileM;
        for (int sg = p.group_seg_begin[tc.group]; sg < p.group_seg_begin[tc.group + 1]; ++sg) {
          const int src = p.seg_src[sg];
          const int row = t0 + p.seg_min_shift[sg];
          const int tb = p.seg_tap_begin[sg], te = p.seg_tap_begin[sg + 1];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(a_empty(ai), aph ^ 1u);
            if constexpr (MC == 2) {
              // both CTAs' boxes complete on the leader's barrier, which the leader arms for the sum
              if (mc_rank == 0) mbar_expect_tx_elect(a_full(ai), 2 * a_tx);
              tma_load_3d_2sm_elect(a_ring + ai * p.a_box_bytes, &p.tmA[src], a_full(ai), kc * KB, row, tc.b);
            } else {
              mbar_expect_tx_elect(a_full(ai), a_tx);
              tma_load_3d_elect(a_ring + ai * p.a_box_bytes, &p.tmA[src], a_full(ai), kc * KB, row, tc.b);
            }
            if (++ai == a_stages) { ai = 0; aph ^= 1u; }
            if constexpr (WS) continue;
            for (int t = tb; t < te; ++t) {
              if (w_pre > 0) {
                // requested before the dependency wait
                --w_pre;
                if (++wi == w_stages) { wi = 0; wph ^= 1u; }
                continue;
              }
              mbar_wait(w_empty(wi), wph ^ 1u);
              if constexpr (MC == 2) {
                if (mc_rank == 0) mbar_expect_tx_elect(w_full(wi), L::w_bytes_raw);
                tma_load_2d_2sm_elect(w_ring + wi * W_STAGE_BYTES, &p.tmWh, w_full(wi), (t * p.kchunks + kc) * KB,
                                      tc.n * BN + mc_rank * (BN / 2));
                if (++wi == w_stages) { wi = 0; wph ^= 1u; }
                continue;
              }
              mbar_expect_tx_elect(w_full(wi), L::w_bytes_raw);
              if constexpr (MC == 1) {
                // my half of the slab (BN/2 rows), written into both CTAs
                tma_load_2d_mc_elect(w_ring + wi * L::w_bytes + mc_rank * (L::w_bytes_raw / 2), &p.tmWh, w_full(wi),
                                     (t * p.kchunks + kc) * KB, tc.n * BN + mc_rank * (BN / 2), (uint16_t)3);
              } else {
                tma_load_2d_elect(w_ring + wi * L::w_bytes, &p.tmW, w_full(wi), (t * p.kchunks + kc) * KB, tc.n * BN);
              }
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
    if (MC != 2 || mc_rank == 0) {
      int ai = 0, wi = 0;
      uint32_t aph = 0, wph = 0;
      int it = 0;
      for (int tile = walker; tile < total_tiles; tile += n_walkers, ++it) {
        const TileCoord tc = decode(tile);
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(buf), bphase ^ 1u);
        tc_fence_after();
        SRB_TRACE_AT(1, it, 0);   // accumulator buffer free
        const uint32_t tacc = tmem_base + buf * TBUF;
        uint32_t first = 1;
        for (int sg = p.group_seg_begin[tc.group]; sg < p.group_seg_begin[tc.group + 1]; ++sg) {
          const int min_shift = p.seg_min_shift[sg];
          const int tb = p.seg_tap_begin[sg], te = p.seg_tap_begin[sg + 1];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(a_full(ai), aph);
            if (kc == 0 && sg == p.group_seg_begin[tc.group]) SRB_TRACE_AT(1, it, 1);   // first activation box landed
            const uint32_t a_addr = a_ring + ai * p.a_box_bytes;
            for (int t = tb; t < te; ++t) {
              if constexpr (WS) {
                wi = t * p.kchunks + kc;
                if (it == 0) mbar_wait(w_full(wi), 0u);
              } else {
                mbar_wait(w_full(wi), wph);
              }
              tc_fence_after();
              const uint64_t adesc = umma_smem_desc<SW>(a_addr + (p.tap_shift[t] - min_shift) * (KB * 2));
              const uint64_t wdesc = umma_smem_desc<SW>(w_ring + wi * W_STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < KB / 16; ++k) {
                if constexpr (MC == 2) umma_bf16_2sm_pred(tacc, adesc + 2 * k, wdesc + 2 * k, IDESC, (first && k == 0) ? 0u : 1u);
                else umma_bf16_pred(1u, tacc, adesc + 2 * k, wdesc + 2 * k, IDESC, (first && k == 0) ? 0u : 1u);
              }
              first = 0;
              if constexpr (WS) continue;
              if constexpr (MC == 2) umma_commit_2sm_pred(w_empty(wi), (uint16_t)3);
              else if constexpr (MC == 1) umma_commit_mc_pred(w_empty(wi), (uint16_t)3);
              else umma_commit_pred(1u, w_empty(wi));
              if (++wi == w_stages) { wi = 0; wph ^= 1u; }
            }
            if constexpr (MC == 2) umma_commit_2sm_pred(a_empty(ai), (uint16_t)3);
            else umma_commit_pred(1u, a_empty(ai));
            if (++ai == a_stages) { ai = 0; aph ^= 1u; }
          }
        }
        if constexpr (MC == 2) umma_commit_2sm_pred(tfull_bar(buf), (uint16_t)3);
        else umma_commit_pred(1u, tfull_bar(buf));
        SRB_TRACE_AT(1, it, 2);   // all MMAs of the tile issued
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;           // column half (0 when EW == 4)
    const int lane_base = quarter * 32;
    uint8_t* my_stage = stage_base + (warp - 2) * stage_bytes;
    // RESNORM: residual stream of this warp (see epi_resnorm); chunk s of the stream is chunk s & 3 of the warp's
    // (s >> 2)-th tile
    ResStream rs;
    rs.buf = smem_u32(my_stage);
    rs.gen = my_stage;
    rs.bar = epi_bar + 16u * (warp - 2);
    rs.nb = p.res_bufs;
    rs.seq = 0;
    rs.tail_buf = w_ring + (warp - 2) * (3 * 4096);
    rs.tail_gen = smem_gen + (rs.tail_buf - smem_base);
    rs.tail_bar = epi_bar + 16u * EW + 24u * (warp - 2);
    rs.tail = 0;
    auto issue_load = [&](int s, int bi) {
      if constexpr (EPI == EPI_RESNORM) {
        const int tile = walker + (s >> 2) * n_walkers;
        if (lane == 0 && tile < total_tiles) {
          const TileCoord t2 = decode(tile);
          mbar_expect_tx(rs.bar + 8u * bi, 4096);
          if (p.l2_keep)
            tma_load_3d_hint(rs.buf + bi * 4096, &p.tmR, rs.bar + 8u * bi, half * (256 / NHALF) + (s & 3) * 32,
                             t2.m * kTileM + lane_base, t2.b, kL2EvictLast);
          else
            tma_load_3d(rs.buf + bi * 4096, &p.tmR, rs.bar + 8u * bi, half * (256 / NHALF) + (s & 3) * 32,
                        t2.m * kTileM + lane_base, t2.b);
        }
      }
    };
    if constexpr (EPI == EPI_RESNORM) {
      for (int i = 0; i < rs.nb; ++i) issue_load(i, i);
    }
    // GENERIC: residual stream; chunk s of the stream is chunk s % NC of the warp's (s / NC)-th tile
    constexpr int G_CW = BN < 32 ? BN : 32;
    constexpr int G_NC = (BN / NHALF) / G_CW;
    GenStream gs;
    gs.buf = smem_u32(my_stage);
    gs.gen = my_stage;
    gs.bar = epi_bar + 16u * (warp - 2);
    gs.n_res = p.n_res;
    gs.nbuf = p.gen_nbuf;
    gs.seq = 0;
    gs.nout = 0;
    auto issue_gen = [&](int s, int bi) {
      if constexpr (EPI == EPI_GENERIC) {
        const int tile = walker + (s / G_NC) * n_walkers;
        if (lane == 0 && tile < total_tiles) {
          const TileCoord t2 = decode(tile);
          const int col = t2.n * BN + half * (BN / NHALF) + (s % G_NC) * G_CW;
          mbar_expect_tx(gs.bar + 8u * bi, gs.n_res * G_CW * 64);
          for (int r = 0; r < gs.n_res; ++r)
            tma_load_3d(gs.buf + (bi * gs.n_res + r) * (G_CW * 64), &p.tmRes[r], gs.bar + 8u * bi, col,
                        t2.m * kTileM + lane_base, t2.b);
        }
      }
    };
    if constexpr (EPI == EPI_GENERIC) {
      if (gs.n_res == 1 && !p.gen_lsu) {
        issue_gen(0, 0);
        if (gs.nbuf == 2) issue_gen(1, 1);
      }
    }
    int it = 0;
    for (int tile = walker; tile < total_tiles; tile += n_walkers, ++it) {
      const TileCoord tc = decode(tile);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const int q = tc.m * kTileM + lane_base + lane;
      EpiWarp ew;
      ew.stage = my_stage;
      ew.lane = lane;
      ew.row0 = tc.m * kTileM + lane_base;
      float4 rope_row = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (EPI == EPI_QKV_ROPE) rope_row = rope_row_fetch(p, tc, ew);
      if (warp == 2) SRB_TRACE_AT(2, it, 0);   // epilogue warp ready for the tile
      mbar_wait(tfull_bar(buf), bphase);
      tc_fence_after();
      if (warp == 2) SRB_TRACE_AT(2, it, 1);   // accumulator complete
      if constexpr (EPI == EPI_RESNORM && MC != 2) {
        // CTA's last tile: its MMAs are complete, so the weight ring is idle from here on (no tile follows, and with
        // multicast every slab a peer writes into this ring has been consumed by the MMAs just completed): request the
        // residual chunks the stream has not asked for yet all at once into this warp's slice of it
        if (p.res_tail && tile + n_walkers >= total_tiles) {
          rs.tail = 1;
          if (lane == 0) {
            for (int c = rs.nb; c < 4; ++c) {
              const uint32_t bar = rs.tail_bar + 8u * (c - rs.nb), dst = rs.tail_buf + (c - rs.nb) * 4096;
              mbar_expect_tx(bar, 4096);
              if (p.l2_keep)
                tma_load_3d_hint(dst, &p.tmR, bar, half * (256 / NHALF) + c * 32, tc.m * kTileM + lane_base, tc.b, kL2EvictLast);
              else
                tma_load_3d(dst, &p.tmR, bar, half * (256 / NHALF) + c * 32, tc.m * kTileM + lane_base, tc.b);
            }
          }
        }
      }
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(lane_base) << 16) + buf * TBUF;
      if constexpr (EPI == EPI_GENERIC) {
        if (p.gen_lsu) epi_generic_lsu<BN, NHALF>(p, tacc, tc, q, half, ew);
        else epi_generic<BN, NHALF>(p, tacc, tc, half, ew, gs, issue_gen);
      }
      else if constexpr (EPI == EPI_GLU) epi_glu<NHALF>(p, tacc, tc, q, half, ew);
      else if constexpr (EPI == EPI_RESNORM) epi_resnorm<NHALF>(p, tacc, tc, q, half, red, lane_base + lane, quarter, ew, rs, issue_load);
      else if constexpr (EPI == EPI_QKV_ROPE) epi_qkv_rope<NHALF>(p, tacc, tc, half, ew, extra_base, rope_row);
      else if constexpr (EPI == EPI_ARGMAX) epi_argmax<NHALF>(p, tacc, tc, q, half);
      else epi_euler(p, tacc, tc, q, ew);
      tc_fence_before();
      __syncwarp();
      if (warp == 2) SRB_TRACE_AT(2, it, 2);   // epilogue of the tile done
      if (lane == 0) {
        if constexpr (MC == 2) mbar_arrive_leader(tempty_bar(buf));   // the leader's MMA warp owns both accumulators
        else mbar_arrive(tempty_bar(buf));
      }
    }
    // stores still reading the staging buffers must finish before the CTA releases its shared memory
    if constexpr (EPI == EPI_RESNORM || EPI == EPI_QKV_ROPE || EPI == EPI_GENERIC) {
      if (lane == 0) bulk_wait_all();
    }
  }

  if (warp == 0) SRB_TRACE_AT(0, 0, 3);   // producer finished
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();   // no CTA may exit while its peer can still multicast into it
  if (warp == 1) {
    if constexpr (MC == 2) tmem_dealloc_2sm(tmem_base, TCOLS);
    else tmem_dealloc(tmem_base, TCOLS);
  }
}

}  // namespace srb
