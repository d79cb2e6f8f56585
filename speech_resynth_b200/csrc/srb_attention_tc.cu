// tcgen05 / TMEM attention for the velocity transformer (transformer.py:115-127): 2 heads x d_head 128, scale
// 1/sqrt(128), keys >= len_b masked, no dropout.  One CTA = one (utterance, head, 128-query tile).
//
//   S = Q K^T   : tcgen05.mma M=128 (queries) x N=128 (keys) x K=128 (d), A = Q tile, B = K tile, both K-major
//                 (d contiguous) straight from the (B, N, ld) q|k buffer through TMA (128-byte swizzle);
//   O += P V    : M=128 x N=128 (d) x K=128 (keys), A = P (bf16, written to swizzled smem by the softmax warps),
//                 B = V^T tile.  V is produced already TRANSPOSED ([d][utterance*frames], keys contiguous) by a
//                 swapped-operand GEMM (srb_cfm_v_transposed), so both operands of P V are K-major as well.
//
// Softmax is exact and never rescales O (no TMEM read-modify-write on the critical path).  Two forms:
//   * single pass (the common case).  softmax is shift invariant, and in floating point (fp32 sums, bf16 P: both carry
//     the fp32 exponent range) ANY shift works as long as nothing overflows or the row's largest term underflows.
//     The caller supplies max |q|^2 and max |k|^2 per (utterance, head) (srb_cfm_qk_rope records them in its
//     epilogue); by Cauchy-Schwarz every scaled logit lies in [-B, B], B = |q|max |k|max log2(e)/sqrt(128).  When
//     B <= 100 the kernel uses shift 0: P = exp2(s*scale) in [2^-100, 2^100], one sweep over K and V;
//   * two passes otherwise (or when no bounds are given): pass 1 runs Q K^T over all key tiles and keeps only the row
//     maxima, pass 2 recomputes S and forms P = exp2(s*scale - m) with the FINAL maximum.
// S is double-buffered in TMEM so the tensor core computes S(j+1) while the softmax warps work on S(j).
//
// Warp roles (448 threads): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocator), warps 10-13 output (O / l -> bf16
// -> global, one warp per TMEM lane quarter: the softmax warps hand over the row sums and start the next item at once;
// waiting for the last P V and writing the output was 1.7 of the 6.35 us an item took them), warps 2-9 softmax:
// thread <-> query row (TMEM lane), and the two warps of a lane quarter split the 128 key columns of an S tile (and
// the 128 output columns of O) in halves.  Four softmax warps (one per SM sub-partition) left the kernel waiting on
// them -- ncu: 75 % of the samples on the S / P mbarriers -- so each row's work is shared by two threads; the row
// maximum after pass 1 and the row sum after pass 2 are combined through shared memory, once each.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include "../../include/srb.h"
#include "srb_common.h"
#include "srb_convgemm.cuh"   // EpiWarp / scatter_store coalescing helpers, PTX wrappers

namespace srb {

struct AttnParams {
  CUtensorMap tm_qk;   // 3-D (ld, frames, batch), box (64, 128, 1)
  CUtensorMap tm_vt;   // 2-D (m_pad, 256), box (64, 128)
  const int* lengths;
  const float* qk_norm2_max;   // (B, 2 [q|k], 2 [head], 2 [frequency half]) partial bounds of the squared row norms, or null
  __nv_bfloat16* out;  // (B, N, 256)
  int frames;
  int k_col;           // first column of k in the q|k buffer (256)
#ifdef SRB_TRACE
  unsigned long long* trace;   // debug build: per CTA (< 4) and role (producer, MMA, softmax warp 2) 256 time stamps in order
#endif
  int q_tiles;         // 128-query tiles per utterance
  int n_items;         // batch * 2 heads * q_tiles
};

constexpr int kAttnTile = 128;
constexpr int kHalfBytes = 128 * 128;          // one [128 rows][64 bf16] swizzled half tile = 16 KB
constexpr int kTileBytes = 2 * kHalfBytes;     // 32 KB

struct AttnSmem {
  static constexpr int q = 0;
  static constexpr int k = q + 2 * kTileBytes;         // (Q: 2 buffers, the next item's Q lands while this item runs) 2 stages
  static constexpr int v = k + 2 * kTileBytes;         // 2 stages
  static constexpr int p = v + 2 * kTileBytes;         // 1 buffer (the wait for P V of the previous tile measured 0.1 us)
  static constexpr int red = p + kTileBytes;           // [2][128] floats: row max / row sum exchange between halves
  static constexpr int lsum = red + 1024;              // [2 O buffers][128] floats: row sums handed to the output warps
  static constexpr int items = lsum + 1024;            // kItemCache work descriptors of this CTA
  static constexpr int bars = items + 512;
  static constexpr int n_bars = 26;
  static constexpr int tmem = bars + 8 * n_bars;
  // no alignment slack: the dynamic shared-memory window of a kernel without static shared memory starts 1 KB aligned
  // (checked at kernel entry: the kernel traps otherwise)
  static constexpr int total = tmem + 16;
  static_assert(total <= 232448, "attention kernel shared memory");
};

constexpr float kAttnScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;   // (1/sqrt(128)) * log2(e)

#ifdef SRB_TRACE
#define ATTN_STAMP(role)                                                                      \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x < 4 && lane == 0 && trace_n < 256) {                 \
      unsigned long long t_;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                  \
      p.trace[(blockIdx.x * 3 + (role)) * 256 + trace_n++] = t_;                              \
    }                                                                                         \
  } while (0)
#else
#define ATTN_STAMP(role) do { } while (0)
#endif

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Persistent: CTA c works on items c, c + gridDim.x, ... where an item is one (utterance, head, 128-query tile); the
// K / V / S / P rings and the double-buffered O accumulator run across item boundaries, so while the softmax warps
// finish item i the producer is already fetching Q, K, V of item i + 1 and the tensor core computes its first S tiles.
// (One CTA per item paid ~3.5 us of launch, TMEM allocation and first-load latency per item -- a third of its time.)
struct AttnItem {
  int b, h, q0, len, nkv;
  int two_pass;
};
constexpr int kItemCache = 16;   // 16 x 24 bytes <= 512

__device__ __forceinline__ AttnItem attn_item(const AttnParams& p, int item) {
  AttnItem it;
  const int qt = item % p.q_tiles;
  const int bh = item / p.q_tiles;
  it.h = bh & 1;
  it.b = bh >> 1;
  it.q0 = qt * kAttnTile;
  int len = p.lengths[it.b];
  it.len = len < p.frames ? len : p.frames;
  it.nkv = (it.len + kAttnTile - 1) / kAttnTile;
  it.two_pass = 1;
  if (p.qk_norm2_max != nullptr) {
    // two partial maxima per head (one per rotary frequency half, see epi_qkv_rope); their sum bounds the row norm
    const float* nq = p.qk_norm2_max + ((it.b * 2 + 0) * 2 + it.h) * 2;
    const float* nk = p.qk_norm2_max + ((it.b * 2 + 1) * 2 + it.h) * 2;
    const float q2 = nq[0] + nq[1], k2 = nk[0] + nk[1];
    // 2 % slack covers the bf16 rounding of q and k after the norms were taken; NaN compares false -> two passes
    it.two_pass = (sqrtf(q2 * k2) * kAttnScaleLog2 * 1.02f <= 100.f) ? 0 : 1;
  }
  return it;
}

__global__ void __launch_bounds__(448, 1) attn_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0u) __trap();   // the swizzled tiles below need 1 KB alignment
  uint8_t* smem = smem_raw;
  const uint32_t s_q = sbase + AttnSmem::q, s_k = sbase + AttnSmem::k, s_v = sbase + AttnSmem::v, s_p = sbase + AttnSmem::p;
  const uint32_t bar0 = sbase + AttnSmem::bars;
  auto q_full = [&](int s) { return bar0 + 8u * (24 + s); };
  auto q_empty = [&](int s) { return bar0 + 8u * s; };
  auto k_full = [&](int s) { return bar0 + 8u * (2 + s); };
  auto k_empty = [&](int s) { return bar0 + 8u * (4 + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (6 + s); };
  auto v_empty = [&](int s) { return bar0 + 8u * (8 + s); };
  auto s_full = [&](int s) { return bar0 + 8u * (10 + s); };
  auto s_empty = [&](int s) { return bar0 + 8u * (12 + s); };
  auto p_full = [&](int s) { return bar0 + 8u * (14 + s); };
  auto p_empty = [&](int s) { return bar0 + 8u * (16 + s); };
  auto o_full = [&](int s) { return bar0 + 8u * (18 + s); };
  auto o_empty = [&](int s) { return bar0 + 8u * (20 + s); };
  auto l_ready = [&](int s) { return bar0 + 8u * (22 + s); };
  const uint32_t tmem_slot = sbase + AttnSmem::tmem;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
#ifdef SRB_TRACE
  int trace_n = 0;
#endif
  if (warp <= 2) ATTN_STAMP(warp);   // kernel entry
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
      mbar_init(s_full(s), 1);
      mbar_init(s_empty(s), 8);
      mbar_init(o_full(s), 1);
      mbar_init(o_empty(s), 4);
      mbar_init(l_ready(s), 8);
      mbar_init(p_full(s), 8);
      mbar_init(p_empty(s), 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.tm_qk);
    tma_prefetch_desc(&p.tm_vt);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // q|k, v^T, lengths and the norm bounds are produced by the preceding kernels

  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + AttnSmem::tmem);
  const uint32_t t_s0 = tmem_base, t_o0 = tmem_base + 256;   // S buffers at +0, +128; O buffers at +256, +384
  constexpr uint32_t IDESC = umma_idesc_bf16(128, 128);
  const int first = blockIdx.x, step = gridDim.x;
  // the descriptors of this CTA's first items (lengths, pass count) are resolved once, by one thread each: reading
  // them from global memory at every item start cost each role an L2 round trip per item
  AttnItem* item_cache = reinterpret_cast<AttnItem*>(smem + AttnSmem::items);
  if (threadIdx.x < kItemCache && first + (int)threadIdx.x * step < p.n_items)
    item_cache[threadIdx.x] = attn_item(p, first + threadIdx.x * step);
  __syncthreads();
  auto get_item = [&](int n, int item) { return n < kItemCache ? item_cache[n] : attn_item(p, item); };

  if (warp == 0) {
    // ================= TMA producer =================
    int kst = 0, vst = 0;
    uint32_t kph = 0, vph = 0;
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const AttnItem it = get_item(n, item);
      auto load_k = [&](int j) {
        mbar_wait(k_empty(kst), kph ^ 1u);
        mbar_expect_tx_elect(k_full(kst), kTileBytes);
        tma_load_3d_elect(s_k + kst * kTileBytes, &p.tm_qk, k_full(kst), p.k_col + it.h * 128, j * kAttnTile, it.b);
        tma_load_3d_elect(s_k + kst * kTileBytes + kHalfBytes, &p.tm_qk, k_full(kst), p.k_col + it.h * 128 + 64, j * kAttnTile, it.b);
        if (++kst == 2) { kst = 0; kph ^= 1u; }
      };
      auto load_v = [&](int j) {
        mbar_wait(v_empty(vst), vph ^ 1u);
        mbar_expect_tx_elect(v_full(vst), kTileBytes);
        const int col = it.b * p.frames + j * kAttnTile;
        tma_load_2d_elect(s_v + vst * kTileBytes, &p.tm_vt, v_full(vst), col, it.h * 128);
        tma_load_2d_elect(s_v + vst * kTileBytes + kHalfBytes, &p.tm_vt, v_full(vst), col + 64, it.h * 128);
        if (++vst == 2) { vst = 0; vph ^= 1u; }
      };
      // Q of item n lives in buffer n & 1; the NEXT item's Q is requested before this item's K / V tiles, so it lands
      // long before the tensor core gets there (with one buffer the softmax warps idled ~2 us at every item boundary)
      auto load_q = [&](int nn, const AttnItem& iq) {
        const int qb = nn & 1;
        mbar_wait(q_empty(qb), ((nn >> 1) & 1) ^ 1u);   // every S tile of item nn - 2 has consumed its Q
        mbar_expect_tx_elect(q_full(qb), kTileBytes);
        tma_load_3d_elect(s_q + qb * kTileBytes, &p.tm_qk, q_full(qb), iq.h * 128, iq.q0, iq.b);
        tma_load_3d_elect(s_q + qb * kTileBytes + kHalfBytes, &p.tm_qk, q_full(qb), iq.h * 128 + 64, iq.q0, iq.b);
      };
      if (n == 0) {
        ATTN_STAMP(0);
        load_q(0, it);
      }
      if (item + step < p.n_items) load_q(n + 1, get_item(n + 1, item + step));
      int v_ahead = 0;
      if (it.two_pass) {
        // the V ring is idle during the maxima pass: fill it first
        for (; v_ahead < 2 && v_ahead < it.nkv; ++v_ahead) load_v(v_ahead);
        for (int j = 0; j < it.nkv; ++j) load_k(j);
      }
      for (int j = 0; j < it.nkv; ++j) {
        load_k(j);
        if (j >= v_ahead) load_v(j);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp, elected lane issues) =================
    // S tiles form ONE stream across items (per item: the maxima sweep if it needs two passes, then the main sweep);
    // the stream runs one tile ahead of the P V products, also across item boundaries, so the first S of item n + 1 is
    // being computed while the softmax warps finish item n.
    int kst = 0, vst = 0;
    uint32_t kph = 0, vph = 0;
    int t = 0;    // S tiles issued so far (ring buffer = t & 1)
    int pv = 0;   // P V products issued so far
    // S stream cursor
    int s_n = 0, s_item = first, s_j = 0, s_left = 0;   // item ordinal / id, next tile inside the item, tiles left in it
    AttnItem s_it;
    bool s_open = false;
    auto next_s = [&]() {
      // advance to an item with tiles left
      while (!s_open || s_left == 0) {
        if (s_open) {
          // (the previous item's last S tile has been issued: its Q buffer may be refilled)
          s_item += step;
          ++s_n;
          s_open = false;
        }
        if (s_item >= p.n_items) return;
        s_it = get_item(s_n, s_item);
        s_left = s_it.two_pass ? 2 * s_it.nkv : s_it.nkv;
        s_j = 0;
        mbar_wait(q_full(s_n & 1), (s_n >> 1) & 1);
        tc_fence_after();
        ATTN_STAMP(1);   // item: Q landed
        s_open = true;
        if (s_left == 0) umma_commit_pred(1u, q_empty(s_n & 1));
      }
      const int sb = t & 1;
      mbar_wait(k_full(kst), kph);
      ATTN_STAMP(1);   // S: K landed
      mbar_wait(s_empty(sb), ((t >> 1) & 1) ^ 1u);
      tc_fence_after();
      ATTN_STAMP(1);   // S: buffer free, issuing
      const uint32_t qa = s_q + (s_n & 1) * kTileBytes;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * kHalfBytes + (kk & 3) * 32;
        umma_bf16_pred(1u, t_s0 + sb * 128, umma_smem_desc<128>(qa + off), umma_smem_desc<128>(s_k + kst * kTileBytes + off),
                       IDESC, kk != 0 ? 1u : 0u);
      }
      umma_commit_pred(1u, k_empty(kst));
      umma_commit_pred(1u, s_full(sb));
      if (--s_left == 0) umma_commit_pred(1u, q_empty(s_n & 1));   // Q may be replaced by the item after next
      if (++kst == 2) { kst = 0; kph ^= 1u; }
      ++s_j;
      ++t;
    };
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const AttnItem it = get_item(n, item);
      const int ob = n & 1;
      const uint32_t t_o = t_o0 + ob * 128;
      // S tiles of this item that must exist before its first P V: the maxima sweep and main tile 0
      const int need = (it.two_pass ? it.nkv : 0) + (it.nkv > 0 ? 1 : 0);
      while (s_n < n || (s_n == n && (!s_open || s_j < need))) {
        if (s_n == n && s_open && s_left == 0) break;
        next_s();
        if (s_item >= p.n_items && !s_open) break;
      }
      // the output warps have read the O buffer this item accumulates into (two items ago)
      mbar_wait(o_empty(ob), ((n >> 1) & 1) ^ 1u);
      tc_fence_after();
      for (int j = 0; j < it.nkv; ++j) {
        next_s();                                         // one S tile ahead (possibly the next item's first)
        mbar_wait(p_full(0), pv & 1);
        ATTN_STAMP(1);   // PV: P ready
        mbar_wait(v_full(vst), vph);
        tc_fence_after();
        ATTN_STAMP(1);   // PV: V landed, issuing
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * kHalfBytes + (kk & 3) * 32;
          umma_bf16_pred(1u, t_o, umma_smem_desc<128>(s_p + off), umma_smem_desc<128>(s_v + vst * kTileBytes + off), IDESC,
                         (j != 0 || kk != 0) ? 1u : 0u);
        }
        umma_commit_pred(1u, v_empty(vst));
        umma_commit_pred(1u, p_empty(0));
        if (++vst == 2) { vst = 0; vph ^= 1u; }
        ++pv;
      }
      umma_commit_pred(1u, o_full(ob));
    }
    __syncwarp();
  } else if (warp >= 10) {
    // ================= output warps =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float* lsum = reinterpret_cast<const float*>(smem + AttnSmem::lsum);
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const AttnItem it = get_item(n, item);
      const int ob = n & 1;
      mbar_wait(l_ready(ob), (n >> 1) & 1);
      mbar_wait(o_full(ob), (n >> 1) & 1);
      tc_fence_after();
      const float l = lsum[ob * 128 + row];
      const float inv = (l > 0.f && it.nkv > 0) ? 1.f / l : 0.f;
      const int q = it.q0 + row;
      // thread <-> query row: the row's 128 output columns of this head are 256 contiguous bytes
      uint4* out = reinterpret_cast<uint4*>(p.out + ((long long)it.b * p.frames + q) * 256 + it.h * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(t_o0 + ob * 128 + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (c == 3) {
          // O and the row sum are in registers: the accumulator and lsum[ob] are free for the item after next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty(ob));
        }
        if (q < p.frames) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            out[c * 4 + i] = make_uint4(pack_bf16(__uint_as_float(v[8 * i]) * inv, __uint_as_float(v[8 * i + 1]) * inv),
                                        pack_bf16(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv),
                                        pack_bf16(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv),
                                        pack_bf16(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv));
        }
      }
    }
  } else {
    // ================= softmax warps =================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                     // key-column half of S / output-column half of O
    const int row = quarter * 32 + lane;                  // query row inside the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = kAttnScaleLog2;
    float* red = reinterpret_cast<float*>(smem + AttnSmem::red);
    int t = 0;    // S tiles consumed
    int pv = 0;   // P tiles produced
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const AttnItem it = get_item(n, item);
      const int len = it.len, nkv = it.nkv;
      const int ob = n & 1;
      float m = -INFINITY;
      for (int j = 0; j < (it.two_pass ? nkv : 0); ++j, ++t) {
        const int sb = t & 1;
        mbar_wait(s_full(sb), (t >> 1) & 1);
        tc_fence_after();
        const int key0 = j * kAttnTile + half * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(t_s0 + sb * 128 + lane_addr + half * 64 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (key0 + c * 32 + i < len) m = fmaxf(m, __uint_as_float(v[i]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(sb));
      }
      float m2 = 0.f;               // single pass: shift 0 (see the header comment)
      if (it.two_pass) {
        // row maximum over both halves
        red[half * 128 + row] = m;
        pair_barrier(quarter);
        m = fmaxf(m, red[(half ^ 1) * 128 + row]);
        pair_barrier(quarter);                              // both have read before `red` is reused for the sums
        m2 = m * sl2;               // finite: every utterance has at least one valid key
      }
      float l = 0.f;
      for (int j = 0; j < nkv; ++j, ++t, ++pv) {
        const int sb = t & 1;
        if (warp == 2) ATTN_STAMP(2);   // tile: waiting for S
        mbar_wait(s_full(sb), (t >> 1) & 1);
        tc_fence_after();
        if (warp == 2) ATTN_STAMP(2);   // tile: S ready
        const int key0 = j * kAttnTile + half * 64;
        // this warp's 64 keys are one [128 rows][64 keys] half tile of P: 128-byte rows, 16-byte pieces XOR-swizzled
        uint8_t* prow = smem + AttnSmem::p + half * kHalfBytes + row * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(t_s0 + sb * 128 + lane_addr + half * 64 + c * 32, v);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int key = key0 + c * 32 + 2 * i;
            const float p0 = key < len ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), sl2, -m2)) : 0.f;
            const float p1 = key + 1 < len ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -m2)) : 0.f;
            l += p0 + p1;
            o[i] = pack_bf16(p0, p1);
          }
          if (c == 0) {
            if (warp == 2) ATTN_STAMP(2);   // tile: first 32 columns done
            mbar_wait(p_empty(0), (pv & 1) ^ 1u);   // P V of the previous tile has finished reading the P tile
            if (warp == 2) ATTN_STAMP(2);   // tile: P buffer free
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int piece = c * 4 + i;
            *reinterpret_cast<uint4*>(prow + ((piece ^ (row & 7)) << 4)) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(s_empty(sb));
          mbar_arrive(p_full(0));
        }
        if (warp == 2) ATTN_STAMP(2);   // tile: P published
      }
      // row sum over both halves, handed to the output warps
      red[half * 128 + row] = l;
      pair_barrier(quarter);
      l += red[(half ^ 1) * 128 + row];
      // lsum[ob] was last read by the output warps two items ago (they release it together with the O buffer)
      mbar_wait(o_empty(ob), ((n >> 1) & 1) ^ 1u);
      if (half == 0) reinterpret_cast<float*>(smem + AttnSmem::lsum)[ob * 128 + row] = l;
      __syncwarp();
      if (lane == 0) mbar_arrive(l_ready(ob));
      pair_barrier(quarter);        // `red` is free for the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static PFN_cuTensorMapEncodeTiled_v12000 attn_get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

}  // namespace srb

using namespace srb;

extern "C" int srb_cfm_attention_tc(const void* qk_bf16, int32_t ld, const void* vt_bf16, int64_t m_pad,
                                    const int32_t* lengths, const float* qk_norm2_max, void* o_bf16, int32_t batch,
                                    int32_t frames, void* stream) {
  SRB_REQUIRE(kSplit == 1, "srb_cfm_attention_tc: not available in the tight-precision build");
  if (batch <= 0 || frames <= 0) return 0;
  auto enc = attn_get_encode();
  SRB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  SRB_REQUIRE(ld >= 512 && ld % 8 == 0 && m_pad % 8 == 0 && m_pad >= (int64_t)batch * frames, "srb_cfm_attention_tc: bad strides");
  SRB_REQUIRE(frames % 8 == 0, "srb_cfm_attention_tc: frames must be a multiple of 8 (TMA box origins in v^T must be 16-byte aligned)");
  AttnParams p;
  {
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)frames, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)frames * ld * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p.tm_qk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qk_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(q|k) failed: %d", (int)r);
  }
  {
    // the map ends at the last utterance's last frame: columns beyond (never written by a fused projection) read as zero
    cuuint64_t dims[2] = {(cuuint64_t)batch * (cuuint64_t)frames, 256};
    cuuint64_t strides[1] = {(cuuint64_t)m_pad * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tm_vt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(vt_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(v^T) failed: %d", (int)r);
  }
  p.lengths = lengths;
#ifdef SRB_TRACE
  p.trace = debug_trace_buffer();
#endif
  p.qk_norm2_max = qk_norm2_max;
  p.out = static_cast<__nv_bfloat16*>(o_bf16);
  p.frames = frames;
  p.k_col = 256;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    SRB_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::total));
    configured[dev & 63] = true;
  }
  p.q_tiles = (frames + kAttnTile - 1) / kAttnTile;
  p.n_items = batch * 2 * p.q_tiles;
  int grid = num_sms();
  if (grid > p.n_items) grid = p.n_items;
  SRB_CUDA(launch_pdl(attn_tc_kernel, dim3(grid), dim3(448), AttnSmem::total, (cudaStream_t)stream, p));
  return after_launch("attn_tc_kernel");
}
