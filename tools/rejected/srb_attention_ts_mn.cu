// tcgen05 / TMEM attention for the velocity transformer (transformer.py:109-127), second form: 2 heads x d_head 128,
// scale 1/sqrt(128), keys >= len_b masked, no dropout.  Reads q | k | v straight from the fused (B, N, 768) projection
// buffer (no transposed copy of V) and writes o (B, N, 256).
//
// One work item = one (utterance, head, PAIR of 128-query tiles): the two query tiles share every K and V tile, which
// halves the K/V bytes an SM pulls per unit of work -- the first form (one query tile per item) was bound by exactly
// that stream (5 TB/s of L2 -> SM traffic, timeline in profiles/r01_trace_attention.txt).
//
//   S_t = Q_t K^T : tcgen05.mma M=128 (queries) x N=128 (keys) x K=128 (d); A = Q tile, B = K tile, both K-major
//                   (d contiguous) through TMA (128-byte swizzle); one fp32 S buffer per query tile t in {A, B};
//   P_t           : the softmax warps overwrite S_t IN PLACE with bf16 pairs (tcgen05.st; columns [0,32) and [64,96) of
//                   the buffer) -- P never touches shared memory;
//   O_t += P_t V  : A = P_t FROM TENSOR MEMORY (the "TS" operand form), B = V tile exactly as TMA delivers it from the
//                   row-major buffer: keys x d with d contiguous = an MN-major operand (instruction-descriptor bit 16;
//                   LBO = distance of the two 64-wide d atoms, SBO = 1024).  Both forms verified on a B200 by
//                   tools/probes/umma_ts_mn_probe.cu.
// The two query tiles ping-pong: while the softmax warps work on S_A(j) the tensor core runs P_B V(j-1) and S_B(j), so
// neither side waits for the other (the tensor pipe executes in issue order, which also makes the in-place S -> P -> S
// reuse of a buffer safe).  O is never rescaled (see below), so there is no TMEM read-modify-write on the critical path.
//
// Softmax is exact.  Two forms, chosen per utterance and head:
//   * single pass (the common case).  softmax is shift invariant, and in floating point (fp32 sums, bf16 P: both carry
//     the fp32 exponent range) ANY shift works as long as nothing overflows or the row's largest term underflows.
//     The caller supplies max |q|^2 and max |k|^2 per (utterance, head) (the projection kernel records them in its
//     epilogue); by Cauchy-Schwarz every scaled logit lies in [-B, B], B = |q|max |k|max log2(e)/sqrt(128).  When
//     B <= 100 the kernel uses shift 0: P = exp2(s*scale) in [2^-100, 2^100], one sweep over K and V;
//   * two passes otherwise (or when no bounds are given): pass 1 runs Q K^T over all key tiles and keeps only the row
//     maxima, pass 2 recomputes S and forms P = exp2(s*scale - m) with the FINAL maximum.
//
// Shared memory: seven 32 KB tile slots = Q buffers (one or two pairs) + a unified K/V ring (five or three stages): short
// utterances change items often and want the next item's Q resident early (two Q pairs, three ring stages), long ones
// want the deepest K/V ring (one Q pair, five stages).
// Warp roles (448 threads): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocator), warps 2-9 softmax (see there),
// warps 10-13 output (O / l -> bf16).  What bounds the kernel was measured, not guessed: tcgen05.ld delivers 56 B / clk /
// SM whatever the number of warps (tools/probes/tmem_rate_probe.cu), so reading one fp32 S tile (64 KB) takes 0.62 us --
// more than its MMAs or its 16 K exponentials.  Two earlier softmax arrangements (all warps in lockstep on one tile
// without prefetch; one four-warp group per query tile) ran at 1.05 us per tile because load and arithmetic phases were
// serialised (timelines in profiles/r02_trace_attention2_*.txt).
// Persistent: CTA c works on items c, c + gridDim.x, ...
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include "../../include/srb.h"
#include "srb_common.h"
#include "srb_convgemm.cuh"   // pair_barrier, PTX wrappers

namespace srb {

struct Attn2Params {
  CUtensorMap tm;      // 3-D (ld, frames, batch) view of the q | k | v buffer, box (64, 128, 1)
  const int* lengths;
  const float* qk_norm2_max;   // (B, 2 [q|k], 2 [head], 2 [frequency half]) partial bounds of the squared row norms, or null
  __nv_bfloat16* out;  // (B, N, 256)
  int frames;
  int q_col, k_col, v_col;     // first columns of q, k, v in the buffer (0, 256, 512)
  int q_tiles;         // 128-query tiles per utterance
  int q_pairs;         // ceil(q_tiles / 2)
  int n_items;         // batch * 2 heads * q_pairs
  int q_bufs;          // 1 or 2 Q pairs resident
  int kv_stages;       // 7 - 2 * q_bufs
#ifdef SRB_TRACE
  unsigned long long* trace;   // debug build: per CTA (< 4) and role (producer, MMA, softmax warp 2) 256 time stamps in order
#endif
};

#ifdef SRB_TRACE
#define A2_STAMP(role)                                                                        \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x < 4 && lane == 0 && trace_n < 256) {                 \
      unsigned long long t_;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                  \
      p.trace[(blockIdx.x * 3 + (role)) * 256 + trace_n++] = t_;                              \
    }                                                                                         \
  } while (0)
#else
#define A2_STAMP(role) do { } while (0)
#endif

constexpr int kA2Tile = 128;
constexpr int kA2Half = 128 * 128;            // one [128 rows][64 bf16] swizzled half tile = 16 KB
constexpr int kA2TileBytes = 2 * kA2Half;     // 32 KB
constexpr int kA2Slots = 7;

struct A2Smem {
  static constexpr int tiles = 0;                              // 7 x 32 KB: Q pairs first, then the K/V ring
  static constexpr int red = tiles + kA2Slots * kA2TileBytes;  // [2 halves][128] floats: row max / row sum exchange
  static constexpr int lsum = red + 1024;                      // [2 query tiles][128] floats: row sums for the output warps
  static constexpr int items = lsum + 1024;                    // kA2ItemCache work descriptors of this CTA
  static constexpr int bars = items + 256;
  static constexpr int n_bars = 24;
  static constexpr int tmem = bars + 8 * n_bars;
  // no alignment slack: the dynamic shared-memory window of a kernel without static shared memory starts 1 KB aligned
  // (checked at kernel entry: the kernel traps otherwise)
  static constexpr int total = tmem + 16;
  static_assert(total <= 232448, "attention kernel shared memory");
};

constexpr float kA2ScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;   // (1/sqrt(128)) * log2(e)

__device__ __forceinline__ float a2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (bf16 pairs in 32-bit columns, row = lane) comes from tensor memory
__device__ __forceinline__ void umma_bf16_ts_pred(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// MN-major B operand in 128-byte-swizzled shared memory: rows = K index (keys), 64 N elements (d) per 128-byte row,
// 8-row groups 1024 bytes apart (SBO), the second 64-wide N atom `atom_bytes` further (LBO)
__device__ __forceinline__ uint64_t umma_smem_desc_mn128(uint32_t smem_addr, uint32_t atom_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((atom_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

struct A2Item {
  int b, h, q0, len, nkv;
  short two_pass, nq;
};
constexpr int kA2ItemCache = 10;   // 10 x 24 bytes <= 256

__device__ __forceinline__ A2Item a2_item(const Attn2Params& p, int item) {
  A2Item it;
  const int pr = item % p.q_pairs;
  const int bh = item / p.q_pairs;
  it.h = bh & 1;
  it.b = bh >> 1;
  it.q0 = pr * 2 * kA2Tile;
  it.nq = (pr * 2 + 1 < p.q_tiles) ? 2 : 1;
  int len = p.lengths[it.b];
  it.len = len < p.frames ? len : p.frames;
  it.nkv = (it.len + kA2Tile - 1) / kA2Tile;
  it.two_pass = 1;
  if (p.qk_norm2_max != nullptr) {
    // two partial maxima per head (one per rotary frequency half, see epi_qkv_rope); their sum bounds the row norm
    const float* nq = p.qk_norm2_max + ((it.b * 2 + 0) * 2 + it.h) * 2;
    const float* nk = p.qk_norm2_max + ((it.b * 2 + 1) * 2 + it.h) * 2;
    const float q2 = nq[0] + nq[1], k2 = nk[0] + nk[1];
    // 2 % slack covers the bf16 rounding of q and k after the norms were taken; NaN compares false -> two passes
    it.two_pass = (sqrtf(q2 * k2) * kA2ScaleLog2 * 1.02f <= 100.f) ? 0 : 1;
  }
  return it;
}

__global__ void __launch_bounds__(448, 1) attn2_kernel(const __grid_constant__ Attn2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0u) __trap();   // the swizzled tiles below need 1 KB alignment
  uint8_t* smem = smem_raw;
  const int q_bufs = p.q_bufs, kv_stages = p.kv_stages;
  const uint32_t s_q = sbase + A2Smem::tiles;                               // Q pair b: s_q + b * 64 KB (tile t at + t * 32 KB)
  const uint32_t s_kv = s_q + q_bufs * 2 * kA2TileBytes;                    // ring stage s: s_kv + s * 32 KB
  const uint32_t bar0 = sbase + A2Smem::bars;
  auto q_full = [&](int s) { return bar0 + 8u * s; };                       // 0..1
  auto q_empty = [&](int s) { return bar0 + 8u * (2 + s); };                // 2..3
  auto kv_full = [&](int s) { return bar0 + 8u * (4 + s); };                // 4..8
  auto kv_empty = [&](int s) { return bar0 + 8u * (9 + s); };               // 9..13
  auto s_full = [&](int t) { return bar0 + 8u * (14 + t); };                // 14..15
  auto s_done = [&](int t) { return bar0 + 8u * (16 + t); };                // 16..17
  const uint32_t o_full = bar0 + 8u * 18, o_empty = bar0 + 8u * 19;
  auto l_ready = [&](int t) { return bar0 + 8u * (20 + t); };             // 20..21
  const uint32_t tmem_slot = sbase + A2Smem::tmem;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
#ifdef SRB_TRACE
  int trace_n = 0;
#endif
  if (warp <= 2) A2_STAMP(warp);   // kernel entry
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
      mbar_init(s_full(s), 1);
      mbar_init(s_done(s), 8);
      mbar_init(l_ready(s), 8);
    }
    for (int s = 0; s < 5; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    fence_barrier_init();
    tma_prefetch_desc(&p.tm);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // q | k | v, lengths and the norm bounds are produced by the preceding kernels

  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + A2Smem::tmem);
  const uint32_t t_s0 = tmem_base, t_o0 = tmem_base + 256;   // S_A, S_B at +0, +128; O_A, O_B at +256, +384
  constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128);
  constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, 128) | (1u << 16);   // B (= V) is MN-major
  const int first = blockIdx.x, step = gridDim.x;
  // the descriptors of this CTA's first items (lengths, pass count) are resolved once, by one thread each: reading
  // them from global memory at every item start cost each role an L2 round trip per item
  A2Item* item_cache = reinterpret_cast<A2Item*>(smem + A2Smem::items);
  if (threadIdx.x < kA2ItemCache && first + (int)threadIdx.x * step < p.n_items)
    item_cache[threadIdx.x] = a2_item(p, first + threadIdx.x * step);
  __syncthreads();
  auto get_item = [&](int n, int item) { return n < kA2ItemCache ? item_cache[n] : a2_item(p, item); };

  if (warp == 0) {
    // ================= TMA producer =================
    int st = 0;
    uint32_t ph = 0;
    auto load_tile = [&](int col, int row, int b) {
      mbar_wait(kv_empty(st), ph ^ 1u);
      A2_STAMP(0);   // ring slot free, tile requested
      mbar_expect_tx_elect(kv_full(st), kA2TileBytes);
      tma_load_3d_elect(s_kv + st * kA2TileBytes, &p.tm, kv_full(st), col, row, b);
      tma_load_3d_elect(s_kv + st * kA2TileBytes + kA2Half, &p.tm, kv_full(st), col + 64, row, b);
      if (++st == kv_stages) { st = 0; ph ^= 1u; }
    };
    // Q of item nn lives in pair buffer nn % q_bufs
    auto load_q = [&](int nn, const A2Item& iq) {
      const int qb = nn % q_bufs, use = nn / q_bufs;
      mbar_wait(q_empty(qb), (use & 1) ^ 1u);   // every S tile of the item that used this buffer last has been issued and completed
      mbar_expect_tx_elect(q_full(qb), iq.nq * kA2TileBytes);
      for (int t = 0; t < iq.nq; ++t) {
        const uint32_t dst = s_q + (qb * 2 + t) * kA2TileBytes;
        tma_load_3d_elect(dst, &p.tm, q_full(qb), p.q_col + iq.h * 128, iq.q0 + t * kA2Tile, iq.b);
        tma_load_3d_elect(dst + kA2Half, &p.tm, q_full(qb), p.q_col + iq.h * 128 + 64, iq.q0 + t * kA2Tile, iq.b);
      }
    };
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const A2Item it = get_item(n, item);
      const bool has_next = item + step < p.n_items;
      if (n == 0) load_q(0, it);
      // two Q pairs: the next item's Q is requested before this item's K / V tiles, so it lands long before the tensor
      // core gets there; one pair: it can only follow this item's last S tile, i.e. after this item's loads
      if (q_bufs == 2 && has_next) load_q(n + 1, get_item(n + 1, item + step));
      const int kc = p.k_col + it.h * 128, vc = p.v_col + it.h * 128;
      if (it.two_pass)
        for (int j = 0; j < it.nkv; ++j) load_tile(kc, j * kA2Tile, it.b);
      for (int j = 0; j < it.nkv; ++j) {
        load_tile(kc, j * kA2Tile, it.b);
        load_tile(vc, j * kA2Tile, it.b);
      }
      if (q_bufs == 1 && has_next) load_q(n + 1, get_item(n + 1, item + step));
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp, elected lane issues) =================
    int st = 0;
    uint32_t ph = 0;
    int uses[2] = {0, 0};   // S tiles issued into buffer t so far (its s_full / s_done barriers complete once per use)
    auto ring_next = [&]() { if (++st == kv_stages) { st = 0; ph ^= 1u; } };
    auto issue_s = [&](int t, uint32_t qa, int k_slot) {
      const uint32_t kb = s_kv + k_slot * kA2TileBytes;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * kA2Half + (kk & 3) * 32;
        umma_bf16_pred(1u, t_s0 + t * 128, umma_smem_desc<128>(qa + off), umma_smem_desc<128>(kb + off), IDESC_S, kk != 0 ? 1u : 0u);
      }
      umma_commit_pred(1u, s_full(t));
      ++uses[t];
    };
    auto wait_done = [&](int t) {   // the softmax warps have finished with the latest S tile of buffer t (P written / maxima taken)
      mbar_wait(s_done(t), (uses[t] - 1) & 1);
      tc_fence_after();
    };
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const A2Item it = get_item(n, item);
      const int nq = it.nq, nkv = it.nkv;
      const int qb = n % q_bufs;
      A2_STAMP(1);   // item: waiting for Q
      mbar_wait(q_full(qb), (n / q_bufs) & 1);
      tc_fence_after();
      A2_STAMP(1);   // item: Q landed
      const uint32_t qa0 = s_q + qb * 2 * kA2TileBytes;
      if (nkv == 0) {
        umma_commit_pred(1u, q_empty(qb));
        umma_commit_pred(1u, o_full);
        continue;
      }
      if (it.two_pass) {
        // maxima sweep: S only; a buffer takes its next tile once the softmax warps have read the previous one
        for (int j = 0; j < nkv; ++j) {
          mbar_wait(kv_full(st), ph);
          tc_fence_after();
          for (int t = 0; t < nq; ++t) {
            if (uses[t] > 0) wait_done(t);
            issue_s(t, qa0 + t * kA2TileBytes, st);
          }
          umma_commit_pred(1u, kv_empty(st));
          ring_next();
        }
      }
      // main sweep, prologue: S_t(0)
      mbar_wait(kv_full(st), ph);
      tc_fence_after();
      A2_STAMP(1);   // K_0 landed
      for (int t = 0; t < nq; ++t) {
        if (uses[t] > 0 && it.two_pass) wait_done(t);   // the maxima of the sweep's last tile have been taken
        issue_s(t, qa0 + t * kA2TileBytes, st);
      }
      umma_commit_pred(1u, kv_empty(st));
      ring_next();
      if (nkv == 1) umma_commit_pred(1u, q_empty(qb));
      for (int j = 0; j < nkv; ++j) {
        // ring: stage `st` holds V_j, the stage after it K_{j+1}
        const int v_st = st;
        const uint32_t v_ph = ph;
        ring_next();
        const int k_st = st;
        const uint32_t k_ph = ph;
        if (j + 1 < nkv) ring_next();
        for (int t = 0; t < nq; ++t) {
          A2_STAMP(1);   // waiting for P_t(j)
          wait_done(t);                       // P_t(j) is in tensor memory
          A2_STAMP(1);   // P_t(j) ready
          if (t == 0) {
            mbar_wait(kv_full(v_st), v_ph);
            if (j == 0) mbar_wait(o_empty, (n & 1) ^ 1u);   // the output warps have read the previous item's O
            tc_fence_after();
          }
          const uint32_t vb = s_kv + v_st * kA2TileBytes;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16_ts_pred(t_o0 + t * 128, t_s0 + t * 128 + (kk >> 2) * 64 + (kk & 3) * 8,
                              umma_smem_desc_mn128(vb + kk * 2048, kA2Half), IDESC_PV, (j != 0 || kk != 0) ? 1u : 0u);
          if (t == nq - 1) umma_commit_pred(1u, kv_empty(v_st));
          if (j + 1 < nkv) {
            // S_t(j + 1) right behind P_t V(j): the tensor pipe runs in issue order, so the buffer P_t lives in is
            // overwritten only after that product has read it
            if (t == 0) {
              mbar_wait(kv_full(k_st), k_ph);
              tc_fence_after();
            }
            A2_STAMP(1);   // V_j / K_{j+1} landed, P V issued
            issue_s(t, qa0 + t * kA2TileBytes, k_st);
            if (t == nq - 1) {
              umma_commit_pred(1u, kv_empty(k_st));
              if (j + 2 == nkv) umma_commit_pred(1u, q_empty(qb));   // the item's last S tiles are issued: Q may be replaced
            }
          }
        }
      }
      umma_commit_pred(1u, o_full);
    }
    __syncwarp();
  } else if (warp >= 10) {
    // ================= output warps =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float* lsum = reinterpret_cast<const float*>(smem + A2Smem::lsum);
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const A2Item it = get_item(n, item);
      mbar_wait(l_ready(0), n & 1);
      mbar_wait(o_full, n & 1);
      tc_fence_after();
      for (int t = 0; t < it.nq; ++t) {
        const float l = lsum[t * 128 + row];
        const bool ok = l > 0.f && it.nkv > 0;
        const float inv = ok ? 1.f / l : 0.f;
        const int q = it.q0 + t * kA2Tile + row;
        // thread <-> query row: the row's 128 output columns of this head are 256 contiguous bytes
        uint4* out = reinterpret_cast<uint4*>(p.out + ((long long)it.b * p.frames + q) * 256 + it.h * 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(t_o0 + t * 128 + lane_addr + c * 32, v);
          tmem_ld_wait();
          if (c == 3 && t == it.nq - 1) {
            // O and the row sums are in registers: the accumulators and lsum are free for the next item
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
          }
          if (q < p.frames) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              out[c * 4 + i] = ok ? make_uint4(pack_bf16(__uint_as_float(v[8 * i]) * inv, __uint_as_float(v[8 * i + 1]) * inv),
                                               pack_bf16(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv),
                                               pack_bf16(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv),
                                               pack_bf16(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv))
                                  : make_uint4(0, 0, 0, 0);
          }
        }
      }
    }
  } else {
    // ================= softmax warps =================
    // All eight warps work on ONE S tile at a time, in the order the tensor core produces them (A(0), B(0), A(1), ...):
    // thread <-> query row (TMEM lane), the two warps of a lane quarter split the 128 key columns in halves of 64 = two
    // 32-column loads.  The loads form one software pipeline across chunks AND tiles: while a chunk is processed the
    // next one (possibly the first of the next tile, if its S is complete) is already in flight, so the tcgen05.ld
    // port -- 56 B / clk / SM measured (tools/probes/tmem_rate_probe.cu), i.e. 0.62 us per 64 KB S tile, the bound of this
    // kernel -- never waits for the arithmetic.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                     // key-column half of S
    const int row = quarter * 32 + lane;                  // query row inside the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = kA2ScaleLog2;
    float* red = reinterpret_cast<float*>(smem + A2Smem::red);
    float* lsum = reinterpret_cast<float*>(smem + A2Smem::lsum);
    int usesA = 0, usesB = 0;   // S tiles consumed from buffers A / B
    uint32_t va[32], vb[32];
    bool have_first = false;    // chunk 0 of the next tile of the stream is already in va
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const A2Item it = get_item(n, item);
      const int len = it.len, nkv = it.nkv, nq = it.nq;
      const int sweeps = (it.two_pass && nkv > 0) ? 2 : 1;
      float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;
      float m2A = 0.f, m2B = 0.f;   // single pass: shift 0 (see the header comment)
      for (int sw = 0; sw < sweeps; ++sw) {
        const bool maxima = sweeps == 2 && sw == 0;
        const int ntiles = nkv * nq;
        for (int k = 0; k < ntiles; ++k) {
          const int t = nq == 2 ? (k & 1) : 0;
          const int jj = nq == 2 ? (k >> 1) : k;
          const uint32_t ts = t_s0 + t * 128 + lane_addr + half * 64;
          if (!have_first) {
            if (warp == 2) A2_STAMP(2);   // waiting for S_t(j)
            mbar_wait(s_full(t), (t ? usesB : usesA) & 1);
            tc_fence_after();
            if (warp == 2) A2_STAMP(2);   // S_t(j) ready
            tmem_ld32(ts, va);
          }
          if (t) ++usesB; else ++usesA;
          have_first = false;
          tmem_ld32(ts + 32, vb);          // second chunk in flight while the first is processed
          tmem_ld_wait_dep(va);            // (waits for both; the first has long arrived in the steady state)
          const int nvalid = len - jj * kA2Tile - half * 64;   // valid keys among this thread's 64
          const float m2 = t ? m2B : m2A;
          float lt = 0.f, mt = -INFINITY;
          uint32_t o[16];
          auto chunk = [&](uint32_t (&cur)[32], int c) {
            if (maxima) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i < nvalid) mt = fmaxf(mt, __uint_as_float(cur[i]));
              return;
            }
            if (nvalid >= 64) {
              float l0 = 0.f, l1 = 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float p0 = a2_ex2(fmaf(__uint_as_float(cur[2 * i]), sl2, -m2));
                const float p1 = a2_ex2(fmaf(__uint_as_float(cur[2 * i + 1]), sl2, -m2));
                l0 += p0;
                l1 += p1;
                o[i] = pack_bf16(p0, p1);
              }
              lt += l0 + l1;
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int key = c * 32 + 2 * i;
                const float p0 = key < nvalid ? a2_ex2(fmaf(__uint_as_float(cur[2 * i]), sl2, -m2)) : 0.f;
                const float p1 = key + 1 < nvalid ? a2_ex2(fmaf(__uint_as_float(cur[2 * i + 1]), sl2, -m2)) : 0.f;
                lt += p0 + p1;
                o[i] = pack_bf16(p0, p1);
              }
            }
            // P in place: this warp's 64 keys become 32 packed columns at the start of its own half of the S buffer
            tmem_st16(ts + c * 16, o);
          };
          chunk(va, 0);
          tmem_ld_wait_dep(vb);
          // the stream's next tile (same item and sweep): if its S is already complete, its first chunk goes in flight now
          if (k + 1 < ntiles) {
            const int t2 = nq == 2 ? ((k + 1) & 1) : 0;
            if (mbar_try_wait(s_full(t2), (t2 ? usesB : usesA) & 1)) {
              tc_fence_after();
              tmem_ld32(t_s0 + t2 * 128 + lane_addr + half * 64, va);
              have_first = true;
            }
          }
          chunk(vb, 1);
          if (t) { lB += lt; mB = fmaxf(mB, mt); } else { lA += lt; mA = fmaxf(mA, mt); }
          if (!maxima) tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_done(t));
          if (warp == 2) A2_STAMP(2);   // P_t(j) published
        }
        if (maxima) {
          // row maxima over both halves
          for (int t = 0; t < nq; ++t) {
            float m = t ? mB : mA;
            red[half * 128 + row] = m;
            pair_barrier(quarter);
            m = fmaxf(m, red[(half ^ 1) * 128 + row]);
            pair_barrier(quarter);                              // both have read before `red` is reused
            if (t) m2B = m * sl2; else m2A = m * sl2;           // finite: every utterance has at least one valid key
          }
        }
      }
      // row sums over both halves, handed to the output warps; lsum was last read by the output warps of the previous item
      mbar_wait(o_empty, (n & 1) ^ 1u);
      for (int t = 0; t < nq; ++t) {
        const float l = t ? lB : lA;
        red[half * 128 + row] = l;
        pair_barrier(quarter);
        const float tot = l + red[(half ^ 1) * 128 + row];
        if (half == 0) lsum[t * 128 + row] = tot;
        pair_barrier(quarter);        // `red` is free again
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(l_ready(0));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static PFN_cuTensorMapEncodeTiled_v12000 attn2_get_encode() {
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

extern "C" int srb_cfm_attention_qkv(const void* qkv_bf16, int32_t ld, const int32_t* lengths, const float* qk_norm2_max,
                                     void* o_bf16, int32_t batch, int32_t frames, void* stream) {
  if (batch <= 0 || frames <= 0) return 0;
  auto enc = attn2_get_encode();
  SRB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  SRB_REQUIRE(ld >= 768 && ld % 8 == 0, "srb_cfm_attention_qkv: the q | k | v buffer needs a row pitch >= 768, multiple of 8");
  Attn2Params p;
  {
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)frames, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)frames * ld * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(q|k|v) failed: %d", (int)r);
  }
  p.lengths = lengths;
#ifdef SRB_TRACE
  p.trace = debug_trace_buffer();
#endif
  p.qk_norm2_max = qk_norm2_max;
  p.out = static_cast<__nv_bfloat16*>(o_bf16);
  p.frames = frames;
  p.q_col = 0;
  p.k_col = 256;
  p.v_col = 512;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    SRB_CUDA(cudaFuncSetAttribute(attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2Smem::total));
    configured[dev & 63] = true;
  }
  p.q_tiles = (frames + kA2Tile - 1) / kA2Tile;
  p.q_pairs = (p.q_tiles + 1) / 2;
  p.n_items = batch * 2 * p.q_pairs;
  // short utterances change items often: keep the next item's Q resident (two Q pairs + three K/V stages); long ones
  // get the deeper K/V ring.  SRB_ATTN_QBUFS = 1 | 2 overrides (A/B runs).
  static const int force = [] { const char* e = getenv("SRB_ATTN_QBUFS"); return e ? atoi(e) : 0; }();
  p.q_bufs = force == 1 || force == 2 ? force : (frames <= 1536 ? 2 : 1);
  p.kv_stages = kA2Slots - 2 * p.q_bufs;
  int grid = num_sms();
  if (grid > p.n_items) grid = p.n_items;
  SRB_CUDA(launch_pdl(attn2_kernel, dim3(grid), dim3(448), A2Smem::total, (cudaStream_t)stream, p));
  return after_launch("attn2_kernel");
}
